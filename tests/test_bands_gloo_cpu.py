"""N > 1 host logic on CPU: world_size 2 and 3 over gloo. The strip plan, the exchange steps and the
gather of raytracingdiffusioncurves_b200.distributed run for real — the collective form (render_frame) and the
peer-memory form (render_frame_peer, with shared host memory standing in for NVLink peer memory); what renders
and blurs is the CPU oracle (injected — the product module itself never touches oracle/). The assembled frame
must equal the single-process frame bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle as po
from raytracingdiffusioncurves_b200 import distributed as rd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
XML = os.path.join(ROOT, "tests", "golden", "xmls")


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def oracle_callbacks(scene, width, height, rpp, zoom):
    oracle = po.Oracle("port")

    def render_strips(image, sigma, stride, offset):
        p = po.make_params(width, height, rpp, zoom_factor=zoom, strip_stride=stride if stride > 1 else 0,
                           strip_offset=offset if stride > 1 else 0)
        rows = image.shape[0]
        p_rows = po.make_params(width, height, rpp)  # only to size the oracle's output arrays
        del p_rows
        img, sig = render_packed(oracle, scene, p, rows)
        image.copy_(torch.from_numpy(img))
        sigma.copy_(torch.from_numpy(sig))

    def blur_rows(dest, source, sigma, scratch, rows, row_begin, row_end, halo):
        out = oracle.blur(source[:rows].numpy(), sigma[:rows].numpy(), threads=2)
        dest[row_begin:row_end].copy_(torch.from_numpy(out[row_begin:row_end]))

    return render_strips, blur_rows


def render_packed(oracle, scene, p, rows):
    """oracle_render into buffers of `rows` packed rows (the strips of one rank)."""
    import ctypes as C

    a, keep = po.arrays_from_dict(scene)
    img = np.zeros((rows, p.image_width, 4), np.float32)
    sig = np.zeros((rows, p.image_width), np.float32)
    accel = po.make_accel()
    rc = oracle._render(C.byref(a), C.byref(accel), C.byref(p), img.ctypes.data, sig.ctypes.data, None, 2)
    assert rc == 0
    del keep
    return img, sig


def worker(rank, world, port, scene_file, width, height, rpp, halo, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene = po.ingest_xml(os.path.join(XML, scene_file), True)
        zoom = scene["image_height"] / height
        plan = rd.StripPlan(height, width, world, rank, halo)
        buf = rd.FrameBuffers(plan, torch.device("cpu"))
        render_strips, blur_rows = oracle_callbacks(scene, width, height, rpp, zoom)
        frame = rd.render_frame(buf, render_strips, blur_rows, use_blur=True)
        if rank == 0:
            np.save(out_path, frame.numpy())
            assert frame.shape == (height, width, 4)
        else:
            assert frame is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def single_process_frame(scene_file, width, height, rpp):
    oracle = po.Oracle("port")
    scene = po.ingest_xml(os.path.join(XML, scene_file), True)
    p = po.make_params(width, height, rpp, zoom_factor=scene["image_height"] / height)
    img, sig, _ = oracle.render(scene, p)
    return oracle.blur(img, sig), float(np.nanmax(sig)), float(scene["blur"][: scene["n_blur"]].max())


@pytest.mark.parametrize("world,scene_file,height,case", [
    (2, "DiffusionCurvePack/lady_bug.xml", 60, "blur: all-gather, local band blur, gather"),
    (3, "DiffusionCurvePack/lady_bug.xml", 70, "blur, 3 ranks, ragged last strip"),
    (2, "DiffusionCurvePack/face.xml", 40, "blur reach deeper than a band"),
    (2, "arch.xml", 37, "no blur: gather of packed strips, uneven strip counts"),
    (3, "arch.xml", 2 * rd.STRIP, "fewer strips than ranks"),
])
def test_strips_reassemble_bit_exactly(world, scene_file, height, case, tmp_path):
    po.build()
    width, rpp = 36, 6
    want, sigma_seen, sigma_bound = single_process_frame(scene_file, width, height, rpp)
    assert sigma_seen <= sigma_bound + 1e-6  # the static bound really bounds the blur map
    halo = rd.halo_rows(sigma_bound)
    out = str(tmp_path / "frame.npy")
    mp.spawn(worker, args=(world, free_port(), scene_file, width, height, rpp, halo, out), nprocs=world, join=True)
    got = np.load(out)
    assert np.array_equal(got[..., :3].view(np.uint32), want[..., :3].view(np.uint32)), case


def test_strip_plan_arithmetic():
    assert [rd.row_band(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert rd.halo_rows(0.0) == 0 and rd.halo_rows(1.5) == 5 and rd.halo_rows(16) == 48
    S = rd.STRIP
    p = rd.StripPlan(1080, 1920, 8, 3, 0)
    n = -(-1080 // S)  # strips
    assert p.n_strips == n and p.packed_rows == -(-n // 8) * S and p.local_strips == len(range(3, n, 8))
    assert rd.StripPlan(1080, 1920, 8, 7, 0).local_strips == len(range(7, n, 8))
    idx = p.source_index()
    last = (1079 // S)
    assert idx[0] == 0 and idx[S] == p.packed_rows and idx[8 * S] == S
    assert idx[1079] == (last % 8) * p.packed_rows + (last // 8) * S + 1079 % S
    assert len(set(idx.tolist())) == 1080
    q = rd.StripPlan(100, 8, 3, 1, 5)
    bi = q.band_index()
    assert bi[0] == 0 and bi[34] == q.max_band_rows and bi[99] == 2 * q.max_band_rows + 32
    assert rd.StripPlan(2 * S, 8, 3, 2, 0).local_strips == 0


def test_oracle_strips_equal_rows_of_the_full_frame():
    oracle = po.Oracle("port")
    scene = po.ingest_xml(os.path.join(XML, "arch.xml"), True)
    full, _, _ = oracle.render(scene, po.make_params(20, 50, 8, zoom_factor=10.0))
    p = po.make_params(20, 50, 8, zoom_factor=10.0, strip_stride=2, strip_offset=1)
    S = rd.STRIP
    mine = [t for t in range(-(-50 // S)) if t % 2 == 1]  # strips of rank 1 of 2
    img, _ = render_packed(oracle, scene, p, len(mine) * S)
    for k, t in enumerate(mine):
        rows = min(S, 50 - t * S)
        assert np.array_equal(img[k * S:k * S + rows].view(np.uint32), full[t * S:t * S + rows].view(np.uint32))


# ---- the peer-memory form (distributed.render_frame_peer) with shared host memory standing in for NVLink peer memory ----

class SharedPeerBuffers:
    """What rdc_peer_frames (csrc/peer.cu) is on GPUs: every rank can store into every rank's frame buffers. Here the
    buffers are shared-memory CPU tensors and an 'address' is the tensor itself; render_frame_peer is the Python statement of
    the sequence rdc_peer_render_frame enqueues (which buffer, which barrier, in which order)."""

    def __init__(self, plan, images, sigmas, frames):
        self.plan = plan
        self.image_ptrs, self.sigma_ptrs, self.frame_ptrs = images, sigmas, frames  # [rank], [rank], [turn][rank]
        self.full_image, self.full_sigma = images[plan.rank], sigmas[plan.rank]
        self.frames = [frames[0][plan.rank], frames[1][plan.rank]]
        self.scratch = None
        self.turn = 0
        self.barriers = 0

    def barrier(self):
        self.barriers += 1
        dist.barrier()


def peer_worker(rank, world, port, scene_file, width, height, rpp, halo, images, sigmas, frames, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = po.Oracle("port")
        scene = po.ingest_xml(os.path.join(XML, scene_file), True)
        zoom = scene["image_height"] / height
        plan = rd.StripPlan(height, width, world, rank, halo)
        buf = SharedPeerBuffers(plan, images, sigmas, frames)
        frame_no = [0]

        def render_to(image_targets, sigma_targets, stride, offset):
            p = po.make_params(width, height, rpp, zoom_factor=zoom, strip_stride=stride, strip_offset=offset, frame=frame_no[0])
            img, sig = render_packed(oracle, scene, p, plan.packed_rows)
            for k, t in enumerate(plan.strips_of(offset)):  # every finished row straight to its place in every target
                rows = min(rd.STRIP, height - t * rd.STRIP)
                for image, sigma in zip(image_targets, sigma_targets):
                    image[t * rd.STRIP:t * rd.STRIP + rows].copy_(torch.from_numpy(img[k * rd.STRIP:k * rd.STRIP + rows]))
                    sigma[t * rd.STRIP:t * rd.STRIP + rows].copy_(torch.from_numpy(sig[k * rd.STRIP:k * rd.STRIP + rows]))

        def blur_rows(dest, source, sigma, scratch, rows, row_begin, row_end, halo_rows):
            out = oracle.blur(source[:rows].numpy(), sigma[:rows].numpy(), threads=2)
            dest[row_begin:row_end].copy_(torch.from_numpy(out[row_begin:row_end]))

        hooks = []
        got = []
        for f in range(3):  # three frames: both frame buffers are used, one of them twice
            frame_no[0] = f
            frame = rd.render_frame_peer(buf, render_to, blur_rows, use_blur=True, before_barrier=lambda: hooks.append(f))
            if rank == 0:
                assert frame is frames[f % 2][0]
                got.append(frame.numpy().copy())
            else:
                assert frame is None
        assert hooks == [0, 1, 2] and buf.barriers == (6 if halo > 0 else 3)
        if rank == 0:
            np.save(out_path, np.stack(got))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,scene_file,height,case", [
    (2, "DiffusionCurvePack/lady_bug.xml", 44, "blur: stores into every rank's frame, band blur into rank 0's"),
    (3, "arch.xml", 37, "no blur: stores into rank 0's frame only"),
])
def test_peer_memory_form_reassembles_bit_exactly(world, scene_file, height, case, tmp_path):
    po.build()
    width, rpp = 36, 6
    oracle = po.Oracle("port")
    scene = po.ingest_xml(os.path.join(XML, scene_file), True)
    sigma_bound = float(scene["blur"][: scene["n_blur"]].max())
    halo = rd.halo_rows(sigma_bound)
    shared = lambda *shape: torch.full(shape, float("nan"), dtype=torch.float32).share_memory_()  # noqa: E731
    images = [shared(height, width, 4) for _ in range(world)]
    sigmas = [shared(height, width) for _ in range(world)]
    frames = [[shared(height, width, 4) for _ in range(world)] for _ in range(2)]
    out = str(tmp_path / "frames.npy")
    mp.spawn(peer_worker, args=(world, free_port(), scene_file, width, height, rpp, halo, images, sigmas, frames, out),
             nprocs=world, join=True)
    got = np.load(out)
    for f in range(3):
        p = po.make_params(width, height, rpp, zoom_factor=scene["image_height"] / height, frame=f)
        img, sig, _ = oracle.render(scene, p)
        want = oracle.blur(img, sig) if halo > 0 else img
        assert np.array_equal(got[f][..., :3].view(np.uint32), want[..., :3].view(np.uint32)), (case, f)
    assert not np.array_equal(got[0].view(np.uint32), got[1].view(np.uint32))  # the frame number keys the generator
