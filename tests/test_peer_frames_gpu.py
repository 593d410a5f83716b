"""The native multi-GPU path (rdc_peer_*, csrc/peer.cu) on whatever GPUs the box has: several ranks in ONE process —
each with its own rdc_scene, stream and peer frames, on device rank % device_count — run the real protocol (stores
into the consumers' frames, barrier kernels spinning on flags in peer memory, per-rank copies into one host frame).
The assembled frame must equal the one-call frame bit for bit, with and without blur, over several frames in a row
(both frame buffers, both host slots)."""
import os

import numpy as np
import pytest

from helpers import bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch

    assert torch.cuda.is_available(), "these tests need a CUDA device"
    from raytracingdiffusioncurves_b200 import api as _api

    return _api


class Ranks:
    def __init__(self, api, path, width, height, world):
        import torch

        self.api, self.torch, self.world = api, torch, world
        n_dev = torch.cuda.device_count()
        self.devices = [r % n_dev for r in range(world)]
        self.host = api.HostScene.from_xml_file(path)
        self.scenes, self.streams, self.frames = [], [], []
        for r, d in enumerate(self.devices):
            with torch.cuda.device(d):
                s = torch.cuda.Stream()
                self.streams.append(s)
                self.scenes.append(api.Scene(self.host.arrays, None, s.cuda_stream))
                self.frames.append(api.PeerFrames(width, height, r, world))
        api.PeerFrames.connect_local(self.frames)

    def reserve(self, params):
        """rdc_scene_reserve on every rank: with one host thread feeding every rank, nothing may allocate (and so wait for
        the device) between two ranks' barrier kernels."""
        for r, scene, _, stream in self.each():
            scene.reserve(params, False, stream)

    def each(self):
        for r, d in enumerate(self.devices):
            with self.torch.cuda.device(d):
                yield r, self.scenes[r], self.frames[r], self.streams[r].cuda_stream

    def sync(self):
        for d in set(self.devices):
            self.torch.cuda.synchronize(d)

    def close(self):
        self.sync()
        for f in self.frames:
            f.close()


def one_call_frame(api, path, width, height, rpp, zoom, frame, blur):
    import torch

    host = api.HostScene.from_xml_file(path)
    stream = torch.cuda.current_stream().cuda_stream
    scene = api.Scene(host.arrays, None, stream)
    out = torch.empty((height, width, 4), dtype=torch.float32).pin_memory()
    scene.render_frame_to_host(api.default_frame_params(width, height, rpp, zoom_factor=zoom, frame=frame), blur, out.data_ptr(), stream)
    return out.numpy().copy()


CASES = [(2, "DiffusionCurvePack/lady_bug.xml", 96, 76, "blur"), (3, "arch.xml", 80, 61, "no blur, ragged last strip"),
         (3, "DiffusionCurvePack/face.xml", 72, 50, "blur reach deeper than a band"), (8, "arch.xml", 64, 40, "fewer strips than ranks")]


@pytest.mark.parametrize("world,name,width,height,case", CASES, ids=[c[4] for c in CASES])
def test_device_consumer_frame_equals_one_call_frame(world, name, width, height, case, xml_dir, api):
    import torch

    path = os.path.join(xml_dir, name)
    rpp, zoom = 8, 512 / height
    ranks = Ranks(api, path, width, height, world)
    halo = ranks.host.halo_rows(2)
    ranks.reserve(api.default_frame_params(width, height, rpp, zoom_factor=zoom))
    try:
        for f in range(3):
            ptr = 0
            for r, scene, frames, stream in ranks.each():
                p = api.default_frame_params(width, height, rpp, zoom_factor=zoom, frame=f)
                got = frames.render_frame(scene, p, True, halo, stream)
                if r == 0:
                    ptr = got
                else:
                    assert got == 0
            ranks.sync()
            for _, _, frames, _ in ranks.each():
                frames.status()
            out = torch.empty((height, width, 4), dtype=torch.float32)
            with torch.cuda.device(ranks.devices[0]):
                import ctypes

                rc = ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ptr),
                                                            ctypes.c_size_t(out.numel() * 4), 2)
                assert rc == 0
            want = one_call_frame(api, path, width, height, rpp, zoom, f, True)
            assert np.array_equal(bits(out.numpy()[..., :3]), bits(want[..., :3])), (case, f)
    finally:
        ranks.close()


@pytest.mark.parametrize("world,name,width,height,case", CASES, ids=[c[4] for c in CASES])
def test_host_consumer_frame_equals_one_call_frame(world, name, width, height, case, xml_dir, api):
    import torch

    path = os.path.join(xml_dir, name)
    rpp, zoom = 8, 512 / height
    ranks = Ranks(api, path, width, height, world)
    halo = ranks.host.halo_rows(2)
    hosts = [torch.full((height, width, 4), float("nan"), dtype=torch.float32).pin_memory() for _ in range(4)]
    ranks.reserve(api.default_frame_params(width, height, rpp, zoom_factor=zoom))
    try:
        for f in range(4):  # enqueued back to back: copies of one frame overlap the next frame's rendering
            for r, scene, frames, stream in ranks.each():
                p = api.default_frame_params(width, height, rpp, zoom_factor=zoom, frame=f)
                frames.frame_to_host(scene, p, True, halo, hosts[f].data_ptr(), stream)
        for _, _, frames, _ in ranks.each():
            frames.wait()
        ranks.sync()
        for f in range(4):
            want = one_call_frame(api, path, width, height, rpp, zoom, f, True)
            assert np.array_equal(bits(hosts[f].numpy()[..., :3]), bits(want[..., :3])), (case, f)
    finally:
        ranks.close()


@pytest.mark.parametrize("name", ["arch.xml", "DiffusionCurvePack/lady_bug.xml"])
def test_group_mode_through_peer_frames(name, xml_dir, api):
    """Eight units per tile (group mode: a block's warps share a tile's table, partial sums in shared memory) with the finished
    pixels stored into the consumers' frames: three ranks, 128 rays per pixel, the summation order pinned on both sides."""
    import ctypes

    import torch

    path = os.path.join(xml_dir, name)
    width, height, rpp, world = 96, 72, 128, 3
    zoom = 512 / height
    ranks = Ranks(api, path, width, height, world)
    halo = ranks.host.halo_rows(2)
    ranks.reserve(api.default_frame_params(width, height, rpp, zoom_factor=zoom, units_per_tile=8))
    try:
        ptr = 0
        for r, scene, frames, stream in ranks.each():
            got = frames.render_frame(scene, api.default_frame_params(width, height, rpp, zoom_factor=zoom, units_per_tile=8), True, halo, stream)
            ptr = got if r == 0 else ptr
        ranks.sync()
        out = torch.empty((height, width, 4), dtype=torch.float32)
        with torch.cuda.device(ranks.devices[0]):
            assert ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(out.numel() * 4), 2) == 0
        host = api.HostScene.from_xml_file(path)
        s = torch.cuda.current_stream().cuda_stream
        scene = api.Scene(host.arrays, None, s)
        want = torch.empty((height, width, 4), dtype=torch.float32).pin_memory()
        for units in (8, 4):  # group mode and the global partial sums add in the same order: the same bits
            scene.render_frame_to_host(api.default_frame_params(width, height, rpp, zoom_factor=zoom, units_per_tile=units), True, want.data_ptr(), s)
            if units == 8:
                assert np.array_equal(bits(out.numpy()[..., :3]), bits(want.numpy()[..., :3]))
    finally:
        ranks.close()


def test_barrier_reports_a_missing_rank(api):
    """A rank that never arrives raises the error flag after the time limit instead of wedging the GPU."""
    import torch

    a, b = api.PeerFrames(16, 16, 0, 2), api.PeerFrames(16, 16, 1, 2)
    api.PeerFrames.connect_local([a, b])
    a.barrier(torch.cuda.current_stream().cuda_stream)  # rank 1 never calls it
    torch.cuda.synchronize()
    with pytest.raises(api.RdcError):
        a.status()
    a.close()
    b.close()


def test_shared_host_frame_is_visible_across_handles(api):
    name = f"/rdc_test_{os.getpid()}"
    a = api.HostFrame(name, 4096, True)
    b = api.HostFrame(name, 4096, False)
    a.numpy((1024,))[:] = np.arange(1024, dtype=np.float32)
    assert np.array_equal(b.numpy((1024,)), np.arange(1024, dtype=np.float32))
    b.owner = False
    b.close()
    a.close()


def test_optixhello_gpus_flag_renders_the_same_image(xml_dir, tmp_path):
    """OptixHello --gpus N (N ranks in one process; on a one-GPU box they share the device): same pixels as one GPU."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "raytracingdiffusioncurves_b200", "OptixHello")
    outs = []
    for gpus in (1, 3):
        out = tmp_path / f"lady_{gpus}.f32"
        r = subprocess.run([exe, "tests/golden/xmls/DiffusionCurvePack/lady_bug.xml", "16", "--width", "160", "--height", "120", "--frames", "2",
                            "--gpus", str(gpus), "--dump-f32", str(out)], cwd=root, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "Average frame time  : " in r.stdout
        outs.append(np.fromfile(out, np.float32))
    assert outs[0].size == 160 * 120 * 4 and np.array_equal(bits(outs[0]), bits(outs[1]))
