"""N > 1 host logic on CPU: world_size 2 and 3 over gloo. The band plan, the blur-halo exchange and the
gather of raytracingdiffusioncurves_b200.distributed run for real; what renders and blurs a band is the
CPU oracle (injected — the product module itself never touches oracle/). The assembled frame must equal the
single-process frame bit for bit."""
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

    def render_band(image_rows, sigma_rows, row_begin, row_end):
        p = po.make_params(width, height, rpp, zoom_factor=zoom, row_begin=row_begin, row_end=row_end)
        img, sig, _ = oracle.render(scene, p, threads=2)
        image_rows.copy_(torch.from_numpy(img))
        sigma_rows.copy_(torch.from_numpy(sig))

    def blur_rows(dest, source, sigma, scratch, rows, row_begin, row_end):
        out = oracle.blur(source[:rows].numpy(), sigma[:rows].numpy(), threads=2)
        dest[row_begin:row_end].copy_(torch.from_numpy(out[row_begin:row_end]))

    return render_band, blur_rows


def worker(rank, world, port, scene_file, width, height, rpp, halo, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene = po.ingest_xml(os.path.join(XML, scene_file), True)
        zoom = scene["image_height"] / height
        plan = rd.BandPlan(height, width, world, rank, halo)
        bands = rd.FrameBands(plan, torch.device("cpu"))
        render_band, blur_rows = oracle_callbacks(scene, width, height, rpp, zoom)
        frame = rd.render_frame(bands, render_band, blur_rows, use_blur=True)
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
    (2, "DiffusionCurvePack/lady_bug.xml", 60, "halo exchange"),      # max blur stop 7 -> halo 21 <= 30 rows
    (3, "DiffusionCurvePack/lady_bug.xml", 66, "halo exchange, 3 ranks"),
    (2, "DiffusionCurvePack/face.xml", 40, "halo deeper than a band"),   # max blur stop 16 -> halo 48 > 20 rows
    (2, "arch.xml", 31, "no blur, uneven bands"),
])
def test_bands_reassemble_bit_exactly(world, scene_file, height, case, tmp_path):
    po.build()
    width, rpp = 36, 6
    want, sigma_seen, sigma_bound = single_process_frame(scene_file, width, height, rpp)
    assert sigma_seen <= sigma_bound + 1e-6  # the static halo bound really bounds the blur map
    halo = rd.halo_rows(sigma_bound)
    out = str(tmp_path / "frame.npy")
    mp.spawn(worker, args=(world, free_port(), scene_file, width, height, rpp, halo, out), nprocs=world, join=True)
    got = np.load(out)
    assert np.array_equal(got[..., :3].view(np.uint32), want[..., :3].view(np.uint32)), case


def test_band_plan_arithmetic():
    assert [rd.row_band(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert rd.halo_rows(0.0) == 0 and rd.halo_rows(1.5) == 5 and rd.halo_rows(16) == 48
    p = rd.BandPlan(1080, 1920, 8, 3, 48)
    assert (p.begin, p.end, p.top, p.bottom, p.exchange, p.buffer_rows) == (405, 540, 48, 48, True, 231)
    edge = rd.BandPlan(1080, 1920, 8, 0, 48)
    assert (edge.top, edge.bottom) == (0, 48)
    deep = rd.BandPlan(64, 32, 8, 2, 48)
    assert not deep.exchange and deep.top == 0 and deep.bottom == 0
    single = rd.BandPlan(64, 32, 1, 0, 48)
    assert not single.exchange and single.buffer_rows == 64
