"""Parity at BASELINE.json's real sizes: the CUDA path, through the C ABI, against the CPU oracle on the same rays.

Every primary ray's first-hit chord index bit-exact, per-pixel RGB within 1e-4 absolute, PSNR printed.
The oracle is oracle/_ref — the reference's own DeviceCode.cu compiled for the host — wherever the switches are
the ones it is compiled with, and the restated port (bit-identical to _ref, tests/test_oracle_cpu.py) where they
differ (trace depth 31). Both narrow the chords a ray is tested against with the uniform grid of oracle_common.h
(== brute force bit for bit, tests/test_oracle_cpu.py), which is what makes these sizes affordable on the host:

  (a) config 2, the headline: arch.xml 1920x1080 @128 — the FULL frame, all 265 M rays, in 120-row bands;
  (b) config 1: arch.xml 512x512 @128 — the full frame;
  (c) config 3: every bundled XML at 3840x2160 @256 — three 4-row bands (whole tile rows) of the real frame each;
  (d) config 4: PortalDemo.xml at 1920x1080 @128, trace depth 31 and 2 — three 8-row bands;
  (e) config 5: the real 100 000-curve synthetic scene at 8192x8192 @512 — three 4-row bands of the real frame
      (local run table, deferred rays, tree), plus the same bands with the route forced to the tree.
"""
import os

import numpy as np
import pytest

from helpers import GpuRenderer, all_scene_files, compare_images, copy_params, XML_DIR
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

RGB_TOL = 1e-4


@pytest.fixture(scope="module")
def api():
    import torch

    assert torch.cuda.is_available(), "these tests need a CUDA device"
    from raytracingdiffusioncurves_b200 import api as _api

    return _api


@pytest.fixture(scope="module")
def shipped_oracle(port_oracle):
    """oracle/_ref when it travelled with the snapshot (shipped switches only), else the port."""
    return po.Oracle("reference") if po.Oracle.reference_available() else port_oracle


def check_band(api, renderer, oracle, scene, params, label, want_stats=False):
    oimg, oblur, ohits = oracle.render(scene, params, want_hits=True, search="grid")
    out = renderer.render(copy_params(params, api.FrameParams), want_hits=True, want_stats=want_stats)
    wrong = int((out["hits"] != ohits).sum())
    assert wrong == 0, f"{label}: {wrong} of {ohits.size} first hits differ from the oracle"
    d, psnr = compare_images(out["image"], oimg, RGB_TOL)
    m = ~np.isnan(oblur)
    assert np.array_equal(np.isnan(out["blur_map"]), np.isnan(oblur))
    assert np.max(np.abs(out["blur_map"][m] - oblur[m]), initial=0) <= 1e-4 * max(1.0, float(np.nanmax(oblur, initial=0)))
    return d, psnr, int((ohits != 0xFFFFFFFF).sum()), out


def test_headline_full_frame_every_ray(xml_dir, api, shipped_oracle):
    """(a) config 2: all 265 420 800 primary rays of the headline frame."""
    path = os.path.join(xml_dir, "arch.xml")
    scene = po.ingest_xml(path, True)
    r = GpuRenderer(path)
    w, h, n = 1920, 1080, 128
    worst, mse_sum, rays, hits = 0.0, 0.0, 0, 0
    for b in range(0, h, 120):
        p = po.make_params(w, h, n, zoom_factor=512 / h, row_begin=b, row_end=b + 120)
        d, psnr, nh, _ = check_band(api, r, shipped_oracle, scene, p, f"rows {b}:{b + 120}")
        worst = max(worst, d)
        mse_sum += 0.0 if psnr == float("inf") else 10 ** (-psnr / 10)
        rays += 120 * w * n
        hits += nh
    assert rays == 265_420_800
    psnr = float("inf") if mse_sum == 0 else 10 * np.log10(1.0 / (mse_sum / 9))
    print(f"headline frame vs {shipped_oracle.kind}: {rays} rays ({hits} hits) bit-exact, max |rgb diff| {worst:.2e}, PSNR {psnr:.1f} dB")


def test_config1_full_frame(xml_dir, api, shipped_oracle):
    """(b) config 1: arch.xml at its own 512x512, 128 rays per pixel, zoom 1."""
    path = os.path.join(xml_dir, "arch.xml")
    scene = po.ingest_xml(path, True)
    p = po.make_params(512, 512, 128)
    d, psnr, nh, _ = check_band(api, GpuRenderer(path), shipped_oracle, scene, p, "512x512")
    print(f"config 1 vs {shipped_oracle.kind}: 33 554 432 rays ({nh} hits) bit-exact, max |rgb diff| {d:.2e}, PSNR {psnr:.1f} dB")


ALL_SCENES = [os.path.relpath(f, XML_DIR) for f in all_scene_files()]


@pytest.mark.parametrize("name", ALL_SCENES)
def test_config3_bands_of_the_4k_frame(name, xml_dir, api, shipped_oracle):
    """(c) config 3: 3840x2160 @256 framing of every bundled scene; whole tile rows at 1/4, 1/2 and 3/4 of the height."""
    path = os.path.join(xml_dir, name)
    scene = po.ingest_xml(path, True)
    r = GpuRenderer(path)
    w, h, n = 3840, 2160, 256
    zoom = scene["image_height"] / h
    total = 0
    for b in (540, 1080, 1620):
        p = po.make_params(w, h, n, zoom_factor=zoom, row_begin=b, row_end=b + 4)
        d, psnr, nh, _ = check_band(api, r, shipped_oracle, scene, p, f"{name} rows {b}:{b + 4}")
        total += nh
    assert total > 0, "the bands never meet the scene"
    print(f"{name}: 3 x {4 * w * n} rays of the 4K frame bit-exact ({total} hits), last band max |rgb diff| {d:.2e}, PSNR {psnr:.1f} dB")


@pytest.mark.parametrize("depth", [2, 31])
def test_config4_portals_at_1080p(depth, xml_dir, api, port_oracle, shipped_oracle):
    """(d) config 4: PortalDemo.xml 1920x1080 @128; depth 31 against the port (the reference build is compiled for 2)."""
    path = os.path.join(xml_dir, "PortalDemo.xml")
    scene = po.ingest_xml(path, True)
    oracle = shipped_oracle if depth == 2 else port_oracle
    r = GpuRenderer(path)
    w, h, n = 1920, 1080, 128
    for b in (264, 536, 808):
        p = po.make_params(w, h, n, zoom_factor=512 / h, max_trace_depth=depth, row_begin=b, row_end=b + 8)
        d, psnr, nh, _ = check_band(api, r, oracle, scene, p, f"depth {depth} rows {b}:{b + 8}")
        print(f"PortalDemo depth {depth} rows {b}:{b + 8} vs {oracle.kind}: {nh} hits, max |rgb diff| {d:.2e}, PSNR {psnr:.1f} dB")


@pytest.fixture(scope="module")
def synth100k(api, tmp_path_factory):
    """BASELINE.json configs[4]: rdc_synth_xml(100000, 8192, 8192), ingested once by the product (the Python restatement
    of the loader takes half a minute on 76 MB of XML; ingest parity has its own tests)."""
    xml = api.synth_xml(100000, 8192, 8192)
    f = tmp_path_factory.mktemp("synth") / "synth100k.xml"
    f.write_bytes(xml)
    r = GpuRenderer(str(f))
    return r, r.host.to_numpy()


@pytest.mark.parametrize("route", ["auto", "tree"])
def test_config5_bands_of_the_8k_frame(route, synth100k, api, port_oracle):
    """(e) config 5: the real 100 k-curve scene at 8192x8192 @512: tile rows near the top, the middle and the bottom."""
    r, scene = synth100k
    assert r.scene.stats.n_curves == 100000 and r.scene.stats.n_chords > 1_000_000
    w = h = 8192
    n = 512
    for b in (1024, 4096, 7168):
        p = po.make_params(w, h, n, zoom_factor=1.0, row_begin=b, row_end=b + 4,
                           route=api.ROUTE_AUTO if route == "auto" else api.ROUTE_TREE)
        d, psnr, nh, out = check_band(api, r, port_oracle, scene, p, f"synth rows {b}:{b + 4}", want_stats=True)
        traced, _, _, _, deferred, gathered = out["stats"]
        if route == "auto":
            assert gathered > 0 and 0 < deferred < traced, "the local run table (with deferred rays) was not exercised"
        else:
            assert gathered == 0 and deferred == 0
        print(f"synth100k 8192^2@512 rows {b}:{b + 4} ({route}): {4 * w * n} rays bit-exact ({nh} hits, {deferred} deferred to the tree), "
              f"max |rgb diff| {d:.2e}, PSNR {psnr:.1f} dB")


def test_blur_parity_at_1080p(xml_dir, api, port_oracle):
    """The blur at a real frame size and the scene's real sigmas (face.xml: up to 16 output pixels, 97 taps)."""
    path = os.path.join(xml_dir, "DiffusionCurvePack/face.xml")
    r = GpuRenderer(path)
    w, h, n = 1920, 1080, 8
    out = r.render(api.default_frame_params(w, h, n, zoom_factor=512 / h), blur=True)
    assert out["max_sigma"] > 8.0
    want = port_oracle.blur(out["image"], out["blur_map"])[..., :3]
    got = out["blurred"][..., :3]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    m = ~np.isnan(want)
    d = float(np.max(np.abs(got[m] - want[m]), initial=0))
    assert d <= 1e-5
    print(f"blur 1920x1080, sigma up to {out['max_sigma']:.1f}: max |diff| {d:.2e}")
