"""Host-side pieces of the C ABI that need no GPU: the image comparison behind tools/rdc_diff.py (rdc_psnr), the blur reach
of a scene (rdc_host_scene_halo_rows), and the scene cache loader's treatment of damaged files."""
import json
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from raytracingdiffusioncurves_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
XML = os.path.join(ROOT, "tests", "golden", "xmls")


def test_psnr_is_rgb_only_and_nan_aware():
    rng = np.random.default_rng(3)
    a = rng.random((20, 30, 4), dtype=np.float32)
    b = a.copy()
    assert api.psnr(a, b) == (float("inf"), 0.0)
    b[..., 3] += 0.5  # alpha is not compared (the reference never writes image.w, DeviceCode.cu:176-178)
    assert api.psnr(a, b) == (float("inf"), 0.0)
    b[..., :3] += np.float32(0.01)
    p, m = api.psnr(a, b)
    assert abs(p - 40.0) < 0.01 and abs(m - 0.01) < 1e-6
    a[2, 3, :3] = np.nan  # all rays missed in both images: skipped
    b[2, 3, :3] = np.nan
    p2, _ = api.psnr(a, b)
    assert abs(p2 - 40.0) < 0.01
    b[5, 6, 0] = np.nan  # NaN on one side only counts as a full-scale error
    p3, m3 = api.psnr(a, b)
    assert p3 < p2 and m3 == 1.0
    assert api.lib.rdc_psnr(None, None, 0, None, None) == -1


def test_rdc_diff_tool_compares_a_float_dump_with_an_8_bit_image(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(4)
    img = rng.random((24, 32, 4), dtype=np.float32)
    img[..., 3] = 1.0
    dump = tmp_path / "a.f32"
    img.tofile(dump)
    rgba = api.image_to_rgba8(img, True)  # what OptixHello --out writes for an Orzan save: rows flipped
    png = tmp_path / "a.png"
    assert api.lib.rdc_write_png(str(png).encode(), rgba.ctypes.data, 32, 24) == 0
    side = tmp_path / "side.png"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "rdc_diff.py"), str(dump), str(png), "--size", "32x24", "--flip-a",
                        "--side-by-side", str(side)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["psnr_db"] == float("inf") and out["max_abs"] == 0.0  # the float side is quantised like the screenshot
    with Image.open(side) as f:
        assert f.size == (96, 24)
    unflipped = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "rdc_diff.py"), str(dump), str(png), "--size", "32x24"],
                               capture_output=True, text=True)
    assert json.loads(unflipped.stdout.strip().splitlines()[-1])["psnr_db"] < 20.0


@pytest.mark.parametrize("name,depth,want", [("arch.xml", 2, 0), ("DiffusionCurvePack/face.xml", 2, 48), ("DiffusionCurvePack/lady_bug.xml", 31, 21),
                                             ("PortalDemo.xml", 31, 0)])
def test_halo_rows_follow_the_blur_taps(name, depth, want):
    """helperKernels.cu:65,74: taps reach ceil(3 sigma) rows; sigma is a weighted mean of blur stops (times one more stop per
    portal passed, DeviceCode.cu:311)."""
    host = api.HostScene.from_xml_file(os.path.join(XML, name))
    assert host.halo_rows(depth) == want
    assert host.halo_rows(depth) == int(np.ceil(3.0 * host.max_blur(depth)))


def test_scene_cache_loader_rejects_damaged_files(tmp_path):
    host = api.HostScene.from_xml_file(os.path.join(XML, "DiffusionCurvePack/zephyr.xml"))
    good = tmp_path / "good.rdc"
    host.save(str(good))
    again = api.HostScene.from_cache(str(good)).to_numpy()
    mine = host.to_numpy()
    assert all(np.array_equal(np.asarray(again[k]), np.asarray(mine[k])) for k in mine)
    data = good.read_bytes()
    cases = {
        "truncated": data[: len(data) // 2],
        "bad magic": b"NOTACACHE" + data[9:],
        "huge count": data[:36] + struct.pack("<Q", 1 << 40) + data[44:],  # first vector claims 2^40 elements: no allocation follows
        "zero size": data[:8] + struct.pack("<ii", 0, 512) + data[16:],
        "empty": b"",
    }
    for what, blob in cases.items():
        bad = tmp_path / "bad.rdc"
        bad.write_bytes(blob)
        with pytest.raises(api.RdcError) as err:
            api.HostScene.from_cache(str(bad))
        assert err.value.code in (-2, -3), what
