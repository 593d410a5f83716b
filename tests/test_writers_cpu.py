"""Headless image writers: the F11 screenshot's pixel conversion and file formats
(glfw_events.cpp:73-94), checked by decoding the files with an independent decoder."""
import numpy as np
import pytest

from raytracingdiffusioncurves_b200 import api

Image = pytest.importorskip("PIL.Image")


def make_pattern(h=123, w=201):  # neither size a multiple of the 8x8 JPEG block
    y, x = np.mgrid[0:h, 0:w]
    img = np.zeros((h, w, 4), np.uint8)
    img[..., 0] = x * 255 // (w - 1)
    img[..., 1] = y * 255 // (h - 1)
    img[..., 2] = ((x + y) % 64) * 4
    img[..., 3] = 255
    if h > 80 and w > 120:
        img[40:80, 50:120, :3] = np.random.default_rng(0).integers(0, 256, (40, 70, 3))  # a block of noise: every AC symbol class
    return img


def test_jpeg_decodes_to_the_image(tmp_path):
    img = make_pattern()
    h, w = img.shape[:2]
    smooth = np.ones((h, w), bool)
    smooth[36:84, 46:124] = False
    sizes = []
    for quality, min_psnr, max_smooth_err in ((100, 48.0, 3), (90, 34.0, 24)):
        path = tmp_path / f"q{quality}.jpg"
        assert api.lib.rdc_write_jpg(str(path).encode(), img.ctypes.data, w, h, quality) == 0
        with Image.open(path) as f:
            assert f.format == "JPEG" and f.size == (w, h)
            dec = np.asarray(f.convert("RGB")).astype(np.float64)
        err = dec - img[..., :3]
        psnr = 10 * np.log10(255.0 ** 2 / np.mean(err ** 2))
        assert psnr >= min_psnr, (quality, psnr)
        assert np.abs(err[smooth]).max() <= max_smooth_err
        sizes.append(path.stat().st_size)
    assert sizes[1] < sizes[0]
    # the reference hands stb an out-of-range quality (the stride): treated as 100
    path = tmp_path / "q_stride.jpg"
    assert api.lib.rdc_write_jpg(str(path).encode(), img.ctypes.data, w, h, w * 4) == 0
    assert path.stat().st_size == sizes[0]
    assert api.lib.rdc_write_jpg(str(path).encode(), None, w, h, 90) == -1


def test_png_and_ppm_are_lossless(tmp_path):
    img = make_pattern(37, 53)
    h, w = img.shape[:2]
    img[..., 3] = (np.arange(w) * 4 % 256).astype(np.uint8)
    png, ppm = tmp_path / "a.png", tmp_path / "a.ppm"
    assert api.lib.rdc_write_png(str(png).encode(), img.ctypes.data, w, h) == 0
    assert api.lib.rdc_write_ppm(str(ppm).encode(), img.ctypes.data, w, h) == 0
    with Image.open(png) as f:
        assert np.array_equal(np.asarray(f.convert("RGBA")), img)
    with Image.open(ppm) as f:
        assert np.array_equal(np.asarray(f.convert("RGB")), img[..., :3])


def test_screenshot_conversion_clamps_and_flips():
    """min(v*255, 255) per channel, NaN (all-miss pixel) -> 0, rows flipped for Orzan saves (glfw_events.cpp:73-92)."""
    f = np.array([[[0.0, 0.5, 1.0, 1.0], [2.0, float("nan"), 0.25, 1.0]],
                  [[1.0, 1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 0.0]]], np.float32)
    out = np.zeros((2, 2, 4), np.uint8)
    assert api.lib.rdc_image_to_rgba8(f.ctypes.data, 2, 2, 0, out.ctypes.data) == 0
    assert out[0, 0].tolist() == [0, 127, 255, 255] and out[0, 1].tolist() == [255, 0, 63, 255]
    flipped = np.zeros_like(out)
    assert api.lib.rdc_image_to_rgba8(f.ctypes.data, 2, 2, 1, flipped.ctypes.data) == 0
    assert np.array_equal(flipped, out[::-1])
