#!/usr/bin/env python
"""rdc_diff — compare two renderings of the same view (the diff tool).

    python tools/rdc_diff.py A B [--size WxH] [--flip-a] [--flip-b] [--side-by-side out.png]

A and B are float dumps written by `OptixHello --dump-f32` (RGBA float32, row 0 first; needs --size) or 8-bit images
(PNG / JPG / PPM, e.g. the reference's F11 screenshots). A float image is first put through the screenshot's conversion,
min(v * 255, 255) (glfw_events.cpp:73-94, rdc_image_to_rgba8), when the other side is 8-bit. Prints one JSON line:
PSNR (rdc_psnr) and the largest absolute difference, on the [0,1] scale.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load(path, size, flip):
    from raytracingdiffusioncurves_b200 import api

    if path.endswith((".f32", ".raw", ".bin")):
        if not size:
            raise SystemExit(f"{path}: a float dump needs --size WxH")
        w, h = (int(v) for v in size.lower().split("x"))
        img = np.fromfile(path, np.float32).reshape(h, w, 4)
        return (img[::-1].copy() if flip else img), True
    from PIL import Image

    rgb = np.asarray(Image.open(path).convert("RGB"), np.float32) / 255.0
    img = np.concatenate([rgb, np.ones(rgb.shape[:2] + (1,), np.float32)], -1)
    return (img[::-1].copy() if flip else img), False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("a")
    ap.add_argument("b")
    ap.add_argument("--size")
    ap.add_argument("--flip-a", action="store_true", help="flip A vertically (Orzan saves are rendered bottom-up, glfw_events.cpp:92)")
    ap.add_argument("--flip-b", action="store_true")
    ap.add_argument("--side-by-side")
    args = ap.parse_args()
    from raytracingdiffusioncurves_b200 import api

    a, a_float = load(args.a, args.size, args.flip_a)
    b, b_float = load(args.b, args.size, args.flip_b)
    if a.shape != b.shape:
        raise SystemExit(f"sizes differ: {a.shape[1]}x{a.shape[0]} vs {b.shape[1]}x{b.shape[0]}")
    if a_float != b_float:  # bring the float side to the screenshot's 8-bit scale
        def quantise(x):
            return api.image_to_rgba8(x, False).astype(np.float32) / 255.0
        a, b = (quantise(a), b) if a_float else (a, quantise(b))
    psnr, max_abs = api.psnr(a, b)
    print(json.dumps({"a": args.a, "b": args.b, "width": a.shape[1], "height": a.shape[0], "psnr_db": psnr, "max_abs": max_abs,
                      "mean_abs": float(np.nanmean(np.abs(a[..., :3] - b[..., :3])))}))
    if args.side_by_side:
        from PIL import Image

        def u8(x):
            return np.nan_to_num(np.clip(x[..., :3], 0, 1) * 255).astype(np.uint8)
        diff = np.clip(np.abs(np.nan_to_num(a[..., :3]) - np.nan_to_num(b[..., :3])) * 4, 0, 1)
        Image.fromarray(np.concatenate([u8(a), u8(b), (diff * 255).astype(np.uint8)], 1)).save(args.side_by_side)


if __name__ == "__main__":
    main()
