"""BASELINE.json configs[2]: every bundled XML at one frame size with the blur pass on.
    python tools/sweep_scenes.py [width height rpp] [--modes]
Prints one line per scene: ms per frame (render, blur), Grays/s, traversal mode the library picked.
--modes also times the frame with the local run table / whole-scene table switched off (tuning)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from helpers import all_scene_files, XML_DIR  # noqa: E402
from raytracingdiffusioncurves_b200 import api  # noqa: E402


def time_frames(scene, width, height, rpp, zoom, frames=3, route=0):
    stream = torch.cuda.current_stream().cuda_stream
    image = torch.empty((height, width, 4), dtype=torch.float32, device="cuda")
    sigma = torch.empty((height, width), dtype=torch.float32, device="cuda")
    scratch = torch.empty_like(image)
    flag = torch.zeros((1,), dtype=torch.float32, device="cuda")
    t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    best = None
    for f in range(frames):
        flag.zero_()
        p = api.default_frame_params(width, height, rpp, zoom_factor=zoom, frame=f, route=route)
        p.max_sigma = flag.data_ptr()
        t0.record()
        scene.render(p, image.data_ptr(), sigma.data_ptr(), stream)
        t1.record()
        api.gaussian_blur(image.data_ptr(), image.data_ptr(), sigma.data_ptr(), scratch.data_ptr(), width, height, 0, height,
                          flag.data_ptr(), stream)
        t2.record()
        torch.cuda.synchronize()
        r = (t0.elapsed_time(t1), t1.elapsed_time(t2), float(flag.item()))
        if f > 0 and (best is None or r[0] + r[1] < best[0] + best[1]):
            best = r
    return best


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    width, height, rpp = (int(v) for v in args[:3]) if len(args) >= 3 else (3840, 2160, 256)
    modes = "--modes" in sys.argv
    rows, total = [], 0.0
    for path in all_scene_files():
        name = os.path.relpath(path, XML_DIR)
        host = api.HostScene.from_xml_file(path)
        run_length = int(os.environ.get("RDC_RUN_LENGTH", "0"))  # tuning: chords per leaf (0 = the library's choice)
        tree = int(os.environ.get("RDC_TREE", "0"))  # api.TREE_*: 0 automatic, 1 Morton radix tree, 2 surface-area heuristic
        scene = api.Scene(host.arrays, api.default_accel_options(run_length=run_length, tree=tree), torch.cuda.current_stream().cuda_stream)
        zoom = host.arrays.image_height / height
        render_ms, blur_ms, smax = time_frames(scene, width, height, rpp, zoom)
        row = {"scene": name, "runs": scene.stats.n_runs, "chords": scene.stats.n_chords, "render_ms": round(render_ms, 3),
               "blur_ms": round(blur_ms, 3), "max_sigma": round(smax, 2),
               "grays_per_s": round(width * height * rpp / (render_ms + blur_ms) / 1e6, 2)}
        if modes:
            row["tree_render_ms"] = round(time_frames(scene, width, height, rpp, zoom, route=api.ROUTE_TREE)[0], 3)
            row["local_render_ms"] = round(time_frames(scene, width, height, rpp, zoom, route=api.ROUTE_LOCAL_TABLE)[0], 3)
            row["cut_render_ms"] = round(time_frames(scene, width, height, rpp, zoom, route=api.ROUTE_CUT_TABLE)[0], 3)
        total += render_ms + blur_ms
        rows.append(row)
        print(json.dumps(row), flush=True)
    print(json.dumps({"frame": [width, height, rpp], "scenes": len(rows), "total_ms": round(total, 2),
                      "grays_per_s": round(len(rows) * width * height * rpp / total / 1e6, 2)}), flush=True)


if __name__ == "__main__":
    main()
