"""Renders a few frames of a bench workload and exits — the short command ncu wraps.
    python tools/profile_frame.py [workload] [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from raytracingdiffusioncurves_b200 import api  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    spec, width, height, rpp, depth = bench.WORKLOADS[workload]
    # RDC_PROFILE_SIZE=WxHxN: same scene and framing, smaller frame (keeps ncu replays short)
    if os.environ.get("RDC_PROFILE_SIZE"):
        width, height, rpp = (int(v) for v in os.environ["RDC_PROFILE_SIZE"].split("x"))
    zoom = bench.workload_zoom(spec, height)
    kind, payload = bench.scene_source(spec)
    host = api.HostScene.from_xml_file(payload) if kind == "file" else api.HostScene.from_xml_text(payload)
    stream = torch.cuda.current_stream().cuda_stream
    run_length = int(os.environ.get("RDC_RUN_LENGTH", "0"))
    scene = api.Scene(host.arrays, api.default_accel_options(run_length=run_length, tree=int(os.environ.get("RDC_TREE", "0"))), stream)
    # RDC_PROFILE_ROWS=a:b: only that band of the frame (full-size frames of the largest workloads take seconds)
    row_begin, row_end = 0, height
    if os.environ.get("RDC_PROFILE_ROWS"):
        row_begin, row_end = (int(v) for v in os.environ["RDC_PROFILE_ROWS"].split(":"))
    rows = row_end - row_begin
    route = int(os.environ.get("RDC_PROFILE_ROUTE", "0"))  # api.ROUTE_*: 0 automatic, 1 tree, 2 local run table
    # RDC_PROFILE_STRIPS=stride:offset: one rank's share of a multi-GPU frame (strips t % stride == offset), on this one GPU
    stride, offset = (int(v) for v in os.environ.get("RDC_PROFILE_STRIPS", "0:0").split(":"))
    units = int(os.environ.get("RDC_PROFILE_UNITS", "0"))  # work units per tile: 0 automatic, 1, 2, 4
    extra = dict(route=route, strip_stride=stride, strip_offset=offset, units_per_tile=units)
    image = torch.empty((rows, width, 4), dtype=torch.float32, device="cuda")
    sigma = torch.empty((rows, width), dtype=torch.float32, device="cuda")
    scratch = torch.empty_like(image)
    flag = torch.zeros((1,), dtype=torch.float32, device="cuda")
    t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    for f in range(frames):
        flag.zero_()
        p = api.default_frame_params(width, height, rpp, zoom_factor=zoom, max_trace_depth=depth, frame=f, row_begin=row_begin,
                                     row_end=row_end, **extra)
        p.max_sigma = flag.data_ptr()
        t0.record()
        scene.render(p, image.data_ptr(), sigma.data_ptr(), stream)
        t1.record()
        api.gaussian_blur(image.data_ptr(), image.data_ptr(), sigma.data_ptr(), scratch.data_ptr(), width, rows, 0, rows,
                          flag.data_ptr(), stream)
        t2.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t2)
        print(f"frame {f}: {ms:.3f} ms (render {t0.elapsed_time(t1):.3f}, blur {t1.elapsed_time(t2):.3f}), "
              f"{rows * width * rpp / ms / 1e6:.2f} Grays/s, chords {scene.stats.n_chords}, max sigma {flag.item():.3f}")
    if os.environ.get("RDC_PROFILE_STATS"):
        stats = torch.zeros((9,), dtype=torch.int64, device="cuda")
        stats[6:8] = torch.iinfo(torch.int64).max
        p = api.default_frame_params(width, height, rpp, zoom_factor=zoom, max_trace_depth=depth, frame=0, row_begin=row_begin,
                                     row_end=row_end, **extra)
        p.stats = stats.data_ptr()
        scene.render(p, image.data_ptr(), sigma.data_ptr(), stream)
        torch.cuda.synchronize()
        rays = float(width) * rows * rpp
        names = ("traced", "boxes", "chords", "shaded", "deferred", "query_nodes")
        host = stats.cpu().tolist()
        print("per primary ray: " + ", ".join(f"{k} {v / rays:.4f}" for k, v in zip(names, host[:6])) + f"; runs {scene.stats.n_runs}")
        print(f"counting build timeline: first warp out of work after {(host[7] - host[6]) * 1e-3:.1f} us, last warp out after "
              f"{(host[8] - host[6]) * 1e-3:.1f} us")


if __name__ == "__main__":
    main()
