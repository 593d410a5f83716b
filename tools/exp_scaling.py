"""Kernel-time experiments on one GPU: time of k_render vs share of the frame (strip stride), and the
cost of pixels whose rays are all culled."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from raytracingdiffusioncurves_b200 import api

def timeit(scene, p, image, sigma, reps=20):
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        scene.render(p, image.data_ptr(), sigma.data_ptr(), stream)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        scene.render(p, image.data_ptr(), sigma.data_ptr(), stream)
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps

def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "arch_1080p_128rpp"
    spec, w, h, rpp, depth = bench.WORKLOADS[wl]
    zoom = bench.workload_zoom(spec, h)
    kind, payload = bench.scene_source(spec)
    host = api.HostScene.from_xml_file(payload) if kind == "file" else api.HostScene.from_xml_text(payload)
    scene = api.Scene(host.arrays, None, torch.cuda.current_stream().cuda_stream)
    image = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
    sigma = torch.empty((h, w), dtype=torch.float32, device="cuda")
    for stride in (1, 2, 4, 8, 16):
        for off in (0, stride - 1):
            p = api.default_frame_params(w, h, rpp, zoom_factor=zoom, max_trace_depth=depth, strip_stride=stride, strip_offset=off)
            print(f"{wl} stride {stride} offset {off}: {timeit(scene, p, image, sigma):.4f} ms", flush=True)
    for rows in (1080, 540, 270, 135, 64, 16):
        if rows > h: continue
        b = (h - rows) // 2 // 16 * 16
        p = api.default_frame_params(w, h, rpp, zoom_factor=zoom, max_trace_depth=depth, row_begin=b, row_end=b + rows)
        print(f"{wl} centre band of {rows} rows: {timeit(scene, p, image, sigma):.4f} ms", flush=True)
    p = api.default_frame_params(w, h, rpp, zoom_factor=zoom, max_trace_depth=depth, offset_x=1e5)
    print(f"{wl} looking away (every ray culled): {timeit(scene, p, image, sigma):.4f} ms", flush=True)
    p = api.default_frame_params(w, h, 8, zoom_factor=zoom, max_trace_depth=depth)
    print(f"{wl} 8 rays per pixel: {timeit(scene, p, image, sigma):.4f} ms", flush=True)

main()
