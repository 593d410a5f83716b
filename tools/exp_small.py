import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, time
import bench
from raytracingdiffusioncurves_b200 import api
spec, w, h, rpp, depth = bench.WORKLOADS["arch_1080p_128rpp"]
zoom = bench.workload_zoom(spec, h)
host = api.HostScene.from_xml_file(bench.scene_source(spec)[1])
stream = torch.cuda.current_stream().cuda_stream
scene = api.Scene(host.arrays, None, stream)
image = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
sigma = torch.empty((h, w), dtype=torch.float32, device="cuda")
cases = {"band16": dict(row_begin=528, row_end=544), "away": dict(offset_x=1e5), "full": {}, "band16_top": dict(row_begin=0, row_end=16)}
for name, kw in cases.items():
    p = api.default_frame_params(w, h, rpp, zoom_factor=zoom, **kw)
    for _ in range(3):
        scene.render(p, image.data_ptr(), sigma.data_ptr(), stream)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = time.perf_counter()
    t0.record()
    for _ in range(10):
        scene.render(p, image.data_ptr(), sigma.data_ptr(), stream)
    t1.record()
    c1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name}: gpu {t0.elapsed_time(t1)/10:.4f} ms per launch, host issue {(c1-c0)/10*1e3:.4f} ms per launch", flush=True)
