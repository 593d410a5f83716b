"""render vs blur time per workload on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from raytracingdiffusioncurves_b200 import api

def ev():
    return torch.cuda.Event(enable_timing=True)

for name, rpp_override in [("xml:DiffusionCurvePack/lady_bug.xml", 16), ("xml:DiffusionCurvePack/dolphin.xml", 16), ("xml:DiffusionCurvePack/face.xml", 16), ("xml:DiffusionCurvePack/fille.xml", 16)]:
    for (w, h) in [(1920, 1080), (3840, 2160)]:
        zoom = bench.workload_zoom(name, h)
        host = api.HostScene.from_xml_file(bench.scene_source(name)[1])
        stream = torch.cuda.current_stream().cuda_stream
        scene = api.Scene(host.arrays, None, stream)
        image = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
        out = torch.empty_like(image)
        scratch = torch.empty_like(image)
        sigma = torch.empty((h, w), dtype=torch.float32, device="cuda")
        flag = torch.zeros((1,), dtype=torch.float32, device="cuda")
        p = api.default_frame_params(w, h, rpp_override, zoom_factor=zoom)
        p.max_sigma = flag.data_ptr()
        a, b, c = ev(), ev(), ev()
        for it in range(3):
            a.record()
            scene.render(p, image.data_ptr(), sigma.data_ptr(), stream)
            b.record()
            api.gaussian_blur(out.data_ptr(), image.data_ptr(), sigma.data_ptr(), scratch.data_ptr(), w, h, 0, h, flag.data_ptr(), stream)
            c.record()
            torch.cuda.synchronize()
        nz = float((sigma > 0).float().mean().item())
        print(f"{name.split('/')[-1]} {w}x{h}@{rpp_override}: render {a.elapsed_time(b):.3f} ms, blur {b.elapsed_time(c):.3f} ms, max sigma {flag.item():.2f}, mean sigma {float(torch.nan_to_num(sigma).mean().item()):.2f}, sigma>0 on {nz:.2f} of pixels", flush=True)
