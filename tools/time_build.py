"""Set-up cost of a scene on the device: upload + chords + runs + tree (rdc_accel_build), per tree builder.
    python tools/time_build.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from raytracingdiffusioncurves_b200 import api  # noqa: E402

XML = os.path.join(ROOT, "tests", "golden", "xmls")


def main():
    stream = torch.cuda.current_stream().cuda_stream
    scenes = [("arch.xml", None), ("DiffusionCurvePack/lady_bug.xml", None), ("DiffusionCurvePack/dolphin.xml", None), ("synth100k", api.synth_xml(100000, 8192, 8192))]
    for name, text in scenes:
        t0 = time.perf_counter()
        host = api.HostScene.from_xml_text(text) if text else api.HostScene.from_xml_file(os.path.join(XML, name))
        ingest = (time.perf_counter() - t0) * 1e3
        for tree in (api.TREE_MORTON, api.TREE_SAH, api.TREE_AUTO):
            times = []
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                scene = api.Scene(host.arrays, api.default_accel_options(tree=tree), stream)
                torch.cuda.synchronize()
                times.append((time.perf_counter() - t0) * 1e3)
                st = scene.stats
                scene.close()
            print(f"{name}: ingest {ingest:.1f} ms; tree {tree}: build {min(times):.1f} ms (first {times[0]:.1f}), runs {st.n_runs}, depth {st.bvh_depth}", flush=True)


if __name__ == "__main__":
    main()
