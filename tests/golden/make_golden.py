"""Generates the committed golden vectors. Run in the authoring container (needs /root/reference):

    python tests/golden/make_golden.py

* xml_dump_sha256.json — sha256 of the element-tree dump that the reference's rapidxml (parse<0>) produces
  for every bundled scene file (oracle/_ref/ref_xml_dump).
* render_<scene>.npz — image, blur map and per-ray first-hit chord ids produced by the reference's own
  DeviceCode.cu compiled for the host (oracle/_ref/libref_oracle.so), shipped switches
  (Orzan on, AA on, MAX_TRACE_DEPTH 2), seed 0, frame 0.
* ingest_arch.json — the known-answer arrays of SURVEY.md Appendix B.4.
"""
import glob
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

RENDER_CASES = {
    # name: (file, width, height, rays per pixel, zoom)
    "arch": ("arch.xml", 40, 40, 32, 512 / 40),
    "portal": ("PortalDemo.xml", 40, 40, 32, 512 / 40),
    "lady_bug": ("DiffusionCurvePack/lady_bug.xml", 32, 32, 16, 16.0),
    "weight_demo": ("weight_demo.xml", 32, 32, 16, 16.0),
    "drape": ("DiffusionCurvePack/drape.xml", 24, 24, 16, 512 / 24),
}


def main():
    po.build()
    xml_dir = os.path.join(HERE, "xmls")
    files = sorted(glob.glob(os.path.join(xml_dir, "*.xml")) + glob.glob(os.path.join(xml_dir, "*", "*.xml")))
    sums = {os.path.relpath(f, xml_dir): hashlib.sha256(po.ref_xml_dump(f).encode()).hexdigest() for f in files}
    with open(os.path.join(HERE, "xml_dump_sha256.json"), "w") as fh:
        json.dump(sums, fh, indent=1, sort_keys=True)
    ref = po.Oracle("reference")
    for name, (f, w, h, n, zoom) in RENDER_CASES.items():
        scene = po.ingest_xml(os.path.join(xml_dir, f), True)
        p = po.make_params(w, h, n, zoom_factor=zoom)
        image, blur, hits = ref.render(scene, p, want_hits=True)
        np.savez_compressed(os.path.join(HERE, f"render_{name}.npz"), image=image, blur_map=blur, hits=hits,
                            meta=np.array([w, h, n, zoom], np.float64), scene_file=np.array(f))
        print(name, "miss fraction", float((hits == 0xFFFFFFFF).mean()))


if __name__ == "__main__":
    main()
