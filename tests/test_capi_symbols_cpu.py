"""The C-ABI library loads without a GPU and exports every symbol include/rdc_b200.h declares."""
import ctypes
import os
import re

from raytracingdiffusioncurves_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "rdc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 24 and "gaussianBlur" in names and "rdc_render" in names
    lib = ctypes.CDLL(api.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in rdc_b200.h but not exported"
        assert n in api.PROTOTYPES, f"{n} has no ctypes prototype in api.py"


def test_defaults_equal_the_reference_knobs():
    i = api.default_ingest_options()
    assert (i.use_diffusion_curve_save, i.default_weight_degree, i.endcap_size) == (1, 0.5, 8.0)
    a = api.default_accel_options()
    assert abs(a.curve_width - 1e-3) < 1e-9
    p = api.default_frame_params(640, 480, 128)
    assert (p.zoom_factor, p.offset_x, p.offset_y, p.frame) == (1.0, 0.0, 0.0, 0)
    assert (p.use_diffusion_curve_save, p.use_aa, p.max_trace_depth) == (1, 1, 2)
    assert (p.row_begin, p.row_end) == (0, 480)
    assert b"sm_100a" in api.lib.rdc_version()


def test_params_header_keeps_the_reference_switch_names():
    text = open(os.path.join(ROOT, "include", "params.h")).read()
    for macro in ("USE_DIFFUSION_CURVE_SAVE", "USE_BLUR", "USE_AA", "USE_DENOISER", "MAX_TRACE_DEPTH"):
        assert re.search(rf"#define\s+{macro}\b", text)


def test_null_arguments_fail_without_crashing():
    assert api.lib.rdc_ingest_xml_file(None, None, None) == -1
    assert api.lib.rdc_render(None, None, None, None, None) == -1
    assert "null" in api.last_error()


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "raytracingdiffusioncurves_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cpp", ".cu")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r'#\s*include[^\n]*oracle|^\s*(from|import)\s+[^\n]*oracle|dlopen[^\n]*oracle|CDLL[^\n]*oracle',
                                     text, flags=re.M), f"{f} reaches into oracle/"
