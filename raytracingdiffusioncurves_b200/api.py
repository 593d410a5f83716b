"""ctypes binding of librdc_b200.so (include/rdc_b200.h) — the host-side mirror used by tests and bench.py.

The reference is a C++ executable, so the product's host code is C++ (csrc/, the ``OptixHello`` program);
this module only marshals pointers. Device memory comes from the caller (torch tensors or raw pointers);
nothing here computes anything and there is no CPU path: if the CUDA library is missing, import fails.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RDC_B200_LIB") or os.path.join(_HERE, "librdc_b200.so")  # override: kernel experiments

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
        "raytracingdiffusioncurves_b200 has no CPU fallback."
    )

_lib = C.CDLL(LIB_PATH)

u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)


class IngestOptions(C.Structure):
    _fields_ = [("use_diffusion_curve_save", C.c_int), ("default_weight_degree", C.c_float), ("endcap_size", C.c_float)]


class SceneArrays(C.Structure):
    _fields_ = [
        ("image_width", C.c_int), ("image_height", C.c_int),
        ("n_vertices", C.c_uint32), ("n_segments", C.c_uint32), ("n_curves", C.c_uint32),
        ("vertices", f32p), ("segment_indices", u32p), ("curve_map", u32p), ("curve_index", u32p),
        ("curve_connect", i32p), ("curve_map_inverse", u32p),
        ("n_color_left", C.c_uint32), ("n_color_right", C.c_uint32), ("n_blur", C.c_uint32),
        ("n_weight", C.c_uint32), ("n_weight_degree", C.c_uint32),
        ("color_left_index", u32p), ("color_left", f32p), ("color_left_u", f32p),
        ("color_right_index", u32p), ("color_right", f32p), ("color_right_u", f32p),
        ("blur_index", u32p), ("blur", f32p), ("blur_u", f32p),
        ("weight_index", u32p), ("weight", f32p), ("weight_u", f32p),
        ("weight_degree_index", u32p), ("weight_degree", f32p), ("weight_degree_u", f32p),
    ]


class AccelOptions(C.Structure):
    _fields_ = [("curve_width", C.c_float), ("flatness_tolerance", C.c_float), ("max_chords_per_segment", C.c_int),
                ("run_length", C.c_int), ("shading_records", C.c_int), ("tree", C.c_int)]


class SceneInfo(C.Structure):
    _fields_ = [
        ("n_segments", C.c_uint32), ("n_curves", C.c_uint32), ("n_chords", C.c_uint32), ("n_runs", C.c_uint32),
        ("n_nodes", C.c_uint32), ("bvh_depth", C.c_uint32), ("has_portals", C.c_int), ("device_bytes", C.c_uint64),
        ("traversal_bytes", C.c_uint64), ("pad", C.c_float),
    ]


class FrameParams(C.Structure):
    _fields_ = [
        ("image_width", C.c_uint32), ("image_height", C.c_uint32), ("number_of_rays_per_pixel", C.c_float),
        ("zoom_factor", C.c_float), ("offset_x", C.c_float), ("offset_y", C.c_float),
        ("frame", C.c_uint32), ("seed", C.c_uint32), ("row_begin", C.c_uint32), ("row_end", C.c_uint32),
        ("strip_stride", C.c_uint32), ("strip_offset", C.c_uint32),
        ("use_diffusion_curve_save", C.c_int), ("use_aa", C.c_int), ("max_trace_depth", C.c_int),
        ("traversal", C.c_int), ("hit_ids", C.c_void_p), ("max_sigma", C.c_void_p), ("stats", C.c_void_p),
        ("route", C.c_int), ("units_per_tile", C.c_uint32), ("local_radius", C.c_float),
    ]


TRAVERSAL_LBVH = 0
TRAVERSAL_BRUTE_FORCE = 1
ROUTE_AUTO, ROUTE_TREE, ROUTE_LOCAL_TABLE, ROUTE_CUT_TABLE = 0, 1, 2, 3
TREE_AUTO, TREE_MORTON, TREE_SAH = 0, 1, 2

# every symbol include/rdc_b200.h declares, with its prototype (tests check the library exports them all)
PROTOTYPES = {
    "rdc_default_ingest_options": (None, [C.POINTER(IngestOptions)]),
    "rdc_ingest_xml_file": (C.c_int, [C.c_char_p, C.POINTER(IngestOptions), C.POINTER(C.c_void_p)]),
    "rdc_ingest_xml_memory": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(IngestOptions), C.POINTER(C.c_void_p)]),
    "rdc_host_scene_arrays": (C.c_int, [C.c_void_p, C.POINTER(SceneArrays)]),
    "rdc_host_scene_destroy": (None, [C.c_void_p]),
    "rdc_host_scene_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "rdc_host_scene_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "rdc_xml_dump_file": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "rdc_free": (None, [C.c_void_p]),
    "rdc_default_accel_options": (None, [C.POINTER(AccelOptions)]),
    "rdc_accel_build": (C.c_int, [C.POINTER(SceneArrays), C.POINTER(AccelOptions), C.c_void_p, C.POINTER(C.c_void_p)]),
    "rdc_scene_get_info": (C.c_int, [C.c_void_p, C.POINTER(SceneInfo)]),
    "rdc_scene_download_chords": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rdc_scene_destroy": (None, [C.c_void_p]),
    "rdc_default_frame_params": (None, [C.POINTER(FrameParams), C.c_uint32, C.c_uint32, C.c_float]),
    "rdc_scene_reserve": (C.c_int, [C.c_void_p, C.POINTER(FrameParams), C.c_int, C.c_void_p]),
    "rdc_render": (C.c_int, [C.c_void_p, C.POINTER(FrameParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "rdc_render_to_frames": (C.c_int, [C.c_void_p, C.POINTER(FrameParams), C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                       C.c_void_p]),
    "rdc_peer_frames_create": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "rdc_peer_frames_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rdc_peer_frames_connect_ipc": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rdc_peer_frames_connect_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "rdc_peer_frames_destroy": (None, [C.c_void_p]),
    "rdc_peer_barrier": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rdc_peer_status": (C.c_int, [C.c_void_p]),
    "rdc_peer_render_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(FrameParams), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_void_p)]),
    "rdc_peer_frame_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(FrameParams), C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "rdc_peer_frames_wait": (C.c_int, [C.c_void_p]),
    "rdc_host_frame_open": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "rdc_host_frame_close": (C.c_int, [C.c_char_p, C.c_void_p, C.c_size_t, C.c_int]),
    "rdc_host_scene_halo_rows": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "gaussianBlur": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "setFloatDevice": (None, [C.c_void_p, C.c_uint, C.c_float, C.c_void_p]),
    "setupCurand": (None, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rdc_gaussian_blur": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p]),
    "rdc_gaussian_blur_band": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p, C.c_void_p]),
    "rdc_render_frame_to_host": (C.c_int, [C.c_void_p, C.POINTER(FrameParams), C.c_int, C.c_void_p, C.c_void_p]),
    "rdc_render_frame_to_host_async": (C.c_int, [C.c_void_p, C.POINTER(FrameParams), C.c_int, C.c_void_p, C.c_void_p]),
    "rdc_frame_wait": (C.c_int, [C.c_void_p]),
    "rdc_image_to_rgba8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "rdc_write_ppm": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_int]),
    "rdc_write_png": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_int]),
    "rdc_write_jpg": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "rdc_psnr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rdc_view_scroll": (None, [C.POINTER(FrameParams), C.c_double]),
    "rdc_view_drag": (None, [C.POINTER(FrameParams), C.c_double, C.c_double]),
    "rdc_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]),
    "rdc_synth_xml": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "rdc_microbench_fp32": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    "rdc_microbench_l2": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    "rdc_last_error_string": (C.c_char_p, []),
    "rdc_version": (C.c_char_p, []),
}
for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(_lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args

lib = _lib


class RdcError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = _lib.rdc_last_error_string().decode("utf-8", "replace")
        super().__init__(f"{where} failed with code {code}: {msg}")


def _check(code: int, where: str) -> None:
    if code != 0:
        raise RdcError(code, where)


def last_error() -> str:
    return _lib.rdc_last_error_string().decode("utf-8", "replace")


def default_ingest_options(**overrides) -> IngestOptions:
    o = IngestOptions()
    _lib.rdc_default_ingest_options(C.byref(o))
    for k, v in overrides.items():
        setattr(o, k, v)
    return o


def default_accel_options(**overrides) -> AccelOptions:
    o = AccelOptions()
    _lib.rdc_default_accel_options(C.byref(o))
    for k, v in overrides.items():
        setattr(o, k, v)
    return o


def default_frame_params(width: int, height: int, rays_per_pixel: float, **overrides) -> FrameParams:
    p = FrameParams()
    _lib.rdc_default_frame_params(C.byref(p), width, height, float(rays_per_pixel))
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


_FAMILIES = ("color_left", "color_right", "blur", "weight", "weight_degree")


class HostScene:
    """Result of the ingest: the structure-of-arrays scene in host memory (reference Params arrays)."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)
        self.arrays = SceneArrays()
        _check(_lib.rdc_host_scene_arrays(self._h, C.byref(self.arrays)), "rdc_host_scene_arrays")

    @classmethod
    def from_xml_file(cls, path: str, options: IngestOptions | None = None) -> "HostScene":
        h = C.c_void_p()
        _check(_lib.rdc_ingest_xml_file(os.fsencode(path), C.byref(options) if options else None, C.byref(h)),
               "rdc_ingest_xml_file")
        return cls(h.value)

    @classmethod
    def from_xml_text(cls, text: bytes, options: IngestOptions | None = None) -> "HostScene":
        h = C.c_void_p()
        _check(_lib.rdc_ingest_xml_memory(text, len(text), C.byref(options) if options else None, C.byref(h)),
               "rdc_ingest_xml_memory")
        return cls(h.value)

    @classmethod
    def from_cache(cls, path: str) -> "HostScene":
        h = C.c_void_p()
        _check(_lib.rdc_host_scene_load(os.fsencode(path), C.byref(h)), "rdc_host_scene_load")
        return cls(h.value)

    def save(self, path: str) -> None:
        _check(_lib.rdc_host_scene_save(self._h, os.fsencode(path)), "rdc_host_scene_save")

    def close(self) -> None:
        if self._h:
            _lib.rdc_host_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def halo_rows(self, max_trace_depth: int = 0) -> int:
        """rdc_host_scene_halo_rows: the blur's reach in rows for any frame of this scene (0: no blur)."""
        out = C.c_int()
        _check(_lib.rdc_host_scene_halo_rows(self._h, max_trace_depth, C.byref(out)), "rdc_host_scene_halo_rows")
        return out.value

    def max_blur(self, max_trace_depth: int = 0) -> float:
        """Upper bound of any blur_map value: sigma is a weighted mean of blur stops; each portal passed
        multiplies it by another stop (DeviceCode.cu:311)."""
        a = self.arrays
        if a.n_blur == 0:
            return 0.0
        m = float(np.ctypeslib.as_array(a.blur, shape=(a.n_blur,)).max())
        portals = bool((np.ctypeslib.as_array(a.curve_connect, shape=(a.n_curves,)) >= 0).any())
        if portals and m > 1.0:
            m = m ** (max_trace_depth + 1)
        return max(m, 0.0)

    def to_numpy(self) -> dict:
        """Copies of every array (sentinels included), keyed like struct Params."""
        a = self.arrays

        def arr(ptr, shape, dtype):
            n = int(np.prod(shape))
            if n == 0:
                return np.zeros(shape, dtype)
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True).reshape(shape)

        colour_len = max(a.n_color_left, a.n_color_right) + 2
        out = {
            "image_width": a.image_width, "image_height": a.image_height,
            "vertices": arr(a.vertices, (a.n_vertices, 3), np.float32),
            "segment_indices": arr(a.segment_indices, (a.n_segments,), np.uint32),
            "curve_map": arr(a.curve_map, (a.n_segments,), np.uint32),
            "curve_index": arr(a.curve_index, (a.n_segments,), np.uint32),
            "curve_connect": arr(a.curve_connect, (a.n_curves,), np.int32),
            "curve_map_inverse": arr(a.curve_map_inverse, (a.n_curves,), np.uint32),
        }
        for fam in _FAMILIES:
            n = getattr(a, "n_" + fam)
            padded = colour_len if fam.startswith("color") else n + 2
            stride = 3 if fam.startswith("color") else 1
            out["n_" + fam] = n
            out[fam + "_index"] = arr(getattr(a, fam + "_index"), (a.n_curves, 2), np.uint32)
            out[fam] = arr(getattr(a, fam), (padded, stride) if stride == 3 else (padded,), np.float32)
            out[fam + "_u"] = arr(getattr(a, fam + "_u"), (padded,), np.float32)
        return out


def xml_dump(path: str) -> str:
    p = C.c_void_p()
    _check(_lib.rdc_xml_dump_file(os.fsencode(path), C.byref(p)), "rdc_xml_dump_file")
    try:
        return C.string_at(p).decode("utf-8", "replace")
    finally:
        _lib.rdc_free(p)


def synth_xml(n_curves: int, width: int, height: int, seed: int = 0x5EEDC0DE) -> bytes:
    p = C.c_void_p()
    n = C.c_size_t()
    _check(_lib.rdc_synth_xml(n_curves, width, height, seed, C.byref(p), C.byref(n)), "rdc_synth_xml")
    try:
        return C.string_at(p, n.value)
    finally:
        _lib.rdc_free(p)


@dataclass
class SceneStats:
    n_segments: int
    n_curves: int
    n_chords: int
    n_runs: int
    n_nodes: int
    bvh_depth: int
    has_portals: bool
    device_bytes: int
    traversal_bytes: int
    pad: float


class Scene:
    """Device-resident scene + LBVH (one per device); wraps rdc_accel_build / rdc_render."""

    def __init__(self, arrays: SceneArrays, options: AccelOptions | None = None, stream: int = 0):
        h = C.c_void_p()
        _check(_lib.rdc_accel_build(C.byref(arrays), C.byref(options) if options else None, C.c_void_p(stream), C.byref(h)),
               "rdc_accel_build")
        self._h = h
        info = SceneInfo()
        _check(_lib.rdc_scene_get_info(self._h, C.byref(info)), "rdc_scene_get_info")
        self.stats = SceneStats(info.n_segments, info.n_curves, info.n_chords, info.n_runs, info.n_nodes, info.bvh_depth,
                                bool(info.has_portals), info.device_bytes, info.traversal_bytes, info.pad)

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def close(self) -> None:
        if self._h:
            _lib.rdc_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def chords(self):
        n = self.stats.n_chords
        geom = np.empty((n, 4), np.float32)
        ids = np.empty((n, 3), np.uint32)
        _check(_lib.rdc_scene_download_chords(self._h, geom.ctypes.data, ids.ctypes.data), "rdc_scene_download_chords")
        return geom, ids

    def reserve(self, params: FrameParams, host_frames: bool = False, stream: int = 0) -> None:
        """rdc_scene_reserve: allocate the handle's scratch for frames like `params` ahead of the first render."""
        _check(_lib.rdc_scene_reserve(self._h, C.byref(params), int(host_frames), C.c_void_p(stream)), "rdc_scene_reserve")

    def render(self, params: FrameParams, image_ptr: int, blur_map_ptr: int, stream: int = 0) -> None:
        """Enqueue one frame (rows [row_begin,row_end)) on `stream`. Pointers are device addresses."""
        _check(_lib.rdc_render(self._h, C.byref(params), C.c_void_p(image_ptr), C.c_void_p(blur_map_ptr), C.c_void_p(stream)),
               "rdc_render")

    def render_to_frames(self, params: FrameParams, image_ptrs, blur_map_ptrs, stream: int = 0) -> None:
        """Like render, but every finished pixel is stored at its place in each of the given FULL frames (device
        addresses, peers' memory included): rdc_render_to_frames."""
        n = len(image_ptrs)
        images = (C.c_void_p * n)(*image_ptrs)
        sigmas = (C.c_void_p * n)(*blur_map_ptrs)
        _check(_lib.rdc_render_to_frames(self._h, C.byref(params), n, images, sigmas, C.c_void_p(stream)), "rdc_render_to_frames")

    def render_frame_to_host(self, params: FrameParams, use_blur: bool, host_image_ptr: int, stream: int = 0) -> None:
        _check(_lib.rdc_render_frame_to_host(self._h, C.byref(params), int(use_blur), C.c_void_p(host_image_ptr),
                                             C.c_void_p(stream)), "rdc_render_frame_to_host")


def _scene_async(self, params, use_blur, host_image_ptr, stream=0):
    _check(_lib.rdc_render_frame_to_host_async(self._h, C.byref(params), int(use_blur), C.c_void_p(host_image_ptr), C.c_void_p(stream)),
           "rdc_render_frame_to_host_async")


def _scene_wait(self):
    _check(_lib.rdc_frame_wait(self._h), "rdc_frame_wait")


Scene.render_frame_to_host_async = _scene_async
Scene.frame_wait = _scene_wait


PEER_HANDLE_BYTES = 320  # RDC_PEER_HANDLE_BYTES


class PeerFrames:
    """One rank's view of the frames the GPUs of a box share (rdc_peer_frames_*). Pointer marshalling only."""

    def __init__(self, width: int, height: int, rank: int, world: int):
        h = C.c_void_p()
        _check(_lib.rdc_peer_frames_create(width, height, rank, world, C.byref(h)), "rdc_peer_frames_create")
        self._h, self.rank, self.world, self.width, self.height = h, rank, world, width, height

    def export_handles(self) -> bytes:
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        _check(_lib.rdc_peer_frames_export(self._h, buf), "rdc_peer_frames_export")
        return buf.raw

    def connect_ipc(self, all_handles: bytes) -> None:
        assert len(all_handles) == PEER_HANDLE_BYTES * self.world
        _check(_lib.rdc_peer_frames_connect_ipc(self._h, all_handles), "rdc_peer_frames_connect_ipc")

    @staticmethod
    def connect_local(frames) -> None:
        arr = (C.c_void_p * len(frames))(*[f._h for f in frames])
        _check(_lib.rdc_peer_frames_connect_local(arr, len(frames)), "rdc_peer_frames_connect_local")

    def barrier(self, stream: int = 0) -> None:
        _check(_lib.rdc_peer_barrier(self._h, C.c_void_p(stream)), "rdc_peer_barrier")

    def status(self) -> None:
        _check(_lib.rdc_peer_status(self._h), "rdc_peer_status")

    def render_frame(self, scene: "Scene", params: FrameParams, use_blur: bool, halo_rows: int, stream: int = 0,
                     wait_event: int = 0) -> int:
        """Device consumer: returns the device address of the finished frame on rank 0, 0 elsewhere."""
        out = C.c_void_p()
        _check(_lib.rdc_peer_render_frame(scene.handle, self._h, C.byref(params), int(use_blur), halo_rows, C.c_void_p(wait_event),
                                          C.c_void_p(stream), C.byref(out)), "rdc_peer_render_frame")
        return out.value or 0

    def frame_to_host(self, scene: "Scene", params: FrameParams, use_blur: bool, halo_rows: int, host_frame_ptr: int,
                      stream: int = 0) -> None:
        _check(_lib.rdc_peer_frame_to_host(scene.handle, self._h, C.byref(params), int(use_blur), halo_rows,
                                           C.c_void_p(host_frame_ptr), C.c_void_p(stream)), "rdc_peer_frame_to_host")

    def wait(self) -> None:
        _check(_lib.rdc_peer_frames_wait(self._h), "rdc_peer_frames_wait")

    def close(self) -> None:
        if self._h:
            _lib.rdc_peer_frames_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostFrame:
    """A host frame every rank of a box can address: POSIX shared memory, pinned (rdc_host_frame_open)."""

    def __init__(self, name: str, nbytes: int, create: bool):
        p = C.c_void_p()
        _check(_lib.rdc_host_frame_open(name.encode(), nbytes, int(create), C.byref(p)), "rdc_host_frame_open")
        self.name, self.nbytes, self.ptr, self.owner = name, nbytes, p.value, create

    def numpy(self, shape):
        return np.ctypeslib.as_array(C.cast(self.ptr, f32p), shape=(self.nbytes // 4,)).reshape(shape)

    def close(self) -> None:
        if self.ptr:
            _lib.rdc_host_frame_close(self.name.encode(), C.c_void_p(self.ptr), self.nbytes, int(self.owner))
            self.ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gaussian_blur(dest_ptr: int, src_ptr: int, sigma_ptr: int, scratch_ptr: int, width: int, height: int,
                  row_begin: int = 0, row_end: int | None = None, max_sigma_ptr: int = 0, stream: int = 0) -> None:
    _check(_lib.rdc_gaussian_blur(C.c_void_p(dest_ptr), C.c_void_p(src_ptr), C.c_void_p(sigma_ptr), C.c_void_p(scratch_ptr),
                                  width, height, row_begin, height if row_end is None else row_end,
                                  C.c_void_p(max_sigma_ptr), C.c_void_p(stream)), "rdc_gaussian_blur")


def gaussian_blur_band(dest_ptr: int, src_ptr: int, sigma_ptr: int, scratch_ptr: int, width: int, height: int,
                       row_begin: int, row_end: int, halo_rows: int, max_sigma_ptr: int = 0, stream: int = 0) -> None:
    _check(_lib.rdc_gaussian_blur_band(C.c_void_p(dest_ptr), C.c_void_p(src_ptr), C.c_void_p(sigma_ptr), C.c_void_p(scratch_ptr),
                                       width, height, row_begin, row_end, halo_rows, C.c_void_p(max_sigma_ptr),
                                       C.c_void_p(stream)), "rdc_gaussian_blur_band")


def image_to_rgba8(image: np.ndarray, flip: bool) -> np.ndarray:
    h, w, _ = image.shape
    img = np.ascontiguousarray(image, np.float32)
    out = np.empty((h, w, 4), np.uint8)
    _check(_lib.rdc_image_to_rgba8(img.ctypes.data, w, h, int(flip), out.ctypes.data), "rdc_image_to_rgba8")
    return out


def psnr(a: np.ndarray, b: np.ndarray) -> tuple[float, float]:
    """rdc_psnr on two [H, W, 4] float images: (PSNR in dB, largest absolute RGB difference)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape and a.shape[-1] == 4
    p, m = C.c_double(), C.c_double()
    _check(_lib.rdc_psnr(a.ctypes.data, b.ctypes.data, a.size // 4, C.byref(p), C.byref(m)), "rdc_psnr")
    return p.value, m.value


def row_band(height: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [begin,end) of rank `rank` of `world`: contiguous bands, remainder spread over the first ranks."""
    base, rem = divmod(height, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def cuda_callbacks(scene: "Scene", make_params, stream: int = 0):
    """(render_strips, blur_rows) for distributed.render_frame, bound to the CUDA entry points.
    make_params() -> FrameParams of the frame being rendered (whole image; the strip fields are set here)."""

    def render_strips(image, sigma, stride, offset):
        p = make_params()
        p.strip_stride, p.strip_offset = (stride, offset) if stride > 1 else (0, 0)
        scene.render(p, image.data_ptr(), sigma.data_ptr(), stream)

    def blur_rows(dest, source, sigma, scratch, height, row_begin, row_end, halo):
        dest_ptr = dest if isinstance(dest, int) else dest.data_ptr()  # an address: a frame in a peer's memory
        gaussian_blur_band(dest_ptr, source.data_ptr(), sigma.data_ptr(), scratch.data_ptr(), source.shape[1], height,
                           row_begin, row_end, halo, 0, stream)

    return render_strips, blur_rows
