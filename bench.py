#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native diffusion-curve ray tracer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one frame of the hot path: ray generation + closest-hit traversal + shading + normalisation
(`rdc_render`) followed by the variable-sigma blur (`rdc_gaussian_blur`), on BASELINE.json's configs[1]:
xmls/arch.xml at 1920x1080, 128 rays per pixel, Orzan flag / blur / per-ray jitter as shipped, denoiser
off. The metric is Grays/s = primary rays per second (W*H*rpp / t_frame / 1e9), whole job.

N > 1: the image is dealt out in 8-row strips, round-robin over the ranks (scene + tree replicated). The
frame reaches rank 0 through peer memory: the render kernel of every rank stores its finished pixels straight
into rank 0's frame over NVLink (rdc_render_to_frames on symmetric memory), then one barrier; scenes with blur
store the rendered frame into every rank's buffer, every rank blurs one contiguous band and the blur stores it
into rank 0's frame (raytracingdiffusioncurves_b200/distributed.py: render_frame_peer). Where symmetric memory
cannot be set up, or with RDC_BENCH_NCCL=1, the same split runs on NCCL gather / all-gather (render_frame).
Total work is fixed as N grows ("strong" scaling); every exchange is inside the timed region.

Prints ONE JSON line (rank 0). Nothing here reads /root/reference.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene, width, height, rays per pixel, max trace depth)
    "arch_1080p_128rpp": ("xml:arch.xml", 1920, 1080, 128, 2),            # BASELINE.json configs[1] (headline)
    "arch_512_128rpp": ("xml:arch.xml", 512, 512, 128, 2),                # configs[0]
    "portal_1080p_depth31": ("xml:PortalDemo.xml", 1920, 1080, 128, 31),  # configs[3]
    "ladybug_1080p_128rpp": ("xml:DiffusionCurvePack/lady_bug.xml", 1920, 1080, 128, 2),   # dense bundled scene
    "dolphin_4k_256rpp": ("xml:DiffusionCurvePack/dolphin.xml", 3840, 2160, 256, 2),        # configs[2], largest bundled scene
    "synth100k_8k_512rpp": ("synth:100000:8192", 8192, 8192, 512, 2),     # configs[4]
    "synth100k_2k_64rpp": ("synth:100000:8192", 2048, 2048, 64, 2),       # configs[4] geometry, smaller frame
}
DEFAULT_WORKLOAD = "arch_1080p_128rpp"
XML_DIR = os.path.join(ROOT, "tests", "golden", "xmls")

# algorithmic work per ray, FMA = 2 flops
F_GEN, F_NODE, F_SEG, F_SHADE, F_ACC = 40.0, 30.0, 20.0, 100.0, 10.0

# From the committed `ncu --set full` capture of k_render on the headline workload
# (profiles/r01c_k_render_arch_ncu_summary.txt): DRAM bytes per launch and issue-slot utilisation.
NCU_CAPTURE = {
    # DRAM bytes: ncu pass with the partial sums' lines discarded after use (the shipped default), profiles/r01c_discard_partials.log;
    # before that the same launch wrote 155 MB (profiles/r01c_k_render_arch_ncu_summary.txt, which also holds the issue-slot figure)
    "arch_1080p_128rpp": {"dram_bytes": 3.64e6 + 30.32e6, "issue_active": 0.8130, "source": "profiles/r01c_discard_partials.log, profiles/r01c_k_render_arch_ncu_summary.txt"},
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def workload_zoom(spec, height):
    """zoom = xml image height / output height, without needing the product library."""
    kind, _, rest = spec.partition(":")
    if kind == "synth":
        return float(rest.split(":")[1]) / height
    import re

    with open(os.path.join(XML_DIR, rest), "rb") as fh:
        head = fh.read(4096).decode("utf-8", "replace")
    return float(re.search(r'image_height="(\d+)"', head).group(1)) / height


def scene_source(spec):
    """Returns (kind, payload): ('file', path) or ('text', xml bytes)."""
    kind, _, rest = spec.partition(":")
    if kind == "xml":
        return "file", os.path.join(XML_DIR, rest)
    n, size = rest.split(":")
    from raytracingdiffusioncurves_b200 import api

    return "text", api.synth_xml(int(n), int(size), int(size))


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def workload_config(name, n_gpus):
    """What both arms print under "config": the workload, nothing that varies from run to run."""
    spec, width, height, rpp, depth = WORKLOADS[name]
    return {"workload": name, "scene": spec, "width": width, "height": height, "rays_per_pixel": rpp, "blur": True, "aa": True,
            "orzan": True, "max_trace_depth": depth, "zoom": workload_zoom(spec, height),
            "parallelism": "single GPU" if n_gpus == 1 else f"{RDC_STRIP_ROWS}-row strips dealt round-robin over {n_gpus} GPUs, frame on rank 0",
            "l2": "flushed between timed steps (256 MiB write)"}


RDC_STRIP_ROWS = 8  # include/rdc_b200.h


def host_threads():
    """Host threads the CPU arm may use: the process's affinity mask, NOT OMP_NUM_THREADS (torchrun exports
    OMP_NUM_THREADS=1 to every rank, which would put the CPU arm on one core and inflate every N > 1 ratio)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_run(kind_pref, scene_spec, width, height, rpp, depth, zoom, target_seconds, whole_frame=False):
    """Times the CPU implementation on a bounded band of the workload (or the whole frame). Returns the cpu_baseline dict."""
    from oracle import pyoracle as po

    kind = kind_pref if (kind_pref == "port" or po.Oracle.reference_available()) else "port"
    if kind == "reference" and depth != 2:  # noqa
        kind = "port"  # the reference build is compiled with MAX_TRACE_DEPTH 2
    oracle = po.Oracle(kind)
    skind, payload = scene_source(scene_spec)
    if skind == "text":
        path = "/tmp/rdc_bench_synth.xml"
        with open(path, "wb") as fh:
            fh.write(payload)
    else:
        path = payload
    scene = po.ingest_xml(path, True)
    mid = height // 2

    def run(rows):
        b = max(0, mid - rows // 2)
        p = po.make_params(width, height, rpp, zoom_factor=zoom, max_trace_depth=depth, row_begin=b, row_end=min(height, b + rows))
        t0 = time.perf_counter()
        oracle.render(scene, p, threads=threads)
        return time.perf_counter() - t0, (p.row_end - p.row_begin)

    threads = host_threads()
    if whole_frame:
        t, rows = run(height)
    else:
        t_probe, rows_probe = run(max(threads, 8))
        rows = int(max(rows_probe, min(height, rows_probe * target_seconds / max(t_probe, 1e-6))))
        t, rows = run(rows)
    rays = float(rows) * width * rpp
    return {
        "value": rays / t / 1e9, "unit": "Grays/s", "cores": threads,
        "kind": "reference" if kind == "reference" else "port",
        "sample": f"{rows} centre rows of the {width}x{height}@{rpp} frame ({rays / 1e6:.1f} M rays, {t:.1f} s), render only, "
                  + ("reference DeviceCode.cu compiled for the host (oracle/_ref)" if kind == "reference"
                     else "restated oracle (oracle/oracle_port.cpp)") + ", brute-force closest hit, OpenMP",
        "ms_per_frame_extrapolated": t / rows * height * 1e3,
    }


def run_reference(args):
    """--impl reference: the reference's own CPU-runnable implementation of the path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec, width, height, rpp, depth = WORKLOADS[args.workload]
    zoom = workload_zoom(spec, height)
    # a step is one whole frame while that keeps the run within a few minutes (the headline frame takes ~4 s on 16 threads);
    # for the heavier workloads each step is a band of the frame sized from a probe
    per_step_seconds = max(1.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    probe = cpu_oracle_run("reference", spec, width, height, rpp, depth, zoom, 1.0)
    whole = probe["ms_per_frame_extrapolated"] * 1e-3 * (args.steps + args.warmup) <= 240.0
    results = []
    for i in range(args.warmup + args.steps):
        r = cpu_oracle_run("reference", spec, width, height, rpp, depth, zoom, per_step_seconds, whole_frame=whole)
        if i >= args.warmup:
            results.append(r)
    value = sum(r["value"] for r in results) / len(results)
    base = results[-1]
    base["value"] = value
    line = {
        "impl": "reference", "metric": "Grays/s", "value": value, "unit": "Grays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(width) * height * rpp / (value * 1e9) * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "bundled scene file" if spec.startswith("xml:") else "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "details": {"step": "one whole frame on the host cores" if whole else
                    "a bounded band of the frame on the host cores; ms_per_step is the extrapolated full frame",
                    "threads_from": "os.sched_getaffinity (OMP_NUM_THREADS is ignored: torchrun sets it to 1)"},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "Grays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from raytracingdiffusioncurves_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    if os.environ.get("NCCL_DEBUG"):
        # NCCL logs to stdout unless told otherwise; stdout carries the one JSON line, so its log (rank counts included) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream().cuda_stream

    spec, width, height, rpp, depth = WORKLOADS[args.workload]
    kind, payload = scene_source(spec)
    t_setup = time.perf_counter()
    host = api.HostScene.from_xml_file(payload) if kind == "file" else api.HostScene.from_xml_text(payload)
    scene = api.Scene(host.arrays, None, stream)
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t_setup) * 1e3
    zoom = float(host.arrays.image_height) / height  # the XML frame stays visible vertically

    row_begin, row_end = api.row_band(height, rank, world)
    rows = row_end - row_begin
    flag = torch.zeros((1,), dtype=torch.float32, device=dev)
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)  # > 126 MB L2
    step_box = [0]

    def params_for(step, b=row_begin, e=row_end, **kw):
        return api.default_frame_params(width, height, rpp, zoom_factor=zoom, max_trace_depth=depth, frame=step,
                                        row_begin=b, row_end=e, **kw)

    if world == 1:
        band = torch.empty((height, width, 4), dtype=torch.float32, device=dev)
        sigma_band = torch.empty((height, width), dtype=torch.float32, device=dev)
        scratch = torch.empty((height, width, 4), dtype=torch.float32, device=dev)
        out_frame = torch.empty((height, width, 4), dtype=torch.float32, device=dev)

        def frame_step(step):
            """One frame, inputs resident: render -> blur."""
            flag.zero_()
            p = params_for(step)
            p.max_sigma = flag.data_ptr()
            scene.render(p, band.data_ptr(), sigma_band.data_ptr(), stream)
            api.gaussian_blur(out_frame.data_ptr(), band.data_ptr(), sigma_band.data_ptr(), scratch.data_ptr(), width, height, 0,
                              height, flag.data_ptr(), stream)
            return out_frame
        launches_per_step = 3
    else:
        from raytracingdiffusioncurves_b200 import distributed as rd

        halo = rd.halo_rows(host.max_blur(depth))
        plan = rd.StripPlan(height, width, world, rank, halo)
        bands = rd.FrameBuffers(plan, dev)
        render_strips, blur_rows = api.cuda_callbacks(scene, lambda: params_for(step_box[0], 0, height), stream)
        band, sigma_band = bands.local_image, bands.local_sigma
        # Peer-memory form (render kernel stores into the consumers' frames over NVLink, no NCCL on the data path) unless
        # symmetric memory is unavailable on this box or RDC_BENCH_NCCL=1 asks for the NCCL gather.
        peers, peer_error = None, None
        if os.environ.get("RDC_BENCH_NCCL") != "1":
            try:
                peers = rd.PeerFrameBuffers(plan, dev)
                render_to, blur_rows_peer = api.cuda_peer_callbacks(scene, lambda: params_for(step_box[0], 0, height), stream)
            except Exception as exc:  # noqa: BLE001
                peers, peer_error = None, f"{type(exc).__name__}: {exc}"
        ok = torch.tensor([1 if peers is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # every rank takes the same path
        if int(ok.item()) == 0:
            peers = None
        exchange = "peer memory (symmetric memory, stores from the render/blur kernels)" if peers is not None else \
            "NCCL gather / all-gather" + (f" (peer path unavailable: {peer_error})" if peer_error else "")

        hook_box = [None]  # e2e: rank 0's "the copy-out of the frame before last is done" wait (see render_frame_peer)

        def frame_step(step):
            """One frame over all ranks: render my strips -> (all-gather, local band blur) -> gather to rank 0."""
            step_box[0] = step
            if peers is not None:
                return rd.render_frame_peer(peers, render_to, blur_rows_peer, use_blur=True, before_barrier=hook_box[0])
            return rd.render_frame(bands, render_strips, blur_rows, use_blur=True)
        launches_per_step = 1 + (2 if halo > 0 else 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-timed, inputs resident, L2 flushed between steps ----------------------------
    for s in range(args.warmup):
        frame_step(s)
    sampler = ClockSampler(local_rank)
    sampler.start()  # (NVML start-up takes milliseconds and differs from rank to rank: keep it in front of the barrier)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()  # all ranks enter the timed region together: a rank that came early would count its wait for the others
    wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.zero_()  # evict L2 between timed iterations (untimed)
        starts[s].record()
        frame_step(args.warmup + s)
        ends[s].record()
    barrier()
    wall = time.perf_counter() - wall0
    sampler.stop_flag = True
    sampler.join()
    step_ms = [a.elapsed_time(b) for a, b in zip(starts, ends)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    rays_per_frame = float(width) * height * rpp
    value = rays_per_frame / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: host buffers in and out, every step ---------------------------------------------------
    n_e2e = max(3, min(args.steps, 50))
    host_img = torch.empty((height, width, 4), dtype=torch.float32).pin_memory() if rank == 0 else None
    if world == 1:
        # the public host-buffer call: frame parameters in, pinned image out, every step. The pipelined form
        # enqueues the copy of frame f on the handle's copy stream while frame f+1 renders; two pinned buffers
        # are filled in turn, the clock stops when the last frame is in host memory.
        host_imgs = [host_img, torch.empty((height, width, 4), dtype=torch.float32).pin_memory()]
        for s in range(2):
            scene.render_frame_to_host(params_for(s), True, host_img.data_ptr(), stream)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(n_e2e):
            scene.render_frame_to_host_async(params_for(1000 + s), True, host_imgs[s % 2].data_ptr(), stream)
        scene.frame_wait()
        t_e2e = (time.perf_counter() - t0) / n_e2e
        t0 = time.perf_counter()
        for s in range(n_e2e):
            scene.render_frame_to_host(params_for(2000 + s), True, host_img.data_ptr(), stream)
        t_sync = (time.perf_counter() - t0) / n_e2e
        e2e_api = ("rdc_render_frame_to_host_async + rdc_frame_wait (render + blur + copy to pinned host memory, the copy of one "
                   f"frame overlapping the next frame's rendering); frame-by-frame rdc_render_frame_to_host: {t_sync * 1e3:.3f} ms")
    else:
        if peers is not None:
            # pipelined like the single-GPU call: rank 0 copies frame f to pinned host memory on its own stream while
            # frame f+1 is rendered into the other frame buffer; two pinned buffers in turn; the clock stops when the
            # last frame is in host memory.
            main, copy_stream = torch.cuda.current_stream(), torch.cuda.Stream()
            host_imgs = [host_img, torch.empty((height, width, 4), dtype=torch.float32).pin_memory()] if rank == 0 else None
            rendered = [torch.cuda.Event() for _ in range(2)]
            copied = [torch.cuda.Event() for _ in range(2)]
            count = [0]

            def wait_for_copy_before_last():
                c = count[0]
                if rank == 0 and c >= 1:
                    main.wait_event(copied[(c - 1) % 2])  # frame c-1's copy-out: its buffer is the next frame's target
            hook_box[0] = wait_for_copy_before_last

            def e2e_step(step):
                frame = frame_step(step)
                c = count[0]
                if rank == 0:
                    rendered[c % 2].record(main)
                    copy_stream.wait_event(rendered[c % 2])
                    with torch.cuda.stream(copy_stream):
                        host_imgs[c % 2].copy_(frame, non_blocking=True)
                        copied[c % 2].record(copy_stream)
                count[0] = c + 1

            def e2e_finish():
                torch.cuda.synchronize()
        else:
            def e2e_step(step):
                frame = frame_step(step)
                if rank == 0:
                    host_img.copy_(frame[:height], non_blocking=True)
                torch.cuda.synchronize()

            def e2e_finish():
                pass
        for s in range(2):
            e2e_step(s)
        e2e_finish()
        barrier()
        t0 = time.perf_counter()
        for s in range(n_e2e):
            e2e_step(1000 + s)
        e2e_finish()
        barrier()
        t_local = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
        t_e2e = float(t_local.item())
        e2e_api = ("distributed.render_frame_peer" if peers is not None else "distributed.render_frame") + \
            " (strip render, exchange, band blur, frame on rank 0) + copy of the frame to pinned host memory on rank 0" + \
            (", the copy of one frame overlapping the next frame's rendering" if peers is not None else "")
    e2e = {"value": rays_per_frame / t_e2e / 1e9, "unit": "Grays/s", "ms_per_step": t_e2e * 1e3,
           "h2d_bytes_per_step": ctypes.sizeof(api.FrameParams) * world, "d2h_bytes_per_step": height * width * 16, "steps": n_e2e,
           "api": e2e_api}

    # ---- roofline of the dominant kernel (k_render), measured live ---------------------------------
    roofline = None
    if rank == 0:
        # work per ray from the counting build (one frame)
        stats = torch.zeros((6,), dtype=torch.int64, device=dev)
        def my_share(step):
            q = params_for(step, 0, height)
            if world > 1:
                q.strip_stride, q.strip_offset = world, rank
            return q

        from raytracingdiffusioncurves_b200.distributed import STRIP as rd_strip

        my_rows = sum(min(rd_strip, height - t * rd_strip) for t in range(rank, (height + rd_strip - 1) // rd_strip, world))
        p = my_share(0)
        p.stats = stats.data_ptr()
        scene.render(p, band.data_ptr(), sigma_band.data_ptr(), stream)
        torch.cuda.synchronize()
        traced, nodes, chords, shaded, deferred, gathered = [float(x) for x in stats.cpu().tolist()]
        nodes += gathered
        band_rays = float(my_rows) * width * rpp
        n_node, n_seg, n_hit = nodes / band_rays, chords / band_rays, shaded / band_rays
        f_ray = F_GEN + n_node * F_NODE + n_seg * F_SEG + n_hit * F_SHADE + F_ACC
        # the kernel alone, CUDA events on its stream
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(5, min(args.steps, 30))
        pk = my_share(1)
        pk.max_sigma = flag.data_ptr()
        k0.record()
        for _ in range(reps):
            scene.render(pk, band.data_ptr(), sigma_band.data_ptr(), stream)
        k1.record()
        torch.cuda.synchronize()
        kernel_ms = k0.elapsed_time(k1) / reps
        # FP32 peak: dependent-FFMA microbenchmark on the same box, same moment
        sink = torch.zeros((4,), dtype=torch.float32, device=dev)
        flops = ctypes.c_double()
        api.lib.rdc_microbench_fp32(2048, 2, sink.data_ptr(), ctypes.byref(flops), stream)
        torch.cuda.synchronize()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        api.lib.rdc_microbench_fp32(2048, 10, sink.data_ptr(), ctypes.byref(flops), stream)
        m1.record()
        torch.cuda.synchronize()
        fp32_peak = flops.value * 10 / (m0.elapsed_time(m1) * 1e-3) / 1e12
        achieved = band_rays * f_ray / (kernel_ms * 1e-3) / 1e12
        hbm_peak = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                hbm_peak = json.load(fh).get("hbm_gbs")
        except Exception:
            pass
        alg_bytes = float(my_rows) * width * 20.0
        roofline = {
            "bound": "fp32", "kernel": "k_render", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": achieved / fp32_peak,
            "traffic": NCU_CAPTURE.get(args.workload, {}).get("dram_bytes") if world == 1 else None,
            "issue_slots": NCU_CAPTURE.get(args.workload) if world == 1 else None,
            "peak_source": "dependent-FFMA microbenchmark (rdc_microbench_fp32) run in this process; MEASURED_PEAKS.json has no FP32 figure",
            "kernel_ms": kernel_ms, "flops_per_ray": f_ray,
            "per_ray": {"nodes": n_node, "chords": n_seg, "hits_shaded": n_hit, "rays_traced": traced / band_rays,
                        "deferred_to_tree": deferred / band_rays, "table_query_nodes": gathered / band_rays},
            "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                    "frac": (alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None,
                    "note": "output only (16 B image + 4 B sigma per pixel): the path is not HBM-bound; the partial sums of split work "
                            "units stay in L2 (their lines are discarded once added up, see DESIGN.md 3.1)"},
        }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_oracle_run("reference", spec, width, height, rpp, depth, zoom, args.cpu_seconds)

    if rank == 0:
        st = scene.stats
        line = {
            "metric": "Grays/s", "value": value, "unit": "Grays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "bundled scene file (tests/golden/xmls)" if kind == "file" else "synthetic (rdc_synth_xml, SplitMix64 0x5EEDC0DE)",
            "config": workload_config(args.workload, world),
            "details": {"curves": st.n_curves, "segments": st.n_segments, "chords": st.n_chords, "runs": st.n_runs, "bvh_depth": st.bvh_depth,
                        "exchange": exchange if world > 1 else None, "setup_ms": setup_ms,
                        "wall_ms_per_step_incl_flush": wall / args.steps * 1e3},
            "clocks": sampler.result(),
            "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps * world,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
