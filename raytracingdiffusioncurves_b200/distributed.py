"""Partitioning of one frame over the GPUs of a box.

One process per GPU (torch.distributed: NCCL on GPUs, gloo in the CPU tests). The curve set and its tree are
replicated. Random numbers are keyed by the GLOBAL pixel index, so the assembled frame has the single-GPU frame's hits
whatever the split — and its pixels bit for bit when rdc_frame_params::units_per_tile is pinned.

Rendering is dealt out in STRIPS of 8 rows, round-robin: rank r renders strips t with t % world == r. A
contiguous band per rank (the obvious split) is badly balanced on sparse scenes — in arch.xml the rows that
look into the arch cost several times the rows that look away from it — while interleaved strips sample
every region of the image on every rank.

The data path has at most two exchange steps:

  no blur in the scene (largest blur stop 0):
      gather      every rank's packed strips -> rank 0, which puts the rows in order (one indexed copy)
  blur:
      all-gather  packed strips (image + sigma) -> every rank has the whole rendered frame;
                  each rank blurs ONE contiguous band of it locally (horizontal pass on the band plus the
                  ceil(3*sigma_max) rows the vertical pass can reach, helperKernels.cu:65,74);
      gather      blurred bands -> rank 0

On GPUs the product path is the peer-memory form of the same plan, and it is NOT in this module: it is C++ behind
the C ABI (csrc/peer.cu — rdc_peer_frames_*, rdc_peer_render_frame, rdc_peer_frame_to_host; api.PeerFrames marshals the
pointers). Every rank holds the device address of every other rank's frame buffers (CUDA IPC or peer access), the
render kernel stores each finished pixel straight into its place in the consumer's frame — rank 0's when the scene has
no blur, everybody's when it has (the all-gather) — and the band blur stores straight into rank 0's finished frame. The
stores ARE the collectives; what is left is one barrier kernel per exchange step. `render_frame_peer` below is the
executable statement of that protocol (which buffer, which barrier, in which order) that the gloo CPU tests run with
shared host memory standing in for peer memory; `render_frame` is the collective form (NCCL / gloo) of the same plan.

What renders and what blurs is injected (`render_strips`, `blur_rows`): the product passes the CUDA entry
points (api.cuda_callbacks), the CPU tests pass the oracle. This module only moves rows.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

STRIP = 8  # RDC_STRIP_ROWS (include/rdc_b200.h): the render kernel's tile height


def row_band(height: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [begin,end) of `rank`: contiguous bands, the remainder spread over the first ranks."""
    base, rem = divmod(height, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def halo_rows(sigma_max: float) -> int:
    """Reach of the blur taps, helperKernels.cu:65,74: k in [-ceil(3 sigma), +ceil(3 sigma)]."""
    if not (sigma_max > 0.0):
        return 0
    return int(math.ceil(3.0 * sigma_max))


class StripPlan:
    """Who renders which rows, and where a row sits in the packed per-rank buffers."""

    def __init__(self, height: int, width: int, world: int, rank: int, halo: int):
        self.height, self.width, self.world, self.rank, self.halo = height, width, world, rank, halo
        self.n_strips = (height + STRIP - 1) // STRIP
        self.local_strips = (self.n_strips - rank + world - 1) // world if rank < self.n_strips else 0
        self.packed_rows = ((self.n_strips + world - 1) // world) * STRIP  # same on every rank (gather needs equal sizes)
        self.band = row_band(height, rank, world)
        self.max_band_rows = row_band(height, 0, world)[1]

    def strips_of(self, rank: int):
        return range(rank, self.n_strips, self.world)

    def source_index(self) -> torch.Tensor:
        """For every row of the frame: its row in the concatenation of all ranks' packed buffers."""
        y = torch.arange(self.height, dtype=torch.int64)
        t = y // STRIP
        return (t % self.world) * self.packed_rows + (t // self.world) * STRIP + y % STRIP

    def band_index(self) -> torch.Tensor:
        """For every row of the frame: its row in the concatenation of all ranks' padded blurred bands."""
        idx = torch.empty(self.height, dtype=torch.int64)
        for r in range(self.world):
            b, e = row_band(self.height, r, self.world)
            idx[b:e] = r * self.max_band_rows + torch.arange(e - b)
        return idx


class FrameBuffers:
    """Per-rank buffers for StripPlan. Tensors live on `device`."""

    def __init__(self, plan: StripPlan, device):
        p = self.plan = plan
        f32 = dict(dtype=torch.float32, device=device)
        multi = p.world > 1
        blur = p.halo > 0
        rows = p.packed_rows if multi else p.height
        self.local_image = torch.zeros((rows, p.width, 4), **f32)
        self.local_sigma = torch.zeros((rows, p.width), **f32)
        self.frame = None
        if not multi:
            if blur:
                self.scratch = torch.zeros((p.height, p.width, 4), **f32)
                self.blurred = torch.zeros((p.height, p.width, 4), **f32)
            return
        self.source_index = p.source_index().to(device)
        if blur:
            self.all_image = torch.zeros((p.world * p.packed_rows, p.width, 4), **f32)
            self.all_sigma = torch.zeros((p.world * p.packed_rows, p.width), **f32)
            self.full_image = torch.zeros((p.height, p.width, 4), **f32)
            self.full_sigma = torch.zeros((p.height, p.width), **f32)
            self.scratch = torch.zeros((p.height, p.width, 4), **f32)
            self.blurred = torch.zeros((p.height, p.width, 4), **f32)
            self.band_pad = torch.zeros((p.max_band_rows, p.width, 4), **f32)
            self.band_index = p.band_index().to(device)
            if p.rank == 0:
                self.all_bands = torch.zeros((p.world * p.max_band_rows, p.width, 4), **f32)
                self.frame = torch.zeros((p.height, p.width, 4), **f32)
        elif p.rank == 0:
            self.all_image = torch.zeros((p.world * p.packed_rows, p.width, 4), **f32)
            self.frame = torch.zeros((p.height, p.width, 4), **f32)


def render_frame(buf: FrameBuffers, render_strips, blur_rows, use_blur: bool = True):
    """One frame over all ranks. Returns the finished frame [H, W, 4] on rank 0, None elsewhere.

    render_strips(image, sigma, stride, offset): renders strips t % stride == offset of the full frame,
        packed one after the other into `image` / `sigma` (rdc_render with strip_stride / strip_offset).
    blur_rows(dest, source, sigma, scratch, height, row_begin, row_end, halo): blurs rows
        [row_begin,row_end) of a `height`-row frame into the same rows of dest (rdc_gaussian_blur_band).
    """
    p = buf.plan
    blur = use_blur and p.halo > 0
    render_strips(buf.local_image, buf.local_sigma, p.world, p.rank)
    if p.world == 1:
        if not blur:
            return buf.local_image
        blur_rows(buf.blurred, buf.local_image, buf.local_sigma, buf.scratch, p.height, 0, p.height, p.halo)
        return buf.blurred
    if not blur:
        parts = list(buf.all_image.view(p.world, p.packed_rows, p.width, 4).unbind(0)) if p.rank == 0 else None
        dist.gather(buf.local_image, parts, dst=0)
        if p.rank != 0:
            return None
        torch.index_select(buf.all_image, 0, buf.source_index, out=buf.frame)
        return buf.frame
    dist.all_gather_into_tensor(buf.all_image, buf.local_image)
    dist.all_gather_into_tensor(buf.all_sigma, buf.local_sigma)
    torch.index_select(buf.all_image, 0, buf.source_index, out=buf.full_image)
    torch.index_select(buf.all_sigma, 0, buf.source_index, out=buf.full_sigma)
    b, e = p.band
    blur_rows(buf.blurred, buf.full_image, buf.full_sigma, buf.scratch, p.height, b, e, p.halo)
    buf.band_pad[: e - b].copy_(buf.blurred[b:e])
    parts = list(buf.all_bands.view(p.world, p.max_band_rows, p.width, 4).unbind(0)) if p.rank == 0 else None
    dist.gather(buf.band_pad, parts, dst=0)
    if p.rank != 0:
        return None
    torch.index_select(buf.all_bands, 0, buf.band_index, out=buf.frame)
    return buf.frame


def render_frame_peer(buf, render_to, blur_rows, use_blur: bool = True, before_barrier=None):
    """One frame over all ranks through peer memory. Returns the finished frame [H, W, 4] on rank 0, None elsewhere.

    render_to(image_ptrs, sigma_ptrs, stride, offset): renders strips t % stride == offset and stores every pixel
        at its place in each of the full frames addressed (rdc_render_to_frames).
    blur_rows: as in render_frame; `dest` is a device address here (rank 0's frame).
    before_barrier(): called on every rank right before the frame's first barrier is enqueued — rank 0 makes its
        stream wait there for whatever still reads the frame buffer the NEXT call will be written into (the frame
        this call returned two calls ago); the peers only start that next frame after this barrier.

    Buffer reuse is safe without an opening barrier: the frame of call c goes to frames[c % 2], whose last reader
    (call c-2's consumer on rank 0) is ordered before call c-1's barrier by before_barrier; full_image / full_sigma
    are read by the band blur of a call, which every rank finishes before that call's closing barrier."""
    p = buf.plan
    blur = use_blur and p.halo > 0
    turn = buf.turn
    buf.turn ^= 1
    out_ptr = buf.frame_ptrs[turn][0]
    if not blur:
        render_to([out_ptr], [buf.sigma_ptrs[0]], p.world, p.rank)  # the gather
        if before_barrier is not None:
            before_barrier()
        buf.barrier()
        return buf.frames[turn] if p.rank == 0 else None
    render_to(buf.image_ptrs, buf.sigma_ptrs, p.world, p.rank)  # the all-gather
    if before_barrier is not None:
        before_barrier()
    buf.barrier()
    b, e = p.band
    blur_rows(out_ptr, buf.full_image, buf.full_sigma, buf.scratch, p.height, b, e, p.halo)  # the gather
    buf.barrier()
    return buf.frames[turn] if p.rank == 0 else None
