"""Row-band partitioning of one frame over the GPUs of a box.

One process per GPU (torch.distributed: NCCL on GPUs, gloo in the CPU tests). The curve set and its tree are
replicated; rank r renders rows [b_r, e_r) with random numbers keyed by the GLOBAL pixel index, so the union
of the bands is bit-identical to the single-GPU frame. The data path has two exchange steps and nothing
else:

  1. blur halo  — the vertical blur pass of a band reads up to ceil(3*sigma_max) rows beyond it. Each rank
     sends that many of its rendered top/bottom rows (image + sigma) to its neighbours (point-to-point),
     then blurs its own band locally. Skipped entirely when the scene has no blur (sigma_max == 0).
  2. gather     — the finished bands go to rank 0.

What renders and what blurs is injected (`render_band`, `blur_rows`): the product passes the CUDA entry
points (see bench.py), the CPU tests pass the oracle. This module only moves rows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.distributed as dist


def row_band(height: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [begin,end) of `rank`: contiguous bands, the remainder spread over the first ranks."""
    base, rem = divmod(height, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


@dataclass
class BandPlan:
    height: int
    width: int
    world: int
    rank: int
    halo: int          # rows the vertical blur may reach beyond a band: ceil(3*sigma_max)
    begin: int = 0
    end: int = 0
    top: int = 0       # halo rows actually held above / below the band (0 at the image border)
    bottom: int = 0
    max_rows: int = 0  # largest band (gather padding)
    exchange: bool = False  # halo exchange + local blur; False = blur the gathered frame on rank 0

    def __post_init__(self):
        self.begin, self.end = row_band(self.height, self.rank, self.world)
        self.max_rows = row_band(self.height, 0, self.world)[1]
        min_rows = self.height // self.world
        # neighbours only: a halo deeper than the smallest band would need rows from two ranks away
        self.exchange = self.world > 1 and 0 < self.halo <= min_rows
        if self.exchange:
            self.top = min(self.halo, self.begin)
            self.bottom = min(self.halo, self.height - self.end)

    @property
    def rows(self) -> int:
        return self.end - self.begin

    @property
    def buffer_rows(self) -> int:
        return self.top + self.rows + self.bottom


def halo_rows(sigma_max: float) -> int:
    """Reach of the blur taps, helperKernels.cu:65,74: k in [-ceil(3 sigma), +ceil(3 sigma)]."""
    if not (sigma_max > 0.0):
        return 0
    return int(math.ceil(3.0 * sigma_max))


class FrameBands:
    """Buffers + the two exchange steps for one rank. Tensors live on `device`."""

    def __init__(self, plan: BandPlan, device, pin_result: bool = False):
        self.plan = plan
        p = plan
        self.image = torch.zeros((p.buffer_rows, p.width, 4), dtype=torch.float32, device=device)
        self.sigma = torch.zeros((p.buffer_rows, p.width), dtype=torch.float32, device=device)
        self.blurred = torch.zeros((p.buffer_rows, p.width, 4), dtype=torch.float32, device=device)
        self.scratch = torch.zeros((p.buffer_rows, p.width, 4), dtype=torch.float32, device=device)
        self.send_pad = torch.zeros((p.max_rows, p.width, 4), dtype=torch.float32, device=device)
        self.frame = None
        self.frame_sigma = None
        if p.rank == 0:
            self.frame = torch.zeros((p.world * p.max_rows, p.width, 4), dtype=torch.float32, device=device)
            if p.world > 1 and not p.exchange:
                self.frame_sigma = torch.zeros((p.world * p.max_rows, p.width), dtype=torch.float32, device=device)
                self.sigma_pad = torch.zeros((p.max_rows, p.width), dtype=torch.float32, device=device)
                self.frame_scratch = torch.zeros((p.height, p.width, 4), dtype=torch.float32, device=device)
                self.frame_out = torch.zeros((p.height, p.width, 4), dtype=torch.float32, device=device)
        elif p.world > 1 and not p.exchange:
            self.sigma_pad = torch.zeros((p.max_rows, p.width), dtype=torch.float32, device=device)

    # views -------------------------------------------------------------------------------------------
    def own(self, t):
        p = self.plan
        return t[p.top:p.top + p.rows]

    # step 1 ------------------------------------------------------------------------------------------
    def exchange_halos(self):
        """Send my rendered border rows to the neighbours, receive theirs into my halo rows."""
        p = self.plan
        if not p.exchange:
            return
        ops = []
        own_img, own_sig = self.own(self.image), self.own(self.sigma)
        if p.rank > 0:  # neighbour above: it needs my first rows as its bottom halo; I need its last rows
            up_rows = min(p.halo, p.rows)
            ops += [dist.P2POp(dist.isend, own_img[:up_rows].contiguous(), p.rank - 1),
                    dist.P2POp(dist.isend, own_sig[:up_rows].contiguous(), p.rank - 1),
                    dist.P2POp(dist.irecv, self.image[:p.top], p.rank - 1),
                    dist.P2POp(dist.irecv, self.sigma[:p.top], p.rank - 1)]
        if p.rank < p.world - 1:
            down_rows = min(p.halo, p.rows)
            ops += [dist.P2POp(dist.isend, own_img[p.rows - down_rows:].contiguous(), p.rank + 1),
                    dist.P2POp(dist.isend, own_sig[p.rows - down_rows:].contiguous(), p.rank + 1),
                    dist.P2POp(dist.irecv, self.image[p.top + p.rows:], p.rank + 1),
                    dist.P2POp(dist.irecv, self.sigma[p.top + p.rows:], p.rank + 1)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    # step 2 ------------------------------------------------------------------------------------------
    def gather(self, band_rows):
        """band_rows: this rank's finished rows [rows, W, 4]. Returns the whole frame on rank 0, else None."""
        p = self.plan
        if p.world == 1:
            self.frame[:p.rows].copy_(band_rows)
            return self.frame[:p.height]
        self.send_pad[:p.rows].copy_(band_rows)
        parts = list(self.frame.view(p.world, p.max_rows, p.width, 4).unbind(0)) if p.rank == 0 else None
        dist.gather(self.send_pad, parts, dst=0)
        if p.rank != 0:
            return None
        if p.height % p.world == 0:
            return self.frame
        return torch.cat([self.frame.view(p.world, p.max_rows, p.width, 4)[r, : row_band(p.height, r, p.world)[1] - row_band(p.height, r, p.world)[0]]
                          for r in range(p.world)])

    def gather_sigma(self):
        p = self.plan
        self.sigma_pad[:p.rows].copy_(self.own(self.sigma))
        parts = list(self.frame_sigma.view(p.world, p.max_rows, p.width).unbind(0)) if p.rank == 0 else None
        dist.gather(self.sigma_pad, parts, dst=0)
        if p.rank != 0:
            return None
        if p.height % p.world == 0:
            return self.frame_sigma
        return torch.cat([self.frame_sigma.view(p.world, p.max_rows, p.width)[r, : row_band(p.height, r, p.world)[1] - row_band(p.height, r, p.world)[0]]
                          for r in range(p.world)])


def render_frame(bands: FrameBands, render_band, blur_rows, use_blur: bool = True):
    """One frame over all ranks. Returns the finished frame [H, W, 4] on rank 0, None elsewhere.

    render_band(image_rows, sigma_rows, row_begin, row_end): fills the band's rows (global row numbers).
    blur_rows(dest, source, sigma, scratch, height, row_begin, row_end): blurs rows [row_begin,row_end) of
        a buffer of `height` rows, clamping at the buffer's edges (rdc_gaussian_blur's contract).
    """
    p = bands.plan
    render_band(bands.own(bands.image), bands.own(bands.sigma), p.begin, p.end)
    if not use_blur or p.halo == 0:
        return bands.gather(bands.own(bands.image))
    if p.world == 1 or p.exchange:
        bands.exchange_halos()
        blur_rows(bands.blurred, bands.image, bands.sigma, bands.scratch, p.buffer_rows, p.top, p.top + p.rows)
        return bands.gather(bands.own(bands.blurred))
    # halo deeper than a band: assemble the rendered frame on rank 0 and blur it there
    frame = bands.gather(bands.own(bands.image))
    sigma = bands.gather_sigma()
    if p.rank != 0:
        return None
    frame = frame.contiguous()
    blur_rows(bands.frame_out, frame, sigma.contiguous(), bands.frame_scratch, p.height, 0, p.height)
    return bands.frame_out
