"""Independent check of the intersector's geometric claim (rdc_math.h "Chord subdivision").

Every parity test compares the product with an oracle that shares rdc_math.h, so an error in the chord
subdivision or in rdc_ray_chord would be invisible to them. Here the chord hit (chord id, t, u) the shared
arithmetic produces is compared with the curve itself — the uniform cubic B-spline of DeviceCode.cu:64-75 —
evaluated in float64 by this file's own basis functions:

 * position: the curve point at the reported parameter u lies within the flatness tolerance (0.05 XML pixels)
   of the reported hit point O + t D;
 * root: a double-precision Newton iteration on cross(D, B(u) - O) = 0 started at u converges to a crossing u*
   with |u - u*| |B'(u*)| and |t - t*| of the order tolerance / |sin(angle between ray and curve)|;
 * closest hit: brute force over a 4x finer float64 polyline of every segment finds the same segment at the
   same distance (up to that bound), and agrees on which rays miss.
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

TOL = 0.05  # rdc_accel_options::flatness_tolerance default
MISS = 0xFFFFFFFF


def basis(u):
    u = np.asarray(u, np.float64)
    return np.stack([(1 - u) ** 3, 3 * u ** 3 - 6 * u ** 2 + 4, -3 * u ** 3 + 3 * u ** 2 + 3 * u + 1, u ** 3], -1) / 6.0


def dbasis(u):
    u = np.asarray(u, np.float64)
    return np.stack([-3 * (1 - u) ** 2, 9 * u ** 2 - 12 * u, -9 * u ** 2 + 6 * u + 3, 3 * u ** 2], -1) / 6.0


def control_points(scene, seg):
    first = scene["segment_indices"][seg].astype(np.int64)
    idx = first[:, None] + np.arange(4)[None, :]
    return scene["vertices"][idx][..., :2].astype(np.float64)  # [n, 4, 2]


def random_rays(geom, n, seed):
    """Origins in and a little around the box of the chords, directions uniform."""
    rng = np.random.default_rng(seed)
    v = geom.reshape(-1, 2)
    lo, hi = v.min(0), v.max(0)
    m = 0.1 * (hi - lo)
    o = rng.uniform(lo - m, hi + m, size=(n, 2))
    a = rng.uniform(0, 2 * np.pi, size=n)
    return np.concatenate([o, np.cos(a)[:, None], np.sin(a)[:, None]], 1).astype(np.float32)


SCENES = ["arch.xml", "PortalDemo.xml", "weight_demo.xml", "DiffusionCurvePack/lady_bug.xml", "DiffusionCurvePack/dolphin.xml",
          "DiffusionCurvePack/roses_spirales.xml", "DiffusionCurvePack/face.xml"]


@pytest.mark.parametrize("name", SCENES)
def test_chord_hit_lies_on_the_curve(name, xml_dir, port_oracle):
    scene = po.ingest_xml(os.path.join(xml_dir, name), True)
    geom, cids = port_oracle.chords(scene)
    rays = random_rays(geom, 100_000, 7)
    ids, t, u = port_oracle.trace_rays(scene, rays, search="grid")
    hit = ids != MISS
    assert hit.sum() > 20_000
    r64 = rays[hit].astype(np.float64)
    O, D = r64[:, :2], r64[:, 2:]
    t64, u64 = t[hit].astype(np.float64), u[hit].astype(np.float64)
    seg = cids[ids[hit], 0]
    cp = control_points(scene, seg)
    p_ray = O + t64[:, None] * D
    scale = 1.0 + np.abs(p_ray).max(1) + t64
    # (1) position: curve point at the reported parameter vs reported hit point
    p_curve = np.einsum("nk,nkc->nc", basis(u64), cp)
    dev = np.linalg.norm(p_curve - p_ray, axis=1)
    assert np.all(dev <= TOL + 4e-6 * scale), f"max |B(u) - (O + tD)| = {dev.max():.4f}"
    # (2) a true crossing next to the chord hit: cross(D, B(.) - O) changes sign inside the window the error bound
    #     allows around u (intermediate value theorem), and the crossing found there by bisection has the reported t
    def f_of(uu):
        bb = np.einsum("nk,nkc->nc", basis(uu), cp) - O
        return D[:, 0] * bb[:, 1] - D[:, 1] * bb[:, 0]

    db = np.einsum("nk,nkc->nc", dbasis(u64), cp)
    speed = np.linalg.norm(db, axis=1)
    sin_theta = np.abs(D[:, 0] * db[:, 1] - D[:, 1] * db[:, 0]) / np.maximum(speed * np.linalg.norm(D, axis=1), 1e-30)
    bound = 1.5 * TOL / np.maximum(sin_theta, 1e-9) + 4e-6 * scale
    du = bound / np.maximum(speed, 1e-30)
    lo, hi = u64 - du, u64 + du
    # rays that cross steeply, away from the ends of the segment's parameter range (beyond them lies another segment
    # or nothing at all — the chord model and the curve may then disagree on hit/miss within the tolerance)
    steep = (sin_theta >= 0.2) & (lo >= 0.0) & (hi <= 1.0)
    assert steep.mean() > 0.6
    # sample the window (the Orzan pack holds segments that double back on themselves: an even number of crossings
    # inside a window shows no sign change between its ends) and take the crossing nearest to u
    grid = np.linspace(0.0, 1.0, 33)
    samples = lo[:, None] + (hi - lo)[:, None] * grid[None, :]
    fs = np.stack([f_of(samples[:, j]) for j in range(len(grid))], 1)
    flips = (fs[:, 1:] > 0) != (fs[:, :-1] > 0)
    bracket = flips.any(1)
    assert np.all(bracket[steep]), f"{(~bracket & steep).sum()} steep hits have no true crossing within the error bound"
    centre = 0.5 * (samples[:, 1:] + samples[:, :-1])
    nearest = np.argmin(np.where(flips, np.abs(centre - u64[:, None]), np.inf), axis=1)
    rows = np.arange(len(u64))
    lo, hi, flo = samples[rows, nearest], samples[rows, nearest + 1], fs[rows, nearest]
    for _ in range(50):
        mid = 0.5 * (lo + hi)
        fm = f_of(mid)
        left = (fm > 0) != (flo > 0)
        hi = np.where(left, mid, hi)
        lo = np.where(left, lo, mid)
        flo = np.where(left, flo, fm)
    us = 0.5 * (lo + hi)
    b = np.einsum("nk,nkc->nc", basis(us), cp) - O
    t_star = np.einsum("nc,nc->n", b, D) / np.einsum("nc,nc->n", D, D)
    along = (np.abs(us - u64) * speed)[steep]
    dt = np.abs(t_star - t64)[steep]
    assert np.all(dt <= bound[steep]), f"worst distance error {np.max(dt / bound[steep]):.2f} x bound"
    print(f"{name}: {hit.sum()} hits, max |B(u)-hit| {dev.max():.4f} px, median {np.median(dev):.5f}; steep rays "
          f"{steep.sum()}: max |dt| {dt.max():.4f}, max along-curve {along.max():.4f} px")


@pytest.mark.parametrize("name", ["arch.xml", "PortalDemo.xml", "DiffusionCurvePack/lady_bug.xml"])
def test_closest_chord_is_the_closest_curve_crossing(name, xml_dir, port_oracle):
    scene = po.ingest_xml(os.path.join(xml_dir, name), True)
    geom, cids = port_oracle.chords(scene)
    rays = random_rays(geom, 20_000, 11)
    ids, t, u = port_oracle.trace_rays(scene, rays)
    # a 4x finer polyline of every segment, in float64, tested by brute force with this file's own arithmetic
    nseg = len(scene["segment_indices"])
    K = np.zeros(nseg, np.int64)
    np.maximum.at(K, cids[:, 0], cids[:, 2])
    pts, pseg, pu = [], [], []
    cp_all = control_points(scene, np.arange(nseg))
    for s in range(nseg):
        m = 4 * int(K[s])
        uu = np.linspace(0.0, 1.0, m + 1)
        pts.append(np.einsum("mk,kc->mc", basis(uu), cp_all[s]))
        pseg.append(np.full(m + 1, s))
        pu.append(uu)
    pts, pseg, pu = np.concatenate(pts), np.concatenate(pseg), np.concatenate(pu)
    joins = pseg[1:] == pseg[:-1]  # consecutive points of one segment form a fine chord
    best_t = np.full(len(rays), np.inf)
    best_seg = np.full(len(rays), -1, np.int64)
    best_sin = np.ones(len(rays))
    r64 = rays.astype(np.float64)
    for lo in range(0, len(rays), 1000):
        R = r64[lo:lo + 1000]
        w = pts[None, :, :] - R[:, None, :2]
        e = R[:, None, 2] * w[..., 1] - R[:, None, 3] * w[..., 0]
        ea, eb = e[:, :-1], e[:, 1:]
        cross = ((ea > 0) != (eb > 0)) & joins[None, :]
        sfrac = np.where(cross, ea / np.where(cross, ea - eb, 1.0), 0.0)
        hx = w[:, :-1, 0] + sfrac * (w[:, 1:, 0] - w[:, :-1, 0])
        hy = w[:, :-1, 1] + sfrac * (w[:, 1:, 1] - w[:, :-1, 1])
        tt = np.where(cross, hx * R[:, None, 2] + hy * R[:, None, 3], np.inf)
        tt = np.where(tt > 0, tt, np.inf)
        j = np.argmin(tt, axis=1)
        rows = np.arange(len(R))
        best_t[lo:lo + 1000] = tt[rows, j]
        best_seg[lo:lo + 1000] = np.where(np.isfinite(tt[rows, j]), pseg[j], -1)
        cdx, cdy = pts[j + 1, 0] - pts[j, 0], pts[j + 1, 1] - pts[j, 1]
        best_sin[lo:lo + 1000] = np.abs(R[:, 2] * cdy - R[:, 3] * cdx) / np.maximum(np.hypot(cdx, cdy), 1e-30)
    hit, fine_hit = ids != MISS, np.isfinite(best_t)
    t64 = np.where(hit, t.astype(np.float64), np.inf)
    seg = np.where(hit, cids[np.where(hit, ids, 0), 0], -1)
    bound = 1.5 * TOL / np.maximum(best_sin, 1e-9) + 4e-6 * (1.0 + np.where(fine_hit, best_t, 0.0))
    agree = (hit == fine_hit) & (~hit | ((seg == best_seg) & (np.abs(best_t - t64) <= bound)))
    # Where the two models differ, the ray must GRAZE the curve: some point of the curve in front of the farther of the two
    # answers lies within 2 x tolerance of the ray while the curve runs nearly parallel to it there — the one situation in
    # which polylines of different resolution legitimately disagree (one crosses, the other passes by).
    odd = np.nonzero(~agree | (best_sin < 0.2))[0]
    odd = odd[~agree[odd]]
    seg_dir = pts[1:] - pts[:-1]
    seg_len = np.maximum(np.hypot(seg_dir[:, 0], seg_dir[:, 1]), 1e-30)
    for i in odd:
        O, D = r64[i, :2], r64[i, 2:]
        w = pts[:-1] - O
        dist = np.abs(D[0] * w[:, 1] - D[1] * w[:, 0])
        along_ray = w[:, 0] * D[0] + w[:, 1] * D[1]
        sin_here = np.abs(D[0] * seg_dir[:, 1] - D[1] * seg_dir[:, 0]) / seg_len
        reach = max(min(t64[i], 1e30) if hit[i] else 0.0, best_t[i] if fine_hit[i] else 0.0) + 1.0
        witness = joins & (dist <= 2 * TOL) & (along_ray > -1.0) & (along_ray < reach) & (sin_here < 0.35)
        # ... or the ray STARTS within the tolerance of the curve: one polyline passes in front of the origin, the other behind
        proj = np.clip((-(w[:, 0] * seg_dir[:, 0] + w[:, 1] * seg_dir[:, 1])) / seg_len ** 2, 0.0, 1.0)
        near = np.hypot(w[:, 0] + proj * seg_dir[:, 0], w[:, 1] + proj * seg_dir[:, 1])
        witness = witness | (joins & (near <= 2 * TOL))
        assert witness.any(), f"ray {i}: chord model {seg[i]}@{t64[i]:.4f} vs curve {best_seg[i]}@{best_t[i]:.4f} without a grazing contact"
    assert len(odd) < 0.01 * len(rays)
    both = hit & fine_hit & agree
    print(f"{name}: {len(rays)} rays, {hit.sum()} hits, {len(odd)} grazing disagreements, same segment and distance on the rest; "
          f"max |dt| {np.abs(best_t - t64)[both].max():.4f} px")
