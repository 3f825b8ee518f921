"""Seeded synthetic workloads for the parity tests and bench.py.

Frames: a pool of distinct camera frames — 25 % uniform noise, 50 % Gaussian-blurred noise (sigma 2 px,
stretched to 0..255: gives weak-edge chains so the edge filter's hysteresis actually works), 25 %
track-like scenes (grey ground, white and yellow lane lines, sky) so both default colour ranges fire.
Car states: points scattered around a recorded centre line, with exact waypoint copies (ties) and far
outliers (the reference's distance-100 sentinel, track_data_process.py:93).
Only numpy is used so the same bytes come out on every box.
"""
from __future__ import annotations

import numpy as np

SEED = 20261018


def _blur_axis(a: np.ndarray, k: np.ndarray, axis: int) -> np.ndarray:
    r = len(k) // 2
    pad = [(0, 0)] * a.ndim
    pad[axis] = (r, r)
    ap = np.pad(a, pad, mode="reflect")
    out = np.zeros_like(a)
    n = a.shape[axis]
    for i, kv in enumerate(k):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(i, i + n)
        out += kv * ap[tuple(sl)]
    return out


def _gauss(sigma: float) -> np.ndarray:
    r = int(3 * sigma + 0.5)
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def _noise(rng, n, h, w):
    return rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)


def _smooth(rng, n, h, w, sigma=2.0):
    a = rng.random((n, h, w, 3))
    k = _gauss(sigma)
    a = _blur_axis(_blur_axis(a, k, 1), k, 2)
    lo = a.min(axis=(1, 2, 3), keepdims=True)
    hi = a.max(axis=(1, 2, 3), keepdims=True)
    return np.clip((a - lo) / (hi - lo) * 255.0 + 0.5, 0, 255).astype(np.uint8)


def _track(rng, n, h, w):
    out = np.empty((n, h, w, 3), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    for i in range(n):
        img = np.clip(90 + 12 * rng.standard_normal((h, w, 1)) + 4 * rng.standard_normal((h, w, 3)), 0, 255)
        horizon = int(h * (0.28 + 0.1 * rng.random()))
        sky = np.array([135, 190, 235]) + 6 * rng.standard_normal((horizon, w, 3))
        img[:horizon] = np.clip(sky, 0, 255)
        vx = w * (0.3 + 0.4 * rng.random())                    # vanishing point
        for colour, base in (((250, 250, 250), 0.08 + 0.2 * rng.random()), ((240, 220, 40), 0.55 + 0.3 * rng.random()),
                             ((250, 250, 250), 0.9 + 0.2 * rng.random())):
            xb = w * base                                       # x position at the bottom row
            t = (yy - horizon) / max(1, (h - 1 - horizon))      # 0 at horizon, 1 at bottom
            xc = vx + (xb - vx) * t + 6 * np.sin(t * 3 + rng.random() * 6)
            half = 0.6 + 1.6 * t
            on = (yy >= horizon) & (np.abs(xx - xc) <= half)
            img[on] = colour
        out[i] = img.astype(np.uint8)
    return out


def frame_pool(n: int, h: int = 120, w: int = 160, seed: int = SEED) -> np.ndarray:
    """(n,h,w,3) uint8 pool: frames i%4==0 noise, ==1/2 smooth, ==3 track-like."""
    rng = np.random.default_rng(seed + h * 1000 + w)
    kinds = np.arange(n) % 4
    out = np.empty((n, h, w, 3), np.uint8)
    for kind, gen in ((0, _noise), (1, _smooth), (2, _smooth), (3, _track)):
        sel = np.nonzero(kinds == kind)[0]
        if len(sel):
            out[sel] = gen(rng, len(sel), h, w)
    return out


def expand_indices(n_total: int, pool_size: int, start: int = 0):
    """Frame i of a batch = pool[i % P] + brightness offset ((i // P) % 32 - 16), saturating (global index i)."""
    i = np.arange(start, start + n_total, dtype=np.int64)
    return (i % pool_size).astype(np.int64), ((i // pool_size) % 32 - 16).astype(np.int16)


def expand_numpy(pool: np.ndarray, n_total: int, start: int = 0) -> np.ndarray:
    idx, off = expand_indices(n_total, pool.shape[0], start)
    return np.clip(pool[idx].astype(np.int16) + off[:, None, None, None], 0, 255).astype(np.uint8)


def expand_torch(pool_dev, n_total: int, start: int = 0, chunk: int = 4096):
    """Same expansion on the device (a uint8 CUDA tensor pool); avoids pushing the batch over PCIe."""
    import torch

    p = pool_dev.shape[0]
    out = torch.empty((n_total,) + tuple(pool_dev.shape[1:]), dtype=torch.uint8, device=pool_dev.device)
    for s in range(0, n_total, chunk):
        e = min(n_total, s + chunk)
        i = torch.arange(start + s, start + e, device=pool_dev.device, dtype=torch.int64)
        off = ((i // p) % 32 - 16).to(torch.int16)
        out[s:e] = (pool_dev[i % p].to(torch.int16) + off[:, None, None, None]).clamp_(0, 255).to(torch.uint8)
    return out


def synthetic_track(n_wp: int = 1200, seed: int = SEED) -> np.ndarray:
    """A closed centre line with repeated points (recorded tracks contain duplicates)."""
    rng = np.random.default_rng(seed + 7)
    t = np.sort(rng.random(n_wp)) * 2 * np.pi
    r = 40 + 8 * np.sin(3 * t) + 3 * np.cos(7 * t)
    wp = np.stack([50 + r * np.cos(t), 0.55 + 0.05 * np.sin(5 * t), 50 + r * np.sin(t)], axis=1)
    wp = np.round(wp, 5)                                         # JSON-like decimal text
    dup = rng.random(n_wp) < 0.3
    dup[0] = False
    for i in np.nonzero(dup)[0]:
        wp[i] = wp[i - 1]
    return wp


def car_states(waypoints: np.ndarray, n: int, seed: int = 4):
    """xyz f64 (n,3), cur speed f64, model speed f32 (pre x20), model steering f32."""
    rng = np.random.default_rng(seed)
    wp = np.asarray(waypoints, np.float64)
    base = wp[rng.integers(0, wp.shape[0], size=n)]
    xyz = base + rng.standard_normal((n, 3)) * np.array([1.5, 0.05, 1.5])
    kind = rng.random(n)
    exact = kind < 0.01
    xyz[exact] = base[exact]
    far = (kind >= 0.01) & (kind < 0.02)
    xyz[far] += np.array([400.0, 50.0, -300.0])
    cur = rng.random(n) * 20.0
    model_spd = (cur / 20.0 + 0.1 * rng.standard_normal(n)).astype(np.float32)
    steer = (rng.random(n) * 2.6 - 1.3).astype(np.float32)
    return np.ascontiguousarray(xyz), cur, model_spd, steer
