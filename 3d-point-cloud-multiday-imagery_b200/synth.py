"""Synthetic multi-day height-map stacks (SURVEY.md section 8(d) "Synthetic inputs").

The reference ships no data for this path (its sample rasters are git-LFS stubs and the
stereo front end needs the external ASP binary), so benchmarks and parity tests run on
stacks shaped like what ``members/rafael/disparity/disparity.py:213-224`` writes: one
float32 height raster per pair/day, NaN where undefined, a few out-of-range sentinels
(``plugin.py:151`` rejects ``|h| > MAX_DISP/2``).
"""
from __future__ import annotations

import numpy as np
import torch

MAX_ABS_HEIGHT = 144.0  # MAX_DISP / 2, members/rafael/disparity/constants.py:54 + plugin.py:151


def make_stack(D: int, H: int, W: int, seed: int = 0, device="cpu", nan_frac=0.02,
               sentinel_frac=0.005, n_buildings=64) -> torch.Tensor:
    """float32 ``[D, H, W]``: tilted ground + building plateaus + per-day bias + noise.

    ground ``0.002*x - 0.001*y``; ``n_buildings`` axis-aligned plateaus (height U[3,40],
    side U[16,128] px, clipped to the raster); day ``d`` adds ``0.25*d`` and N(0, 0.3^2)
    noise; ``nan_frac`` of pixels become NaN and ``sentinel_frac`` become +-200.
    """
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    ys = torch.arange(H, device=dev, dtype=torch.float32)[:, None]
    xs = torch.arange(W, device=dev, dtype=torch.float32)[None, :]
    base = 0.002 * xs - 0.001 * ys
    base = base.contiguous()
    # building layout comes from a host RNG so it is identical on every device
    rs = np.random.RandomState(int(seed) + 7919)
    for _ in range(n_buildings):
        bw = int(rs.randint(16, 129))
        bh = int(rs.randint(16, 129))
        x0 = int(rs.randint(0, max(1, W - 1)))
        y0 = int(rs.randint(0, max(1, H - 1)))
        base[y0 : min(H, y0 + bh), x0 : min(W, x0 + bw)] += float(rs.uniform(3.0, 40.0))
    out = torch.empty((D, H, W), device=dev, dtype=torch.float32)
    for d in range(D):
        noise = torch.randn((H, W), generator=g, device=dev, dtype=torch.float32) * 0.3
        day = base + (0.25 * d) + noise
        u = torch.rand((H, W), generator=g, device=dev, dtype=torch.float32)
        day = torch.where(u < nan_frac, torch.full_like(day, float("nan")), day)
        hi = u > (1.0 - sentinel_frac)
        sign = torch.where(u > (1.0 - 0.5 * sentinel_frac), 1.0, -1.0)
        day = torch.where(hi, 200.0 * sign, day)
        out[d] = day
    return out


def make_stack_range(D: int, H: int, W: int, pix_begin: int, pix_count: int, seed: int = 0, device="cpu",
                     nan_frac=0.02, sentinel_frac=0.005, n_buildings=64) -> torch.Tensor:
    """float32 ``[pix_count]``: the pixels ``[pix_begin, pix_begin + pix_count)`` of a conceptual
    ``[D, H, W]`` stack of the same kind as ``make_stack`` -- a rank's row band of a stack that is
    too large to generate whole (BASELINE.json configs[2]: 20 x 8192 x 8192, sharded by row
    bands).  ``pix_begin`` / ``pix_count`` must be multiples of ``W``.  The building layout
    depends on ``seed`` only; the noise of a piece depends on (seed, day, first row), so a band
    is reproducible but the stack as a whole depends on how it was cut."""
    assert pix_begin % W == 0 and pix_count % W == 0 and pix_begin + pix_count <= D * H * W
    dev = torch.device(device)
    rs = np.random.RandomState(int(seed) + 7919)
    boxes = []
    for _ in range(n_buildings):
        bw, bh = int(rs.randint(16, 129)), int(rs.randint(16, 129))
        x0, y0 = int(rs.randint(0, max(1, W - 1))), int(rs.randint(0, max(1, H - 1)))
        boxes.append((x0, y0, bw, bh, float(rs.uniform(3.0, 40.0))))
    out = torch.empty(pix_count, device=dev, dtype=torch.float32)
    xs = torch.arange(W, device=dev, dtype=torch.float32)[None, :]
    row = pix_begin // W
    end = (pix_begin + pix_count) // W
    pos = 0
    max_rows = max(1, (64 << 20) // W)  # pieces of at most 64 Mi pixels keep the temporaries small
    while row < end:
        d, r0 = divmod(row, H)
        r1 = min(H, r0 + (end - row), r0 + max_rows)
        ys = torch.arange(r0, r1, device=dev, dtype=torch.float32)[:, None]
        piece = (0.002 * xs - 0.001 * ys).contiguous()
        for x0, y0, bw, bh, hgt in boxes:
            a, b = max(y0, r0), min(H, y0 + bh, r1)
            if a < b:
                piece[a - r0:b - r0, x0:min(W, x0 + bw)] += hgt
        g = torch.Generator(device=dev)
        g.manual_seed((int(seed) * 1000003 + d * 8191 + r0) & 0x7FFFFFFF)
        piece += 0.25 * d + torch.randn(piece.shape, generator=g, device=dev, dtype=torch.float32) * 0.3
        u = torch.rand(piece.shape, generator=g, device=dev, dtype=torch.float32)
        piece = torch.where(u < nan_frac, torch.full_like(piece, float("nan")), piece)
        sign = torch.where(u > (1.0 - 0.5 * sentinel_frac), 1.0, -1.0)
        piece = torch.where(u > (1.0 - sentinel_frac), 200.0 * sign, piece)
        n = (r1 - r0) * W
        out[pos:pos + n] = piece.reshape(-1)
        pos += n
        row += r1 - r0
        del piece, u, sign
    return out


def init_from_points(points_xyz: np.ndarray, k: int, seed: int = 0) -> np.ndarray:
    """``points[RandomState(seed).choice(N, K, replace=False)]`` as float64."""
    n = points_xyz.shape[0]
    idx = np.random.RandomState(int(seed)).choice(n, int(k), replace=False)
    return np.ascontiguousarray(points_xyz[np.sort(idx)], dtype=np.float64)
