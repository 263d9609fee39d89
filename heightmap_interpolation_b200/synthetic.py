"""Seeded synthetic heightmaps and masks for parity tests and benchmarks.

The reference ships no data; these generators are the ones SURVEY.md
section 8(d) defines.  Everything is produced on the host with NumPy ``default_rng`` in row
blocks, so a grid of any size is reproducible and identical for the CUDA path and the oracle.
Masks follow the reference's convention: True == known cell
(/root/reference/heightmap_interpolation/inpainting/fd_pde_inpainter.py:165-166).
"""
import numpy as np

_BLOCK = 1024  # rows generated per block


def _shape(n):
    if isinstance(n, (tuple, list)):
        return int(n[0]), int(n[1])
    return int(n), int(n)


def gaussian_hill(n, dtype=np.float64):
    """100 * exp(-((x-C/2)^2 + (y-R/2)^2) / (2 (min(R,C)/6)^2))."""
    rows, cols = _shape(n)
    s = min(rows, cols) / 6.0
    x = np.arange(cols, dtype=np.float64)
    gx = np.exp(-((x - cols / 2.0) ** 2) / (2.0 * s * s))
    out = np.empty((rows, cols), dtype=dtype)
    for r0 in range(0, rows, _BLOCK):
        y = np.arange(r0, min(rows, r0 + _BLOCK), dtype=np.float64)
        gy = np.exp(-((y - rows / 2.0) ** 2) / (2.0 * s * s))
        out[r0:r0 + len(y)] = (100.0 * gy[:, None] * gx[None, :]).astype(dtype)
    return out


def bathymetry(n, seed=7, dtype=np.float64, row_range=None):
    """-2500 + sum_o (1200/2^o) sin(2 pi 2^o x/C + phi_o) cos(2 pi 2^o y/R + psi_o), o = 0..5.
    `row_range=(a, b)` returns only rows [a, b) of the grid (identical values), for row slabs."""
    rows, cols = _shape(n)
    a, b = (0, rows) if row_range is None else (int(row_range[0]), int(row_range[1]))
    rng = np.random.default_rng(seed)
    phi = rng.uniform(0.0, 2.0 * np.pi, 6)
    psi = rng.uniform(0.0, 2.0 * np.pi, 6)
    x = np.arange(cols, dtype=np.float64)
    sx = [np.sin(2.0 * np.pi * (2 ** o) * x / cols + phi[o]) for o in range(6)]
    out = np.empty((b - a, cols), dtype=dtype)
    for r0 in range(a, b, _BLOCK):
        y = np.arange(r0, min(b, r0 + _BLOCK), dtype=np.float64)
        acc = np.full((len(y), cols), -2500.0)
        for o in range(6):
            cy = np.cos(2.0 * np.pi * (2 ** o) * y / rows + psi[o])
            acc += (1200.0 / 2 ** o) * cy[:, None] * sx[o][None, :]
        out[r0 - a:r0 - a + len(y)] = acc.astype(dtype)
    return out


def mask_random(n, p, seed=0, row_range=None):
    """Known where rng.random > p (so a fraction p of the cells is unknown).  `row_range=(a, b)`
    returns only rows [a, b) of the same mask (the generator is advanced past the other blocks)."""
    rows, cols = _shape(n)
    a, b = (0, rows) if row_range is None else (int(row_range[0]), int(row_range[1]))
    rng = np.random.default_rng(seed)
    out = np.empty((b - a, cols), dtype=bool)
    for r0 in range(0, rows, _BLOCK):
        nb = min(rows, r0 + _BLOCK) - r0
        if r0 >= b:
            break
        if r0 + nb <= a:
            rng.bit_generator.advance(nb * cols)       # one 64-bit draw per double
            continue
        blk = rng.random((nb, cols)) > p
        lo, hi = max(a, r0), min(b, r0 + nb)
        out[lo - a:hi - a] = blk[lo - r0:hi - r0]
    return out


def mask_tracks(n, line_w=8, gap=16, cross_every=256):
    """Survey tracks: known iff row % (line_w+gap) < line_w or col % cross_every < line_w."""
    rows, cols = _shape(n)
    r = (np.arange(rows) % (line_w + gap)) < line_w
    c = (np.arange(cols) % cross_every) < line_w
    return r[:, None] | c[None, :]


def mask_holes(n, count=64, radius=32, seed=1):
    """`count` discs of unknown cells at rng.integers centres."""
    rows, cols = _shape(n)
    rng = np.random.default_rng(seed)
    cy = rng.integers(0, rows, count)
    cx = rng.integers(0, cols, count)
    out = np.ones((rows, cols), dtype=bool)
    for y0, x0 in zip(cy, cx):
        ya, yb = max(0, y0 - radius), min(rows, y0 + radius + 1)
        xa, xb = max(0, x0 - radius), min(cols, x0 + radius + 1)
        yy, xx = np.ogrid[ya:yb, xa:xb]
        out[ya:yb, xa:xb] &= ((yy - y0) ** 2 + (xx - x0) ** 2) > radius * radius
    return out


def zero_unknown(image, mask):
    """What the reference CLI does before inpainting (apps/interpolate_netcdf4.py:192)."""
    image = image.copy()
    image[~mask] = 0
    return image
