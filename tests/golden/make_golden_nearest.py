"""Golden fixture for the 'nearest' initialiser, produced by the UNMODIFIED reference
(heightmap_interpolation/inpainting/initializer.py:19-34, scipy.interpolate.griddata 'nearest').

    python tests/golden/make_golden_nearest.py        # build container only

Writes nearest.npz: per case the input image/mask, the reference's initialised image, the exact
squared distance to the nearest known cell and whether that cell is unique (brute force), plus one
harmonic solve started from the reference's initialisation (result + update count).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference, run_reference, surface   # noqa: E402


def brute(mask):
    """Exact squared distance to the nearest known cell and the number of known cells at that distance."""
    rows, cols = mask.shape
    ky, kx = np.nonzero(mask)
    d2 = np.zeros(mask.shape, dtype=np.int64)
    cnt = np.ones(mask.shape, dtype=np.int64)
    for i, j in zip(*np.nonzero(~mask)):
        dd = (ky - i) ** 2 + (kx - j) ** 2
        m = dd.min()
        d2[i, j] = m
        cnt[i, j] = int((dd == m).sum())
    return d2, cnt


def main():
    classes = load_reference()
    from heightmap_interpolation.inpainting.initializer import Initializer
    rng = np.random.default_rng(11)
    out, meta = {}, []
    cases = []
    # random 50 % mask with a few big holes, edges and corners unknown
    m = rng.random((97, 131)) > 0.5
    m[30:60, 40:90] = False
    m[0, :9] = False
    m[-1, -9:] = False
    m[:, 0] = False
    cases.append(("nn00", 97, 131, "f64", m))
    # sparse known cells (5 %): long searches
    m = rng.random((64, 200)) > 0.95
    cases.append(("nn01", 64, 200, "f32", m))
    # survey tracks
    m = np.zeros((120, 90), dtype=bool)
    m[::17, :] = True
    m[:, ::31] = True
    cases.append(("nn02", 120, 90, "f64", m))
    for key, rows, cols, dt, mask in cases:
        dtype = np.float64 if dt == "f64" else np.float32
        img = surface(rows, cols, 5, dtype)
        img[~mask] = 0
        init = Initializer("nearest").initialize(img.copy(), mask)
        d2, cnt = brute(mask)
        out[key + "_image"] = img
        out[key + "_mask"] = mask
        out[key + "_init"] = init
        out[key + "_d2"] = d2
        out[key + "_unique"] = cnt == 1
        meta.append("%s|%d|%d|%s" % (key, rows, cols, dt))
    # one solve from the reference's nearest initialisation (CLI-style options)
    img, mask = out["nn00_image"].copy(), out["nn00_mask"]
    opts = dict(update_step_size=0.2, term_criteria="absolute_percent", term_thres=1e-6, term_check_iters=10,
                max_iters=100000, init_with="nearest", convolver="opencv")
    res, updates, diff, term, caller_after, inp = run_reference(classes["harmonic"], img, mask, **opts)
    out["nn00_solve"] = res
    out["nn00_solve_meta"] = np.array(["%d|%r|%d|%r" % (updates, diff, int(term), float(inp.term_thres))])
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(HERE, "nearest.npz"), **out)
    print("wrote nearest.npz:", meta, out["nn00_solve_meta"])


if __name__ == "__main__":
    main()
