"""'nearest' initialiser: hmi_init_nearest_host (B200) against scipy.interpolate.griddata (the reference's
implementation, inpainting/initializer.py:19-34) — wall time and agreement."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from heightmap_interpolation_b200 import synthetic
from heightmap_interpolation_b200.inpainting.initializer import Initializer

for n, do_scipy in ((1024, True), (2048, True), (4096, False), (8192, False)):
    for mname, mask in (("random 40 %", synthetic.mask_random(n, 0.4, seed=3)),
                        ("tracks 8/16/256", synthetic.mask_tracks(n, 8, 16, 256)),
                        ("64 holes r=%d" % (n // 64), synthetic.mask_holes(n, 64, n // 64, seed=1))):
        img = synthetic.zero_unknown(synthetic.bathymetry(n, dtype=np.float32), mask)
        Initializer("nearest").initialize(img.copy(), mask)          # warm-up
        a = img.copy(); t = time.perf_counter(); Initializer("nearest").initialize(a, mask); tg = time.perf_counter() - t
        line = "%5d^2 fp32 %-16s B200 %.2f ms" % (n, mname, 1e3 * tg)
        if do_scipy:
            b = img.copy(); t = time.perf_counter(); Initializer("nearest", nearest_backend="scipy").initialize(b, mask)
            ts = time.perf_counter() - t
            line += "   scipy %.2f s (%.0fx)   cells differing (equidistant ties): %.2f %%" % (
                ts, ts / tg, 100.0 * float((a != b).mean()))
        print(line, flush=True)
