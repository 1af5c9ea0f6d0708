import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import halide_b200 as hb
from oracle import pyoracle as po
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import run_local_laplacian, u16_frame
l = hb.load_library()
for (h, w) in [(16, 16), (33, 47)]:
    img = u16_frame((3, h, w), h * 1000 + w)
    want = po.local_laplacian(img, 8, 1 / 7, 1.0)
    for mask in (7, 6, 5, 3, 0):
        l.halide_b200_ll_force_generic(mask)
        got = run_local_laplacian(hb, img, 8, 1 / 7, 1.0)
        bad = np.argwhere(got != want)
        print(h, w, "mask", mask, "mismatches", len(bad), bad[:6].tolist(), [(int(got[tuple(b)]), int(want[tuple(b)])) for b in bad[:4]])
l.halide_b200_ll_force_generic(0)
