"""Quick check of the tcgen05 conv_layer path against the SIMT path (both through the C ABI)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import halide_b200 as hb
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_conv_layer_gpu import make, run
l = hb.load_library()
inp, filt, bias = make(0, 1.0)
l.halide_b200_conv_use_tensor_cores(0)
ref = run(hb, inp, filt, bias)
l.halide_b200_conv_use_tensor_cores(1)
got = run(hb, inp, filt, bias)
err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30)
print("tc vs simt: max rel err", err.max(), "mean", err.mean(), "nonzero", np.count_nonzero(got), "of", got.size)
bad = np.argwhere(err > 1e-4)
print("bad count", len(bad), bad[:5].tolist())
if len(bad):
    b = tuple(bad[0]); print(got[b], ref[b])
torch.cuda.synchronize()
import subprocess
