import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import test_gpu_conv as T
from read_b200 import _lib as L
for name, srcs, cout, k, elu, kw in T.TC_CASES[:3]:
    got, want = T.run_conv(srcs, cout, k, 1, elu, True, L.CONV_TCGEN05, **kw)[:2]
    nan = ~torch.isfinite(got)
    err = (got - want).abs()
    err[nan] = 0
    print(name, "shape", tuple(got.shape), "nan count", int(nan.sum()), "max err (finite)", float(err.max()))
    m = nan.any(dim=1)[0]            # [H,W]
    e = (err.max(dim=1)[0][0] > 0.05)
    for y in range(min(m.shape[0], 20)):
        print("".join("N" if m[y, x] else ("x" if e[y, x] else ".") for x in range(m.shape[1])))
