"""A/B of the two G = 8 3DmFV kernels on one GPU: timing at several batch sizes and output agreement.
    python tools/fv_ab.py            (DPD_FV_IMPL is switched per call)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdist_b200 import dpdist_util  # noqa: E402


def run(x, impl):
    os.environ["DPD_FV_IMPL"] = impl
    return dpdist_util.get_3dmfv_tf(x, n_gaussians=512, sigma=0.125, flatten=False)


g = torch.Generator().manual_seed(5)
for n, N in ((2048, 64), (16384, 64), (65536, 64), (2048, 512), (2051, 200)):
    big = (torch.rand((n, N, 3), generator=g) * 1.6 - 0.8).cuda()
    res = {}
    for impl in ("old", "new"):
        for _ in range(3):
            out = run(big, impl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20 if n <= 16384 else 5
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            out = run(big, impl)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[impl] = out
        print("%s clouds %d N %d: %.1f us  %.2f M clouds/s  %.0f GB/s" % (impl, n, N, ms * 1e3, n / ms / 1e3,
                                                                     n * 4 * (3 * N + 20 * 512) / ms / 1e6), flush=True)
    d = (res["old"] - res["new"]).abs()
    mm = [1, 5, 6, 7, 8, 9, 10, 14, 15, 16, 17, 18, 19]
    print("   max |old - new| = %.3g (max/min channels %.3g), finite %s" % (float(d.max()), float(d[:, :, mm].max()),
                                                                          bool(torch.isfinite(res["new"]).all())), flush=True)
    del big, res
