"""Runs the G = 8 3DmFV kernel a few times on one batch (for ncu captures):  python tools/fv_one.py [clouds] [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdist_b200 import dpdist_util  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
g = torch.Generator().manual_seed(5)
big = (torch.rand((n, N, 3), generator=g) * 1.6 - 0.8).cuda()
for _ in range(4):
    out = dpdist_util.get_3dmfv_tf(big, n_gaussians=512, sigma=0.125, flatten=False)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
