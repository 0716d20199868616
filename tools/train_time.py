"""Times the DPDist training step (forward + backward + Adam) on one GPU and prints the per-kernel breakdown.
    python tools/train_time.py [pairs_per_step]          (DPD_TC_BWD=0 selects the fp32 SIMT backward)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdist_b200 import _lib, synthetic, train  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
tr = train.DPDistTrainer(dev, seed=1)
pcA, pcB, lab = synthetic.uniform_batch(2, pairs, 64)
a, b, l = (torch.tensor(x, device=dev) for x in (pcA, pcB, lab))
for _ in range(3):
    tr.step(a, b, l)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    loss = tr.step(a, b, l)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("DPD_TC_BWD=%s  %d pairs/step: %.3f ms/step, %.0f pairs/s, loss %.5f" % (
    os.environ.get("DPD_TC_BWD", "1"), pairs, ms, pairs / ms * 1e3, float(loss)))
lib = _lib.load()
lib.dpd_profile_enable(1)
_lib.profile_read(reset=True)
for _ in range(3):
    tr.step(a, b, l)
torch.cuda.synchronize()
prof = _lib.profile_read(reset=True)
lib.dpd_profile_enable(0)
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print("  %-28s %8.3f ms/step  (%d launches/step)" % (k, v[0] / 3, v[1] // 3))

# DPDist as a frozen loss: gradient into input1 only (the PCRNet-ours / AUE use), same batch
from dpdist_b200.dpdist_loss import DPDistLoss  # noqa: E402
dl = DPDistLoss(num_point=64, device=dev, seed=1)
x = a.clone().requires_grad_(True)
for _ in range(3):
    x.grad = None
    dl.loss(x, b).backward()
torch.cuda.synchronize()
e0.record()
for _ in range(n):
    x.grad = None
    dl.loss(x, b).backward()
e1.record()
torch.cuda.synchronize()
print("frozen-loss forward + backward into input1, %d pairs: %.3f ms" % (pairs, e0.elapsed_time(e1) / n))
lib.dpd_profile_enable(1)
_lib.profile_read(reset=True)
for _ in range(3):
    x.grad = None
    dl.loss(x, b).backward()
torch.cuda.synchronize()
prof = _lib.profile_read(reset=True)
lib.dpd_profile_enable(0)
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:12]:
    print("  %-28s %8.3f ms/step  (%d launches/step)" % (k, v[0] / 3, v[1] // 3))
