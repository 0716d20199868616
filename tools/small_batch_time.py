"""Step time of the reference's own training configuration (global batch 16 pairs) eager vs CUDA-graph replay."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdist_b200 import synthetic, train  # noqa: E402

dev = torch.device("cuda", 0)
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pcA, pcB, lab = synthetic.chair_batch(1, pairs, 64)
a, b, l = (torch.tensor(x, device=dev) for x in (pcA, pcB, lab))
for graph in (False, True):
    tr = train.DPDistTrainer(dev, seed=1, cuda_graph=graph)
    for _ in range(6):
        tr.step(a, b, l)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        loss = tr.step(a, b, l)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("batch %d, cuda_graph=%s: %.3f ms/step, %.0f pairs/s, loss %.5f" % (pairs, graph, ms, pairs / ms * 1e3, float(loss)))
