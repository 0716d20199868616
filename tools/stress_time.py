"""BASELINE configs[4] (stress): 4096 pairs, N = NP = 512, G = 8, k = 5 on one GPU -- forward time, evals/s, 3DmFV GB/s."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdist_b200 import _lib, dpdist_and_aue as MODEL, dpdist_util, tf_util  # noqa: E402

dev = torch.device("cuda", 0)
B, N = 4096, 512
g = torch.Generator().manual_seed(3)
pcA = (torch.rand((B, N, 3), generator=g) * 1.6 - 0.8).to(dev)
pcB = (torch.rand((B, N, 3), generator=g) * 1.6 - 0.8).to(dev)
store = tf_util.VariableStore(device=dev, seed=1)


def step():
    with tf_util.use_store(store):
        return MODEL.get_model(pcA, pcB, False, bn=0, Embedding_Size=512, k=5, sigma3dmfv=0.125)[0]


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
evals = 2 * B * N
print("config E forward: %.2f ms per batch of %d pairs (%.2f M rows) -> %.1f M evals/s" % (ms, B, evals / 1e6, evals / ms / 1e3))
lib = _lib.load()
lib.dpd_profile_enable(1)
_lib.profile_read(reset=True)
step()
torch.cuda.synchronize()
prof = _lib.profile_read(reset=True)
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:8]:
    print("  %-28s %8.3f ms  (%d launches)" % (k, v[0], v[1]))
fv_ms = prof.get("fv_g8", (0, 1))[0]
print("3DmFV at N=512: %.3f ms for %d clouds = %.0f GB/s algorithmic (47,104 B/cloud)" % (fv_ms, 2 * B, 2 * B * 47104 / fv_ms / 1e6))
