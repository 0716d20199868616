"""Forward step (configs[1]: 1024 pairs, N=NP=64, G=8, k=5) timing with the per-kernel device times of the library's
profiler.  Environment switches (DPD_TC_SEG_HEAD, DPD_FV_IMPL, DPD_TC_TRACE ...) are read by the library at first use.
    python tools/fwd_time.py [steps] [pairs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdist_b200 import _lib, dpdist_and_aue as MODEL, synthetic, tf_util  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda", 0)
store = tf_util.VariableStore(device=dev, seed=1)
sets = []
for s in range(4):
    a, b, _ = synthetic.uniform_batch(seed=2 + s, batch=pairs, num_point=64)
    sets.append((torch.tensor(a, device=dev), torch.tensor(b, device=dev)))


def step(i):
    a, b = sets[i % 4]
    with tf_util.use_store(store):
        return MODEL.get_model(a, b, False, bn=0, Embedding_Size=512, k=5, sigma3dmfv=0.125)[0]


for i in range(10):
    out = step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    out = step(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print("env %s: %.3f ms/step, %.2f M evals/s  (checksum %.6f)" % (
    {k: v for k, v in os.environ.items() if k.startswith("DPD_")}, ms, pairs * 128 / ms / 1e3,
    float(out["pred_listAB"].double().sum())))
lib = _lib.load()
lib.dpd_profile_enable(1)
_lib.profile_read(reset=True)
for i in range(20):
    step(i)
torch.cuda.synchronize()
for k, v in sorted(_lib.profile_read(reset=True).items(), key=lambda kv: -kv[1][0])[:6]:
    print("   %-30s %8.4f ms/launch" % (k, v[0] / max(v[1], 1)))
