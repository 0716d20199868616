"""Interleaved A/B of one library switch inside ONE process (drift of clocks / power state cancels): the forward step of
configs[1] is timed in alternating blocks with the environment variable set to each value.
    python tools/ab_env.py DPD_TC_SEG_HEAD 4 6 [blocks] [steps per block] [idle seconds before every block]
With an idle time (e.g. 1.0) every block is a cold burst at the boost clock - what a 20-step bench run measures; without
it the blocks run back to back at the power-capped clock.  Only switches the library reads per call can be compared this way (DPD_TC_SEG_HEAD, DPD_FV_IMPL)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdist_b200 import dpdist_and_aue as MODEL, synthetic, tf_util  # noqa: E402

var, va, vb = sys.argv[1], sys.argv[2], sys.argv[3]
blocks = int(sys.argv[4]) if len(sys.argv) > 4 else 12
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 40
idle = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
dev = torch.device("cuda", 0)
store = tf_util.VariableStore(device=dev, seed=1)
sets = []
for s in range(4):
    a, b, _ = synthetic.uniform_batch(seed=2 + s, batch=1024, num_point=64)
    sets.append((torch.tensor(a, device=dev), torch.tensor(b, device=dev)))


def step(i):
    a, b = sets[i % 4]
    with tf_util.use_store(store):
        return MODEL.get_model(a, b, False, bn=0, Embedding_Size=512, k=5, sigma3dmfv=0.125)[0]


tot = {va: 0.0, vb: 0.0}
for v in (va, vb):
    os.environ[var] = v
    for i in range(20):
        step(i)
torch.cuda.synchronize()
for blk in range(blocks):
    for v in ((va, vb) if blk % 2 == 0 else (vb, va)):
        os.environ[var] = v
        if idle > 0:
            import time
            time.sleep(idle)
            for i in range(3):
                step(i)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        tot[v] += e0.elapsed_time(e1)
for v in (va, vb):
    ms = tot[v] / (blocks * steps)
    print("%s=%s: %.4f ms/step, %.2f M evals/s" % (var, v, ms, 131072 / ms / 1e3))
print("ratio %s/%s = %.4f" % (va, vb, tot[va] / tot[vb]))
