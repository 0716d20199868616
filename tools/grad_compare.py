"""Weight gradients of one training batch (1024 pairs, 131072 rows) saved to disk; run once with DPD_TC_BWD=1 and once with
DPD_TC_BWD=0, then `python tools/grad_compare.py diff a.pt b.pt` prints the deviation of the tensor-core backward from
the fp32 SIMT backward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if sys.argv[1] == "diff":
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    for n in sorted(a):
        d = (a[n].double() - b[n].double()).abs().max().item()
        print("%-52s max|g| %.3e  max deviation %.3e  (%.2e of max)" % (n, b[n].abs().max().item(), d, d / b[n].abs().max().item()))
else:
    from dpdist_b200 import dpdist_and_aue as MODEL, synthetic, tf_util
    dev = torch.device("cuda", 0)
    store = tf_util.VariableStore(device=dev, seed=1)
    pcA, pcB, lab = synthetic.uniform_batch(2, 1024, 64)
    a, b, l = (torch.tensor(x, device=dev) for x in (pcA, pcB, lab))
    tf_util.clear_collections()
    with tf_util.use_store(store):
        pred, ep, _ = MODEL.get_model(a, b, True, bn=0, Embedding_Size=512, k=5, sigma3dmfv=0.125)
        MODEL.get_loss(pred, ep, l)
    tf_util.get_collection("loss_samples")[-1].backward()
    torch.save({n: p.grad.cpu() for n, p in store.vars.items()}, sys.argv[1])
    print("saved", sys.argv[1])
