"""conv_version 3 head (the reference's other implicit net, --implicit_net_type 3; utils/dpdist_util.py:640-687,
resnet3d :394-410) against the CPU oracle: forward and the gradients of loss_samples w.r.t. its 16 variables."""
import numpy as np
import pytest
import torch

from dpdist_b200 import dpdist_and_aue as MODEL, synthetic, tf_util
from oracle import dpdist_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GAINS = (8.0,) + (1.5,) * 5 + (2.0, 1.0)


@pytest.mark.parametrize("G,k,H,B,N", [(8, 5, 1024, 2, 64), (4, 3, 128, 3, 16)])
def test_cv3_forward_and_gradients_match_the_oracle(G, k, H, B, N):
    pcA, pcB, labels = synthetic.uniform_batch(80 + G, B, N, outside_frac=0.05)
    labels = labels * 3.0
    var = O.init_cv3_variables(k=k, mlp=(H, H, H), seed=4, gain=GAINS, bias_std=0.05, out_bias=1.0)
    v = {n: t.clone().requires_grad_(True) for n, t in var.items()}
    kw = dict(Embedding_Size=G ** 3, k=k, sigma3dmfv=1.0 / G)
    with O.tf_cpu_numerics():
        p, _, _ = O.get_model(torch.tensor(pcA), torch.tensor(pcB), v, conv_version=3, **kw)
        loss, _ = O.get_loss(p, {}, torch.tensor(labels))
    loss.backward()
    store = tf_util.VariableStore(device=DEV)
    store.load_state_dict(var, strict=False)
    tf_util.clear_collections()
    with tf_util.use_store(store):
        pred, ep, _ = MODEL.get_model(torch.tensor(pcA, device=DEV), torch.tensor(pcB, device=DEV), True, bn=0, conv_version=3,
                                      localSNmlp=[H, H, H], reuse=True, **kw)
        MODEL.get_loss(pred, ep, torch.tensor(labels, device=DEV))
    lg = tf_util.get_collection("loss_samples")[-1]
    lg.backward()
    torch.cuda.synchronize()
    assert sorted(store.names()) == sorted(var)                       # same 16 TF variable names
    for key in ("pred_listAB", "pred_listBA"):
        g, w = pred[key].detach().cpu(), p[key].detach()
        assert g.shape == w.shape
        assert float((g - w).abs().max()) <= 1e-4 * float(w.abs().max()) + 1e-5, key
    spread = p["pred_listAB"][..., 0]
    assert float(spread[spread > 0].std()) > 0.05                     # the comparison is not atol-dominated
    assert abs(float(lg) - float(loss)) <= 1e-5 * max(1.0, abs(float(loss)))
    for n, t in v.items():
        got = store.vars[n].grad
        assert got is not None and got.shape == t.grad.shape, n
        scale = float(t.grad.abs().max())
        assert scale > 0, n
        err = float((got.cpu() - t.grad).abs().max())
        assert err <= 5e-4 * scale, "%s: %.3e vs scale %.3e" % (n, err, scale)


def test_cv3_inference_equals_training_forward_and_needs_no_saved_state():
    pcA, pcB, _ = synthetic.uniform_batch(91, 2, 32)
    var = O.init_cv3_variables(k=3, mlp=(128, 128, 128), seed=6, gain=GAINS, bias_std=0.05, out_bias=1.0)
    store = tf_util.VariableStore(device=DEV)
    store.load_state_dict(var, strict=False)
    a, b = torch.tensor(pcA, device=DEV), torch.tensor(pcB, device=DEV)
    kw = dict(bn=0, conv_version=3, Embedding_Size=64, k=3, sigma3dmfv=0.25, localSNmlp=[128] * 3, reuse=True)
    with tf_util.use_store(store):
        p_train, _, _ = MODEL.get_model(a, b, True, **kw)
        with torch.no_grad():
            p_eval, _, _ = MODEL.get_model(a, b, False, **kw)
    assert p_train["pred_listAB"].requires_grad and not p_eval["pred_listAB"].requires_grad
    assert torch.equal(p_train["pred_listAB"].detach(), p_eval["pred_listAB"])
