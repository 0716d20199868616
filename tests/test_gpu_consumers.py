"""GPU tests of the consumers' side of the path: the name-addressed DPDistLoss module,
TF-named checkpoints, and one PCRNet-ours training step whose gradients come through the DPDist loss."""
import numpy as np
import pytest
import torch

from dpdist_b200 import dpdist_and_aue as MODEL, pcrnet_ours, synthetic, tf_util
from dpdist_b200.dpdist_loss import DPDistLoss
from oracle import dpdist_oracle as O
from tolerances import assert_close, assert_grad_close, assert_out_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _loss_module(var):
    m = DPDistLoss(num_point=64, device=DEV, seed=1)
    m.load_tf_state_dict({n: t.numpy() for n, t in var.items()})
    return m


def test_graph_names_run_and_checkpoints(tmp_path):
    var = O.unit_scale_variables(4)
    m = _loss_module(var)
    assert sorted(m.tf_variables()) == sorted(var)                                 # pc_compare/dpdist_local/mapper_conv*/...
    pcA, pcB, _ = synthetic.uniform_batch(3, 4, 64)
    with O.tf_cpu_numerics():
        want, _, _ = O.get_model(torch.tensor(pcA), torch.tensor(pcB), var)
    o1, o2 = m.run(["g1/pc_compare/output1:0", "pc_compare/output2:0"],
                   {"input1:0": pcA, "g1/input2:0": pcB, "Placeholder:0": False, "add_noise:0": np.zeros_like(pcA)})
    assert o1.shape == (4, 64, 1, 3) and o2.shape == (4, 64, 1, 3)
    assert_out_close(o1, want["pred_listAB"], "output1")
    assert_out_close(o2, want["pred_listBA"], "output2")
    with pytest.raises(KeyError):
        m.run("pc_compare/output3:0", {"input1": pcA, "input2": pcB})
    with pytest.raises(KeyError):
        m.run("pc_compare/output1:0", {"input1": pcA, "input2": pcB, "input3": pcB})
    with pytest.raises(ValueError):
        m.run("pc_compare/output1:0", {"input1": pcA[:, :32], "input2": pcB[:, :32]})    # static shapes, like the graph
    # save as a TF V2 checkpoint, restore into a fresh module (different random init) -> identical outputs
    prefix = m.save(str(tmp_path / "model.ckpt"))
    m2 = DPDistLoss(num_point=64, device=DEV, seed=99).restore(prefix)
    p1 = m2.run("pc_compare/output1:0", {"input1": pcA, "input2": pcB})
    assert torch.equal(p1, o1)
    m3 = DPDistLoss(num_point=64, device=DEV, seed=98).restore(m.save(str(tmp_path / "model.npz")))
    assert torch.equal(m3.run("pc_compare/output1:0", {"input1": pcA, "input2": pcB}), o1)
    # add_noise is added to input1 before the encoder only (models/dpdist_and_aue.py:45)
    noise = np.full_like(pcA, 0.01)
    n1 = m.run("pc_compare/output1:0", {"input1": pcA, "input2": pcB, "add_noise": noise})
    with O.tf_cpu_numerics():
        wn, _, _ = O.get_model(torch.tensor(pcA), torch.tensor(pcB), var, add_noise=torch.tensor(noise))
    assert_out_close(n1, wn["pred_listAB"], "output1 with add_noise")


def test_loss_gradient_matches_oracle_and_leaves_the_variables_alone():
    var = O.unit_scale_variables(4)
    m = _loss_module(var)
    pcA, pcB, _ = synthetic.chair_batch(5, 3, 64)
    a = torch.tensor(pcA, dtype=torch.float64, requires_grad=True)
    p, _, _ = O.get_model(a, torch.tensor(pcB, dtype=torch.float64), {n: t.double() for n, t in var.items()})
    want = (p["pred_listAB"][..., 0].mean() + p["pred_listBA"][..., 0].mean()) / 2
    want.backward()
    x = torch.tensor(pcA, device=DEV, requires_grad=True)
    loss = m.loss(x, torch.tensor(pcB, device=DEV))
    loss.backward()
    assert abs(loss.item() - want.item()) <= 1e-5 * max(1.0, abs(want.item()))
    assert_grad_close(x.grad, a.grad.numpy(), "d loss / d input1 (chairs)", frac=0.95)
    assert all(p.grad is None for p in m.parameters())


def test_pcrnet_step_trains_only_the_pose_network():
    var = O.unit_scale_variables(4)
    m = _loss_module(var)
    before = {n: v.detach().clone() for n, v in m.tf_variables().items()}
    tr = pcrnet_ours.IterativePCRNetOurs(m, max_loops=3, learning_rate=1e-3, device=DEV, seed=0)
    rng = np.random.default_rng(0)
    tpl = pcrnet_ours.synthetic_templates(16, 64, seed=1)
    src = pcrnet_ours.apply_transformation(tpl, pcrnet_ours.generate_poses(16, rng))
    w0 = [p.detach().clone() for p in tr.net.parameters()]
    loss, T, moved = tr.train_step(torch.tensor(src, device=DEV), torch.tensor(tpl, device=DEV))
    assert torch.isfinite(loss) and T.shape == (16, 4, 4) and moved.shape == (16, 64, 3)
    assert any(not torch.equal(a, b) for a, b in zip(w0, tr.net.parameters()))      # the pose network moved
    for n, v in m.tf_variables().items():                                          # DPDist stayed frozen
        assert torch.equal(v, before[n]), n
    # the accumulated transform reproduces the moved cloud: moved = R src + t
    chk = torch.einsum("bij,bnj->bni", T[:, :3, :3], torch.tensor(src, device=DEV)) + T[:, None, :3, 3]
    assert_close(chk, moved, 1e-4, 1e-5, "accumulated transformation")
    T2, moved2 = tr.register(torch.tensor(src, device=DEV), torch.tensor(tpl, device=DEV))
    assert torch.isfinite(moved2).all()


def test_pcrnet_graph_step_matches_eager():
    """The captured PCRNet-ours batch step replays to the same pose-network weights as the eager one (dropout off)."""
    var = O.unit_scale_variables(4)
    rng = np.random.default_rng(0)
    tpl = pcrnet_ours.synthetic_templates(16, 64, seed=1)
    src = pcrnet_ours.apply_transformation(tpl, pcrnet_ours.generate_poses(16, rng))
    s, t = torch.tensor(src, device=DEV), torch.tensor(tpl, device=DEV)
    res = []
    for graph in (False, True):
        m = _loss_module(var)
        tr = pcrnet_ours.IterativePCRNetOurs(m, max_loops=3, learning_rate=1e-3, device=DEV, seed=0, cuda_graph=graph)
        tr.net.dp4.p = 0.0                                   # no dropout: the two runs must see the same network
        # the same optimizer arithmetic in both runs (capturable Adam keeps its step count on the device)
        tr.opt = torch.optim.Adam(tr.net.parameters(), lr=1e-3, eps=1e-8, capturable=True)
        losses = [tr.train_step(s, t)[0].clone() for _ in range(6)]
        torch.cuda.synchronize()
        res.append((torch.stack(losses).cpu(), [p.detach().cpu().clone() for p in tr.net.parameters()]))
    assert tr._graph is not None
    assert torch.allclose(res[0][0], res[1][0], rtol=1e-4, atol=1e-7), (res[0][0], res[1][0])
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), float((a - b).abs().max())
