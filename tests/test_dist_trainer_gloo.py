"""world_size-2 gloo test (CPU) of the trainer's data-parallel step THROUGH its hook path: the head's backward writes
into the trainer's flat gradient buffer and reports layers 4..1, the trainer all-reduces two buckets and applies Adam to
the reduced buffer.  The CUDA head is replaced by the CPU oracle wrapped in an autograd node that follows the same
sink protocol as dpdist_util._head_backward, and the Adam launch by a torch restatement of optim.cu; everything else
(FlatState, _GradSink, buckets, averaging, schedule) is the product code.  Semantics under test:
train_multi_gpu_pc_compare_dist.py:241-251 (slicing), :936-974 (average_gradients), :216 (Adam)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dpdist_b200 import dpdist_util, synthetic, tf_util, train
from oracle import dpdist_oracle as O

CFG = dict(Embedding_Size=64, k=3, sigma3dmfv=0.25)   # G=4: small enough for the literal oracle on CPU
MLP = (64, 64, 64)
NAMES = train.DPDistTrainer.HEAD


class _OracleHead(torch.autograd.Function):
    """CPU stand-in for dpdist_util._ModelTrainFunction with the same gradient-sink protocol."""

    @staticmethod
    def forward(ctx, pcA, pcB, *weights):
        ctx.sink = dpdist_util._grad_sink()
        ctx.inputs = (pcA, pcB)
        ctx.save_for_backward(*weights)
        with torch.no_grad():
            var = {n: w for n, w in zip(NAMES, weights)}
            p, _, _ = O.get_model(pcA, pcB, var, **CFG)
        return p["pred_listAB"], p["pred_listBA"]

    @staticmethod
    def backward(ctx, g_ab, g_ba):
        weights = ctx.saved_tensors
        with torch.enable_grad():
            var = {n: w.detach().clone().requires_grad_(True) for n, w in zip(NAMES, weights)}
            p, _, _ = O.get_model(*ctx.inputs, var, **CFG)
            tot = (p["pred_listAB"] * g_ab).sum() + (p["pred_listBA"] * g_ba).sum()
            gs = torch.autograd.grad(tot, [var[n] for n in NAMES])
        bufs = ctx.sink.buffers([tuple(w.shape) for w in weights])
        for layer in (4, 3, 2, 1):
            for i in (2 * layer - 2, 2 * layer - 1):
                bufs[i].copy_(gs[i])
            ctx.sink.ready(layer)
        return (None, None) + tuple(bufs)


class _CpuTrainer(train.DPDistTrainer):
    def _create_variables(self):
        var = O.init_variables(k=3, mlp=MLP, seed=2, bias_std=0.05, weight_gain=(40.0, 2.0, 2.0, 1.0), out_bias=1.0)
        self.store.load_state_dict(var, strict=False)

    def _adam(self, flat):      # optim.cu:adam_dev_kernel
        b1, b2, eps = train.ADAM_BETA1, train.ADAM_BETA2, train.ADAM_EPS
        with torch.no_grad():
            flat.m.mul_(b1).add_(flat.grad, alpha=1 - b1)
            flat.v.mul_(b2).addcmul_(flat.grad, flat.grad, value=1 - b2)
            flat.param.sub_(self._lr_t * flat.m / (flat.v.sqrt() + eps))


def _get_model(pcA, pcB, is_training, **kw):
    store = tf_util.default_store()
    ab, ba = _OracleHead.apply(pcA, pcB, *[store.vars[n] for n in NAMES])
    return {"pred_listAB": ab, "pred_listBA": ba}, {}, {}


def _run(rank, world, steps=3):
    train.MODEL.get_model = _get_model
    pcA, pcB, labels = synthetic.uniform_batch(3, 4, 16)
    labels = labels * 3.0
    tr = _CpuTrainer("cpu", store=tf_util.VariableStore(device="cpu"), mlp=MLP, **CFG)
    mine = [torch.tensor(train.shard(x, rank, world)) for x in (pcA, pcB, labels)]
    losses = [float(tr.step(*mine)) for _ in range(steps)]
    return tr, losses


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tr, losses = _run(rank, world)
    assert tr.ranks_consistent()
    torch.save({n: v.detach().clone() for n, v in tr.store.vars.items()}, os.path.join(out_dir, "w%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_hooked_step_equals_full_batch_step(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    saved = train.MODEL.get_model
    try:
        tr, losses = _run(0, 1)
    finally:
        train.MODEL.get_model = saved
    assert np.isfinite(losses).all() and tr.batch == 3
    w0, w1 = torch.load(tmp_path / "w0.pt"), torch.load(tmp_path / "w1.pt")
    init = O.init_variables(k=3, mlp=MLP, seed=2, bias_std=0.05, weight_gain=(40.0, 2.0, 2.0, 1.0), out_bias=1.0)
    for n in NAMES:
        assert torch.equal(w0[n], w1[n]), n                                   # every rank applied the same averaged gradient
        full = tr.store.vars[n].detach()
        moved = float((full - init[n]).abs().max())
        assert moved > 0, n                                                   # the step did something
        # same mean gradient up to fp32 summation order; Adam's first steps move every weight by ~lr, so compare moves
        assert float((w0[n] - full).abs().max()) <= 0.05 * moved + 1e-9, n


def test_flat_state_layout_and_sink():
    store = tf_util.VariableStore(device="cpu")
    var = O.init_variables(k=3, mlp=MLP, seed=1)
    store.load_state_dict(var, strict=False)
    named = {n: v for n, v in store.vars.items()}
    before = {n: v.detach().clone() for n, v in named.items()}
    flat = train.FlatState(named)
    assert len(flat.buckets) == 2 and flat.buckets[0][1] == flat.buckets[1][0] and flat.buckets[1][1] == flat.total
    lo, hi = flat.buckets[1]
    for n, p in named.items():
        assert torch.equal(p.detach(), before[n]) and store.vars[n] is p      # same Parameter objects, same values
        off = flat.offset(n)
        assert off % train.FlatState.ALIGN == 0
        assert (lo <= off < hi) == ("/mapper_conv1/" in n)                    # layer 1 alone in the last bucket
        p.data.add_(1.0)
        assert torch.equal(flat.param[off:off + p.numel()].view(p.shape), p.detach())   # storage really is shared
    sink = train._GradSink(flat, NAMES, lambda layer: None)
    bufs = sink.buffers([tuple(named[n].shape) for n in NAMES])
    bufs[0].fill_(2.0)
    assert float(flat.grad[flat.offset(NAMES[0])]) == 2.0
