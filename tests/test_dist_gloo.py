"""world_size-2 gloo test (CPU) of the data-parallel logic: batch slicing + gradient averaging give the
single-process full-batch gradient (the semantics of average_gradients with equal towers,
train_multi_gpu_pc_compare_dist.py:241-251, 936-974).  Gradients come from the CPU oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dpdist_b200 import synthetic, train
from oracle import dpdist_oracle as O

CFG = dict(Embedding_Size=64, k=3, sigma3dmfv=0.25)   # G=4: small enough for the literal oracle on CPU


def _grads(pcA, pcB, labels, var):
    v = {k: t.clone().requires_grad_(True) for k, t in var.items()}
    p, _, _ = O.get_model(torch.tensor(pcA), torch.tensor(pcB), v, **CFG)
    loss, _ = O.get_loss(p, {}, torch.tensor(labels))
    loss.backward()
    return [v[k].grad for k in sorted(v)]


def _make(seed):
    pcA, pcB, labels = synthetic.uniform_batch(seed, 4, 16)
    var = O.init_variables(k=3, mlp=(64, 64, 64), seed=2, bias_std=0.05, weight_gain=(40.0, 2.0, 2.0, 1.0), out_bias=1.0)
    return pcA, pcB, labels * 3.0, var


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pcA, pcB, labels, var = _make(3)
    mine = [train.shard(x, rank, world) for x in (pcA, pcB, labels)]
    g = _grads(*mine, var)
    train.average_gradients(g)
    torch.save(g, os.path.join(out_dir, "g%d.pt" % rank))
    dist.destroy_process_group()


def test_sharded_average_equals_full_batch_gradient(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    pcA, pcB, labels, var = _make(3)
    full = _grads(pcA, pcB, labels, var)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    for a, b, f in zip(g0, g1, full):
        assert torch.equal(a, b)                                   # every rank holds the same averaged gradient
        assert float((a - f).abs().max()) <= 1e-6 * max(1e-6, float(f.abs().max())) + 1e-9


def test_shard_requires_divisible_batch_and_lr_schedule():
    x = np.zeros((6, 2))
    assert train.shard(x, 1, 3).shape == (2, 2)
    try:
        train.shard(x, 0, 4)
        raise SystemError("expected an assertion")
    except AssertionError:
        pass
    assert train.get_learning_rate(0) == 1e-4
    assert train.get_learning_rate(300 * 512 - 1) == 1e-4 and train.get_learning_rate(300 * 512) == 5e-5
    assert train.get_learning_rate(10 ** 9) == 1e-7                # tf.maximum(lr, 1e-7), :987


def test_assemble_batch_follows_the_reference():
    pts, lab = synthetic.dataset_batch(1, 3, 64)                    # [3, 384, 3], [3, 256]
    pcA, pcB, lab_ab = train.assemble_batch(pts, lab, 64)
    assert pcA.shape == (3, 64, 3) and pcB.shape == (3, 64, 3) and lab_ab.shape == (3, 64)
    assert np.array_equal(pcA, pts[:, :64])                         # S_A = first half of the surface points
    assert np.array_equal(pcB[:, :32], pts[:, 64:96])               # S_B[:32]
    assert np.array_equal(pcB[:, 32:48], pts[:, 128:144])           # close[:16]
    assert np.array_equal(pcB[:, 48:64], pts[:, 256 + 16:256 + 32]) # far[16:32]
    assert np.all(lab_ab[:, :32] == 0)
    assert np.array_equal(lab_ab[:, 32:48], lab[:, :16]) and np.array_equal(lab_ab[:, 48:], lab[:, 128 + 16:128 + 32])
