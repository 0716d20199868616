"""Two NCCL ranks on two GPUs: the data-parallel training step the scaling runs time (tensor-core backward, gradients
written into the trainer's flat buffer, two all-reduce buckets overlapping the backward, one Adam launch) against the
single-GPU full-batch step.  Reference semantics: equal towers + average_gradients,
train_multi_gpu_pc_compare_dist.py:241-251, 936-974.  Skipped when the box has one GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dpdist_b200 import synthetic, tf_util, train

pytestmark = pytest.mark.gpu
STEPS = 3


def _batch(pairs):
    pcA, pcB, labels = synthetic.uniform_batch(41, pairs, 64, outside_frac=0.05)
    return pcA, pcB, labels * 3.0


def _steps(dev, rank, world, pairs, graph):
    pcA, pcB, labels = _batch(pairs)
    tr = train.DPDistTrainer(dev, seed=7, cuda_graph=graph)
    mine = [torch.tensor(np.ascontiguousarray(train.shard(x, rank, world)), device=dev) for x in (pcA, pcB, labels)]
    n = STEPS + (4 if graph else 0)                 # graph mode: 3 eager warm-up steps, capture, replays
    losses = [tr.step(*mine) for _ in range(n)]
    torch.cuda.synchronize(dev)
    return tr, torch.stack(losses).cpu()


def _worker(rank, world, port, out_dir, pairs, graph):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    import datetime
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=120))
    tr, losses = _steps(dev, rank, world, pairs, graph)
    ok = tr.ranks_consistent()
    torch.save({"w": {n: v.detach().cpu() for n, v in tr.store.vars.items()}, "losses": losses, "consistent": ok,
                "graph": tr._graph is not None}, os.path.join(out_dir, "r%d.pt" % rank))
    tr.close()                       # captured NCCL kernels must be released before the communicator goes away
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("pairs,graph", [(64, False), (16, True)])
def test_two_rank_nccl_step_equals_single_gpu_full_batch_step(tmp_path, pairs, graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.spawn(_worker, args=(2, port, str(tmp_path), pairs, graph), nprocs=2, join=False)
    import time
    deadline = time.time() + 240
    while not ctx.join(timeout=5):
        if time.time() > deadline:
            for p in ctx.processes:
                p.kill()
            pytest.fail("the two NCCL ranks did not finish within 240 s")
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert r0["consistent"] and r1["consistent"] and r0["graph"] == graph
    tr, losses = _steps(torch.device("cuda", 0), 0, 1, pairs, graph)
    init = tf_util.VariableStore(device="cpu", seed=7)
    # the global loss is the mean of the two towers' losses (equal slices)
    both = (r0["losses"] + r1["losses"]) / 2
    assert torch.allclose(both, losses, rtol=2e-4, atol=1e-6), (both, losses)
    for n, w0 in r0["w"].items():
        assert torch.equal(w0, r1["w"][n]), n                       # bit-identical weights on every rank
        full = tr.store.vars[n].detach().cpu()
        scale = float(full.abs().max())
        # after a few Adam steps every weight has moved by ~steps*lr; a different summation order of the same mean
        # gradient can flip the sign of near-zero gradient entries, so compare against the size of the move
        moved = len(losses) * 1e-4
        bad = ((w0 - full).abs() > 0.05 * moved).double().mean()
        assert float(bad) < 2e-3, (n, float(bad))
        assert float((w0 - full).abs().max()) <= 2.5 * moved, n
