"""DPDist training step: data-parallel towers, gradient averaging, Adam, staircase LR.

Mirrors the DPDist branch of the reference trainer `train_multi_gpu_pc_compare_dist.py`:
  batch assembly            train_one_epoch_3d            :732-778
  towers / batch slicing    :125-126, :237-258
  gradient averaging        average_gradients             :936-974
  optimizer + LR schedule   :209-216, :976-990 (AdamOptimizer defaults, staircase decay, floor 1e-7)
Re-expressed B200-first: one process per GPU, every rank holds the variables in HBM (the reference
pins them to /cpu:0 and averages on the host over PCIe), one NCCL all-reduce(mean) per layer, issued
from the backward pass as soon as that layer's gradients exist, Adam applied redundantly per rank.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, dpdist_and_aue as MODEL, dpdist_util, tf_util

ADAM_BETA1, ADAM_BETA2, ADAM_EPS = 0.9, 0.999, 1e-8     # TF-semantics tf.train.AdamOptimizer defaults


def get_learning_rate(batch, base_lr=0.0001, decay_step=300 * 512, decay_rate=0.5):
    """:976-990: exponential_decay(base_lr, batch, DECAY_STEP, DECAY_RATE, staircase=True), clipped at 1e-7."""
    lr = base_lr * decay_rate ** (int(batch) // int(decay_step))
    return max(lr, 0.0000001)


def assemble_batch(batch_data, batch_label, NUM_POINT):
    """:749-766.  batch_data [bsize, 3*npoints, 3] = surface | close | far (modelnet_dataset.py:136-139),
    batch_label [bsize, 2*npoints] = GT distances of close | far  ->  (pcA, pcB, labels_AB)."""
    H_NUM_POINT = int(NUM_POINT / 2)
    split_off_surface = 0.5
    batch_data = np.split(batch_data, 3, 1)                    # surface, close, far
    batch_surface = np.split(batch_data[0], 2, 1)              # two clouds from the same surface S_A, S_B
    bsize = batch_data[0].shape[0]
    pcA = batch_surface[0][:, :NUM_POINT]
    batch_label = np.split(batch_label, 2, 1)                  # GT distances of close and far points
    q = int(H_NUM_POINT * split_off_surface)
    labels_AB = np.concatenate([np.zeros([bsize, H_NUM_POINT]), batch_label[0][:, :q],
                                batch_label[1][:, q:H_NUM_POINT]], 1)
    batch_off = np.concatenate([batch_data[1][:, :q], batch_data[2][:, q:H_NUM_POINT]], 1)
    pcB = np.concatenate([batch_surface[1][:, :H_NUM_POINT], batch_off], 1)
    return pcA.astype(np.float32), pcB.astype(np.float32), labels_AB.astype(np.float32)


def shard(array, rank, world):
    """tf.slice(x, [i*DEVICE_BATCH_SIZE, ...], [DEVICE_BATCH_SIZE, ...]) (:241-251); batch must divide (:125)."""
    n = array.shape[0]
    if n % world != 0:
        raise AssertionError("BATCH_SIZE % NUM_GPUS != 0 (train_multi_gpu_pc_compare_dist.py:125)")
    per = n // world
    return array[rank * per:(rank + 1) * per]


def average_gradients(grads, group=None):
    """:936-974 as a collective: every rank ends with mean_over_towers(grad).  Works on CUDA (NCCL) and
    CPU (gloo) tensors.  Returns the async work handles (empty if not distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return []
    world = dist.get_world_size(group)
    works = []
    for g in grads:
        if g.is_cuda:
            works.append(dist.all_reduce(g, op=dist.ReduceOp.AVG, group=group, async_op=True))
        else:                                   # gloo has no AVG
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            g.div_(world)
    return works


class DPDistTrainer:
    """One rank of the DPDist trainer.  `step(pcA, pcB, labels_AB)` takes this rank's slice of the global
    batch as CUDA tensors and performs forward, loss_samples (:260-262), backward, gradient averaging and
    the Adam update; returns the local loss_samples as a tensor (no host sync)."""

    def __init__(self, device, base_lr=0.0001, decay_step=300 * 512, decay_rate=0.5, seed=1, store=None,
                 Embedding_Size=512, k=5, sigma3dmfv=0.125, mlp=(1024, 1024, 1024), overlap_allreduce=True,
                 cuda_graph=False):
        self.device = torch.device(device)
        self.store = store if store is not None else tf_util.VariableStore(device=self.device, seed=seed)
        self.base_lr, self.decay_step, self.decay_rate = base_lr, decay_step, decay_rate
        self.kw = dict(bn=0, Embedding_Size=Embedding_Size, k=k, sigma3dmfv=sigma3dmfv, localSNmlp=list(mlp))
        self.batch = 0                      # the 'batch' global step variable (:201)
        self.m, self.v = {}, {}
        self.overlap = overlap_allreduce
        self._works = []
        # cuda_graph=True (single process): after three eager steps the whole step (forward, backward, Adam) is captured
        # once per input shape and replayed; only the Adam rate scalar and the inputs are rewritten per step.  For the
        # reference's launch-bound batch of 16 pairs.
        self.cuda_graph = bool(cuda_graph)
        self._graph = None
        self._eager_steps = 0
        # warm-up steps and the capture share one side stream (the pattern torch documents for whole-step capture): autograd
        # remembers the stream a parameter's gradient accumulator was created on, and a capture must never make the legacy
        # default stream wait on it
        self._side = torch.cuda.Stream(device=self.device) if self.cuda_graph else None

    def variables(self):
        return self.store.trainable_variables("pc_compare")

    def _on_grad_ready(self, layer, grads):
        # per-layer all-reduce on NCCL's stream, overlapping the remaining backward kernels
        self._works += average_gradients(grads)

    def _graph_step(self, pcA, pcB, labels_AB):
        lib = _lib.load()
        gs = self._graph
        if gs is None or gs["shape"] != tuple(pcA.shape):
            params = self.variables()
            gs = {"shape": tuple(pcA.shape), "a": pcA.clone(), "b": pcB.clone(), "l": labels_AB.clone(),
                  "lr_t": torch.zeros(1, device=self.device), "params": params}
            for p in params:
                if id(p) not in self.m:
                    self.m[id(p)] = torch.zeros_like(p)
                    self.v[id(p)] = torch.zeros_like(p)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=self._side):
                tf_util.clear_collections()
                with tf_util.use_store(self.store):
                    pred, end_points, _ = MODEL.get_model(gs["a"], gs["b"], True, **self.kw)
                    MODEL.get_loss(pred, end_points, gs["l"])
                loss = tf_util.get_collection("loss_samples")[-1]
                # autograd.grad, not backward(): no AccumulateGrad nodes, whose streams (the default stream of the eager
                # steps) would make the legacy stream wait on the capturing one and invalidate the capture
                grads = torch.autograd.grad(loss, params)
                gs["grads"] = grads
                stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                with torch.no_grad():
                    for p, g in zip(params, grads):
                        rc = lib.dpd_adam_step_dev(p.data_ptr(), g.data_ptr(), self.m[id(p)].data_ptr(),
                                                   self.v[id(p)].data_ptr(), p.numel(), gs["lr_t"].data_ptr(),
                                                   ADAM_BETA1, ADAM_BETA2, ADAM_EPS, stream)
                        _lib.check(rc, "dpd_adam_step_dev")
                        p.add_(0)
                gs["loss"] = loss.detach()
            gs["graph"] = graph
            self._graph = gs
        self.batch += 1
        lr = get_learning_rate(self.batch - 1, self.base_lr, self.decay_step, self.decay_rate)
        # pageable source: the runtime stages it before returning, so the host may run ahead of the device safely
        gs["lr_t"].copy_(torch.tensor([lib.dpd_adam_lr_t(lr, ADAM_BETA1, ADAM_BETA2, self.batch)]))
        gs["a"].copy_(pcA, non_blocking=True)
        gs["b"].copy_(pcB, non_blocking=True)
        gs["l"].copy_(labels_AB, non_blocking=True)
        gs["graph"].replay()
        return gs["loss"].clone()        # the graph's own output buffer is overwritten by the next replay

    def step(self, pcA, pcB, labels_AB, add_noise=0):
        distributed_now = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if self.cuda_graph and not distributed_now and not torch.is_tensor(add_noise) and add_noise == 0:
            if self._eager_steps >= 3:
                return self._graph_step(pcA, pcB, labels_AB)
            self._eager_steps += 1
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                loss = self._eager_step(pcA, pcB, labels_AB, add_noise)
            cur.wait_stream(self._side)
            return loss
        return self._eager_step(pcA, pcB, labels_AB, add_noise)

    def _eager_step(self, pcA, pcB, labels_AB, add_noise=0):
        lib = _lib.load()
        tf_util.clear_collections()
        with tf_util.use_store(self.store):
            pred, end_points, _ = MODEL.get_model(pcA, pcB, True, add_noise=add_noise, **self.kw)
            MODEL.get_loss(pred, end_points, labels_AB)
        loss = tf_util.get_collection("loss_samples")[-1]      # total_loss_samples (:262-263)
        params = self.variables()
        for p in params:
            p.grad = None
        self._works = []
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        dpdist_util.GRAD_READY_HOOK = self._on_grad_ready if (self.overlap and distributed) else None
        try:
            loss.backward()
        finally:
            dpdist_util.GRAD_READY_HOOK = None
        grads = [p.grad for p in params]
        if distributed and not self.overlap:
            self._works = average_gradients(grads)
        for w in self._works:
            w.wait()
        self.batch += 1
        lr = get_learning_rate(self.batch - 1, self.base_lr, self.decay_step, self.decay_rate)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        with torch.no_grad():
            for p, g in zip(params, grads):
                key = id(p)
                if key not in self.m:
                    self.m[key] = torch.zeros_like(p)
                    self.v[key] = torch.zeros_like(p)
                rc = lib.dpd_adam_step(p.data_ptr(), g.data_ptr(), self.m[key].data_ptr(), self.v[key].data_ptr(),
                                       p.numel(), lr, ADAM_BETA1, ADAM_BETA2, ADAM_EPS, self.batch, stream)
                _lib.check(rc, "dpd_adam_step")
                p.add_(0)      # bump the tensor version: the packed-weight cache keys on it
        return loss.detach()
