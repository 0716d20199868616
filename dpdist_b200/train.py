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


def get_bn_decay(batch, decay_step=300 * 512, init_decay=0.5, decay_rate=0.5, clip=0.99):
    """:172-175, :992-1000: min(BN_DECAY_CLIP, 1 - BN_INIT_DECAY * BN_DECAY_DECAY_RATE^floor(batch / DECAY_STEP))."""
    return min(clip, 1.0 - init_decay * decay_rate ** (int(batch) // int(decay_step)))


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


class FlatState:
    """The trainable variables of one tower re-homed into ONE flat fp32 buffer, with gradient and Adam-moment buffers of
    the same layout.  The parameters stay the same `torch.nn.Parameter` objects (the VariableStore, checkpoints and the
    TF names are untouched); only their storage moves.

    Layout = the order in which the backward pass finishes the layers: bucket 0 = every variable of layers 4, 3, 2 (and
    any variable that is not a layer-1 conv variable), bucket 1 = mapper_conv1/{weights,biases}.  So the data-parallel
    exchange (`average_gradients`, :936-974) is at most two all-reduces per step, the first one in flight while the
    largest product (dW1, 10.25 MB of the 18.67 MB) is still being computed, and Adam is one launch over the buffer.
    Every segment starts on a 256-byte boundary; the padding stays zero (zero gradient -> zero Adam update)."""
    ALIGN = 64      # floats

    def __init__(self, named_params):
        names = list(named_params.keys())
        last = [n for n in names if "/mapper_conv1/" in n and not "/bn/" in n]
        first = [n for n in reversed(names) if n not in last]       # layer 4 first, as the backward produces them
        self.order = first + last
        dev = named_params[names[0]].device
        offs, off = {}, 0
        for n in self.order:
            offs[n] = off
            off += -(-named_params[n].numel() // self.ALIGN) * self.ALIGN
            if n == first[-1]:
                self.split = off
        self.total = off
        self.param = torch.zeros(off, device=dev, dtype=torch.float32)
        self.grad = torch.zeros_like(self.param)
        self.m = torch.zeros_like(self.param)
        self.v = torch.zeros_like(self.param)
        self.grad_view, self.params = {}, {}
        with torch.no_grad():
            for n in self.order:
                p = named_params[n]
                view = self.param[offs[n]:offs[n] + p.numel()].view(p.shape)
                view.copy_(p.detach())
                p.data = view                                       # same Parameter object, storage inside the flat buffer
                self.grad_view[n] = self.grad[offs[n]:offs[n] + p.numel()].view(p.shape)
                self.params[n] = p
        self.offs = offs
        self.accum = None           # gradient accumulator of micro-batched steps, allocated on first use
        self.buckets = [(0, self.split), (self.split, self.total)] if last and first else [(0, self.total)]

    def offset(self, name):
        return self.offs[name]


class _GradSink:
    """What the head's backward writes into when a trainer is driving it: gradient buffers it does not allocate itself,
    and a callback per finished layer.  Travels on the VariableStore (`store.grad_sink`), is captured by the autograd
    node at forward time and used by its backward - no process-wide state, several trainers can coexist."""

    def __init__(self, flat, head_names, on_ready):
        self.views = [flat.grad_view[n] for n in head_names]       # [w1, b1, w2, b2, w3, b3, w4, b4]
        self.on_ready = on_ready

    def buffers(self, shapes):
        if [tuple(v.shape) for v in self.views] != [tuple(s) for s in shapes]:
            raise ValueError("gradient sink does not match the head's variables")
        return self.views

    def ready(self, layer):
        self.on_ready(layer)


class DPDistTrainer:
    """One rank of the DPDist trainer.  `step(pcA, pcB, labels_AB)` takes this rank's slice of the global
    batch as CUDA tensors and performs forward, loss_samples (:260-262), backward, gradient averaging and
    the Adam update; returns the local loss_samples as a tensor (no host sync)."""

    HEAD = ["pc_compare/dpdist_local/mapper_conv%d/%s" % (l, w) for l in (1, 2, 3, 4) for w in ("weights", "biases")]

    def __init__(self, device, base_lr=0.0001, decay_step=300 * 512, decay_rate=0.5, seed=1, store=None,
                 Embedding_Size=512, k=5, sigma3dmfv=0.125, mlp=(1024, 1024, 1024), overlap_allreduce=True,
                 cuda_graph=False, group=None, bn=0, conv_version=1):
        self.device = torch.device(device)
        self.store = store if store is not None else tf_util.VariableStore(device=self.device, seed=seed)
        self.base_lr, self.decay_step, self.decay_rate = base_lr, decay_step, decay_rate
        # bn truthy = the reference's --BN 1: batch norm after every conv, batch statistics per tower (not synchronised
        # across towers, utils/tf_util.py:573-577), gamma / beta trained with the other variables, bn_decay of :992-1000
        self.kw = dict(bn=int(bn), Embedding_Size=Embedding_Size, k=k, sigma3dmfv=sigma3dmfv, localSNmlp=list(mlp))
        if int(conv_version) != 1:          # --implicit_net_type 3: the 3-D CNN head (utils/dpdist_util.py:640-687)
            self.kw["conv_version"] = int(conv_version)
        self.batch = 0                      # the 'batch' global step variable (:201)
        self.overlap = overlap_allreduce
        self.group = group
        self.flat = None
        self._works, self._reduced = [], set()
        self._lr_t = torch.zeros(1, device=self.device)
        # cuda_graph=True: after three eager steps the whole step (forward, backward, gradient all-reduce, Adam) is
        # captured once per input shape and replayed; only the Adam rate scalar and the inputs are rewritten per step.
        # For the reference's launch-bound batch of 16 pairs, on one GPU or sharded over several (NCCL inside the graph).
        self.cuda_graph = bool(cuda_graph)
        self._graph = None
        self._eager_steps = 0
        # warm-up steps and the capture share one side stream (the pattern torch documents for whole-step capture): autograd
        # remembers the stream a parameter's gradient accumulator was created on, and a capture must never make the legacy
        # default stream wait on it
        self._side = torch.cuda.Stream(device=self.device) if self.cuda_graph else None

    # ---- state -------------------------------------------------------------------------------------------------------
    def variables(self):
        return self.store.trainable_variables("pc_compare")

    def _named(self):
        return {n: v for n, v in self.store.vars.items() if v.requires_grad and n.startswith("pc_compare")}

    def _world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    def _create_variables(self):
        """tf.get_variable at graph-build time: the head's variables exist before the first step."""
        mlp = self.kw["localSNmlp"]
        G, _ = dpdist_util._fv_grid(self.kw["Embedding_Size"], 3)
        with tf_util.use_store(self.store), tf_util.variable_scope('pc_compare'):
            if self.kw.get("conv_version", 1) == 3:
                dpdist_util._cv3_variables(dpdist_util.FV_CHANNELS[True], self.kw["k"], 3, mlp, None)
                return
            dpdist_util._head_variables(dpdist_util.FV_CHANNELS[True] * self.kw["k"] ** 3, 3, mlp, None, bn=bool(self.kw["bn"]))

    def _ensure_flat(self):
        named = self._named()
        if not named or (self.kw.get("conv_version", 1) == 1 and not all(n in named for n in self.HEAD)):
            self._create_variables()
            named = self._named()
        f = self.flat
        if f is None or set(named) != set(f.params) or any(named[n] is not f.params[n] for n in named):
            self.flat = FlatState(named)
            if f is not None:                   # variables were added (e.g. batch norm): keep the moments of the old ones
                for n, p in f.params.items():
                    if named.get(n) is p:
                        for new, old in ((self.flat.m, f.m), (self.flat.v, f.v)):
                            new[self.flat.offset(n):][:p.numel()] = old[f.offset(n):][:p.numel()]
            self._weights_changed()
        return self.flat

    @property
    def m(self):
        """{id(param): first-moment view} (for callers that inspect the optimizer state)."""
        f = self._ensure_flat()
        return {id(p): f.m[f.offset(n):][:p.numel()].view(p.shape) for n, p in f.params.items()}

    @property
    def v(self):
        f = self._ensure_flat()
        return {id(p): f.v[f.offset(n):][:p.numel()].view(p.shape) for n, p in f.params.items()}

    # ---- one step ----------------------------------------------------------------------------------------------------
    def _reduce_bucket(self, i):
        lo, hi = self.flat.buckets[i]
        self._works += average_gradients([self.flat.grad[lo:hi]], self.group)
        self._reduced.add(i)

    def _on_layer_ready(self, layer):
        # called from the head's backward as soon as a layer's gradients are complete (4, 3, 2, 1): the first bucket goes
        # out while dW1 is still being computed, on NCCL's own stream
        if not self.overlap or self._world() == 1 or len(self.flat.buckets) < 2:
            return
        if layer == 2:
            self._reduce_bucket(0)
        elif layer == 1:
            self._reduce_bucket(1)

    MAX_ROWS = 1 << 18      # rows (2 * pairs * NP) one DPD_HEAD_TRAIN call keeps activations for (head.cu MAX_CHUNK_ROWS)

    def _forward_backward_update(self, pcA, pcB, labels_AB, add_noise):
        flat = self._ensure_flat()
        pairs, NP = pcA.shape[0], pcB.shape[1]
        per = max(1, self.MAX_ROWS // (2 * NP))
        if pairs > per:
            # The training forward keeps every activation of its rows, so it is bounded to one row chunk.  Larger tower
            # batches (config E: N = NP = 512) run as micro-batches whose gradients are accumulated with the weights
            # n_i / n: the loss is a mean over the batch, so the sum is the gradient of the whole batch.
            if flat.accum is None:
                flat.accum = torch.zeros_like(flat.grad)
            flat.accum.zero_()
            total = None
            for lo in range(0, pairs, per):
                hi = min(pairs, lo + per)
                noise = add_noise[lo:hi] if torch.is_tensor(add_noise) else add_noise
                loss = self._forward_backward(flat, pcA[lo:hi], pcB[lo:hi], labels_AB[lo:hi], noise, reduce=False)
                w = (hi - lo) / pairs
                flat.accum.add_(flat.grad, alpha=w)
                total = loss * w if total is None else total + loss * w
            flat.grad.copy_(flat.accum)
            loss = total
        else:
            loss = self._forward_backward(flat, pcA, pcB, labels_AB, add_noise, reduce=True)
        if self._world() > 1:
            for i in range(len(flat.buckets)):
                if i not in self._reduced:
                    self._reduce_bucket(i)
            for w in self._works:
                w.wait()                                            # stream-ordered: the compute stream waits, not the host
        self._works = []
        self._adam(flat)
        return loss.detach()

    def _forward_backward(self, flat, pcA, pcB, labels_AB, add_noise, reduce):
        """Forward, loss_samples and the gradients of every trainable variable into flat.grad.  `reduce`: let the head's
        backward start the all-reduce of a bucket as soon as its layers are done."""
        tf_util.clear_collections()
        head_ok = all(n in flat.grad_view for n in self.HEAD)
        self.store.grad_sink = _GradSink(flat, self.HEAD, self._on_layer_ready if reduce else (lambda layer: None)) if head_ok else None
        self._works, self._reduced = [], set()
        try:
            kw = dict(self.kw, bn_decay=get_bn_decay(max(self.batch - 1, 0), self.decay_step)) if self.kw["bn"] else self.kw
            with tf_util.use_store(self.store):
                pred, end_points, _ = MODEL.get_model(pcA, pcB, True, add_noise=add_noise, **kw)
                MODEL.get_loss(pred, end_points, labels_AB)
            loss = tf_util.get_collection("loss_samples")[-1]      # total_loss_samples (:262-263)
            names = list(flat.params.keys())
            # autograd.grad, not backward(): no AccumulateGrad nodes (they would clone the buffers the all-reduce is
            # running on, and their streams would invalidate a graph capture)
            grads = torch.autograd.grad(loss, [flat.params[n] for n in names], allow_unused=True)
        finally:
            self.store.grad_sink = None
        with torch.no_grad():
            for n, g in zip(names, grads):
                view = flat.grad_view[n]
                if g is None:
                    view.zero_()
                elif g.data_ptr() != view.data_ptr():               # a variable outside the head's sink (batch norm, ...)
                    if self._reduced:
                        raise RuntimeError("gradient of %s was not written into the trainer's buffer" % n)
                    view.copy_(g)
        return loss

    def _adam(self, flat):
        """One TF-semantics Adam launch over the whole flat buffer; the bias-corrected rate is read from device memory."""
        lib = _lib.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = lib.dpd_adam_step_dev(flat.param.data_ptr(), flat.grad.data_ptr(), flat.m.data_ptr(), flat.v.data_ptr(),
                                   flat.total, self._lr_t.data_ptr(), ADAM_BETA1, ADAM_BETA2, ADAM_EPS, stream)
        _lib.check(rc, "dpd_adam_step_dev")

    def _set_rate(self):
        self.batch += 1
        lr = get_learning_rate(self.batch - 1, self.base_lr, self.decay_step, self.decay_rate)
        # pageable source: the runtime stages it before returning, so the host may run ahead of the device safely
        self._lr_t.copy_(torch.tensor([_lib.load().dpd_adam_lr_t(lr, ADAM_BETA1, ADAM_BETA2, self.batch)]))

    def _weights_changed(self):
        # the packed-weight caches key on this counter (the flat Adam launch does not touch the tensors' version counters)
        self.store.weights_generation = getattr(self.store, "weights_generation", 0) + 1

    def _graph_step(self, pcA, pcB, labels_AB):
        gs = self._graph
        if gs is None or gs["shape"] != tuple(pcA.shape):
            self._ensure_flat()
            gs = {"shape": tuple(pcA.shape), "a": pcA.clone(), "b": pcB.clone(), "l": labels_AB.clone()}
            graph = torch.cuda.CUDAGraph()
            self._weights_changed()              # the capture starts with a weight re-pack, so every replay does
            # distributed: NCCL's watchdog thread polls events while this thread captures; only this thread's calls may
            # invalidate the capture
            mode = "thread_local" if self._world() > 1 else "global"
            with torch.cuda.graph(graph, stream=self._side, capture_error_mode=mode):
                gs["loss"] = self._forward_backward_update(gs["a"], gs["b"], gs["l"], 0)
            gs["graph"] = graph
            self._graph = gs
        self._set_rate()
        gs["a"].copy_(pcA, non_blocking=True)
        gs["b"].copy_(pcB, non_blocking=True)
        gs["l"].copy_(labels_AB, non_blocking=True)
        gs["graph"].replay()
        self._weights_changed()
        return gs["loss"].clone()        # the graph's own output buffer is overwritten by the next replay

    def step(self, pcA, pcB, labels_AB, add_noise=0):
        if self.cuda_graph and not self.kw["bn"] and "conv_version" not in self.kw and not torch.is_tensor(add_noise) and add_noise == 0:
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
        self._set_rate()
        loss = self._forward_backward_update(pcA, pcB, labels_AB, add_noise)
        self._weights_changed()
        return loss

    def close(self):
        """Drop the captured graph.  Call before dist.destroy_process_group(): a communicator cannot be torn down while a
        live CUDA graph still holds its captured collective kernels (the destroy call then never returns)."""
        torch.cuda.synchronize(self.device)
        self._graph = None
        import gc
        gc.collect()
        torch.cuda.synchronize(self.device)

    def weights_checksum(self):
        """Order-independent 64-bit checksum of the parameter bits (sum of the words as integers): equal on every rank
        iff the ranks hold bit-identical weights."""
        f = self._ensure_flat()
        return f.param.view(torch.int32).to(torch.int64).sum()

    def ranks_consistent(self):
        """True iff every rank of the group holds bit-identical weights (one all-reduce of the checksum)."""
        if self._world() == 1:
            return True
        c = self.weights_checksum().reshape(1)
        lo, hi = c.clone(), c.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        return bool((lo == hi).item())
