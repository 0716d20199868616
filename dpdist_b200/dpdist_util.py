"""Host-side mirror of the reference's op library `utils/dpdist_util.py` for the hot path.

Same function names, argument order and defaults as the reference; torch CUDA tensors in and
out; every op is one or more calls into libdpdist_b200.so (include/dpdist_b200.h).  There is no
CPU implementation here: a CPU tensor or a missing extension raises.

Reference anchors (relative to the reference repo):
  get_3dmfv_tf      utils/dpdist_util.py:22-141
  local_z / _3d     utils/dpdist_util.py:850-854, 911-960
  get_grid_centers  utils/dpdist_util.py:982-992
  DPDist            utils/dpdist_util.py:412-544, 688-700 (conv_version 1, k > 0)
  get_loss          utils/dpdist_util.py:962-980
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import tf_util

FV_CHANNELS = {True: 20, False: 7}


# --------------------------------------------------------------------------------------
# host-side grid tables, built exactly as the reference builds them (numpy fp64 -> fp32)
# --------------------------------------------------------------------------------------
def _fv_grid(n_gaussians, D=3):
    if D != 3:
        raise NotImplementedError("only NUM_DIMS == 3 is on the DPDist hot path")
    grid_size = int(np.ceil(np.power(n_gaussians, 1 / 3)))                    # :41
    if grid_size ** 3 != n_gaussians:
        raise ValueError("n_gaussians=%d is not a perfect cube (the reference's graph fails with a "
                         "shape error at utils/dpdist_util.py:73)" % n_gaussians)
    l = np.linspace(-1, 1, grid_size, False) + (1 / grid_size)                # :42
    return grid_size, np.ascontiguousarray(l.astype(np.float32))              # :50 tf.constant(x, tf.float32)


def get_grid_centers(Embedding_Size, NUM_DIMS=2):
    """utils/dpdist_util.py:982-992 (numpy, as in the reference)."""
    if NUM_DIMS == 2:
        vec_size = int(np.floor(np.sqrt(Embedding_Size)))
    else:
        vec_size = int(np.ceil(np.power(Embedding_Size, 1 / 3)))
    grid_step = 2 / vec_size
    l = np.arange(-1, 1, grid_step) + grid_step / 2
    if NUM_DIMS == 2:
        return np.meshgrid(l, l)
    return np.meshgrid(l, l, l)


def _assign_tables(C):
    """Per-axis centre / interval tables from the [V,3] fp32 centre list C, the reference's way:
    grid_size = |C[0].z - C[1].z| / 2 (:468); lo = C - grid_size, hi = C + grid_size in fp32 (:478-487).
    C[g] = (l[i1], l[i0], l[i2]) for g = i0*G^2 + i1*G + i2, so C[:G, 2] is the axis vector l."""
    cached = getattr(C, "_dpd_tables", None)
    if cached is not None:
        return cached
    Cn = C.detach().cpu().numpy() if torch.is_tensor(C) else np.asarray(C)   # device sync: only for foreign C
    Cn = Cn.astype(np.float32)
    V = Cn.shape[0]
    G = int(np.round(np.power(V, 1 / 3)))
    if G ** 3 != V:
        raise ValueError("C must list G^3 voxel centres")
    l = np.ascontiguousarray(Cn[:G, 2])
    gs = np.float32(np.abs(Cn[0][2] - Cn[1][2]) / np.float32(2))
    lo = np.ascontiguousarray((l - gs).astype(np.float32))
    hi = np.ascontiguousarray((l + gs).astype(np.float32))
    # sanity: C really is the meshgrid('xy') product of l (otherwise the separable search is invalid)
    i = np.arange(V)
    exp = np.stack([l[(i // G) % G], l[i // (G * G)], l[i % G]], -1)
    if not np.array_equal(exp, Cn):
        raise ValueError("C is not the reference's get_grid_centers layout")
    return G, l, lo, hi


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_cuda(t, name):
    if not torch.is_tensor(t) or not t.is_cuda:
        raise _lib.DPDistNativeError("%s must be a CUDA tensor: dpdist_b200 has no CPU path" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32" % name)
    return t.contiguous()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


# --------------------------------------------------------------------------------------
# 3DmFV
# --------------------------------------------------------------------------------------
def _fv_forward(points, G, l, sigma, full_fv, flatten):
    lib = _lib.load()
    B, N, _ = points.shape
    C = FV_CHANNELS[bool(full_fv)]
    V = G ** 3
    out = torch.empty((B, C * V) if flatten else (B, V, C), device=points.device, dtype=torch.float32)
    with torch.cuda.device(points.device):
        rc = lib.dpd_fv_forward(_ptr(points), B, N, G, _lib.fptr(l), float(sigma), int(bool(full_fv)),
                                int(bool(flatten)), _ptr(out), _stream())
    _lib.check(rc, "dpd_fv_forward")
    return out


class _FvFunction(torch.autograd.Function):
    """get_3dmfv_tf with its gradient w.r.t. the points (dpd_fv_backward): the path PCRNet-ours and the AUE
    task differentiate through when DPDist is their loss (iterative_PCRNet_ours.py:229-257)."""

    @staticmethod
    def forward(ctx, points, G, l, sigma, full_fv, flatten):
        ctx.save_for_backward(points)
        ctx.cfg = (G, l, float(sigma), bool(full_fv), bool(flatten))
        return _fv_forward(points, G, l, sigma, full_fv, flatten)

    @staticmethod
    def backward(ctx, grad_fv):
        (points,) = ctx.saved_tensors
        G, l, sigma, full_fv, flatten = ctx.cfg
        lib = _lib.load()
        grad_fv = grad_fv.contiguous().float()
        B, N, _ = points.shape
        grad_points = torch.empty_like(points)
        with torch.cuda.device(points.device):
            rc = lib.dpd_fv_backward(_ptr(points), B, N, G, _lib.fptr(l), sigma, int(full_fv), int(flatten),
                                     _ptr(grad_fv), _ptr(grad_points), _stream())
        _lib.check(rc, "dpd_fv_backward")
        return grad_points, None, None, None, None, None


def get_3dmfv_tf(points, n_gaussians=9, sigma=0.0625, flatten=True, normalize=True, full_fv=True):
    """points [B,N,3] -> fv [B, n_gaussians, 20|7] (flatten=False) or [B, (20|7)*n_gaussians].
    `normalize` is accepted and ignored: the reference hard-wires it to True (:111).
    Differentiable w.r.t. `points`."""
    points = _check_cuda(points, "points")
    if points.dim() != 3 or points.shape[-1] != 3:
        raise ValueError("points must be [B,N,3]")
    G, l = _fv_grid(n_gaussians, 3)
    if torch.is_grad_enabled() and points.requires_grad:
        return _FvFunction.apply(points, G, l, sigma, full_fv, flatten)
    return _fv_forward(points, G, l, sigma, full_fv, flatten)


# --------------------------------------------------------------------------------------
# local patches
# --------------------------------------------------------------------------------------
class LocalPatches:
    """What `local_z` returns in place of the reference's [B,V,k^3*E] tensor (5.12 MB per cloud at
    the defaults): the FV tensor plus k.  `DPDist` gathers patch rows on the fly from `fv`;
    `materialize()` produces the reference's dense tensor on request (embedding_set, parity tests)."""

    def __init__(self, fv, k):
        self.fv, self.k = fv, int(k)

    @property
    def shape(self):
        B, V, E = self.fv.shape
        return torch.Size((B, V, self.k ** 3 * E))

    def materialize(self):
        lib = _lib.load()
        fv = _check_cuda(self.fv, "fv")
        B, V, E = fv.shape
        G = int(np.round(np.power(V, 1 / 3)))
        out = torch.empty((B, V, self.k ** 3 * E), device=fv.device, dtype=torch.float32)
        with torch.cuda.device(fv.device):
            rc = lib.dpd_local_patches(_ptr(fv), B, G, E, self.k, _ptr(out), _stream())
        _lib.check(rc, "dpd_local_patches")
        return out


def local_z(net, is_training, reuse=False, NUM_DIMS=2, k=3, overlap=True):
    """utils/dpdist_util.py:850-854."""
    if NUM_DIMS == 2:
        raise NotImplementedError("local_z_2d is not on the DPDist hot path (NUM_DIMS=3)")
    return local_z_3d(net, is_training, reuse=reuse, NUM_DIMS=NUM_DIMS, k=k, overlap=overlap)


def local_z_3d(net, is_training, reuse=False, NUM_DIMS=3, k=3, overlap=True):
    """net [B,V,E] -> (LocalPatches standing for [B,V,k^3*E], C [V,3] fp32 voxel centres)."""
    net = _check_cuda(net, "net")
    num_vox = net.shape[1]
    grid_len = int(np.round(np.power(num_vox, 1 / 3)))                        # :916
    if grid_len ** NUM_DIMS != num_vox:
        net = net[:, :int(grid_len ** NUM_DIMS), :].contiguous()              # :918
    return LocalPatches(net, k), _centers_tensor(num_vox, NUM_DIMS, net.device)


_CENTERS = {}


def _centers_tensor(num_vox, NUM_DIMS, device):
    """C [V,3] fp32 on `device` (:925-929), built once per (grid, device); carries the host interval tables so
    that DPDist needs no D2H copy."""
    key = (int(num_vox), int(NUM_DIMS), str(device))
    Ct = _CENTERS.get(key)
    if Ct is None:
        X, Y, Z = get_grid_centers(num_vox, NUM_DIMS)
        C = np.stack([X, Y, Z], -1).astype(np.float32).reshape(-1, NUM_DIMS)      # :927-929
        Ct = torch.from_numpy(C).to(device)
        Ct._dpd_tables = _assign_tables(C)
        _CENTERS[key] = Ct
    return Ct


# --------------------------------------------------------------------------------------
# voxel assignment
# --------------------------------------------------------------------------------------
def get_pc_grid_binary_mask_from_centers(Centers, point_cloud):
    """utils/dpdist_util.py:459-492 without the [B,N,V] mask and [B,N,V,3] offsets: returns the three
    things the reference gathers out of them (:436-447): (binary_vect [B,N], offset [B,N,3],
    argmax [B,N] int32)."""
    lib = _lib.load()
    pc = _check_cuda(point_cloud, "point_cloud")
    B, N, _ = pc.shape
    G, l, lo, hi = _assign_tables(Centers)
    idx = torch.empty((B, N), device=pc.device, dtype=torch.int32)
    mask = torch.empty((B, N), device=pc.device, dtype=torch.float32)
    off = torch.empty((B, N, 3), device=pc.device, dtype=torch.float32)
    with torch.cuda.device(pc.device):
        rc = lib.dpd_voxel_assign(_ptr(pc), B, N, G, _lib.fptr(l), _lib.fptr(lo), _lib.fptr(hi),
                                  _ptr(idx), _ptr(mask), _ptr(off), _stream())
    _lib.check(rc, "dpd_voxel_assign")
    return mask, off, idx


# --------------------------------------------------------------------------------------
# implicit distance head
# --------------------------------------------------------------------------------------
class _PackedHead:
    """Kernel-layout copy of the four conv layers, rebuilt when any variable changes.  The cache key is
    the storage address of the variables (kept alive here so addresses cannot be recycled; detach() views
    share it) plus their in-place version counters, so optimizer steps and load_state_dict both invalidate it."""

    def __init__(self):
        self.key = None
        self.owners = None
        self.blob = None
        self.ws = None

    def get(self, lib, cfg, weights, ws_list):
        # weights_generation: bumped by whoever updates the variables behind autograd's back (the trainer's flat Adam
        # launch, CUDA-graph replays), which the tensors' own version counters do not see
        key = (cfg.G, cfg.C, cfg.k, cfg.H, cfg.flags, getattr(tf_util.default_store(), "weights_generation", 0)) + \
            tuple((w.data_ptr(), w._version) for w in weights)
        if key != self.key:
            nbytes = lib.dpd_head_packed_bytes(ctypes.byref(cfg))
            if nbytes == 0:
                _lib.check(-1, "dpd_head_packed_bytes")
            if self.blob is None or self.blob.numel() < nbytes or self.blob.device != ws_list[0].device:
                self.blob = torch.empty(nbytes, device=ws_list[0].device, dtype=torch.uint8)
            rc = lib.dpd_head_pack_weights(ctypes.byref(cfg), *[_ptr(w) for w in ws_list], _ptr(self.blob), _stream())
            _lib.check(rc, "dpd_head_pack_weights")
            self.key = key
            self.owners = list(weights)
        return self.blob

    def workspace(self, lib, cfg, device):
        nbytes = lib.dpd_head_workspace_bytes(ctypes.byref(cfg))
        if nbytes == 0:
            _lib.check(-1, "dpd_head_workspace_bytes")
        if self.ws is None or self.ws.numel() < nbytes or self.ws.device != device:
            self.ws = torch.empty(nbytes, device=device, dtype=torch.uint8)
        return self.ws


_PACKED = {}
HEAD_IMPL = _lib.HEAD_AUTO   # module-level switch: _lib.HEAD_AUTO / HEAD_SIMT / HEAD_TC


def _head_call(fv, query, tables, weights, k, flags, idx=None):
    """One dpd_head_forward call; returns (out, cfg, cache) so a backward can find the activations."""
    lib = _lib.load()
    n_clouds, V, Cc = fv.shape
    NP = query.shape[1]
    G, l, lo, hi = tables
    ws_list = [_check_cuda(w.detach(), "variable") for w in weights]
    w1 = ws_list[0]
    H = w1.shape[-1]
    if w1.numel() != (3 + k ** 3 * Cc) * H:
        raise ValueError("mapper_conv1/weights has %d elements, expected (3+k^3*C)*H = %d" % (w1.numel(), (3 + k ** 3 * Cc) * H))
    cfg = _lib.HeadConfig(n_clouds, NP, G, Cc, int(k), H, flags)
    out = torch.empty((n_clouds, NP, 3), device=fv.device, dtype=torch.float32)
    with torch.cuda.device(fv.device):
        cache = _PACKED.setdefault((fv.device, cfg.flags), _PackedHead())
        blob = cache.get(lib, cfg, list(weights), ws_list)
        ws = cache.workspace(lib, cfg, fv.device)
        rc = lib.dpd_head_forward(ctypes.byref(cfg), _ptr(fv), _ptr(query), _lib.fptr(l), _lib.fptr(lo),
                                  _lib.fptr(hi), _ptr(blob), _ptr(out),
                                  _ptr(idx) if idx is not None else None, _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "dpd_head_forward")
    cache.generation = getattr(cache, "generation", 0) + 1
    return out, cfg, cache


def _grad_sink():
    """The gradient sink of the store in use, if a trainer installed one (train._GradSink): the head's backward then
    writes the variables' gradients into the trainer's flat buffer and reports every finished layer (4..1), so the
    data-parallel exchange of a layer can start while the next one is computed.  Captured at forward time - the
    backward runs on autograd's thread, where the store stack of the caller is not visible."""
    return getattr(tf_util.default_store(), "grad_sink", None)


class _HeadFunction(torch.autograd.Function):
    """Autograd node of the head.  Gradients w.r.t. the 8 variables (train_multi_gpu_pc_compare_dist.py:274-277)
    and / or w.r.t. the inputs fv and query (DPDist as a loss for another network, iterative_PCRNet_ours.py:229-257);
    only what the graph asks for is computed."""

    @staticmethod
    def forward(ctx, fv, query, tables, k, impl, *weights):
        ctx.inputs_need = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        flags = impl | _lib.HEAD_TRAIN | (_lib.HEAD_INPUT_GRAD if ctx.inputs_need else 0)
        out, cfg, cache = _head_call(fv, query, tables, weights, k, flags)
        ctx.fv, ctx.cfg, ctx.cache, ctx.generation = fv.detach(), cfg, cache, cache.generation
        ctx.shapes = [tuple(w.shape) for w in weights]
        ctx.query_shape = tuple(query.shape)
        ctx.sink = _grad_sink()
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return _head_backward(ctx, grad_out)


def _head_backward(ctx, grad_out, n_leading=5):
    """Shared backward of the head nodes: (grad fv, grad query, None x (n_leading - 2), *weight grads)."""
    if True:
        cache, cfg, fv = ctx.cache, ctx.cfg, ctx.fv
        if cache.generation != ctx.generation:
            raise RuntimeError("the activations of this forward were overwritten by a later DPDist call; call backward() first")
        lib = _lib.load()
        grad_out = grad_out.contiguous().float()
        need_w = ctx.needs_input_grad[n_leading:]
        # a layer's weight and bias gradients come out of one product: compute both if either is asked for
        need_layer = [bool(need_w[2 * i] or need_w[2 * i + 1]) for i in range(4)]
        sink = getattr(ctx, "sink", None) if all(need_layer) else None
        if sink is not None:
            grads = list(sink.buffers(ctx.shapes))
        else:
            grads = [torch.empty(s, device=fv.device, dtype=torch.float32) if need_layer[i // 2] else None
                     for i, s in enumerate(ctx.shapes)]
        ptrs = [_ptr(g) if g is not None else None for g in grads]
        grad_fv = grad_query = None
        with torch.cuda.device(fv.device):
            for stage, layer in ((_lib.BWD_L4, 4), (_lib.BWD_L3, 3), (_lib.BWD_L2, 2), (_lib.BWD_L1, 1)):
                if layer == 1 and not need_layer[0]:
                    break
                rc = lib.dpd_head_backward(ctypes.byref(cfg), _ptr(fv), _ptr(cache.blob), _ptr(grad_out), stage, *ptrs,
                                           _ptr(cache.ws), cache.ws.numel(), _stream())
                _lib.check(rc, "dpd_head_backward")
                if sink is not None:
                    sink.ready(layer)
            if ctx.inputs_need:
                grad_fv = torch.empty_like(fv)
                grad_query = torch.empty(ctx.query_shape, device=fv.device, dtype=torch.float32)
                rc = lib.dpd_head_backward_inputs(ctypes.byref(cfg), _ptr(cache.blob), _ptr(grad_fv), _ptr(grad_query),
                                                  _ptr(cache.ws), cache.ws.numel(), _stream())
                _lib.check(rc, "dpd_head_backward_inputs")
        wg = tuple(g if need_w[i] else None for i, g in enumerate(grads))
        return (grad_fv if ctx.needs_input_grad[0] else None, grad_query if ctx.needs_input_grad[1] else None) + \
            (None,) * (n_leading - 2) + wg


class _ModelTrainFunction(torch.autograd.Function):
    """3DmFV + head in ONE library call (dpd_model_forward with DPD_HEAD_TRAIN) as an autograd node for the 8 variables:
    the training step of a DPDist whose inputs are data (no gradient into the clouds).  Saves the separate |fv|max and
    (hi, lo) split passes of the staged path; the backward is the head's."""

    @staticmethod
    def forward(ctx, points, query, n_gaussians, sigma, full_fv, k, mlp_h, impl, *weights):
        lib = _lib.load()
        n_clouds, N, _ = points.shape
        G, l = _fv_grid(n_gaussians, 3)
        V, Cc = G ** 3, FV_CHANNELS[bool(full_fv)]
        Ct = _centers_tensor(V, 3, points.device)
        _, cl, lo, hi = Ct._dpd_tables
        ws_list = [_check_cuda(w.detach(), "variable") for w in weights]
        cfg = _lib.HeadConfig(n_clouds, query.shape[1], G, Cc, int(k), int(mlp_h), impl | _lib.HEAD_TRAIN)
        fv = torch.empty((n_clouds, V, Cc), device=points.device, dtype=torch.float32)
        out = torch.empty((n_clouds, query.shape[1], 3), device=points.device, dtype=torch.float32)
        with torch.cuda.device(points.device):
            cache = _PACKED.setdefault((points.device, cfg.flags), _PackedHead())
            blob = cache.get(lib, cfg, list(weights), ws_list)
            ws = cache.workspace(lib, cfg, points.device)
            rc = lib.dpd_model_forward(ctypes.byref(cfg), _ptr(points), N, float(sigma), _lib.fptr(l), _ptr(query),
                                       _lib.fptr(cl), _lib.fptr(lo), _lib.fptr(hi), _ptr(blob), _ptr(fv), _ptr(out), None,
                                       _ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "dpd_model_forward")
        cache.generation = getattr(cache, "generation", 0) + 1
        ctx.fv, ctx.cfg, ctx.cache, ctx.generation = fv, cfg, cache, cache.generation
        ctx.shapes = [tuple(w.shape) for w in weights]
        ctx.sink = _grad_sink()
        ctx.inputs_need = False
        ctx.query_shape = tuple(query.shape)
        ctx.mark_non_differentiable(fv)
        return out, fv

    @staticmethod
    def backward(ctx, grad_out, _grad_fv):
        return _head_backward(ctx, grad_out, n_leading=8)


def model_forward_train(points, query, n_gaussians, sigma, full_fv, k, mlp, reuse=None, impl=None):
    """model_forward with gradients w.r.t. the 8 variables (`points` / `query` are data): -> (fv, out, C)."""
    points = _check_cuda(points, "points")
    query = _check_cuda(query, "query")
    _check_head_options(k, 1, 3, False, 'relu', mlp)
    G, _ = _fv_grid(n_gaussians, 3)
    Cc = FV_CHANNELS[bool(full_fv)]
    weights = _head_variables(Cc * k ** 3, 3, mlp, reuse)
    flags = HEAD_IMPL if impl is None else impl
    out, fv = _ModelTrainFunction.apply(points, query, n_gaussians, sigma, full_fv, k, mlp[0], flags, *weights)
    return fv, out, _centers_tensor(G ** 3, 3, points.device)


def head_forward(fv, query, C, weights, k, impl=None, return_idx=False):
    """out[c,q,:] = mask * relu6(MLP([query - centre | patch_k(fv[c], voxel(query))])) / 3.
    fv [n_clouds,V,Cc], query [n_clouds,NP,3], weights = [w1,b1,w2,b2,w3,b3,w4,b4] in the
    reference's HWIO layouts.  Differentiable w.r.t. the weights, fv and query."""
    fv = _check_cuda(fv, "fv")
    query = _check_cuda(query, "query")
    if query.shape[0] != fv.shape[0]:
        raise ValueError("fv and query disagree on the number of clouds")
    tables = _assign_tables(C)
    if tables[0] ** 3 != fv.shape[1]:
        raise ValueError("C and fv disagree on the grid")
    impl = HEAD_IMPL if impl is None else impl
    needs_grad = torch.is_grad_enabled() and (fv.requires_grad or query.requires_grad or
                                              any(getattr(w, "requires_grad", False) for w in weights))
    if needs_grad and not return_idx:
        return _HeadFunction.apply(fv, query, tables, k, impl, *weights)
    idx = torch.empty(query.shape[:2], device=fv.device, dtype=torch.int32) if return_idx else None
    out, _, _ = _head_call(fv, query, tables, weights, k, impl, idx)
    return (out, idx) if return_idx else out


def _as_fv(embedding, k):
    """Accept what `local_z` returns (LocalPatches) or the reference's dense [B,V,k^3*E] tensor; in a
    dense patch tensor the FV record of voxel v is its own centre tap (a0=a1=a2=(k-1)//2)."""
    if isinstance(embedding, LocalPatches):
        if embedding.k != k:
            raise ValueError("DPDist k=%d but the patches were extracted with k=%d" % (k, embedding.k))
        return embedding.fv
    emb = _check_cuda(embedding, "embedding")
    E = emb.shape[2] // (k ** 3)
    pb = (k - 1) // 2
    centre = (pb * k + pb) * k + pb
    return emb[:, :, centre * E:(centre + 1) * E].contiguous()


def _check_head_options(k, conv_version, NUM_DIMS, bn, output_act, mlp):
    if k <= 0:
        raise NotImplementedError("k == 0 (global embedding) is not the DPDist hot path")
    if conv_version not in (1, 3):
        raise NotImplementedError("conv_version %d: the shared-MLP head (1) and the 3-D CNN head (3) are implemented" % conv_version)
    if conv_version == 3 and bn:
        raise NotImplementedError("conv_version 3 with batch norm is not implemented")
    if NUM_DIMS != 3:
        raise NotImplementedError("NUM_DIMS must be 3")
    if output_act != 'relu':
        raise NotImplementedError("output_act must be 'relu' (models/dpdist_and_aue.py:74)")
    if len(mlp) != 3 or not (mlp[0] == mlp[1] == mlp[2]):
        raise NotImplementedError("mlp must be three equal widths (reference default [1024,1024,1024])")


def _head_variables(E, NUM_DIMS, mlp, reuse, bn=False):
    """The 8 variables of scope 'dpdist_local' (:514-545), created or looked up exactly as DPDist does; with `bn`
    also the four batch-norm variables of every layer (returned as a second list of 4-tuples)."""
    H = mlp[0]
    with tf_util.variable_scope('dpdist_local', reuse=reuse):                 # :514
        w1, b1 = tf_util.conv2d_variables(1, H, [1, E + NUM_DIMS], 'mapper_conv1', reuse=reuse)   # :516-521
        w2, b2 = tf_util.conv2d_variables(H, mlp[1], [1, 1], 'mapper_conv2', reuse=reuse)         # :529-533
        w3, b3 = tf_util.conv2d_variables(mlp[1], mlp[2], [1, 1], 'mapper_conv3', reuse=reuse)    # :535-539
        w4, b4 = tf_util.conv2d_variables(mlp[2], NUM_DIMS, [1, 1], 'mapper_conv4', reuse=reuse)  # :541-545
        bns = []
        if bn:
            for scope, ch in (('mapper_conv1', H), ('mapper_conv2', mlp[1]), ('mapper_conv3', mlp[2]), ('mapper_conv4', NUM_DIMS)):
                with tf_util.variable_scope(scope, reuse=reuse):
                    bns.append(tf_util.batch_norm_variables(ch, reuse=reuse))
    if bn:
        return [w1, b1, w2, b2, w3, b3, w4, b4], bns
    return [w1, b1, w2, b2, w3, b3, w4, b4]


def _fold_batch_norm(weights, bns):
    """Inference-mode batch norm (is_training False: moving statistics, utils/tf_util.py:221-224) is a per-channel affine
    map after conv + bias, so it folds into the layer:  W' = W * s,  b' = (b - mean) * s + beta,  s = gamma / sqrt(var + eps).
    The folded tensors are cached on the store that owns the variables, keyed on the versions of the inputs (the
    packed-weight cache keys on them in turn)."""
    store = tf_util.default_store()
    folded = store.__dict__.setdefault("_bn_folded", {})
    gen = getattr(store, "weights_generation", 0)
    out = []
    for i, (beta, gamma, mean, var) in enumerate(bns):
        w, b = weights[2 * i], weights[2 * i + 1]
        key = (gen,) + tuple((t.data_ptr(), t._version) for t in (w, b, beta, gamma, mean, var))
        hit = folded.get(i)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                s = gamma * torch.rsqrt(var + tf_util.BN_EPSILON)
                hit = (key, (w.detach() * s).contiguous(), ((b.detach() - mean) * s + beta).contiguous(), (w, b, beta, gamma, mean, var))
            folded[i] = hit
        out += [hit[1], hit[2]]
    return out


class _BnTrainHead(torch.autograd.Function):
    """The head with training-mode batch norm (`--BN 1`, is_training True): conv + bias, batch statistics, affine map and
    activation per layer (utils/tf_util.py:213-227, 558-577), layer by layer in fp32 through dpd_layer_forward /
    dpd_bn_forward, and the gradient graph over it through dpd_bn_backward / dpd_layer_backward.
    Inputs: fv [2B,V,C], query [2B,NP,3], the 8 conv variables (reference layouts) and (gamma, beta) of the 4 layers.
    Outputs: out [2B,NP,3] and, not differentiable, (batch mean, biased batch variance) of every layer for the moving
    averages."""

    @staticmethod
    def forward(ctx, fv, query, tables, k, *params):
        lib = _lib.load()
        weights, affine = params[:8], params[8:]
        fv = _check_cuda(fv.detach(), "fv")
        query = _check_cuda(query.detach(), "query")
        n_clouds, V, Cc = fv.shape
        NP = query.shape[1]
        rows = n_clouds * NP
        G, l, lo, hi = tables
        dev = fv.device
        H = weights[0].shape[-1]
        E = k ** 3 * Cc
        Kp1 = -(-(E + 3) // 32) * 32
        idx = torch.empty(rows, device=dev, dtype=torch.int32)
        mask = torch.empty(rows, device=dev, dtype=torch.float32)
        off = torch.empty((rows, 3), device=dev, dtype=torch.float32)
        st = _stream()
        with torch.cuda.device(dev):
            _lib.check(lib.dpd_voxel_assign(_ptr(query), 1, rows, G, _lib.fptr(l), _lib.fptr(lo), _lib.fptr(hi), _ptr(idx), _ptr(mask),
                                            _ptr(off), st), "dpd_voxel_assign")
            w1 = weights[0].detach().reshape(E + 3, H)
            w1p = torch.cat([w1[3:], w1[:3], w1.new_zeros((Kp1 - E - 3, H))], 0).contiguous()     # patch | offset | pad
            ws_bytes = max(lib.dpd_layer_workspace_bytes(rows, Kp1, H), lib.dpd_layer_workspace_bytes(rows, H, H),
                           lib.dpd_layer_workspace_bytes(rows, H, 3))
            ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
            zs, ys, stats = [], [], []
            x = None
            for layer in range(4):
                wl = w1p if layer == 0 else weights[2 * layer].detach().reshape(H, -1).contiguous()
                bl = weights[2 * layer + 1].detach().contiguous()
                K, N = wl.shape
                z = torch.empty((rows, N), device=dev, dtype=torch.float32)
                if layer == 0:
                    rc = lib.dpd_layer_forward(None, rows, K, _ptr(wl), _ptr(bl), N, 0, _ptr(z), _ptr(fv), _ptr(idx), _ptr(off), NP, G, Cc, k, st)
                else:
                    rc = lib.dpd_layer_forward(_ptr(x), rows, K, _ptr(wl), _ptr(bl), N, 0, _ptr(z), None, None, None, 0, 0, 0, 0, st)
                _lib.check(rc, "dpd_layer_forward")
                gamma, beta = affine[2 * layer].detach().contiguous(), affine[2 * layer + 1].detach().contiguous()
                y = torch.empty_like(z)
                mean = torch.empty(N, device=dev, dtype=torch.float32)
                var = torch.empty(N, device=dev, dtype=torch.float32)
                rc = lib.dpd_bn_forward(_ptr(z), rows, N, _ptr(gamma), _ptr(beta), tf_util.BN_EPSILON, 1 if layer < 3 else 0, _ptr(y),
                                        _ptr(mean), _ptr(var), _ptr(ws), ws.numel(), st)
                _lib.check(rc, "dpd_bn_forward")
                zs.append(z); ys.append(y); stats += [mean, var]
                x = y
        y4 = ys[3]
        out = (torch.clamp(y4, 0.0, 6.0) / 3.0 * mask[:, None]).view(n_clouds, NP, 3)          # :690-691, :697-698
        ctx.cfg = (rows, NP, G, Cc, k, H, E, Kp1)
        ctx.keep = (fv, idx, off, mask, w1p, zs, ys, stats, ws, [w.detach() for w in weights], [a.detach() for a in affine])
        ctx.mark_non_differentiable(*stats)
        return (out,) + tuple(stats)

    @staticmethod
    def backward(ctx, grad_out, *_unused):
        lib = _lib.load()
        rows, NP, G, Cc, k, H, E, Kp1 = ctx.cfg
        fv, idx, off, mask, w1p, zs, ys, stats, ws, weights, affine = ctx.keep
        dev = fv.device
        st = _stream()
        y4 = ys[3]
        dy = (grad_out.reshape(rows, 3).float() * mask[:, None] * ((y4 > 0) & (y4 < 6)).float() / 3.0).contiguous()    # relu6' / 3
        gw, gb, gg, gbeta = [None] * 4, [None] * 4, [None] * 4, [None] * 4
        with torch.cuda.device(dev):
            for layer in (3, 2, 1, 0):
                z = zs[layer]
                N = z.shape[1]
                gamma, beta = affine[2 * layer].contiguous(), affine[2 * layer + 1].contiguous()
                dz = torch.empty_like(z)
                gg[layer] = torch.empty(N, device=dev, dtype=torch.float32)
                gbeta[layer] = torch.empty(N, device=dev, dtype=torch.float32)
                rc = lib.dpd_bn_backward(_ptr(z), _ptr(dy), rows, N, _ptr(gamma), _ptr(beta), _ptr(stats[2 * layer]), _ptr(stats[2 * layer + 1]),
                                         tf_util.BN_EPSILON, 1 if layer < 3 else 0, _ptr(dz), _ptr(gg[layer]), _ptr(gbeta[layer]),
                                         _ptr(ws), ws.numel(), st)
                _lib.check(rc, "dpd_bn_backward")
                gb[layer] = torch.empty(N, device=dev, dtype=torch.float32)
                if layer == 0:
                    gw[0] = torch.empty((E + 3, H), device=dev, dtype=torch.float32)
                    rc = lib.dpd_layer_backward(None, rows, Kp1, _ptr(w1p), H, _ptr(dz), _ptr(gw[0]), _ptr(gb[0]), None, _ptr(fv), _ptr(idx),
                                                _ptr(off), NP, G, Cc, k, _ptr(ws), ws.numel(), st)
                else:
                    wl = weights[2 * layer].reshape(H, -1).contiguous()
                    gw[layer] = torch.empty_like(wl)
                    dx = torch.empty((rows, H), device=dev, dtype=torch.float32)
                    rc = lib.dpd_layer_backward(_ptr(ys[layer - 1]), rows, H, _ptr(wl), N, _ptr(dz), _ptr(gw[layer]), _ptr(gb[layer]), _ptr(dx),
                                                None, None, None, 0, 0, 0, 0, _ptr(ws), ws.numel(), st)
                    dy = dx
                _lib.check(rc, "dpd_layer_backward")
        grads = []
        for layer in range(4):
            grads += [gw[layer].view(weights[2 * layer].shape), gb[layer]]
        for layer in range(4):
            grads += [gg[layer], gbeta[layer]]
        return (None, None, None, None) + tuple(grads)


def _cv3_variables(Cc, k, NUM_DIMS, mlp, reuse):
    """The 16 variables of the conv_version 3 head, scope 'dpdist_local_cnn_fc' (utils/dpdist_util.py:647-687):
    mapper_conv0 (1x1x1, C -> 64), mapper_conv1_1 / 1_2 / 2_1 / 2_2 (3x3x3, 64 -> 64, the two resnet3d blocks :394-410),
    mapper_conv3 (1x1x1, 64 -> 16), mapper_conv5 (16 k^3 + 3 -> mlp[2]), mapper_conv6 (-> 3)."""
    out = []
    with tf_util.variable_scope('dpdist_local_cnn_fc', reuse=reuse):
        out += tf_util.conv3d_variables(Cc, 64, [1, 1, 1], 'mapper_conv0', reuse=reuse)
        for blk in ('mapper_conv1', 'mapper_conv2'):
            for part in ('_1', '_2'):
                out += tf_util.conv3d_variables(64, 64, [3, 3, 3], blk + part, reuse=reuse)
        out += tf_util.conv3d_variables(64, 16, [1, 1, 1], 'mapper_conv3', reuse=reuse)
        out += tf_util.conv2d_variables(16 * k ** 3 + NUM_DIMS, mlp[2], [1, 1], 'mapper_conv5', reuse=reuse)
        out += tf_util.conv2d_variables(mlp[2], NUM_DIMS, [1, 1], 'mapper_conv6', reuse=reuse)
    return out


def _pad_cols(t, n):
    return t if t.shape[1] == n else torch.nn.functional.pad(t, (0, n - t.shape[1]))


def _pad_rows(t, n):
    return t if t.shape[0] == n else torch.nn.functional.pad(t, (0, 0, 0, n - t.shape[0]))


class _Cv3Head(torch.autograd.Function):
    """conv_version 3 head (utils/dpdist_util.py:640-687), the reference's other implicit net (--implicit_net_type 3):
    per query the row [offset | patch] is cut as the REFERENCE cuts it -- net[:, :, :E] / net[:, :, E:] with the three
    offsets first (:455), so the "patch volume" is [offset | patch[0:E-3]] reshaped to [k,k,k,C] and the "offset" fed to the
    FC layer is the last three patch values (:641-642); a drop-in reproduces that.  Then 1x1x1 conv (C -> 64), two residual
    blocks of two 3x3x3 SAME convs (:394-410), 1x1x1 conv (64 -> 16), flatten, concat, FC (mlp[2]) and the output layer.
    Every convolution is one fp32 library call: a 3x3x3 SAME conv over a [k,k,k,64] volume is dpd_layer_forward's gathered
    layer with fv = the volumes, grid k, patch edge 3 (extract_volume_patches semantics, zero padding); 1x1x1 convs and FC
    are dense layers.  115 MFLOP per query (12 x the default head), SIMT fp32: correct, not fast.
    Gradients w.r.t. the 16 variables (training); the clouds are data."""

    @staticmethod
    def forward(ctx, fv, query, tables, k, *weights):
        lib = _lib.load()
        fv = _check_cuda(fv.detach(), "fv")
        query = _check_cuda(query.detach(), "query")
        n_clouds, V, Cc = fv.shape
        NP = query.shape[1]
        rows = n_clouds * NP
        G, l, lo, hi = tables
        dev = fv.device
        k3, E = k ** 3, k ** 3 * Cc
        M = rows * k3
        W = [w.detach() for w in weights]
        H = W[12].shape[-1]
        st = _stream()
        idx = torch.empty(rows, device=dev, dtype=torch.int32)
        mask = torch.empty(rows, device=dev, dtype=torch.float32)
        off = torch.empty((rows, 3), device=dev, dtype=torch.float32)
        keep = {}

        def dense(x, w2d, b, act):
            Kp = -(-w2d.shape[0] // 16) * 16
            x, w2d = _pad_cols(x, Kp).contiguous(), _pad_rows(w2d, Kp).contiguous()
            z = torch.empty((x.shape[0], w2d.shape[1]), device=dev, dtype=torch.float32)
            _lib.check(lib.dpd_layer_forward(_ptr(x), x.shape[0], Kp, _ptr(w2d), _ptr(b.contiguous()), w2d.shape[1], act, _ptr(z),
                                             None, None, None, 0, 0, 0, 0, st), "dpd_layer_forward")
            return x, w2d, z

        Kc = -(-(27 * 64 + 3) // 16) * 16
        idx_c = torch.arange(k3, device=dev, dtype=torch.int32).repeat(rows)
        off_c = torch.zeros((M, 3), device=dev, dtype=torch.float32)

        def conv(x, w5d, b):
            wp = _pad_rows(w5d.reshape(27 * 64, 64), Kc).contiguous()
            z = torch.empty((M, 64), device=dev, dtype=torch.float32)
            _lib.check(lib.dpd_layer_forward(None, M, Kc, _ptr(wp), _ptr(b.contiguous()), 64, 1, _ptr(z), _ptr(x), _ptr(idx_c),
                                             _ptr(off_c), k3, k, 64, 3, st), "dpd_layer_forward (conv3d)")
            return wp, z

        def add(a, b):
            out = a.clone()
            _lib.check(lib.dpd_add_inplace(_ptr(out), _ptr(b), out.numel(), st), "dpd_add_inplace")
            return out

        with torch.cuda.device(dev):
            _lib.check(lib.dpd_voxel_assign(_ptr(query), 1, rows, G, _lib.fptr(l), _lib.fptr(lo), _lib.fptr(hi), _ptr(idx), _ptr(mask),
                                            _ptr(off), st), "dpd_voxel_assign")
            R = torch.empty((rows, E + 3), device=dev, dtype=torch.float32)
            _lib.check(lib.dpd_gather_rows(_ptr(fv), _ptr(idx), _ptr(off), rows, NP, G, Cc, k, _ptr(R), st), "dpd_gather_rows")
            x0 = R[:, :E].reshape(M, Cc)                       # :641,644-646  (offset | patch[:E-3]) as [k,k,k,C]
            netD = R[:, E:E + 3]                               # :642          the last three patch values
            x0p, w0p, y0 = dense(x0, W[0].reshape(Cc, 64), W[1], 1)
            w11, a = conv(y0, W[2], W[3])
            w12, b_ = conv(a, W[4], W[5])
            r1 = add(b_, y0)
            w21, c = conv(r1, W[6], W[7])
            w22, d = conv(c, W[8], W[9])
            r2 = add(d, r1)
            _, w3p, y3 = dense(r2, W[10].reshape(64, 16), W[11], 1)
            f = torch.cat([y3.view(rows, 16 * k3), netD], 1)                        # :671-673
            fp, w5p, y5 = dense(f, W[12].reshape(16 * k3 + 3, H), W[13], 1)
            w6 = W[14].reshape(H, 3).contiguous()
            z6 = torch.empty((rows, 3), device=dev, dtype=torch.float32)
            _lib.check(lib.dpd_layer_forward(_ptr(y5), rows, H, _ptr(w6), _ptr(W[15].contiguous()), 3, 0, _ptr(z6), None, None, None,
                                             0, 0, 0, 0, st), "dpd_layer_forward")
        out = (torch.clamp(z6, 0.0, 6.0) / 3.0 * mask[:, None]).view(n_clouds, NP, 3)          # :690-691, :697-698
        if any(ctx.needs_input_grad[4:]):
            ctx.dims = (rows, M, k, k3, Cc, H, Kc)
            ctx.keep = dict(mask=mask, x0p=x0p, w0p=w0p, y0=y0, a=a, b_=b_, r1=r1, c=c, d=d, r2=r2, y3=y3, fp=fp, w5p=w5p, y5=y5, w6=w6,
                            z6=z6, w11=w11, w12=w12, w21=w21, w22=w22, w3p=w3p, idx_c=idx_c, off_c=off_c,
                            shapes=[tuple(w.shape) for w in weights])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        rows, M, k, k3, Cc, H, Kc = ctx.dims
        K = ctx.keep
        dev = K["y0"].device
        st = _stream()
        extra = 768 << 20          # room for the gathered layers' input gradients (dx1 of a group of volumes)
        ws_bytes = max(lib.dpd_layer_workspace_bytes(M, Kc, 128), lib.dpd_layer_workspace_bytes(rows, K["fp"].shape[1], H),
                       lib.dpd_layer_workspace_bytes(M, 64, 128)) + M * 12 + extra
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)

        def relu_bwd(dy, y):
            _lib.check(lib.dpd_relu_backward(_ptr(dy), _ptr(y), dy.numel(), st), "dpd_relu_backward")
            return dy

        def dense_bwd(x, w2d, dz, want_dx):
            """x [R,K], w2d [K,N], dz [R,N] -> gw [K,N], gb [N], dx [R,K] (N padded to a multiple of 128 for the product)."""
            R, Kd = x.shape
            N = w2d.shape[1]
            Np = N if N <= 4 else -(-N // 128) * 128
            wq, dzq = _pad_cols(w2d, Np).contiguous(), _pad_cols(dz, Np).contiguous()
            gw = torch.empty((Kd, Np), device=dev, dtype=torch.float32)
            gb = torch.empty(Np, device=dev, dtype=torch.float32)
            dx = torch.empty((R, Kd), device=dev, dtype=torch.float32) if want_dx else None
            _lib.check(lib.dpd_layer_backward(_ptr(x), R, Kd, _ptr(wq), Np, _ptr(dzq), _ptr(gw), _ptr(gb), _ptr(dx) if want_dx else None,
                                              None, None, None, 0, 0, 0, 0, _ptr(ws), ws.numel(), st), "dpd_layer_backward")
            return gw[:, :N], gb[:N], dx

        def conv_bwd(x, wp, dz, want_dx):
            """3x3x3 conv: x [M,64] volumes, wp [Kc,64] packed, dz [M,64] -> gw [3,3,3,64,64], gb, dx [M,64]."""
            wq, dzq = _pad_cols(wp, 128).contiguous(), _pad_cols(dz, 128).contiguous()
            gw = torch.empty((27 * 64 + 3, 128), device=dev, dtype=torch.float32)       # reference row order: 3 offset rows first
            gb = torch.empty(128, device=dev, dtype=torch.float32)
            dx = torch.empty((M, 64), device=dev, dtype=torch.float32) if want_dx else None
            _lib.check(lib.dpd_layer_backward(None, M, Kc, _ptr(wq), 128, _ptr(dzq), _ptr(gw), _ptr(gb), _ptr(dx) if want_dx else None,
                                              _ptr(x), _ptr(K["idx_c"]), _ptr(K["off_c"]), k3, k, 64, 3, _ptr(ws), ws.numel(), st),
                       "dpd_layer_backward (conv3d)")
            return gw[3:, :64].reshape(3, 3, 3, 64, 64), gb[:64], dx

        def add_(a, b):
            _lib.check(lib.dpd_add_inplace(_ptr(a), _ptr(b), a.numel(), st), "dpd_add_inplace")
            return a

        with torch.cuda.device(dev):
            z6 = K["z6"]
            dz6 = (grad_out.reshape(rows, 3).float() * K["mask"][:, None] * ((z6 > 0) & (z6 < 6)).float() / 3.0).contiguous()
            g14, g15, dy5 = dense_bwd(K["y5"], K["w6"], dz6, True)
            g12, g13, df = dense_bwd(K["fp"], K["w5p"], relu_bwd(dy5, K["y5"]), True)
            dy3 = df[:, :16 * k3].reshape(M, 16).contiguous()
            g10, g11, dr2 = dense_bwd(K["r2"], K["w3p"][:64], relu_bwd(dy3, K["y3"]), True)
            dr1 = dr2.clone()                                                    # r2 = d + r1
            g8, g9, dc = conv_bwd(K["c"], K["w22"], relu_bwd(dr2, K["d"]), True)
            g6, g7, t = conv_bwd(K["r1"], K["w21"], relu_bwd(dc, K["c"]), True)
            add_(dr1, t)
            dy0 = dr1.clone()                                                    # r1 = b_ + y0
            g4, g5, da = conv_bwd(K["a"], K["w12"], relu_bwd(dr1, K["b_"]), True)
            g2, g3, t = conv_bwd(K["y0"], K["w11"], relu_bwd(da, K["a"]), True)
            add_(dy0, t)
            g0, g1, _ = dense_bwd(K["x0p"], K["w0p"], relu_bwd(dy0, K["y0"]), False)
        sh = K["shapes"]
        grads = [g0[:Cc].reshape(sh[0]), g1, g2.reshape(sh[2]), g3, g4.reshape(sh[4]), g5, g6.reshape(sh[6]), g7, g8.reshape(sh[8]), g9,
                 g10.reshape(sh[10]), g11, g12[:16 * k3 + 3].reshape(sh[12]), g13, g14.reshape(sh[14]), g15]
        return (None, None, None, None) + tuple(g.contiguous() for g in grads)


def model_forward(points, query, n_gaussians, sigma, full_fv, k, mlp, reuse=None, impl=None):
    """One dpd_model_forward call: 3DmFV of `points` [2B,N,3] (rows [A | B]) and the head evaluated at `query`
    [2B,NP,3] (rows [pcB | pcA]) -> (fv [2B,V,C], out [2B,NP,3], C [V,3]).  Inference only (no autograd node);
    results equal get_3dmfv_tf + local_z + DPDist, which get_model uses whenever gradients are needed."""
    lib = _lib.load()
    points = _check_cuda(points, "points")
    query = _check_cuda(query, "query")
    n_clouds, N, D = points.shape
    if D != 3 or query.shape[0] != n_clouds or query.shape[2] != 3:
        raise ValueError("points must be [2B,N,3] and query [2B,NP,3]")
    _check_head_options(k, 1, 3, False, 'relu', mlp)
    G, l = _fv_grid(n_gaussians, 3)
    V, Cc = G ** 3, FV_CHANNELS[bool(full_fv)]
    Ct = _centers_tensor(V, 3, points.device)
    _, cl, lo, hi = Ct._dpd_tables
    weights = _head_variables(Cc * k ** 3, 3, mlp, reuse)
    ws_list = [_check_cuda(w.detach(), "variable") for w in weights]
    H = mlp[0]
    NP = query.shape[1]
    flags = HEAD_IMPL if impl is None else impl
    cfg = _lib.HeadConfig(n_clouds, NP, G, Cc, int(k), H, flags)
    fv = torch.empty((n_clouds, V, Cc), device=points.device, dtype=torch.float32)
    out = torch.empty((n_clouds, NP, 3), device=points.device, dtype=torch.float32)
    with torch.cuda.device(points.device):
        cache = _PACKED.setdefault((points.device, cfg.flags), _PackedHead())
        blob = cache.get(lib, cfg, list(weights), ws_list)
        ws = cache.workspace(lib, cfg, points.device)
        rc = lib.dpd_model_forward(ctypes.byref(cfg), _ptr(points), N, float(sigma), _lib.fptr(l), _ptr(query),
                                   _lib.fptr(cl), _lib.fptr(lo), _lib.fptr(hi), _ptr(blob), _ptr(fv), _ptr(out), None,
                                   _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "dpd_model_forward")
    cache.generation = getattr(cache, "generation", 0) + 1
    return fv, out, Ct


def DPDist(point_cloud, point_cloudB, embedding,
           embeddingB, C, is_training, bn_decay=None, reuse=None,
           bn=True, wd=0.0,
           sig=True, Embedding_Size=512,
           NUM_DIMS=2, mlp=[32, 16, 16], k=3, conv_version=1, output_act='relu'):
    """utils/dpdist_util.py:412-700 for k > 0, conv_version 1, NUM_DIMS 3; `bn` truthy adds batch norm after every conv
    (training mode: batch statistics and moving-average updates; inference: moving statistics).
    Returns [pred_AB, pred_BA], each [B,NP,1,3]: queries point_cloudB against A's field and
    queries point_cloud against B's field (:494-500), masked to the unit cube (:697-698)."""
    _check_head_options(k, conv_version, NUM_DIMS, bn, output_act, mlp)
    pcA = _check_cuda(point_cloud, "point_cloud")
    pcB = _check_cuda(point_cloudB, "point_cloudB")
    fvA, fvB = _as_fv(embedding, k), _as_fv(embeddingB, k)
    if pcA.shape != pcB.shape:
        raise ValueError("point_cloud and point_cloudB must have the same shape (the reference concatenates "
                         "their rows on the batch axis, utils/dpdist_util.py:511)")
    B, NP, _ = pcA.shape
    E = fvA.shape[2] * k ** 3
    H = mlp[0]
    training = not (isinstance(is_training, (bool, int)) and not is_training)
    if conv_version == 3:
        weights = _cv3_variables(fvA.shape[2], k, NUM_DIMS, mlp, reuse)
        if not training:
            weights = [w.detach() for w in weights]
        out = _Cv3Head.apply(torch.cat([fvA, fvB], 0), torch.cat([pcB, pcA], 0), _assign_tables(C), k, *weights)
        out = out.view(2, B, NP, 1, 3)
        return [out[0], out[1]]
    if bn and training:
        # batch statistics (utils/tf_util.py:221-224, 558-577): layer-by-layer fp32 path, _BnTrainHead
        weights, bns = _head_variables(E, NUM_DIMS, mlp, reuse, bn=True)
        fv_all = torch.cat([fvA, fvB], 0)
        query = torch.cat([pcB, pcA], 0)
        affine = [t for (beta, gamma, _, _) in bns for t in (gamma, beta)]
        res = _BnTrainHead.apply(fv_all, query, _assign_tables(C), k, *weights, *affine)
        out, stats = res[0], res[1:]
        # moving averages, updated in place as with updates_collections=None [TF-semantics: the fused batch norm feeds
        # them the unbiased batch variance]; decay = bn_decay, 0.9 if None (utils/tf_util.py:569)
        d = 0.9 if bn_decay is None else float(bn_decay)
        n = fv_all.shape[0] * NP
        with torch.no_grad():
            for i, (_, _, mm, mv) in enumerate(bns):
                mm.mul_(d).add_(stats[2 * i], alpha=1.0 - d)
                mv.mul_(d).add_(stats[2 * i + 1], alpha=(1.0 - d) * n / max(n - 1, 1))
        out = out.view(2, B, NP, 1, 3)
        return [out[0], out[1]]
    if bn:
        # inference: the moving statistics fold into the layers
        weights, bns = _head_variables(E, NUM_DIMS, mlp, reuse, bn=True)
        weights = _fold_batch_norm(weights, bns)
    else:
        weights = _head_variables(E, NUM_DIMS, mlp, reuse)
    fv_all = torch.cat([fvA, fvB], 0)             # rows [A-field | B-field]  (:511)
    query = torch.cat([pcB, pcA], 0)              # A's field is queried at B's points and vice versa (:494,498)
    # is_training only switches batch norm in the reference (off here); a literal False/0 additionally
    # freezes the variables: no weight gradients, and no activations kept unless the INPUTS ask for gradients
    # (DPDist as a loss for another network)
    if isinstance(is_training, (bool, int)) and not is_training:
        weights = [w.detach() for w in weights]
    out = head_forward(fv_all, query, C, weights, k)
    out = out.view(2, B, NP, 1, 3)
    return [out[0], out[1]]                                                    # :695


def get_loss(pred_set, end_points, labels, loss_type='l1_dist'):
    """utils/dpdist_util.py:962-980.  Like the reference it RETURNS (loss_samples, loss_pred) where
    `loss_samples` is the squeezed prediction pred_AB[...,0] (:967-968, :980), and publishes the
    two scalar losses through the collections 'loss_samples' (mean |pred - labels|, :972-974) and
    'loss_pred' (:976-979), which is where the trainer reads them (train...py:262-265)."""
    pred_listAB = pred_set['pred_listAB']
    pred_listBA = pred_set['pred_listBA']
    if loss_type != 'l1_dist':
        raise NotImplementedError("only 'l1_dist' is implemented in the reference")
    loss_samples = pred_listAB[:, :, :, 0]
    loss_samples = loss_samples.squeeze()
    loss = torch.mean(torch.abs(loss_samples - labels))
    tf_util.add_to_collection('loss_samples', loss)
    loss_pred = (torch.mean(pred_listAB[:, :, :, 0]) + torch.mean(pred_listBA[:, :, :, 0])) / 2
    tf_util.add_to_collection('loss_pred', loss_pred)
    return loss_samples, loss_pred
