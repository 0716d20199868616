"""CPU ORACLE for the DPDist hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A literal torch-CPU restatement of the reference's TF1 graph for the path
3DmFV -> local patches -> voxel assignment -> implicit distance MLP -> loss.
It materialises the same [B,N,V,3] tiles and the [B,V,k^3*20] patch tensor the
reference's graph materialises, op for op, so that it is the reference's
*algorithm* (and its cost) on the CPU.

PARITY UNPINNED: TensorFlow 1.x cannot be installed here (Python 3.12, no
network) and the reference ships no tests, golden vectors, checkpoints or data
(SURVEY.md section 8c).  This restatement is therefore checked only against
(i) an fp64 twin of itself, (ii) an independent separable numpy formulation
(oracle/fv_separable_np.py), (iii) the reference's own explicit pad-and-slice
patch loop (utils/dpdist_util.py:932-957) and (iv) analytic invariants.  Every
parity claim in this repo means "matches this restatement", not "matches TF1".

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  dpdist_b200/ never does.

All file:line citations are relative to /root/reference/.
TF-semantics = behaviour of the un-vendored TensorFlow dependency
("TensorFlow >= 1.14", README.md:39) restated from its published definition.
"""
import contextlib
import math

import numpy as np
import torch

LOG_2PI = math.log(2.0 * math.pi)


@contextlib.contextmanager
def tf_cpu_numerics():
    """TF-semantics: TF CPU worker threads run with flush-to-zero/denormals-are-zero.

    Only matters inside the 1e-12 power-normalisation clamp band.
    """
    ok = torch.set_flush_denormal(True)
    try:
        yield ok
    finally:
        torch.set_flush_denormal(False)


# --------------------------------------------------------------------------------------
# grids
# --------------------------------------------------------------------------------------
def fv_grid_size(n_gaussians, D=3):
    """utils/dpdist_util.py:38-41."""
    if D == 2:
        return int(np.sqrt(n_gaussians))
    return int(np.ceil(np.power(n_gaussians, 1 / 3)))


def fv_gmm_centers(n_gaussians):
    """GMM means as the reference builds them, utils/dpdist_util.py:41-50.

    np.meshgrid default 'xy' indexing => flat g = i0*G^2 + i1*G + i2 has
    mu_g = (l[i1], l[i0], l[i2]).  Returns fp64 [V,3]; the reference casts to fp32.
    """
    grid_size = fv_grid_size(n_gaussians, 3)
    l = np.linspace(-1, 1, grid_size, False) + (1 / grid_size)
    x, y, z = np.meshgrid(l, l, l)
    return np.stack([x.flatten(), y.flatten(), z.flatten()]).T


def get_grid_centers(Embedding_Size, NUM_DIMS=2):
    """utils/dpdist_util.py:982-992 (numpy in the reference too)."""
    if NUM_DIMS == 2:
        vec_size = int(np.floor(np.sqrt(Embedding_Size)))
    else:
        vec_size = int(np.ceil(np.power(Embedding_Size, 1 / 3)))
    grid_step = 2 / vec_size
    l = np.arange(-1, 1, grid_step) + grid_step / 2
    if NUM_DIMS == 2:
        return np.meshgrid(l, l)
    return np.meshgrid(l, l, l)


# --------------------------------------------------------------------------------------
# 3DmFV
# --------------------------------------------------------------------------------------
def get_3dmfv(points, n_gaussians=9, sigma=0.0625, flatten=True, normalize=True, full_fv=True):
    """utils/dpdist_util.py:22-141, tile for tile.  points [B,N,3] (fp32 or fp64 CPU tensor)."""
    dt = points.dtype
    n_batches, n_points, D = points.shape
    assert D == 3
    x = fv_gmm_centers(n_gaussians)                                   # :41-48
    if x.shape[0] != n_gaussians:
        raise ValueError("n_gaussians must be a perfect cube (reference: TF shape error at :58/:73)")
    w = torch.ones(n_gaussians, dtype=dt) / n_gaussians               # :49
    mu = torch.tensor(x, dtype=torch.float64).to(dt)                  # :50 tf.constant(x, float32)
    sig = sigma * torch.ones(n_gaussians, D, dtype=dt)                # :51

    batch_sig = sig[None, None].expand(n_batches, n_points, -1, -1)   # :54-55
    batch_mu = mu[None, None].expand(n_batches, n_points, -1, -1)     # :56-57
    batch_w = w[None, None].expand(n_batches, n_points, -1)           # :58
    batch_points = points[:, :, None, :].expand(-1, -1, n_gaussians, -1)  # :59

    nd = D * 3 if full_fv else D
    w_per_batch_per_d = w[None, :, None].expand(n_batches, -1, nd)    # :62-65

    # :69-71  TF-semantics MultivariateNormalDiag.prob = exp(log_prob),
    # log_prob = sum_d[-0.5 z^2 - 0.5 ln 2pi] - sum_d ln sigma, no stabilisation.
    zz = (batch_points - batch_mu) / batch_sig
    log_prob = (-0.5 * zz * zz - 0.5 * LOG_2PI).sum(-1) - torch.log(batch_sig).sum(-1)
    p_per_point = torch.exp(log_prob)

    w_p = p_per_point * batch_w                                       # :73
    Q = w_p / w_p.sum(-1, keepdim=True)                               # :74
    Q_per_d = Q[..., None]                                            # :75

    d_pi_all = ((Q - batch_w) / (torch.sqrt(batch_w) * n_points))[..., None]   # :78
    d_pi_max = d_pi_all.amax(dim=1)                             # :80
    d_pi_mean = d_pi_all.mean(dim=1)                                  # :81
    d_pi = torch.cat([d_pi_mean, d_pi_max], 2) if full_fv else d_pi_mean      # :82-85

    d_mu_all = Q_per_d * (batch_points - batch_mu) / batch_sig        # :87
    d_mu_all_max = d_mu_all.amax(dim=1)                         # :89
    d_mu_all_min = d_mu_all.amin(dim=1)                         # :90
    d_mu_all_mean = d_mu_all.mean(dim=1)                              # :91
    if full_fv:
        d_mu_all_full = torch.cat([d_mu_all_mean, d_mu_all_max, d_mu_all_min], 2)   # :94
    else:
        d_mu_all_full = d_mu_all_mean
    d_mu = (1 / torch.sqrt(w_per_batch_per_d)) * d_mu_all_full        # :98

    d_sig_all = Q_per_d * (torch.pow((batch_points - batch_mu) / batch_sig, 2) - 1)  # :100
    d_sig_all_max = d_sig_all.amax(dim=1)     # amax/amin: gradient split evenly among ties, like TF reduce_max
    d_sig_all_min = d_sig_all.amin(dim=1)
    d_sig_all_mean = d_sig_all.mean(dim=1)
    if full_fv:
        d_sig_all_full = torch.cat([d_sig_all_mean, d_sig_all_max, d_sig_all_min], 2)  # :106
    else:
        d_sig_all_full = d_sig_all_mean
    d_sigma = (1 / torch.sqrt(2 * w_per_batch_per_d)) * d_sig_all_full  # :109

    normalize = True                                                  # :111 (hard-wired)
    if normalize:
        alpha = 0.5
        epsilon = 1e-12

        def pnorm(t):                                                 # :118-121
            return torch.sign(t) * torch.pow(torch.clamp_min(torch.abs(t), epsilon), alpha)

        def l2n(t):                                                   # :124-126 TF-semantics l2_normalize(dim=1)
            ss = (t * t).sum(dim=1, keepdim=True)
            return t * torch.rsqrt(torch.clamp_min(ss, 1e-12))

        d_pi, d_mu, d_sigma = l2n(pnorm(d_pi)), l2n(pnorm(d_mu)), l2n(pnorm(d_sigma))
    if flatten:                                                       # :127-132
        fv = torch.cat([d_pi.transpose(1, 2).reshape(n_batches, -1),
                        d_mu.transpose(1, 2).reshape(n_batches, -1),
                        d_sigma.transpose(1, 2).reshape(n_batches, -1)], dim=1)
    else:                                                             # :133-137 (two transposes cancel)
        fv = torch.cat([d_pi, d_mu, d_sigma], dim=2)
    return fv


# --------------------------------------------------------------------------------------
# local patches
# --------------------------------------------------------------------------------------
def _same_pads(k):
    """TF-semantics 'SAME', stride 1: pad_total = k-1, before = (k-1)//2, after = k//2."""
    return (k - 1) // 2, k // 2


def local_z_3d(net, k=3, explicit_loop=False):
    """utils/dpdist_util.py:911-960.  net [B,V,E] -> (patches [B,V,k^3*E], C [V,3] fp32).

    explicit_loop=False follows the TF14 branch (:921-930, tf.extract_volume_patches,
    SAME); explicit_loop=True follows the reference's own pad-and-slice loop
    (:932-957), which is the in-tree definition of the patch element order.
    """
    batch_size, num_vox, E = net.shape
    grid_len = int(np.round(np.power(num_vox, 1 / 3)))                # :916
    net = net[:, :int(grid_len ** 3), :]                              # :918
    net = net.reshape(batch_size, grid_len, grid_len, grid_len, E)    # :919
    X, Y, Z = get_grid_centers(num_vox, 3)
    if not explicit_loop:
        pb, pa = _same_pads(k)
        padded = torch.nn.functional.pad(net, (0, 0, pb, pa, pb, pa, pb, pa))
        # extract_volume_patches: depth of the output = (a0, a1, a2, c) C-order
        u = padded.unfold(1, k, 1).unfold(2, k, 1).unfold(3, k, 1)    # [B,G,G,G,E,k,k,k]
        u = u.permute(0, 1, 2, 3, 5, 6, 7, 4)                         # [B,G,G,G,k,k,k,E]
        output = u.reshape(batch_size, grid_len ** 3, -1)             # :930
        C = np.stack([X, Y, Z], -1)                                   # :927
        C = torch.tensor(C.astype(np.float32)).reshape(-1, 3)         # :928-929
    else:
        kh = int(np.floor(k / 2))                                     # :934
        padded = torch.nn.functional.pad(net, (0, 0, kh, kh, kh, kh, kh, kh))   # :935-940
        output, C = [], []
        for ii in range(grid_len):
            for jj in range(grid_len):
                for ll in range(grid_len):
                    output.append(padded[:, ii:ii + 2 * kh + 1, jj:jj + 2 * kh + 1, ll:ll + 2 * kh + 1, :])
                    C.append([X[ii, jj, ll], Y[ii, jj, ll], Z[ii, jj, ll]])     # :951-954
        C = torch.tensor(np.array(C)).to(torch.float32)               # tf.stack of fp64 -> used as fp32 downstream
        output = torch.stack(output, 1).reshape(batch_size, grid_len ** 3, -1)  # :956-957
    return output, C.to(net.dtype) if net.dtype == torch.float64 else C


def local_z(net, is_training=None, reuse=False, NUM_DIMS=3, k=3, overlap=True):
    """utils/dpdist_util.py:850-854 (3-D branch only)."""
    assert NUM_DIMS == 3
    return local_z_3d(net, k=k)


# --------------------------------------------------------------------------------------
# voxel assignment + head
# --------------------------------------------------------------------------------------
def get_pc_grid_binary_mask_from_centers(Centers, point_cloud):
    """utils/dpdist_util.py:459-492.  Returns (binary_vect [B,N,V], offsets [B,N,V,3], argmax [B,N] int64)."""
    V = Centers.shape[0]
    grid_size = torch.abs(Centers[0][2] - Centers[1][2]) / 2           # :468
    Cc = Centers[None, None]                                           # :470-471
    pc = point_cloud[:, :, None, :].expand(-1, -1, V, -1)              # :472
    dt = point_cloud.dtype
    A = (pc[..., 0] > Cc[..., 0] - grid_size).to(dt)                   # :478
    Bm = (pc[..., 0] <= Cc[..., 0] + grid_size).to(dt)
    Cm = (pc[..., 1] > Cc[..., 1] - grid_size).to(dt)
    Dm = (pc[..., 1] <= Cc[..., 1] + grid_size).to(dt)
    binary_vect = A * Bm * Cm * Dm                                     # :482
    E = (pc[..., 2] > Cc[..., 2] - grid_size).to(dt)                   # :486-487
    F = (pc[..., 2] <= Cc[..., 2] + grid_size).to(dt)
    binary_vect = binary_vect * E * F                                  # :488
    # TF-semantics tf.math.argmax returns the FIRST maximal index; torch.argmax does not
    # promise that, so take the first index where the row max is attained explicitly.
    is_max = binary_vect == binary_vect.max(dim=2, keepdim=True).values
    ar = torch.arange(V)[None, None].expand_as(is_max)
    argmax = torch.where(is_max, ar, torch.full_like(ar, V)).min(dim=2).values   # :490
    return binary_vect, pc - Cc, argmax                                # :491-492


def get_emb_and_concat(offsets, embedding, argmax, bv):
    """utils/dpdist_util.py:434-457.  embedding [B,V,E] (the reference's [B,1,V,E])."""
    B, NP = argmax.shape
    bi = torch.arange(B)[:, None].expand(B, NP)
    ni = torch.arange(NP)[None, :].expand(B, NP)
    bv_g = bv[bi, ni, argmax]                                          # :436-440
    bv_g = bv_g[..., None, None].expand(-1, -1, 1, 3)                  # :441
    new_pc = offsets[bi, ni, argmax]                                   # :443-447
    new_emb = embedding[bi, argmax]                                    # :449-453
    new_in = torch.cat([new_pc, new_emb], -1)                          # :455 offset FIRST
    return new_in, bv_g


MLP_SCOPES = ["mapper_conv1", "mapper_conv2", "mapper_conv3", "mapper_conv4"]
VAR_PREFIX = "pc_compare/dpdist_local/"


def xavier_uniform_hwio(shape, gen, dtype=torch.float32):
    """TF-semantics tf.contrib.layers.xavier_initializer() (uniform) for an HWIO kernel:
    fan_in = kh*kw*Cin, fan_out = kh*kw*Cout, limit = sqrt(6/(fan_in+fan_out)).
    (utils/tf_util.py:90-91)."""
    rf = int(np.prod(shape[:-2]))
    fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit).to(dtype)


def init_variables(k=5, channels=20, mlp=(1024, 1024, 1024), NUM_DIMS=3, seed=1, dtype=torch.float32,
                   bias_std=0.0, weight_gain=1.0, out_bias=0.0):
    """Variables of DPDist conv_version 1 under their TF names and HWIO shapes
    (utils/dpdist_util.py:514-544; utils/tf_util.py:199-218).  biases are 0 in the
    reference; bias_std / weight_gain (scalar or 4-tuple) / out_bias let tests make the
    activations O(1) so ReLU, relu6 saturation and the rtol are all exercised
    (at the Xavier init the outputs are ~1e-4 and any comparison is atol-dominated)."""
    gen = torch.Generator().manual_seed(seed)
    gains = weight_gain if isinstance(weight_gain, (tuple, list)) else (weight_gain,) * 4
    E = k ** 3 * channels
    dims = [(1, E + NUM_DIMS, 1, mlp[0]), (1, 1, mlp[0], mlp[1]), (1, 1, mlp[1], mlp[2]), (1, 1, mlp[2], NUM_DIMS)]
    out = {}
    for i, (scope, shp) in enumerate(zip(MLP_SCOPES, dims)):
        out[VAR_PREFIX + scope + "/weights"] = xavier_uniform_hwio(shp, gen, dtype) * gains[i]
        b = torch.randn(shp[-1], generator=gen, dtype=torch.float64) * bias_std
        if i == 3:
            b = b + out_bias
        out[VAR_PREFIX + scope + "/biases"] = b.to(dtype)
    return out


def unit_scale_variables(seed=1, dtype=torch.float32, **kw):
    """Test weights with O(1) activations in every layer and outputs spread over [0,2]."""
    return init_variables(seed=seed, dtype=dtype, bias_std=0.05, weight_gain=(600.0, 2.0, 2.0, 1.0), out_bias=1.0, **kw)


def conv2d_1xw(x, weights, biases, relu):
    """utils/tf_util.py:161-228 for the two shapes the path uses: a [1,W] VALID conv over an
    input whose width is exactly W and channel count 1 (layer 1), and 1x1 convs.
    x [..., Cin_total]; weights HWIO."""
    kh, kw, cin, cout = weights.shape
    y = x @ weights.reshape(kh * kw * cin, cout) + biases              # conv2d + bias_add :213-219
    return torch.relu(y) if relu else y                                # :226-227


BN_EPSILON = 0.001   # TF-semantics: tf.contrib.layers.batch_norm default


def bn_inference_variables(seed=2, mlp=(1024, 1024, 1024), dtype=torch.float32):
    """Random moving statistics / affine parameters under the TF names of tf.contrib.layers.batch_norm(scope='bn')
    inside each conv scope (utils/tf_util.py:221-224, 573-577)."""
    gen = torch.Generator().manual_seed(seed)
    out = {}
    for scope, ch in zip(MLP_SCOPES, list(mlp) + [3]):
        p = VAR_PREFIX + scope + "/bn/"
        out[p + "beta"] = (torch.randn(ch, generator=gen, dtype=torch.float64) * 0.1).to(dtype)
        out[p + "gamma"] = (1.0 + 0.2 * torch.randn(ch, generator=gen, dtype=torch.float64)).to(dtype)
        out[p + "moving_mean"] = (torch.randn(ch, generator=gen, dtype=torch.float64) * 0.2).to(dtype)
        out[p + "moving_variance"] = (0.5 + torch.rand(ch, generator=gen, dtype=torch.float64)).to(dtype)
    return out


BN_TRAINING = True   # bn="train" is implemented (tests/test_tf_golden.py looks at this)


def get_bn_decay(batch, decay_step=300 * 512):
    """train_multi_gpu_pc_compare_dist.py:172-175, 992-1000: min(0.99, 1 - 0.5 * 0.5^floor(batch / DECAY_STEP))."""
    return min(0.99, 1.0 - 0.5 * 0.5 ** (int(batch) // int(decay_step)))


def DPDist(point_cloud, point_cloudB, embedding, embeddingB, C, variables, output_act="relu", bn=False, bn_decay=None,
           bn_updates=None):
    """utils/dpdist_util.py:412-544,688-700, conv_version 1, k>0.

    bn False: no batch norm (the reference default, --BN 0).  bn True: inference-mode batch norm (moving statistics).
    bn "train": training-mode batch norm -- tf.contrib.layers.batch_norm(is_training=True, decay=bn_decay,
    updates_collections=None), utils/tf_util.py:558-577 [TF-semantics: the fused implementation normalises with the biased
    batch variance and feeds the moving average with the unbiased one; epsilon 0.001]; the new moving statistics are
    returned through the dict `bn_updates` ({variable name: tensor}) instead of being assigned in place.

    embedding / embeddingB are the [B,V,k^3*20] patch tensors from local_z.
    Returns [pred_AB, pred_BA], each [B,NP,1,3]."""
    bv, net, argmax = get_pc_grid_binary_mask_from_centers(C, point_cloudB)      # :494
    net, binary_vect = get_emb_and_concat(net, embedding, argmax, bv)            # :495-496
    bvB, netB, argmaxB = get_pc_grid_binary_mask_from_centers(C, point_cloud)    # :498
    netB, binary_vectB = get_emb_and_concat(netB, embeddingB, argmaxB, bvB)      # :499-500
    x = torch.cat([net, netB], 0)                                                # :511  [2B,NP,E+3]
    for i, scope in enumerate(MLP_SCOPES):                                       # :516-544
        x = conv2d_1xw(x, variables[VAR_PREFIX + scope + "/weights"],
                       variables[VAR_PREFIX + scope + "/biases"], relu=(i < 3) and not bn)
        if bn == "train":   # batch statistics over every row of the tower (moments over axes [0,1,2] of NHWC)
            p = VAR_PREFIX + scope + "/bn/"
            n = x.shape[0] * x.shape[1]
            mean = x.mean(dim=(0, 1))
            var = ((x - mean) ** 2).mean(dim=(0, 1))
            x = (x - mean) * torch.rsqrt(var + BN_EPSILON) * variables[p + "gamma"] + variables[p + "beta"]
            if bn_updates is not None:
                d = 0.9 if bn_decay is None else bn_decay                        # utils/tf_util.py:569
                bn_updates[p + "moving_mean"] = (d * variables[p + "moving_mean"] + (1 - d) * mean).detach()
                bn_updates[p + "moving_variance"] = (d * variables[p + "moving_variance"] + (1 - d) * var * (n / max(n - 1, 1))).detach()
            if i < 3:
                x = torch.relu(x)
        elif bn:   # inference-mode batch norm between bias_add and the activation (utils/tf_util.py:219-227)
            p = VAR_PREFIX + scope + "/bn/"
            x = (x - variables[p + "moving_mean"]) * torch.rsqrt(variables[p + "moving_variance"] + BN_EPSILON) \
                * variables[p + "gamma"] + variables[p + "beta"]
            if i < 3:
                x = torch.relu(x)
    x = x[:, :, None, :]                                                         # [2B,NP,1,3]
    if output_act == "relu":
        x = torch.clamp(x, 0.0, 6.0) / 3                                         # :690-691
    else:
        raise NotImplementedError("only output_act='relu' is reachable (models/dpdist_and_aue.py:74)")
    B = point_cloud.shape[0]
    return [x[:B] * binary_vect, x[B:] * binary_vectB]                           # :695-698


CV3_PREFIX = "pc_compare/dpdist_local_cnn_fc/"
CV3_LAYERS = [("mapper_conv0", (1, 1, 1, None, 64)), ("mapper_conv1_1", (3, 3, 3, 64, 64)), ("mapper_conv1_2", (3, 3, 3, 64, 64)),
              ("mapper_conv2_1", (3, 3, 3, 64, 64)), ("mapper_conv2_2", (3, 3, 3, 64, 64)), ("mapper_conv3", (1, 1, 1, 64, 16))]


def init_cv3_variables(k=5, channels=20, mlp=(1024, 1024, 1024), seed=1, dtype=torch.float32, gain=1.0, bias_std=0.0, out_bias=0.0):
    """Variables of the conv_version 3 head under their TF names (utils/dpdist_util.py:647-687; conv3d kernels are DHWIO,
    utils/tf_util.py:347-353).  gain / bias_std / out_bias as in init_variables."""
    gen = torch.Generator().manual_seed(seed)
    out = {}
    shapes = [(n, tuple(channels if d is None else d for d in shp)) for n, shp in CV3_LAYERS]
    shapes += [("mapper_conv5", (1, 1, 16 * k ** 3 + 3, mlp[2])), ("mapper_conv6", (1, 1, mlp[2], 3))]
    gains = gain if isinstance(gain, (tuple, list)) else (gain,) * len(shapes)
    for (name, shp), gn in zip(shapes, gains):
        out[CV3_PREFIX + name + "/weights"] = xavier_uniform_hwio(shp, gen, dtype) * gn
        b = torch.randn(shp[-1], generator=gen, dtype=torch.float64) * bias_std
        if name == "mapper_conv6":
            b = b + out_bias
        out[CV3_PREFIX + name + "/biases"] = b.to(dtype)
    return out


def _conv3d_same(x, w, b, relu=True):
    """tf.nn.conv3d(NDHWC, DHWIO kernel, stride 1, 'SAME') + bias_add (+ relu), utils/tf_util.py:355-371."""
    y = torch.nn.functional.conv3d(x.permute(0, 4, 1, 2, 3), w.permute(4, 3, 0, 1, 2), padding=[(d - 1) // 2 for d in w.shape[:3]])
    y = y.permute(0, 2, 3, 4, 1) + b
    return torch.relu(y) if relu else y


def DPDist_cv3(point_cloud, point_cloudB, embedding, embeddingB, C, variables, k):
    """utils/dpdist_util.py:412-511, 640-700 with conv_version 3 (NUM_DIMS 3, bn off), as written: the row
    [offset (3) | patch (E)] is sliced at E, so net_E = [offset | patch[:E-3]] and net_D = patch[E-3:] (:641-642, :455)."""
    bv, net, argmax = get_pc_grid_binary_mask_from_centers(C, point_cloudB)
    net, binary_vect = get_emb_and_concat(net, embedding, argmax, bv)
    bvB, netB, argmaxB = get_pc_grid_binary_mask_from_centers(C, point_cloud)
    netB, binary_vectB = get_emb_and_concat(netB, embeddingB, argmaxB, bvB)
    x = torch.cat([net, netB], 0)                                                # :511  [2B,NP,E+3]
    B2, NP, _ = x.shape
    E = embedding.shape[2]
    net_E, net_D = x[:, :, :E], x[:, :, E:]                                      # :641-642
    v = net_E.reshape(B2 * NP, k, k, k, -1)                                      # :644-646
    P = CV3_PREFIX
    get = lambda n: (variables[P + n + "/weights"], variables[P + n + "/biases"])
    v = _conv3d_same(v, *get("mapper_conv0"))                                    # :648-652
    for blk in ("mapper_conv1", "mapper_conv2"):                                 # resnet3d :394-410, :653-662
        t = _conv3d_same(v, *get(blk + "_1"))
        t = _conv3d_same(t, *get(blk + "_2"))
        v = t + v
    v = _conv3d_same(v, *get("mapper_conv3"))                                    # :663-667
    f = torch.cat([v.reshape(B2, NP, -1), net_D], -1)                            # :668-673
    w5, b5 = get("mapper_conv5")
    w6, b6 = get("mapper_conv6")
    f = torch.relu(f @ w5.reshape(w5.shape[2], w5.shape[3]) + b5)                # :680-684
    f = f @ w6.reshape(w6.shape[2], w6.shape[3]) + b6                            # :686-690
    out = (torch.clamp(f, 0.0, 6.0) / 3)[:, :, None, :]                          # :690-691
    B = point_cloud.shape[0]
    return [out[:B] * binary_vect, out[B:] * binary_vectB]                       # :695-698


def get_loss(pred_set, end_points, labels, loss_type="l1_dist"):
    """utils/dpdist_util.py:962-980."""
    pred_listAB, pred_listBA = pred_set["pred_listAB"], pred_set["pred_listBA"]
    assert loss_type == "l1_dist"
    loss_samples = pred_listAB[:, :, :, 0].squeeze()
    loss = torch.mean(torch.abs(loss_samples - labels))
    loss_pred = (torch.mean(pred_listAB[:, :, :, 0]) + torch.mean(pred_listBA[:, :, :, 0])) / 2
    return loss, loss_pred


def get_model(pcA, pcB, variables, Embedding_Size=512, k=5, full_fv=True, sigma3dmfv=0.125, add_noise=0, bn=False,
              bn_decay=None, bn_updates=None, conv_version=1):
    """models/dpdist_and_aue.py:31-86 (3dmfv encoder, k>0, conv_version 1)."""
    pcA_noise = pcA + add_noise                                                  # :45
    embedding_A = get_3dmfv(pcA_noise, n_gaussians=Embedding_Size, flatten=False,
                            full_fv=full_fv, normalize=True, sigma=sigma3dmfv)   # :56-58
    embedding_B = get_3dmfv(pcB, n_gaussians=Embedding_Size, flatten=False,
                            full_fv=full_fv, normalize=True, sigma=sigma3dmfv)   # :59-61
    fvA, fvB = embedding_A, embedding_B
    embedding_A, C = local_z(embedding_A, k=k)                                   # :64
    embedding_B, _ = local_z(embedding_B, k=k)                                   # :65
    C = C.to(pcA.dtype)
    if conv_version == 3:
        net = DPDist_cv3(pcA, pcB, embedding_A, embedding_B, C, variables, k)
    else:
        net = DPDist(pcA, pcB, embedding_A, embedding_B, C, variables, bn=bn, bn_decay=bn_decay, bn_updates=bn_updates)   # :69-75
    pred_set = {"pred_listAB": net[0], "pred_listBA": net[1]}                    # :80-81
    embedding_set = {"embedding_A": embedding_A, "embedding_B": embedding_B}
    return pred_set, {"fvA": fvA, "fvB": fvB, "C": C}, embedding_set


def forward_chunked(pcA, pcB, variables, chunk=8, **kw):
    """get_model over batch chunks so the literal patch tensor stays bounded
    (5.12 MB per cloud).  Pairs are independent with BN off, so this is exact."""
    outs_ab, outs_ba = [], []
    for s in range(0, pcA.shape[0], chunk):
        p, _, _ = get_model(pcA[s:s + chunk], pcB[s:s + chunk], variables, **kw)
        outs_ab.append(p["pred_listAB"])
        outs_ba.append(p["pred_listBA"])
    return torch.cat(outs_ab), torch.cat(outs_ba)


# --------------------------------------------------------------------------------------
# data side (test infrastructure for dpdist_b200/data.py)
# --------------------------------------------------------------------------------------
def nearest_distance(point_set, neg_set):
    """dataset_sample_with_gt.py:90-91: dist = cdist(point_set, neg_set_rand); dist_AB = dist.min(0)  (float64)."""
    from scipy.spatial.distance import cdist
    d = cdist(np.asarray(point_set, dtype=np.float64), np.asarray(neg_set, dtype=np.float64))
    return d.min(0), d.argmin(0)


def rotate_shift(batch_data, angles, shifts):
    """provider.rotate_point_cloud (provider.py:32-50) + provider.shift_point_cloud (:200-211) with given draws."""
    out = np.zeros(batch_data.shape, dtype=np.float32)
    for k in range(batch_data.shape[0]):
        cosval, sinval = np.cos(angles[k]), np.sin(angles[k])
        rotation_matrix = np.array([[cosval, 0, sinval], [0, 1, 0], [-sinval, 0, cosval]])
        out[k] = np.dot(batch_data[k].reshape((-1, 3)), rotation_matrix)
        out[k] += shifts[k]
    return out


def assemble_batch(batch_data, batch_label, NUM_POINT):
    """train_multi_gpu_pc_compare_dist.py:749-766, literally."""
    H_NUM_POINT = int(NUM_POINT / 2)
    split_off_surface = 0.5
    batch_data = np.split(batch_data, 3, 1)
    batch_surface = np.split(batch_data[0], 2, 1)
    bsize = batch_data[0].shape[0]
    pcA = batch_surface[0][:, :NUM_POINT]
    batch_label = np.split(batch_label, 2, 1)
    labels_AB = np.concatenate(
        [np.zeros([bsize, H_NUM_POINT]), batch_label[0][:, :int(H_NUM_POINT * split_off_surface)],
         batch_label[1][:, int(H_NUM_POINT * split_off_surface):H_NUM_POINT]], 1)
    batch_off = np.concatenate([batch_data[1][:, :int(H_NUM_POINT * split_off_surface)],
                                batch_data[2][:, int(H_NUM_POINT * split_off_surface):H_NUM_POINT]], 1)
    pcB = np.concatenate([batch_surface[1][:, :H_NUM_POINT], batch_off], 1)
    return pcA, pcB, labels_AB
