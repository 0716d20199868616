"""GPU parity tests: the CUDA path, called through the C ABI (ctypes) and the reference-shaped
Python API, against the CPU oracle on the same seeded inputs.  Run with -m gpu on the B200 box."""
import ctypes
import os

import numpy as np
import pytest
import torch

from dpdist_b200 import _lib, dpdist_and_aue as MODEL, dpdist_util, synthetic, tf_util
from oracle import dpdist_oracle as O
from tolerances import assert_close, assert_fv_close, assert_out_close

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def _cloud(seed, B, N, scale=0.9):
    rng = np.random.default_rng(seed)
    return rng.uniform(-scale, scale, size=(B, N, 3)).astype(np.float32)


def _oracle_fv(pts, V, sigma, **kw):
    with O.tf_cpu_numerics():
        return O.get_3dmfv(torch.tensor(pts), V, sigma, **kw)


def _store_from(var):
    store = tf_util.VariableStore(device=DEV)
    store.load_state_dict(var, strict=False)
    return store


# ------------------------------------------------------------------ 3DmFV
@pytest.mark.parametrize("G,N,sigma", [(8, 64, 0.125), (8, 1, 0.125), (8, 7, 0.125), (8, 200, 0.125), (8, 512, 0.125),
                                       (5, 64, 0.2), (3, 33, 0.25), (4, 64, 0.25), (8, 64, 0.25), (10, 40, 0.1)])
@pytest.mark.parametrize("full_fv", [True, False])
def test_fv_matches_oracle(G, N, sigma, full_fv):
    pts = _cloud(G * 1000 + N, 3, N)
    want = _oracle_fv(pts, G ** 3, sigma, flatten=False, full_fv=full_fv)
    got = dpdist_util.get_3dmfv_tf(torch.tensor(pts, device=DEV), n_gaussians=G ** 3, sigma=sigma, flatten=False, full_fv=full_fv)
    assert got.shape == want.shape
    assert_fv_close(got, want, "fv G=%d N=%d" % (G, N))


def test_fv_flatten_and_batch_independence():
    pts = _cloud(1, 5, 64)
    d = torch.tensor(pts, device=DEV)
    a = dpdist_util.get_3dmfv_tf(d, n_gaussians=512, sigma=0.125, flatten=False)
    b = dpdist_util.get_3dmfv_tf(d, n_gaussians=512, sigma=0.125, flatten=True)
    assert torch.equal(b.view(5, 20, 512), a.transpose(1, 2))
    one = dpdist_util.get_3dmfv_tf(d[2:3].contiguous(), n_gaussians=512, sigma=0.125, flatten=False)
    assert torch.equal(one[0], a[2])                               # no cross-cloud coupling, deterministic
    assert_fv_close(b, _oracle_fv(pts, 512, 0.125, flatten=True), "fv flatten")


def test_fv_chair_clouds_and_far_points():
    pcA, pcB, _ = synthetic.chair_batch(3, 4, 64)
    pts = np.concatenate([pcA, pcB], 0)
    pts[0, :4] = [[1.4, 0, 0], [0, -1.2, 0.3], [0.99, 0.99, 0.99], [-1.0, -1.0, -1.0]]   # outside the cube: legal FV input
    assert_fv_close(dpdist_util.get_3dmfv_tf(torch.tensor(pts, device=DEV), n_gaussians=512, sigma=0.125, flatten=False),
                    _oracle_fv(pts, 512, 0.125, flatten=False), "fv chairs")


def test_fv_large_batch_properties():
    """BASELINE size (2048 clouds): size-independent properties instead of an oracle run."""
    pcA, pcB, _ = synthetic.uniform_batch(2, 1024, 64)
    pts = torch.tensor(np.concatenate([pcA, pcB], 0), device=DEV)
    fv = dpdist_util.get_3dmfv_tf(pts, n_gaussians=512, sigma=0.125, flatten=False)
    assert torch.isfinite(fv).all()
    ss = (fv.double() ** 2).sum(1)
    assert torch.allclose(ss, torch.ones_like(ss), atol=1e-5)              # every channel L2-normalised
    perm = torch.randperm(64, generator=torch.Generator().manual_seed(0)).to(DEV)
    fv_p = dpdist_util.get_3dmfv_tf(pts[:, perm].contiguous(), n_gaussians=512, sigma=0.125, flatten=False)
    assert_close(fv_p, fv, 1e-4, 1e-5, "permutation invariance")           # sums re-associate; sqrt amplifies near 0
    sub = _oracle_fv(pts[1000:1004].cpu().numpy(), 512, 0.125, flatten=False)
    assert_fv_close(fv[1000:1004], sub, "fv slice of the large batch")


# ------------------------------------------------------------------ voxel assignment / patches
@pytest.mark.parametrize("G", [8, 5, 3, 4])
def test_voxel_assign_bit_exact(G):
    V = G ** 3
    X, Y, Z = O.get_grid_centers(V, 3)
    C = torch.tensor(np.stack([X, Y, Z], -1).astype(np.float32).reshape(-1, 3))
    rng = np.random.default_rng(G)
    pc = rng.uniform(-1.2, 1.2, size=(4, 300, 3)).astype(np.float32)
    l = (np.arange(-1, 1, 2 / G) + 1 / G).astype(np.float32)
    edges = np.concatenate([l - np.float32(1 / G), l + np.float32(1 / G), [np.float32(-1), np.float32(1)]]).astype(np.float32)
    # points exactly on cell edges, one ulp either side, and non-finite coordinates
    e = rng.choice(edges, size=(4, 120, 3)).astype(np.float32)
    pc[:, :40] = e[:, :40]
    pc[:, 40:80] = np.nextafter(e[:, 40:80], np.float32(2))
    pc[:, 80:120] = np.nextafter(e[:, 80:120], np.float32(-2))
    # nextafter(0) is a denormal: the TF1 CPU runtime runs DAZ, IEEE GPUs do not; use the smallest normals instead
    tiny = np.float32(1.1754944e-38)
    pc[(np.abs(pc) < tiny) & (pc > 0)] = tiny
    pc[(np.abs(pc) < tiny) & (pc < 0)] = -tiny
    pc[0, 120] = [np.nan, 0, 0]; pc[0, 121] = [np.inf, 0, 0]; pc[0, 122] = [0, -np.inf, 0]
    bv, off, am = O.get_pc_grid_binary_mask_from_centers(C, torch.tensor(pc))
    bi = torch.arange(4)[:, None].expand(4, 300); ni = torch.arange(300)[None].expand(4, 300)
    want_mask, want_off = bv[bi, ni, am], off[bi, ni, am]
    mask, goff, idx = dpdist_util.get_pc_grid_binary_mask_from_centers(C.to(DEV), torch.tensor(pc, device=DEV))
    assert idx.dtype == torch.int32
    assert torch.equal(idx.cpu().long(), am)
    assert torch.equal(mask.cpu(), want_mask)
    fin = torch.isfinite(want_off)
    assert torch.equal(goff.cpu()[fin], want_off[fin])
    assert int((want_mask == 0).sum()) > 0 and int((want_mask == 1).sum()) > 0


@pytest.mark.parametrize("G,k,C", [(8, 5, 20), (5, 3, 20), (4, 4, 7), (3, 5, 20), (8, 1, 20)])
def test_local_patches_bit_exact(G, k, C):
    fv = torch.randn(3, G ** 3, C, generator=torch.Generator().manual_seed(G * k))
    want, Cw = O.local_z_3d(fv, k=k)
    lp, Cg = dpdist_util.local_z(fv.to(DEV), None, NUM_DIMS=3, k=k)
    assert lp.shape == want.shape
    assert torch.equal(lp.materialize().cpu(), want)
    assert torch.equal(Cg.cpu(), Cw)


# ------------------------------------------------------------------ head + full model
def _run_model(pcA, pcB, var, impl, k=5, V=512, sigma=0.125, H=1024, fused=True):
    dpdist_util.HEAD_IMPL = impl
    MODEL.FUSED_INFERENCE = fused
    try:
        with tf_util.use_store(_store_from(var)):
            p, _, emb = MODEL.get_model(torch.tensor(pcA, device=DEV), torch.tensor(pcB, device=DEV), False, bn=0,
                                        Embedding_Size=V, k=k, localSNmlp=[H, H, H], sigma3dmfv=sigma, reuse=True)
        torch.cuda.synchronize()
        return p, emb
    finally:
        dpdist_util.HEAD_IMPL = _lib.HEAD_AUTO
        MODEL.FUSED_INFERENCE = True


IMPLS = [_lib.HEAD_SIMT, _lib.HEAD_AUTO]


@pytest.mark.parametrize("impl", IMPLS)
def test_model_anchor_matches_golden_and_oracle(impl):
    z = np.load(os.path.join(GOLDEN, "anchor_A.npz"))
    var = O.unit_scale_variables(int(z["weight_seed"]))
    p, emb = _run_model(z["pcA"], z["pcB"], var, impl)
    assert p["pred_listAB"].shape == (1, 64, 1, 3)
    assert_fv_close(emb["embedding_A"].fv, z["fvA"], "fvA vs golden")
    assert_fv_close(emb["embedding_B"].fv, z["fvB"], "fvB vs golden")
    assert_out_close(p["pred_listAB"], z["pred_AB"], "pred_AB vs golden")
    assert_out_close(p["pred_listBA"], z["pred_BA"], "pred_BA vs golden")
    with O.tf_cpu_numerics():
        po, _, _ = O.get_model(torch.tensor(z["pcA"]), torch.tensor(z["pcB"]), var)
    assert_out_close(p["pred_listAB"], po["pred_listAB"], "pred_AB vs oracle")
    assert_out_close(p["pred_listBA"], po["pred_listBA"], "pred_BA vs oracle")
    tf_util.clear_collections()
    MODEL.get_loss(p, {}, torch.tensor(z["labels"], device=DEV))
    assert abs(float(tf_util.get_collection("loss_samples")[-1]) - float(z["loss"])) < 1e-5


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("B,seed", [(16, 7), (3, 8)])
def test_model_batch_matches_oracle(impl, B, seed):
    pcA, pcB, _ = synthetic.uniform_batch(seed, B, 64, outside_frac=0.05)
    var = O.unit_scale_variables(seed)
    p, _ = _run_model(pcA, pcB, var, impl)
    with O.tf_cpu_numerics():
        ab, ba = O.forward_chunked(torch.tensor(pcA), torch.tensor(pcB), var, chunk=8)
    assert_out_close(p["pred_listAB"], ab, "pred_AB")
    assert_out_close(p["pred_listBA"], ba, "pred_BA")
    outside = (np.abs(pcB) > 1.0).any(-1)
    assert outside.any() and float(p["pred_listAB"][torch.tensor(outside, device=DEV)].abs().max()) == 0.0


@pytest.mark.parametrize("impl", IMPLS)
def test_model_xavier_init_matches_oracle(impl):
    """Reference initial state (Xavier weights, zero biases): outputs ~1e-4, comparison is atol-level."""
    pcA, pcB, _ = synthetic.chair_batch(1, 4, 64)
    var = O.init_variables(seed=5)
    p, _ = _run_model(pcA, pcB, var, impl)
    with O.tf_cpu_numerics():
        ab, ba = O.forward_chunked(torch.tensor(pcA), torch.tensor(pcB), var, chunk=4)
    assert_out_close(p["pred_listAB"], ab)
    assert_out_close(p["pred_listBA"], ba)


@pytest.mark.parametrize("G,k,H", [(5, 3, 256), (4, 5, 512), (8, 3, 256)])
def test_model_other_grids_simt(G, k, H):
    pcA, pcB, _ = synthetic.uniform_batch(G + k, 4, 32, outside_frac=0.05)
    var = O.init_variables(k=k, mlp=(H, H, H), seed=2, bias_std=0.05, weight_gain=(600.0, 2.0, 2.0, 1.0), out_bias=1.0)
    p, _ = _run_model(pcA, pcB, var, _lib.HEAD_AUTO, k=k, V=G ** 3, sigma=1.0 / G, H=H)
    with O.tf_cpu_numerics():
        po, _, _ = O.get_model(torch.tensor(pcA), torch.tensor(pcB), var, Embedding_Size=G ** 3, k=k, sigma3dmfv=1.0 / G)
    assert_out_close(p["pred_listAB"], po["pred_listAB"])
    assert_out_close(p["pred_listBA"], po["pred_listBA"])


@pytest.mark.parametrize("impl,G,k,H", [(_lib.HEAD_AUTO, 8, 5, 1024), (_lib.HEAD_SIMT, 8, 5, 1024), (_lib.HEAD_AUTO, 5, 3, 256),
                                        (_lib.HEAD_TC_TF32, 8, 5, 1024)])
def test_one_call_forward_equals_the_three_stages(impl, G, k, H):
    """dpd_model_forward (3DmFV emitting the tensor-core operand copy, |fv| <= 1 bound) is bit-identical to
    dpd_fv_forward -> dpd_head_forward (measured |fv|max, separate split)."""
    pcA, pcB, _ = synthetic.uniform_batch(11 + G, 6, 64, outside_frac=0.05)
    var = O.init_variables(k=k, mlp=(H, H, H), seed=3, bias_std=0.05, weight_gain=(600.0, 2.0, 2.0, 1.0), out_bias=1.0)
    kw = dict(k=k, V=G ** 3, sigma=1.0 / G, H=H)
    pf, ef = _run_model(pcA, pcB, var, impl, fused=True, **kw)
    ps, es = _run_model(pcA, pcB, var, impl, fused=False, **kw)
    assert torch.equal(ef["embedding_A"].fv, es["embedding_A"].fv) and torch.equal(ef["embedding_B"].fv, es["embedding_B"].fv)
    assert torch.equal(pf["pred_listAB"], ps["pred_listAB"]) and torch.equal(pf["pred_listBA"], ps["pred_listBA"])
    assert float(pf["pred_listAB"].abs().max()) > 0


def test_dense_patch_tensor_is_accepted_like_the_reference():
    pcA, pcB, _ = synthetic.uniform_batch(4, 2, 64)
    var = O.unit_scale_variables(2)
    a, b = torch.tensor(pcA, device=DEV), torch.tensor(pcB, device=DEV)
    with tf_util.use_store(_store_from(var)), tf_util.variable_scope("pc_compare", reuse=True):
        fvA = dpdist_util.get_3dmfv_tf(a, 512, 0.125, flatten=False)
        fvB = dpdist_util.get_3dmfv_tf(b, 512, 0.125, flatten=False)
        eA, C = dpdist_util.local_z(fvA, False, NUM_DIMS=3, k=5)
        eB, _ = dpdist_util.local_z(fvB, False, NUM_DIMS=3, k=5)
        lazy = dpdist_util.DPDist(a, b, eA, eB, C, False, bn=False, NUM_DIMS=3, mlp=[1024] * 3, k=5, reuse=True)
        dense = dpdist_util.DPDist(a, b, eA.materialize(), eB.materialize(), C, False, bn=False, NUM_DIMS=3,
                                   mlp=[1024] * 3, k=5, reuse=True)
    assert torch.equal(lazy[0], dense[0]) and torch.equal(lazy[1], dense[1])


def test_full_size_batch_properties():
    """BASELINE config B (1024 pairs): run the whole path; check range, mask, determinism, and a slice vs the oracle."""
    pcA, pcB, _ = synthetic.uniform_batch(2, 1024, 64)
    var = O.unit_scale_variables(4)
    p, _ = _run_model(pcA, pcB, var, _lib.HEAD_AUTO)
    ab, ba = p["pred_listAB"], p["pred_listBA"]
    assert ab.shape == (1024, 64, 1, 3) and torch.isfinite(ab).all() and torch.isfinite(ba).all()
    assert float(ab.min()) >= 0 and float(ab.max()) <= 2.0
    p2, _ = _run_model(pcA, pcB, var, _lib.HEAD_AUTO)
    assert torch.equal(p2["pred_listAB"], ab)
    with O.tf_cpu_numerics():
        oab, oba = O.forward_chunked(torch.tensor(pcA[500:508]), torch.tensor(pcB[500:508]), var, chunk=8)
    assert_out_close(ab[500:508], oab)
    assert_out_close(ba[500:508], oba)
    # pairs are independent: the same pair alone gives the same answer
    p1, _ = _run_model(pcA[500:501], pcB[500:501], var, _lib.HEAD_AUTO)
    assert_out_close(p1["pred_listAB"], ab[500:501])


def test_stress_config_E_properties():
    """BASELINE config E (4096 pairs, N = NP = 512, G = 8, k = 5: 4.19 M query rows in 16 row chunks, 336 MB of
    3DmFV tensors) on one GPU.  The oracle cannot run this size; check size-independent properties and two pairs
    against the oracle."""
    rng = np.random.default_rng(3)
    B, N = 4096, 512
    pcA = rng.uniform(-0.8, 0.8, size=(B, N, 3)).astype(np.float32)
    pcB = rng.uniform(-0.8, 0.8, size=(B, N, 3)).astype(np.float32)
    pcB[:, :8, 0] = 1.25                                   # 8 queries per cloud outside the unit cube
    var = O.unit_scale_variables(4)
    p, emb = _run_model(pcA, pcB, var, _lib.HEAD_AUTO)
    ab, ba = p["pred_listAB"], p["pred_listBA"]
    assert ab.shape == (B, N, 1, 3) and ba.shape == (B, N, 1, 3)
    assert torch.isfinite(ab).all() and torch.isfinite(ba).all()
    assert float(ab.min()) >= 0 and float(ab.max()) <= 2.0 and float(ba.max()) <= 2.0
    assert float(ab[:, :8].abs().max()) == 0.0             # masked queries (:697-698); pcB queries A's field
    assert float(ab[:, 8:].max()) > 0.0
    fv = emb["embedding_A"].fv
    assert fv.shape == (B, 512, 20)
    # every channel of every cloud is L2-normalised over the Gaussians (:124-126)
    nrm = fv.double().square().sum(1).sqrt()
    assert float((nrm - 1).abs().max()) < 1e-5
    # pairs are independent and row chunks are invisible: pairs from the first, a middle and the last chunk alone
    for i in (0, 2049, B - 1):
        p1, _ = _run_model(pcA[i:i + 1], pcB[i:i + 1], var, _lib.HEAD_AUTO)
        assert_out_close(p1["pred_listAB"], ab[i:i + 1], "pair %d alone" % i)
        assert_out_close(p1["pred_listBA"], ba[i:i + 1], "pair %d alone" % i)
    with O.tf_cpu_numerics():
        oab, oba = O.forward_chunked(torch.tensor(pcA[2049:2051]), torch.tensor(pcB[2049:2051]), var, chunk=1)
    assert_out_close(ab[2049:2051], oab, "config E vs oracle (AB)")
    assert_out_close(ba[2049:2051], oba, "config E vs oracle (BA)")


# ------------------------------------------------------------------ size-independent properties
def test_fv_permutation_and_mirror_properties():
    rng = np.random.default_rng(21)
    pts = rng.uniform(-0.8, 0.8, size=(5, 64, 3)).astype(np.float32)
    x = torch.tensor(pts, device=DEV)
    fv = dpdist_util.get_3dmfv_tf(x, n_gaussians=512, sigma=0.125, flatten=False)
    # (1) the encoding is a set function: max / min statistics do not depend on the point order at all, means only
    #     through fp32 summation order (the per-channel L2 norm couples the channels' Gaussians, hence a tolerance)
    perm = torch.tensor(rng.permutation(64), device=DEV)
    fvp = dpdist_util.get_3dmfv_tf(x[:, perm].contiguous(), n_gaussians=512, sigma=0.125, flatten=False)
    assert_close(fvp, fv, 1e-5, 2e-6, "fv of permuted points")
    mm = [1, 5, 6, 7, 8, 9, 10, 14, 15, 16, 17, 18, 19]              # max / min channels (:134-137)
    assert torch.equal(fvp[:, :, mm], fv[:, :, mm])
    # (2) mirroring the cloud in x mirrors the grid in i1 (centre x = l[i1], :47-48), negates the d/dmu_x channels and
    #     swaps their max and min; every other channel is unchanged (up to the summation order of the 8 cell weights).
    xm = x.clone()
    xm[:, :, 0] = -xm[:, :, 0]
    fvm = dpdist_util.get_3dmfv_tf(xm, n_gaussians=512, sigma=0.125, flatten=False).view(5, 8, 8, 8, 20).flip(2).reshape(5, 512, 20)
    same = [0, 1, 3, 4, 6, 7, 9, 10] + list(range(11, 20))
    assert_close(fvm[:, :, same], fv[:, :, same], 1e-5, 2e-6, "mirror: unchanged channels")
    assert_close(fvm[:, :, 2], -fv[:, :, 2], 1e-5, 2e-6, "mirror: mean d/dmu_x")
    assert_close(fvm[:, :, 5], -fv[:, :, 8], 1e-5, 2e-6, "mirror: max d/dmu_x <-> -min")
    assert_close(fvm[:, :, 8], -fv[:, :, 5], 1e-5, 2e-6, "mirror: min d/dmu_x <-> -max")


@pytest.mark.parametrize("impl", IMPLS)
def test_head_rows_are_independent(impl):
    """Permuting the queries of a cloud permutes its outputs bit-exactly; a query outside the cube gives exactly 0."""
    pcA, pcB, _ = synthetic.uniform_batch(12, 3, 64, outside_frac=0.1)
    var = O.unit_scale_variables(6)
    p, _ = _run_model(pcA, pcB, var, impl)
    perm = np.random.default_rng(0).permutation(64)
    p2, _ = _run_model(pcA, pcB[:, perm], var, impl)
    assert torch.equal(p2["pred_listAB"], p["pred_listAB"][:, perm])
    outside = torch.tensor((np.abs(pcB) > 1).any(-1))
    assert outside.any() and float(p["pred_listAB"][outside].abs().max()) == 0.0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process():
    """One process per GPU is the deployment model, but nothing may be tied to device 0 (opt-in shared-memory sizes,
    packed-weight caches, streams)."""
    pcA, pcB, _ = synthetic.uniform_batch(5, 4, 64)
    var = O.unit_scale_variables(4)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        store = tf_util.VariableStore(device=dev)
        store.load_state_dict(var, strict=False)
        a = torch.tensor(pcA, device=dev, requires_grad=True)
        with tf_util.use_store(store):
            pred, _, _ = MODEL.get_model(a, torch.tensor(pcB, device=dev), False, bn=0, Embedding_Size=512, k=5,
                                         sigma3dmfv=0.125, reuse=True)
        loss = pred["pred_listAB"][..., 0].mean() + pred["pred_listBA"][..., 0].mean()
        loss.backward()
        torch.cuda.synchronize(dev)
        outs.append((pred["pred_listAB"].detach().cpu(), a.grad.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("impl", IMPLS)
def test_inference_batch_norm_folds_into_the_layers(impl):
    """bn=True, is_training=False: the moving statistics of tf.contrib.layers.batch_norm applied after every conv + bias
    (utils/tf_util.py:221-224) -- folded into the packed weights here, applied literally in the oracle."""
    pcA, pcB, _ = synthetic.uniform_batch(14, 3, 64)
    var = dict(O.unit_scale_variables(7))
    var.update(O.bn_inference_variables(3))
    with O.tf_cpu_numerics():
        want, _, _ = O.get_model(torch.tensor(pcA), torch.tensor(pcB), var, bn=True)
    dpdist_util.HEAD_IMPL = impl
    try:
        store = _store_from(var)
        with tf_util.use_store(store):
            p, _, _ = MODEL.get_model(torch.tensor(pcA, device=DEV), torch.tensor(pcB, device=DEV), False, bn=1,
                                      Embedding_Size=512, k=5, sigma3dmfv=0.125, reuse=True)
    finally:
        dpdist_util.HEAD_IMPL = _lib.HEAD_AUTO
    assert_out_close(p["pred_listAB"], want["pred_listAB"], "bn inference AB")
    assert_out_close(p["pred_listBA"], want["pred_listBA"], "bn inference BA")
    names = [n for n in store.names() if "/bn/" in n]
    assert len(names) == 16 and not store.vars[O.VAR_PREFIX + "mapper_conv1/bn/moving_mean"].requires_grad


# ------------------------------------------------------------------ points far outside the cube
def test_fv_far_points_get_the_limit_value_not_nan():
    """0/0 policy.  The reference evaluates exp() unshifted (utils/dpdist_util.py:69-74): for a point farther than
    ~13 sigma from every Gaussian all 512 densities underflow, Q = 0/0 = NaN and the whole cloud's FV is NaN (the fp32
    oracle reproduces that).  The G = 8 kernel shifts every axis by its smallest exponent before exp(), which is the
    same number wherever the reference is finite and the mathematical limit where it is not: checked against the fp64
    twin of the oracle, in which nothing underflows at these distances."""
    rng = np.random.default_rng(8)
    pts = rng.uniform(-0.8, 0.8, size=(4, 64, 3)).astype(np.float32)
    pts[0, 0] = [3.0, 3.0, 3.0]           # joint underflow in the reference: NaN there
    pts[1, :3] = [[2.9, 0.1, -0.2], [-3.5, -3.5, 0.0], [0.0, 0.0, 4.5]]
    pts[2, 5] = [1.9, 1.9, 1.9]           # every axis alone is fine, the product underflows in fp32
    with O.tf_cpu_numerics():
        ref32 = O.get_3dmfv(torch.tensor(pts), 512, 0.125, flatten=False)
    assert not torch.isfinite(ref32[0]).all()                      # the reference's behaviour, for the record
    got = dpdist_util.get_3dmfv_tf(torch.tensor(pts, device=DEV), n_gaussians=512, sigma=0.125, flatten=False)
    assert torch.isfinite(got).all()
    want = O.get_3dmfv(torch.tensor(pts, dtype=torch.float64), 512, 0.125, flatten=False)
    assert torch.isfinite(want).all()
    assert_fv_close(got, want, "fv with far points vs the fp64 twin")
    assert_fv_close(got[3], ref32[3], "an ordinary cloud in the same batch")


def test_far_points_generic_kernel_and_backward_follow_the_same_policy():
    rng = np.random.default_rng(9)
    pts = rng.uniform(-0.8, 0.8, size=(2, 40, 3)).astype(np.float32)
    pts[0, 0] = [3.0, -3.0, 2.5]
    got = dpdist_util.get_3dmfv_tf(torch.tensor(pts, device=DEV), n_gaussians=125, sigma=0.2, flatten=False)
    want = O.get_3dmfv(torch.tensor(pts, dtype=torch.float64), 125, 0.2, flatten=False)
    assert torch.isfinite(got).all()
    assert_fv_close(got, want, "generic kernel with a far point vs the fp64 twin")
    for V, sigma in ((512, 0.125), (125, 0.2)):                    # gradient into the clouds stays finite as well
        x = torch.tensor(pts, device=DEV, requires_grad=True)
        fv = dpdist_util.get_3dmfv_tf(x, n_gaussians=V, sigma=sigma, flatten=False)
        g = torch.Generator(device="cpu").manual_seed(1)
        (fv * torch.randn(fv.shape, generator=g).to(DEV)).sum().backward()
        assert torch.isfinite(x.grad).all()
