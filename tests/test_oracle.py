"""CPU tests of the oracle itself: fp64 twin, independent separable formulation, the reference's
own explicit patch loop, analytic invariants, and the committed golden fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import dpdist_oracle as O
from oracle.fv_separable_np import fv_separable
from dpdist_b200 import synthetic as S
from tolerances import assert_fv_close, assert_out_close

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _cloud(seed, B=2, N=64, scale=0.8):
    rng = np.random.default_rng(seed)
    return rng.uniform(-scale, scale, size=(B, N, 3)).astype(np.float32)


@pytest.mark.parametrize("G,N,sigma", [(8, 64, 0.125), (5, 33, 0.2), (3, 7, 0.25), (8, 1, 0.125)])
@pytest.mark.parametrize("full_fv", [True, False])
def test_fv_literal_matches_separable_and_fp64(G, N, sigma, full_fv):
    pts = _cloud(G * 100 + N, 2, N)
    t = torch.tensor(pts)
    with O.tf_cpu_numerics():
        lit32 = O.get_3dmfv(t, G ** 3, sigma, flatten=False, full_fv=full_fv)
    lit64 = O.get_3dmfv(t.double(), G ** 3, sigma, flatten=False, full_fv=full_fv)
    sep64 = fv_separable(pts, G ** 3, sigma, full_fv=full_fv)
    assert lit32.shape == (2, G ** 3, 20 if full_fv else 7)
    assert_fv_close(lit32, lit64, "literal fp32 vs fp64")
    assert_fv_close(sep64, lit64, "separable vs literal (fp64)")


def test_fv_flatten_layout():
    pts = torch.tensor(_cloud(3, 2, 16))
    a = O.get_3dmfv(pts, 27, 0.25, flatten=False)
    b = O.get_3dmfv(pts, 27, 0.25, flatten=True)
    assert b.shape == (2, 20 * 27)
    # flatten: [pi channels x V | mu channels x V | sigma channels x V], channel-major (:129-132)
    assert torch.equal(b.view(2, 20, 27), a.transpose(1, 2))


def test_fv_invariants():
    G, N = 8, 64
    pts = _cloud(11, 1, N)
    fv = O.get_3dmfv(torch.tensor(pts).double(), G ** 3, 0.125, flatten=False)[0]
    # every channel is L2-normalised over the V Gaussians
    assert torch.allclose((fv * fv).sum(0), torch.ones(20, dtype=torch.float64), atol=1e-9)
    # permutation invariance w.r.t. point order
    perm = np.random.default_rng(0).permutation(N)
    fv_p = O.get_3dmfv(torch.tensor(pts[:, perm]).double(), G ** 3, 0.125, flatten=False)[0]
    assert torch.allclose(fv, fv_p, atol=1e-12)
    # mirror x -> -x: Gaussian (i0,i1,i2) <-> (i0,G-1-i1,i2); d_mu_x channels flip sign and swap max<->min
    m = pts.copy(); m[..., 0] *= -1
    fv_m = O.get_3dmfv(torch.tensor(m).double(), G ** 3, 0.125, flatten=False)[0].view(G, G, G, 20).flip(1).reshape(-1, 20)
    assert torch.allclose(fv_m[:, 2], -fv[:, 2], atol=1e-9)      # mu mean x
    assert torch.allclose(fv_m[:, 5], -fv[:, 8], atol=1e-9)      # mu max x <-> -mu min x
    assert torch.allclose(fv_m[:, 3], fv[:, 3], atol=1e-9)       # mu mean y unchanged
    assert torch.allclose(fv_m[:, 11], fv[:, 11], atol=1e-9)     # sigma mean x unchanged


def test_fv_axis_order_meshgrid_xy():
    # a single point sitting on Gaussian (i0,i1,i2)'s centre: mu = (l[i1], l[i0], l[i2])
    G = 4
    l = np.linspace(-1, 1, G, False) + 1 / G
    i0, i1, i2 = 1, 3, 2
    pt = np.array([[[l[i1], l[i0], l[i2]]]], dtype=np.float32)
    fv = O.get_3dmfv(torch.tensor(pt).double(), G ** 3, 0.125, flatten=False)[0]
    g = i0 * G * G + i1 * G + i2
    assert int(torch.argmax(fv[:, 1])) == g          # pi max peaks at the Gaussian the point sits on
    assert abs(float(fv[g, 2])) < 1e-9 and abs(float(fv[g, 3])) < 1e-9 and abs(float(fv[g, 4])) < 1e-9


@pytest.mark.parametrize("G,k", [(8, 5), (5, 3), (4, 1), (3, 5)])
def test_patches_unfold_equals_reference_loop(G, k):
    net = torch.randn(2, G ** 3, 6, generator=torch.Generator().manual_seed(G + k))
    a, Ca = O.local_z_3d(net, k=k, explicit_loop=False)
    b, Cb = O.local_z_3d(net, k=k, explicit_loop=True)
    assert a.shape == (2, G ** 3, k ** 3 * 6)
    assert torch.equal(a, b)
    assert torch.equal(Ca, Cb)
    pb = (k - 1) // 2
    centre = (pb * k + pb) * k + pb
    assert torch.equal(a[:, :, centre * 6:(centre + 1) * 6], net)


def test_voxel_assignment_semantics():
    G = 8
    X, Y, Z = O.get_grid_centers(G ** 3, 3)
    C = torch.tensor(np.stack([X, Y, Z], -1).astype(np.float32).reshape(-1, 3))
    pc = torch.tensor([[[0.1, -0.3, 0.6],      # interior
                        [-0.75, -0.75, -0.75], # an edge between cells: (lo, hi] -> lower cell
                        [1.0, 1.0, 1.0],       # upper boundary is inside (<=)
                        [-1.0, 0.0, 0.0],      # lower boundary is outside (>)
                        [1.5, 0.0, 0.0]]],     # outside
                      dtype=torch.float32)
    bv, off, am = O.get_pc_grid_binary_mask_from_centers(C, pc)
    inside = bv.sum(-1)[0]
    assert inside.tolist() == [1, 1, 1, 0, 0]
    assert am[0, 3] == 0 and am[0, 4] == 0
    l = np.arange(-1, 1, 0.25) + 0.125
    g = int(am[0, 0]); i0, i1, i2 = g // 64, (g // 8) % 8, g % 8
    assert abs(l[i1] - 0.1) <= 0.125 and abs(l[i0] + 0.3) <= 0.125 and abs(l[i2] - 0.6) <= 0.125
    g = int(am[0, 1]); assert (g // 64, (g // 8) % 8, g % 8) == (0, 0, 0)
    g = int(am[0, 2]); assert (g // 64, (g // 8) % 8, g % 8) == (7, 7, 7)


def test_head_mask_and_range():
    pcA, pcB, _ = S.uniform_batch(5, 2, 64, outside_frac=0.2)
    var = O.unit_scale_variables(3)
    p, aux, _ = O.get_model(torch.tensor(pcA), torch.tensor(pcB), var)
    for key, q in (("pred_listAB", pcB), ("pred_listBA", pcA)):
        out = p[key]
        assert out.shape == (2, 64, 1, 3)
        assert float(out.min()) >= 0.0 and float(out.max()) <= 2.0
        outside = (np.abs(q) > 1.0).any(-1) | (q <= -1.0).any(-1)
        assert outside.any()
        assert float(out[torch.tensor(outside)].abs().max()) == 0.0


def test_xavier_limits_follow_tf_fans():
    var = O.init_variables(seed=0)
    w1 = var[O.VAR_PREFIX + "mapper_conv1/weights"]
    assert tuple(w1.shape) == (1, 2503, 1, 1024)
    lim = np.sqrt(6.0 / (2503 + 2503 * 1024))
    assert float(w1.abs().max()) <= lim and float(w1.abs().max()) > 0.99 * lim
    n = sum(v.numel() for v in var.values())
    assert n == 4666371


def test_loss_definition():
    B, NP = 3, 8
    g = torch.Generator().manual_seed(0)
    ab, ba = torch.rand(B, NP, 1, 3, generator=g), torch.rand(B, NP, 1, 3, generator=g)
    lab = torch.rand(B, NP, generator=g)
    loss, lp = O.get_loss({"pred_listAB": ab, "pred_listBA": ba}, {}, lab)
    assert abs(float(loss) - float((ab[..., 0, 0] - lab).abs().mean())) < 1e-7
    assert abs(float(lp) - 0.5 * float(ab[..., 0].mean() + ba[..., 0].mean())) < 1e-7


def test_golden_fixture_pins_the_oracle():
    """tests/golden/anchor_A.npz was written by tests/golden/make_golden.py from this oracle (config A
    of BASELINE.json).  It pins the restatement against drift; it is NOT a TF1 output (parity unpinned)."""
    z = np.load(os.path.join(GOLDEN, "anchor_A.npz"))
    var = O.unit_scale_variables(int(z["weight_seed"]))
    chk = float(sum(v.double().abs().sum() for v in var.values()))
    assert abs(chk - float(z["weight_checksum"])) <= 1e-6 * chk, "torch RNG drifted: regenerate the fixture"
    with O.tf_cpu_numerics():
        p, aux, _ = O.get_model(torch.tensor(z["pcA"]), torch.tensor(z["pcB"]), var)
    assert_fv_close(aux["fvA"], z["fvA"], "fvA")
    assert_fv_close(aux["fvB"], z["fvB"], "fvB")
    assert_out_close(p["pred_listAB"], z["pred_AB"], "pred_AB")
    assert_out_close(p["pred_listBA"], z["pred_BA"], "pred_BA")
    _, _, am = O.get_pc_grid_binary_mask_from_centers(aux["C"], torch.tensor(z["pcB"]))
    assert np.array_equal(am.numpy().astype(np.int32), z["idx_B"])


def test_fv_is_a_set_function_hypothesis():
    """Property test (hypothesis): the oracle's encoding does not depend on the point order (fp64: to rounding)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=10, deadline=None)
    @given(st.integers(0, 10 ** 6), st.integers(2, 40))
    def check(seed, n):
        rng = np.random.default_rng(seed)
        pts = torch.tensor(rng.uniform(-0.9, 0.9, size=(1, n, 3)))
        perm = torch.tensor(rng.permutation(n))
        a = O.get_3dmfv(pts, 27, 0.3, flatten=False)
        b = O.get_3dmfv(pts[:, perm], 27, 0.3, flatten=False)
        assert float((a - b).abs().max()) < 1e-12
    check()
