"""CPU tests of the host-side mirror of the reference interface (names, shapes, variable names)."""
import inspect

import numpy as np
import pytest
import torch

from dpdist_b200 import dpdist_and_aue as MODEL
from dpdist_b200 import dpdist_util, synthetic, tf_util
from oracle import dpdist_oracle as O


def test_signatures_match_the_reference():
    # models/dpdist_and_aue.py:31-35
    sig = inspect.signature(MODEL.get_model)
    names = list(sig.parameters)
    assert names[:17] == ["pcA", "pcB", "is_training", "bn_decay", "wd", "bn", "Embedding_Size", "pn", "sig", "k",
                          "overlap", "localSNmlp", "full_fv", "sigma3dmfv", "conv_version", "add_noise", "reuse"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert d["Embedding_Size"] == 512 and d["k"] == 0 and d["sigma3dmfv"] == 0.125 and d["bn"] is True
    # utils/dpdist_util.py:22, :850, :412-416, :962, :982
    d = {k: v.default for k, v in inspect.signature(dpdist_util.get_3dmfv_tf).parameters.items()}
    assert d == {"points": inspect._empty, "n_gaussians": 9, "sigma": 0.0625, "flatten": True, "normalize": True, "full_fv": True}
    assert list(inspect.signature(dpdist_util.local_z).parameters) == ["net", "is_training", "reuse", "NUM_DIMS", "k", "overlap"]
    assert list(inspect.signature(dpdist_util.DPDist).parameters) == [
        "point_cloud", "point_cloudB", "embedding", "embeddingB", "C", "is_training", "bn_decay", "reuse", "bn", "wd",
        "sig", "Embedding_Size", "NUM_DIMS", "mlp", "k", "conv_version", "output_act"]
    assert list(inspect.signature(MODEL.get_loss).parameters) == ["pred_set", "end_points", "labels", "loss_type"]
    assert list(inspect.signature(MODEL.placeholder_inputs).parameters)[:3] == ["batch_size", "num_point", "NUM_DIMS"]


def test_placeholders():
    a, b, lab, lba = MODEL.placeholder_inputs(16, 64, NUM_DIMS=3, device="cpu")
    assert a.shape == (16, 64, 3) and b.shape == (16, 64, 3) and lab.shape == (16, 64) and lba.shape == (16, 64)


def test_grid_tables_match_reference_construction():
    for V in (512, 125, 27, 64):
        X, Y, Z = dpdist_util.get_grid_centers(V, 3)
        Xo, Yo, Zo = O.get_grid_centers(V, 3)
        assert np.array_equal(X, Xo) and np.array_equal(Z, Zo)
        C = np.stack([X, Y, Z], -1).astype(np.float32).reshape(-1, 3)
        G, l, lo, hi = dpdist_util._assign_tables(C)
        assert G ** 3 == V
        gs = np.float32(abs(C[0][2] - C[1][2]) / np.float32(2))
        assert np.array_equal(lo, (l - gs).astype(np.float32)) and np.array_equal(hi, (l + gs).astype(np.float32))
        # FV grid (linspace form, :42) and voxel centres (arange form, :987) coincide in fp32 at these sizes
        Gf, lf = dpdist_util._fv_grid(V)
        assert Gf == G and np.allclose(lf, l, atol=1e-7)
    with pytest.raises(ValueError):
        dpdist_util._fv_grid(500)
    with pytest.raises(ValueError):
        dpdist_util._assign_tables(np.random.default_rng(0).random((27, 3)).astype(np.float32))


def test_variable_names_shapes_and_reuse():
    store = tf_util.VariableStore(device="cpu", seed=0)
    with tf_util.use_store(store), tf_util.variable_scope("pc_compare"), tf_util.variable_scope("dpdist_local"):
        w1, b1 = tf_util.conv2d_variables(1, 1024, [1, 2503], "mapper_conv1")
        w2, b2 = tf_util.conv2d_variables(1024, 1024, [1, 1], "mapper_conv2")
        tf_util.conv2d_variables(1024, 1024, [1, 1], "mapper_conv3")
        w4, b4 = tf_util.conv2d_variables(1024, 3, [1, 1], "mapper_conv4")
        w1b, _ = tf_util.conv2d_variables(1, 1024, [1, 2503], "mapper_conv1", reuse=True)
    assert w1b is w1
    assert sorted(store.names()) == sorted(O.init_variables().keys())
    assert tuple(w1.shape) == (1, 2503, 1, 1024) and tuple(w2.shape) == (1, 1, 1024, 1024) and tuple(w4.shape) == (1, 1, 1024, 3)
    assert float(b1.abs().max()) == 0.0 and float(b4.abs().max()) == 0.0          # tf.constant_initializer(0.0)
    assert float(w1.abs().max()) <= np.sqrt(6.0 / (2503 + 2503 * 1024)) + 1e-9    # TF fan computation
    assert float(w2.abs().max()) <= np.sqrt(6.0 / 2048) + 1e-9
    assert sum(v.numel() for v in store.trainable_variables("pc_compare")) == 4666371
    with tf_util.use_store(tf_util.VariableStore(device="cpu")), tf_util.variable_scope("pc_compare", reuse=True):
        with pytest.raises(ValueError):
            tf_util.conv2d_variables(1, 8, [1, 4], "mapper_conv1")
    sd = store.state_dict()
    s2 = tf_util.VariableStore(device="cpu")
    s2.load_state_dict(sd)
    assert torch.equal(s2.vars["pc_compare/dpdist_local/mapper_conv2/weights"], w2)


def test_get_loss_returns_what_the_reference_returns():
    tf_util.clear_collections()
    g = torch.Generator().manual_seed(0)
    ab, ba = torch.rand(4, 8, 1, 3, generator=g), torch.rand(4, 8, 1, 3, generator=g)
    lab = torch.rand(4, 8, generator=g)
    ls, lp = MODEL.get_loss({"pred_listAB": ab, "pred_listBA": ba}, {}, lab)
    assert ls.shape == (4, 8) and torch.equal(ls, ab[:, :, 0, 0])          # utils/dpdist_util.py:967-968,980
    o_loss, o_lp = O.get_loss({"pred_listAB": ab, "pred_listBA": ba}, {}, lab)
    assert torch.allclose(tf_util.get_collection("loss_samples")[-1], o_loss)
    assert torch.allclose(lp, o_lp) and torch.allclose(tf_util.get_collection("loss_pred")[-1], o_lp)


def test_synthetic_shapes_follow_the_dataset():
    pts, lab = synthetic.dataset_batch(0, 2, num_point=64)
    assert pts.shape == (2, 384, 3) and lab.shape == (2, 256)           # modelnet_dataset.py:177-178
    pcA, pcB, l = synthetic.anchor_pair(0)
    assert pcA.shape == (1, 64, 3) and pcB.shape == (1, 64, 3) and l.shape == (1, 64)
    assert np.all(l[:, :32] == 0) and np.all(l[:, 32:] > 0)             # train...py:759-761
    a, b, _ = synthetic.uniform_batch(2, 8, 64)
    assert (np.abs(b) > 1).any()
    a2, _, _ = synthetic.uniform_batch(2, 8, 64)
    assert np.array_equal(a, a2)


def test_pcrnet_host_geometry():
    """helper.py:539-570 / :229-262 / :309-329 restated: quaternion -> rotation, Euler poses, transform accumulation."""
    import torch
    from dpdist_b200 import pcrnet_ours as P
    rng = np.random.default_rng(0)
    ang = 0.7
    q = torch.tensor([[np.cos(ang / 2), 0.0, 0.0, np.sin(ang / 2)]], dtype=torch.float32)     # rotation about z
    pts = torch.tensor(rng.normal(size=(1, 5, 3)), dtype=torch.float32)
    t = torch.tensor([[0.1, -0.2, 0.3]])
    got = P.transformation_quat_tensor(pts, q, t)
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    np.testing.assert_allclose(got[0].numpy(), pts[0].numpy() @ Rz.T + t.numpy(), atol=1e-6)
    poses = P.generate_poses(8, rng)
    assert np.abs(poses[:, :3]).max() <= 0.01 and np.abs(poses[:, 3:]).max() <= np.pi / 4
    R = P.euler_to_matrix(poses)
    np.testing.assert_allclose(np.einsum("bij,bkj->bik", R, R), np.tile(np.eye(3), (8, 1, 1)), atol=1e-12)
    moved = P.apply_transformation(rng.normal(size=(8, 4, 3)), poses)
    assert moved.dtype == np.float32 and moved.shape == (8, 4, 3)
    # compose: T <- M(pose) @ T, un-normalised quaternion is normalised (transforms3d.quat2mat behaviour)
    pose7 = torch.tensor([[0.1, 0.2, 0.3, 2 * np.cos(ang / 2), 0, 0, 2 * np.sin(ang / 2)]], dtype=torch.float32)
    T, Rm = P.compose(torch.eye(4)[None], pose7)
    np.testing.assert_allclose(Rm[0].numpy(), Rz, atol=1e-6)
    np.testing.assert_allclose(T[0, :3, 3].numpy(), [0.1, 0.2, 0.3], atol=1e-7)
    assert abs(float(P.normalize_quat(pose7[:, 3:]).norm()) - 1) < 1e-6
