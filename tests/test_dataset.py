"""Dataset reader (reference file formats) -- CPU part: a tiny fake ModelNet tree in tmp_path."""
import os

import numpy as np
import pytest

from dpdist_b200 import modelnet_dataset as MD


def make_tree(root, n_train=5, n_test=2, n_surface=300, n_neg=400, seed=0):
    rng = np.random.default_rng(seed)
    os.makedirs(root / "chair")
    os.makedirs(root / "table")
    (root / "modelnet40_shape_names.txt").write_text("chair\ntable\n")
    ids = {"train": ["chair_%04d" % i for i in range(n_train)] + ["table_0001"],
           "test": ["chair_%04d" % (100 + i) for i in range(n_test)]}
    for split, lst in ids.items():
        (root / ("modelnet40_%s.txt" % split)).write_text("\n".join(lst) + "\n")
        for x in lst:
            name = "_".join(x.split("_")[:-1])
            base = str(root / name / x)
            surf = rng.uniform(-0.8, 0.8, size=(n_surface, 3))
            np.savetxt(base + ".txt", np.concatenate([surf / 0.8, rng.normal(size=(n_surface, 3))], 1), fmt="%.6f", delimiter=",")
            np.savetxt(base + "_dist_c_scaled.txt", surf, fmt="%.6f", delimiter=",")
            for kind, lo, hi in (("l", 0.001, 0.1), ("u", 0.1, 1.0)):
                pts = rng.uniform(-1, 1, size=(n_neg, 3))
                d = rng.uniform(lo, hi, size=(n_neg, 1))
                np.savetxt(base + "_10000_dist_c_neg_%s.txt" % kind, np.concatenate([pts, d], 1), fmt="%.6f", delimiter=",")
    return ids


def test_reader_follows_the_reference_layout(tmp_path):
    ids = make_tree(tmp_path)
    ds = MD.ModelNetDataset(root=str(tmp_path), npoints=128, split="train", batch_size=4, class_choice=["chair"], seed=1)
    assert len(ds) == 5 and ds.num_channel() == 3                       # the table is filtered out (:55-66)
    assert ds.num_batches == 2
    pts, cls, lab = ds[0]
    assert pts.shape == (3 * 128, 3) and lab.shape == (2 * 128,) and pts.dtype == np.float32
    base = ds.datapath[0][1][:-4]
    surf = np.loadtxt(base + "_dist_c_scaled.txt", delimiter=",").astype(np.float32)
    negl = np.loadtxt(base + "_10000_dist_c_neg_l.txt", delimiter=",").astype(np.float32)
    negu = np.loadtxt(base + "_10000_dist_c_neg_u.txt", delimiter=",").astype(np.float32)
    # first access: surface[:np] | near[:np] | a random subset of the far set, labels = their distances (:136-139)
    assert np.array_equal(pts[:128], surf[:128]) and np.array_equal(pts[128:256], negl[:128, :3])
    assert np.array_equal(lab[:128], negl[:128, 3])
    far = {tuple(r) for r in negu.round(6)}
    assert all(tuple(r) in far for r in np.concatenate([pts[256:], lab[128:, None]], 1).round(6))
    # cached access: ONE permutation for the three point sets and both label sets (:99-110)
    pts2, _, lab2 = ds[0]
    perm = [int(np.where((pts[:128] == p).all(1))[0][0]) for p in pts2[:128]]
    assert sorted(perm) == list(range(128))
    assert np.array_equal(pts2[128:256], pts[128:256][perm]) and np.array_equal(pts2[256:], pts[256:][perm])
    assert np.array_equal(lab2[:128], lab[:128][perm]) and np.array_equal(lab2[128:], lab[128:][perm])
    # batches: smaller last batch, epoch end, reset (:170-187)
    n = []
    while ds.has_next_batch():
        d, l = ds.next_batch(augment=True)
        assert d.shape[1:] == (384, 3) and l.shape[1:] == (256,)
        n.append(d.shape[0])
    assert n == [4, 1]
    ds.reset()
    assert ds.has_next_batch()
    test = MD.ModelNetDataset(root=str(tmp_path), npoints=128, split="test", batch_size=4, class_choice=["chair"])
    assert len(test) == 2 and test.shuffle is False


def test_augmentation_is_rigid_per_shape():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(3, 50, 3))
    y = MD.shift_point_cloud(MD.rotate_point_cloud(x.copy(), rng), rng=rng)
    for k in range(3):
        d0 = np.linalg.norm(x[k][:, None] - x[k][None], axis=-1)
        d1 = np.linalg.norm(y[k][:, None] - y[k][None], axis=-1)
        assert np.allclose(d0, d1, atol=1e-5)                           # rotation + shift preserve distances
        assert np.allclose((y[k] - y[k].mean(0))[:, 1], (x[k] - x[k].mean(0))[:, 1], atol=1e-5)   # about the up (y) axis
