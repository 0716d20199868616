"""GPU tests of the data side: ground-truth distances and batch assembly + augmentation against the
reference's numpy / scipy formulation restated in the oracle."""
import os

import numpy as np
import pytest
import torch

from dpdist_b200 import _lib, data as D, synthetic
from oracle import dpdist_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,S,Q", [(1, 10000, 20000), (3, 777, 1025), (2, 1, 5), (4, 1024, 1)])
def test_nearest_distance_matches_cdist_min(B, S, Q):
    rng = np.random.default_rng(S + Q)
    surf = (rng.uniform(-1, 1, size=(B, S, 3)) * 0.8).astype(np.float32)
    qry = rng.uniform(-1, 1, size=(B, Q, 3)).astype(np.float32)
    if Q > 2:
        qry[0, 2] = surf[0, S // 2]                                     # a query exactly on the surface sample: distance 0
    dist, arg = D.nearest_distance(torch.tensor(surf, device=DEV), torch.tensor(qry, device=DEV), return_index=True)
    dist2 = D.nearest_distance(torch.tensor(surf, device=DEV), torch.tensor(qry, device=DEV))
    assert torch.equal(dist, dist2)
    for b in range(B):
        want, warg = O.nearest_distance(surf[b], qry[b])
        got = dist[b].cpu().double().numpy()
        # the generator writes %.6f (dataset_sample_with_gt.py:124-126); fp32 on d^2 gives ~1e-7 relative
        assert np.abs(got - want).max() <= 1e-6 * np.maximum(want, 1e-3).max(), np.abs(got - want).max()
        a = arg[b].cpu().numpy()
        # the index may differ only where two surface points are equidistant to fp32 rounding
        dsel = np.linalg.norm(surf[b][a].astype(np.float64) - qry[b].astype(np.float64), axis=1)
        assert np.abs(dsel - want).max() <= 1e-6
        assert (a == warg).mean() > 0.999
    if Q > 2:
        assert float(dist[0, 2]) == 0.0 and int(arg[0, 2]) == S // 2


def test_generate_points_with_gt_sets():
    S, near, far, _, _ = synthetic.chair_item(3, 64, dense=4096)
    surf = torch.tensor(synthetic._sample_box_surface(np.random.default_rng(0), 5000).astype(np.float32), device=DEV)
    g = torch.Generator(device=DEV).manual_seed(0)
    neg_l, neg_u = D.generate_points_with_gt(surf, num_neg_points=2000, generator=g)
    assert neg_l.shape == (2000, 4) and neg_u.shape == (2000, 4)
    assert float(neg_l[:, 3].min()) > 0.001 and float(neg_l[:, 3].max()) < 0.1            # min_eps < d < 2 eps
    assert float(neg_u[:-200, 3].min()) > 0.1
    assert float(neg_u[-200:, :3].norm(dim=1).min()) > 1.0                                # last 10 %: outside the unit ball
    assert float(neg_u[:-200, :3].norm(dim=1).max()) <= 1.0 + 1e-6
    for rows in (neg_l[:300], neg_u[-300:]):                                              # the stored distance is the GT
        want, _ = O.nearest_distance(surf.cpu().numpy(), rows[:, :3].cpu().numpy())
        assert np.abs(rows[:, 3].cpu().double().numpy() - want).max() <= 2e-6


@pytest.mark.parametrize("bsize,num_point", [(16, 64), (5, 32), (3, 512)])
def test_assemble_batch_matches_the_trainer_and_the_augmentation(bsize, num_point):
    pts, lab = synthetic.dataset_batch(4, bsize, num_point)                               # [bsize, 3*npoints, 3], [bsize, 2*npoints]
    rng = np.random.default_rng(1)
    angles = rng.uniform(0, 2 * np.pi, size=bsize).astype(np.float32)
    shifts = rng.uniform(-0.1, 0.1, size=(bsize, 3)).astype(np.float32)
    wa, wb, wl = O.assemble_batch(pts, lab, num_point)
    a, b, l = D.assemble_batch(torch.tensor(pts, device=DEV), torch.tensor(lab, device=DEV), num_point)
    assert np.array_equal(a.cpu().numpy(), wa) and np.array_equal(b.cpu().numpy(), wb)     # pure gather: bit-exact
    assert np.array_equal(l.cpu().numpy(), wl.astype(np.float32))
    aug = O.rotate_shift(pts, angles.astype(np.float64), shifts.astype(np.float64))
    wa, wb, wl = O.assemble_batch(aug, lab, num_point)
    a, b, l = D.assemble_batch(torch.tensor(pts, device=DEV), torch.tensor(lab, device=DEV), num_point,
                               angle=torch.tensor(angles, device=DEV), shift=torch.tensor(shifts, device=DEV))
    assert np.abs(a.cpu().numpy() - wa).max() <= 2e-6 and np.abs(b.cpu().numpy() - wb).max() <= 2e-6
    assert np.array_equal(l.cpu().numpy(), wl.astype(np.float32))
    # augmentation is rigid: GT distances stay valid
    d0 = torch.cdist(torch.tensor(pts[:, :num_point]), torch.tensor(pts[:, :num_point]))
    ang, sh = D.random_augmentation(bsize, DEV)
    assert ang.shape == (bsize,) and sh.shape == (bsize, 3) and float(sh.abs().max()) <= 0.1


def test_data_entry_points_validate():
    lib = _lib.load()
    x = torch.zeros((1, 12, 3), device=DEV)
    assert lib.dpd_nearest_distance(None, 1, 4, x.data_ptr(), 4, x.data_ptr(), None, None) == -1
    assert lib.dpd_assemble_batch(x.data_ptr(), x.data_ptr(), 1, 3, 4, None, None, x.data_ptr(), x.data_ptr(), x.data_ptr(), None) == -1
    assert b"npoints" in lib.dpd_last_error()


def test_generator_reader_and_trainer_on_a_tiny_dataset(tmp_path):
    """dataset_sample_with_gt -> ModelNetDataset -> next_batch_device -> training driver, on a fake ModelNet tree."""
    from test_dataset import make_tree
    from dpdist_b200 import dataset_sample_with_gt as GEN, modelnet_dataset as MD, train_multi_gpu_pc_compare_dist as DRV, tf_checkpoint
    make_tree(tmp_path, n_train=4, n_test=2, n_surface=600)
    for f in tmp_path.rglob("*_dist_c_*"):          # keep only the raw shapes: the generator must produce the rest
        f.unlink()
    n = GEN.generate_points_with_gt(str(tmp_path), num_neg_points=500, cur_cls=["chair"], seed=0, verbose=False)
    assert n == 6
    assert GEN.generate_points_with_gt(str(tmp_path), num_neg_points=500, cur_cls=["chair"], verbose=False) == 0   # already there
    base = str(tmp_path / "chair" / "chair_0000")
    raw = np.loadtxt(base + ".txt", delimiter=",")[:, :3]
    pos = np.loadtxt(base + "_dist_c_scaled.txt", delimiter=",")
    assert np.allclose(pos, raw * 0.8, atol=1e-6)
    negl = np.loadtxt(base + "_500_dist_c_neg_l.txt", delimiter=",")
    want, _ = O.nearest_distance(pos, negl[:, :3])
    assert negl.shape == (500, 4) and np.abs(negl[:, 3] - want).max() < 5e-6 and negl[:, 3].max() < 0.1
    # the reader expects the 10^4-point file names; link them for this miniature
    for f in list(tmp_path.rglob("*_500_dist_c_*")):
        os.rename(f, str(f).replace("_500_", "_10000_"))
    ds = MD.ModelNetDataset(root=str(tmp_path), npoints=128, split="train", batch_size=4, class_choice=["chair"], device=DEV, seed=3)
    pcA, pcB, lab = ds.next_batch_device(64, augment=True)
    assert pcA.shape == (4, 64, 3) and pcB.shape == (4, 64, 3) and lab.shape == (4, 64)
    assert float(lab[:, :32].abs().max()) == 0.0 and float(lab[:, 32:48].max()) < 0.1 and float(lab[:, 48:].min()) > 0.0
    # labels are distances to the (augmented) surface: rigid motion keeps them valid
    d = D.nearest_distance(torch.cat([pcA, pcB[:, :32]], 1), pcB[:, 32:])
    assert float((lab[:, 32:] - d).max()) < 1e-5      # a subsample of the surface (96 of 600 points) can only over-estimate
    DRV.main(["--data_root", str(tmp_path), "--batch_size", "4", "--max_epoch", "2", "--log_dir", str(tmp_path / "log")])
    ck = tf_checkpoint.list_variables(str(tmp_path / "log" / "model.ckpt"))
    assert "pc_compare/dpdist_local/mapper_conv1/weights" in ck
