"""Seeded synthetic inputs with the reference's shapes (no dataset ships with the reference).

Mirrors what `ModelNetDataset.next_batch` + `train_one_epoch_3d` hand to the graph
(modelnet_dataset.py:170-187; train_multi_gpu_pc_compare_dist.py:732-778):
  item   = [S (2*np surface) | near (np) | far (np)] points, labels = [gt near (np) | gt far (np)]
  pcA    = first `np` surface points
  pcB    = [np/2 of the second surface half | np/4 near-surface | np/4 far]
  labels = [0 x np/2 | gt_near | gt_far]
All generators are numpy, deterministic in `seed`, fp32.
"""
import numpy as np

# a "chair": seat, back, 4 legs (axis-aligned boxes; centre, half-extent)
_CHAIR_BOXES = np.array([
    [[0.0, 0.0, 0.0], [0.45, 0.05, 0.45]],      # seat
    [[0.0, 0.45, -0.40], [0.45, 0.45, 0.05]],   # back
    [[-0.38, -0.40, -0.38], [0.05, 0.38, 0.05]],
    [[0.38, -0.40, -0.38], [0.05, 0.38, 0.05]],
    [[-0.38, -0.40, 0.38], [0.05, 0.38, 0.05]],
    [[0.38, -0.40, 0.38], [0.05, 0.38, 0.05]],
], dtype=np.float64)


def _sample_box_surface(rng, n):
    """Uniform-ish samples on the union of the chair boxes' surfaces, scaled into radius <= 0.8
    (dataset_sample_with_gt.py:82 scales clouds by 0.8 inside the unit sphere)."""
    areas = []
    for c, h in _CHAIR_BOXES:
        areas.append(8 * (h[0] * h[1] + h[1] * h[2] + h[0] * h[2]))
    areas = np.array(areas) / np.sum(areas)
    which = rng.choice(len(_CHAIR_BOXES), size=n, p=areas)
    pts = np.empty((n, 3))
    for i, b in enumerate(which):
        c, h = _CHAIR_BOXES[b]
        fa = np.array([h[1] * h[2], h[0] * h[2], h[0] * h[1]])
        ax = rng.choice(3, p=fa / fa.sum())
        p = c + (rng.random(3) * 2 - 1) * h
        p[ax] = c[ax] + h[ax] * (1 if rng.random() < 0.5 else -1)
        pts[i] = p
    r = np.linalg.norm(_CHAIR_BOXES[:, 0] + np.sign(_CHAIR_BOXES[:, 0]) * _CHAIR_BOXES[:, 1], axis=1).max()
    return pts * (0.8 / r)


def _dist_to_set(q, s):
    d = np.sqrt(((q[:, None, :] - s[None, :, :]) ** 2).sum(-1))
    return d.min(1)


def chair_item(seed=0, num_point=64, dense=2048):
    """One dataset item shaped like modelnet_dataset.py:136-139 with npoints = 2*num_point:
    points [3*npoints... ] -> returns (S [2np,3], near [np,3], far [np,3], gt_near [np], gt_far [np])."""
    rng = np.random.default_rng(seed)
    surf_dense = _sample_box_surface(rng, dense)
    S = _sample_box_surface(rng, 2 * num_point)
    base = _sample_box_surface(rng, num_point)
    dirs = rng.normal(size=(num_point, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    near = base + dirs * rng.uniform(0.001, 0.1, size=(num_point, 1))
    far = rng.normal(size=(num_point, 3))
    far = far / np.linalg.norm(far, axis=1, keepdims=True) * rng.random((num_point, 1)) ** (1 / 3)
    return (S.astype(np.float32), near.astype(np.float32), far.astype(np.float32),
            _dist_to_set(near, surf_dense).astype(np.float32), _dist_to_set(far, surf_dense).astype(np.float32))


def assemble_pair(S, near, far, gt_near, gt_far, num_point=64):
    """train_multi_gpu_pc_compare_dist.py:752-766 for one item."""
    np_, h, q = num_point, num_point // 2, num_point // 4
    S_A, S_B = S[:np_], S[np_:2 * np_]
    pcA = S_A[:np_]
    pcB = np.concatenate([S_B[:h], near[:q], far[q:2 * q]], 0)
    labels = np.concatenate([np.zeros(h, np.float32), gt_near[:q], gt_far[q:2 * q]], 0)
    return pcA, pcB, labels


def anchor_pair(seed=0, num_point=64):
    """Config A: one chair pair, N = NP = num_point -> (pcA [1,N,3], pcB [1,N,3], labels_AB [1,N])."""
    pcA, pcB, lab = assemble_pair(*chair_item(seed, num_point), num_point=num_point)
    return pcA[None], pcB[None], lab[None]


def chair_batch(seed, batch, num_point=64):
    out = [assemble_pair(*chair_item(seed * 100003 + i, num_point, dense=512), num_point=num_point)
           for i in range(batch)]
    return tuple(np.stack([o[j] for o in out]) for j in range(3))


def uniform_batch(seed, batch, num_point=64, outside_frac=0.02):
    """Config B / E: points U(-0.8,0.8)^3 shifted by U(-0.1,0.1) per cloud (provider.py:200-211 range),
    `outside_frac` of the queries pushed outside [-1,1]^3 to exercise the in-cube mask.
    labels: |distance to nearest pcA point| as a stand-in ground truth."""
    rng = np.random.default_rng(seed)

    def cloud():
        p = rng.uniform(-0.8, 0.8, size=(batch, num_point, 3))
        p += rng.uniform(-0.1, 0.1, size=(batch, 1, 3))
        out = rng.random((batch, num_point)) < outside_frac
        ax = rng.integers(0, 3, size=(batch, num_point))
        push = rng.uniform(1.0, 1.3, size=(batch, num_point)) * rng.choice([-1.0, 1.0], size=(batch, num_point))
        bi, ni = np.nonzero(out)
        p[bi, ni, ax[bi, ni]] = push[bi, ni]
        return p.astype(np.float32)

    pcA, pcB = cloud(), cloud()
    labels = rng.uniform(0.0, 0.3, size=(batch, num_point)).astype(np.float32)
    return pcA, pcB, labels


def dataset_batch(seed, bsize, num_point=64):
    """Shaped like ModelNetDataset.next_batch (modelnet_dataset.py:170-187):
    points [bsize, 3*npoints, 3], labels [bsize, 2*npoints], npoints = 2*num_point."""
    npoints = 2 * num_point
    pts = np.zeros((bsize, 3 * npoints, 3), np.float32)
    lab = np.zeros((bsize, 2 * npoints), np.float32)
    for i in range(bsize):
        S, near, far, gn, gf = chair_item(seed * 100003 + i, npoints, dense=512)
        pts[i] = np.concatenate([S[:npoints], near, far], 0)
        lab[i] = np.concatenate([gn, gf], 0)
    return pts, lab
