"""Data side of the DPDist path on the GPU: ground-truth distance generation and batch assembly.

Reference anchors:
  generate_points_with_gt   dataset_sample_with_gt.py:60-135 (near / far query sets with their distance to the surface)
  uniform_sampeling         dataset_sample_with_gt.py:141-189 ('dropped_coordinates' ball sampling, 'cube')
  assemble + augment        train_multi_gpu_pc_compare_dist.py:749-766; modelnet_dataset.py:82-95; provider.py:32-50, 200-211
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .dpdist_util import _check_cuda, _ptr, _stream


def nearest_distance(surface, query, return_index=False):
    """cdist(surface, query).min(0) per cloud (dataset_sample_with_gt.py:90-91): surface [B,S,3], query [B,Q,3]
    -> dist [B,Q] (and the index of the nearest surface point)."""
    lib = _lib.load()
    surface, query = _check_cuda(surface, "surface"), _check_cuda(query, "query")
    if surface.dim() != 3 or query.dim() != 3 or surface.shape[0] != query.shape[0] or surface.shape[2] != 3 or query.shape[2] != 3:
        raise ValueError("surface must be [B,S,3] and query [B,Q,3]")
    B, S, _ = surface.shape
    Q = query.shape[1]
    dist = torch.empty((B, Q), device=query.device, dtype=torch.float32)
    arg = torch.empty((B, Q), device=query.device, dtype=torch.int32) if return_index else None
    with torch.cuda.device(query.device):
        rc = lib.dpd_nearest_distance(_ptr(surface), B, S, _ptr(query), Q, _ptr(dist), _ptr(arg) if arg is not None else None,
                                      _stream())
    _lib.check(rc, "dpd_nearest_distance")
    return (dist, arg) if return_index else dist


def assemble_batch(batch_data, batch_label, NUM_POINT, angle=None, shift=None):
    """train_one_epoch_3d's batch (train...py:749-766) from a device-resident dataset batch, with the dataset's
    augmentation (rotation about y by `angle` [bsize], then `shift` [bsize,3]) applied in the same kernel.
    batch_data [bsize, 3*npoints, 3], batch_label [bsize, 2*npoints] -> (pcA, pcB, labels_AB)."""
    lib = _lib.load()
    data, label = _check_cuda(batch_data, "batch_data"), _check_cuda(batch_label, "batch_label")
    bsize, n3, _ = data.shape
    if n3 % 3 != 0 or label.shape != (bsize, 2 * (n3 // 3)):
        raise ValueError("batch_data must be [bsize, 3*npoints, 3] and batch_label [bsize, 2*npoints]")
    npoints = n3 // 3
    dev = data.device
    pcA = torch.empty((bsize, NUM_POINT, 3), device=dev, dtype=torch.float32)
    pcB = torch.empty((bsize, NUM_POINT, 3), device=dev, dtype=torch.float32)
    lab = torch.empty((bsize, NUM_POINT), device=dev, dtype=torch.float32)
    a = _check_cuda(angle, "angle") if angle is not None else None
    s = _check_cuda(shift, "shift") if shift is not None else None
    if a is not None and a.numel() != bsize or s is not None and s.numel() != 3 * bsize:
        raise ValueError("angle must be [bsize] and shift [bsize,3]")
    with torch.cuda.device(dev):
        rc = lib.dpd_assemble_batch(_ptr(data), _ptr(label), bsize, npoints, int(NUM_POINT), _ptr(a) if a is not None else None,
                                    _ptr(s) if s is not None else None, _ptr(pcA), _ptr(pcB), _ptr(lab), _stream())
    _lib.check(rc, "dpd_assemble_batch")
    return pcA, pcB, lab


def random_augmentation(bsize, device, generator=None, shift_range=0.1):
    """The random draws of ModelNetDataset._augment_batch_data: angle ~ U(0, 2 pi) (provider.py:42),
    shift ~ U(-0.1, 0.1)^3 (provider.py:208), one per item."""
    angle = torch.rand(bsize, device=device, generator=generator) * (2 * np.pi)
    shift = (torch.rand((bsize, 3), device=device, generator=generator) * 2 - 1) * shift_range
    return angle, shift


def uniform_sampeling(shape, type="dropped_coordinates", vmin=-1.0, vmax=1.0, device="cuda", generator=None):
    """dataset_sample_with_gt.py:141-189, the two variants the generator uses: uniform in the unit ball by dropping two
    coordinates of a point on S^4 ('dropped_coordinates'), and uniform in the cube."""
    B, N, D = shape
    if type == "cube":
        return torch.rand((B, N, D), device=device, generator=generator) * (vmax - vmin) + vmin
    if type != "dropped_coordinates":
        raise NotImplementedError("only 'dropped_coordinates' and 'cube' are used by generate_points_with_gt")
    g = torch.randn((B, N, 5), device=device, generator=generator)
    return g[..., 2:5] / g.norm(dim=-1, keepdim=True)


def generate_points_with_gt(point_set, eps=0.05, min_eps=0.001, num_neg_points=10 ** 4, generator=None, batch=50000):
    """dataset_sample_with_gt.py:79-127 for one shape: point_set [S,3] (already scaled by 0.8) ->
    (neg_set_l [num_neg,4], neg_set_u [num_neg,4]) rows = (x, y, z, distance to the surface); near set: min_eps < d < 2 eps,
    far set: d > 2 eps with its last 10 % replaced by points of the cube outside the unit ball."""
    surf = _check_cuda(point_set, "point_set")[None]
    dev = surf.device
    f = 2
    near, far, size_l, size_u = [], [], 0, 0
    while size_l < num_neg_points:
        cand = uniform_sampeling([1, batch, 3], device=dev, generator=generator)
        d = nearest_distance(surf, cand)[0]
        rows = torch.cat([cand[0], d[:, None]], -1)
        ind_l = (d > min_eps) & (d < f * eps)
        near.append(rows[ind_l])
        size_l += int(ind_l.sum())
        if size_u < num_neg_points:
            ind_u = d > f * eps
            far.append(rows[ind_u])
            size_u += int(ind_u.sum())
    neg_set_l = torch.cat(near)[:num_neg_points]
    neg_set_u = torch.cat(far)[:num_neg_points]
    n_out = int(num_neg_points * 0.1)
    outs, size_uu = [], 0
    while size_uu < n_out:
        cand = uniform_sampeling([1, batch, 3], type="cube", device=dev, generator=generator)[0]
        cand = cand[cand.square().sum(-1).sqrt() > 1]
        d = nearest_distance(surf, cand[None])[0]
        outs.append(torch.cat([cand, d[:, None]], -1))
        size_uu += cand.shape[0]
    neg_set_u[-n_out:] = torch.cat(outs)[:n_out]
    return neg_set_l, neg_set_u
