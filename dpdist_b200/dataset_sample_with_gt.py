"""Ground-truth generator: for every shape, the scaled surface cloud plus near / far query points with their distance
to it -- the three text files `modelnet_dataset.ModelNetDataset` reads.  Mirrors the reference script
dataset_sample_with_gt.py:60-135 with the scipy `cdist(...).min(0)` replaced by `dpd_nearest_distance` on the GPU.

    python -m dpdist_b200.dataset_sample_with_gt --data_root data/modelnet40_normal_resampled [--only_chair] [--classes chair,table]

Per shape `<root>/<shape>/<id>.txt` (csv, >= 3 columns: x,y,z[,normals]) it writes
    <id>_dist_c_scaled.txt               the cloud scaled by 0.8                                   (:82, :124)
    <id>_10000_dist_c_neg_l.txt          10^4 rows x,y,z,d with 0.001 < d < 0.1 (near the surface)  (:92-97)
    <id>_10000_dist_c_neg_u.txt          10^4 rows x,y,z,d with d > 0.1, the last 10 % outside the unit ball  (:99-121)
Note: the reference assigns both output names to one variable (`fn_neg`, :73-74), so as shipped it writes the near set
and then the far set to the SAME `_neg_u` file and never creates `_neg_l`, which its own loader requires
(modelnet_dataset.py:124-125).  This generator writes the two files the loader reads.
"""
import argparse
import os
import time

import numpy as np
import torch

from . import data as D


def get_data_files(data_root, split='train'):
    """dataset_sample_with_gt.py:191-203."""
    shape_ids = [line.rstrip() for line in open(os.path.join(data_root, 'modelnet40_%s.txt' % split))]
    names = ['_'.join(x.split('_')[0:-1]) for x in shape_ids]
    return [(names[i], os.path.join(data_root, names[i], shape_ids[i]) + '.txt') for i in range(len(shape_ids))]


def output_names(path, num_neg_points=10 ** 4):
    base = path[:-4]
    return (base + '_dist_c_scaled.txt', base + '_%d_dist_c_neg_l.txt' % num_neg_points,
            base + '_%d_dist_c_neg_u.txt' % num_neg_points)


def generate_points_with_gt(data_root, eps=0.05, min_eps=0.001, num_neg_points=10 ** 4, cur_cls=(), device="cuda:0",
                            seed=None, splits=('test', 'train'), verbose=True):
    gen = torch.Generator(device=device)
    if seed is not None:
        gen.manual_seed(seed)
    done = 0
    for split in splits:
        for shape, path in get_data_files(data_root, split):
            if cur_cls and shape not in cur_cls:
                continue
            fn_pos, fn_l, fn_u = output_names(path, num_neg_points)
            if all(os.path.exists(f) for f in (fn_pos, fn_l, fn_u)):
                if verbose:
                    print('data already exist for: {}'.format(path))
                continue
            t0 = time.time()
            point_set = np.loadtxt(path, delimiter=',').astype(np.float32)[:, 0:3] * np.float32(0.8)     # :79-82
            neg_l, neg_u = D.generate_points_with_gt(torch.from_numpy(point_set).to(device), eps=eps, min_eps=min_eps,
                                                     num_neg_points=num_neg_points, generator=gen)
            np.savetxt(fn_pos, point_set, fmt='%.6f', delimiter=',')                                     # :124-126
            np.savetxt(fn_l, neg_l.cpu().numpy(), fmt='%.6f', delimiter=',')
            np.savetxt(fn_u, neg_u.cpu().numpy(), fmt='%.6f', delimiter=',')
            done += 1
            if verbose:
                print('processing time: {:.3f} s  {}'.format(time.time() - t0, path))
    return done


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--data_root', default='data/modelnet40_normal_resampled')
    ap.add_argument('--only_chair', action='store_true')
    ap.add_argument('--classes', default='', help='comma-separated shape names (default: all)')
    ap.add_argument('--num_neg_points', type=int, default=10 ** 4)
    ap.add_argument('--seed', type=int, default=None)
    args = ap.parse_args(argv)
    cls = ['chair'] if args.only_chair else [c for c in args.classes.split(',') if c]
    n = generate_points_with_gt(args.data_root, num_neg_points=args.num_neg_points, cur_cls=cls, seed=args.seed)
    print('generated ground truth for %d shapes' % n)


if __name__ == '__main__':
    main()
