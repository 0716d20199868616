"""Host-side mirror of the reference's dataset reader `modelnet_dataset.py` (ModelNet shapes with ground-truth
distances) plus a device-resident fast path for the trainer.

File formats and item layout are the reference's (modelnet_dataset.py:30-146):
  <root>/modelnet40_shape_names.txt, modelnet40_{train,test}.txt (modelnet10_* with modelnet10=True)
  <root>/<shape>/<id>_dist_c_scaled.txt                  surface points, csv x,y,z            (:116-121)
  <root>/<shape>/<id>_10000_dist_c_neg_l.txt             near points, csv x,y,z,distance      (:124-126)
  <root>/<shape>/<id>_10000_dist_c_neg_u.txt             far points,  csv x,y,z,distance      (:127-128)
  item = [surface[:npoints] | near[:npoints] | far[perm[:npoints]]] (3*npoints x 3), labels = [gt near | gt far] (:136-139)
  every access re-shuffles the npoints of each of the three sets with ONE permutation (:99-110)
`next_batch(augment)` returns numpy arrays exactly like the reference ([bsize, 3*npoints, 3], [bsize, 2*npoints]).
`next_batch_device(num_point, augment)` is the B200 path: items live in HBM after their first read, the per-item
permutation is a device gather, and slicing + rotation + shift run in one kernel (dpd_assemble_batch) -> (pcA, pcB,
labels_AB) ready for DPDistTrainer.step.  Files are produced by `python -m dpdist_b200.dataset_sample_with_gt`.
"""
import os

import numpy as np
import torch

NUM_NEG_POINTS = 10 ** 4      # modelnet_dataset.py:123


def pc_normalize(pc):
    """modelnet_dataset.py:22-28."""
    centroid = np.mean(pc, axis=0)
    pc = pc - centroid
    m = np.max(np.sqrt(np.sum(pc ** 2, axis=1)))
    return pc / m


def rotate_point_cloud(batch_data, rng=np.random):
    """provider.py:32-50: one random rotation about the up axis per shape."""
    rotated_data = np.zeros(batch_data.shape, dtype=np.float32)
    for k in range(batch_data.shape[0]):
        rotation_angle = rng.uniform() * 2 * np.pi
        cosval, sinval = np.cos(rotation_angle), np.sin(rotation_angle)
        rotation_matrix = np.array([[cosval, 0, sinval], [0, 1, 0], [-sinval, 0, cosval]])
        rotated_data[k, ...] = np.dot(batch_data[k, ...].reshape((-1, 3)), rotation_matrix)
    return rotated_data


def shift_point_cloud(batch_data, shift_range=0.1, rng=np.random):
    """provider.py:200-211: one random shift per shape."""
    B = batch_data.shape[0]
    shifts = rng.uniform(-shift_range, shift_range, (B, 3))
    for batch_index in range(B):
        batch_data[batch_index, :, :] += shifts[batch_index, :]
    return batch_data


class ModelNetDataset:
    def __init__(self, root, batch_size=32, npoints=1024, split='train', normalize=False, normal_channel=False,
                 modelnet10=False, cache_size=15000, shuffle=None, class_choice=None, device=None, seed=None):
        if normal_channel:
            raise NotImplementedError("normal_channel=True is not used by the DPDist trainer (train...py:184)")
        self.root, self.batch_size, self.npoints = root, batch_size, npoints
        self.normalize, self.split, self.normal_channel = normalize, split, normal_channel
        prefix = 'modelnet10' if modelnet10 else 'modelnet40'
        self.catfile = os.path.join(self.root, prefix + '_shape_names.txt')
        self.cat = [line.rstrip() for line in open(self.catfile)]
        self.classes = dict(zip(self.cat, range(len(self.cat))))
        assert split == 'train' or split == 'test'
        ids = [line.rstrip() for line in open(os.path.join(self.root, '%s_%s.txt' % (prefix, split)))]
        self.datapath = []
        for x in ids:                                                   # :55-66 category selection
            name = '_'.join(x.split('_')[0:-1])
            if class_choice and name not in class_choice:
                continue
            self.datapath.append((name, os.path.join(self.root, name, x) + '.txt'))
        self.cache_size = cache_size
        self.cache = {}
        self.shuffle = (split == 'train') if shuffle is None else shuffle
        self.rng = np.random.default_rng(seed) if seed is not None else np.random.default_rng()
        self.shuffle_points_ind = np.arange(self.npoints)
        self.device = torch.device(device) if device is not None else None
        self._dev_items = {}                                            # index -> (points [3*npoints,3], labels [2*npoints]) in HBM
        self.reset()

    # ---- items -------------------------------------------------------------------------------------------
    def _load(self, index):
        """:111-142 (first access of an item)."""
        npoints = self.npoints
        fn = self.datapath[index]
        base = fn[1][:-4]
        point_set = np.loadtxt(base + '_dist_c_scaled.txt', delimiter=',').astype(np.float32)[0:npoints, :]
        neg_l = np.loadtxt(base + '_%d_dist_c_neg_l.txt' % NUM_NEG_POINTS, delimiter=',').astype(np.float32)
        neg_u = np.loadtxt(base + '_%d_dist_c_neg_u.txt' % NUM_NEG_POINTS, delimiter=',').astype(np.float32)
        shuff_ind_u = np.arange(len(neg_u))
        self.rng.shuffle(shuff_ind_u)        # the last 10 % of the far set lie outside the unit ball: mix them in (:130-133)
        pts = np.concatenate([point_set[:npoints, :3], neg_l[:npoints, :3], neg_u[shuff_ind_u[:npoints], :3]], 0)
        labels = np.concatenate([neg_l[:npoints, 3], neg_u[shuff_ind_u[:npoints], 3]], 0)
        if self.normalize:
            pts[:, 0:3] = pc_normalize(pts[:, 0:3])
        cls = np.array([self.classes[fn[0]]]).astype(np.int32)
        return pts.astype(np.float32), cls, labels.astype(np.float32)

    def _get_item(self, index):
        """:98-146: cached items are re-shuffled with one permutation for the three point sets and the two label sets."""
        shuff_ind = self.shuffle_points_ind
        self.rng.shuffle(shuff_ind)
        npoints = self.npoints
        if index in self.cache:
            point_set, cls, labels = self.cache[index]
            point_set = point_set.reshape(3, npoints, 3)[:, shuff_ind].reshape(3 * npoints, 3)
            labels = labels.reshape(2, npoints)[:, shuff_ind].reshape(2 * npoints)
            return point_set, cls, labels
        point_set, cls, labels = self._load(index)
        if len(self.cache) < self.cache_size:
            self.cache[index] = (point_set, cls, labels)
        return point_set, cls, labels

    def __getitem__(self, index):
        return self._get_item(index)

    def __len__(self):
        return len(self.datapath)

    def num_channel(self):
        return 3

    # ---- batches -----------------------------------------------------------------------------------------
    def reset(self):
        self.idxs = np.arange(0, len(self.datapath))
        if self.shuffle:
            self.rng.shuffle(self.idxs)
        self.num_batches = (len(self.datapath) + self.batch_size - 1) // self.batch_size
        self.batch_idx = 0

    def has_next_batch(self):
        return self.batch_idx < self.num_batches

    def _augment_batch_data(self, batch_data):
        """:82-95 for normal_channel False: rotation about the up axis, then a shift."""
        rotated = rotate_point_cloud(batch_data, self.rng)
        rotated[:, :, 0:3] = shift_point_cloud(rotated[:, :, 0:3], rng=self.rng)
        return rotated

    def next_batch(self, augment=False):
        """:170-187.  The returned batch may be smaller than batch_size."""
        start_idx = self.batch_idx * self.batch_size
        end_idx = min((self.batch_idx + 1) * self.batch_size, len(self.datapath))
        bsize = end_idx - start_idx
        batch_data = np.zeros((bsize, self.npoints * 3, self.num_channel()))
        batch_label = np.zeros((bsize, self.npoints * 2), dtype=np.float32)
        for i in range(bsize):
            ps, _, labels = self._get_item(self.idxs[i + start_idx])
            batch_data[i] = ps
            batch_label[i] = labels
        self.batch_idx += 1
        if augment:
            batch_data = self._augment_batch_data(batch_data)
        return batch_data, batch_label

    def next_batch_device(self, num_point, augment=False):
        """The same batch, assembled for the DPDist trainer on the GPU: -> (pcA, pcB, labels_AB) [bsize, num_point, ...].
        Items are uploaded once; the per-access point permutation, the surface / close / far slicing
        (train_multi_gpu_pc_compare_dist.py:749-766) and the augmentation run on the device."""
        from . import data as D
        if self.device is None:
            raise ValueError("construct the dataset with device=... to use next_batch_device")
        start_idx = self.batch_idx * self.batch_size
        end_idx = min((self.batch_idx + 1) * self.batch_size, len(self.datapath))
        npoints = self.npoints
        pts, labs = [], []
        for i in range(start_idx, end_idx):
            index = int(self.idxs[i])
            if index not in self._dev_items:
                p, _, l = self._load(index)
                self._dev_items[index] = (torch.from_numpy(p).to(self.device), torch.from_numpy(l).to(self.device))
            pts.append(self._dev_items[index][0])
            labs.append(self._dev_items[index][1])
        self.batch_idx += 1
        bsize = len(pts)
        data = torch.stack(pts)                                  # [bsize, 3*npoints, 3]
        label = torch.stack(labs)                                # [bsize, 2*npoints]
        # one permutation per item for its three point sets and two label sets (:99-110)
        perm = torch.stack([torch.from_numpy(self.rng.permutation(npoints)) for _ in range(bsize)]).to(self.device)
        data = torch.gather(data.view(bsize, 3, npoints, 3), 2, perm[:, None, :, None].expand(bsize, 3, npoints, 3)).reshape(bsize, 3 * npoints, 3)
        label = torch.gather(label.view(bsize, 2, npoints), 2, perm[:, None, :].expand(bsize, 2, npoints)).reshape(bsize, 2 * npoints)
        angle = shift = None
        if augment:
            angle = torch.from_numpy((self.rng.uniform(size=bsize) * 2 * np.pi).astype(np.float32)).to(self.device)
            shift = torch.from_numpy(self.rng.uniform(-0.1, 0.1, (bsize, 3)).astype(np.float32)).to(self.device)
        return D.assemble_batch(data.contiguous(), label.contiguous(), num_point, angle=angle, shift=shift)
