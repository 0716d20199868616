"""PCRNet registration trained with the DPDist loss (BASELINE config D): the consumer either side of the hot path.

Mirrors the reference driver pcrnet-registration/iterative_PCRNet_ours.py:
  network       models/ipcr_model.py:198-231 (`pointnet`: shared MLP 64-64-64-128-1024 + max pool over the points),
                :268-279 (`get_pose`: fc 1024-512-256, dropout keep 0.7, fc 7 = translation | quaternion)
  transform     helper.py:539-570 (`transformation_quat_tensor`), quaternion normalised as :213-220 of the driver
  loss          the serialized DPDist graph with input1 = transformed source, input2 = template,
                loss = (mean(output1[...,0]) + mean(output2[...,0])) / 2                         :229-254
  iteration     MAX_LOOPS-1 pose refinements without training, then one trained step             :407-471
  poses         generate_poses_ours.py:15-18 ranges, helper.py:229-262 (`apply_transformation`, R = Rx Ry Rz)

What is ours here is the loss: `DPDistLoss` runs the 3DmFV / patch-gather MLP kernels forward and their hand-written
backward (dpd_head_backward_inputs, dpd_fv_backward) into `input1`.  The small pose network itself is host-side
plumbing expressed with torch.nn (cuBLAS-sized layers: 2 x 16 x 64 points), as SURVEY.md 8 f3 scopes it.
"""
import argparse
import math
import time

import numpy as np
import torch

from . import synthetic
from .dpdist_loss import DPDistLoss


def transformation_quat_tensor(data, quat, translation):
    """helper.py:539-570, batched: data [B,N,3], quat [B,4] = (q0,q1,q2,q3), translation [B,3]."""
    q0, q1, q2, q3 = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    R = torch.stack([
        torch.stack([q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)], -1),
        torch.stack([2 * (q1 * q2 + q0 * q3), q0 * q0 + q2 * q2 - q1 * q1 - q3 * q3, 2 * (q2 * q3 - q0 * q1)], -1),
        torch.stack([2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), q0 * q0 + q3 * q3 - q1 * q1 - q2 * q2], -1)], 1)
    return torch.einsum("bij,bnj->bni", R, data) + translation[:, None, :]


def normalize_quat(predicted_quat):
    """iterative_PCRNet_ours.py:213-220: q / (|q| + 1e-7)."""
    return predicted_quat / (predicted_quat.square().sum(1, keepdim=True).sqrt() + 0.0000001)


def euler_to_matrix(poses):
    """helper.py:229-262: R = Rx @ Ry @ Rz from poses[:, 3:6] = (rx, ry, rz); -> [B,3,3] float64 numpy."""
    out = np.zeros((poses.shape[0], 3, 3))
    for i, (rx, ry, rz) in enumerate(poses[:, 3:6]):
        Rx = np.array([[1, 0, 0], [0, np.cos(rx), -np.sin(rx)], [0, np.sin(rx), np.cos(rx)]])
        Ry = np.array([[np.cos(ry), 0, np.sin(ry)], [0, 1, 0], [-np.sin(ry), 0, np.cos(ry)]])
        Rz = np.array([[np.cos(rz), -np.sin(rz), 0], [np.sin(rz), np.cos(rz), 0], [0, 0, 1]])
        out[i] = Rx @ Ry @ Rz
    return out


def apply_transformation(datas, poses):
    """helper.py:229-262: rotate by R(poses[:,3:6]) then translate by poses[:,0:3]."""
    R = euler_to_matrix(poses)
    return (np.einsum("bij,bnj->bni", R, datas) + poses[:, None, 0:3]).astype(np.float32)


def generate_poses(n, rng, max_t=0.01, max_deg=45.0):
    """generate_poses_ours.py:15-18: translation U(+-max_t), Euler angles U(+-max_deg)."""
    t = rng.uniform(-max_t, max_t, size=(n, 3))
    r = rng.uniform(-max_deg, max_deg, size=(n, 3)) * (np.pi / 180)
    return np.concatenate([t, r], 1)


def chamfer_dist(pc, rec_pc):
    """iterative_PCRNet_ours.py:166-187 (squared distances, mean of the two directions)."""
    d = torch.cdist(rec_pc, pc).square()
    return (d.min(2).values.mean() + d.min(1).values.mean()) / 2.0


class PCRNet(torch.nn.Module):
    """models/ipcr_model.py `pointnet` + `get_pose` (bn off, max pooling, quaternion head)."""

    def __init__(self, out_features=1024):
        super().__init__()
        dims = [3, 64, 64, 64, 128, out_features]
        self.convs = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])   # 1x3 / 1x1 convs
        self.fc1 = torch.nn.Linear(2 * out_features, 1024)
        self.fc2 = torch.nn.Linear(1024, 512)
        self.fc3 = torch.nn.Linear(512, 256)
        self.dp4 = torch.nn.Dropout(p=0.3)                                    # keep_prob 0.7
        self.fc4 = torch.nn.Linear(256, 7)
        for m in self.modules():                                              # tf_util: xavier weights, zero biases
            if isinstance(m, torch.nn.Linear):
                torch.nn.init.xavier_uniform_(m.weight)
                torch.nn.init.zeros_(m.bias)

    def features(self, pc):
        x = pc
        for c in self.convs:
            x = torch.relu(c(x))
        return x.max(dim=1).values

    def forward(self, source, template):
        f = torch.cat([self.features(source), self.features(template)], 1)
        x = torch.relu(self.fc1(f))
        x = torch.relu(self.fc2(x))
        x = torch.relu(self.fc3(x))
        return self.fc4(self.dp4(x))                                          # [B,7] = translation(3) | quaternion(4)


def compose(TRANSFORMATIONS, pose7):
    """helper.py:309-329 on the device: T <- [R(q) t; 0 1] @ T.  The reference passes the raw quaternion to
    transforms3d.quat2mat, which normalises it."""
    q = pose7[:, 3:7] / pose7[:, 3:7].norm(dim=1, keepdim=True).clamp_min(1e-12)
    eye = torch.eye(3, device=pose7.device).expand(pose7.shape[0], 3, 3)
    R = transformation_quat_tensor(eye, q, torch.zeros_like(pose7[:, :3])).transpose(1, 2)
    M = torch.zeros((pose7.shape[0], 4, 4), device=pose7.device)
    M[:, :3, :3], M[:, :3, 3], M[:, 3, 3] = R, pose7[:, :3], 1.0
    return M @ TRANSFORMATIONS, R


class IterativePCRNetOurs:
    """One trainer: `train_step(source, template)` = iterative_PCRNet_ours.py:407-471 for one batch."""

    def __init__(self, dpdist, max_loops=8, learning_rate=0.001, train_single=False, device=None, seed=0, cuda_graph=False):
        self.dpdist = dpdist
        self.device = torch.device(device) if device is not None else next(dpdist.parameters()).device
        torch.manual_seed(seed)
        self.net = PCRNet().to(self.device)
        # cuda_graph=True: the whole batch step (max_loops pose refinements, the DPDist loss forward and backward, Adam) is
        # captured once per batch shape and replayed -- at batch 16 the step is otherwise bound by ~150 tiny launches
        self.cuda_graph = bool(cuda_graph)
        self.opt = torch.optim.Adam(self.net.parameters(), lr=learning_rate, eps=1e-8,       # TF AdamOptimizer defaults
                                    capturable=self.cuda_graph)
        self.max_loops, self.train_single = int(max_loops), bool(train_single)
        self._side = torch.cuda.Stream(device=self.device) if self.cuda_graph else None
        self._graph, self._eager_steps = None, 0

    def _predict(self, source, template):
        pose = self.net(source, template)
        quat = normalize_quat(pose[:, 3:7])
        return pose, transformation_quat_tensor(source, quat, pose[:, 0:3])

    def _trained_step(self, source, template):
        pose, moved = self._predict(source, template)
        loss = self.dpdist.loss(moved, template)                 # gradients reach the pose network through input1 only
        params = [p for p in self.net.parameters()]
        # autograd.grad (no gradient accumulators): identical in eager mode, and safe inside a stream capture
        for p, g in zip(params, torch.autograd.grad(loss, params)):
            p.grad = g
        self.opt.step()
        return pose.detach(), loss.detach()

    def train_step(self, source, template):
        if not self.cuda_graph:
            return self._train_step(source, template)
        if self._eager_steps < 3:                                # warm-up on the stream the capture will use
            self._eager_steps += 1
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                out = self._train_step(source, template)
            cur.wait_stream(self._side)
            return out
        gs = self._graph
        if gs is None or gs["shape"] != tuple(source.shape):
            gs = {"shape": tuple(source.shape), "src": source.clone(), "tpl": template.clone()}
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=self._side):
                gs["out"] = self._train_step(gs["src"], gs["tpl"])
            gs["graph"] = graph
            self._graph = gs
        gs["src"].copy_(source, non_blocking=True)
        gs["tpl"].copy_(template, non_blocking=True)
        gs["graph"].replay()
        return tuple(t.clone() for t in gs["out"])     # the graph's output buffers are overwritten by the next replay

    def _train_step(self, source, template):
        B = source.shape[0]
        T = torch.eye(4, device=self.device).repeat(B, 1, 1)
        self.net.train()
        loss = None
        for _ in range(self.max_loops - 1):
            if self.train_single:
                pose, loss = self._trained_step(source, template)
            else:
                with torch.no_grad():
                    pose = self.net(source, template)
            T, R = compose(T, pose)
            source = torch.einsum("bij,bnj->bni", R, source) + pose[:, None, 0:3]        # helper.py:327-328
        pose, loss = self._trained_step(source, template)
        T, R = compose(T, pose)
        source = torch.einsum("bij,bnj->bni", R, source) + pose[:, None, 0:3]
        return loss, T, source

    @torch.no_grad()
    def register(self, source, template):
        """Evaluation loop (results_itrPCRNet_no_stop.eval_network): max_loops refinements, no training."""
        self.net.eval()
        T = torch.eye(4, device=self.device).repeat(source.shape[0], 1, 1)
        for _ in range(self.max_loops):
            pose = self.net(source, template)
            T, R = compose(T, pose)
            source = torch.einsum("bij,bnj->bni", R, source) + pose[:, None, 0:3]
        return T, source


def synthetic_templates(n, num_point, seed):
    """Chair-like templates shaped like the reference's templates array [n, num_point, 3] (data_txt_to_hdf5.py:39-51)."""
    out = np.zeros((n, num_point, 3), np.float32)
    for i in range(n):
        out[i] = synthetic.chair_item(seed * 7919 + i, num_point, dense=64)[0][:num_point]
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description="PCRNet with the DPDist loss on synthetic chair templates (config D)")
    ap.add_argument("--batch_size", type=int, default=16)
    ap.add_argument("--num_point", type=int, default=64)
    ap.add_argument("--max_loops", type=int, default=8)
    ap.add_argument("--learning_rate", type=float, default=0.001)
    ap.add_argument("--max_epoch", type=int, default=2)
    ap.add_argument("--num_templates", type=int, default=64)
    ap.add_argument("--train_single", type=int, default=0)
    ap.add_argument("--cuda_graph", type=int, default=0)
    ap.add_argument("--model_path", default="", help="DPDist checkpoint (TF V2 prefix or .npz); random init if empty")
    args = ap.parse_args(argv)
    dev = torch.device("cuda", 0)
    dpd = DPDistLoss(num_point=args.num_point, device=dev, seed=1)
    if args.model_path:
        dpd.restore(args.model_path)
    tr = IterativePCRNetOurs(dpd, args.max_loops, args.learning_rate, bool(args.train_single), dev, cuda_graph=bool(args.cuda_graph))
    rng = np.random.default_rng(0)
    templates = synthetic_templates(args.num_templates, args.num_point, seed=3)
    for epoch in range(args.max_epoch):
        t0, losses = time.time(), []
        for s in range(0, args.num_templates - args.batch_size + 1, args.batch_size):
            tpl = templates[s:s + args.batch_size]
            src = apply_transformation(tpl, generate_poses(args.batch_size, rng))
            loss, _, moved = tr.train_step(torch.tensor(src, device=dev), torch.tensor(tpl, device=dev))
            losses.append(loss)
        torch.cuda.synchronize()
        print("epoch %03d  mean DPDist loss %.6f  (%.1f ms / batch)" % (
            epoch, float(torch.stack(losses).mean()), (time.time() - t0) * 1e3 / max(1, len(losses))))


if __name__ == "__main__":
    main()
