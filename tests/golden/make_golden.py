"""Writes tests/golden/anchor_A.npz: BASELINE.json config A (one chair pair, N=NP=64, G=8, k=5,
128 queries) through the fp64 twin of oracle/dpdist_oracle.py, stored as fp32.

    python tests/golden/make_golden.py

The reference cannot run here (TF1) and ships no vectors, so these are ORACLE
outputs: they pin the restatement and give the GPU box a committed anchor; they are not TF1 output.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import dpdist_oracle as O  # noqa: E402
from dpdist_b200 import synthetic as S  # noqa: E402

WEIGHT_SEED = 1


def main():
    pcA, pcB, labels = S.anchor_pair(seed=0, num_point=64)
    var = O.unit_scale_variables(WEIGHT_SEED)
    var64 = {k: v.double() for k, v in var.items()}
    p, aux, _ = O.get_model(torch.tensor(pcA).double(), torch.tensor(pcB).double(), var64)
    # indices from the fp32 path (they are defined by fp32 comparisons)
    _, _, idx_B = O.get_pc_grid_binary_mask_from_centers(aux["C"].float(), torch.tensor(pcB))
    _, _, idx_A = O.get_pc_grid_binary_mask_from_centers(aux["C"].float(), torch.tensor(pcA))
    loss, loss_pred = O.get_loss(p, {}, torch.tensor(labels).double())
    # gradients of the consumers' loss (mean out1[...,0] + mean out2[...,0]) / 2 into both clouds and of
    # the training loss into the variables, through the same fp64 twin (amax / amin: TF tie-splitting)
    a = torch.tensor(pcA).double().requires_grad_(True)
    b = torch.tensor(pcB).double().requires_grad_(True)
    v = {k: t.clone().requires_grad_(True) for k, t in var64.items()}
    p2, _, _ = O.get_model(a, b, v)
    ((p2["pred_listAB"][..., 0].mean() + p2["pred_listBA"][..., 0].mean()) / 2).backward(retain_graph=True)
    g_in1, g_in2 = a.grad.clone(), b.grad.clone()
    for t in v.values():
        t.grad = None
    O.get_loss(p2, {}, torch.tensor(labels).double())[0].backward()
    grad_sums = np.array([float(v[k].grad.abs().sum()) for k in sorted(v)], dtype=np.float64)
    g_w4 = v[O.VAR_PREFIX + "mapper_conv4/weights"].grad.numpy().astype(np.float32)
    g_b1 = v[O.VAR_PREFIX + "mapper_conv1/biases"].grad.numpy().astype(np.float32)
    np.savez_compressed(
        os.path.join(HERE, "anchor_A.npz"),
        pcA=pcA, pcB=pcB, labels=labels,
        fvA=aux["fvA"].numpy().astype(np.float32), fvB=aux["fvB"].numpy().astype(np.float32),
        pred_AB=p["pred_listAB"].numpy().astype(np.float32), pred_BA=p["pred_listBA"].numpy().astype(np.float32),
        idx_A=idx_A.numpy().astype(np.int32), idx_B=idx_B.numpy().astype(np.int32),
        loss=np.float32(loss), loss_pred=np.float32(loss_pred),
        grad_input1=g_in1.numpy().astype(np.float32), grad_input2=g_in2.numpy().astype(np.float32),
        grad_abs_sums=grad_sums, grad_w4=g_w4, grad_b1=g_b1,
        weight_seed=np.int64(WEIGHT_SEED),
        weight_checksum=np.float64(sum(v.double().abs().sum() for v in var.values())))
    print("wrote anchor_A.npz; loss %.6f loss_pred %.6f" % (float(loss), float(loss_pred)))


if __name__ == "__main__":
    main()
