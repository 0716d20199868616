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
    np.savez_compressed(
        os.path.join(HERE, "anchor_A.npz"),
        pcA=pcA, pcB=pcB, labels=labels,
        fvA=aux["fvA"].numpy().astype(np.float32), fvB=aux["fvB"].numpy().astype(np.float32),
        pred_AB=p["pred_listAB"].numpy().astype(np.float32), pred_BA=p["pred_listBA"].numpy().astype(np.float32),
        idx_A=idx_A.numpy().astype(np.int32), idx_B=idx_B.numpy().astype(np.int32),
        loss=np.float32(loss), loss_pred=np.float32(loss_pred),
        weight_seed=np.int64(WEIGHT_SEED),
        weight_checksum=np.float64(sum(v.double().abs().sum() for v in var.values())))
    print("wrote anchor_A.npz; loss %.6f loss_pred %.6f" % (float(loss), float(loss_pred)))


if __name__ == "__main__":
    main()
