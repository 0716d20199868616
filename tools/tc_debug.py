"""GPU bring-up script for the tcgen05 path (not a pytest): run under `timeout`."""
import ctypes
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dpdist_b200 import _lib

lib = _lib.load()
dev = "cuda:0"


def gemm_case(M, K, N, seed=0, f16=1):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(K, N, generator=g) / np.sqrt(K)
    b = torch.randn(N, generator=g) * 0.1
    ref = torch.relu(a.double() @ w.double() + b.double())
    ad, wd, bd = a.to(dev), w.to(dev), b.to(dev)
    out = torch.full((M, N), float("nan"), device=dev)
    scratch = torch.empty(8 * (M * K + N * K) + 1024, dtype=torch.uint8, device=dev)
    rc = lib.dpd_debug_tc_gemm(ad.data_ptr(), M, K, wd.data_ptr(), N, bd.data_ptr(), out.data_ptr(), scratch.data_ptr(),
                               scratch.numel(), f16, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "dpd_debug_tc_gemm")
    torch.cuda.synchronize()
    o = out.cpu().double()
    err = (o - ref).abs()
    fp32 = torch.relu(ad @ wd + bd).cpu().double()
    print(("f16 " if f16 else "tf32") + " gemm M=%d K=%d N=%d: max|err| %.3e (torch fp32 matmul: %.3e) nan=%d ref_max %.3f" % (
        M, K, N, float(err.max()), float((fp32 - ref).abs().max()), int(torch.isnan(o).sum()), float(ref.max())), flush=True)
    if float(err.max()) > 1e-4:
        bad = (err > 1e-4).nonzero()
        print("  first bad entries:", bad[:8].tolist(), "rows bad:", sorted(set(bad[:, 0].tolist()))[:16],
              "cols bad:", sorted(set(bad[:, 1].tolist()))[:16], flush=True)
        print("  got", o[bad[0, 0], bad[0, 1]].item(), "want", ref[bad[0, 0], bad[0, 1]].item())
    return float(err.max())


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
    if which == "gemm":
        for f16 in (0, 1):
            gemm_case(128, 64, 256, f16=f16)
            gemm_case(256, 128, 512, f16=f16)
            gemm_case(300, 1024, 1024, f16=f16)
            gemm_case(128 * 150, 1024, 1024, f16=f16)
    elif which == "model":
        from dpdist_b200 import dpdist_and_aue as MODEL, dpdist_util, synthetic, tf_util
        from oracle import dpdist_oracle as O
        pcA, pcB, _ = synthetic.uniform_batch(7, 4, 64, outside_frac=0.05)
        var = O.unit_scale_variables(7)
        outs = {}
        for impl in (_lib.HEAD_SIMT, _lib.HEAD_TC, _lib.HEAD_TC_TF32):
            dpdist_util.HEAD_IMPL = impl
            store = tf_util.VariableStore(device=dev)
            store.load_state_dict(var, strict=False)
            with tf_util.use_store(store):
                p, _, _ = MODEL.get_model(torch.tensor(pcA, device=dev), torch.tensor(pcB, device=dev), False, bn=0,
                                          Embedding_Size=512, k=5, sigma3dmfv=0.125, reuse=True)
            torch.cuda.synchronize()
            outs[impl] = torch.cat([p["pred_listAB"], p["pred_listBA"]]).cpu().double()
        with O.tf_cpu_numerics():
            po, _, _ = O.get_model(torch.tensor(pcA).double(), torch.tensor(pcB).double(), {k: v.double() for k, v in var.items()})
        ref = torch.cat([po["pred_listAB"], po["pred_listBA"]])
        for impl, o in outs.items():
            print("impl %d vs fp64 oracle: max|err| %.3e" % (impl, float((o - ref).abs().max())), flush=True)
        print("tc f16 vs tc tf32: max|diff| %.3e (0 would mean the same code path)" % float((outs[2] - outs[3]).abs().max()), flush=True)
        print("tc f16 vs simt: max|diff| %.3e; tc tf32 vs simt: %.3e" % (float((outs[1] - outs[2]).abs().max()), float((outs[1] - outs[3]).abs().max())), flush=True)
