#!/usr/bin/env python
"""bench.py -- DPDist hot path on B200: patch-query distance evals/sec at N=64, K(k)=5, G=8.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (3DmFV of both clouds -> voxel assignment -> patch-gather
MLP -> masked distances) over one batch of BASELINE.json configs[1]: 1024 synthetic cloud pairs,
N = NP = 64 points, 512 Gaussians (G=8), k=5, MLP 1024x3 -> 131072 distance evals per GPU per step.
Pairs are independent, so ranks shard the batch (weak scaling: 1024 pairs per GPU), no collective.

One JSON line on rank 0.  `value` = evals/s with inputs resident in HBM; `e2e` = the same through
the reference-shaped API (dpdist_and_aue.get_model) with HOST buffers, H2D and D2H inside the
timed region; `roofline` = the dominant kernel against the measured peak; `fv_kernel` = the
3DmFV kernel against the measured HBM copy bandwidth; `cpu_baseline` = the CPU oracle (literal
restatement of the reference) on this box's host cores.
--impl reference times that CPU restatement (TensorFlow 1.x cannot run here) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "patch-query distance evals/sec at N=64,K=5"
UNIT = "evals/s"
CFG = dict(pairs_per_gpu=1024, N=64, NP=64, G=8, k=5, H=1024, sigma=0.125, C=20)
FLOPS_PER_EVAL = 2 * ((3 + CFG["C"] * CFG["k"] ** 3) * CFG["H"] + 2 * CFG["H"] ** 2 + 3 * CFG["H"])   # 9,326,592
FV_BYTES_PER_CLOUD = 4 * (3 * CFG["N"] + CFG["C"] * CFG["G"] ** 3)                                        # 41,728


def shared_config(world):
    """`config` of both arms (ours and --impl reference): the same workload, described the same way."""
    return {"workload": "configs[1]: batch=%d synthetic pairs per GPU, N=NP=64, G=8 (512 Gaussians), k=5, MLP 1024x3, forward-only" % CFG["pairs_per_gpu"],
            "evals_per_step": CFG["pairs_per_gpu"] * 2 * CFG["NP"] * world,
            "l2": "per-step working set (activations ~1 GB on the GPU, the 10.5 GB patch tensor on the CPU) exceeds the 126 MB L2; inputs rotate over 4 batches",
            "weights": "Xavier-uniform random init (TF fan rules), zero biases, seed 1"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic(kernel):
    """(bytes per launch, source) of `kernel` from profiles/ncu_traffic.json: {kernel: {"bytes": dram read + write of one
    launch, "source": "<ncu report / command / commit>"}}, written by tools/ncu_traffic.py from an `ncu --set full`
    capture of `python bench.py`; (None, reason) if the kernel is not in it."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as fh:
            d = json.load(fh)
        e = d[kernel]
        return float(e["bytes"]), e.get("source", "profiles/ncu_traffic.json")
    except (OSError, KeyError, ValueError):
        return None, "no ncu capture of this kernel under profiles/ (ncu_traffic.json)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs(rank, n_sets):
    from dpdist_b200 import synthetic
    sets = []
    for s in range(n_sets):
        pcA, pcB, _ = synthetic.uniform_batch(seed=2 + 1000 * rank + s, batch=CFG["pairs_per_gpu"], num_point=CFG["N"])
        sets.append((pcA, pcB))
    return sets


# ------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle (literal restatement of the reference's TF1 graph)
# ------------------------------------------------------------------------------------------
def cpu_reference_rate(pairs, chunk, repeats=1, variables=None):
    from oracle import dpdist_oracle as O       # bench.py's cpu legs are the one place outside tests/ that may
    from dpdist_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    var = variables if variables is not None else O.init_variables(seed=1)
    pcA, pcB, _ = synthetic.uniform_batch(seed=2, batch=pairs, num_point=CFG["N"])
    a, b = torch.tensor(pcA), torch.tensor(pcB)
    best = None
    with O.tf_cpu_numerics():
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.forward_chunked(a, b, var, chunk=chunk, Embedding_Size=CFG["G"] ** 3, k=CFG["k"], sigma3dmfv=CFG["sigma"])
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return pairs * 2 * CFG["NP"] / best, best


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the host CPU (TF1 itself cannot be installed:
    Python 3.12, no network).  Rank 0 only."""
    if rank != 0:
        return
    from oracle import dpdist_oracle as O
    # one step = the whole configs[1] batch (1024 pairs = 131072 evals), evaluated in chunks of 8 pairs because the
    # literal patch tensor of the reference is 5.12 MB per cloud (10.5 GB for the batch).  About 2 s per step on 16 cores.
    pairs = CFG["pairs_per_gpu"]
    if os.environ.get("DPD_BENCH_REF_PAIRS"):
        pairs = int(os.environ["DPD_BENCH_REF_PAIRS"])
    var = O.init_variables(seed=1)
    for _ in range(args.warmup):
        cpu_reference_rate(pairs, 8, variables=var)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_rate(pairs, 8, variables=var)
    dt = time.perf_counter() - t0
    evals = pairs * 2 * CFG["NP"] * args.steps
    v = evals / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": shared_config(1) if pairs == CFG["pairs_per_gpu"] else dict(shared_config(1), evals_per_step=pairs * 2 * CFG["NP"],
                                                                             workload="configs[1] sample: %d pairs per step" % pairs),
        "notes": "CPU restatement of the reference TF1 graph (oracle/dpdist_oracle.py, torch CPU, literal tiles and patch tensor, "
                 "8 pairs per chunk); TF1 itself is not installable here (DESIGN.md 6)",
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d steps x %d pairs (%d evals)" % (args.steps, pairs, evals)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    from dpdist_b200 import _lib, dpdist_and_aue as MODEL, dpdist_util, tf_util
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    dist = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        # a rank that fails must not leave the others waiting for NCCL's default ten minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # random-init weights of the reference architecture (Xavier, zero biases), same on every rank
    store = tf_util.VariableStore(device=dev, seed=1)
    n_sets = 4
    host_sets = make_inputs(rank, n_sets)
    pinned = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in host_sets]
    dev_sets = [(a.to(dev), b.to(dev)) for a, b in pinned]
    kw = dict(bn=0, Embedding_Size=CFG["G"] ** 3, k=CFG["k"], sigma3dmfv=CFG["sigma"], localSNmlp=[CFG["H"]] * 3)

    def step_resident(i):
        a, b = dev_sets[i % n_sets]
        with tf_util.use_store(store):
            pred, _, _ = MODEL.get_model(a, b, False, **kw)
        return pred

    out_host = [torch.empty((CFG["pairs_per_gpu"], CFG["NP"], 1, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    in_dev = [torch.empty((CFG["pairs_per_gpu"], CFG["N"], 3), dtype=torch.float32, device=dev) for _ in range(2)]

    def step_e2e(i):
        a, b = pinned[i % n_sets]
        in_dev[0].copy_(a, non_blocking=True)
        in_dev[1].copy_(b, non_blocking=True)
        with tf_util.use_store(store):
            pred, _, _ = MODEL.get_model(in_dev[0], in_dev[1], False, **kw)
        out_host[0].copy_(pred["pred_listAB"], non_blocking=True)
        out_host[1].copy_(pred["pred_listBA"], non_blocking=True)
        torch.cuda.current_stream().synchronize()     # the caller reads the distances on the host every step
        return out_host

    # The same end-to-end step as a steady stream of batches: the H2D copy of batch i+1 and the D2H read of batch i-1 run on
    # their own streams (separate copy engines) while batch i computes; the host still receives every batch's distances
    # (it waits for batch i-1's D2H event inside step i).  This is how a caller that feeds batches continuously uses the API.
    h2d_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    pin_dev = [[torch.empty((CFG["pairs_per_gpu"], CFG["N"], 3), dtype=torch.float32, device=dev) for _ in range(2)] for _ in range(2)]
    pout_host = [[torch.empty((CFG["pairs_per_gpu"], CFG["NP"], 1, 3), dtype=torch.float32).pin_memory() for _ in range(2)] for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_comp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    pipe_state = {"n": 0}

    def step_e2e_pipelined(i):
        n = pipe_state["n"]
        j = n % 2
        comp = torch.cuda.current_stream()
        a, b = pinned[i % n_sets]
        with torch.cuda.stream(h2d_stream):
            if n >= 2:
                h2d_stream.wait_event(ev_comp[j])          # the compute that read this input buffer two batches ago
            pin_dev[j][0].copy_(a, non_blocking=True)
            pin_dev[j][1].copy_(b, non_blocking=True)
            ev_in[j].record(h2d_stream)
        comp.wait_event(ev_in[j])
        with tf_util.use_store(store):
            pred, _, _ = MODEL.get_model(pin_dev[j][0], pin_dev[j][1], False, **kw)
        ev_comp[j].record(comp)
        if n >= 2:
            ev_out[j].synchronize()                        # batch n-2's distances are on the host before its buffer is reused
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(ev_comp[j])
            for t in (pred["pred_listAB"], pred["pred_listBA"]):
                t.record_stream(d2h_stream)
            pout_host[j][0].copy_(pred["pred_listAB"], non_blocking=True)
            pout_host[j][1].copy_(pred["pred_listBA"], non_blocking=True)
            ev_out[j].record(d2h_stream)
        if n >= 1:
            ev_out[1 - j].synchronize()                    # the caller consumes batch n-1's distances now
        pipe_state["n"] = n + 1
        return pout_host[j]

    def _pipe_finish():        # the timed region ends only when the last batch's distances are on the host
        for e in ev_out:
            torch.cuda.current_stream().wait_event(e)
    step_e2e_pipelined.finish = _pipe_finish

    evals_per_step_rank = CFG["pairs_per_gpu"] * 2 * CFG["NP"]

    def timed(step_fn, steps, warmup, sample_clocks=False):
        for i in range(warmup):
            step_fn(i)
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        l0 = lib.dpd_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step_fn(warmup + i)
        if hasattr(step_fn, "finish"):
            step_fn.finish()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = lib.dpd_launch_count() - l0
        clocks = sampler.stop() if sampler else None
        barrier()
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks

    ms, launches, clocks = timed(step_resident, args.steps, args.warmup, sample_clocks=True)
    value = evals_per_step_rank * world * args.steps / (ms * 1e-3)
    # The same step for >= 2 s: K = 20 steps last 60 ms, a burst at the boost clock; under sustained load the part settles
    # at its power cap.  Both are reported; `value` stays the contract's "exactly K steps".
    # DPD_BENCH_PROFILE=1 (ncu launch lists / captures of this very command): only the headline region and the per-kernel
    # times, none of the sub-records
    profile_only = bool(os.environ.get("DPD_BENCH_PROFILE"))
    ss_steps = args.steps if profile_only else max(args.steps, int(np.ceil(2000.0 / max(ms / args.steps, 1e-3))))
    if dist is not None:
        t = torch.tensor([ss_steps], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ss_steps = int(t.item())
    ms_ss, _, clocks_ss = timed(step_resident, ss_steps, 0, sample_clocks=True)
    steady_state = {"value": evals_per_step_rank * world * ss_steps / (ms_ss * 1e-3), "unit": UNIT, "steps": ss_steps,
                    "ms_per_step": ms_ss / ss_steps, "seconds": ms_ss * 1e-3, "clocks": clocks_ss}
    ms_e2e_sync, _, _ = timed(step_e2e, args.steps, max(3, args.warmup // 2))
    e2e_sync_value = evals_per_step_rank * world * args.steps / (ms_e2e_sync * 1e-3)
    ms_e2e, _, _ = timed(step_e2e_pipelined, args.steps, max(3, args.warmup // 2))
    e2e_value = evals_per_step_rank * world * args.steps / (ms_e2e * 1e-3)

    # per-kernel device times, measured live with CUDA events on the launching stream
    lib.dpd_profile_enable(1)
    _lib.profile_read(reset=True)
    psteps = args.steps if profile_only else max(args.steps, 50)
    for i in range(psteps):
        step_resident(i)
    torch.cuda.synchronize()
    prof = _lib.profile_read(reset=True)
    lib.dpd_profile_enable(0)
    peaks = measured_peaks()
    total_ms = sum(v[0] for v in prof.values()) or 1.0
    kernels = {k: {"ms_per_launch": v[0] / max(v[1], 1), "launches_per_step": v[1] / psteps, "share": v[0] / total_ms}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    dom = next(iter(kernels)) if kernels else None
    roofline, fv_kernel = None, None
    head_flops = {"l1": 2 * (3 + CFG["C"] * CFG["k"] ** 3) * CFG["H"], "l23": 2 * CFG["H"] ** 2}
    if dom is not None:
        per_launch_s = kernels[dom]["ms_per_launch"] * 1e-3
        if "l1" in dom or "gather" in dom:
            fl = head_flops["l1"] * evals_per_step_rank
        elif "fused" in dom:
            fl = FLOPS_PER_EVAL * evals_per_step_rank
        else:
            fl = head_flops["l23"] * evals_per_step_rank
        ach = fl / per_launch_s / 1e12
        # the kernel is timed inside a long step -> sustained peak
        peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        # dram__bytes_read.sum + dram__bytes_write.sum per launch cannot be measured inside this run (it needs ncu's
        # replay); it is read from the committed capture of this command on the current kernels, or null
        traffic, traffic_src = ncu_traffic(dom)
        roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "tensor_work_frac": 3 * ach / peak, "peak_source": peaks["source"] + ", dense bf16 sustained; the fp32-accurate fp16x3 split issues 3 tensor "
                    "passes per algorithmic flop, so frac <= 0.333 (DESIGN.md 4.2)", "tensor_passes_per_flop": 3, "share_of_step": kernels[dom]["share"]}
    fvk = [k for k in kernels if k.startswith("fv")]
    if fvk:
        s = kernels[fvk[0]]["ms_per_launch"] * 1e-3
        ach = 2 * CFG["pairs_per_gpu"] * FV_BYTES_PER_CLOUD / s / 1e9
        fv_traffic, fv_traffic_src = ncu_traffic(fvk[0])
        fv_kernel = {"kernel": fvk[0], "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": ach / peaks["hbm_gbs"], "traffic": fv_traffic, "traffic_source": fv_traffic_src,
                     "note": "nominally HBM-bound; in practice bound by fp32 / max-min issue slots of the all-pairs loop "
                             "(DESIGN.md 4.1)", "clouds_per_launch": 2 * CFG["pairs_per_gpu"],
                     "bytes_per_cloud": FV_BYTES_PER_CLOUD, "peak_source": peaks["source"]}

    # steady-state 3DmFV bandwidth: 65536 clouds per launch (32 x the in-step launch) so launch and tail effects amortise,
    # through the C ABI into a preallocated output (2.7 GB > L2: every launch writes through to HBM)
    fv_large = None
    if fv_kernel is not None and not profile_only:
        import ctypes
        g = torch.Generator(device="cpu").manual_seed(5)
        big = (torch.rand((65536, CFG["N"], 3), generator=g) * 1.6 - 0.8).to(dev)
        out_big = torch.empty((big.shape[0], CFG["G"] ** 3, CFG["C"]), device=dev, dtype=torch.float32)
        _, lgrid = dpdist_util._fv_grid(CFG["G"] ** 3, 3)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

        def fv_call():
            rc = lib.dpd_fv_forward(ctypes.c_void_p(big.data_ptr()), big.shape[0], CFG["N"], CFG["G"], _lib.fptr(lgrid), CFG["sigma"], 1, 0,
                                    ctypes.c_void_p(out_big.data_ptr()), stream)
            _lib.check(rc, "dpd_fv_forward")
        for _ in range(3):
            fv_call()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        ev0.record()
        for _ in range(reps):
            fv_call()
        ev1.record()
        torch.cuda.synchronize()
        s = ev0.elapsed_time(ev1) * 1e-3 / reps
        ach = big.shape[0] * FV_BYTES_PER_CLOUD / s / 1e9
        fv_large = {"clouds_per_launch": int(big.shape[0]), "ms_per_launch": s * 1e3, "achieved": ach, "unit": "GB/s",
                    "peak": peaks["hbm_gbs"], "frac": ach / peaks["hbm_gbs"], "clouds_per_s": big.shape[0] / s,
                    "note": "dpd_fv_forward into a preallocated output, CUDA events around 10 back-to-back launches; the all-pairs "
                            "arithmetic itself caps this kernel at 0.27 of the HBM peak (DESIGN.md 4.1)"}
        fv_kernel["steady_state"] = fv_large
        del big, out_big

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def event_time(fn, n, warm):
        for _ in range(warm):
            fn()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return max_over_ranks(t0.elapsed_time(t1) / n)

    def guarded(fn):
        """Sub-records are informational and must never break the bench line."""
        try:
            return fn()
        except Exception as e:      # noqa: BLE001
            torch.cuda.synchronize()
            return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    # BASELINE configs[2] (training loop) beside the headline: forward + tensor-core backward + two-bucket gradient
    # all-reduce + one Adam launch on 1024 synthetic pairs per GPU (weak scaling).
    def run_train():
        from dpdist_b200 import train as TR, synthetic
        trainer = TR.DPDistTrainer(dev, seed=1)
        pa, pb, lab = synthetic.uniform_batch(2 + 1000 * rank, CFG["pairs_per_gpu"], CFG["N"])
        ta, tb, tl = (torch.tensor(x, device=dev) for x in (pa, pb, lab))
        tsteps = 30
        tms = event_time(lambda: trainer.step(ta, tb, tl), tsteps, 3)
        consistent = bool(trainer.ranks_consistent())
        trainer.close()
        return {"workload": "configs[2]: DPDist training step, %d pairs per GPU, Adam, gradient all-reduce (2 buckets)" % CFG["pairs_per_gpu"],
                "ms_per_step": tms, "pairs_per_s": CFG["pairs_per_gpu"] * world / (tms * 1e-3), "steps": tsteps,
                # every rank must hold bit-identical weights after the averaged updates (:936-974)
                "ranks_consistent": consistent}

    # The reference's own training configuration (global batch 16, train...py:57) sharded over the ranks: strong
    # scaling of a launch-bound step, eager and as one captured CUDA graph (NCCL all-reduce inside the graph).
    def run_strong16():
        from dpdist_b200 import train as TR, synthetic
        if 16 % world != 0:
            return {"skipped": "16 pairs do not divide over %d ranks" % world}
        pa, pb, lab = synthetic.chair_batch(7, 16, CFG["N"])
        mine = [torch.tensor(np.ascontiguousarray(TR.shard(x, rank, world)), device=dev) for x in (pa, pb, lab)]
        out = {"workload": "configs[2] at the reference batch: 16 pairs global (%d per GPU), N=NP=64" % (16 // world), "scaling": "strong"}
        for name, graph in (("eager", False), ("cuda_graph", True)):
            trainer = TR.DPDistTrainer(dev, seed=1, cuda_graph=graph)
            ms_ = event_time(lambda: trainer.step(*mine), 50, 8)
            out[name] = {"ms_per_step": ms_, "pairs_per_s": 16 / (ms_ * 1e-3), "ranks_consistent": bool(trainer.ranks_consistent()),
                         "captured": trainer._graph is not None}
            trainer.close()        # the captured NCCL kernels must be gone before the communicator is destroyed
            del trainer
        return out

    # BASELINE configs[4] (stress): 4096 pairs, N = NP = 512 over the ranks (strong scaling), forward only, swept over
    # the Gaussian grid G and the patch edge k as SURVEY 8(d) asks.
    def run_stress():
        from dpdist_b200 import synthetic
        total_pairs, Np = 4096, 512
        if os.environ.get("DPD_BENCH_STRESS_PAIRS"):
            total_pairs = int(os.environ["DPD_BENCH_STRESS_PAIRS"])
        if total_pairs % world != 0:
            return {"skipped": "%d pairs do not divide over %d ranks" % (total_pairs, world)}
        pairs = total_pairs // world
        g = torch.Generator(device="cpu").manual_seed(11 + rank)
        sa = (torch.rand((pairs, Np, 3), generator=g) * 1.6 - 0.8).to(dev)
        sb = (torch.rand((pairs, Np, 3), generator=g) * 1.6 - 0.8).to(dev)
        out = {"workload": "configs[4]: %d pairs global (%d per GPU), N=NP=%d, forward-only" % (total_pairs, pairs, Np),
               "scaling": "strong", "sweep": []}
        for G, k in ((8, 5), (8, 3), (5, 5), (5, 3)):
            st = tf_util.VariableStore(device=dev, seed=1)
            kw2 = dict(bn=0, Embedding_Size=G ** 3, k=k, sigma3dmfv=1.0 / G, localSNmlp=[CFG["H"]] * 3)

            def one():
                with tf_util.use_store(st):
                    return MODEL.get_model(sa, sb, False, **kw2)[0]
            ms_ = event_time(one, 3, 1)
            evals = 2 * total_pairs * Np
            flops = 2 * ((3 + CFG["C"] * k ** 3) * CFG["H"] + 2 * CFG["H"] ** 2 + 3 * CFG["H"])
            out["sweep"].append({"G": G, "k": k, "sigma": 1.0 / G, "ms_per_batch": ms_, "evals_per_s": evals / (ms_ * 1e-3),
                                 "tflops_algorithmic": evals * flops / (ms_ * 1e-3) / 1e12})
            del st
        return out

    # BASELINE configs[3]: PCRNet trained with the DPDist loss (iterative_PCRNet_ours.py), batch 16, 8 pose refinements,
    # single GPU by specification: every rank would run the same thing, so rank 0's figure is reported.
    def run_pcrnet():
        from dpdist_b200 import pcrnet_ours as P
        from dpdist_b200.dpdist_loss import DPDistLoss
        out = {"workload": "configs[3]: PCRNet-ours training step, batch 16 x 64 points, 8 refinements, DPDist loss fwd + bwd into input1"}
        tpl = torch.tensor(P.synthetic_templates(16, 64, seed=3), device=dev)
        rng = np.random.default_rng(0)
        src = torch.tensor(P.apply_transformation(tpl.cpu().numpy(), P.generate_poses(16, rng)), device=dev)
        for name, graph in (("eager", False), ("cuda_graph", True)):
            dpd = DPDistLoss(num_point=64, device=dev, seed=1)
            with torch.no_grad():      # a "trained-like" DPDist: outputs spread over (0, 2) instead of the ~1e-4 of a fresh Xavier init
                for n_, v_ in dpd.store.vars.items():
                    if n_.endswith("mapper_conv1/weights"):
                        v_.mul_(600.0)
                    elif n_.endswith("weights") and "conv4" not in n_:
                        v_.mul_(2.0)
                    elif n_.endswith("mapper_conv4/biases"):
                        v_.add_(1.0)
            tr = P.IterativePCRNetOurs(dpd, 8, 0.001, False, dev, cuda_graph=graph)
            for _ in range(5):
                tr.train_step(src, tpl)
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(20):
                loss = tr.train_step(src, tpl)[0]
            t1.record()
            torch.cuda.synchronize()
            out[name] = {"ms_per_step": t0.elapsed_time(t1) / 20, "loss": float(loss)}
        return out

    train_info, configs = None, None
    if not profile_only:
        train_info = guarded(run_train)
        configs = {"strong_scaling_batch16": guarded(run_strong16), "stress": guarded(run_stress)}
        if rank == 0:
            configs["pcrnet_ours"] = guarded(run_pcrnet)
    barrier()

    cpu_baseline = None
    if rank == 0 and world == 1 and not os.environ.get("DPD_BENCH_NO_CPU") and not profile_only:
        # bounded sample of the same workload on this box's host cores (about 10-30 s of CPU work)
        v1, t1 = cpu_reference_rate(16, 8)
        pairs = int(min(16384, max(32, 12.0 / max(t1 / 16, 1e-6))))
        pairs -= pairs % 8
        v, t = cpu_reference_rate(pairs, 8)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                        "sample": "%d pairs (%d evals) of configs[1], literal torch-CPU restatement, %.1f s" % (pairs, pairs * 128, t)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": shared_config(world),
            "notes": "head: fp16x3 tcgen05, cta_group::2 pairs (fp32-grade results, 3 tensor passes per algorithmic flop); 3DmFV: fv_g8_ws_kernel",
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "mode": "batches streamed: H2D of batch i+1 and D2H of batch i-1 overlap the compute of batch i on separate "
                            "streams; the host receives every batch's distances",
                    "synchronous": {"value": e2e_sync_value, "ms_per_step": ms_e2e_sync / args.steps,
                                    "note": "one batch at a time: copy in, compute, copy out, synchronise"},
                    "h2d_bytes_per_step": 2 * CFG["pairs_per_gpu"] * CFG["N"] * 3 * 4,
                    "d2h_bytes_per_step": 2 * CFG["pairs_per_gpu"] * CFG["NP"] * 3 * 4},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "steady_state": steady_state,
            "roofline": roofline,
            "fv_kernel": fv_kernel,
            "kernels": kernels,
            "train": train_info,
            "configs": configs,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        # the line is out; never let communicator teardown hold the process (and the GPU box) hostage
        sys.stdout.flush()
        threading.Timer(30.0, lambda: os._exit(0)).start()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: dpdist_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
