"""Reads the timelines DPD_TC_TRACE=<prefix> makes the 2-CTA GEMM kernel write (cluster 0, leader CTA, one stamping thread
per warp role; head_tc.cu launch2) and prints where every role spends its cycles.
    DPD_TC_TRACE=gpurun_out/trace python tools/fwd_time.py 30 ; python tools/tc_trace.py gpurun_out/trace.*.bin"""
import sys

import numpy as np

TAGS = {1: "mma wait full", 2: "mma full ready", 3: "mma issued", 4: "mma wait seg_empty", 5: "mma seg_empty ready",
        6: "epi wait seg_full", 7: "epi seg_full ready", 8: "epi drained", 9: "epi store begin",
        11: "gather wait empty", 12: "gather empty ready", 13: "gather issued", 14: "tma wait empty", 15: "tma empty ready"}


def intervals(tags, clk, a, b):
    """durations between every tag a and the next tag b"""
    out = []
    t0 = None
    for t, c in zip(tags, clk):
        if t == a:
            t0 = c
        elif t == b and t0 is not None:
            out.append(c - t0)
            t0 = None
    return np.array(out, dtype=np.int64)


def report(path):
    raw = np.fromfile(path, dtype=np.uint64).reshape(4, 65536)
    print("==", path)
    roles = {}
    for r, name in enumerate(("tma", "mma", "epi", "gather")):
        v = raw[r][raw[r] != 0]
        tags = (v >> np.uint64(56)).astype(np.int64)
        clk = (v & np.uint64((1 << 56) - 1)).astype(np.int64)
        roles[name] = (tags, clk)
    tags, clk = roles["mma"]
    if len(clk) == 0:
        print("  empty")
        return
    total = clk[-1] - clk[0]
    kb = int((tags == 3).sum())
    wf = intervals(tags, clk, 1, 2)
    iss = intervals(tags, clk, 2, 3)
    se = intervals(tags, clk, 4, 5)
    print("  mma thread: %d K-blocks in %d cycles = %.0f cycles per K-block (ideal 1536)" % (kb, total, total / max(kb, 1)))
    print("    waiting for full   : %5.1f %% of the time, mean %.0f, p50 %.0f, p90 %.0f, max %d cycles per K-block" % (
        100.0 * wf.sum() / total, wf.mean(), np.percentile(wf, 50), np.percentile(wf, 90), wf.max()))
    print("    issuing 12 MMAs    : %5.1f %%, mean %.0f cycles per K-block" % (100.0 * iss.sum() / total, iss.mean()))
    print("    waiting seg_empty  : %5.1f %%, mean %.0f, max %d cycles per segment (%d segments)" % (
        100.0 * se.sum() / total, se.mean(), se.max(), len(se)))
    # K-block period histogram: time between consecutive 'issued' stamps
    issued = clk[tags == 3]
    per = np.diff(issued)
    print("    period between issues: p10 %.0f p50 %.0f p90 %.0f p99 %.0f" % tuple(np.percentile(per, [10, 50, 90, 99])))
    tags, clk = roles["epi"]
    if len(clk):
        w = intervals(tags, clk, 6, 7)
        d = intervals(tags, clk, 7, 8)
        # store time: from tag 9 to the next tag 6
        st = intervals(tags, clk, 9, 6)
        tot = clk[-1] - clk[0]
        print("  epilogue warp: waiting seg_full %5.1f %%, draining a segment mean %.0f cycles, tile store mean %.0f max %d cycles (%d tiles)" % (
            100.0 * w.sum() / tot, d.mean(), st.mean() if len(st) else 0, st.max() if len(st) else 0, len(st)))
    tags, clk = roles["gather"]
    if len(clk):
        w = intervals(tags, clk, 11, 12)
        g = intervals(tags, clk, 12, 13)
        tot = clk[-1] - clk[0]
        print("  gather thread: waiting empty %5.1f %% (mean %.0f), issuing a K-block mean %.0f p90 %.0f cycles" % (
            100.0 * w.sum() / tot, w.mean(), g.mean(), np.percentile(g, 90)))
    tags, clk = roles["tma"]
    if len(clk):
        w = intervals(tags, clk, 14, 15)
        tot = clk[-1] - clk[0]
        print("  tma thread   : waiting empty %5.1f %% (mean %.0f cycles)" % (100.0 * w.sum() / tot, w.mean()))
    # latency from the gather thread's / tma thread's "issued" to the MMA's "full ready" of the same K-block
    gt, gc = roles["gather"]
    mt, mc = roles["mma"]
    gi = gc[gt == 13]
    mf = mc[mt == 2]
    n = min(len(gi), len(mf))
    if n:
        lat = mf[:n] - gi[:n]
        print("  gather issued -> full ready: mean %.0f p50 %.0f p90 %.0f cycles" % (lat.mean(), np.percentile(lat, 50), np.percentile(lat, 90)))
    tt, tcl = roles["tma"]
    ti = tcl[tt == 15]
    n = min(len(ti), len(mf))
    if n:
        lat = mf[:n] - ti[:n]
        print("  tma empty ready (loads issued) -> full ready: mean %.0f p50 %.0f p90 %.0f cycles" % (lat.mean(), np.percentile(lat, 50), np.percentile(lat, 90)))
    me = mc[mt == 3]
    n = min(len(ti) - 3, len(me))
    if n > 0:
        lat = ti[3:3 + n] - me[:n]
        print("  mma issued (commit) of K-block i -> tma sees empty for K-block i+3: mean %.0f p50 %.0f p90 %.0f cycles" % (
            lat.mean(), np.percentile(lat, 50), np.percentile(lat, 90)))


for p in sys.argv[1:]:
    report(p)
