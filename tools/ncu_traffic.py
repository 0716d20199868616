"""profiles/ncu_traffic.json from an `ncu --set full` report of `python bench.py`:
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ;  python tools/ncu_traffic.py raw.csv [more.csv ...] "<how it was captured>"
Per kernel (DPD_LAUNCH name, see MAP): dram__bytes_read.sum + dram__bytes_write.sum of one launch (mean over the captured
launches), which bench.py quotes as roofline.traffic."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# kernel function name (regex) + discriminator -> the name bench.py's per-kernel profile uses
MAP = [(r"fv_g8_ws_kernel", "fv_g8_ws"), (r"fv_g8_kernel", "fv_g8"), (r"fv_generic_kernel", "fv_generic")]


def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


def main():
    raws = [a for a in sys.argv[1:] if a.endswith(".csv")]
    how = " ".join(a for a in sys.argv[1:] if not a.endswith(".csv"))
    out = {}
    acc = {}
    for raw in raws:
        rows = list(csv.reader(open(raw)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            rd = float(r[col["dram__bytes_read.sum"]]) * unit_scale(units[col["dram__bytes_read.sum"]])
            wr = float(r[col["dram__bytes_write.sum"]]) * unit_scale(units[col["dram__bytes_write.sum"]])
            dur = float(r[col["gpu__time_duration.sum"]])
            key = None
            for pat, k in MAP:
                if re.search(pat, name):
                    key = k
            if key is None and "tc_gemm2_kernel<1>" in name:
                key = "tc_gemm2_gather_l1_f16"
            if key is None and "tc_gemm2_kernel<0>" in name:
                # layers 2 and 3 share one instantiation: layer 2 writes the (hi, lo) activations (~0.5 GB at the bench size),
                # layer 3 with the fused output layer writes one float4 per row and column slice
                key = "tc_gemm2_dense_f16" if wr > 0.2 * rd else "tc_gemm2_dense_l3_l4_f16"
            if key is None:
                continue
            acc.setdefault(key, []).append((rd, wr, dur))
    for k, v in acc.items():
        n = len(v)
        out[k] = {"bytes": sum(a + b for a, b, _ in v) / n, "read": sum(a for a, _, _ in v) / n, "write": sum(b for _, b, _ in v) / n,
                  "launches": n, "duration_us_under_ncu": sum(d for _, _, d in v) / n, "source": how}
    dst = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
