"""Extract the numbers bench.py's `roofline` object quotes (DRAM bytes per launch, tensor-pipe utilisation, duration)
from an `ncu --set full` report into profiles/roofline_kernels.json, keyed by kernel shape, naming the capture file.

    ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > profiles/r02_ncu_full_gemm.csv
    python scripts/ncu_extract.py profiles/r02_ncu_full_gemm.csv conv_gemm_512_1024_k4s1_b32=conv_gemm_persistent_kernel:0
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    out_path = os.path.join(ROOT, "profiles", "roofline_kernels.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for spec in sys.argv[2:]:
        key, sel = spec.split("=")
        name, idx = sel.split(":")
        match = [r for r in data if name in r[col["Kernel Name"]]]
        r = match[int(idx)]
        f = lambda m: float(r[col[m]].replace(",", "")) if m in col and r[col[m]] not in ("", "n/a") else None
        unit = lambda m: rows[1][col[m]] if m in col else ""
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = f("dram__bytes_read.sum") * scale.get(unit("dram__bytes_read.sum"), 1)
        wr = f("dram__bytes_write.sum") * scale.get(unit("dram__bytes_write.sum"), 1)
        dur = f("gpu__time_duration.sum")
        tp = None
        for m in ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                  "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"):
            if f(m) is not None:
                tp = f(m)
                break
        out[key] = {"kernel": r[col["Kernel Name"]][:120], "dram_bytes": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
                    "duration": dur, "duration_unit": unit("gpu__time_duration.sum"), "tensor_pipe_pct": tp,
                    "capture": os.path.relpath(path, ROOT)}
        print(key, out[key])
    json.dump(out, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
