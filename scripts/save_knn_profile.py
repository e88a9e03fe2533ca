#!/usr/bin/env python
"""Turn an ncu capture of the dominant kNN kernel (one launch of the full workload) into the small JSON
bench.py reads for `roofline.traffic` / `roofline_issue`, stamped with the hash of the kernel source the
capture was taken from (bench.py ignores it when the source has changed since).

    python scripts/save_knn_profile.py gpurun_out/x.ncu-rep profiles/r02_knn_duo_c4_ncu.json [csv-out]
"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def num(key):
    if key not in d:  # a reduced section set was captured
        return float("nan")
    u, v = d[key]
    x = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "us": 1e-3, "ms": 1.0, "s": 1e3,
             "ns": 1e-6}.get(u)
    return x * scale if scale is not None else x


src = os.path.join(ROOT, "ptv_interpolation_b200", "csrc", "knn_duo.cu")
with open(src, "rb") as f:
    sha = hashlib.sha256(f.read()).hexdigest()
stalls = {h.split("stalled_")[1].split("_per")[0]: float(v.replace(",", "")) for h, (u, v) in d.items()
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")}
prof = {
    "kernel": d["Kernel Name"][1],
    "kernel_source": "ptv_interpolation_b200/csrc/knn_duo.cu",
    "kernel_source_sha256": sha,
    "capture": os.path.basename(rep),
    "kernel_ms_under_ncu": num("gpu__time_duration.sum"),
    "registers_per_thread": int(num("launch__registers_per_thread")) if "launch__registers_per_thread" in d else None,
    "dynamic_smem_bytes": num("launch__shared_mem_per_block_dynamic"),
    "grid_size": int(num("launch__grid_size")) if "launch__grid_size" in d else None,
    "warp_inst_per_launch": num("smsp__inst_executed.sum"),
    "threads_per_inst": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "warps_eligible_per_cycle": num("smsp__warps_eligible.avg.per_cycle_active"),
    "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
    "dram_bytes_read": num("dram__bytes_read.sum"),
    "dram_bytes_write": num("dram__bytes_write.sum"),
    "lsu_wavefronts_pct": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "fp64_pipe_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "stalls_per_issue": dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8]),
}
with open(out, "w") as f:
    json.dump(prof, f, indent=1)
print(json.dumps(prof, indent=1))
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        f.write(raw)
