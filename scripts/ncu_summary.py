"""Summarise an .ncu-rep (first kernel): key metrics + hottest CUDA source lines.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [nlines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__shared_mem_per_block_dynamic',
        'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']
for k in keys:
    if k in d: print(f"{k:72s} {d[k][1]} {d[k][0]}")
st = [(h, float(v.replace(',', ''))) for h, (u, v) in d.items() if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
print("stalls per issue:", ", ".join(f"{h.split('stalled_')[1].split('_per')[0]}={v:.2f}" for h, v in sorted(st, key=lambda x: -x[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
out = []
for r in rows[3:]:
    if len(r) > 8 and r[0] != '' and r[2] == '-':
        try: out.append((float(r[7]), float(r[6]), r[0], r[1], float(r[8])))
        except ValueError: pass
tot = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print(f"total warp-inst {tot:.4g}, samples {ts:.0f}")
for n, s, ln, code, ti in sorted(out, key=lambda x: -x[1])[:nl]:
    print(f"{100*n/tot:5.1f}% inst {100*s/ts:5.1f}% samp thr/inst {ti/max(n,1):4.1f} L{ln:>4}: {code.strip()[:95]}")
