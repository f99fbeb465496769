#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into a small JSON: python tools/ncu_summary.py rep [kernel_row_index | all] > out.json
(`all`: one compact JSON line per captured launch)"""
import csv, json, subprocess, sys
rep = sys.argv[1]; ALL = len(sys.argv) > 2 and sys.argv[2] == "all"; idx = 0 if ALL or len(sys.argv) <= 2 else int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]; units = rows[1]
if ALL:
    want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
            'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
    for r in rows[2:]:
        if len(r) < len(h): continue
        o = {k.split('.')[0].replace('__', '_'): (r[h.index(k)] + (' ' + units[h.index(k)] if units[h.index(k)] else '')) for k in want if k in h}
        st = {k.replace('smsp__pcsamp_warps_issue_stalled_', ''): float(r[i]) for i, k in enumerate(h)
              if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued') and r[i] not in ('', 'n/a')}
        tot = sum(st.values()) or 1
        o['top_stalls'] = {k: round(v / tot, 2) for k, v in sorted(st.items(), key=lambda x: -x[1])[:3]}
        print(json.dumps(o))
    sys.exit(0)
r = rows[2 + idx]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active']
out = {k: [r[h.index(k)], units[h.index(k)]] for k in keys if k in h}
st = {k.replace('smsp__pcsamp_warps_issue_stalled_', ''): int(float(r[i])) for i, k in enumerate(h)
      if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued') and r[i] not in ('', 'n/a')}
tot = sum(st.values()) or 1
out['stall_samples_share'] = {k: round(v / tot, 3) for k, v in sorted(st.items(), key=lambda x: -x[1]) if v / tot > 0.01}
print(json.dumps(out, indent=1))
