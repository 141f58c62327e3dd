#!/usr/bin/env python
"""Summarise an ncu raw/source CSV pair of k_multipoles: pipe utilisation and per-point instruction mix."""
import collections
import csv
import json
import sys

raw, src = sys.argv[1], sys.argv[2]
npoints = float(sys.argv[3]) if len(sys.argv) > 3 else 65536 * 150000
out_json = sys.argv[4] if len(sys.argv) > 4 else None
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg',
        'sm__cycles_elapsed.avg.per_second', 'lts__t_bytes.sum', 'launch__grid_size', 'launch__block_size',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']
met = {}
for h, u, v in zip(hdr, units, vals):
    if h in keep:
        met[h] = f"{v} {u}".strip()
        print(f"{h:90s} {v} {u}")
rows = list(csv.reader(open(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
mx = max(float(r[ix['Instructions Executed']] or 0) for r in data)
hot = [r for r in data if float(r[ix['Instructions Executed']] or 0) > 0.05 * mx]
cls = collections.Counter()
wf = 0.0
for r in hot:
    parts = r[ix['Source']].split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    cls[op.split('.')[0]] += float(r[ix['Instructions Executed']])
    wf += float(r[ix['L1 Wavefronts Shared']] or 0)
wp = npoints / 32
mix = {k: round(v / wp, 2) for k, v in cls.most_common()}
print('hot-loop warp instructions per quadrature point:', mix)
tot = sum(cls.values()) / wp
fp64 = sum(v for k, v in cls.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX')) / wp
print(f'total {tot:.1f}  fp64 {fp64:.1f}  smem wavefronts/pt {wf / wp:.1f}')
if out_json:
    json.dump({'metrics': met, 'hot_loop_warp_instr_per_point': mix, 'total_per_point': tot,
               'fp64_per_point': fp64, 'smem_wavefronts_per_point': wf / wp}, open(out_json, 'w'), indent=1)
