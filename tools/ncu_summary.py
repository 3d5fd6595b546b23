#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics per kernel and the hottest source lines.
usage: tools_ncu_summary.py rep [kernel-substr] [nlines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else None; nl = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_shared_atom.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.sum','smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    if kern and kern not in r[idx['Kernel Name']]: continue
    print('---', r[idx['Kernel Name']][:100])
    for w in want:
        if w in idx: print('  ', w, r[idx[w]], rows[1][idx[w]])
    st = [(h, float(r[i].replace(',', ''))) for h, i in idx.items() if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i] not in ('', 'n/a')]
    st.sort(key=lambda x: -x[1])
    print('   stalls:', [(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), round(v, 2)) for h, v in st[:8]])
if kern:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    cur = None; out = []; tot = 0; ts = 0; seen_kernel = 0
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
        if len(r) == 2 and r[0] == 'Function Name':
            continue
        if len(r) < 8 or r[0] in ('Line No',): continue
        if r[0] != '':
            try:
                out.append((cur, int(r[0]), r[1].strip()[:120], int(r[7]), int(r[4]))); tot += int(r[7]); ts += int(r[4])
            except Exception: pass
    print('total inst', tot, 'samples', ts)
    out.sort(key=lambda x: -x[4])
    for o in out[:nl]:
        print(f"{o[0]}:{o[1]:4d} inst={o[3]/max(tot,1)*100:5.1f}% samp={o[4]/max(ts,1)*100:5.1f}%  {o[2]}")
