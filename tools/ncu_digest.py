#!/usr/bin/env python
"""Digest of an ncu report: headline metrics + per-region instruction mix / stall samples of the attention kernel.
usage: python tools/ncu_digest.py gpurun_out/x.ncu-rep [--top 40]"""
import csv, subprocess, sys, collections, io

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.max.per_second',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_pipe_uniform.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for h, u, v in zip(hdr, units, vals):
    if h in want or any(h.endswith(w) for w in want):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
isrc, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
data = rows[2:]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "sass instrs", len(data))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top]
for i in sorted(idx):
    r = data[i]
    print(f"{i:5d} samples={int(r[isamp]):6d} ({100*int(r[isamp])/tot:4.1f}%) exec={r[iex]:>10s}  {r[isrc].strip()[:90]}")
