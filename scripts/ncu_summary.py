#!/usr/bin/env python
"""Summarise an ncu report: python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'launch__shared_mem_per_block_dynamic',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
idx = [hdr.index(w) for w in want if w in hdr]
ki = hdr.index('Kernel Name')
stall = [i for i, h in enumerate(hdr) if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
seen = set()
for r in rows[2:]:
    key = (r[ki][:14], r[hdr.index('launch__grid_size')], r[hdr.index('gpu__time_duration.sum')][:4])
    if key in seen:
        continue
    seen.add(key)
    print(r[ki][:40])
    for i in idx:
        print(f"    {hdr[i]:75s} {r[i]} {rows[1][i]}")
    tot = sum(float(r[i]) for i in stall) or 1
    print('    stalls:', [(hdr[i].replace('smsp__pcsamp_warps_issue_stalled_', ''), round(float(r[i]) / tot, 2)) for i in sorted(stall, key=lambda i: -float(r[i]))[:7]])

# --traffic-json FILE [--batch B]: per-kernel DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every
# captured launch, for bench.py's roofline.traffic
if '--traffic-json' in sys.argv:
    import json
    out_path = sys.argv[sys.argv.index('--traffic-json') + 1]
    batch = int(sys.argv[sys.argv.index('--batch') + 1]) if '--batch' in sys.argv else None
    ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    unit_r, unit_w = rows[1][ir], rows[1][iw]
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    per = {}
    for r in rows[2:]:
        name = r[ki].split('<')[0].replace('void ', '').strip()
        b = float(r[ir].replace(',', '')) * mult.get(unit_r, 1) + float(r[iw].replace(',', '')) * mult.get(unit_w, 1)
        per.setdefault(name, []).append(b)
    doc = {"source": "ncu --set full --clock-control none capture %s: dram__bytes_read.sum + dram__bytes_write.sum per launch" % sys.argv[1]}
    for k, v in per.items():
        doc[k] = {"batch": batch, "launches_captured": len(v), "dram_bytes_per_launch": [round(x) for x in v], "avg_dram_bytes_per_launch": sum(v) / len(v)}
    json.dump(doc, open(out_path, 'w'), indent=1)
    print('wrote', out_path)
