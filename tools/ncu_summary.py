#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): key raw metrics + top stall sites.  Usage: python tools/ncu_summary.py rep [ntop]"""
import csv, subprocess, sys, io

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "smsp__cycles_active.avg", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_srcunit_tex.sum", "lts__t_bytes.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in KEYS:
        if k in d:
            print(f"{k} = {d[k]} {units[hdr.index(k)]}")
    print()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ix = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
print("total samples", tot)
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:ntop]:
    s = int(r[ix["# Samples"]])
    st = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{s:6d} {100 * s / tot:5.1f}%  {r[ix['Address']][-5:]}  {r[ix['Source']][:80]:80s} {st}")
