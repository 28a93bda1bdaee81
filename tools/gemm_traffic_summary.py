#!/usr/bin/env python
"""Aggregate the per-launch ncu pass over the tcgen05 GEMM / conv launches of one UNet step
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
      -k regex:tc_gemm --clock-control none --csv --log-file gemm_traffic.csv python tools/profile_step.py
into (a) the JSON bench.py reads for `roofline.traffic` and (b) a per-template table with the time-weighted tensor-pipe activity.
Usage: python tools/gemm_traffic_summary.py gemm_traffic.csv out.json"""
import csv
import gzip
import json
import re
import sys
from collections import defaultdict

src, dst = sys.argv[1], sys.argv[2]
op = gzip.open if src.endswith(".gz") else open
rows = list(csv.reader(l for l in op(src, "rt") if l.startswith('"')))
hdr = rows[0]
ix = {k: i for i, k in enumerate(hdr)}
launch = defaultdict(dict)
for r in rows[1:]:
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    name = r[ix["Metric Name"]]
    if name == "gpu__time_duration.sum":
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)        # -> us
    if name.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    launch[r[ix["ID"]]][name] = v
    launch[r[ix["ID"]]]["kernel"] = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("ia2p::", "").strip()
T = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
by = defaultdict(lambda: [0, 0.0, 0.0])
tot_t = tot_w = rd = wr = 0.0
for d in launch.values():
    t = d["gpu__time_duration.sum"]
    by[d["kernel"]][0] += 1
    by[d["kernel"]][1] += t
    by[d["kernel"]][2] += t * d[T]
    tot_t += t
    tot_w += t * d[T]
    rd += d["dram__bytes_read.sum"]
    wr += d["dram__bytes_write.sum"]
n = len(launch)
out = dict(launches=n, dram_bytes_per_launch=(rd + wr) / n, dram_read_bytes_total=rd, dram_write_bytes_total=wr,
           serialized_time_us_total=tot_t, tensor_pipe_active_pct_time_weighted=tot_w / tot_t,
           per_template={k: dict(launches=c, ms=round(t / 1e3, 3), tensor_pipe_active_pct=round(w / t, 1)) for k, (c, t, w) in by.items()},
           source="ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
                  "sm__pipe_tensor_cycles_active... -k regex:tc_gemm --clock-control none over one eager UNet step (tools/profile_step.py)")
json.dump(out, open(dst, "w"), indent=1)
print("| template | launches | ms | tensor pipe active |\n|---|---:|---:|---:|")
for k, (c, t, w) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {c} | {t / 1e3:.1f} | {w / t:.1f} % |")
print(f"| all {n} | | {tot_t / 1e3:.1f} | **{tot_w / tot_t:.1f} %** |")
