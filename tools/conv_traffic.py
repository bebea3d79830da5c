#!/usr/bin/env python3
"""DRAM traffic of the tensor-core convolution launches of one benchmark step, from an ncu csv:
    ncu --profile-from-start off -k regex:conv_tc --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/conv_dram.csv python tools/step_once.py
    python tools/conv_traffic.py gpurun_out/conv_dram.csv profiles/traffic.json"""
import csv
import json
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
n = 0
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}
for row in csv.DictReader(lines):
    m = row.get("Metric Name")
    if m in tot:
        tot[m] += float(row["Metric Value"].replace(",", "")) * scale.get(row["Metric Unit"], 1)
        n += m == "gpu__time_duration.sum"
out = {"conv_tc_kernel_launches_per_step": n, "conv_tc_kernel_dram_bytes_per_step": tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"],
       "dram_read_bytes": tot["dram__bytes_read.sum"], "dram_write_bytes": tot["dram__bytes_write.sum"],
       "ncu_time_us": tot["gpu__time_duration.sum"], "source": "ncu dram__bytes_{read,write}.sum over the conv_tc launches of one step (tools/step_once.py)"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
