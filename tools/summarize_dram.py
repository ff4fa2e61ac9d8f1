"""DRAM traffic per GEMM launch from `tools/collect_profiles.sh dram` (ncu --metrics dram__bytes_read.sum,
dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm over one training step) -> the JSON bench.py reads for
`roofline.traffic`.  usage: python tools/summarize_dram.py gpurun_out/<tag>_gemm_dram.csv > profiles/rNN_gemm_traffic.json"""
import csv
import json
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
per = {}
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    name = row["Metric Name"]
    if name.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    else:
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}[unit]
    per.setdefault(row["ID"], {})[name] = v
n = len(per)
rd = sum(p["dram__bytes_read.sum"] for p in per.values()) / n
wr = sum(p["dram__bytes_write.sum"] for p in per.values()) / n
us = sum(p["gpu__time_duration.sum"] for p in per.values()) / n
print(json.dumps({
    "launches": n, "dram_read_bytes_per_launch": rd, "dram_write_bytes_per_launch": wr,
    "dram_bytes_per_launch": rd + wr, "avg_us_per_launch_under_ncu": us,
    "how": "ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
           "-k regex:gemm over the GEMM launches of one training step (ViT-B/32, 1024 pairs, packed text tower, 8-bit "
           "QuickGELU' save; tools/collect_profiles.sh dram), cold-cache serialised"}, indent=1))
