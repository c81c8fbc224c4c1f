"""Extracts per-launch DRAM traffic / duration of the dominant kernel from an ncu report into a small JSON that
bench.py quotes as roofline.traffic.  usage: ncu_traffic.py rep.ncu-rep out.json"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    i = hdr.index(name)
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(units[i], 1.0)
    return [float(r[i]) * scale for r in data]


rd, wr, dur = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
grid = [int(float(r[hdr.index("launch__grid_size")])) for r in data]
res = {"report": rep, "kernel": "k_optimize", "launches": [
    {"grid_blocks": g, "duration_s": d, "dram_read_bytes": a, "dram_write_bytes": b, "dram_bytes": a + b}
    for g, d, a, b in zip(grid, dur, rd, wr)]}
res["dram_bytes_per_launch_mean"] = sum(x["dram_bytes"] for x in res["launches"]) / len(res["launches"])
res["patches_per_launch_mean"] = sum(grid) * 4 / len(grid)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
