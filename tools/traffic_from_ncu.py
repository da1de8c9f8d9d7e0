#!/usr/bin/env python
"""profiles/traffic.json from an ncu CSV of `--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
over the launches of one bench step: DRAM bytes per launch (average over the step's launches), read by bench.py into
`roofline.traffic`.      python tools/traffic_from_ncu.py gpurun_out/x.csv launches_per_step [out.json]"""
import csv
import json
import sys


def main():
    path, per_step = sys.argv[1], int(sys.argv[2])
    out = sys.argv[3] if len(sys.argv) > 3 else "profiles/traffic.json"
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    per = {}
    for r in rd:
        k = r["ID"]
        d = per.setdefault(k, {"name": r["Kernel Name"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6}.get(unit, 1)
        d[r["Metric Name"]] = v * scale
    ids = sorted(per, key=int)
    ids = ids[-per_step:]  # the last full step captured
    rd_b = sum(per[i].get("dram__bytes_read.sum", 0) for i in ids)
    wr_b = sum(per[i].get("dram__bytes_write.sum", 0) for i in ids)
    t_ns = sum(per[i].get("gpu__time_duration.sum", 0) for i in ids)
    names = {}
    for i in ids:
        n = per[i]["name"].split("(")[0][-60:]
        a = names.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += per[i].get("gpu__time_duration.sum", 0)
    res = {"dram_bytes_per_launch": round((rd_b + wr_b) / len(ids)), "launches": len(ids), "dram_read_bytes_step": rd_b,
           "dram_write_bytes_step": wr_b, "ncu_time_us_step": t_ns / 1e3,
           "kernels": {k: {"launches": v[0], "ncu_us": round(v[1] / 1e3, 1)} for k, v in names.items()}, "source": path}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
