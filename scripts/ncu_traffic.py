#!/usr/bin/env python
"""DRAM traffic of one bench launch -> profiles/r2_traffic.json (read by bench.py: roofline.traffic).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:whfast_steps \
        -s 1 -c 1 --csv --log-file gpurun_out/traffic.csv python scripts/profile_launch.py --steps 2000 --systems 65536 --arithmetic hybrid
    python scripts/ncu_traffic.py gpurun_out/traffic.csv c4_trappist1 65536 2000 hybrid
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, workload, systems, spc, arithmetic = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    iname, iunit, ival = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, ms = 0.0, None
    for r in rows[1:]:
        if r[iname].startswith("dram__bytes"):
            total += float(r[ival].replace(",", "")) * scale[r[iunit]]
        if r[iname] == "gpu__time_duration.sum":
            ms = float(r[ival].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[iunit]]
    out = os.path.join(ROOT, "profiles", "r2_traffic.json")
    data = json.load(open(out)) if os.path.exists(out) else []
    data = [d for d in data if (d["workload"], d["systems"], d["steps_per_call"], d["arithmetic"]) != (workload, systems, spc, arithmetic)]
    data.append({"workload": workload, "systems": systems, "steps_per_call": spc, "arithmetic": arithmetic, "dram_bytes": total,
                 "launch_ms_under_ncu": ms, "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch (scripts/ncu_traffic.py)"})
    json.dump(data, open(out, "w"), indent=1)
    print("%s x%d, %d steps, %s: %.1f MB per launch" % (workload, systems, spc, arithmetic, total / 1e6))


if __name__ == "__main__":
    main()
