#!/usr/bin/env python
"""Turns ncu CSV launch lists (`ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum ...`)
into profiles/r02_traffic.json, the file bench.py reads `roofline.traffic` from.

    python tools/ncu_traffic.py compact2x_1080p gpurun_out/r02_pipe_traffic.csv 16 tc_pipe_kernel \
           valar_540p gpurun_out/r02_valar_traffic.csv 4 ""

Per entry: key, CSV, frames per captured pass, kernel-name filter ("" = every kernel in the CSV = one whole pass)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_traffic.json")


def parse(path, name_filter):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hdr_i = next(i for i, r in enumerate(rows) if "Metric Name" in r and "Kernel Name" in r)
    hdr = rows[hdr_i]
    k, m, u, v = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    idc = hdr.index("ID")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
             "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
    launches = set()
    for r in rows[hdr_i + 1:]:
        if len(r) <= max(k, m, u, v) or r[m] not in tot or (name_filter and name_filter not in r[k]):
            continue
        tot[r[m]] += float(r[v].replace(",", "")) * scale[r[u]]
        launches.add(r[idc])
    return tot, len(launches)


def main(argv):
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for i in range(0, len(argv), 4):
        key, path, frames, filt = argv[i], argv[i + 1], int(argv[i + 2]), argv[i + 3]
        tot, n = parse(path, filt)
        data[key] = {"kernel": filt or "all kernels of one pass", "launches_summed": n, "frames_per_launch": frames,
                     "dram_read_bytes": tot["dram__bytes_read.sum"], "dram_write_bytes": tot["dram__bytes_write.sum"],
                     "gpu_time_ms_under_ncu": tot["gpu__time_duration.sum"],
                     "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, "
                               "profiles/" + os.path.basename(path)}
        print(key, json.dumps(data[key]))
    json.dump(data, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1:])
