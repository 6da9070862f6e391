#!/usr/bin/env python
"""Reads dram__bytes_read.sum + dram__bytes_write.sum of the first kernel in an .ncu-rep and records it in
profiles/r02_traffic.json under KEY (e.g. cm_scan_bwd@conmamba_small_ctc_fwdbwd_b32x15s) - bench.py's roofline.traffic.

    python tools/ncu_traffic.py gpurun_out/x.ncu-rep KEY
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    rep, key = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    tot = sum(to_bytes(m[k][1], m[k][0]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[key] = tot
    d[key + ":source"] = "%s (%s)" % (os.path.basename(rep), m.get("Kernel Name", ("", "?"))[1])
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    print(key, tot)


if __name__ == "__main__":
    main()
