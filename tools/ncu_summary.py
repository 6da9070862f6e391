#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_x.txt [--top 25]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, zip(units, r))) for r in rows[2:]]


def source(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], hdr=None, rows=[])
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    return blocks


def main():
    path, dst = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 20
    lines = ["ncu summary of %s" % path, ""]
    for k in raw(path):
        lines.append("== %s" % k.get("Kernel Name", ("", "?"))[1])
        for key in KEYS:
            if key in k:
                lines.append("  %-70s %s %s" % (key, k[key][1], k[key][0]))
        lines.append("  -- warp stall reasons (per issue active)")
        for h, (u, v) in k.items():
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                lines.append("  %-70s %s" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
        lines.append("")
    for b in source(path)[:1]:
        ix = {h: i for i, h in enumerate(b["hdr"])}

        def f(r, key):
            try:
                return float(r[ix[key]])
            except Exception:
                return 0.0
        tot = sum(f(r, "# Samples") for r in b["rows"]) or 1.0
        lines.append("== hottest SASS instructions by stall samples: %s (%d instructions, %d samples)" % (b["name"], len(b["rows"]), tot))
        for r in sorted(b["rows"], key=lambda r: -f(r, "# Samples"))[:top]:
            reasons = sorted(((f(r, k), k) for k in ix if k.startswith("stall_") and "Not Issued" not in k), reverse=True)[:2]
            lines.append("  %5.1f%%  %-60s %s" % (100 * f(r, "# Samples") / tot, r[ix["Source"]][:60],
                                                " ".join("%s=%d" % (k, v) for v, k in reasons if v > 0)))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
