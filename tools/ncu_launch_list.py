#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` launch list by kernel into profiles/.

    python tools/ncu_launch_list.py gpurun_out/final_launches.csv profiles/r01_bench_launches_final_ncu.txt "<command>"
"""
import collections
import csv
import sys


def main():
    src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for r in data:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        us = v / 1e3 if r[mu] in ("nsecond", "ns") else v if r[mu] in ("usecond", "us") else v * 1e3
        a = agg.setdefault(r[kn], [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
        n += 1
    # ncu prints the innermost namespace only: the kernels of namespace cm::<x> appear as <x>::name
    inner = ("cm::", "sp::", "spb::", "lc::", "lcb::", "wgb::", "wgf::", "dft::", "stem::", "ctc::", "lna::", "tma::")
    ours = sum(us for k, (c, us) in agg.items() if any(t in k for t in inner))
    with open(dst, "w") as f:
        f.write("ncu launch list of `%s` (first %d launches of the process, gpu__time_duration.sum, --clock-control none)\n" % (cmd, n))
        f.write("per-launch times are cold-cache and serialised by the profiler: the SHARE of each kernel is what is comparable "
                "with bench.py's kernel_time_share_ms\n")
        f.write("total %.1f us over %d launches; hand-written sm_100a kernels (namespace cm) %.1f us = %.1f %%\n\n"
                % (tot, n, ours, 100 * ours / tot))
        for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:80]:
            f.write("%9.1f us %5.1f%% %5d x  %s\n" % (us, 100 * us / tot, c, k[:150]))
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    main()
