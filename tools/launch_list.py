#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum --csv` log into the launch list kept under profiles/:
the kernels of the LAST replay of the step graph with their share of the step, and launch counts per kernel.
usage: tools/launch_list.py launches.csv "command line" > profiles/rNN_launch_list_X.txt"""
import collections
import csv
import re
import sys

STEP = ("quads_list_kernel", "quads_kernel", "prepared_forward_kernel", "ray_sweep_kernel", "residual_kernel", "ne_rows_kernel",
        "backproject_wruns_kernel", "backproject_wsegments_kernel", "backproject_combine_short_kernel",
        "backproject_combine_kernel", "prepared_adjoint_kernel", "finish_gradient_kernel", "finish_compact_kernel",
        "adjoint_runs_kernel", "peer_reduce_expand_kernel", "zero_kernel", "mul_kernel", "ne_from_m_kernel")


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)
    return name


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    i_name, i_val, i_metric = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    i_unit = hdr.index("Metric Unit")
    launches = []
    for r in rows:
        if r is hdr or r[i_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[i_val].replace(",", ""))
        v = {"ns": v / 1e6, "us": v / 1e3, "usecond": v / 1e3, "nsecond": v / 1e6, "msecond": v, "ms": v, "s": v * 1e3, "second": v * 1e3}[r[i_unit]]
        launches.append((short(r[i_name]), v))
    print("# ncu --metrics gpu__time_duration.sum --clock-control none")
    print("# command: %s" % (sys.argv[2] if len(sys.argv) > 2 else "?"))
    print("# %d launches captured in total; below: the kernels of the LAST replay of the step graph "
          "(cold-cache, serialised times)" % len(launches))
    # last replay: walk back from the end while the kernels belong to the step, stop when a kernel repeats
    last, seen = [], set()
    for name, ms in reversed(launches):
        base = name.split("<")[0]
        if base not in STEP or base in seen:
            if last:
                break
            continue
        seen.add(base)
        last.append((name, ms))
    last.reverse()
    tot = sum(ms for _, ms in last)
    print("%-52s %8s %7s" % ("kernel", "ms", "share"))
    for name, ms in last:
        print("%-52s %8.4f %6.1f%%" % (name[:52], ms, 100 * ms / tot if tot else 0))
    print("%-52s %8.4f" % ("sum", tot))
    print()
    print("# launches per kernel over the whole run (set-up, warm-up, component timing, steps):")
    cnt = collections.Counter(n for n, _ in launches)
    for name, c in cnt.most_common(24):
        print("%-60s %5d" % (name[:60], c))


if __name__ == "__main__":
    main()
