#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics per kernel and the
hot-loop instruction mix / stall reasons from the source page.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [warp_iters_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__block_size', 'launch__grid_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'smsp__cycles_active.avg', 'sm__inst_executed_pipe_lsu.sum', 'gpc__cycles_elapsed.max',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio']


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    witers = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = ncu(rep, "raw")
    hdr, units, data = raw[0], raw[1], raw[2:]
    for d in data:
        print("=" * 100)
        print(d[hdr.index('Kernel Name')])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-70s %s %s" % (k, d[i], units[i]))
    rows = ncu(rep, "source")
    kern, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            kern.append(cur)
            continue
        if r and r[0] == 'Address':
            cur['hdr'] = r
            continue
        if cur is not None and r:
            cur['rows'].append(r)
    seen = set()
    for k in kern:
        if k['name'] in seen:
            continue
        seen.add(k['name'])
        h = k['hdr']
        iS, iE, iN = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
        print("=" * 100)
        print(k['name'])
        st = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
        tot = collections.Counter()
        for r in k['rows']:
            for c in st:
                tot[c] += int(r[h.index(c)])
        s = sum(tot.values()) or 1
        print("  stalls %:", {c[6:]: round(100 * v / s, 1) for c, v in tot.most_common(9)})
        total_inst = sum(int(r[iE]) for r in k['rows'])
        mx = max(int(r[iE]) for r in k['rows'])
        op = collections.Counter()
        for r in k['rows']:
            e = int(r[iE])
            if e > 0.5 * mx:
                toks = [t for t in r[iS].split() if not t.startswith('@')]
                op[toks[0].split('.')[0]] += e
        hot = sum(op.values())
        norm = witers or (mx / 2.0)
        print("  total warp-inst %d, hot-loop %d; per warp-iteration (norm %.0f): total %.1f hot %.1f"
              % (total_inst, hot, norm, total_inst / norm, hot / norm))
        print("  hot mix:", ", ".join("%s %.1f" % (n, c / norm) for n, c in op.most_common(30)))
        for r in sorted(k['rows'], key=lambda r: -int(r[iN]))[:14]:
            stalls = {c[6:]: int(r[h.index(c)]) for c in st if int(r[h.index(c)]) > 0}
            best = sorted(stalls.items(), key=lambda x: -x[1])[:2]
            print("   %6s %9s  %-58s %s" % (r[iN], r[iE], r[iS].strip()[:58], best))


if __name__ == "__main__":
    main()
