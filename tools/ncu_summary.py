#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` (per-launch key metrics) and a launch-list CSV into profiles/ text files.
usage: python tools/ncu_summary.py raw <prof_raw.csv>     |    python tools/ncu_summary.py launches <launches.csv>"""
import collections, csv, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')
             and 'not_issued' not in h]
    for r in rows[2:]:
        print('----', r[ix['Kernel Name']].replace('psk::', '')[:90])
        for k in KEYS:
            if k in ix:
                print(f'  {k:66s} {r[ix[k]]} {units[ix[k]]}')
        st = sorted(((float(r[ix[h]].replace(',', '') or 0), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')])
                     for h in stall), reverse=True)[:6]
        print('  top stalls (warps per issue): ' + ', '.join(f'{n} {v:.2f}' for v, n in st))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[ix['Metric Name']] != 'gpu__time_duration.sum':
            continue
        v = float(r[ix['Metric Value']].replace(',', ''))
        u = r[ix['Metric Unit']]
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(u, 1.0)
        a = agg[r[ix['Kernel Name']].replace('psk::', '')[:60]]
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{k:62s} n={n:3d} total={t:9.1f} us  avg={t / n:8.1f} us share={t / tot:.3f}')
    print(f'total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches')


if __name__ == '__main__':
    {'raw': raw, 'launches': launches}[sys.argv[1]](sys.argv[2])
