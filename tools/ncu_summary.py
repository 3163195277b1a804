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


MSG_KERNELS = {"k_rotconv": "rotconv", "k_warp_direct": "warp_in", "k_resample_bilinear": "resample (in + back)",
               "k_gauss_xy": "gauss_xy", "k_conv_cols_tma": "conv_cols", "k_conv_rows": "conv_rows", "k_epilogue": "epilogue"}


def traffic(path, n_messages):
    """DRAM bytes (read + write) of the message kernels in a `--set full` raw CSV that covers exactly one image, per message."""
    import json
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    units = rows[1]

    def val(r, k):
        v = float(r[ix[k]].replace(',', '') or 0)
        return v * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(units[ix[k]], 1.0)
    by = collections.defaultdict(lambda: [0, 0.0, 0.0])
    other = 0.0
    for r in rows[2:]:
        name = r[ix['Kernel Name']]
        b = val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum')
        t = float(r[ix['gpu__time_duration.sum']].replace(',', '') or 0) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(units[ix['gpu__time_duration.sum']], 1.0)
        for key, klass in MSG_KERNELS.items():
            if key in name:
                by[klass][0] += 1
                by[klass][1] += b
                by[klass][2] += t
                break
        else:
            other += b
    total = sum(v[1] for v in by.values())
    out = {"message": {"dram_bytes": round(total / n_messages), "messages": n_messages, "launches": sum(v[0] for v in by.values()),
                       "by_kernel": {k: {"launches": v[0], "dram_bytes_per_message": round(v[1] / n_messages),
                                         "cold_us_per_message": round(v[2] / n_messages, 2)} for k, v in by.items()}},
           "other_kernels_dram_bytes_per_image": round(other),
           "_source": "ncu --set full --clock-control none over the launches of one image (cold caches, serialised); dram__bytes_read.sum + "
                      "dram__bytes_write.sum of the message kernels divided by the %d messages of the image" % n_messages}
    print(json.dumps(out, indent=1))


def lastimage(path):
    """Prints "<launches in the list> <launches of the last image>": the last image starts at the first launch of its ingest."""
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    ix = {h: i for i, h in enumerate(rows[0])}
    names = [r[ix['Kernel Name']] for r in rows[1:] if r[ix['Metric Name']] == 'gpu__time_duration.sum']
    last_scatter = max(i for i, n in enumerate(names) if 'k_ingest' in n)
    # the ingest of an image is [k_fill] k_set_int k_ingest_*: the fill is skipped from the second image on the same
    # lattice on, so walk back over whichever of the two are there
    start = last_scatter
    while start > 0 and ('k_fill' in names[start - 1] or 'k_set_int' in names[start - 1]):
        start -= 1
    print(len(names), len(names) - start)


if __name__ == '__main__':
    if sys.argv[1] == 'lastimage':
        lastimage(sys.argv[2])
    elif sys.argv[1] == 'traffic':
        traffic(sys.argv[2], int(sys.argv[3]))
    else:
        {'raw': raw, 'launches': launches}[sys.argv[1]](sys.argv[2])
