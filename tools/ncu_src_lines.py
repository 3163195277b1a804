#!/usr/bin/env python
"""Per-source-line instruction counts from an `ncu --page source --csv` dump.
usage: python tools/ncu_src_lines.py <src.csv> <kernel substr> [top N]"""
import csv, collections, sys

path, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fn = None
hdr = None
per = {}
seen = set()
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == 'Function Name':
        fn = row[1]
        key = fn
        k = 0
        while key in seen:
            k += 1
            key = f'{fn}#{k}'
        seen.add(key)
        fn = key
        per[fn] = collections.Counter()
        continue
    if row[0] == 'Line No':
        hdr = {h: i for i, h in enumerate(row)}
        continue
    if fn is None or hdr is None or sub not in fn or row[0] in ('File Path',):
        continue
    try:
        n = int(row[hdr['Instructions Executed']].replace(',', '') or 0)
    except (ValueError, IndexError):
        continue
    per[fn][(row[0], row[1].strip()[:110])] += n
for fn, c in per.items():
    if sub not in fn or not c:
        continue
    tot = sum(c.values())
    print(f'== {fn}  total warp-instr {tot}')
    for (ln, src), n in c.most_common(top):
        print(f'  {n:10d} {100.0 * n / tot:5.1f}%  {ln:>5}  {src}')
