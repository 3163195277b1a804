#!/usr/bin/env python
"""Executed-instruction mix by SASS opcode from an `ncu --page source --csv` dump (first matching launch).
usage: python tools/ncu_src_ops.py <src.csv> <kernel substr>"""
import csv, collections, re, sys

path, sub = sys.argv[1], sys.argv[2]
fn, hdr, c, done = None, None, collections.Counter(), False
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] in ('Function Name', 'Kernel Name'):
        if fn and sub in fn and c:
            break
        fn = row[1]
        continue
    if row[0] in ('Line No', 'Address'):
        hdr = {}
        for i, h in enumerate(row):
            hdr.setdefault(h, i)
        src_cols = [i for i, h in enumerate(row) if h == 'Source']
        continue
    if fn is None or hdr is None or sub not in fn:
        continue
    try:
        n = int(row[hdr['Instructions Executed']].replace(',', '') or 0)
    except (ValueError, IndexError):
        continue
    sass = row[src_cols[-1]]
    if not row[hdr['Address']].strip() or sass.strip() in ('', '-'):
        continue  # source-line aggregate rows
    sass = re.sub(r'^\s*@!?U?P\w+\s+', '', sass.strip())
    op = sass.split()[0].split('.')[0] if sass else '?'
    c[op] += n
tot = sum(c.values())
print(fn, 'total warp-instr', tot)
for op, n in c.most_common(25):
    print(f'  {n:10d} {100.0 * n / tot:5.1f}%  {op}')
