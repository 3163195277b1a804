#!/usr/bin/env python
"""Static SASS instruction mix of the kernels whose (mangled) name contains a substring.
usage: python tools/sass_mix.py <substr> [lib] [--dump]"""
import collections, re, subprocess, sys

sub = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith('--') else 'partapp_b200/libpsinfer.so'
dump = '--dump' in sys.argv
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
fn = None
mix = collections.defaultdict(collections.Counter)
pat = re.compile(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;?\s*/\*')
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = m.group(1)
        continue
    if fn is None or sub not in fn:
        continue
    m = pat.match(line)
    if not m:
        continue
    ins = re.sub(r'^@!?U?P\w+\s+', '', m.group(2))
    if dump:
        print(m.group(1), m.group(2))
    op = ins.split()[0].rstrip(';')
    mix[fn][op.split('.')[0]] += 1
for f, c in mix.items():
    print(f, sum(c.values()))
    print('  ' + ', '.join(f'{k}:{v}' for k, v in c.most_common(30)))
