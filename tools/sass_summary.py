#!/usr/bin/env python
"""profiles/r02_sass_summary.txt: per kernel of libdetzero_b200.so, how many tcgen05 / TMEM / TMA / cp.async SASS instructions it
holds (cuobjdump -sass): UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store, LDGSTS = cp.async."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'detzero_b200', 'libdetzero_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
pats = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'LDGSTS', 'UBLKCP', 'SYNCS']
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace('(anonymous namespace)::', '')
        cur = re.sub(r'\(.*', '', cur)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for p in pats:
            if re.search(r'\b' + p + r'\b|\b' + p + r'\.', line):
                counts[cur][p] += 1
rows = [(k, v) for k, v in counts.items() if sum(v.values())]
w = max(len(k) for k, _ in rows)
lines = ['# cuobjdump -sass detzero_b200/libdetzero_b200.so | per-kernel counts (kernels without any of these are omitted)',
         '%-*s ' % (w, 'kernel') + ' '.join('%8s' % p for p in pats)]
for k, v in sorted(rows):
    lines.append('%-*s ' % (w, k) + ' '.join('%8d' % v[p] for p in pats))
tot = collections.Counter()
for _, v in rows:
    tot.update(v)
lines.append('%-*s ' % (w, 'TOTAL') + ' '.join('%8d' % tot[p] for p in pats))
txt = '\n'.join(lines) + '\n'
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r02_sass_summary.txt')
open(dst, 'w').write(txt)
print(txt)
