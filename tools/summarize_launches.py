"""summarize an ncu gpu__time_duration launch list (csv) by kernel name: count, total us; optional --list prints every launch"""
import csv, sys, collections
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(rows); hdr = next(r)
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
agg = collections.OrderedDict(); tot = 0.0
for x in r:
    name = x[ki].split('(')[0][:48]; us = float(x[vi].replace(',', '')) / 1000
    if '--list' in sys.argv: print('%-50s %8.1f  %s' % (name, us, x[gi]))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; tot += us
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print('%-50s x%-3d %8.1f us' % (k, c, t))
print('total %.1f us over %d launches' % (tot, sum(c for c, _ in agg.values())))
