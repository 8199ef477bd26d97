"""Per-kernel shares of an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file ...)."""
import csv, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors='replace')) if r and r[0].isdigit()]
hdr = None
for r in csv.reader(open(path, errors='replace')):
    if r and r[0] == 'ID':
        hdr = r
        break
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows:
    name = r[ik].split('(')[0].split('::')[-1]
    v = float(r[iv].replace(',', ''))
    us = v / 1e3 if r[iu] in ('ns', 'nsecond') else (v if r[iu] in ('us', 'usecond') else v * 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print('%s: %d launches, %.1f us total (cold-cache, serialised: compare shares, not absolutes)' % (path, sum(a[0] for a in agg.values()), tot))
print('%-66s %5s %10s %7s' % ('kernel', 'n', 'us', 'share'))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-66s %5d %10.1f %6.1f%%' % (k, n, us, 100 * us / tot))
