"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
tot, cnt = collections.Counter(), collections.Counter()
for x in csv.DictReader(lines):
    if x['Metric Name'] != 'gpu__time_duration.sum':
        continue
    k = x['Kernel Name'].split('(')[0].replace('unnamed>::', '').replace('void ', '')[:64]
    v = float(x['Metric Value'].replace(',', ''))
    v = v / 1000.0 if x['Metric Unit'] == 'ns' else v
    tot[k] += v; cnt[k] += 1
s = sum(tot.values())
print('%s: %d launches, %.1f us total (cold-cache, serialised: compare shares, not absolutes)' % (path, sum(cnt.values()), s))
print('%-66s %6s %10s %7s' % ('kernel', 'n', 'us', 'share'))
for k, v in tot.most_common():
    print('%-66s %6d %10.1f %6.1f%%' % (k, cnt[k], v, 100 * v / s))
