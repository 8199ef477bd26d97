"""Per-region digest of an `ncu --page source --csv --print-source sass` export: executed warp instructions, stall
samples by reason, and the instruction mix, split at the barriers (BAR.SYNC) of the kernel.
   ncu -i rep.ncu-rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > src.csv
   python tools/ncu_source_hot.py src.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
name = rows[0][1]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
regions, cur = [], {'inst': 0, 'samples': 0, 'st': collections.Counter(), 'mix': collections.Counter(), 'n': 0, 'first': None}
def num(x):
    try: return float(x)
    except ValueError: return 0.0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix['Source']]
    op = src.split()[0] if src.split() else '?'
    if op.startswith('@'): op = src.split()[1]
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LDG', 'STG', 'LDS', 'STS', 'BAR')) and '.' in op else '')
    cur['inst'] += num(r[ix['Instructions Executed']]); cur['samples'] += num(r[ix['# Samples']]); cur['n'] += 1
    cur['mix'][op] += num(r[ix['Instructions Executed']])
    for s in stalls: cur['st'][s] += num(r[ix[s]])
    if cur['first'] is None: cur['first'] = r[ix['Address']]
    if src.strip().startswith('BAR') or ' BAR.' in src:
        regions.append(cur); cur = {'inst': 0, 'samples': 0, 'st': collections.Counter(), 'mix': collections.Counter(), 'n': 0, 'first': None}
regions.append(cur)
tot_i = sum(g['inst'] for g in regions); tot_s = sum(g['samples'] for g in regions)
print(name); print('total warp instructions %.4g, samples %d' % (tot_i, tot_s))
for k, g in enumerate(regions):
    if not g['n']: continue
    print('\nregion %d: %d SASS lines, %.1f%% of executed instructions, %.1f%% of samples' % (k, g['n'], 100 * g['inst'] / tot_i, 100 * g['samples'] / max(tot_s, 1)))
    print('  stalls: ' + ', '.join('%s %.0f%%' % (s[6:], 100 * v / max(g['samples'], 1)) for s, v in g['st'].most_common(6)))
    print('  mix   : ' + ', '.join('%s %.1f%%' % (o, 100 * v / max(g['inst'], 1)) for o, v in g['mix'].most_common(14)))
