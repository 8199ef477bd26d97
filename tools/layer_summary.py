"""Summarise a tools/profile_layers.py table: time per kernel group x resolution stage."""
import collections, re, sys
def load(path):
    mode = None; rows = []
    for l in open(path).read().splitlines():
        if l.startswith('=='): mode = l.split()[1].rstrip(':'); continue
        if l.startswith('group') or not l.strip(): continue
        p = l.split()
        rows.append((mode, p[0], p[1], int(p[2]), float(p[3]), float(p[4]), float(p[5])))
    return rows
def stage(layer):
    m = re.search(r'expanded_conv(_(\d+))?/(\w+)', layer)
    if layer.endswith('/Conv'): return 'A:257'
    if not m: return 'E:head'
    b = int(m.group(2) or 0); k = m.group(3)
    if b == 0: return 'A:257'
    if b == 1: return 'B:129' if k == 'project' else 'A:257'
    if b == 2: return 'B:129'
    if b == 3: return 'C:65' if k == 'project' else 'B:129'
    if b in (4, 5): return 'C:65'
    if b == 6: return 'D:33' if k == 'project' else 'C:65'
    return 'D:33'
for path in sys.argv[1:]:
    rows = load(path)
    for mode in ('train', 'infer'):
        tab = collections.defaultdict(lambda: collections.defaultdict(float))
        for r in rows:
            if r[0] == mode: tab[r[1]][stage(r[2])] += r[4]
        st = sorted({s for g in tab.values() for s in g})
        print(path, mode); print('%-16s' % 'group' + ''.join('%9s' % s for s in st) + '%9s' % 'total')
        tot = collections.defaultdict(float)
        for g, v in sorted(tab.items(), key=lambda kv: -sum(kv[1].values())):
            print('%-16s' % g + ''.join('%9.0f' % v.get(s, 0) for s in st) + '%9.0f' % sum(v.values()))
            for s in st: tot[s] += v.get(s, 0)
        print('%-16s' % 'TOTAL' + ''.join('%9.0f' % tot[s] for s in st) + '%9.0f' % sum(tot.values()))
