"""Per-kernel counts of the Blackwell tensor-path SASS mnemonics in the built library (no GPU needed):
UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA loads / stores (cp.async.bulk.tensor), LDTM = tcgen05.ld (TMEM -> registers),
UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, UBLKCP = cp.async.bulk (1-D)."""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'ams_b200', 'lib', 'libams_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
ops = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'UTCBAR', 'UTCATOMSWS', 'UBLKCP', 'SYNCS', 'FFMA2', 'HMMA']
cur, tab = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r'ams::\(anonymous namespace\)::|ams::', '', cur).split('(')[0]
        tab[cur] = collections.Counter()
        continue
    if cur:
        for o in ops:
            if re.search(r'\b' + o + r'\b|\b' + o + r'\.', line):
                tab[cur][o] += 1
print('%s: sm_100a SASS mnemonic counts per kernel (tensor-path kernels only)' % os.path.basename(lib))
print('%-64s' % 'kernel' + ''.join('%11s' % o for o in ops))
tot = collections.Counter()
for k, c in tab.items():
    tot.update(c)
    if c['UTCHMMA'] or c['UTMALDG'] or c['LDTM'] or c['UBLKCP']:
        print('%-64s' % k[:63] + ''.join('%11d' % c[o] for o in ops))
print('%-64s' % 'TOTAL (all kernels)' + ''.join('%11d' % tot[o] for o in ops))
