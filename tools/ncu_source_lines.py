"""Per-source-line digest of `ncu -i rep --page source --csv --print-source cuda,sass [--launch-skip K --launch-count 1]`:
executed warp instructions and stall samples of every CUDA source line (needs -lineinfo), top lines first, plus range sums.
   python tools/ncu_source_lines.py src.csv [file-substring] [lo-hi ...]"""
import csv, sys, collections
path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ''
ranges = [tuple(int(v) for v in a.split('-')) for a in sys.argv[3:]]
fname, hdr, ix = None, None, None
lines = {}
tot_i = tot_s = 0.0
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == 'File Path': fname = r[1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No':
        hdr = r; ix = {}
        for i, h in enumerate(hdr): ix.setdefault(h, i)
        continue
    if hdr is None or not r[0].isdigit(): continue
    def f(k):                                       # index from the END: unescaped quotes in source text can add columns
        v = r[ix[k] - len(hdr)]
        try: return float(v)
        except ValueError: return 0.0
    inst, smp = f('Instructions Executed'), f('# Samples')
    st = {h[6:]: f(h) for h in hdr if h.startswith('stall_') and 'Not Issued' not in h}
    tot_i += inst; tot_s += smp
    lines[(fname, int(r[0]))] = (inst, smp, st, r[1])
print('total warp instructions %.4g, samples %.0f' % (tot_i, tot_s))
sel = [(k, v) for k, v in lines.items() if want in k[0]]
for (fn, ln), (inst, smp, st, src) in sorted(sel, key=lambda kv: -kv[1][0])[:40]:
    top = ', '.join('%s %.0f%%' % (a, 100 * b / max(smp, 1)) for a, b in sorted(st.items(), key=lambda t: -t[1])[:3])
    print('%s:%d  inst %5.2f%%  samples %5.2f%%  [%s]  %s' % (fn.split('/')[-1], ln, 100 * inst / tot_i, 100 * smp / max(tot_s, 1), top, src.strip()[:90]))
for lo, hi in ranges:
    i = sum(v[0] for k, v in sel if lo <= k[1] <= hi); s = sum(v[1] for k, v in sel if lo <= k[1] <= hi)
    print('lines %d-%d: inst %.1f%%, samples %.1f%%' % (lo, hi, 100 * i / tot_i, 100 * s / max(tot_s, 1)))
