"""Attribute an `ncu --page source --csv --print-source sass` dump to CUDA source functions / lines.
    python scripts/ncu_funcs.py <sass.csv> <nvdisasm -g -c output> <kernel section substring> [func-to-list]
"""
import csv, re, sys, collections, os

def fun_map(root, fname):
    p = os.path.join(root, fname)
    if not os.path.exists(p): return []
    st = []
    for i, s in enumerate(open(p), 1):
        if ('__device__' in s or '__global__' in s):
            m = re.search(r'\b(\w+)\s*\(', s.split('__device__')[-1].split('__global__')[-1])
            if m: st.append((i, m.group(1)))
    return st

def main():
    sass_csv, dis, sym = sys.argv[1:4]
    listf = sys.argv[4] if len(sys.argv) > 4 else None
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'catint_b200', 'csrc')
    on = False; cur = ('?', 0); off = {}
    for ln in open(dis):
        if ln.startswith('.text.'):
            on = sym in ln; continue
        if not on: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m: off[int(m.group(1), 16)] = cur
    maps = {}
    def func(key):
        f, l = key
        if f not in maps: maps[f] = fun_map(root, f)
        name = '?'
        for st, nm in maps[f]:
            if st <= l: name = nm
        return f.split('.')[0][-8:] + ':' + name
    rows = list(csv.reader(open(sass_csv)))
    hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
    hdr = rows[hi]
    ia, isrc, ie, ism = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    ist = hdr.index('Warp Stall Sampling (All Samples)')
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_')]
    data = [r for r in rows[hi + 1:] if len(r) > ie and r[ia].startswith('0x')]
    base = int(data[0][ia], 16)
    byf = collections.defaultdict(lambda: [0, 0, 0]); byl = collections.defaultdict(lambda: [0, 0, 0])
    stf = collections.defaultdict(collections.Counter)
    ti = ts = 0
    listing = []
    for r in data:
        a = int(r[ia], 16) - base
        key = off.get(a, ('?', 0))
        n = int(float(r[ie] or 0)); s = int(float(r[ism] or 0))
        f = func(key)
        byf[f][0] += n; byf[f][1] += s; byf[f][2] += 1
        byl[key][0] += n; byl[key][1] += s; byl[key][2] += 1
        for i, h in stall_cols:
            v = int(float(r[i] or 0))
            if v: stf[f][h] += v
        ti += n; ts += s
        if listf and listf in f:
            top = sorted(((int(float(r[i] or 0)), h[6:]) for i, h in stall_cols), reverse=True)[:2]
            listing.append('%6x %5d %9d %4d  %-60s %s' % (a, key[1], n, s, r[isrc].strip()[:60], [t for t in top if t[0]]))
    print('total warp instr %d samples %d static %d' % (ti, ts, len(data)))
    print('--- by function: inst%  samples%  static  cycles/instr-ish  top stalls')
    for f, (n, s, c) in sorted(byf.items(), key=lambda kv: -kv[1][1])[:24]:
        tops = ', '.join('%s %.0f%%' % (h[6:], 100.0 * v / max(s, 1)) for h, v in stf[f].most_common(4))
        print('  %-32s %5.1f%% %5.1f%% %5d   %s' % (f, 100.0 * n / ti, 100.0 * s / ts, c, tops))
    print('--- by line')
    for key, (n, s, c) in sorted(byl.items(), key=lambda kv: -kv[1][1])[:30]:
        print('  %-18s:%4d  %5.1f%%  %5.1f%%  %4d' % (key[0], key[1], 100.0 * n / ti, 100.0 * s / ts, c))
    if listing:
        print('--- listing of', listf); print('\n'.join(listing))

main()
