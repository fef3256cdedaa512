"""Code-size breakdown of one kernel: SASS instructions per source function / line (needs -lineinfo).
    python scripts/sass_lines.py <nvdisasm -g -c output> <kernel substring>
"""
import re, sys, collections

def main():
    path, want = sys.argv[1], sys.argv[2]
    cnt = collections.Counter(); inl = collections.Counter(); cur = None; on = False; ctx = None
    for l in open(path):
        if l.startswith('.text.'):
            on = want in l; continue
        if not on: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            ctx = (m.group(3).split('/')[-1], int(m.group(4))) if m.group(3) else None
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l) and cur:
            cnt[cur] += 1
            if ctx: inl[(cur[0], ctx)] += 1
    tot = sum(cnt.values()); print('total instr', tot, '%.1f KB' % (tot * 16 / 1024))
    byfile = collections.Counter()
    for (f, ln), c in cnt.items(): byfile[f] += c
    print(dict(byfile))
    # by function: map line -> function via a crude scan of the sources
    import os
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'catint_b200', 'csrc')
    byfun = collections.Counter()
    for f in byfile:
        p = os.path.join(root, f)
        if not os.path.exists(p): continue
        starts = []
        for i, s in enumerate(open(p), 1):
            m = re.match(r'(?:__device__|__global__|template|static|inline).*?\b(\w+)\s*\(', s)
            if m and ('__device__' in s or '__global__' in s): starts.append((i, m.group(1)))
        for (ff, ln), c in cnt.items():
            if ff != f: continue
            name = '?'
            for st, nm in starts:
                if st <= ln: name = nm
            byfun[(f, name)] += c
    for k, c in byfun.most_common(25): print('%7d %5.1f%%  %s:%s' % (c, 100.0 * c / tot, k[0], k[1]))
    print('-- top lines')
    for (f, ln), c in cnt.most_common(25): print('%7d %s:%d' % (c, f, ln))

main()
