"""Join an `ncu --page source --csv` SASS dump with nvdisasm line info and rank CUDA source
lines by executed instructions / stall samples.
    python scripts/ncu_lines.py <sass.csv> <nvdisasm -gi -c output> <kernel symbol substring> [top]
"""
import csv, re, sys, collections

def main():
    sass_csv, dis, sym = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # ---- nvdisasm: offset -> (file,line)
    lines = open(dis).read().split('\n')
    infn = False; cur = ('?', 0); off2line = {}
    for ln in lines:
        if ln.startswith('//---') and '.text.' in ln:
            infn = sym in ln
            continue
        if not infn:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
        if m:
            off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
    rows = list(csv.reader(open(sass_csv)))
    hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
    hdr = rows[hi]
    ia, ie, ism = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    data = [r for r in rows[hi + 1:] if len(r) > ie and r[ia]]
    base = int(data[0][ia], 16) if data[0][ia].startswith('0x') else int(data[0][ia])
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    opc = collections.Counter()
    for r in data:
        a = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
        key, txt = off2line.get(a - base, (('?', 0), ''))
        n = int(float(r[ie] or 0)); s = int(float(r[ism] or 0))
        agg[key][0] += n; agg[key][1] += s; agg[key][2] += 1
        tot_i += n; tot_s += s
        op = txt.split()[0] if txt else '?'
        if op.startswith('@'):
            op = txt.split()[1] if len(txt.split()) > 1 else op
        opc[op.split('.')[0]] += n
    print('total warp instructions %d, samples %d, static SASS %d' % (tot_i, tot_s, len(data)))
    print('--- by opcode'); 
    for op, n in opc.most_common(18):
        print('  %-10s %5.1f%%' % (op, 100.0 * n / tot_i))
    print('--- by source line (inst%, stall-sample%, static count)')
    for key, (n, s, c) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('  %-18s:%4d  %5.1f%%  %5.1f%%  %4d' % (key[0], key[1], 100.0 * n / tot_i, 100.0 * s / max(tot_s, 1), c))

if __name__ == '__main__':
    main()
