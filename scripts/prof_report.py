"""usage: python scripts/prof_report.py <tag> <kernel symbol substring>  (expects gpurun_out/prof_<tag>.ncu-rep and build/obj/pnp_inst_9.o)"""
import csv, os, subprocess, sys
tag, sym = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
os.makedirs('/tmp/probe/sass', exist_ok=True)
subprocess.run('cd /tmp/probe/sass && rm -f *.cubin && cuobjdump -xelf all /root/repo/build/obj/pnp_inst_9.o > /dev/null 2>&1 && nvdisasm -gi -c pnp_inst.sm_100a.cubin > dis_%s.txt 2>/dev/null' % tag, shell=True)
subprocess.run('ncu -i gpurun_out/prof_%s.ncu-rep --page source --csv 2>/dev/null > gpurun_out/src_%s.csv' % (tag, tag), shell=True)
subprocess.run([sys.executable, 'scripts/ncu_lines.py', 'gpurun_out/src_%s.csv' % tag, '/tmp/probe/sass/dis_%s.txt' % tag, sym, '12'])
rows = list(csv.reader(open('gpurun_out/src_%s.csv' % tag)))
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hi]; ia, isrc, ie, ism = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
st = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[hi + 1:] if len(r) > ie and r[ia]]
tot = sum(int(float(r[ism] or 0)) for r in data)
agg = {}
for r in data:
    for i in st:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(float(r[i] or 0))
print('stall totals:', {k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -int(float(r[ism] or 0)))[:top]:
    t2 = sorted(((int(float(r[i] or 0)), hdr[i]) for i in st), reverse=True)[:2]
    print('%5.2f%% exec %9s %-52s %s' % (100 * int(float(r[ism])) / tot, r[ie], r[isrc][:52], t2))
