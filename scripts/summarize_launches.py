"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split('(')[0][:90]
    t = float(r[14].replace(',', ''))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print('%-92s %6s %14s %7s' % ('kernel', 'calls', 'time_ns', 'share'))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-92s %6d %14.0f %6.2f%%' % (k, n, t, 100 * t / tot))
print('total %.3f ms over %d launches' % (tot / 1e6, len(rows)))
