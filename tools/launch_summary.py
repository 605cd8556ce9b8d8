"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): time and launches per kernel."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
iname, ival, iunit = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if r[hdr.index('Metric Name')] != 'gpu__time_duration.sum':
        continue
    v = float(r[ival].replace(',', ''))
    u = r[iunit]
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}.get(u, 1e-6)
    n = r[iname].split('(')[0]
    tot[n] += v
    cnt[n] += 1
s = sum(tot.values())
print('%-70s %8s %10s %8s %7s' % ('kernel', 'launches', 'total ms', 'ms/launch', 'share'))
for n, v in tot.most_common():
    print('%-70s %8d %10.3f %8.3f %6.1f%%' % (n[:70], cnt[n], v, v / cnt[n], 100 * v / s))
print('%-70s %8d %10.3f' % ('all', sum(cnt.values()), s))
