"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/launch_table.py launches.csv [first_id last_id]"""
import csv
import sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iid = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('ID')
lo = int(sys.argv[2]) if len(sys.argv) > 2 else -1
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
tot = OrderedDict()
for r in rows[1:]:
    try:
        i = int(r[iid]); v = float(r[iv].replace(',', ''))
    except ValueError:
        continue
    if not (lo <= i <= hi):
        continue
    unit = r[hdr.index('Metric Unit')]
    us = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    name = r[ik].split('(')[0][-60:]
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1; t[1] += us
total = sum(t[1] for t in tot.values())
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('%-62s %5d launches %10.1f us  %5.1f %%' % (name, n, us, 100 * us / total))
print('%-62s %5s          %10.1f us' % ('total', '', total))
