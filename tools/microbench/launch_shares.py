"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        h = r; start = i; break
ki = h.index('Kernel Name'); mi = h.index('Metric Value'); ui = h.index('Metric Unit')
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[start + 1:]:
    if len(r) <= mi: continue
    t = float(r[mi].replace(',', ''))
    if r[ui] == 'ns': t /= 1e3
    elif r[ui] == 'ms': t *= 1e3
    name = re.sub(r'\(.*$', '', r[ki]).replace('thb::', '').replace('<unnamed>::', '').replace('(anonymous namespace)::', '')
    name = re.sub(r'\(int\)', '', name)
    tot[name] += t; cnt[name] += 1
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-60s n=%5d total=%10.1f us avg=%8.2f us share=%5.1f%%" % (k[:60], cnt[k], v, v / cnt[k], 100 * v / s))
print("sum %.1f us" % s)
