"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel count, time, share."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
H = rows[0]
ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split('(')[0]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(',', ''))
tot = sum(a[1] for a in agg.values())
print('# %s: %d launches, %.3f ms total (cold-cache, serialised under ncu)' % (sys.argv[1], len(rows) - 1, tot / 1e6))
for n, a in agg.items():
    print('%-28s n=%4d total_ms=%9.3f share=%.3f avg_us=%9.1f' % (n, a[0], a[1] / 1e6, a[1] / tot, a[1] / a[0] / 1e3))
