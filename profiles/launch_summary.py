"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / total (us)."""
import csv, collections, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki][:70]; v = float(r[vi].replace(',', '')); u = r[ui]
        v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
        agg.setdefault(name, []).append(v)
    print(f)
    tot = sum(sum(v) for k, v in agg.items() if k.startswith(('k_', 'void k_')))
    for k, v in agg.items():
        if k.startswith(('k_', 'void k_')):
            print('  %-70s n=%3d mean=%9.1f us total=%10.1f us share=%5.1f%%' % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
