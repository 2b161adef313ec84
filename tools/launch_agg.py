#!/usr/bin/env python3
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv): total / count / mean per kernel."""
import csv, collections, sys
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if len(r) > 5 and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d['Metric Value'].replace(',', ''))
            except ValueError:
                continue
            u = d['Metric Unit']
            v = v / 1e3 if u == 'us' else v / 1e6 if u == 'ns' else v * 1e3 if u == 's' else v
            k = d['Kernel Name'][:70]
            agg[k][0] += 1
            agg[k][1] += v
    print("==", f)
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:10.3f} ms {n:5d}  {t / n * 1e3:9.1f} us  {k}")
