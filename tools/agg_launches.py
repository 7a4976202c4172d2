"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
    name = re.sub(r'\(.*', '', row['Kernel Name']); name = re.sub(r'^void ', '', name)
    name = re.sub(r'b200::\(anonymous namespace\)::|b200::|<unnamed>::', '', name)
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f"{'total us':>12} {'count':>6} {'avg us':>9} {'share':>6}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:12.1f} {n:6d} {t/n:9.2f} {100*t/tot:5.1f}%  {k}")
print(f"{tot:12.1f} total")
