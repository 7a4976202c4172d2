"""Top stalled SASS instructions from `ncu -i X.ncu-rep --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    try: data.append((int(r[ci['# Samples']]), r))
    except Exception: pass
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for n, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[ci[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    idx = rows.index(r) - hi
    print(f"{n:5d} {100*n/tot:5.1f}%  #{idx:5d} exec={r[ci['Instructions Executed']]:>7s} {r[ci['Source']].strip()[:70]:70s} {st}")
