"""Aggregate an `ncu --page source --csv` export: per-SASS-instruction stall samples, hottest first by address order.
usage: ncu -i X.ncu-rep --page source --csv --launch-skip K --launch-count 1 > src.csv; python profiles/ncu_source_stalls.py src.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
out, tot = [], 0
for n, r in enumerate(rows[2:]):
    try:
        s = int(r[ix['# Samples']])
    except (ValueError, IndexError):
        continue
    tot += s
    out.append((n, s, r))
print('instructions', len(out), 'total samples', tot)
agg = {}
for n, s, r in out:
    for h in stalls:
        agg[h] = agg.get(h, 0) + int(r[ix[h]] or 0)
print('by reason:', sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for n, s, r in out:
    if s >= tot * minpct / 100:
        st = {h: int(r[ix[h]] or 0) for h in stalls}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(n, s, f"{100*s/tot:.1f}%", r[ix['Source']][:80], top)
