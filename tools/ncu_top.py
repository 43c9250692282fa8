"""Print the hottest SASS lines (warp-stall samples) of an `ncu --page source --csv` export."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # which kernel section of the export
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = heads[which]
stop = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
print(rows[hi - 1][:2])
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:stop] if len(r) == len(hdr)]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
print('stall totals:', sorted(((v, k) for k, v in agg.items()), reverse=True)[:8])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for k, r in enumerate(data):
    r.append(k)
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:n]:
    s = int(r[ix['# Samples']])
    st = sorted(((int(r[ix[h]]), h.replace('stall_', '')) for h in stalls), reverse=True)[:2]
    print(f"{r[-1]:5d} {s:6d} {100 * s / tot:5.1f}%  {r[ix['Source']].strip()[:64]:64s} {st}")
