"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: total time, launches, share."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; ki, mi = H.index("Kernel Name"), H.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # launches to skip at the start (set-up)
agg, n = collections.defaultdict(lambda: [0, 0.0]), 0
for r in rows[hdr + 1:]:
    if len(r) <= mi:
        continue
    try:
        v = float(r[mi].replace(",", ""))
    except ValueError:
        continue
    n += 1
    if n <= skip:
        continue
    k = r[ki].split("(")[0].replace("void ", "").replace("hedit::", "")[:70]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t / 1e6:9.3f} ms {c:5d}x {100 * t / tot:5.1f}%  avg {t / c / 1e3:7.1f} us  {k}")
