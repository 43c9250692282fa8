"""Target of the whole-forward DRAM-traffic capture: ONE 40-sample SD-1.5 UNet forward after one warm-up forward.
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file X python tools/forward_dram.py
  python tools/forward_dram.py --summarise X   -> total bytes of the second forward against SURVEY 8d's 1.08 GB per sample-forward"""
import collections, csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
S = 40
if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
    rows = list(csv.reader(open(sys.argv[2], errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[hdr]; ki, ni, ui, vi, idi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Unit"), H.index("Metric Value"), H.index("ID")
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}
    per = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        d = per.setdefault(int(r[idi]), {"name": r[ki].split("(")[0].replace("void ", "").replace("hedit::", "")[:48]})
        d[r[ni]] = float(r[vi].replace(",", "")) * mult.get(r[ui], 1)
    ids = list(per)
    half = ids[len(ids) // 2:]           # the second forward (weights load + warm-up forward come first; both forwards launch the same kernels)
    fw = [per[i] for i in ids if "gpu__time_duration.sum" in per[i]]
    # the forward's kernel count = launches after the last weight-conversion kernel, halved
    last_setup = max((n for n, i in enumerate(ids) if "cvt" in per[i]["name"] or "cast_w" in per[i]["name"]), default=-1)
    body = ids[last_setup + 1:]
    second = body[len(body) // 2:]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i in second:
        d = per[i]; a = agg[d["name"]]
        a[0] += 1; a[1] += d.get("dram__bytes_read.sum", 0); a[2] += d.get("dram__bytes_write.sum", 0); a[3] += d.get("gpu__time_duration.sum", 0)
    rd, wr, t = sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values()), sum(a[3] for a in agg.values())
    print(f"one forward of {S} samples: {len(second)} kernels, DRAM read {rd / 1e9:.2f} GB + write {wr / 1e9:.2f} GB = {(rd + wr) / 1e9 / S:.3f} GB per sample-forward "
          f"(SURVEY 8d minimum: 1.08 GB incl. 1.72 GB of weights once per launch = {1.72 / S:.3f} GB per sample at {S}); kernel time under ncu {t / 1e6:.2f} ms")
    for k, a in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2]))[:14]:
        print(f"  {a[0]:4d}x  read {a[1] / 1e9:7.3f} GB  write {a[2] / 1e9:7.3f} GB  {a[3] / 1e6:7.3f} ms  {k}")
    sys.exit(0)
import torch
import hedit_b200
cfg = dict(in_channels=4, out_channels=4, sample_size=64, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, heads=8,
           cross_attention_dim=768, norm_groups=32, ctx_len=77)
eng = hedit_b200.UNetEngine(cfg, max_samples=S, max_contexts=S)
eng.load_random_weights(0)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(S, 4, 64, 64, generator=g, device="cuda"); ctx = torch.randn(S, 77, 768, generator=g, device="cuda")
for _ in range(2):
    eng.forward(x, 500.0, ctx)
    torch.cuda.synchronize()
