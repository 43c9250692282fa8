#!/bin/bash
mkdir -p gpurun_out
for b in 1 2 4 8; do
  timeout 600 python bench.py --batch $b --steps 1 --warmup 1 --no-cpu-baseline --profile > gpurun_out/bench_b$b.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b$b.json"))
tot=sum(v["ms"] for v in d["kernel_breakdown_ms_per_forward"].values())
print("batch $b: %.3f img/s  frac %.3f  forward(5B samples) %.2f ms -> %.3f ms/sample" % (d["value"], d["roofline"]["frac"], tot, tot/(5*$b)), {k: round(v["ms"]/(5*$b),4) for k,v in list(d["kernel_breakdown_ms_per_forward"].items())[:8]})
PY
done
