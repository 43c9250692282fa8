#!/usr/bin/env python
"""Validates bench.py's CPU-arm extrapolation (one timestep of one image x T) against a FULL run of the oracle port of the reference
loop: BASELINE.json configs[0] (SD-1.5 geometry, 1 image, T = 10, implicit h-Edit-R + P2P) end to end on this host's cores.
    python tests/cpu_extrapolation_check.py  ->  profiles/r02_cpu_extrapolation.json
Lives under tests/ because it executes the oracle (test infrastructure)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

T = 10
threads = os.cpu_count() or 1
bench.cpu_reference_sample(T, 1, threads)                       # warm-up (weights, thread pools)
per_ts = [bench.cpu_reference_sample(T, 1, threads) for _ in range(3)]
t0 = time.perf_counter()
full_per_ts = bench.cpu_reference_sample(T, T, threads)         # all T timesteps of the same edit
full = time.perf_counter() - t0
est = sum(per_ts) / len(per_ts) * T
out = {"T": T, "threads": threads, "one_timestep_s": per_ts, "extrapolated_loop_s": est, "full_loop_s": full_per_ts * T,
       "full_wall_s_incl_model_build": full, "extrapolated_over_full": est / (full_per_ts * T),
       "note": "bench.py's cpu arm times 1 of T timesteps (9 UNet sample-forwards) and multiplies by T; the loop has no per-step state that "
               "changes the work (same 9 sample-forwards every step), so the ratio should be ~1"}
print(json.dumps(out))
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_cpu_extrapolation.json"), "w"), indent=1)
