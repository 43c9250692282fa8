#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --batch 1 --steps 2 --warmup 3 --no-cpu-baseline --no-single-image > gpurun_out/r2p_bench_b1.json 2> gpurun_out/r2p_bench_b1.err; echo "bench b1 rc=$?"; python tools/show_bench.py gpurun_out/r2p_bench_b1.json | head -1 | cut -c1-200
HEDIT_LOOP_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_b1_launches.csv python bench.py --steps 1 --warmup 1 --batch 1 --timesteps 3 --no-cpu-baseline --no-single-image > gpurun_out/r2p_ncu.log 2>&1; echo "ncu rc=$?"
python tools/shape_prof.py --samples 2 > gpurun_out/r2p_shape_prof_s2.log 2>&1; head -3 gpurun_out/r2p_shape_prof_s2.log
python tools/shape_prof.py --samples 5 > gpurun_out/r2p_shape_prof_s5.log 2>&1; head -3 gpurun_out/r2p_shape_prof_s5.log
