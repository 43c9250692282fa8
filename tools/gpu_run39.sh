#!/bin/bash
# Round deliverables v6: smoke, default bench (with cpu baseline), reference arm, ncu launch list of the default bench
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python bench.py --profile > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cat gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --batch 8 --timesteps 4 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
python tools/show_bench.py gpurun_out/bench_default.json
