#!/bin/bash
# Round deliverables: full GPU test-suite, default bench (with cpu baseline), reference arm, ncu launch list, ncu full captures.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python bench.py --profile > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --batch 8 --timesteps 4 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 60 -c 3 -o gpurun_out/prof_unet_gemm python bench.py --steps 1 --warmup 1 --batch 8 --timesteps 2 --no-cpu-baseline > gpurun_out/ncu_gemm_full.log 2>&1; echo "ncu gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:self_attn2 -s 4 -c 2 -o gpurun_out/prof_unet_attn python bench.py --steps 1 --warmup 1 --batch 8 --timesteps 2 --no-cpu-baseline > gpurun_out/ncu_attn_full.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out | tail -20
