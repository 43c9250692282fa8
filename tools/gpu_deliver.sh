#!/bin/bash
# Round deliverables: full GPU test-suite, smoke, default bench (with cpu baseline), reference arm, ncu launch list, ncu full captures.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python bench.py --profile > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
if [ -n "$HEDIT_REF" ]; then timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --batch 8 --timesteps 4 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/prof_conv python tools/op_bench.py conv --iters 1 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:self_attn2 -s 2 -c 1 -o gpurun_out/prof_attn2 python tools/op_bench.py attn --iters 1 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out | tail -20
