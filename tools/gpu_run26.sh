#!/bin/bash
# Re-entry check: GPU suite, per-operator micro-benchmarks, default bench, ncu --set full captures of the top kernels.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/op_bench.py all --iters 10 > gpurun_out/op_bench.log 2>&1; echo "op_bench rc=$?"; cat gpurun_out/op_bench.log
timeout 900 python bench.py --profile > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
for k in conv:gemm_bf16 attn:self_attn2 linear:gemm_bf16 cross:cross_attn2; do
  what=${k%%:*}; rx=${k##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o gpurun_out/prof_$what -f python tools/op_bench.py $what --iters 1 > gpurun_out/ncu_$what.log 2>&1; echo "ncu $what rc=$?"
  ncu -i gpurun_out/prof_$what.ncu-rep --page raw --csv > gpurun_out/prof_${what}_raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -30
