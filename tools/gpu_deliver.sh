#!/bin/bash
# Round deliverables: GPU tests, smoke, default bench (+ per-op profile), ncu launch list, ncu --set full of the conv and attention kernels
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1500 python bench.py --profile > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --batch 8 --timesteps 4 --no-cpu-baseline --no-single-image > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
for k in conv:gemm_bf16 attn:self_attn4; do
  what=${k%%:*}; rx=${k##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o gpurun_out/prof_$what -f python tools/op_bench.py $what --iters 1 > gpurun_out/ncu_$what.log 2>&1; echo "ncu $what rc=$?"
  ncu -i gpurun_out/prof_$what.ncu-rep --page raw --csv > gpurun_out/prof_${what}_raw.csv 2>/dev/null
done
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/forward_dram.csv python tools/forward_dram.py > gpurun_out/forward_dram.log 2>&1; echo "ncu forward dram rc=$?"
python tools/forward_dram.py --summarise gpurun_out/forward_dram.csv > gpurun_out/forward_dram_summary.txt 2>&1; head -3 gpurun_out/forward_dram_summary.txt | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 0 -c 1 -o gpurun_out/prof_qkv -f python tools/op_bench.py linear --iters 1 > gpurun_out/ncu_qkv.log 2>&1; echo "ncu qkv rc=$?"
ncu -i gpurun_out/prof_qkv.ncu-rep --page raw --csv > gpurun_out/prof_qkv_raw.csv 2>/dev/null
python tools/show_bench.py gpurun_out/bench_default.json 2>/dev/null | head -30 | cut -c1-500
