#!/bin/bash
# Runs each GPU parity test group in its own process (a device-side trap poisons the CUDA context) and collects logs.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
: > gpurun_out/summary.txt
run() { # name timeout cmd...
  local name=$1 to=$2; shift 2
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  echo "$name rc=$?" >> gpurun_out/summary.txt
}
for t in test_linear test_conv3x3 test_group_norm test_layer_norm test_self_attention test_self_attention_injection test_cross_attention_p2p; do
  run "ops_$t" 400 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "$t" -s --tb=short
done
run unet_tiny 600 python -m pytest tests/test_gpu_unet.py -m gpu -q -s --tb=short -k "forward_tiny"
for t in tiny_refine_blend tiny_reference_schedule tiny_replace_mos2 tiny_noblend sd15_config1; do
  run "loop_$t" 900 python -m pytest tests/test_gpu_unet.py -m gpu -q -s --tb=short -k "$t"
done
cat gpurun_out/summary.txt
grep -h -E "passed|failed|error" gpurun_out/*.log | tail -40
if [ -n "$HEDIT_BENCH" ]; then
  timeout 900 python bench.py --steps 1 --warmup 1 --profile --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
  echo "bench rc=$?" >> gpurun_out/summary.txt
  tail -c 3000 gpurun_out/bench_quick.json
  tail -5 gpurun_out/bench_quick.err
fi
