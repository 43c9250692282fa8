#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x > gpurun_out/ops.log 2>&1; echo "ops rc=$?"; tail -2 gpurun_out/ops.log
timeout 300 python tools/op_bench.py attn
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q -s --tb=short -k "sd15 or h_edit_step or mos2" > gpurun_out/loop.log 2>&1; echo "loop rc=$?"; grep -E "rel |passed|failed|Error" gpurun_out/loop.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:self_attn2 -s 2 -c 1 -o gpurun_out/prof_attn2 python tools/op_bench.py attn --iters 1 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 900 python bench.py --steps 1 --warmup 2 --profile --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_quick.err
