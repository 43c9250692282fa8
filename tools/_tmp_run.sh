mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_ops.py -m gpu -q --tb=short -s 2>&1 | grep -E "tiny unet forward|sd15_config1|passed|failed|Error" | tail -6
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile > gpurun_out/bench_up.json 2>/dev/null; python tools/show_bench.py gpurun_out/bench_up.json 2>/dev/null | head -12
