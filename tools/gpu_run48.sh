#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -q --tb=short -k "masactrl_explicit" 2>&1 | tail -3
timeout 600 python tools/config_bench.py --config 3 > gpurun_out/config3.json 2> gpurun_out/config3.err; tail -2 gpurun_out/config3.err; cat gpurun_out/config3.json
timeout 900 python tools/config_bench.py --config 5 > gpurun_out/config5.json 2> gpurun_out/config5.err; tail -2 gpurun_out/config5.err; cat gpurun_out/config5.json
