#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/style_bench.py --batch 8 --K 3 > gpurun_out/style_bench.json 2> gpurun_out/style_bench.err; echo "rc=$?"; tail -3 gpurun_out/style_bench.err; cat gpurun_out/style_bench.json
