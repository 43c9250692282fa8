#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/e2e_image_bench.py > gpurun_out/e2e_image.json 2> gpurun_out/e2e_image.err; echo "rc=$?"; tail -3 gpurun_out/e2e_image.err; cat gpurun_out/e2e_image.json
