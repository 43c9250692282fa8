#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2r_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2r_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2r_bench.json').read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'single_image', d['single_image']['value'], 'frac', d['roofline']['frac'])"
