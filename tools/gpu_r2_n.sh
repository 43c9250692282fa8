#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_compat.py -q -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log | cut -c1-300
for v in 1 0; do
  HEDIT_PDL=$v timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_pdl$v.json 2> gpurun_out/r2n_bench_pdl$v.err; echo "bench pdl=$v rc=$?"
  python tools/show_bench.py gpurun_out/r2n_bench_pdl$v.json 2>/dev/null | head -1 | cut -c1-120; python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_pdl$v.json')); print('single_image', d['single_image']['value'], 'e2e', d['e2e']['value'])"
  tail -2 gpurun_out/r2n_bench_pdl$v.err
done
