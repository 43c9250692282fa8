#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vae_launches.csv python tools/vae_prof.py 2 > gpurun_out/vae_prof.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/vae_launches.csv
