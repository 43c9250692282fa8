#!/bin/bash
mkdir -p gpurun_out
HEDIT_NET_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t_face_launches.csv python tools/face_prof.py > /dev/null 2>&1; echo "ncu list rc=$?"
HEDIT_NET_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 60 -c 3 -o gpurun_out/prof_face_gemm -f python tools/face_prof.py > gpurun_out/ncu_face_gemm.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_face_gemm.ncu-rep --page raw --csv > gpurun_out/prof_face_gemm_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_face_gemm_raw.csv 2>&1 | head -60
