#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/face_launches.csv python tools/face_prof.py > /dev/null 2>&1; echo rc=$?
