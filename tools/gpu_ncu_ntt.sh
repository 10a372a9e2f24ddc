#!/bin/bash
# ncu --set full capture of the headline transforms (one K1 and one K3 launch at 2^20 items) with raw and source pages exported
# to csv on the box (the .ncu-rep stays there).  usage: bash tools/gpu_ncu_ntt.sh <tag> [extra args of tools/ncu_ntt.py]
TAG=${1:-rXX}; shift
mkdir -p gpurun_out
python tools/ncu_ntt.py "$@" > gpurun_out/ncu_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"ntt|matvec" -s 3 -c 3 -o /tmp/prof_$TAG python tools/ncu_ntt.py "$@" > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > gpurun_out/src_$TAG.csv 2>/dev/null
ls -la gpurun_out/*$TAG*
