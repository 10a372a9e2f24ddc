#!/bin/bash
# Executed-instruction evidence for bench.py's roofline (profiles/r02_executed_counts.json): ncu source counters (per SASS line
# executed counts) of the headline transforms, of the dense kernel and of one K4 call, exported as source-page csv.
# usage: bash tools/gpu_ncu_counts.sh <tag>
TAG=${1:-rXX}
mkdir -p gpurun_out
SEC="--section SourceCounters --section LaunchStats --section SpeedOfLight --section WarpStateStats --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis"
python tools/ncu_ntt.py --what dense --reps 2 > gpurun_out/ncu_plain_dense_$TAG.log 2>&1 && \
ncu $SEC --clock-control none --import-source on -k regex:"matvec" -c 2 -o /tmp/dense_$TAG python tools/ncu_ntt.py --what dense --reps 2 > gpurun_out/ncu_dense_$TAG.log 2>&1
ncu -i /tmp/dense_$TAG.ncu-rep --page source --csv > gpurun_out/src_dense_$TAG.csv 2>/dev/null
ncu -i /tmp/dense_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_dense_$TAG.csv 2>/dev/null
python tools/ncu_ntt.py --what k4 --log2 17 --reps 2 > gpurun_out/ncu_plain_k4_$TAG.log 2>&1 && \
ncu $SEC --clock-control none --import-source on -c 2000 -o /tmp/k4_$TAG python tools/ncu_ntt.py --what k4 --log2 17 --reps 2 > gpurun_out/ncu_k4_$TAG.log 2>&1
ncu -i /tmp/k4_$TAG.ncu-rep --page source --csv > gpurun_out/src_k4_$TAG.csv 2>/dev/null
ncu -i /tmp/k4_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_k4_$TAG.csv 2>/dev/null
cat gpurun_out/ncu_plain_k4_$TAG.log
ls -la gpurun_out/*$TAG*; du -sh gpurun_out
