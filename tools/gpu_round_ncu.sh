#!/bin/bash
# ncu evidence of one GPU visit, sized to fit the 64 MiB return limit: launch lists (csv) and `--set full` captures exported to
# raw csv on the box (the .ncu-rep files stay on the box unless they are small).
# usage: bash tools/gpu_round_ncu.sh <tag>
TAG=${1:-rXX}
mkdir -p gpurun_out
SHORT="python bench.py --steps 2 --warmup 3 --log2-batch 20 --log2-e2e-batch 16 --no-cpu-baseline --no-robust-leg"
$SHORT > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $SHORT > gpurun_out/ncu_list.log 2>&1
$SHORT > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"matvec|ntt" -s 4 -c 3 -o /tmp/prof_$TAG $SHORT > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>/dev/null
K4="python tools/bench_configs.py --which c4 --log2-c4 17"
$K4 > gpurun_out/plain_k4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_k4_$TAG.csv $K4 > gpurun_out/ncu_list_k4.log 2>&1
$K4 > gpurun_out/plain_k4b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"bm_segment|omega_kernel|staged_|permute|ntt_kernel" -s 100 -c 24 -o /tmp/prof_k4_$TAG $K4 > gpurun_out/ncu_full_k4.log 2>&1
ncu -i /tmp/prof_k4_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_k4_$TAG.csv 2>/dev/null
HBMPC_STAGED_PROF=1 python tools/bench_configs.py --which c4 --log2-c4 17 2>&1 | grep staged_decode > gpurun_out/k4_stages_$TAG.log
HBMPC_NO_SPECULATION=1 HBMPC_STAGED_PROF=1 python tools/bench_configs.py --which c3r --log2 20 > gpurun_out/c3r_nospec_$TAG.json 2>> gpurun_out/k4_stages_$TAG.log
ls -la /tmp/*.ncu-rep gpurun_out; du -sh gpurun_out
