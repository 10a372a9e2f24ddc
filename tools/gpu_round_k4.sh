#!/bin/bash
# K4 (robust decoder) measurements of one GPU visit: configs 2/3r/4/5, stage timings of the staged decoder, ncu captures.
# usage: bash tools/gpu_round_k4.sh <tag>
TAG=${1:-rXX}
mkdir -p gpurun_out
python tools/bench_configs.py --which c2,c3r,c4,c5 --log2-c4 20 > gpurun_out/configs_$TAG.json 2> gpurun_out/configs_$TAG.err; echo "configs rc=$?"
HBMPC_STAGED_PROF=1 python tools/bench_configs.py --which c4 --log2-c4 17 2>&1 | grep staged_decode > gpurun_out/k4_stages_$TAG.log
HBMPC_NO_SPECULATION=1 HBMPC_STAGED_PROF=1 python tools/bench_configs.py --which c3r --log2 20 > gpurun_out/c3r_nospec_$TAG.json 2>> gpurun_out/k4_stages_$TAG.log
K4="python tools/bench_configs.py --which c4 --log2-c4 17"
$K4 > gpurun_out/plain_k4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_k4_$TAG.csv $K4 > gpurun_out/ncu_list_k4.log 2>&1
$K4 > gpurun_out/plain_k4b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"bm_segment|omega_kernel|staged_|permute" -s 60 -c 16 -o gpurun_out/prof_k4_$TAG $K4 > gpurun_out/ncu_full_k4.log 2>&1
tail -2 gpurun_out/ncu_full_k4.log
cat gpurun_out/configs_$TAG.json | head -80
