#!/bin/bash
# One GPU visit: smoke, default bench, reference arm, launch list, ncu full captures of the two hot kernels.
# usage: bash tools/gpu_round.sh <tag>
TAG=${1:-rXX}
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_$TAG.json; cat gpurun_out/bench_ref_$TAG.json
SHORT="python bench.py --steps 2 --warmup 3 --log2-batch 20 --log2-e2e-batch 16 --no-cpu-baseline --no-robust-leg"
$SHORT > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $SHORT > gpurun_out/ncu_list.log 2>&1
$SHORT > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"matvec|ntt" -s 4 -c 4 -o gpurun_out/prof_$TAG $SHORT > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
