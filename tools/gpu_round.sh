#!/bin/bash
# One GPU visit: smoke, default bench, reference arm, launch list of the bench command.  Executed-instruction counts and the
# --set full captures come from tools/gpu_ncu_counts.sh.  Everything written to gpurun_out/ is small (gpurun returns <= 64 MiB).
# usage: bash tools/gpu_round.sh <tag>
TAG=${1:-rXX}
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_$TAG.json; cat gpurun_out/bench_ref_$TAG.json
SHORT="python bench.py --steps 2 --warmup 3 --log2-batch 20 --log2-e2e-batch 16 --no-cpu-baseline --no-configs"
$SHORT > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $SHORT > gpurun_out/ncu_list_$TAG.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
