#!/bin/bash
# One GPU visit: smoke, default bench, reference arm.  The ncu evidence (launch lists, --set full summaries) comes from
# tools/gpu_round_ncu.sh, which keeps the large .ncu-rep files on the box (gpurun returns at most 64 MiB).
# usage: bash tools/gpu_round.sh <tag>
TAG=${1:-rXX}
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_$TAG.json; cat gpurun_out/bench_ref_$TAG.json
ls -la gpurun_out
