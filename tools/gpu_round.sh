#!/bin/bash
# One GPU visit: smoke, default bench, launch list, ncu full capture of the matvec kernel.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
SHORT="python bench.py --steps 2 --warmup 1 --log2-batch 20 --log2-e2e-batch 16 --no-cpu-baseline"
$SHORT > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_list.log 2>&1
$SHORT > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:matvec -s 2 -c 3 -o gpurun_out/prof_matvec $SHORT > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
