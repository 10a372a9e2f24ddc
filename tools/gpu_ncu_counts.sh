#!/bin/bash
# Executed-instruction evidence for bench.py's roofline (profiles/r02_executed_counts.json): ncu source counters (per SASS line
# executed counts) of the headline transforms, of the dense kernel and of one K4 call.  The source pages are reduced ON THE BOX to
# small per-kernel JSON summaries (tools/ncu_exec_counts.py): gpurun returns at most 64 MiB.
# usage: bash tools/gpu_ncu_counts.sh <tag>
TAG=${1:-rXX}
mkdir -p gpurun_out
SEC="--section SourceCounters --section LaunchStats --section SpeedOfLight --section WarpStateStats --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis"
# headline transforms: one K1 and one K3 launch at 2^20 items (ncu --set full: also the dram traffic of profiles/ncu_traffic_per_2p20.json)
python tools/ncu_ntt.py > gpurun_out/ncu_plain_ntt_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"ntt" -s 2 -c 2 -o /tmp/ntt_$TAG python tools/ncu_ntt.py > gpurun_out/ncu_ntt_$TAG.log 2>&1
ncu -i /tmp/ntt_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_ntt_$TAG.csv 2>/dev/null
ncu -i /tmp/ntt_$TAG.ncu-rep --page source --csv > /tmp/src_ntt_$TAG.csv 2>/dev/null
python tools/ncu_exec_counts.py /tmp/src_ntt_$TAG.csv --items 1048576 --json gpurun_out/exec_ntt_$TAG.json > gpurun_out/exec_ntt_$TAG.txt
# dense kernel (43 senders, flags): one launch at 2^20 chunks
python tools/ncu_ntt.py --what dense --reps 2 > gpurun_out/ncu_plain_dense_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"matvec" -s 1 -c 1 -o /tmp/dense_$TAG python tools/ncu_ntt.py --what dense --reps 2 > gpurun_out/ncu_dense_$TAG.log 2>&1
ncu -i /tmp/dense_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_dense_$TAG.csv 2>/dev/null
ncu -i /tmp/dense_$TAG.ncu-rep --page source --csv > /tmp/src_dense_$TAG.csv 2>/dev/null
python tools/ncu_exec_counts.py /tmp/src_dense_$TAG.csv --items 1048576 --json gpurun_out/exec_dense_$TAG.json > gpurun_out/exec_dense_$TAG.txt
# K4: every kernel of ONE robust_interpolate_batch call (n=128, t=42, 2^17 codewords, e~U{0..42}); the call is bracketed by
# cudaProfilerStart/Stop in tools/ncu_ntt.py
python tools/ncu_ntt.py --what k4 --log2 17 --reps 2 > gpurun_out/ncu_plain_k4_$TAG.log 2>&1 && \
ncu $SEC --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --import-source on --profile-from-start off -o /tmp/k4_$TAG python tools/ncu_ntt.py --what k4 --log2 17 --reps 2 > gpurun_out/ncu_k4_$TAG.log 2>&1
ncu -i /tmp/k4_$TAG.ncu-rep --page raw --csv > /tmp/raw_k4_$TAG.csv 2>/dev/null
python tools/ncu_summary.py /tmp/raw_k4_$TAG.csv gpurun_out/raw_k4_summary_$TAG.json > /dev/null 2>&1
python - /tmp/raw_k4_$TAG.csv gpurun_out/k4_traffic_$TAG.json <<'PYEOF'
# DRAM traffic and duration of the WHOLE captured K4 call (all launches), per kernel name and in total
import csv, json, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name):
    return hdr.index(name) if name in hdr else None
ir, iw, it, ik = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum"), col("Kernel Name")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
per = collections.OrderedDict()
for r in data:
    if ir is None or iw is None:
        break
    e = per.setdefault(r[ik].split("(")[0], {"launches": 0, "dram_bytes": 0.0, "ms": 0.0})
    e["launches"] += 1
    e["dram_bytes"] += float(r[ir] or 0) * scale.get(units[ir], 1) + float(r[iw] or 0) * scale.get(units[iw], 1)
    if it is not None:
        e["ms"] += float(r[it] or 0) * tscale.get(units[it], 1)
tot = {"launches": sum(e["launches"] for e in per.values()), "dram_bytes": sum(e["dram_bytes"] for e in per.values()), "ms": sum(e["ms"] for e in per.values())}
json.dump({"what": "one robust_interpolate_batch call, n=128, t=42, 2^17 codewords, e~U{0..42}: dram__bytes_read.sum + dram__bytes_write.sum and gpu__time_duration.sum over all its launches (ncu, serialised)", "codewords": 131072, "total": tot, "per_kernel": per}, open(sys.argv[2], "w"), indent=1)
print(tot)
PYEOF
ncu -i /tmp/k4_$TAG.ncu-rep --page source --csv > /tmp/src_k4_$TAG.csv 2>/dev/null
python tools/ncu_exec_counts.py /tmp/src_k4_$TAG.csv --items 131072 --aggregate --json gpurun_out/exec_k4_$TAG.json > gpurun_out/exec_k4_$TAG.txt
cat gpurun_out/ncu_plain_k4_$TAG.log
ls -la gpurun_out/*$TAG* /tmp/*.csv; du -sh gpurun_out
