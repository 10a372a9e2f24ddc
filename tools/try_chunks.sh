#!/bin/bash
for mb in 4 8 16 24 48 96; do
  echo "== chunk ${mb} MB"
  HBMPC_CHUNK_MB=$mb python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']['value'], d['e2e']['ms_per_step'])"
done
