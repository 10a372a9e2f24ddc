#!/bin/bash
for f in build/variants/*.so; do
  echo "== $f"
  HBMPC_LIB=$PWD/$f python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['breakdown']['gen_ms'], d['breakdown']['recon_ms'])"
done
