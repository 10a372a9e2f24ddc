#!/bin/bash
# end-to-end (host-buffer) leg of bench.py for several pipeline chunk sizes
for mb in 8 16 32 64 128; do
  echo -n "chunk ${mb} MB: "
  HBMPC_CHUNK_MB=$mb python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['e2e']['value']/1e9,3), 'G shares/s', round(d['e2e']['ms_per_step'],1), 'ms', d['e2e'].get('pcie_gbs'))"
done
