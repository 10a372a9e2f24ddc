#!/bin/bash
for f in build/variants/*.so; do
  echo "== $f"
  HBMPC_LIB=$PWD/$f python tools/bench_configs.py --which c4 --log2-c4 17 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read())['c4_n128_t42']; print({k:(round(v['ms'],2), round(v['codewords_per_s']/1e6,2)) for k,v in d.items() if k!='B'})"
done
