#!/bin/bash
# Short device-resident bench of the in-tree library and of every variant under build/variants/ (tools/build_variant.sh).
# usage: bash tools/gpu_variants.sh [--configs]
mkdir -p gpurun_out
EXTRA="--no-configs"; [ "$1" == "--configs" ] && EXTRA="--log2-c4 17 --log2-c5 18"
run() {
  HBMPC_LIB=$1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --log2-e2e-batch 16 $EXTRA 2>gpurun_out/variants.err | tail -1 | \
    python -c "
import sys,json
d=json.loads(sys.stdin.read()); b=d['breakdown']; c=d.get('configs',{}); f=c.get('c3_first_call',{}); c4=c.get('c4',{}); c5=c.get('c5',{})
print('$2', round(d['value']/1e9,2), 'G/s gen', round(b['gen_ms'],3), 'recon', round(b['recon_ms'],3), '| 43 senders: coeffs', round(f.get('coeffs_ms',0),2), 'secrets', round(f.get('secrets_ms',0),2), 'flags', round(f.get('coeffs_flags_ms',0),2), 'dense', round(f.get('dense_flags_ms',0),2),
 '| c2', round(c.get('c2',{}).get('shares_per_s',0)/1e9,1), 'G/s | c4 2^17 Mcw/s', {k: round(v['codewords_per_s']/1e6,1) for k,v in c4.items() if isinstance(v,dict) and 'codewords_per_s' in v}, '| c5 2^18 ms', c5.get('total_ms'))"
}
run "" in-tree | tee gpurun_out/variants.log
for f in build/variants/*.so; do [ -e "$f" ] && run $PWD/$f $(basename $f) | tee -a gpurun_out/variants.log; done
