#!/bin/bash
# Short device-resident bench of the in-tree library (with / without ntt16x_kernel) and of every variant under build/variants/.
# usage: bash tools/gpu_variants.sh
mkdir -p gpurun_out
run() {
  HBMPC_LIB=$1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-robust-leg --log2-e2e-batch 16 2>gpurun_out/variants.err | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); b=d['breakdown']; print('$2', round(d['value']/1e9,2), 'G/s gen', round(b['gen_ms'],2), 'recon', round(b['recon_ms'],2), '43:', round(b['recon_43_senders_ms'],2), 'flags', round(b['recon_43_senders_flags_ms'],2), 'dense', round(b['recon_43_senders_dense_ms'],2))"
}
run "" in-tree | tee gpurun_out/variants.log
HBMPC_NTT16X=0 run "" in-tree-no16x | tee -a gpurun_out/variants.log
for f in build/variants/*.so; do [ -e "$f" ] && run $PWD/$f $(basename $f) | tee -a gpurun_out/variants.log; done
