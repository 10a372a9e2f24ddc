#!/bin/bash
# A/B timing of the headline launches under environment knobs: bash tools/gpu_ab.sh "<env A>" "<env B>" ...
# prints gen_ms / recon_ms (and the 43-sender first-call times) of a short device-resident bench run for each setting
for E in "$@"; do
  echo "== $E"
  env $E python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --log2-e2e-batch 16 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); b=d['breakdown']; print('gen_ms %.3f recon_ms %.3f value %.2f G' % (b['gen_ms'], b['recon_ms'], d['value']/1e9))
    elif l: print(l[:300])
"
done
