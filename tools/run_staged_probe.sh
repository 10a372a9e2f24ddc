set -x
python -m pytest tests/test_gpu_staged_decoder.py -x -q 2>&1 | tail -15
HBMPC_STAGED_PROF=1 HBMPC_STAGED_SEG=16 python tools/bench_configs.py --which c4 --log2-c4 17 2>&1 | grep -E 'staged_decode|"ms"|uniform|adversarial|recovered'
HBMPC_NO_SPECULATION=1 HBMPC_STAGED_PROF=1 python tools/bench_configs.py --which c3r --log2 20 2>&1 | tail -16
python tools/bench_configs.py --which c3r --log2 20 2>&1 | tail -14
