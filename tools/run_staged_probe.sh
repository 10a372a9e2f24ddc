python -m pytest tests/test_gpu_staged_decoder.py -x -q 2>&1 | tail -3
python tools/bench_configs.py --which c4 --log2-c4 17 2>&1 | grep -E '"ms"|per_s"|recovered'
python tools/bench_configs.py --which c4 --log2-c4 20 2>&1 | grep -E '"ms"|per_s"|recovered'
python tools/bench_configs.py --which c3r --log2 20 2>&1 | grep -E 'staged_decode|"ms"|recovered|chunks_per_s"' | tail -8
