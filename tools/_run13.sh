export PATH=/usr/local/cuda/bin:$PATH
timeout 600 python -m pytest tests/test_gpu_scorer.py -m gpu -q --tb=line -x 2>&1 | tail -4
timeout 300 python tools/gpu_ladder.py bench_scorer 2>&1 | grep -v "^====="
