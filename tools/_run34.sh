export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests/test_gpu_scorer.py -m gpu -q --tb=short -x -k full_size 2>&1 | tail -8
