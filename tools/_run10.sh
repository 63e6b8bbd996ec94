export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 300 python tools/gpu_ladder.py gemm_min 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_scorer.py -m gpu -q --tb=line -x 2>&1 | tail -12
timeout 300 python tools/gpu_ladder.py bench_gemm bench_scorer 2>&1 | grep -v "^====="
LPI_SCORER_PAIR=0 timeout 300 python tools/gpu_ladder.py bench_scorer 2>&1 | grep -v "^====="
