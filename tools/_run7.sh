export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --tb=short 2>&1 | tail -15
python tools/bench_train.py --batch 64
python tools/bench_train.py --batch 64 --fwd-only
python tools/gpu_ladder.py bench_gemm 2>&1 | grep -v "^=====" 

