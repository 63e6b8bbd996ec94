export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py tests/test_gpu_scorer.py -m gpu -q --tb=short 2>&1 | tail -5
python tools/bench_train.py --batch 64 | tee gpurun_out/train_b64.json
python tools/bench_train.py --batch 256 --steps 5
