export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_learner.py -m gpu -q --tb=short 2>&1 | tail -3
python tools/parity_margins.py fp16
LPI_DH_F32=1 python tools/parity_margins.py fp16
python tools/bench_train.py --batch 64
LPI_DH_F32=1 python tools/bench_train.py --batch 64
