export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
python tools/bench_train.py --batch 64 | tee gpurun_out/train_b64.json
python tools/bench_train.py --batch 64 --text-precision tf32 | tee gpurun_out/train_b64_tf32.json
python tools/bench_train.py --batch 64 --fwd-only | tee gpurun_out/train_b64_fwd.json
