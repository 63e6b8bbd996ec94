export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python tools/bench_train.py --batch 64 | tee gpurun_out/train_b64.json
python tools/bench_train.py --batch 64 --fwd-only | tee gpurun_out/train_b64_fwd.json
python tools/bench_train.py --batch 256 --steps 5 | tee gpurun_out/train_b256.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 900 --csv --log-file gpurun_out/train_launches.csv python tools/bench_train.py --batch 64 --steps 2 --warmup 3 > gpurun_out/ncu_train.log 2>&1
tail -2 gpurun_out/ncu_train.log
