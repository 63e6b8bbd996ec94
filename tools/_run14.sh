export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --tb=short 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -3 | cut -c1-1800
