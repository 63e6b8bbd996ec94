export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q --tb=short 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_n$N.err | tail -1 | tee gpurun_out/bench_n$N.json | cut -c1-3000
tail -3 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
