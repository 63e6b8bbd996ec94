export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_n$N.err | tail -1 | tee gpurun_out/bench_n$N.json | cut -c1-2600
tail -3 gpurun_out/bench_n$N.err
