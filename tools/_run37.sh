export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scorer.py tests/test_abi.py -m gpu -q --tb=short -x 2>&1 | tail -4
for G in 5000000 625000; do
python bench.py --steps 10 --warmup 3 --skip-train --skip-cpu --gallery $G 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print($G, 'ms', d['ms_per_step'], 'kern', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['recall'], d['gpu_launches'])"
done
