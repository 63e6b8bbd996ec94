export PATH=/usr/local/cuda/bin:$PATH
for SR in 4096 32768 65536; do for G in 5000000 625000; do
LPI_SEED_ROWS=$SR python bench.py --steps 10 --warmup 3 --skip-train --skip-cpu --gallery $G 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print($SR, $G, 'ms', round(d['ms_per_step'],3), 'kern', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],2))"
done; done
