export PATH=/usr/local/cuda/bin:$PATH
for cfg in "8 1" "8 2" "4 2" "16 2" "16 4"; do set -- $cfg
python bench.py --steps 5 --warmup 3 --skip-cpu --skip-train --e2e-chunks $1 --e2e-streams $2 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', d['e2e'], d['ms_per_step'])"
done
