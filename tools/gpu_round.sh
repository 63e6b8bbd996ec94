#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), micro-benches, ncu launch list + full capture.
set -u
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench.json; tail -5 gpurun_out/bench.err


echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_|topk|recall|split|l2_norm" -c 200 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-train > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log
echo "== ncu full (scorer kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 3 -c 1 -f -o gpurun_out/scorer_full \
  python bench.py --steps 1 --warmup 3 --skip-cpu --skip-train > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu -i gpurun_out/scorer_full.ncu-rep --page raw --csv > gpurun_out/scorer_full_raw.csv 2>/dev/null
ls -la gpurun_out
