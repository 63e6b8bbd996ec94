#!/bin/bash
# One GPU-box visit (gpurun -- 'bash tools/gpu_round.sh'): parity tests, smoke, bench (both arms), train-step bench, ncu launch lists and
# one full capture of the dominant kernel.  Everything lands in gpurun_out/ (scratch); summaries worth keeping are copied to profiles/.
set -u
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench (reference arm, then ours)"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-400
timeout 900 python bench.py 2> gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json | cut -c1-3000; tail -3 gpurun_out/bench.err
echo "== train step"; python tools/bench_train.py --batch 64 | tee gpurun_out/train_b64.json; python tools/bench_train.py --batch 256 --steps 5 | tee gpurun_out/train_b256.json
echo "== parity margins"; python tools/parity_margins.py fp16 2>&1 | tee gpurun_out/parity_margins.txt
echo "== ncu launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_|topk|recall|split|l2_norm" -c 200 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-train > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 850 --csv --log-file gpurun_out/train_launches.csv \
  python tools/bench_train.py --batch 64 --steps 2 --warmup 3 > gpurun_out/ncu_train.log 2>&1
echo "== ncu full (scorer main pass: the 2nd gemm_pair launch of a step)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 7 -c 1 -f -o gpurun_out/scorer_full \
  python bench.py --steps 1 --warmup 3 --skip-cpu --skip-train > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu -i gpurun_out/scorer_full.ncu-rep --page raw --csv > gpurun_out/scorer_full_raw.csv 2>/dev/null
ls -la gpurun_out
