#!/bin/bash
# One GPU-box visit (gpurun -- 'bash tools/gpu_round.sh'): parity tests, smoke, bench (both arms), ncu launch lists and full captures of the
# dominant kernels.  Everything lands in gpurun_out/ (scratch, <= 64 MiB: .ncu-rep files stay in /tmp, only CSV exports travel); summaries
# worth keeping are copied to profiles/.
set -u
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|NUMA" >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.log
echo "== bench (reference arm, then ours)"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-300
timeout 1500 python bench.py 2> gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json | cut -c1-400; tail -3 gpurun_out/bench.err
echo "== ncu launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_|topk|recall|merge" -c 200 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-train --skip-sweep > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 850 --csv --log-file gpurun_out/train_launches.csv \
  python tools/bench_train.py --batch 64 --steps 2 --warmup 3 > gpurun_out/ncu_train.log 2>&1
echo "== ncu full: scorer main pass at 5 M rows (N = 1) and at the shards of the 2 / 4 / 8-GPU runs"
for n in 5000000 2500000 1250000 625000; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 7 -c 1 -f -o /tmp/scorer_full_$n \
    python bench.py --gallery $n --steps 1 --warmup 3 --skip-cpu --skip-train --skip-sweep > gpurun_out/ncu_full_$n.log 2>&1
  ncu -i /tmp/scorer_full_$n.ncu-rep --page raw --csv > gpurun_out/scorer_full_raw_$n.csv 2>/dev/null
done
echo "== ncu full: one forward + one backward vision layer of the training step"
timeout 900 ncu --set full --clock-control none -k regex:"gemm_pair|attn_|layernorm" -s 470 -c 10 -f -o /tmp/train_fwd python tools/bench_train.py --batch 64 --steps 1 --warmup 2 > gpurun_out/ncu_train_fwd.log 2>&1
ncu -i /tmp/train_fwd.ncu-rep --page raw --csv > gpurun_out/train_fwd_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:"gemm_pair|attn_|layernorm" -s 640 -c 12 -f -o /tmp/train_bwd python tools/bench_train.py --batch 64 --steps 1 --warmup 2 > gpurun_out/ncu_train_bwd.log 2>&1
ncu -i /tmp/train_bwd.ncu-rep --page raw --csv > gpurun_out/train_bwd_raw.csv 2>/dev/null
ls -la gpurun_out | head -40
