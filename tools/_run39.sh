export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_|topk|recall|split|l2_norm" -c 200 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-train > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 7 -c 1 -f -o gpurun_out/scorer_full \
  python bench.py --steps 1 --warmup 3 --skip-cpu --skip-train > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/scorer_full.ncu-rep --page raw --csv > gpurun_out/scorer_full_raw.csv 2>/dev/null
ls -la gpurun_out | head -30
