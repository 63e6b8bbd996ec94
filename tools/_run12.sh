export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 3 -c 1 -f -o gpurun_out/scorer_full python bench.py --steps 1 --warmup 3 --skip-cpu --skip-train > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/scorer_full.ncu-rep --page raw --csv > gpurun_out/scorer_full_raw.csv 2>/dev/null
ncu -i gpurun_out/scorer_full.ncu-rep --page source --csv > gpurun_out/scorer_full_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
