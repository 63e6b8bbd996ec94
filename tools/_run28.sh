export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 850 --csv --log-file gpurun_out/train_launches.csv python tools/bench_train.py --batch 64 --steps 2 --warmup 3 > gpurun_out/ncu_train.log 2>&1
tail -2 gpurun_out/ncu_train.log
