export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
python tools/bench_train.py --batch 64
python tools/bench_train.py --batch 64 --fwd-only
python tools/bench_train.py --batch 256
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lpi|gemm_|attn|layernorm|head_|assemble|im2col|prompt|clip_|sgemm|sgd|row_mean|add_row|gram|task_loss|axpy" -s 1500 -c 700 --csv --log-file gpurun_out/train_launches.csv python tools/bench_train.py --batch 64 --steps 2 --warmup 3 > gpurun_out/ncu_train.log 2>&1
