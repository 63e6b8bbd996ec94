export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
K='regex:gemm_|attn_|layernorm|head_|assemble|prompt_|im2col|sgemm|clip_|task_|sgd|gram|row_mean|add_row|nearest|inject|sum_prompt|l2_'
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 400 -c 12 -f -o gpurun_out/train_fwd python tools/bench_train.py --batch 64 --steps 1 --warmup 2 > gpurun_out/ncu_tf.log 2>&1
tail -2 gpurun_out/ncu_tf.log
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 600 -c 12 -f -o gpurun_out/train_bwd python tools/bench_train.py --batch 64 --steps 1 --warmup 2 > gpurun_out/ncu_tb.log 2>&1
tail -2 gpurun_out/ncu_tb.log
ls -la gpurun_out/*.ncu-rep
