export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
echo "== fast act"; python tools/parity_margins.py fp16 tf32 bf16 2>&1 | tee gpurun_out/parity_margins.txt
echo "== precise act"; LPI_F16_PRECISE_ACT=1 python tools/parity_margins.py fp16 2>&1 | tee -a gpurun_out/parity_margins.txt
python tools/bench_train.py --batch 64
LPI_F16_PRECISE_ACT=1 python tools/bench_train.py --batch 64
