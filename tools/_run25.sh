export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
LPI_ATTN_TC=1 timeout 300 python tools/attn_debug.py 2>&1 | grep -v -i warn | tee gpurun_out/attn_tc.log
python tools/attn_trace.py 2>&1 | tail -4
