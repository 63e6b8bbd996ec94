export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
LPI_ATTN_TC=1 timeout 300 python tools/attn_debug.py 2>&1 | tee gpurun_out/attn_tc.log
LPI_ATTN_TC=0 timeout 300 python tools/attn_debug.py 2>&1 | tail -4 | tee gpurun_out/attn_legacy.log
