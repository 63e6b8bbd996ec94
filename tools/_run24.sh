export PATH=/usr/local/cuda/bin:$PATH
python tools/attn_trace.py 2>&1 | tail -6
