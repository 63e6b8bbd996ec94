"""Times the CTA-pair GEMM on the vision-tower shapes (CUDA events, 30 launches after warm-up).  With LPI_GEMM_PROBE=1 the kernel skips its
A-tile TMA loads (results are WRONG): the time difference is what the shared-memory fill of the streamed operand costs.
The probe only exists in a library built with -DLPI_DEBUG_PROBE (never the shipped one):
    nvcc <flags of __graft_entry__.NVCC_FLAGS> -DLPI_DEBUG_PROBE -shared lpi_b200/csrc/*.cu -o /tmp/liblpi_probe.so
    LPI_LIB_PATH=/tmp/liblpi_probe.so LPI_GEMM_PROBE=1 python tools/gemm_probe.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpi_b200 import ops
print("LPI_GEMM_PROBE =", os.environ.get("LPI_GEMM_PROBE", "0"))
for (M, N, K, epi, name) in [(13632, 2304, 768, ops.EPI_BIAS_BF16, "qkv"), (13632, 768, 768, ops.EPI_BF16, "do"),
                             (13632, 3072, 768, ops.EPI_BF16, "fc plain"), (13632, 768, 3072, ops.EPI_F32, "dh2")]:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    kw = dict(bias=bias) if epi == ops.EPI_BIAS_BF16 else {}
    for _ in range(5):
        ops.gemm(a, w, epi, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        ops.gemm(a, w, epi, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    print(f"{name:9s} M{M} N{N} K{K}: {us:7.1f} us  {2.0 * M * N * K / us / 1e6:7.0f} TFLOP/s", flush=True)
