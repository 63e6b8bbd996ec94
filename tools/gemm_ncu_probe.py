"""ncu target: one launch of this repo's GEMM and one cuBLAS launch per encoder shape (after one warm-up each), so that a single
`ncu --set full` pass shows both kernels side by side (cycles, tensor-pipe activity, L2 / shared-memory throughput, stalls).
    ncu --set full --clock-control none -o gpurun_out/gemm_probe python tools/gemm_ncu_probe.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpi_b200 import ops

dev = torch.device("cuda")
M = 13632
SHAPES = [("dh2", 768, 3072, ops.EPI_BF16), ("qkv", 2304, 768, ops.EPI_BIAS_BF16), ("fc_gelu", 3072, 768, ops.EPI_BIAS_GELU_BF16),
          ("dgelu", 3072, 768, ops.EPI_DGELU_BF16), ("c_proj", 768, 3072, ops.EPI_BIAS_RESID_F32)]
sel = sys.argv[1:] or [s[0] for s in SHAPES]
for name, N, K, epi in SHAPES:
    if name not in sel:
        continue
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
    kw = {}
    if epi in (ops.EPI_BIAS_BF16, ops.EPI_BIAS_GELU_BF16, ops.EPI_BIAS_RESID_F32):
        kw["bias"] = torch.randn(N, device=dev)
    if epi == ops.EPI_BIAS_RESID_F32:
        kw["resid"] = torch.randn(M, N, device=dev); kw["out"] = torch.empty(M, N, device=dev)
    if epi == ops.EPI_BIAS_GELU_BF16:
        kw["out2"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if epi == ops.EPI_DGELU_BF16:
        kw["aux"] = torch.randn(M, N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        ops.gemm(x, w, epi, **kw)
        torch.matmul(x, w.t(), out=out)
    torch.cuda.synchronize()
    print(name, "done", flush=True)
