"""Profiling target: a few launches of the attention kernels at the training shape (B=64, L=213, H=12)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpi_b200 import ops
B, L, H = 64, int(os.environ.get("ATT_L", "213")), 12
causal = False
D = H * 64
qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
d_out = torch.randn(B * L, D, device="cuda").bfloat16()
for _ in range(4):
    out, lse = ops.attn_fwd(qkv, B, L, H, causal)
    ops.attn_bwd(qkv, out, d_out, lse, B, L, H, causal)
torch.cuda.synchronize()
