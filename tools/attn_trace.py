"""Prints the clock64 timeline recorded by the attention kernels (debug aid)."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpi_b200 import ops, _lib
B, L, H = 64, 213, 12
D = H * 64
qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
d_out = torch.randn(B * L, D, device="cuda").bfloat16()
for _ in range(3):
    out, lse = ops.attn_fwd(qkv, B, L, H, False)
    ops.attn_bwd(qkv, out, d_out, lse, B, L, H, False)
torch.cuda.synchronize()
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
_lib.lib().lpi_debug_attn_trace(C.c_void_p(buf.data_ptr()))
def show(tag):
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    for c in range(2):
        ev = [(i, t[64 * c + i]) for i in range(64) if t[64 * c + i]]
        if not ev: continue
        t0 = min(v for _, v in ev)
        print(tag, "CTA", "first" if c == 0 else "last", " ".join(f"{i}:{v - t0}" for i, v in sorted(ev, key=lambda kv: kv[1])))
    buf.zero_()
out, lse = ops.attn_fwd(qkv, B, L, H, False); show("fwd")
ops.attn_bwd(qkv, out, d_out, lse, B, L, H, False); show("bwd")
