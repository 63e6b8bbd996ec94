"""Attention kernel check + timing (debug aid): errors of out / lse / dq / dk / dv vs an fp32 torch reference."""
import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpi_b200 import ops


def ref_attention(qkv, B, L, H, causal):
    D = H * 64
    q, k, v = qkv.float().view(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / 8.0
    if causal:
        s = s + torch.full((L, L), float("-inf"), device=s.device).triu_(1)
    return (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, D), s


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def main():
    print("LPI_ATTN_TC =", os.environ.get("LPI_ATTN_TC"))
    for (B, L, H, causal) in [(2, 213, 12, False), (3, 77, 8, True), (1, 197, 12, False), (2, 64, 2, True), (1, 1, 1, False),
                              (2, 130, 3, True), (1, 256, 1, False), (1, 16, 1, False), (1, 128, 1, False)]:
        g = torch.Generator().manual_seed(L * 31 + H)
        D = H * 64
        qkv = (torch.randn(B * L, 3 * D, generator=g) * 1.5).cuda().bfloat16()
        d_out = torch.randn(B * L, D, generator=g).cuda().bfloat16()
        qr = qkv.float().requires_grad_(True)
        ref, s = ref_attention(qr, B, L, H, causal)
        ref.backward(d_out.float())
        want_lse = torch.logsumexp(s, -1) * math.log2(math.e)
        try:
            out, lse = ops.attn_fwd(qkv, B, L, H, causal)
            torch.cuda.synchronize()
            e_out = float((out.float() - ref).abs().max() / ref.abs().max())
            e_lse = float((lse.view(B, H, L) - want_lse).abs().max())
            dqkv = ops.attn_bwd(qkv, out, d_out, lse, B, L, H, causal)
            torch.cuda.synchronize()
            parts = [rel(dqkv[:, i * D:(i + 1) * D].float(), qr.grad[:, i * D:(i + 1) * D]) if qr.grad[:, i * D:(i + 1) * D].norm() > 1e-6
                     else float(dqkv[:, i * D:(i + 1) * D].float().abs().max()) for i in range(3)]
            print(f"B{B} L{L} H{H} causal={int(causal)}: out {e_out:.2e} lse {e_lse:.2e} dq {parts[0]:.2e} dk {parts[1]:.2e} dv {parts[2]:.2e}"
                  f" nan={bool(torch.isnan(out.float()).any())}/{bool(torch.isnan(dqkv.float()).any())}", flush=True)
        except Exception as e:
            print(f"B{B} L{L} H{H} causal={int(causal)}: EXC {e!r}"[:300], flush=True)
            return
    for (B, L, H, causal) in [(64, 213, 12, False), (64, 77, 8, True), (256, 213, 12, False)]:
        D = H * 64
        qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
        d_out = torch.randn(B * L, D, device="cuda").bfloat16()
        out, lse = ops.attn_fwd(qkv, B, L, H, causal)
        ops.attn_bwd(qkv, out, d_out, lse, B, L, H, causal)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        n = 20
        e[0].record()
        for _ in range(n):
            out, lse = ops.attn_fwd(qkv, B, L, H, causal)
        e[1].record()
        for _ in range(n):
            ops.attn_bwd(qkv, out, d_out, lse, B, L, H, causal)
        e[2].record()
        torch.cuda.synchronize()
        tf, tb = e[0].elapsed_time(e[1]) / n * 1e3, e[1].elapsed_time(e[2]) / n * 1e3
        fl = 4.0 * L * L * 64 * H * B
        print(f"B{B} L{L} H{H}: fwd {tf:.1f} us ({fl / tf / 1e6:.0f} TFLOP/s)  bwd(+delta) {tb:.1f} us ({2 * fl / tb / 1e6:.0f} TFLOP/s algorithmic)", flush=True)


if __name__ == "__main__":
    main()
