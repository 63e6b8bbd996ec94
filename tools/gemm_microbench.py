"""Encoder GEMM shapes of one training step (8 forward + 8 dgrad, vision fp16 (or bf16) + text fp16) at B = 64 and B = 256: this repo's kernel WITH the
fused epilogue the step uses, against cuBLAS (torch.matmul on the same 16-bit operands, no epilogue) on the same box.
CUDA events, 20 launches after 5 warm-ups; between shapes nothing else runs.  Output: profiles/r2_gemm_microbench.txt

    python tools/gemm_microbench.py [--batches 64,256] [--text-len 77]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpi_b200 import ops  # noqa: E402


def time_us(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def shapes(B, text_len, vision_dtype=torch.float16):
    Mv, Mt = B * 213, B * text_len
    out = []
    for tower, M, D, h in (("vision", Mv, 768, vision_dtype), ("text", Mt, 512, torch.float16)):
        out += [
            (tower, "qkv       fwd", M, 3 * D, D, ops.EPI_BIAS_BF16, h),
            (tower, "out-proj  fwd", M, D, D, ops.EPI_BIAS_RESID_F32, h),
            (tower, "c_fc+GELU fwd", M, 4 * D, D, ops.EPI_BIAS_GELU_BF16, h),
            (tower, "c_proj    fwd", M, D, 4 * D, ops.EPI_BIAS_RESID_F32, h),
            (tower, "dz=dGELU  bwd", M, 4 * D, D, ops.EPI_DGELU_BF16, h),
            (tower, "dh2       bwd", M, D, 4 * D, ops.EPI_BF16, h),
            (tower, "do        bwd", M, D, D, ops.EPI_BF16, h),
            (tower, "dh1       bwd", M, D, 3 * D, ops.EPI_BF16, h),
        ]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="64,256")
    ap.add_argument("--text-len", type=int, default=77)
    ap.add_argument("--vision-dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--only", default="", help="substring filter on the GEMM name (e.g. dGELU)")
    a = ap.parse_args()
    dev = torch.device("cuda")
    print(f"# {torch.cuda.get_device_name(0)}; ours = lpi_gemm_{{bf16,f16}} with the fused epilogue; cuBLAS = torch.matmul(a, w.t()) same operands, no epilogue")
    print(f"# {'tower':6s} {'gemm':14s} {'M':>6s} {'N':>5s} {'K':>5s} | {'ours us':>8s} {'TFLOP/s':>8s} | {'cuBLAS us':>9s} {'TFLOP/s':>8s} | cuBLAS/ours | cuBLAS + eager epilogue us, / ours")
    for B in [int(x) for x in a.batches.split(",")]:
        tot_o = tot_c = tot_f = tot_s = 0.0
        for tower, name, M, N, K, epi, h in shapes(B, a.text_len, torch.float16 if a.vision_dtype == "fp16" else torch.bfloat16):
            if a.only and a.only not in name:
                continue
            g = torch.Generator(device=dev).manual_seed(1)
            x = torch.randn(M, K, device=dev, generator=g).to(h)
            w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).to(h)
            bias = torch.randn(N, device=dev, generator=g)
            kw = {}
            if epi in (ops.EPI_BIAS_BF16, ops.EPI_BIAS_GELU_BF16, ops.EPI_BIAS_RESID_F32):
                kw["bias"] = bias
            if epi == ops.EPI_BIAS_RESID_F32:
                kw["resid"] = torch.randn(M, N, device=dev, generator=g)
                kw["out"] = torch.empty(M, N, device=dev)
            if epi == ops.EPI_BIAS_GELU_BF16:
                kw["out2"] = torch.empty(M, N, device=dev, dtype=h)
            if epi == ops.EPI_DGELU_BF16:
                kw["aux"] = torch.randn(M, N, device=dev, generator=g).to(h)
            if "out" not in kw:
                kw["out"] = torch.empty(M, N, device=dev, dtype=h)
            ours = time_us(lambda: ops.gemm(x, w, epi, **kw))
            wt = w.t()
            ref_out = torch.empty(M, N, device=dev, dtype=h)
            cublas = time_us(lambda: torch.matmul(x, wt, out=ref_out))
            # the same OPERATION through stock torch: cuBLAS (with its own bias epilogue where addmm offers one) + the element-wise kernels the
            # fused epilogue replaces, same operands and outputs
            def stock():
                if epi == ops.EPI_BIAS_BF16:
                    return torch.addmm(bias_h, x, wt, out=ref_out)
                if epi == ops.EPI_BIAS_RESID_F32:
                    y = torch.addmm(bias_h, x, wt, out=ref_out)
                    return torch.add(kw["resid"], y, out=stock_f32)
                if epi == ops.EPI_BIAS_GELU_BF16:
                    z = torch.addmm(bias_h, x, wt, out=ref_out)
                    return torch.mul(z, torch.sigmoid(1.702 * z), out=stock_h)
                if epi == ops.EPI_DGELU_BF16:
                    y = torch.matmul(x, wt, out=ref_out)
                    sg = torch.sigmoid(1.702 * kw["aux"])
                    return torch.mul(y, sg * (1 + 1.702 * kw["aux"] * (1 - sg)), out=stock_h)
                return torch.matmul(x, wt, out=ref_out)
            bias_h = bias.to(h)
            stock_f32 = torch.empty(M, N, device=dev) if epi == ops.EPI_BIAS_RESID_F32 else None
            stock_h = torch.empty(M, N, device=dev, dtype=h)
            stock_us = time_us(stock)
            fl = 2.0 * M * N * K
            if tower == "vision":
                tot_o += ours; tot_c += cublas; tot_f += fl; tot_s += stock_us
            print(f"  {tower:6s} {name:14s} {M:6d} {N:5d} {K:5d} | {ours:8.1f} {fl / ours / 1e6:8.0f} | {cublas:9.1f} {fl / cublas / 1e6:8.0f} | {cublas / ours:5.2f}x"
                  f" | {stock_us:8.1f} {stock_us / ours:5.2f}x", flush=True)
            del x, w, kw, ref_out, stock_f32, stock_h
        print(f"  B={B} vision-layer GEMM aggregate: ours {tot_f / tot_o / 1e6:.0f} TFLOP/s ({tot_o:.0f} us), cuBLAS without epilogues {tot_f / tot_c / 1e6:.0f} TFLOP/s ({tot_c:.0f} us), cuBLAS + eager epilogues {tot_s:.0f} us", flush=True)


if __name__ == "__main__":
    main()
