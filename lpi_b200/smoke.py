"""One small invocation of the hot path on cuda:0, checked against the oracle (driver's smoke())."""
from __future__ import annotations

import numpy as np
import torch


def run() -> None:
    from oracle import lpi_oracle as O          # checker only
    from . import ops, retrieval as R, synthetic as S

    dev = torch.device("cuda", torch.cuda.current_device())
    # 1. Flickr-shaped Recall@K (BASELINE.json configs[1], scaled down): feature path vs the oracle's itm_eval
    img, txt, img2txt, txt2img, cat_i, cat_t = S.make_retrieval_set(200, 5, 512, 5, seed=5)
    s = (img @ txt.t()).numpy()
    want = O.itm_eval(s, np.ascontiguousarray(s.T), txt2img, img2txt, cat_i, cat_t, 5)
    got = R.itm_eval_features(img.to(dev), txt.to(dev), txt2img, img2txt, cat_i, cat_t, 5, precision="fp32")
    assert got == want, f"Recall@K mismatch\n got {got}\nwant {want}"
    got_dense = R.itm_eval(s, np.ascontiguousarray(s.T), txt2img, img2txt, cat_i, cat_t, 5)
    assert got_dense == want, "dense itm_eval mismatch"
    # 2. one encoder-shaped tensor-core GEMM with a fused epilogue vs an fp32 matmul
    g = torch.Generator().manual_seed(0)
    a = torch.randn(426, 768, generator=g).to(dev).bfloat16()
    w = (torch.randn(2304, 768, generator=g) * 768 ** -0.5).to(dev).bfloat16()
    b = torch.randn(2304, generator=g).to(dev)
    out = ops.gemm(a, w, ops.EPI_BIAS_BF16, bias=b)
    ref = a.float() @ w.float().t() + b
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2 * ref.abs().max().item(), f"gemm error {err}"
    torch.cuda.synchronize()
    print("smoke ok: Recall@K bit-exact vs oracle; gemm max err %.3e; lpi kernels launched: %d" % (err, ops.KERNEL_LAUNCHES))
