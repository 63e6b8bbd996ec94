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
    # 3. one LPI training step (both towers fwd, 3 losses, backward to the prompt factors) vs the fixture produced by the REAL reference
    import os

    from . import lpi_step
    from .engine import TextEngine, VisionEngine

    gold = torch.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "model_b4_seed0.pt"),
                      weights_only=False)
    sd = S.make_clip_state_dict(0)
    vision, text = VisionEngine(sd, dev), TextEngine(sd, dev)
    fac = {k: v.to(dev) for k, v in S.make_prompt_factors(0).items()}
    r = lpi_step.train_step(vision, text, fac, S.make_images(4, 0).to(dev), gold["tokens"].to(dev), 1 / 0.07)
    want = gold["step_task1"]
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    assert rel(r["img_f"], want["img_f"]) < 1e-2 and rel(r["txt_f"], want["txt_f"]) < 1e-2, "feature parity"
    lg_err = rel(r["logits"], want["logits"])
    lg_cos = float((r["logits"].double().cpu() - want["logits"].double()).abs().max()) * 0.07
    assert lg_err < 1e-2 and lg_cos < 1e-2, f"logits parity: Frobenius-relative {lg_err:.2e}, max-abs / logit_scale {lg_cos:.2e}"
    for k2, v2 in want["losses"].items():
        assert abs(float(r["losses"][k2]) - v2) < 1e-2 * max(abs(v2), 1e-3), f"loss {k2}"
    worst = max(rel(r["grads"][k2], want["grads"][k2]) for k2 in O.FACTOR_NAMES)
    assert worst < 2e-2, f"prompt-gradient parity {worst}"
    torch.cuda.synchronize()
    print("train step ok: base_loss %.4f, logits rel err %.2e (max-abs/scale %.2e), worst prompt-grad rel err %.2e"
          % (float(r["losses"]["base_loss"]), lg_err, lg_cos, worst))
    print("smoke ok: Recall@K bit-exact vs oracle; gemm max err %.3e; lpi kernels launched: %d" % (err, ops.KERNEL_LAUNCHES))
