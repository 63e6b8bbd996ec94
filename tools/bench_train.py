"""Times one LPI training step (COCO-shaped, BASELINE.json configs[2]) on one GPU: forward of both towers, three losses,
backward to the prompt factors, SGD step.  python tools/bench_train.py [--batch 64] [--steps 10] [--text-precision fp16|tf32|bf16]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpi_b200 import lpi_step, ops, synthetic as S  # noqa: E402
from lpi_b200.engine import TextEngine, VisionEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--text-precision", default="fp16")
    ap.add_argument("--fwd-only", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the step from a CUDA graph (lpi_step.GraphedTrainStep)")
    ap.add_argument("--no-trim", action="store_true", help="run the text tower on all 77 positions (default: up to the batch's last EOT)")
    a = ap.parse_args()
    dev = torch.device("cuda")
    sd = S.make_clip_state_dict(0)
    vision, text = VisionEngine(sd, dev), TextEngine(sd, dev, precision=a.text_precision)
    fac = {k: v.to(dev) for k, v in S.make_prompt_factors(0).items()}
    opt = lpi_step.PromptSGD(fac, 0.05)
    images = S.make_images(a.batch, 0).to(dev)
    tokens_host = S.make_tokens(a.batch, 0)
    text_len = None if a.no_trim else int(tokens_host.argmax(dim=-1).max()) + 1
    tokens = tokens_host.to(dev)

    def step():
        if a.fwd_only:
            vis, txt = lpi_step.reconstruct(fac)
            vision.forward(images, vis.unsqueeze(0))
            text.forward(tokens, txt.unsqueeze(0), text_len=text_len)
            return
        r = lpi_step.train_step(vision, text, fac, images, tokens, 1 / 0.07, text_len=text_len)
        opt.step(r["grads"])

    if a.graph and not a.fwd_only:
        step = lpi_step.GraphedTrainStep(vision, text, fac, opt, images, tokens, 1 / 0.07, text_len=text_len).step
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    n0 = ops.KERNEL_LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / a.steps * 1e3
    ms = e0.elapsed_time(e1) / a.steps
    gflop_pair = 44.05 if a.fwd_only else 89.7
    print(json.dumps({"batch": a.batch, "ms_per_step": ms, "wall_ms_per_step": wall, "pairs_per_s": a.batch / ms * 1e3,
                      "tflops_algorithmic": a.batch * gflop_pair / ms, "launches_per_step": (ops.KERNEL_LAUNCHES - n0) / a.steps,
                      "text_precision": a.text_precision, "fwd_only": a.fwd_only, "text_positions": text_len or 77, "graph": bool(a.graph)}))


if __name__ == "__main__":
    main()
