"""How much of the text tower hides behind the vision tower: CUDA-graph replays of (a) the vision tower alone (forward + backward),
(b) the text tower alone, (c) both on two streams as lpi_step.train_step runs them.  python tools/tower_overlap.py [--batch 64]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpi_b200 import lpi_step, synthetic as S  # noqa: E402
from lpi_b200.engine import TextEngine, VisionEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda")
    sd = S.make_clip_state_dict(0)
    vision, text = VisionEngine(sd, dev), TextEngine(sd, dev)
    fac = {k: v.to(dev) for k, v in S.make_prompt_factors(0).items()}
    images = S.make_images(a.batch, 0).to(dev)
    tokens_host = S.make_tokens(a.batch, 0)
    text_len = int(tokens_host.argmax(dim=-1).max()) + 1
    tokens = tokens_host.to(dev)
    d = torch.randn(a.batch, 512, device=dev) * 1e-2
    fv = (fac["dim_1_share"].unsqueeze(0), fac["dim_2_visual"].unsqueeze(0), fac["dim_3_visual"].unsqueeze(0), 1.0)
    ft = (fac["dim_1_share"].unsqueeze(0), fac["dim_2_textual"].unsqueeze(0), fac["dim_3_textual"].unsqueeze(0), 1.0)

    def run_v():
        tape = {}
        vision.forward(images, None, None, tape, (), factors=fv)
        return vision.backward(tape, d)

    def run_t():
        tape = {}
        text.forward(tokens, None, None, tape, (), text_len=text_len, factors=ft)
        return text.backward(tape, d)

    def run_both():
        main_s = torch.cuda.current_stream()
        side = lpi_step._side_stream(dev)
        side.wait_stream(main_s)
        with torch.cuda.stream(side):
            gt = run_t()
        gv = run_v()
        main_s.wait_stream(side)
        return gv, gt

    res = {}
    for name, fn in (("vision", run_v), ("text", run_t), ("both", run_both)):
        warm = torch.cuda.Stream()
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm):
            fn()
            fn()
        torch.cuda.current_stream().wait_stream(warm)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = fn()
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / a.iters
        del keep
    res["hidden_ms"] = res["vision"] + res["text"] - res["both"]
    print(json.dumps({k: round(v, 4) for k, v in res.items()}))


if __name__ == "__main__":
    main()
