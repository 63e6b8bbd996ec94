"""Prints the parity margins of the fused training step against the real-reference golden fixtures (tests/golden/model_b4_seed0.pt and
model_b64_seed0.pt): feature / logits / loss / prompt-gradient errors per (vision precision, text precision).
Bars (BASELINE.json north_star): features, logits, losses 1e-2 relative; gradients 2e-2.
    python tools/parity_margins.py [vision:text ...]       e.g.  bf16:fp16 fp16:fp16 fp32:fp32 (the exact-fp32 parity mode; bars 1e-5 / 1e-4)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpi_b200 import lpi_step, synthetic as S
from lpi_b200.engine import TextEngine, VisionEngine

FACTOR_NAMES = lpi_step.FACTOR_NAMES


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())


def main():
    dev = torch.device("cuda")
    sd = S.make_clip_state_dict(0)
    sim = np.loadtxt(os.path.join(ROOT, "lpi_b200", "MID", "task_sim_matrix.txt"))
    tgt = torch.tensor((sim[:2, :2] > 0.4).astype(np.int32)).cuda()
    for combo in sys.argv[1:] or ["bf16:fp16", "fp16:fp16"]:
        vp, tp = combo.split(":")
        vision, text = VisionEngine(sd, dev, precision=vp), TextEngine(sd, dev, precision=tp)
        for fixture in ("model_b4_seed0.pt", "model_b64_seed0.pt"):
            path = os.path.join(ROOT, "tests", "golden", fixture)
            if not os.path.isfile(path):
                continue
            g = torch.load(path, weights_only=False)
            B = g["meta"]["B"]
            images = S.make_images(B, g["meta"]["image_seed"]).cuda()
            tokens = g["tokens"].cuda()
            for task in (1, 2):
                if task == 1:
                    fac = {k: v.cuda() for k, v in S.make_prompt_factors(0).items()}
                    r = lpi_step.train_step(vision, text, fac, images, tokens, 1 / 0.07)
                else:
                    fac = {k: v.cuda() for k, v in S.make_prompt_factors(1).items()}
                    prev = [lpi_step.reconstruct({k: v.cuda() for k, v in S.make_prompt_factors(0).items()})]
                    r = lpi_step.train_step(vision, text, fac, images, tokens, 1 / 0.07, prev, tgt)
                want = g[f"step_task{task}"]
                grads = " ".join(f"{k.replace('dim_', 'd')}={rel(r['grads'][k], want['grads'][k]):.2e}" for k in FACTOR_NAMES)
                lg, lw = r["logits"].double().cpu(), want["logits"].double()
                losses = " ".join(f"{k.split('_')[0]}={abs(float(r['losses'][k]) - v) / max(abs(v), 1e-3):.1e}" for k, v in want["losses"].items())
                print(f"vision={vp} text={tp} B={B} task{task}: img_f {rel(r['img_f'], want['img_f']):.2e} txt_f {rel(r['txt_f'], want['txt_f']):.2e} "
                      f"logits fro {rel(lg, lw):.2e} maxabs/scale {float((lg - lw).abs().max()) * 0.07:.2e} (|logit| max {float(lw.abs().max()):.2f}) "
                      f"losses {losses} | {grads}", flush=True)
        del vision, text


if __name__ == "__main__":
    main()
