"""Prints the parity margins of the fused training step against the real-reference golden fixture (tests/golden/model_b4_seed0.pt):
feature / loss / prompt-gradient relative errors per text-tower precision.  Bars: 1e-2 (features, losses), 2e-2 (gradients)."""
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
    g = torch.load(os.path.join(ROOT, "tests", "golden", "model_b4_seed0.pt"), weights_only=False)
    images = S.make_images(g["meta"]["B"], 0).cuda()
    tokens = g["tokens"].cuda()
    vision = VisionEngine(sd, dev)
    for prec in sys.argv[1:] or ["fp16", "tf32", "bf16"]:
        text = TextEngine(sd, dev, precision=prec)
        for task in (1, 2):
            if task == 1:
                fac = {k: v.cuda() for k, v in S.make_prompt_factors(0).items()}
                r = lpi_step.train_step(vision, text, fac, images, tokens, 1 / 0.07)
            else:
                fac = {k: v.cuda() for k, v in S.make_prompt_factors(1).items()}
                prev = [lpi_step.reconstruct({k: v.cuda() for k, v in S.make_prompt_factors(0).items()})]
                sim = np.loadtxt(os.path.join(ROOT, "lpi_b200", "MID", "task_sim_matrix.txt"))
                tgt = torch.tensor((sim[:2, :2] > 0.4).astype(np.int32)).cuda()
                r = lpi_step.train_step(vision, text, fac, images, tokens, 1 / 0.07, prev, tgt)
            want = g[f"step_task{task}"]
            grads = " ".join(f"{k.replace('dim_', 'd')}={rel(r['grads'][k], want['grads'][k]):.2e}" for k in FACTOR_NAMES)
            print(f"text={prec} task{task}: img_f {rel(r['img_f'], want['img_f']):.2e} txt_f {rel(r['txt_f'], want['txt_f']):.2e} "
                  f"base_loss {abs(float(r['losses']['base_loss']) - want['losses']['base_loss']) / want['losses']['base_loss']:.2e} | {grads}", flush=True)
        del text


if __name__ == "__main__":
    main()
