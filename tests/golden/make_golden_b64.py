"""Golden fixture at the BASELINE batch size (configs[0] / configs[2]: B = 64) from the REAL reference (CPU, fp32).

    python tests/golden/make_golden_b64.py          (build container only: needs /root/reference; ~2 min on 8 cores)

Same procedure as make_golden.py (SliNet.forward -> cal_loss -> backward with the reference's freeze policy,
retrieval/methods/sprompt.py:229-237, 300-311) on 64 synthetic 224x224 images and 64 captions -- M = 64 x 213 = 13 632 vision
rows (ragged last 128-row tile) and 64 x 77 text rows, i.e. the GEMM shapes the throughput numbers are quoted on.  Only small
outputs are stored: 2 x [64, 512] features, the 64 x 64 logits, the losses and the 5 284-element factor gradients of a task-1 and
a task-2 step.  Inputs are regenerated from seeds by lpi_b200/synthetic.py on whichever box runs the tests."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_loader as RL  # noqa: E402
from lpi_b200 import synthetic as S  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import train_step_case  # noqa: E402  (same procedure as the B = 4 fixture)

OUT = os.path.dirname(os.path.abspath(__file__))
B = 64
IMAGE_SEED, CAPTION_SEED = 64, 64


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    sd = S.make_clip_state_dict(0)
    net = RL.build_reference_slinet(sd, {0: S.make_prompt_factors(0), 1: S.make_prompt_factors(1)}, numtask=1)
    images, captions = S.make_images(B, IMAGE_SEED), S.make_captions(B, CAPTION_SEED)
    g = {"meta": {"B": B, "weights_seed": 0, "factor_seeds": [0, 1], "image_seed": IMAGE_SEED, "caption_seed": CAPTION_SEED},
         "captions": captions, "tokens": RL.reference_tokenize(captions)}
    g["step_task1"] = train_step_case(net, images, captions)
    print("task 1", g["step_task1"]["losses"], flush=True)
    net.numtask = 2
    g["step_task2"] = train_step_case(net, images, captions)
    print("task 2", g["step_task2"]["losses"], flush=True)
    for k in ("step_task1", "step_task2"):             # keep the file small: fp32 is what the reference produced, nothing to trim but views
        g[k] = {n: (v.contiguous() if torch.is_tensor(v) else v) for n, v in g[k].items()}
    torch.save(g, os.path.join(OUT, "model_b64_seed0.pt"))


if __name__ == "__main__":
    main()
