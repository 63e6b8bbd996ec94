"""Generate the golden fixtures in this directory from the REAL reference (CPU, fp32).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The fixtures hold small outputs (embeddings, losses, 5 284-element grads, recall dicts, token ids);
inputs are regenerated from seeds by lpi_b200/synthetic.py on whichever box runs the tests.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_loader as RL  # noqa: E402
from lpi_b200 import synthetic as S  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FACTORS = S.FACTOR_NAMES


def train_step_case(net, images, captions):
    net.train()
    for n, p in net.named_parameters():
        p.requires_grad_(f"prompts.{net.numtask - 1}." in n)     # freeze policy, sprompt.py:229-237
        p.grad = None
    img_f, txt_f, vp, tp = net(images, captions)
    with RL.in_reference_cwd():
        out = net.cal_loss(img_f, txt_f, vp, tp)
    sum(out["loss"].values()).backward()
    logits = net.logit_scale.exp() * img_f @ txt_f.t()
    return {
        "img_f": img_f.detach().clone(), "txt_f": txt_f.detach().clone(), "logits": logits.detach().clone(),
        "losses": {k: float(v.detach()) for k, v in out["loss"].items()},
        "grads": {k: getattr(net.prompts[net.numtask - 1], k).grad.detach().clone() for k in FACTORS},
    }


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    sd = S.make_clip_state_dict(0)
    fac0, fac1 = S.make_prompt_factors(0), S.make_prompt_factors(1)
    net = RL.build_reference_slinet(sd, {0: fac0, 1: fac1}, numtask=1)

    B = 4
    images, captions = S.make_images(B, 0), S.make_captions(B, 0)
    tokens = RL.reference_tokenize(captions)
    g = {"meta": {"B": B, "weights_seed": 0, "factor_seeds": [0, 1], "image_seed": 0, "caption_seed": 0},
         "captions": captions, "tokens": tokens}

    # --- config 1: train step, task 1 (base + alignment loss) ---------------------------------
    g["step_task1"] = train_step_case(net, images, captions)

    # --- prompts themselves -------------------------------------------------------------------
    with torch.no_grad():
        vis, txt = net.prompts[0]()
    g["prompt0"] = {"vis_sum": float(vis.double().sum()), "txt_sum": float(txt.double().sum()),
                    "vis_l0": vis[0].clone(), "txt_l0": txt[0].clone()}

    # --- eval-path features (a10-a12) -----------------------------------------------------------
    net.eval()
    with torch.no_grad():
        g["extract_vector"] = net.extract_vector(images).clone()
        g["extract_textual_vector"] = net.extract_textual_vector(captions).clone()
        net.numtask = 2
        cat = torch.tensor([0, 1, 1, 0])
        g["visual_interface"] = net.visual_interface(images, cat).clone()
        g["textual_interface"] = net.textual_interface(captions, cat).clone()
        g["interface_cat"] = cat

    # --- train step, task 2 (adds the task loss, slinet.py:160-183) -----------------------------
    net.numtask = 2
    g["step_task2"] = train_step_case(net, images, captions)
    torch.save(g, os.path.join(OUT, "model_b4_seed0.pt"))
    print("losses task1", g["step_task1"]["losses"], "task2", g["step_task2"]["losses"])

    # --- config 2: Recall@K on the Flickr30K-shaped set (a16) -----------------------------------
    ns = RL.load_reference()
    img, txt, img2txt, txt2img, cat_i, cat_t = S.make_retrieval_set()
    s_i2t = (img @ txt.t()).numpy()
    s_t2i = np.ascontiguousarray(s_i2t.T)
    res = ns.sprompt.SPrompts.itm_eval(types.SimpleNamespace(cur_id=4), s_i2t, s_t2i, txt2img, img2txt, cat_i,
                                       np.asarray(cat_t))
    torch.save({"result": res, "meta": {"n_img": 1000, "caps_per_img": 5, "seed": 2, "n_tasks": 5}},
               os.path.join(OUT, "recall_flickr_seed2.pt"))
    print("recall", res)

    # --- small ragged case: 37 images x 3 captions, 3 tasks ---------------------------------------
    img, txt, img2txt, txt2img, cat_i, cat_t = S.make_retrieval_set(37, 3, 512, 3, seed=5, signal=0.1)
    s_i2t = (img @ txt.t()).numpy()
    res = ns.sprompt.SPrompts.itm_eval(types.SimpleNamespace(cur_id=2), s_i2t, np.ascontiguousarray(s_i2t.T), txt2img,
                                       img2txt, cat_i, np.asarray(cat_t))
    torch.save({"result": res, "meta": {"n_img": 37, "caps_per_img": 3, "seed": 5, "n_tasks": 3, "signal": 0.1}},
               os.path.join(OUT, "recall_small_seed5.pt"))

    # --- losses on small random inputs (a8, a9) ---------------------------------------------------
    gen = torch.Generator().manual_seed(77)
    logits = torch.randn(9, 9, generator=gen) * 3
    x = torch.randn(4, 64, generator=gen)
    tgt = (torch.rand(4, 4, generator=gen) > 0.5).int()
    tgt = ((tgt + tgt.t()) > 0).int()
    tgt.fill_diagonal_(1)
    tgt[0, 1] = tgt[1, 0] = 0                                   # every row keeps >=1 negative
    torch.save({"logits": logits, "clip_loss": float(ns.loss.ClipLoss()(logits)), "x": x, "target": tgt,
                "nt_bxent": float(ns.loss.nt_bxent_loss(x, tgt, 0.001))}, os.path.join(OUT, "losses_seed77.pt"))


if __name__ == "__main__":
    main()
