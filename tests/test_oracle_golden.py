"""The oracle restatement (oracle/lpi_oracle.py) against fixtures generated from the REAL reference
(tests/golden/make_golden.py).  CPU only; this is what pins the oracle on every box."""
import os

import numpy as np
import pytest
import torch

from lpi_b200 import synthetic as S
from oracle import lpi_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def test_losses_golden():
    g = torch.load(os.path.join(GOLDEN, "losses_seed77.pt"), weights_only=False)
    assert abs(float(O.clip_loss(g["logits"])) - g["clip_loss"]) < 1e-6
    assert abs(float(O.nt_bxent_loss(g["x"], g["target"], 0.001)) - g["nt_bxent"]) < 1e-6


@pytest.mark.parametrize("name", ["recall_flickr_seed2.pt", "recall_small_seed5.pt"])
def test_itm_eval_golden(name):
    g = torch.load(os.path.join(GOLDEN, name), weights_only=False)
    m = g["meta"]
    img, txt, img2txt, txt2img, cat_i, cat_t = S.make_retrieval_set(m["n_img"], m["caps_per_img"], 512, m["n_tasks"],
                                                                    seed=m["seed"], signal=m.get("signal", 0.15))
    s = (img @ txt.t()).numpy()
    res = O.itm_eval(s, np.ascontiguousarray(s.T), txt2img, img2txt, cat_i, cat_t, m["n_tasks"])
    assert res == g["result"]          # Recall@K bit-exact (Python floats)
    # the top-k formulation gives the same recall as the rank formulation
    _, idx = O.topk_lowest_index(np.ascontiguousarray(s.T), 10)
    hit1 = np.mean([txt2img[t] == idx[t, 0] for t in range(s.shape[1])])
    want = np.mean([g["result"]["mscoco"]["t2i"][c][0] for c in cat_t]) / 100.0
    assert abs(hit1 - want) < 1e-12


def test_prompt_golden(golden_model):
    fac = S.make_prompt_factors(0)
    vis, txt = O.decomposed_prompt(*[fac[k] for k in O.FACTOR_NAMES])
    assert vis.shape == (9, 16, 768) and txt.shape == (9, 16, 512)
    assert torch.allclose(vis[0], golden_model["prompt0"]["vis_l0"], atol=1e-7)
    assert torch.allclose(txt[0], golden_model["prompt0"]["txt_l0"], atol=1e-7)


def test_train_step_golden(golden_model, clip_sd):
    g = golden_model
    images = S.make_images(g["meta"]["B"], 0)
    r = O.train_step(clip_sd, S.make_prompt_factors(0), images, g["tokens"])
    want = g["step_task1"]
    assert _rel(r["img_f"], want["img_f"]) < 1e-5 and _rel(r["txt_f"], want["txt_f"]) < 1e-5
    for k, v in want["losses"].items():
        assert abs(r["losses"][k] - v) < 1e-5 * max(1.0, abs(v))
    for k in O.FACTOR_NAMES:
        assert _rel(r["grads"][k], want["grads"][k]) < 1e-4, k


def test_train_step_task2_golden(golden_model, clip_sd):
    g = golden_model
    images = S.make_images(g["meta"]["B"], 0)
    sim = np.loadtxt(os.path.join(os.path.dirname(GOLDEN), "..", "lpi_b200", "MID", "task_sim_matrix.txt"))
    r = O.train_step(clip_sd, S.make_prompt_factors(1), images, g["tokens"], prev_factors=[S.make_prompt_factors(0)],
                     task_sim=sim)
    want = g["step_task2"]
    assert set(r["losses"]) == {"base_loss", "alignment_loss", "task_loss"}
    for k, v in want["losses"].items():
        assert abs(r["losses"][k] - v) < 1e-5 * max(1.0, abs(v)), k
    for k in O.FACTOR_NAMES:
        assert _rel(r["grads"][k], want["grads"][k]) < 1e-4, k


def test_eval_features_golden(golden_model, clip_sd):
    g = golden_model
    images = S.make_images(g["meta"]["B"], 0)
    with torch.no_grad():
        f = O.l2_normalize(O.vision_forward(clip_sd, images, None))
        assert _rel(f, g["extract_vector"]) < 1e-5
        t = O.l2_normalize(O.text_forward(clip_sd, g["tokens"], None))
        assert _rel(t, g["extract_textual_vector"]) < 1e-5
        cat = g["interface_cat"]
        protos = [O.decomposed_prompt(*[S.make_prompt_factors(s)[k] for k in O.FACTOR_NAMES]) for s in (0, 1)]
        vis = torch.stack([protos[int(c)][0] for c in cat])
        f = O.l2_normalize(O.vision_forward(clip_sd, images, vis))
        assert _rel(f, g["visual_interface"]) < 1e-5
        ctx = torch.stack([protos[int(c)][1][0] for c in cat])
        t = O.l2_normalize(O.text_forward(clip_sd, g["tokens"], ctx))
        assert _rel(t, g["textual_interface"]) < 1e-5


def test_nearest_task_and_sgd():
    gen = torch.Generator().manual_seed(3)
    f = torch.randn(6, 8, generator=gen)
    keys = [torch.randn(5, 8, generator=gen) for _ in range(3)]
    sel = O.nearest_task_l1(f, keys)
    brute = torch.stack([torch.stack([(f - c).abs().sum(1) for c in k]).min(0)[0] for k in keys]).min(0)[1]
    assert torch.equal(sel, brute)
    # SGD against torch.optim.SGD (sprompt.py:253)
    w = torch.randn(7, 4, generator=gen)
    p = torch.nn.Parameter(w.clone())
    opt = torch.optim.SGD([p], lr=0.05, momentum=0.9, weight_decay=2e-4)
    buf, ww = None, w.clone()
    for _ in range(3):
        gr = torch.randn(7, 4, generator=gen)
        p.grad = gr.clone()
        opt.step()
        ww, buf = O.sgd_momentum_step(ww, gr, buf, 0.05)
    assert torch.allclose(ww, p.detach(), atol=1e-6)
