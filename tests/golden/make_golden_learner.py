"""Golden fixture for the whole-learner path (BASELINE.json configs[3], scaled down): the REAL reference SPrompts learner run
on CPU for 2 synthetic tasks (16 train pairs, batch 8, 2 epochs; 6 eval images x 2 captions per task).  Only the 12-task
`for` of incremental_train is restated (it hard-codes COCO files, sprompt.py:154-172); _train -> train_function -> clustering
-> _evaluate_retrieval -> itm_eval run untouched.     python tests/golden/make_golden_learner.py [n_tasks]  (build container only)
n_tasks = 2 (default) -> learner_2task_seed0.pt; n_tasks = 5 -> learner_5task_seed0.pt, the 5-task continual sequence of
BASELINE.json configs[3] (per-task prompt growth, task loss over 2..5 stacked prompts, evaluation over tasks 0..t after every task)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_loader as RL  # noqa: E402
from lpi_b200 import data as D, synthetic as S  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CFG = dict(n_tasks=2, n_train=16, n_eval_images=6, caps_per_image=2, batch_size=8, eval_batch_size=4, epochs=2)


def main():
    if len(sys.argv) > 1:
        CFG["n_tasks"] = int(sys.argv[1])
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ns = RL.load_reference()
    args = RL.reference_args()
    args["epochs"] = CFG["epochs"]
    args["batch_size"] = CFG["batch_size"]
    with RL.in_reference_cwd():
        learner = ns.sprompt.SPrompts(args)
    net = learner._network
    net.clip_model.load_state_dict(S.make_clip_state_dict(0))
    with torch.no_grad():
        for t in range(CFG["n_tasks"]):
            for k, v in S.make_prompt_factors(t).items():
                getattr(net.prompts[t], k).copy_(v)
    loaders = D.make_task_loaders(CFG["n_tasks"], CFG["n_train"], CFG["n_eval_images"], CFG["caps_per_image"], CFG["batch_size"],
                                  CFG["eval_batch_size"])
    g = {"cfg": CFG, "tasks": []}
    step_log = []
    orig_cal = net.cal_loss

    def logged(*a, **k):
        out = orig_cal(*a, **k)
        step_log.append({n: float(v.detach()) for n, v in out["loss"].items()})
        return out

    net.cal_loss = logged
    for t in range(CFG["n_tasks"]):
        learner.cur_id = t
        net.update_fc(0)
        step_log.clear()
        with RL.in_reference_cwd():
            res = learner._train(*loaders[t])
        ds = loaders[t][1].dataset
        with torch.no_grad():
            net.eval()
            imgs = torch.stack(ds.image)
            sel_i = learner.get_visual_task_id(imgs)
            f_i = net.visual_interface(imgs, sel_i)
            sel_t = learner.get_textual_task_id(ds.text)
            f_t = net.textual_interface(ds.text, sel_t)
        g["tasks"].append({
            "result": res, "losses": [dict(s) for s in step_log],
            "factors": {k: getattr(net.prompts[t], k).detach().clone() for k in S.FACTOR_NAMES},
            "keys_visual": learner.all_keys[t].clone(), "keys_textual": learner.textual_all_keys[t].clone(),
            "sel_i": sel_i.clone(), "sel_t": sel_t.clone(), "img_f": f_i.clone(), "txt_f": f_t.clone(),
        })
        print("task", t, res, step_log[-1])
    torch.save(g, os.path.join(OUT, f"learner_{CFG['n_tasks']}task_seed0.pt"))


if __name__ == "__main__":
    main()
