"""BASELINE.json configs[3]: the 5-task continual retrieval sequence on synthetic data (per-task prompt growth, K-Means task keys,
per-task Recall@K evaluation over the tasks seen so far), with per-task checkpoints and the reshandle summary.
python tools/run_continual.py [--tasks 5] [--train 256] [--epochs 2] [--eval-images 100]"""
import argparse, json, os, sys, tempfile, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpi_b200 import data as D, ops, reshandle as RH, synthetic as S  # noqa: E402
from lpi_b200.config import default_args  # noqa: E402
from lpi_b200.sprompt import SPrompts  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=5)
    ap.add_argument("--train", type=int, default=256)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--eval-images", type=int, default=100)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--graph", action="store_true", help="replay the fused step from CUDA graphs (args['graph_step'])")
    a = ap.parse_args()
    work = tempfile.mkdtemp(prefix="lpi_continual_")
    os.chdir(work)
    torch.manual_seed(0)                               # the prompt factors are drawn from torch's global generator (prompts.py:21-25)
    sd = S.make_clip_state_dict(0)
    args = default_args(clip_state_dict=sd, device=[torch.device("cuda")], epochs=a.epochs, batch_size=a.batch, n_tasks=a.tasks,
                        checkpoint_dir=os.path.join(work, "ckpt"), graph_step=a.graph)
    learner = SPrompts(args)
    loaders = D.make_task_loaders(a.tasks, a.train, a.eval_images, 5, a.batch, 128)
    torch.cuda.synchronize()
    n0, t0 = ops.KERNEL_LAUNCHES, time.perf_counter()
    res = learner.incremental_train(loaders)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"config": "5-task continual retrieval sequence (BASELINE.json configs[3])", "tasks": a.tasks, "train_pairs_per_task": a.train,
           "epochs": a.epochs, "graph_step": a.graph, "eval_images_per_task": a.eval_images, "captions_per_image": 5, "wall_s": dt,
           "lpi_kernel_launches": ops.KERNEL_LAUNCHES - n0,
           "final_session": {side: res[a.tasks - 1]["mscoco"][side] for side in ("i2t", "t2i")},
           "summary_i2t": RH.summarize(res, "mscoco", "i2t"), "summary_t2i": RH.summarize(res, "mscoco", "t2i"),
           "checkpoints": sorted(os.listdir(os.path.join(work, "ckpt")))}
    for s in ("summary_i2t", "summary_t2i"):
        out[s] = {k: out[s][k] for k in ("avg_recall", "forgetting", "avg_forgetting")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
