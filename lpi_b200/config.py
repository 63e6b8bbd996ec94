"""Flat-dict run configuration with the reference's keys (retrieval/configs/lpi/coco_lpi.json; argparse + JSON are merged into
one dict, retrieval/main.py:18-20).  Extra keys understood by this build: inject_layers, group, fused_step, n_tasks,
clip_state_dict, task_loaders, checkpoint_dir, resume_from, graph_step."""
from __future__ import annotations


def default_args(**overrides) -> dict:
    args = {
        # learner / model selection (only these values are implemented)
        "model_name": "sprompts", "net_type": "slip", "prompt_type": "lpi", "dataset": "Coco", "prefix": "reproduce",
        # prompt geometry: 16 prompts per layer, 9 prompt layers (hard-coded in slinet.py:46), widths 768 / 512, 12 sessions
        "prompt_length": 16, "visual_dim": 768, "textual_dim": 512, "embd_dim": 768, "total_sessions": 12,
        "prompt_depth": 3,            # read by nothing in the reference; see inject_layers
        "inject_layers": [],          # [] = reference as shipped; [1, 2] = the intended depth-3 additive injection
        # text prompt learner
        "backbonename": "ViT-B/16", "NCTX": 16, "CTXINIT": "", "CSC": False, "CLASS_TOKEN_POSITION": "end",
        # optimisation (sprompt.py:253-254)
        "epochs": 10, "lrate": 0.05, "lrate_decay": 0.1, "weight_decay": 2e-4, "batch_size": 64,
        "init_epoch": 10, "init_lr": 0.05, "init_lr_decay": 0.1, "init_weight_decay": 0.0005,
        # bookkeeping the BaseLearner reads
        "memory_size": 0, "memory_per_class": 0, "fixed_memory": True, "shuffle": False, "EPSILON": 1e-8, "num_workers": 8,
        "device": ["0"], "seed": [1993],
    }
    args.update(overrides)
    return args
