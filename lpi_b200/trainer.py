"""trainer.train(args) with the reference's signature (retrieval/trainer.py:13-90): per-seed loop, fixed torch seeds,
device list -> torch.device, build the learner, incremental_train + after_task.  `args` is the flat dict of
configs/lpi/coco_lpi.json plus `task_loaders` (the data layer is the caller's)."""
from __future__ import annotations

import copy
import logging
import sys

import torch

from . import factory


def train(args):
    seed_list = copy.deepcopy(args["seed"])
    device = copy.deepcopy(args["device"])
    out = []
    for seed in seed_list:
        args["seed"] = seed
        args["device"] = device
        out.append(_train(args))
    return out


def _train(args):
    logging.basicConfig(level=logging.INFO, format="%(asctime)s [%(filename)s] => %(message)s", handlers=[logging.StreamHandler(sys.stdout)])
    _set_random()
    _set_device(args)
    model = factory.get_model(args["model_name"], args)
    n_all = sum(p.numel() for p in model._network.parameters())
    logging.info("All params: {}".format(n_all))
    res = model.incremental_train()
    model.after_task()
    return res


def _set_device(args):
    gpus = []
    for device in args["device"]:
        gpus.append(device if isinstance(device, torch.device) else torch.device("cuda:{}".format(device)))
    args["device"] = gpus


def _set_random():
    torch.manual_seed(1)
    torch.cuda.manual_seed_all(1)
