"""SPrompts -- the LPI learner with the reference's surface (retrieval/methods/sprompt.py:104-646 + the BaseLearner fields
it uses, retrieval/methods/base.py:14-28): incremental_train / after_task / _train / train_function / clustering /
get_visual_task_id / get_textual_task_id / _evaluate_retrieval / itm_eval, plus gather_features (sprompt.py:38-82).

What runs where: encoders, losses, backward, SGD step, task-id selection, similarity + top-k + Recall@K are CUDA kernels
(liblpi_b200.so); K-Means for the task keys stays on the host with sklearn (random_state=0) for parity; data
loading is the caller's business -- pass loaders (any Dataset yielding the reference's item tuples) instead of COCO paths.
Multi-GPU: one process per GPU; pass a torch.distributed group via args['group'] -- the batch is sharded, features are
all-gathered once, the 21 KB prompt gradient is all-reduced once (the reference only supports a single GPU, README.md:13).
"""
from __future__ import annotations

import collections
import json
import logging
import os
from datetime import datetime
from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import optim

from . import lpi_step, ops, retrieval
from ._lib import LpiError
from .loss import ClipLoss
from .slinet import SliNet, load_task_sim_matrix
from . import losses as L


def gather_features(image_features, text_features, local_loss=False, gather_with_grad=False, rank=0, world_size=1, use_horovod=False):
    """sprompt.py:38-82 (open_clip style): all-gather both feature sets; without gather_with_grad the local slot keeps its own
    (grad-carrying) tensors.  Dead code in the reference (never called); here it is the documented DP entry point."""
    import torch.distributed as dist

    if use_horovod:
        raise LpiError("horovod is not supported; use torch.distributed (NCCL)")
    if world_size == 1:
        return image_features, text_features
    if gather_with_grad:
        import torch.distributed.nn as dist_nn       # (a bare `import torch.distributed.nn` here would shadow the module-level `torch`)

        all_image_features = torch.cat(dist_nn.all_gather(image_features), dim=0)
        all_text_features = torch.cat(dist_nn.all_gather(text_features), dim=0)
    else:
        gi = [torch.zeros_like(image_features) for _ in range(world_size)]
        gt = [torch.zeros_like(text_features) for _ in range(world_size)]
        dist.all_gather(gi, image_features)
        dist.all_gather(gt, text_features)
        if not local_loss:
            gi[rank] = image_features
            gt[rank] = text_features
        all_image_features = torch.cat(gi, dim=0)
        all_text_features = torch.cat(gt, dim=0)
    return all_image_features, all_text_features


def gather_rows(t: torch.Tensor, group) -> torch.Tensor:
    """All ranks' rows of a [n_r, d] tensor concatenated rank-major on every rank; n_r may differ per rank (ragged last shard of a
    loader).  Backend-agnostic (NCCL on the box, gloo in the CPU tests)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if world == 1:
        return t
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x) for x in sizes]
    m = max(sizes)
    pad = torch.zeros(m, *t.shape[1:], device=t.device, dtype=t.dtype)
    pad[: t.shape[0]] = t
    out = torch.empty(world * m, *t.shape[1:], device=t.device, dtype=t.dtype)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return torch.cat([out[r * m: r * m + sizes[r]] for r in range(world)], dim=0)


def is_main_rank(group) -> bool:
    """True on the one rank that writes files (checkpoints, ./res/*.json) in a data-parallel run."""
    if group is None:
        return True
    import torch.distributed as dist

    return dist.get_rank(group) == 0


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class FusedSGD(optim.Optimizer):
    """torch.optim.SGD(momentum, weight_decay) semantics (no dampening / nesterov, sprompt.py:253) as one kernel per tensor.
    Parameters without a gradient are skipped, exactly like torch (the reference hands SGD all 149.8 M parameters of which
    5 284 ever receive a gradient)."""

    def __init__(self, params, lr, momentum=0.9, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                first = "momentum_buffer" not in st
                if first:
                    st["momentum_buffer"] = torch.zeros_like(p)
                ops.sgd_momentum_step(p.data, p.grad.contiguous(), st["momentum_buffer"], group["lr"], group["momentum"],
                                      group["weight_decay"], first)


class _CapturedStep:
    """A closure over static input buffers captured into a CUDA graph (all graphs of a learner share one memory pool: they replay one
    at a time).  Calling it copies the new inputs into the captured buffers and replays; the returned tensors are the captured outputs."""

    def __init__(self, fn, static_inputs: dict, owner: "SPrompts"):
        self.inputs = static_inputs
        if owner._graph_pool is None:
            owner._graph_pool = torch.cuda.graph_pool_handle()
        torch.cuda.synchronize()
        n0 = ops.KERNEL_LAUNCHES
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=owner._graph_pool):
            self.out = fn(**self.inputs)
        self.launches = ops.KERNEL_LAUNCHES - n0

    def __call__(self, **new_inputs):
        for k, v in new_inputs.items():
            self.inputs[k].copy_(v, non_blocking=True)
        self.graph.replay()
        ops._count(self.launches)
        return self.out


class SPrompts(object):
    def __init__(self, args):
        # BaseLearner fields (base.py:14-28)
        self._cur_task = -1
        self._known_classes = 0
        self._total_classes = 0
        self._old_network = None
        self._device = args["device"][0] if isinstance(args["device"], (list, tuple)) else args["device"]
        self._multiple_gpus = args["device"] if isinstance(args["device"], (list, tuple)) else [args["device"]]
        if args["net_type"] != "slip":
            raise ValueError("Unknown net: {}.".format(args["net_type"]))
        self._network = SliNet(args)
        self.args = args
        self.EPSILON = args["EPSILON"]
        self.init_epoch, self.init_lr = args["init_epoch"], args["init_lr"]
        self.init_lr_decay, self.init_weight_decay = args["init_lr_decay"], args["init_weight_decay"]
        self.epochs, self.lrate, self.lrate_decay = args["epochs"], args["lrate"], args["lrate_decay"]
        self.batch_size, self.weight_decay, self.num_workers = args["batch_size"], args["weight_decay"], args["num_workers"]
        self.topk = 2
        self.class_num = self._network.class_num
        self.all_keys: List[torch.Tensor] = []
        self.textual_all_keys: List[torch.Tensor] = []
        self.loss = ClipLoss()
        self.cur_id = 0
        self.group = args.get("group")                    # torch.distributed process group for data-parallel training
        self.fused_step = bool(args.get("fused_step", True))
        # opt-in: replay the fused step from CUDA graphs (one per batch shape / text-length bucket / learning rate, see _step_fused)
        self.graph_step = bool(args.get("graph_step", False))
        self._graphs: dict = {}
        self._graph_pool = None
        self._task_consts: dict = {}
        self.n_tasks = int(args.get("n_tasks", args["total_sessions"]))
        self.log = logging.getLogger("lpi_b200")

    # ------------------------------------------------------------------ task loop
    def after_task(self):
        """sprompt.py:145-148 deep-copies the whole 149.8 M-parameter network and never uses it; only the
        bookkeeping is kept."""
        self._known_classes = self._total_classes

    def incremental_train(self, task_loaders: Optional[Sequence] = None):
        """sprompt.py:150-187.  task_loaders[i] = (train_loader, test_loader) for task i; the test loader covers tasks 0..i.
        Returns {task: result dict} and writes ./res/<datetime>.json like the reference."""
        task_loaders = task_loaders if task_loaders is not None else self.args.get("task_loaders")
        if task_loaders is None:
            raise LpiError("pass task_loaders=[(train_loader, test_loader), ...]: the COCO file datasets of the reference "
                           "(utils/data.py) are outside the hot path (SURVEY.md section 8(f) f4)")
        final_res = {}
        start = 0
        resume = self.args.get("resume_from")
        if resume:
            final_res = self.load_checkpoint(resume)
            start = self.cur_id + 1
            self.log.info("resumed after task %d from %s", self.cur_id, resume)
        ckpt_dir = self.args.get("checkpoint_dir")
        for i in range(start, min(self.n_tasks, len(task_loaders))):
            self._cur_task = [i]
            self.cur_id = i
            self._network.update_fc(self._total_classes)
            self.train_loader, self.test_loader = task_loaders[i]
            final_res[i] = self._train(self.train_loader, self.test_loader)
            if ckpt_dir:
                os.makedirs(ckpt_dir, exist_ok=True)
                self.save_checkpoint(os.path.join(ckpt_dir, f"task_{i}.pt"), final_res)
        os.makedirs("./res", exist_ok=True)
        self.save_dict(final_res, f"./res/{datetime.now()}.json")
        return final_res

    # ------------------------------------------------------------------ per-task checkpoint / resume
    def checkpoint_state(self, results: Optional[dict] = None) -> dict:
        """Everything a continual run needs to continue after task `cur_id`: the trainable state (every `prompts.{t}.*` factor and
        the per-task text contexts -- a few hundred KB; CLIP itself is frozen and reloaded from its own state_dict), the K-Means
        task keys (sprompt.py:396-397) and the results so far.  The reference's BaseLearner.save_checkpoint (base.py:57-63) is
        never called and would pickle all 149.8 M parameters."""
        sd = self._network.state_dict()
        keep = {k: v.detach().cpu().clone() for k, v in sd.items()
                if k.startswith("prompts.") or (k.startswith("classifier_pool.") and ".clip_model." not in k)}
        return {"format": "lpi_b200.checkpoint.v1", "cur_id": int(self.cur_id), "numtask": int(self._network.numtask),
                "known_classes": int(self._known_classes), "total_classes": int(self._total_classes), "trainable": keep,
                "all_keys": [k.detach().cpu() for k in self.all_keys],
                "textual_all_keys": [k.detach().cpu() for k in self.textual_all_keys],
                "results": {int(t): r for t, r in (results or {}).items()}}

    def save_checkpoint(self, path: str, results: Optional[dict] = None) -> str:
        """Data-parallel runs: every rank holds the same state (replicated prompts, gathered task keys), rank 0 alone writes, everybody
        leaves together -- concurrent ranks would otherwise race on the shared `.tmp` file."""
        if is_main_rank(self.group):
            tmp = path + ".tmp"
            torch.save(self.checkpoint_state(results), tmp)
            os.replace(tmp, path)                          # a crash mid-write never leaves a truncated checkpoint behind
        self._barrier()
        return path

    def _barrier(self):
        if self.group is not None:
            import torch.distributed as dist

            dist.barrier(group=self.group)

    def load_checkpoint(self, path: str) -> dict:
        """Restores prompts / contexts, task keys and counters; returns the results recorded so far ({task: result dict})."""
        st = torch.load(path, map_location="cpu", weights_only=True)    # v1 holds tensors, ints, lists and dicts of floats only
        if st.get("format") != "lpi_b200.checkpoint.v1":
            raise LpiError(f"{path} is not an lpi_b200 checkpoint")
        own = self._network.state_dict()
        missing = [k for k in st["trainable"] if k not in own]
        if missing:
            raise LpiError(f"checkpoint keys not in this model: {missing[:3]}...")
        with torch.no_grad():
            for k, v in st["trainable"].items():
                own[k].copy_(v.to(own[k].device, own[k].dtype))
        self._task_consts.clear()
        self._graphs.clear()
        self._network.numtask = st["numtask"]
        self.cur_id = st["cur_id"]
        self._cur_task = [self.cur_id]
        self._known_classes, self._total_classes = st["known_classes"], st["total_classes"]
        self.all_keys = [k.to(self._device) for k in st["all_keys"]]
        self.textual_all_keys = [k.to(self._device) for k in st["textual_all_keys"]]
        return {int(t): r for t, r in st["results"].items()}

    def save_dict(self, dictionary, file_path):
        if is_main_rank(self.group):
            with open(file_path, "w") as f:
                json.dump(dictionary, f)
        self._barrier()

    def _train(self, train_loader, test_loader):
        self._task_consts.clear()                              # per-task caches of the fused step (earlier prompts may have been reloaded)
        self._graphs.clear()
        self._network.to(self._device)
        network = self._network
        for name, param in network.named_parameters():          # freeze all but "prompts.{t}." (sprompt.py:229-237)
            param.requires_grad_(False)
            if "prompts" + "." + str(network.numtask - 1) + "." in name:
                param.requires_grad_(True)
        optimizer = FusedSGD(self._network.parameters(), momentum=0.9, lr=self.lrate, weight_decay=self.weight_decay)
        scheduler = optim.lr_scheduler.CosineAnnealingLR(optimizer=optimizer, T_max=self.epochs)
        self.run_epoch = self.epochs
        return self.train_function(train_loader, test_loader, optimizer, scheduler)

    # ------------------------------------------------------------------ hot loop
    def _step_autograd(self, images, captions, optimizer):
        image_features, text_features, visual_prompt, textual_prompt = self._network(images, captions)
        model_out = self._network.cal_loss(image_features, text_features, visual_prompt, textual_prompt)
        loss = sum(l for l in model_out["loss"].values())
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        return model_out["loss"]

    def _task_constants(self, net, device):
        """Per-task constants of the fused step, computed once per task instead of once per step: exp(logit_scale) (a device -> host
        read), the reconstructed prompts of the frozen earlier tasks, the task-similarity targets (slinet.py:167-183)."""
        key = (net.numtask, str(device))
        c = self._task_consts.get(key)
        if c is None:
            t = net.numtask - 1
            prev, target = [], None
            if net.numtask != 1:
                with torch.no_grad():
                    prev = [net.prompts[i]() for i in range(t)]
                sim = load_task_sim_matrix() if net._task_sim is None else net._task_sim
                net._task_sim = sim
                target = torch.tensor((sim[:t + 1, :t + 1] > L.TASK_THRESHOLD).astype(np.int32), device=device)
            c = self._task_consts[key] = (float(net.logit_scale.exp()), prev, target)
        return c

    def _step_fused(self, images, captions, optimizer):
        """Same maths as _step_autograd without the autograd graph; supports the data-parallel group.  With args['graph_step'] the
        step is replayed from a CUDA graph from the second occurrence of a (batch shape, text-length bucket, learning rate) on."""
        net = self._network
        t = net.numtask - 1
        prompt = net.prompts[t]
        factors = {k: getattr(prompt, k).data for k in lpi_step.FACTOR_NAMES}
        tokens = captions if isinstance(captions, torch.Tensor) else net.classifier_pool[t].tokenize(captions)
        scale, prev, target = self._task_constants(net, images.device)
        vision, text = net.image_encoder.engine(), net.clip_model.text_engine()
        inject = tuple(net.clip_model.inject_layers)
        text_len = getattr(tokens, "lpi_text_len", None)

        def run(images, tokens, text_len):
            r = lpi_step.train_step(vision, text, factors, images, tokens, scale, prev, target, inject, self.group, text_len=text_len,
                                    prompt_scale=float(prompt.scale))
            for k in lpi_step.FACTOR_NAMES:
                getattr(prompt, k).grad = r["grads"][k]
            optimizer.step()
            return {k: v.view(()) for k, v in r["losses"].items()}

        images = images.float()
        if self.graph_step:
            # a longer text_len is still exact, so lengths are bucketed to multiples of 8 to bound the number of graphs
            tl = tokens.shape[1] if text_len is None else min(tokens.shape[1], -(-int(text_len) // 8) * 8)
            lr = float(optimizer.param_groups[0]["lr"])
            key = (net.numtask, tuple(images.shape), tuple(tokens.shape), tl, lr)
            slot = self._graphs.get(key)
            if slot is None:                                   # first occurrence: eager (it is also the warm-up of every lazy one-time setup)
                for k in [k for k in self._graphs if k[0] != net.numtask or k[-1] != lr]:
                    del self._graphs[k]                        # graphs of earlier tasks / learning rates are dead
                self._graphs[key] = "seen"
            else:
                if slot == "seen":
                    slot = self._graphs[key] = _CapturedStep(lambda images, tokens: run(images, tokens, tl),
                                                             dict(images=images.clone(), tokens=tokens.clone()), self)
                return slot(images=images, tokens=tokens)
        optimizer.zero_grad()
        return run(images, tokens, text_len)

    def train_function(self, train_loader, test_loader, optimizer, scheduler):
        """sprompt.py:290-334"""
        loss_meter = collections.defaultdict(AverageMeter)
        for epoch in range(self.run_epoch):
            self._network.train()
            for i, (images, captions, _, _) in enumerate(train_loader):
                images = images.to(self._device, non_blocking=True)
                captions = captions if isinstance(captions, torch.Tensor) else list(captions)
                if isinstance(captions, torch.Tensor):
                    captions = captions.to(self._device, non_blocking=True)
                step = self._step_fused if (self.fused_step or self.group is not None) else self._step_autograd
                out = step(images, captions, optimizer)
                for k, v in out.items():
                    loss_meter[k].update(v.detach())              # stays on the device: no per-step sync
                if i % 50 == 0:
                    info = "Task {}, Epoch {}/{}, Batch {}, lr {:.4f} =>, ".format(self.cur_id, epoch + 1, self.run_epoch, i,
                                                                                   optimizer.param_groups[0]["lr"])
                    for k, v in loss_meter.items():
                        info += "{} = {:.4f}, ".format(k, float(v.avg))
                        v.reset()
                    self.log.info(info)
            scheduler.step()
        self.clustering(dataloader=train_loader)
        _, _, final_res = self._evaluate_retrieval(test_loader)
        return final_res

    # ------------------------------------------------------------------ task keys / task-id
    def _task_id(self, feature, keys):
        return ops.nearest_center_l1(feature.float().contiguous(), torch.stack([k.float() for k in keys]).contiguous())

    def get_visual_task_id(self, inputs):
        """sprompt.py:336-351: argmin over tasks of the min L1 distance to the task's 5 centres."""
        with torch.no_grad():
            return self._task_id(self._network.extract_vector(inputs), self.all_keys)

    def get_textual_task_id(self, inputs):
        with torch.no_grad():
            return self._task_id(self._network.extract_textual_vector(inputs), self.textual_all_keys)

    def clustering(self, dataloader):
        """sprompt.py:370-397: un-prompted features of the task's train set -> KMeans(5, random_state=0) centres (host, sklearn)."""
        from sklearn.cluster import KMeans

        vf, tf = [], []
        with torch.no_grad():
            for inputs, captions, _, _ in dataloader:
                inputs = inputs.to(self._device)
                captions = captions.to(self._device) if isinstance(captions, torch.Tensor) else list(captions)
                vf.append(self._network.extract_vector(inputs))
                tf.append(self._network.extract_textual_vector(captions))
        vf = torch.cat(vf, 0)
        tf = torch.cat(tf, 0)
        if self.group is not None:
            # data-parallel: each rank saw its own shard of the task's train set; the keys must come from ALL of it and be identical on
            # every rank (they drive the task-id selection at evaluation time), so the features are gathered (rank-major) ...
            vf, tf = gather_rows(vf.contiguous(), self.group), gather_rows(tf.contiguous(), self.group)
        vf = ops.l2_normalize(vf.contiguous()).cpu().numpy()       # the reference re-normalises (sprompt.py:387-390)
        tf = ops.l2_normalize(tf.contiguous()).cpu().numpy()
        vc = KMeans(n_clusters=5, random_state=0).fit(vf)
        tc = KMeans(n_clusters=5, random_state=0).fit(tf)
        keys_v = torch.tensor(vc.cluster_centers_).to(self._device)
        keys_t = torch.tensor(tc.cluster_centers_).to(self._device)
        if self.group is not None:
            import torch.distributed as dist

            # ... and rank 0's centres are the ones everybody keeps (host K-Means is deterministic for identical inputs, but a different
            # BLAS / thread count on one rank must not be able to split the ranks)
            src = dist.get_global_rank(self.group, 0)
            dist.broadcast(keys_v, src=src, group=self.group)
            dist.broadcast(keys_t, src=src, group=self.group)
        self.all_keys.append(keys_v)
        self.textual_all_keys.append(keys_t)

    # ------------------------------------------------------------------ evaluation
    @torch.no_grad()
    def _evaluate_retrieval(self, data_loader, return_scores: Optional[bool] = None):
        """sprompt.py:433-548 -> (score_matrix_i2t, score_matrix_t2i, final_res).  The dense matrices are only materialised
        (numpy, like the reference) when small (<= 5e7 entries) or when return_scores=True; Recall@K itself never needs them."""
        self._network.eval()
        ds = data_loader.dataset
        texts = ds.text
        texts_cat = torch.as_tensor(ds.text_cat)
        image_feats, category_i = [], []
        for image, img_id, category in data_loader:
            image = image.to(self._device)
            if self.args["prompt_type"] == "clip":
                feat = self._network.extract_vector(image)
            else:
                # get_visual_task_id + visual_interface with the patch embedding shared by both ViT passes and the selection produced by
                # the un-prompted pass's head kernel
                feat, _ = self._network.visual_select_and_encode(image, self.all_keys)
            image_feats.append(feat)
            category_i.extend(int(z) for z in category)
        image_feats = torch.cat(image_feats)
        text_feats = []
        text_bs = 256
        num_text = len(texts)
        for i in range(0, num_text, text_bs):
            text = texts[i:min(num_text, i + text_bs)]
            if isinstance(text, torch.Tensor):
                text = text.to(self._device)
            if self.args["prompt_type"] == "clip":
                tfeat = self._network.extract_textual_vector(text)
            else:
                tsel = self.get_textual_task_id(text)
                tfeat = self._network.textual_interface(text, tsel)
            text_feats.append(tfeat)
        text_feats = torch.cat(text_feats)
        final_res = retrieval.itm_eval_features(image_feats, text_feats, ds.txt2img, ds.img2txt, category_i, texts_cat.tolist(),
                                                self.cur_id + 1, precision="fp32")
        n = image_feats.shape[0] * text_feats.shape[0]
        if return_scores is None:
            return_scores = n <= 50_000_000
        if return_scores:
            s = ops.sgemm(image_feats, text_feats.t())
            s_i2t = s.cpu().numpy()
            return s_i2t, np.ascontiguousarray(s_i2t.T), final_res
        return None, None, final_res

    @torch.no_grad()
    def itm_eval(self, scores_i2t, scores_t2i, txt2img, img2txt, category_i, category_t):
        """sprompt.py:550-646 on dense score matrices (drop-in signature)."""
        return retrieval.itm_eval(scores_i2t, scores_t2i, txt2img, img2txt, category_i, category_t, self.cur_id + 1, device=self._device)
