"""SliNet -- the prompted-CLIP wrapper of the LPI learner with the reference's surface
(retrieval/models/slinet.py:12-234): forward / cal_loss / cal_task_loss / visual_interface / textual_interface /
extract_vector / extract_textual_vector / update_fc / copy / freeze, same attribute and parameter names
(`prompts.{t}.dim_1_share`, `classifier_pool.{t}.ctx`, `clip_model.*`, ...), same loss dict.  Only `prompt_type: lpi`
(and the un-prompted `clip` evaluation path) is implemented -- the S-Prompts / L2P baselines are out of scope."""
from __future__ import annotations

import copy
import os
from typing import List, Sequence, Union

import numpy as np
import torch
import torch.nn as nn

from . import losses as L
from ._lib import LpiError
from .autograd import AlignmentLossFn, ContrastiveFn, TextEncodeFn, _as_table
from .clip import PromptLearner, TextEncoder, cfgc, load_clip_to_cpu
from .loss import ClipLoss, nt_bxent_loss
from .prompts import DecomposedPrompt

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_task_sim_matrix() -> np.ndarray:
    """`./MID/task_sim_matrix.txt` relative to the cwd like the reference (slinet.py:171), else the packaged copy; read ONCE
    (the reference re-reads it every step)."""
    for p in (os.path.join(".", "MID", "task_sim_matrix.txt"), os.path.join(_HERE, "MID", "task_sim_matrix.txt")):
        if os.path.isfile(p):
            return np.loadtxt(p)
    raise FileNotFoundError("MID/task_sim_matrix.txt")


class SliNet(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.cfg = cfgc()
        self.args = args
        self.cfg.backbonename = args["backbonename"]
        self.cfg.NCTX = args["NCTX"]
        self.cfg.CTXINIT = args["CTXINIT"]
        self.cfg.CSC = args["CSC"]
        self.cfg.CLASS_TOKEN_POSITION = args["CLASS_TOKEN_POSITION"]
        if args["prompt_type"] not in ("lpi", "clip"):
            raise LpiError(f"prompt_type {args['prompt_type']!r}: only 'lpi' (and 'clip' evaluation) are in scope")

        clip_model = load_clip_to_cpu(args)
        self.clip_model = clip_model
        self.image_encoder = clip_model.visual
        self.text_encoder = TextEncoder(clip_model)
        self.logit_scale = clip_model.logit_scale
        self.dtype = clip_model.visual.conv1.weight.dtype
        self.prompts = nn.ModuleList([DecomposedPrompt(9, args["prompt_length"], args["visual_dim"], args["textual_dim"])
                                      for _ in range(args["total_sessions"])])
        self.classifier_pool = nn.ModuleList([PromptLearner(self.cfg, self.clip_model) for _ in range(args["total_sessions"])])
        self.class_num = 2
        self.numtask = 0
        self.loss = ClipLoss()
        self.alignment_loss = ClipLoss()
        self.all_keys = []
        self._task_sim = None
        inject = tuple(args.get("inject_layers", ()))          # () = the reference as shipped (model.py:190 is dead code)
        clip_model.inject_layers = inject
        clip_model.visual.inject_layers = inject

    @property
    def device(self):
        return self.logit_scale.device

    @property
    def feature_dim(self):
        return self.image_encoder.output_dim

    # ------------------------------------------------------------------ un-prompted features (K-Means keys, task-id)
    def extract_vector(self, image):
        return self.image_encoder.encode(image.float(), None)[0]

    extract_visual_vector = extract_vector

    def extract_textual_vector(self, text):
        text_prompts, tokenized = self.classifier_pool[self.numtask - 1].extract_vector(text)
        return self.text_encoder.encode(text_prompts, tokenized, None)[0]

    # ------------------------------------------------------------------ training forward
    def forward(self, image, text):
        """slinet.py:109-135.  image [B,3,224,224]; text = list[str] (or pre-tokenised ids [B,77]).
        -> (image_features, text_features, visual_prompt_exp [B,9,16,768], textual_prompt_exp [B,9,16,512])"""
        visual_prompt, textual_prompt = self.prompts[self.numtask - 1]()
        bs = image.shape[0]
        visual_prompt_exp = visual_prompt.expand(bs, -1, -1, -1)
        image_features = self.image_encoder.encode(image.float(), visual_prompt_exp)[0]
        prompts = self.classifier_pool[self.numtask - 1]
        textual_prompt_exp = textual_prompt.expand(bs, -1, -1, -1)
        text_prompts, tokenized_prompts = prompts(text, textual_prompt_exp[:, 0])
        text_features = self.text_encoder.encode(text_prompts, tokenized_prompts, textual_prompt_exp)[0]
        return image_features, text_features, visual_prompt_exp, textual_prompt_exp

    # ------------------------------------------------------------------ losses
    def cal_loss(self, image_featuers, text_features, visual_prompt, textual_prompt):
        """slinet.py:137-165 -> {'loss': {'base_loss', 'alignment_loss'[, 'task_loss']}}"""
        losses = {"base_loss": ContrastiveFn.apply(image_featuers, text_features, float(self.logit_scale.exp()))}
        if self.args["prompt_type"] == "lpi":
            # the batch mean over an `expand`ed prompt is the identity (slinet.py:146-152): use one copy
            vis = visual_prompt[0] if visual_prompt.dim() == 4 else visual_prompt
            txt = textual_prompt[0] if textual_prompt.dim() == 4 else textual_prompt
            losses["alignment_loss"] = AlignmentLossFn.apply(vis, txt)           # already carries the 0.1 weight
            if self.numtask != 1:
                losses["task_loss"] = 0.1 * self.cal_task_loss(self.numtask - 1, None, None)
        return {"loss": losses}

    def cal_task_loss(self, task_id, visual_prompt, textual_prompt):
        """slinet.py:167-183: stack the flattened prompts of tasks 0..task_id (only the last row trains)."""
        if self._task_sim is None:
            self._task_sim = load_task_sim_matrix()
        dev = self.device
        target = torch.tensor((self._task_sim[:task_id + 1, :task_id + 1] > L.TASK_THRESHOLD).astype(np.int32), device=dev)
        vs, ts = [], []
        for i in range(task_id + 1):
            if i == task_id:
                v, t = self.prompts[i]()
            else:
                with torch.no_grad():
                    v, t = self.prompts[i]()
            vs.append(v.reshape(-1))
            ts.append(t.reshape(-1))
        return (nt_bxent_loss(torch.stack(vs), target, L.TASK_TEMPERATURE) + nt_bxent_loss(torch.stack(ts), target, L.TASK_TEMPERATURE)) / 2

    # ------------------------------------------------------------------ evaluation interfaces (per-sample task prompts)
    def _prompt_tables(self):
        with torch.no_grad():
            ps = [p() for p in self.prompts]
        return torch.stack([p[0] for p in ps]), torch.stack([p[1] for p in ps])        # [T,9,16,768], [T,9,16,512]

    def _factor_stacks(self):
        """The factors of ALL tasks stacked ([T, Lp, r], [T, P, r], [T, D, r] per modality) for the kernels that reconstruct prompt rows
        on the fly.  Rebuilt on every call (five tiny concatenations): the fused SGD kernel updates the factors in place through raw
        pointers, so parameter version counters cannot be trusted for caching."""
        with torch.no_grad():
            st = lambda name: torch.stack([getattr(pr, name).detach().float() for pr in self.prompts]).contiguous()
            out = {"d1": st("dim_1_share"), "d2v": st("dim_2_visual"), "d2t": st("dim_2_textual"), "d3v": st("dim_3_visual"),
                   "d3t": st("dim_3_textual")}
        scales = {float(pr.scale) for pr in self.prompts}
        out["scale"] = scales.pop() if len(scales) == 1 else None      # per-task scales differ: fall back to the tables
        return out

    def _fused_eval(self) -> bool:
        """Evaluation interfaces take the fused path (prompt rows reconstructed inside the assembly kernels, no [T,9,16,D] tables) unless
        autograd has to see the prompts or deep injection needs layers >= 1 of the tables."""
        return (not torch.is_grad_enabled()) and len(self.clip_model.inject_layers) == 0 and len({float(pr.scale) for pr in self.prompts}) == 1

    def visual_interface(self, image, image_category):
        """slinet.py:212-220: every sample uses the prompts of its (predicted) task."""
        sel = torch.as_tensor(image_category, device=image.device).to(torch.int32).contiguous()
        if self._fused_eval():
            f = self._factor_stacks()
            return self.image_encoder.engine().forward(image.float(), None, sel, factors=(f["d1"], f["d2v"], f["d3v"], f["scale"]))[0]
        vt, _ = self._prompt_tables()
        from .autograd import VisionEncodeFn
        return VisionEncodeFn.apply(self.image_encoder.engine(), image.float(), vt, sel, tuple(self.image_encoder.inject_layers))[0]

    @torch.no_grad()
    def visual_select_and_encode(self, image, task_keys):
        """get_visual_task_id + visual_interface of one image batch (sprompt.py:336-351, 456-470; slinet.py:212-220) as the evaluation
        loop needs them: the patch embedding is computed once for the un-prompted and the prompted ViT pass, the task-id selection comes
        out of the un-prompted pass's head kernel.  task_keys: list of [5, E] K-Means centres per task.
        -> (features [B, E], selection int64 [B])"""
        centers = torch.stack([k.float() for k in task_keys]).contiguous()
        eng = self.image_encoder.engine()
        if self._fused_eval():
            f = self._factor_stacks()
            feat, sel, _ = eng.select_and_encode(image.float(), centers, None, (f["d1"], f["d2v"], f["d3v"], f["scale"]))
        else:
            vt, _ = self._prompt_tables()
            feat, sel, _ = eng.select_and_encode(image.float(), centers, vt, None, tuple(self.image_encoder.inject_layers))
        return feat, sel.long()

    def textual_interface(self, text, text_category):
        """slinet.py:185-210.  Eval: one batched pass (the reference loops per sample in Python).  Train: classifier_pool ctx."""
        pl = self.classifier_pool[self.numtask - 1]
        if self.training:
            a, b = pl(text, None)
            return self.text_encoder.encode(a, b, None)[0]
        tokenized = text if isinstance(text, torch.Tensor) else pl.tokenize(text)
        sel = torch.as_tensor(text_category, device=tokenized.device).to(torch.int32).contiguous()
        if self._fused_eval():
            f = self._factor_stacks()
            return self.clip_model.text_engine().forward(tokenized, None, sel, factors=(f["d1"], f["d2t"], f["d3t"], f["scale"]))[0]
        _, tt = self._prompt_tables()
        return TextEncodeFn.apply(self.clip_model.text_engine(), tokenized, tt, sel, tuple(self.clip_model.inject_layers))[0]

    def update_fc(self, nb_classes):
        self.numtask += 1

    def copy(self):
        return copy.deepcopy(self)

    def freeze(self):
        for param in self.parameters():
            param.requires_grad = False
        self.eval()
        return self
