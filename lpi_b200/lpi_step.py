"""One LPI training step on the kernels, without torch autograd: forward of both towers, the three losses, the analytic
backward down to the five DecomposedPrompt factors, optional data-parallel exchange, SGD step.

This is the fused equivalent of the reference's hot loop body (retrieval/methods/sprompt.py:300-311):
    SliNet.forward -> SliNet.cal_loss -> sum(losses) -> zero_grad -> backward -> optimizer.step
with the data-parallel recipe of SURVEY.md section 8(e): batch sharded over ranks, ONE all-gather of the [b, 512] image
and text features feeding the global B x B InfoNCE, local dgrad, ONE all-reduce of the flat 5 284-float prompt gradient.
The nn.Module / autograd mirror of the same maths lives in slinet.py; both call the same kernels.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch

from . import losses, ops
from .engine import TextEngine, VisionEngine

FACTOR_NAMES = ("dim_1_share", "dim_2_visual", "dim_2_textual", "dim_3_visual", "dim_3_textual")
OVERLAP_TOWERS = os.environ.get("LPI_OVERLAP_TOWERS", "1") != "0"
_SIDE_STREAMS: Dict = {}


def _side_stream(device) -> "torch.cuda.Stream":
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return st


def reconstruct(factors: Dict[str, torch.Tensor]):
    return ops.prompt_fwd(*[factors[k] for k in FACTOR_NAMES])


def flat_size(factors: Dict[str, torch.Tensor]) -> int:
    return sum(factors[k].numel() for k in FACTOR_NAMES)


def train_step(vision: VisionEngine, text: TextEngine, factors: Dict[str, torch.Tensor], images: torch.Tensor, tokens: torch.Tensor,
               logit_scale: float, prev_prompts: Sequence = (), task_target: Optional[torch.Tensor] = None,
               inject_layers: Sequence[int] = (), group=None, text_len: Optional[int] = None, overlap_towers: Optional[bool] = None) -> Dict:
    """factors: the five fp32 device tensors of the current task's DecomposedPrompt.  images [b,3,224,224] fp32 and
    tokens [b,77] int64 are this rank's slice of the global batch.  prev_prompts: [(vis, txt)] of the frozen earlier
    tasks (task loss, only when non-empty).  Returns losses (0-dim-like device tensors), grads (same keys as factors)
    and the features.  With `group`, features are all-gathered and the gradient all-reduced (sum).  text_len: host-side bound on the
    EOT positions of `tokens` (see TextEngine.forward; output-exact trimming of the padding after the last EOT)."""
    if overlap_towers is None:
        overlap_towers = OVERLAP_TOWERS
    vis, txt = reconstruct(factors)
    vtape, ttape = {}, {}
    side = _side_stream(images.device) if overlap_towers else None
    main = torch.cuda.current_stream()
    if side is not None:
        # the text tower's kernels are small (M = b x ~40 rows: 20-40 of the 74 CTA pairs busy, LayerNorm / attention grids far below one
        # wave), so it runs on a second stream next to the vision tower and fills the SMs that tower's non-persistent kernels leave idle
        side.wait_stream(main)
        with torch.cuda.stream(side):
            txt_f, _ = text.forward(tokens, txt.unsqueeze(0), None, ttape, inject_layers, text_len=text_len)
        img_f, _ = vision.forward(images, vis.unsqueeze(0), None, vtape, inject_layers)
        main.wait_stream(side)
        txt_f.record_stream(main)
    else:
        img_f, _ = vision.forward(images, vis.unsqueeze(0), None, vtape, inject_layers)
        txt_f, _ = text.forward(tokens, txt.unsqueeze(0), None, ttape, inject_layers, text_len=text_len)
    b = img_f.shape[0]
    rank, world = 0, 1
    all_img, all_txt = img_f, txt_f
    if group is not None:
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if world > 1:
            both = torch.cat([img_f, txt_f], dim=1).contiguous()                 # one buffer, one collective
            gathered = torch.empty(world * b, both.shape[1], device=both.device, dtype=both.dtype)
            dist.all_gather_into_tensor(gathered, both, group=group)
            E = img_f.shape[1]
            all_img, all_txt = gathered[:, :E].contiguous(), gathered[:, E:].contiguous()
    base, d_img, d_txt, logits = losses.contrastive_fwd_bwd(all_img, all_txt, logit_scale, rank * b, b)
    if side is not None:
        side.wait_stream(main)
        d_txt.record_stream(side)
        with torch.cuda.stream(side):
            G_txt = text.backward(ttape, d_txt)[0]
        G_vis = vision.backward(vtape, d_img)[0]      # [Lp, P, Dv]: batch-summed (prompts are `expand`ed, slinet.py:119,129)
        main.wait_stream(side)
        G_txt.record_stream(main)
    else:
        G_vis = vision.backward(vtape, d_img)[0]
        G_txt = text.backward(ttape, d_txt)[0]
    if world > 1:
        # replicated terms (alignment / task losses) are added once after the all-reduce, not world times
        enc = _factor_grads(factors, G_vis, G_txt)
        fg = torch.cat([enc[k].reshape(-1) for k in FACTOR_NAMES])      # 5 284 floats = 21 KB: one flat buffer, one collective
        dist.all_reduce(fg, op=dist.ReduceOp.SUM, group=group)
        G_vis = torch.zeros_like(G_vis)
        G_txt = torch.zeros_like(G_txt)
    out_losses = {"base_loss": base}
    out_losses["alignment_loss"] = losses.alignment_fwd_bwd(vis, txt, G_vis, G_txt)
    if len(prev_prompts) > 0:
        vs = torch.stack([p[0].reshape(-1) for p in prev_prompts] + [vis.reshape(-1)])
        ts = torch.stack([p[1].reshape(-1) for p in prev_prompts] + [txt.reshape(-1)])
        out_losses["task_loss"] = losses.task_fwd_bwd(vs, ts, task_target, G_vis, G_txt)
    grads = _factor_grads(factors, G_vis, G_txt)
    if world > 1:
        off = 0
        for k in FACTOR_NAMES:
            n = grads[k].numel()
            grads[k] = grads[k] + fg[off:off + n].view_as(grads[k])
            off += n
    return {"losses": out_losses, "grads": grads, "img_f": img_f, "txt_f": txt_f, "logits": logits}


def _factor_grads(factors, G_vis, G_txt) -> Dict[str, torch.Tensor]:
    outs = ops.prompt_bwd(*[factors[k] for k in FACTOR_NAMES], G_vis.contiguous(), G_txt.contiguous())
    return dict(zip(FACTOR_NAMES, outs))


class PromptSGD:
    """torch.optim.SGD(momentum 0.9, weight_decay 2e-4) + CosineAnnealingLR(T_max = epochs) over the five factors
    (sprompt.py:253-254), as one fused kernel per tensor."""

    def __init__(self, factors: Dict[str, torch.Tensor], lr: float, momentum: float = 0.9, weight_decay: float = 2e-4, t_max: int = 10):
        self.factors = factors
        self.base_lr, self.momentum, self.wd, self.t_max = lr, momentum, weight_decay, t_max
        self.epoch = 0
        self.bufs = {k: torch.zeros_like(factors[k]) for k in FACTOR_NAMES}
        self.first = True

    @property
    def lr(self) -> float:
        import math

        return self.base_lr * 0.5 * (1.0 + math.cos(math.pi * self.epoch / self.t_max))

    def step(self, grads: Dict[str, torch.Tensor]):
        for k in FACTOR_NAMES:
            ops.sgd_momentum_step(self.factors[k], grads[k].contiguous(), self.bufs[k], self.lr, self.momentum, self.wd, self.first)
        self.first = False

    def epoch_end(self):
        self.epoch += 1
