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
from typing import Dict, Optional, Sequence

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


def _record(t: torch.Tensor, stream) -> None:
    """Tell the caching allocator that `t` (allocated on another stream) is in use on `stream`.  Inside a CUDA-graph capture the
    allocation lives in the graph's private pool and the fork / join edges order every use, so nothing is recorded there."""
    if not torch.cuda.is_current_stream_capturing():
        t.record_stream(stream)


def reconstruct(factors: Dict[str, torch.Tensor]):
    return ops.prompt_fwd(*[factors[k] for k in FACTOR_NAMES])


def flat_size(factors: Dict[str, torch.Tensor]) -> int:
    return sum(factors[k].numel() for k in FACTOR_NAMES)


def train_step(vision: VisionEngine, text: TextEngine, factors: Dict[str, torch.Tensor], images: torch.Tensor, tokens: torch.Tensor,
               logit_scale: float, prev_prompts: Sequence = (), task_target: Optional[torch.Tensor] = None,
               inject_layers: Sequence[int] = (), group=None, text_len: Optional[int] = None, overlap_towers: Optional[bool] = None,
               prompt_scale: float = 1.0, want_logits: bool = True) -> Dict:
    """factors: the five fp32 device tensors of the current task's DecomposedPrompt.  images [b,3,224,224] fp32 and
    tokens [b,77] int64 are this rank's slice of the global batch.  prev_prompts: [(vis, txt)] of the frozen earlier
    tasks (task loss, only when non-empty).  Returns losses (0-dim-like device tensors), grads (same keys as factors)
    and the features.  With `group`, features are all-gathered and the gradient all-reduced (sum).  text_len: host-side bound on the
    EOT positions of `tokens` (see TextEngine.forward; output-exact trimming of the padding after the last EOT).
    prompt_scale: DecomposedPrompt.scale (prompts.py:26; 1 in the reference).

    The token sequences take their prompt rows straight from the factors (reconstruction fused into the assembly kernels); the
    [9, 16, D] tables are still reconstructed once per step because the alignment and task losses read all nine layers."""
    if overlap_towers is None:
        overlap_towers = OVERLAP_TOWERS
    vis, txt = reconstruct(factors)
    if prompt_scale != 1.0:
        vis, txt = vis * prompt_scale, txt * prompt_scale
    fv = (factors["dim_1_share"].unsqueeze(0), factors["dim_2_visual"].unsqueeze(0), factors["dim_3_visual"].unsqueeze(0), prompt_scale)
    ft = (factors["dim_1_share"].unsqueeze(0), factors["dim_2_textual"].unsqueeze(0), factors["dim_3_textual"].unsqueeze(0), prompt_scale)
    deep = len(inject_layers) > 0                   # opt-in deep injection reads layers >= 1 of the table
    vtab, ttab = (vis.unsqueeze(0), txt.unsqueeze(0)) if deep else (None, None)
    vtape, ttape = {}, {}
    side = _side_stream(images.device) if overlap_towers else None
    main = torch.cuda.current_stream()
    if side is not None:
        # the text tower's kernels are small (M = b x ~40 rows: 20-40 of the 74 CTA pairs busy, LayerNorm / attention grids far below one
        # wave), so it runs on a second stream next to the vision tower and fills the SMs that tower's non-persistent kernels leave idle
        side.wait_stream(main)
        with torch.cuda.stream(side):
            txt_f, _ = text.forward(tokens, ttab, None, ttape, inject_layers, text_len=text_len, factors=ft)
        img_f, _ = vision.forward(images, vtab, None, vtape, inject_layers, factors=fv)
        main.wait_stream(side)
        _record(txt_f, main)
    else:
        img_f, _ = vision.forward(images, vtab, None, vtape, inject_layers, factors=fv)
        txt_f, _ = text.forward(tokens, ttab, None, ttape, inject_layers, text_len=text_len, factors=ft)
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
    # similarities + InfoNCE + their backward in one launch; the B x B logits are written only when asked for (want_logits)
    base, d_img, d_txt, logits = losses.contrastive_fwd_bwd(all_img, all_txt, logit_scale, rank * b, b, want_logits=want_logits)
    if side is not None:
        side.wait_stream(main)
        _record(d_txt, side)
        with torch.cuda.stream(side):
            G_txt = text.backward(ttape, d_txt)[0]
        G_vis = vision.backward(vtape, d_img)[0]      # [Lp, P, Dv]: batch-summed (prompts are `expand`ed, slinet.py:119,129)
        main.wait_stream(side)
        _record(G_txt, main)
    else:
        G_vis = vision.backward(vtape, d_img)[0]
        G_txt = text.backward(ttape, d_txt)[0]
    if world > 1:
        # replicated terms (alignment / task losses) are added once after the all-reduce, not world times
        enc = _factor_grads(factors, G_vis * prompt_scale if prompt_scale != 1.0 else G_vis, G_txt * prompt_scale if prompt_scale != 1.0 else G_txt)
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
    if prompt_scale != 1.0:                        # G_* are gradients w.r.t. the SCALED tables
        G_vis, G_txt = G_vis * prompt_scale, G_txt * prompt_scale
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


class GraphedTrainStep:
    """train_step + PromptSGD.step captured once into a CUDA graph and replayed: ~400 kernel launches (two streams, the NCCL
    exchange included) become one cudaGraphLaunch, so the step no longer depends on how fast one Python thread can issue launches --
    at 8 ranks per host the eager step was CPU-bound (11.6 ms against 9.0 ms of GPU work).

    Shapes, text_len, the learning rate and the task set are frozen at capture time: build a new object when one of them changes
    (per epoch for the cosine schedule).  New batches are copied into the captured input buffers by step()."""

    def __init__(self, vision: VisionEngine, text: TextEngine, factors: Dict[str, torch.Tensor], opt: "PromptSGD", images: torch.Tensor,
                 tokens: torch.Tensor, logit_scale: float, prev_prompts: Sequence = (), task_target: Optional[torch.Tensor] = None,
                 inject_layers: Sequence[int] = (), group=None, text_len: Optional[int] = None, warmup: int = 2):
        self.vision, self.text, self.factors, self.opt = vision, text, factors, opt
        self.images, self.tokens = images.clone(), tokens.clone()
        self.kw = dict(prev_prompts=prev_prompts, task_target=task_target, inject_layers=inject_layers, group=group, text_len=text_len)
        self.logit_scale = logit_scale
        cur = torch.cuda.current_stream()
        warm = torch.cuda.Stream(device=images.device)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm):                      # one-time lazy setup (function attributes, NCCL channels, SGD first step) happens here
            for _ in range(max(1, warmup)):
                self._eager()
        cur.wait_stream(warm)
        torch.cuda.synchronize()
        n0 = ops.KERNEL_LAUNCHES
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._eager()
        self.launches = ops.KERNEL_LAUNCHES - n0

    def _eager(self) -> Dict:
        r = train_step(self.vision, self.text, self.factors, self.images, self.tokens, self.logit_scale, **self.kw)
        self.opt.step(r["grads"])
        return r

    def prefetch(self, images: torch.Tensor, tokens: torch.Tensor) -> None:
        """Starts the host -> device copy of the NEXT batch (pinned host tensors of the captured shapes) on a copy stream into a staging
        buffer, so that it overlaps the step that is running; the following step() picks the staged batch up with a device-to-device
        copy (38 MB at HBM speed instead of at PCIe speed on the critical path)."""
        if getattr(self, "_stage", None) is None:
            self._stage = (torch.empty_like(self.images), torch.empty_like(self.tokens))
            self._copy_stream = torch.cuda.Stream(device=self.images.device)
            self._staged = None
        cs = self._copy_stream
        # the staging buffers are free once the previous step() has copied them into the captured inputs -- an event recorded right after
        # those two device-to-device copies, NOT the end of the replay that follows them: waiting for the whole stream would put the PCIe
        # copy (0.7 ms for a 64-image batch) between two steps instead of under one
        consumed = getattr(self, "_consumed", None)
        if consumed is not None:
            cs.wait_event(consumed)
        else:
            cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cs):
            self._stage[0].copy_(images, non_blocking=True)
            self._stage[1].copy_(tokens, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        self._staged = ev

    def step(self, images: Optional[torch.Tensor] = None, tokens: Optional[torch.Tensor] = None) -> Dict:
        """Replays the captured step (optionally on a new batch of the captured shape, or on the batch staged by prefetch()); returns the
        captured result tensors."""
        staged = getattr(self, "_staged", None)
        if images is None and tokens is None and staged is not None:
            torch.cuda.current_stream().wait_event(staged)
            self.images.copy_(self._stage[0], non_blocking=True)
            self.tokens.copy_(self._stage[1], non_blocking=True)
            self._staged = None
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream())
        if images is not None:
            self.images.copy_(images, non_blocking=True)
        if tokens is not None:
            self.tokens.copy_(tokens, non_blocking=True)
        self.graph.replay()
        ops._count(self.launches)
        return self.out
