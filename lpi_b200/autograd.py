"""torch.autograd bridges over the CUDA engine, so the nn.Module mirrors (prompts.py, clip.py, loss.py, slinet.py) behave like
the reference's modules under `loss.backward()` (retrieval/methods/sprompt.py:308-311).  Every forward/backward body is a
sequence of liblpi_b200.so kernels; torch only tracks the graph."""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import losses, ops


class DecomposedPromptFn(torch.autograd.Function):
    """prompts.py:38-57 -> (vis [L,P,Dv], txt [L,P,Dt])"""

    @staticmethod
    def forward(ctx, d1, d2v, d2t, d3v, d3t):
        fs = [t.detach().contiguous().float() for t in (d1, d2v, d2t, d3v, d3t)]
        ctx.save_for_backward(*fs)
        return ops.prompt_fwd(*fs)

    @staticmethod
    def backward(ctx, g_vis, g_txt):
        fs = ctx.saved_tensors
        L, P, Dv, Dt = fs[0].shape[0], fs[1].shape[0], fs[3].shape[0], fs[4].shape[0]
        gv = g_vis.contiguous().float() if g_vis is not None else torch.zeros(L, P, Dv, device=fs[0].device)
        gt = g_txt.contiguous().float() if g_txt is not None else torch.zeros(L, P, Dt, device=fs[0].device)
        return tuple(ops.prompt_bwd(*fs, gv, gt))


def _as_table(tokens: Optional[torch.Tensor]):
    """[B, Lp, P, D] per-sample prompt tokens -> (table [T, Lp, P, D], sel int32[B] or None).  An `expand`ed batch
    (stride 0, slinet.py:119,129) collapses to one table row; a materialised batch becomes B rows with sel = arange."""
    if tokens is None:
        return None, None
    if tokens.dim() == 3:
        return tokens.unsqueeze(0), None
    if tokens.stride(0) == 0 or tokens.shape[0] == 1:
        return tokens[0:1], None
    return tokens, torch.arange(tokens.shape[0], device=tokens.device, dtype=torch.int32)


class VisionEncodeFn(torch.autograd.Function):
    """VisionTransformer.forward (model.py:227-259) + L2 norm (slinet.py:122) -> (feat_normalised, z_raw)."""

    @staticmethod
    def forward(ctx, engine, images, table, sel, inject_layers):
        need = table is not None and table.requires_grad
        tape = {} if need else None
        tab = None if table is None else table.detach().contiguous().float()
        feat, z = engine.forward(images.detach().float(), tab, sel, tape, inject_layers)
        ctx.engine, ctx.tape = engine, tape
        ctx.table_shape = None if table is None else table.shape
        return feat, z

    @staticmethod
    def backward(ctx, dfeat, dz):
        if ctx.tape is None:
            return None, None, None, None, None
        G = ctx.engine.backward(ctx.tape, dfeat, dz)
        ctx.tape = None
        return None, None, G.view(ctx.table_shape), None, None


class TextEncodeFn(torch.autograd.Function):
    """PromptLearner splice + TextEncoder.forward (prompt_learner.py:133-163, 52-63) + L2 norm -> (feat, z)."""

    @staticmethod
    def forward(ctx, engine, tokens, table, sel, inject_layers):
        need = table is not None and table.requires_grad
        tape = {} if need else None
        tab = None if table is None else table.detach().contiguous().float()
        feat, z = engine.forward(tokens, tab, sel, tape, inject_layers)
        ctx.engine, ctx.tape = engine, tape
        ctx.table_shape = None if table is None else table.shape
        return feat, z

    @staticmethod
    def backward(ctx, dfeat, dz):
        if ctx.tape is None:
            return None, None, None, None, None
        G = ctx.engine.backward(ctx.tape, dfeat, dz)
        ctx.tape = None
        return None, None, G.view(ctx.table_shape), None, None


class TextEmbeddedFn(torch.autograd.Function):
    """TextEncoder.forward on already-embedded prompts [B, 77, D] (the reference's module boundary,
    prompt_learner.py:52-63): + positional embedding, tower, ln_final, EOT gather, projection."""

    @staticmethod
    def forward(ctx, engine, prompts, tokens, table, sel, inject_layers):
        need = prompts.requires_grad or (table is not None and table.requires_grad)
        tape = {} if need else None
        tab = None if table is None else table.detach().contiguous().float()
        feat, z = engine.forward_embedded(prompts.detach().contiguous().float(), tokens, tab, sel, tape, inject_layers)
        ctx.engine, ctx.tape, ctx.shape = engine, tape, prompts.shape
        ctx.table_shape = None if table is None else table.shape
        return feat, z

    @staticmethod
    def backward(ctx, dfeat, dz):
        if ctx.tape is None:
            return None, None, None, None, None, None
        g, G = ctx.engine.backward_embedded(ctx.tape, dfeat, dz)
        ctx.tape = None
        return None, g.view(ctx.shape), None, (None if G is None else G.view(ctx.table_shape)), None, None


class SpliceFn(torch.autograd.Function):
    """PromptLearner.forward's embedding lookup + context splice (prompt_learner.py:133-163, CLASS_TOKEN_POSITION 'end'):
    prompts[b] = [E[tok[b,0]], ctx[b or shared], E[tok[b,17:]]]; the lookup is under no_grad in the reference."""

    @staticmethod
    def forward(ctx, engine, tokens, ctx_table, sel):
        B, L = tokens.shape
        P = 0 if ctx_table is None else ctx_table.shape[1]
        tab = None if ctx_table is None else ctx_table.detach().contiguous().float()
        x = ops.assemble_text(engine.emb, tokens.contiguous(), engine.zero_pos, tab, sel, B, L, P, engine.width)
        ctx.meta = (B, L, P, engine.width, None if ctx_table is None else ctx_table.shape[0], sel)
        return x.view(B, L, engine.width)

    @staticmethod
    def backward(ctx, g):
        B, L, P, D, T, sel = ctx.meta
        if T is None:
            return None, None, None, None
        d = ops.sum_prompt_rows(g.contiguous().float().view(B * L, D), sel, B, L, P, T, D)
        return None, None, d, None


class ClipLossFn(torch.autograd.Function):
    """ClipLoss.forward(logits) (loss.py:75-87)."""

    @staticmethod
    def forward(ctx, logits):
        loss, d = ops.clip_loss_logits(logits.detach().contiguous().float(), 1.0, logits.requires_grad)
        ctx.d = d
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        return ctx.d * g


class ContrastiveFn(torch.autograd.Function):
    """ClipLoss(scale * I @ T^T) fused with its backward (slinet.py:138-141): the B x B logits are formed once."""

    @staticmethod
    def forward(ctx, img_f, txt_f, scale):
        loss, dI, dT, _ = losses.contrastive_fwd_bwd(img_f.detach().contiguous().float(), txt_f.detach().contiguous().float(), float(scale))
        ctx.save_for_backward(dI, dT)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        dI, dT = ctx.saved_tensors
        return dI * g, dT * g, None


class AlignmentLossFn(torch.autograd.Function):
    """0.1 * ClipLoss(mean_d(vis)/0.01 @ (mean_d(txt)/0.01)^T) (slinet.py:144-158) on [L,P,D] prompt tensors."""

    @staticmethod
    def forward(ctx, vis, txt):
        Gv, Gt = torch.zeros_like(vis, dtype=torch.float32), torch.zeros_like(txt, dtype=torch.float32)
        loss = losses.alignment_fwd_bwd(vis.detach().contiguous().float(), txt.detach().contiguous().float(), Gv, Gt)
        ctx.save_for_backward(Gv, Gt)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        Gv, Gt = ctx.saved_tensors
        return Gv * g, Gt * g


class TaskLossFn(torch.autograd.Function):
    """nt_bxent_loss(x, target, temperature) (loss.py:6-33) with the gradient for the LAST row of x only (the rows of
    earlier tasks are frozen, slinet.py:176-180; other rows get a zero gradient)."""

    @staticmethod
    def forward(ctx, x, target, temperature):
        xs = x.detach().contiguous().float()
        R, n = xs.shape
        loss = torch.zeros(1, device=xs.device, dtype=torch.float32)
        g_last = torch.zeros(n, device=xs.device, dtype=torch.float32)
        ops.task_loss(xs, target.to(torch.int32).contiguous(), float(temperature), 1.0, loss, False, g_last, False)
        ctx.save_for_backward(g_last)
        ctx.R = R
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (g_last,) = ctx.saved_tensors
        G = torch.zeros(ctx.R, g_last.shape[0], device=g_last.device, dtype=torch.float32)
        G[-1] = g_last * g
        return G, None, None
