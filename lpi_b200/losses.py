"""Loss-side composition of the LPI learner on the kernels of liblpi_b200.so (functional layer; the nn.Module
mirrors in loss.py / slinet.py wrap these in autograd Functions).

Reference: retrieval/models/slinet.py:137-183 (cal_loss, cal_task_loss), retrieval/loss/loss.py:6-33, 75-87.
Each function returns the loss value AND the analytic gradients the backward needs,
so one pass over the logits serves both directions.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import ops

ALIGN_TEMPERATURE = 0.01      # slinet.py:144
TASK_TEMPERATURE = 0.001      # slinet.py:182
TASK_THRESHOLD = 0.4          # slinet.py:173
AUX_WEIGHT = 0.1              # slinet.py:158,161


FUSED_INFONCE = __import__("os").environ.get("LPI_FUSED_INFONCE", "1") != "0"     # 0 = the four-launch form (3 fp32 GEMMs + ClipLoss kernels)
# The one-launch kernel recomputes every score instead of storing it (5 n^2 E MACs on warp-reduced dot products against 3 n^2 E in tiled
# GEMMs): measured 49.8 vs 94.1 us at n = 64, 364 vs 211 us at n = 512, 8.4 vs 2.5 ms at n = 4096 -- it is used up to the crossover.
FUSED_INFONCE_MAX_N = int(__import__("os").environ.get("LPI_FUSED_INFONCE_MAX_N", "256"))


def contrastive_fwd_bwd(img_f: torch.Tensor, txt_f: torch.Tensor, scale: float, row0: int = 0, n_local: Optional[int] = None,
                        want_grad: bool = True, want_logits: bool = True):
    """base_loss = ClipLoss(scale * I @ T^T) over the GLOBAL batch (img_f, txt_f: all-gathered [n, E] fp32) and its gradient
    for the local rows [row0, row0 + n_local): dI = scale * G[rows, :] @ T, dT = scale * G[:, rows]^T @ I with
    G = (softmax_rows + softmax_cols - 2 I) / (2n)  (loss.py:75-87; slinet.py:138-141)."""
    n = img_f.shape[0]
    n_local = n if n_local is None else n_local
    if FUSED_INFONCE and n <= FUSED_INFONCE_MAX_N:
        # one cooperative launch: similarities, both LSEs, loss and the local gradient rows; the logits are an optional by-product
        return ops.sim_infonce_fwd_bwd(img_f.contiguous(), txt_f.contiguous(), scale, row0, n_local, 1.0, want_grad, want_logits)
    logits = ops.sgemm(img_f, txt_f.t(), alpha=scale)
    loss, dS = ops.clip_loss_logits(logits, 1.0, want_grad)
    if not want_grad:
        return loss, None, None, logits
    d_img = ops.sgemm(dS[row0:row0 + n_local], txt_f, alpha=scale)
    d_txt = ops.sgemm(dS[:, row0:row0 + n_local].t(), img_f, alpha=scale)
    return loss, d_img, d_txt, logits


def alignment_fwd_bwd(vis: torch.Tensor, txt: torch.Tensor, G_vis: Optional[torch.Tensor], G_txt: Optional[torch.Tensor]):
    """alignment_loss = 0.1 * ClipLoss(V @ U^T), V = mean_d(vis)/0.01 in [L,P], U likewise (slinet.py:144-158).
    Accumulates d loss / d vis, d loss / d txt into G_vis / G_txt ([L,P,D] fp32) when given."""
    L, P, Dv = vis.shape
    Dt = txt.shape[2]
    V = ops.row_mean(vis.reshape(L * P, Dv), 1.0 / ALIGN_TEMPERATURE).view(L, P)
    U = ops.row_mean(txt.reshape(L * P, Dt), 1.0 / ALIGN_TEMPERATURE).view(L, P)
    S = ops.sgemm(V, U.t())
    want = G_vis is not None
    loss, dS = ops.clip_loss_logits(S, AUX_WEIGHT, want)
    if want:
        dV = ops.sgemm(dS, U)
        dU = ops.sgemm(dS.t(), V)
        ops.add_rowconst(G_vis.view(L * P, Dv), dV.view(-1), 1.0 / (ALIGN_TEMPERATURE * Dv), True)
        ops.add_rowconst(G_txt.view(L * P, Dt), dU.view(-1), 1.0 / (ALIGN_TEMPERATURE * Dt), True)
    return loss


def task_fwd_bwd(vis_stack: torch.Tensor, txt_stack: torch.Tensor, target: torch.Tensor, G_vis: Optional[torch.Tensor],
                 G_txt: Optional[torch.Tensor]):
    """task_loss = 0.1 * 1/2 [nt_bxent(vis_stack) + nt_bxent(txt_stack)] (slinet.py:160-183); *_stack = [t+1, L*P*D] flattened
    prompts of tasks 0..t with the current (trainable) task LAST; target int32 [t+1, t+1] = task_sim > 0.4.
    Accumulates the gradient wrt the last row into G_vis / G_txt when given."""
    loss = torch.zeros(1, device=vis_stack.device, dtype=torch.float32)
    w = AUX_WEIGHT * 0.5
    ops.task_loss(vis_stack, target, TASK_TEMPERATURE, w, loss, True, None if G_vis is None else G_vis.view(-1), True)
    ops.task_loss(txt_stack, target, TASK_TEMPERATURE, w, loss, True, None if G_txt is None else G_txt.view(-1), True)
    return loss
