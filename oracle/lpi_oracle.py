"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the LPI continual-retrieval hot path.

This is the *checker*, never the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may import it.  `lpi_b200/` must not.

Parity pin: the reference ships no tests or golden vectors for this path,
so the restatement is pinned against the reference *itself*, imported on CPU in the build
container (oracle/reference_loader.py) -- see tests/test_oracle_vs_reference.py (runs where
/root/reference exists) and the committed fixtures under tests/golden/ which were generated
from the real reference by tests/golden/make_golden.py and are re-checked against this
restatement on every box (tests/test_oracle_golden.py).

Everything is plain torch-on-CPU arithmetic written from the maths of the cited lines
(paths relative to /root/reference/retrieval); gradients come from torch autograd over
this restatement.  dtype follows the inputs (fp32 by default, fp64 for tie audits).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------
# a1  DecomposedPrompt.forward            models/prompts/prompts.py:38-57
# ----------------------------------------------------------------------------------------

def decomposed_prompt(d1: Tensor, d2v: Tensor, d2t: Tensor, d3v: Tensor, d3t: Tensor) -> Tuple[Tensor, Tensor]:
    """vis[l,p,d] = mean_k d1[l,k] d2v[p,k] d3v[d,k]; txt likewise with the *shared* d1."""
    r = d1.shape[1]
    vis = torch.einsum("lk,pk,dk->lpd", d1, d2v, d3v) / r
    txt = torch.einsum("lk,pk,dk->lpd", d1, d2t, d3t) / r
    return vis, txt


# ----------------------------------------------------------------------------------------
# a4  ResidualAttentionBlock              models/clip/model.py:154-196
# ----------------------------------------------------------------------------------------

def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm computed in fp32 (model.py:154-160); biased variance, eps inside sqrt."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def quick_gelu(z: Tensor) -> Tensor:
    """model.py:163-165"""
    return z * torch.sigmoid(1.702 * z)


def attention(x: Tensor, w_in: Tensor, b_in: Tensor, w_out: Tensor, b_out: Tensor, heads: int, causal: bool) -> Tensor:
    """nn.MultiheadAttention as used at model.py:172,183-185: packed in_proj rows ordered q,k,v;
    heads are contiguous 64-wide slices; softmax(q k^T / sqrt(dh) + mask) v; dropout 0.
    x: [B, L, D] (batch-major; the reference uses LND, the maths is layout independent)."""
    B, L, D = x.shape
    dh = D // heads
    qkv = x @ w_in.t() + b_in
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, L, heads, dh).transpose(1, 2)
    k = k.view(B, L, heads, dh).transpose(1, 2)
    v = v.view(B, L, heads, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if causal:  # additive -inf strictly above the diagonal, model.py:347-353
        mask = torch.full((L, L), float("-inf"), dtype=x.dtype).triu_(1)
        s = s + mask
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, L, D)
    return o @ w_out.t() + b_out


def residual_block(x: Tensor, sd: Dict[str, Tensor], pfx: str, heads: int, causal: bool) -> Tensor:
    """model.py:187-196 (the per-layer injection at :190-193 is dead code as shipped)."""
    g = lambda k: sd[pfx + k]
    x = x + attention(layer_norm(x, g("ln_1.weight"), g("ln_1.bias")), g("attn.in_proj_weight"), g("attn.in_proj_bias"),
                      g("attn.out_proj.weight"), g("attn.out_proj.bias"), heads, causal)
    h = layer_norm(x, g("ln_2.weight"), g("ln_2.bias"))
    h = quick_gelu(h @ g("mlp.c_fc.weight").t() + g("mlp.c_fc.bias"))
    return x + h @ g("mlp.c_proj.weight").t() + g("mlp.c_proj.bias")


def _n_layers(sd: Dict[str, Tensor], pfx: str) -> int:
    n = 0
    while f"{pfx}{n}.ln_1.weight" in sd:
        n += 1
    return n


# ----------------------------------------------------------------------------------------
# a3  VisionTransformer.forward           models/clip/model.py:227-259
# ----------------------------------------------------------------------------------------

def vision_forward(sd: Dict[str, Tensor], images: Tensor, prompt_layers: Optional[Tensor] = None,
                   inject_layers: Sequence[int] = ()) -> Tensor:
    """images [B,3,224,224]; prompt_layers [B,Lp,P,Dv] or [Lp,P,Dv] or None -> [B,E].

    Token order: CLS(+pos0) | P prompt rows of layer 0 (NO positional term, model.py:240-248) |
    196 patches (+pos[1:]).  `inject_layers` (default empty = reference as shipped, C1) adds
    prompt_layers[:, l] to rows 1..P at the input of block l (the *intended* semantics of
    model.py:190-193, cf. grounding modeling_bert.py:749-773)."""
    pfx = "visual."
    w = sd[pfx + "conv1.weight"]
    B = images.shape[0]
    D = w.shape[0]
    x = F.conv2d(images, w, stride=w.shape[-1])                # [B, D, 14, 14]
    x = x.reshape(B, D, -1).permute(0, 2, 1)                   # [B, 196, D]
    cls = sd[pfx + "class_embedding"].expand(B, 1, D)
    x = torch.cat([cls, x], dim=1) + sd[pfx + "positional_embedding"]
    P = 0
    if prompt_layers is not None:
        if prompt_layers.dim() == 3:
            prompt_layers = prompt_layers.unsqueeze(0).expand(B, -1, -1, -1)
        P = prompt_layers.shape[2]
        x = torch.cat([x[:, :1], prompt_layers[:, 0], x[:, 1:]], dim=1)
    x = layer_norm(x, sd[pfx + "ln_pre.weight"], sd[pfx + "ln_pre.bias"])
    heads = D // 64
    for l in range(_n_layers(sd, pfx + "transformer.resblocks.")):
        if prompt_layers is not None and l in inject_layers and l != 0:
            x = torch.cat([x[:, :1], x[:, 1:1 + P] + prompt_layers[:, l], x[:, 1 + P:]], dim=1)
        x = residual_block(x, sd, f"{pfx}transformer.resblocks.{l}.", heads, causal=False)
    x = layer_norm(x[:, 0], sd[pfx + "ln_post.weight"], sd[pfx + "ln_post.bias"])
    return x @ sd[pfx + "proj"]


# ----------------------------------------------------------------------------------------
# a5/a6  PromptLearner.forward + TextEncoder.forward   models/clip/prompt_learner.py:128-163, 52-63
# ----------------------------------------------------------------------------------------

def text_forward(sd: Dict[str, Tensor], tokens: Tensor, ctx: Optional[Tensor] = None,
                 prompt_layers: Optional[Tensor] = None, inject_layers: Sequence[int] = ()) -> Tensor:
    """tokens [B,77] int64 (SOT, 16 'X' placeholders, caption, '.', EOT, 0-pad).
    ctx [P,Dt] or [B,P,Dt]: spliced over positions 1..P (CLASS_TOKEN_POSITION='end',
    prompt_learner.py:152-163); None = extract_vector path (raw 'X' embeddings, :118-126).
    Prompt rows DO receive the positional embedding (prompt_learner.py:53).
    Output row = argmax(tokens) (EOT) after ln_final, times text_projection (:57-61)."""
    emb = sd["token_embedding.weight"][tokens]                 # [B,77,Dt], no grad in the reference
    B = tokens.shape[0]
    P = 0
    if ctx is not None:
        if ctx.dim() == 2:
            ctx = ctx.unsqueeze(0).expand(B, -1, -1)
        P = ctx.shape[1]
        emb = torch.cat([emb[:, :1], ctx, emb[:, 1 + P:]], dim=1)
    x = emb + sd["positional_embedding"]
    D = x.shape[-1]
    heads = D // 64
    for l in range(_n_layers(sd, "transformer.resblocks.")):
        if prompt_layers is not None and l in inject_layers and l != 0:
            pl = prompt_layers if prompt_layers.dim() == 4 else prompt_layers.unsqueeze(0).expand(B, -1, -1, -1)
            x = torch.cat([x[:, :1], x[:, 1:1 + P] + pl[:, l], x[:, 1 + P:]], dim=1)
        x = residual_block(x, sd, f"transformer.resblocks.{l}.", heads, causal=True)
    x = layer_norm(x, sd["ln_final.weight"], sd["ln_final.bias"])
    eot = tokens.argmax(dim=-1)
    return x[torch.arange(B), eot] @ sd["text_projection"]


def l2_normalize(x: Tensor) -> Tensor:
    """slinet.py:122,133 -- no epsilon."""
    return x / x.norm(dim=-1, keepdim=True)


# ----------------------------------------------------------------------------------------
# a2  SliNet.forward                      models/slinet.py:109-135
# ----------------------------------------------------------------------------------------

def slinet_forward(sd: Dict[str, Tensor], factors: Dict[str, Tensor], images: Tensor, tokens: Tensor,
                   inject_layers: Sequence[int] = ()):
    vis, txt = decomposed_prompt(factors["dim_1_share"], factors["dim_2_visual"], factors["dim_2_textual"],
                                 factors["dim_3_visual"], factors["dim_3_textual"])
    img_f = l2_normalize(vision_forward(sd, images, vis, inject_layers))
    txt_f = l2_normalize(text_forward(sd, tokens, txt[0], txt, inject_layers))
    return img_f, txt_f, vis, txt


# ----------------------------------------------------------------------------------------
# a8/a9  ClipLoss / nt_bxent_loss         loss/loss.py:75-87, 6-33
# ----------------------------------------------------------------------------------------

def clip_loss(logits: Tensor) -> Tensor:
    n = logits.shape[0]
    lab = torch.arange(n)
    return 0.5 * (F.cross_entropy(logits, lab) + F.cross_entropy(logits.t(), lab))


def nt_bxent_loss(x: Tensor, target: Tensor, temperature: float) -> Tensor:
    """loss.py:6-33 including the double sigmoid (BCE-with-logits applied to an already
    sigmoided value, loss.py:21 -- reference quirk C5, reproduced on purpose)."""
    n = x.shape[0]
    xn = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-8)      # F.cosine_similarity eps=1e-8
    c = xn @ xn.t()
    eye = torch.eye(n, dtype=torch.bool)
    c = torch.where(eye, torch.full_like(c, float("inf")), c)
    z = torch.sigmoid(c / temperature)
    y = target.to(x.dtype)
    ell = torch.clamp(z, min=0) - z * y + torch.log1p(torch.exp(-z.abs()))
    pos = y.bool()
    lp = torch.where(pos, ell, torch.zeros_like(ell)).sum(1) / y.sum(1)
    ln = torch.where(~pos, ell, torch.zeros_like(ell)).sum(1) / (n - y.sum(1))
    return (lp + ln).mean()


# ----------------------------------------------------------------------------------------
# a7  SliNet.cal_loss / cal_task_loss     models/slinet.py:137-183
# ----------------------------------------------------------------------------------------

def cal_loss(img_f: Tensor, txt_f: Tensor, vis: Tensor, txt: Tensor, logit_scale: Tensor,
             prev_vis: Sequence[Tensor] = (), prev_txt: Sequence[Tensor] = (),
             task_sim: Optional[np.ndarray] = None) -> Dict[str, Tensor]:
    """vis [Lp,P,Dv], txt [Lp,P,Dt] (the batch `expand` + batch mean of slinet.py:146-152 is a no-op).
    prev_*: flattened-able prompts of tasks 0..t-1 (frozen); task loss only when present
    (numtask != 1, slinet.py:160)."""
    s = logit_scale.exp()
    losses = {"base_loss": clip_loss(s * img_f @ txt_f.t())}
    temperature = 0.01
    v = vis.mean(-1) / temperature                              # [Lp,P]
    u = txt.mean(-1) / temperature
    losses["alignment_loss"] = 0.1 * clip_loss(v @ u.t())
    if len(prev_vis) > 0:
        t = len(prev_vis)
        tgt = torch.tensor((task_sim[: t + 1, : t + 1] > 0.4).astype(np.int32))
        xv = torch.stack([p.reshape(-1) for p in list(prev_vis) + [vis]])
        xt = torch.stack([p.reshape(-1) for p in list(prev_txt) + [txt]])
        losses["task_loss"] = 0.1 * 0.5 * (nt_bxent_loss(xv, tgt, 0.001) + nt_bxent_loss(xt, tgt, 0.001))
    return losses


# ----------------------------------------------------------------------------------------
# a13  task-id selection                   methods/sprompt.py:336-368
# ----------------------------------------------------------------------------------------

def nearest_task_l1(feat: Tensor, keys: Sequence[Tensor]) -> Tensor:
    """sel[b] = argmin_t min_c sum_d |f[b,d]-K_t[c,d]|  (written ((f-c)**2)**0.5 in the reference);
    torch.min semantics = first occurrence on ties."""
    per_task = []
    for centers in keys:
        d = (feat[:, None, :] - centers[None, :, :].to(feat.dtype)).abs().sum(-1)   # [B,C]
        per_task.append(d.min(1)[0])
    return torch.stack(per_task).min(0)[1]


# ----------------------------------------------------------------------------------------
# a16  itm_eval (Recall@K)                 methods/sprompt.py:550-646
# ----------------------------------------------------------------------------------------

def ranks_by_count(scores: np.ndarray, gt: Sequence[Sequence[int]]) -> np.ndarray:
    """rank = #{j: s_j > s*} + #{j < g*: s_j == s*} with s* the best ground-truth score
    (lowest-index tie rule; equals the reference's argsort position whenever no exact tie
    touches a ground-truth score -- SURVEY.md A10)."""
    out = np.zeros(scores.shape[0], dtype=np.int64)
    for i, row in enumerate(scores):
        best = None
        for g in gt[i]:
            r = int((row > row[g]).sum() + (row[:g] == row[g]).sum())
            best = r if best is None or r < best else best
        out[i] = best
    return out


def itm_eval(scores_i2t: np.ndarray, scores_t2i: np.ndarray, txt2img, img2txt, category_i, category_t,
             task_num: int) -> dict:
    """Result dict exactly as sprompt.py:638-646: {'mscoco': {'i2t': {task:[r1,r5,r10]}, 't2i': {...}}}."""
    n_i, n_t = scores_i2t.shape
    ranks_i = ranks_by_count(scores_i2t, [img2txt[i] for i in range(n_i)])
    ranks_t = ranks_by_count(scores_t2i, [[txt2img[t]] for t in range(n_t)])
    cat_i = np.asarray([int(c) for c in category_i])
    cat_t = np.asarray([int(c) for c in category_t])

    def per_task(ranks, cat):
        res = {}
        for task in range(task_num):
            r = ranks[cat == task]
            res[task] = [100.0 * int((r < k).sum()) / len(r) for k in (1, 5, 10)]
        return res

    return {"mscoco": {"i2t": per_task(ranks_i, cat_i), "t2i": per_task(ranks_t, cat_t)}}


def topk_lowest_index(scores: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Exact top-k per row ordered by (score desc, index asc) -- the build's tie rule."""
    idx = np.lexsort((np.broadcast_to(np.arange(scores.shape[1]), scores.shape), -scores), axis=1)[:, :k]
    return np.take_along_axis(scores, idx, 1), idx


# ----------------------------------------------------------------------------------------
# a17  optimiser                           methods/sprompt.py:253-254
# ----------------------------------------------------------------------------------------

def sgd_momentum_step(w: Tensor, g: Tensor, buf: Optional[Tensor], lr: float, momentum: float = 0.9,
                      weight_decay: float = 2e-4) -> Tuple[Tensor, Tensor]:
    """torch.optim.SGD, no dampening / nesterov: g += wd*w; v = g (first) or m*v+g; w -= lr*v."""
    g = g + weight_decay * w
    buf = g.clone() if buf is None else momentum * buf + g
    return w - lr * buf, buf


def cosine_lr(base_lr: float, epoch: int, t_max: int) -> float:
    """CosineAnnealingLR(T_max=epochs), eta_min 0, closed form."""
    return base_lr * 0.5 * (1.0 + math.cos(math.pi * epoch / t_max))


# ----------------------------------------------------------------------------------------
# convenience: one full train-step oracle (forward, 3 losses, prompt grads)
# ----------------------------------------------------------------------------------------

FACTOR_NAMES = ("dim_1_share", "dim_2_visual", "dim_2_textual", "dim_3_visual", "dim_3_textual")


def train_step(sd: Dict[str, Tensor], factors: Dict[str, Tensor], images: Tensor, tokens: Tensor,
               prev_factors: Sequence[Dict[str, Tensor]] = (), task_sim: Optional[np.ndarray] = None,
               inject_layers: Sequence[int] = ()):
    fac = {k: factors[k].detach().clone().requires_grad_(True) for k in FACTOR_NAMES}
    img_f, txt_f, vis, txt = slinet_forward(sd, fac, images, tokens, inject_layers)
    prev_vis, prev_txt = [], []
    for pf in prev_factors:
        with torch.no_grad():
            pv, pt = decomposed_prompt(*[pf[k] for k in FACTOR_NAMES])
        prev_vis.append(pv)
        prev_txt.append(pt)
    losses = cal_loss(img_f, txt_f, vis, txt, sd["logit_scale"], prev_vis, prev_txt, task_sim)
    total = sum(losses.values())
    total.backward()
    return {
        "img_f": img_f.detach(), "txt_f": txt_f.detach(),
        "losses": {k: float(v.detach()) for k, v in losses.items()},
        "grads": {k: fac[k].grad.detach() for k in FACTOR_NAMES},
    }


# ----------------------------------------------------------------------------------------
# CPU baseline leg of bench.py: the reference's own ranking procedure, literally
# (methods/sprompt.py:509 dense matmul, :559-567 / :597-599 argsort + np.where)
# ----------------------------------------------------------------------------------------

def reference_ranks_argsort(scores: np.ndarray, gt: Sequence[Sequence[int]]) -> np.ndarray:
    """rank[i] = min over ground truths g of the position of g in np.argsort(row)[::-1]."""
    ranks = np.zeros(scores.shape[0])
    for index, score in enumerate(scores):
        inds = np.argsort(score)[::-1]
        rank = 1e20
        for g in gt[index]:
            tmp = np.where(inds == g)[0][0]
            if tmp < rank:
                rank = tmp
        ranks[index] = rank
    return ranks


def dense_scores(queries: Tensor, gallery: Tensor) -> Tensor:
    """`image_feats @ text_feats.t()` (sprompt.py:509) in fp32 on the host cores."""
    return queries.float() @ gallery.float().t()


def audit_topk(got_idx: np.ndarray, want_idx: np.ndarray, q_row: Tensor, gallery: Tensor, tol: float = 4e-7) -> str:
    """Near-tie audit of one query's top-k list (SURVEY.md section 8(d) scorer parity protocol).
    'exact' = identical lists; 'near' = every differing position swaps two items whose fp64 scores
    differ by < tol (fp32 accumulation-order noise); 'bad' otherwise."""
    if np.array_equal(got_idx, want_idx):
        return "exact"
    q64 = q_row.double()
    for a, b in zip(got_idx.tolist(), want_idx.tolist()):
        if a == b:
            continue
        if a < 0 or a >= gallery.shape[0]:
            return "bad"
        sa = float(q64 @ gallery[a].double())
        sb = float(q64 @ gallery[b].double())
        if abs(sa - sb) >= tol:
            return "bad"
    return "near"
