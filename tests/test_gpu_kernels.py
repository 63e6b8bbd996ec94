"""-m gpu unit parity of the tower / loss kernels against plain fp32 torch maths of the same op (floating point: the
tolerance of each check is written next to it) and against the oracle restatement where one exists."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lpi_b200 import losses, ops
from oracle import lpi_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("M,D", [(1, 512), (213 * 3, 768), (1000, 512), (77, 1024)])
def test_layernorm_fwd_bwd(M, D):
    g = torch.Generator().manual_seed(M + D)
    x = (torch.randn(M, D, generator=g) * 2 + 0.3).cuda()
    gamma = (1 + 0.1 * torch.randn(D, generator=g)).cuda()
    beta = (0.05 * torch.randn(D, generator=g)).cuda()
    yf, yb = ops.layernorm_fwd(x, gamma, beta, want_f32=True, want_bf16=True)
    ref = F.layer_norm(x, (D,), gamma, beta, 1e-5)
    assert (yf - ref).abs().max() < 2e-5                       # fp32 LN, two-pass variance
    assert (yb.float() - ref).abs().max() < 2 ** -7 * ref.abs().max()      # bf16 rounding of the output
    dy = torch.randn(M, D, generator=g).cuda()
    xr = x.clone().requires_grad_(True)
    F.layer_norm(xr, (D,), gamma, beta, 1e-5).backward(dy)
    base = torch.randn(M, D, generator=g).cuda()
    gacc = base.clone()
    gb = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    ops.layernorm_bwd(dy, x, gamma, gacc, gb, accumulate=True)
    assert (gacc - (base + xr.grad)).abs().max() < 5e-5
    assert (gb.float() - gacc).abs().max() <= 2 ** -8 * gacc.abs().max()
    g2 = torch.empty_like(x)
    ops.layernorm_bwd(dy, x, gamma, g2, None, accumulate=False)
    assert (g2 - xr.grad).abs().max() < 5e-5


def _ref_attention(qkv, B, L, H, causal):
    D = H * 64
    q, k, v = qkv.float().view(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / 8.0
    if causal:
        s = s + torch.full((L, L), float("-inf"), device=s.device).triu_(1)
    p = torch.softmax(s, -1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * L, D), s


@pytest.mark.parametrize("B,L,H,causal", [(2, 213, 12, False), (3, 77, 8, True), (1, 197, 12, False), (2, 64, 2, True),
                                           (1, 1, 1, False), (2, 130, 3, True), (1, 256, 1, False)])
def test_attention_fwd_bwd(B, L, H, causal):
    g = torch.Generator().manual_seed(L * 31 + H)
    D = H * 64
    qkv = (torch.randn(B * L, 3 * D, generator=g) * 1.5).cuda().bfloat16()
    out, lse = ops.attn_fwd(qkv, B, L, H, causal)
    qr = qkv.float().requires_grad_(True)
    ref, s = _ref_attention(qr, B, L, H, causal)
    # operands are identical bf16 values; differences = bf16 rounding of P and of the output (2^-8 relative each)
    assert (out.float() - ref).abs().max() < 0.02 * ref.abs().max()
    want_lse = torch.logsumexp(s, -1) * math.log2(math.e)            # kernel stores log2-domain LSE of the scaled scores
    assert (lse.view(B, H, L) - want_lse).abs().max() < 1e-3
    d_out = torch.randn(B * L, D, generator=g).cuda().bfloat16()
    ref.backward(d_out.float())
    out2, _, of = ops.attn_fwd(qkv, B, L, H, causal, want_f32=True)         # fp32 hand-over for TF32 consumers
    assert torch.equal(out2, out) and (of - ref).abs().max() < 0.02 * ref.abs().max() and torch.equal(of.bfloat16(), out)
    dq32 = ops.attn_bwd(qkv, out, d_out, lse, B, L, H, causal, f32=True)
    dqkv = ops.attn_bwd(qkv, out, d_out, lse, B, L, H, causal)
    assert dq32.dtype == torch.float32 and torch.equal(dq32.bfloat16(), dqkv)
    assert _rel(dqkv.float(), qr.grad) < 2e-2                          # bf16 P / dS / outputs vs fp32 autograd
    for part in range(3):
        sl = slice(part * D, (part + 1) * D)
        if qr.grad[:, sl].norm() < 1e-6:                               # L = 1: softmax is constant, dq = dk = 0 exactly
            assert dqkv[:, sl].float().abs().max() < 1e-5
        else:
            assert _rel(dqkv[:, sl].float(), qr.grad[:, sl]) < 3e-2, part


@pytest.mark.parametrize("B,L,H,causal", [(3, 77, 8, True), (2, 213, 12, False), (1, 16, 1, False), (2, 130, 3, True)])
def test_attention_fp16(B, L, H, causal):
    """fp16 storage (text tower): tighter than bf16 (P, dS and the outputs carry 10 mantissa bits); the backward is linear in d_out,
    so scaling d_out by 2^10 (the engine's gradient scale) scales dqkv by exactly 2^10."""
    g = torch.Generator().manual_seed(L * 17 + H)
    D = H * 64
    qkv = (torch.randn(B * L, 3 * D, generator=g) * 1.5).cuda().half()
    out, lse = ops.attn_fwd(qkv, B, L, H, causal)
    qr = qkv.float().requires_grad_(True)
    ref, s = _ref_attention(qr, B, L, H, causal)
    assert out.dtype == torch.float16 and (out.float() - ref).abs().max() < 4e-3 * ref.abs().max()
    want_lse = torch.logsumexp(s, -1) * math.log2(math.e)
    assert (lse.view(B, H, L) - want_lse).abs().max() < 1e-3
    d_out = (torch.randn(B * L, D, generator=g) * 1e-3).cuda().half()
    ref.backward(d_out.float())
    dqkv = ops.attn_bwd(qkv, out, d_out, lse, B, L, H, causal)
    assert dqkv.dtype == torch.float16
    d_scaled = (d_out.float() * 1024).half()
    dq_scaled = ops.attn_bwd(qkv, out, d_scaled, lse, B, L, H, causal)
    assert _rel(dq_scaled.float() / 1024, qr.grad) < 4e-3              # scaled path: fp16 rounding only
    assert _rel(dqkv.float(), qr.grad) < 2e-2                          # unscaled 1e-3 gradients already lose bits to fp16 subnormals


def test_layernorm_fp16_shadow_and_grad_scale():
    g = torch.Generator().manual_seed(11)
    M, D = 333, 512
    x = torch.randn(M, D, generator=g).cuda() * 3 + 1
    gamma, beta = (1 + 0.1 * torch.randn(D, generator=g)).cuda(), (0.1 * torch.randn(D, generator=g)).cuda()
    of, oh = ops.layernorm_fwd(x, gamma, beta, want_f32=True, half_dtype=torch.float16)
    assert oh.dtype == torch.float16 and torch.equal(of.half(), oh)
    dy = torch.randn(M, D, generator=g).cuda() * 1e-4
    xr = x.clone().requires_grad_(True)
    F.layer_norm(xr, (D,), gamma, beta, 1e-5).backward(dy)
    base = torch.randn(M, D, generator=g).cuda() * 1e-4
    gacc = base.clone()
    gh = torch.empty(M, D, device="cuda", dtype=torch.float16)
    ops.layernorm_bwd(dy * 1024, x, gamma, gacc, gh, accumulate=True, grad_scale=1024.0)
    assert (gacc - (base + xr.grad)).abs().max() < 5e-5 * 1e-3
    assert (gh.float() / 1024 - gacc).abs().max() <= 2 ** -10 * gacc.abs().max()


def test_im2col_matches_conv():
    g = torch.Generator().manual_seed(5)
    img = torch.randn(3, 3, 224, 224, generator=g).cuda()
    w = (torch.randn(768, 3, 16, 16, generator=g) * 0.02).cuda()
    patches = ops.im2col_patches(img, 16)
    pe = ops.gemm(patches, w.reshape(768, -1).bfloat16().contiguous(), ops.EPI_F32)
    ref = F.conv2d(img.bfloat16().float(), w.bfloat16().float(), stride=16).reshape(3, 768, -1).permute(0, 2, 1).reshape(-1, 768)
    assert (pe - ref).abs().max() < 1e-3 * max(1.0, ref.abs().max().item())


def test_assemble_vision_and_text():
    g = torch.Generator().manual_seed(6)
    B, n_patch, P, D, T = 5, 196, 16, 768, 3
    pe = torch.randn(B * n_patch, D, generator=g).cuda()
    cls, pos = torch.randn(D, generator=g).cuda(), torch.randn(n_patch + 1, D, generator=g).cuda()
    table = torch.randn(T, P, D, generator=g).cuda()
    sel = torch.tensor([2, 0, 1, 1, 2], dtype=torch.int32).cuda()
    gam, bet = (1 + 0.1 * torch.randn(D, generator=g)).cuda(), (0.1 * torch.randn(D, generator=g)).cuda()
    x = ops.assemble_vision(pe, cls, pos, table, sel, gam, bet, B, n_patch, P, D)
    tr = table.clone().requires_grad_(True)
    pre = torch.cat([(cls + pos[0]).expand(B, 1, D), tr[sel.long()], pe.view(B, n_patch, D) + pos[1:]], 1)
    ref = F.layer_norm(pre, (D,), gam, bet, 1e-5)
    assert (x.view(B, -1, D) - ref).abs().max() < 3e-5
    gup = torch.randn(B * (1 + P + n_patch), D, generator=g).cuda()
    ref.backward(gup.view(B, -1, D))
    d = ops.assemble_vision_bwd(gup, table, sel, gam, B, 1 + P + n_patch, P, T, D)
    assert (d - tr.grad).abs().max() < 1e-4
    x0 = ops.assemble_vision(pe, cls, pos, None, None, gam, bet, B, n_patch, 0, D)          # un-prompted CLIP
    ref0 = F.layer_norm(torch.cat([(cls + pos[0]).expand(B, 1, D), pe.view(B, n_patch, D) + pos[1:]], 1), (D,), gam, bet, 1e-5)
    assert (x0.view(B, -1, D) - ref0).abs().max() < 3e-5
    # text
    V, L, Dt = 1000, 77, 512
    emb, tpos = torch.randn(V, Dt, generator=g).cuda(), torch.randn(L, Dt, generator=g).cuda()
    tok = torch.randint(0, V, (B, L), generator=g).cuda()
    ctx = torch.randn(T, P, Dt, generator=g).cuda()
    xt = ops.assemble_text(emb, tok, tpos, ctx, sel, B, L, P, Dt)
    e = emb[tok]
    reft = torch.cat([e[:, :1], ctx[sel.long()], e[:, 1 + P:]], 1) + tpos
    assert torch.equal(xt.view(B, L, Dt), reft)
    assert torch.equal(ops.assemble_text(emb, tok, tpos, None, None, B, L, 0, Dt).view(B, L, Dt), e + tpos)
    gt = torch.randn(B * L, Dt, generator=g).cuda()
    dctx = ops.sum_prompt_rows(gt, sel, B, L, P, T, Dt)
    want = torch.zeros(T, P, Dt, device="cuda").index_add_(0, sel.long(), gt.view(B, L, Dt)[:, 1:1 + P])
    assert (dctx - want).abs().max() < 1e-5
    xi = xt.clone()
    ops.inject_prompt_rows(xi, ctx, sel, B, L, P, Dt)
    refi = xt.view(B, L, Dt).clone()
    refi[:, 1:1 + P] += ctx[sel.long()]
    assert torch.equal(xi.view(B, L, Dt), refi)


@pytest.mark.parametrize("B,D,E", [(1, 768, 512), (7, 768, 512), (64, 512, 512)])
def test_head_fwd_bwd(B, D, E):
    g = torch.Generator().manual_seed(B)
    L = 9
    x = torch.randn(B * L, D, generator=g).cuda()
    rows = (torch.arange(B) * L + torch.randint(0, L, (B,), generator=g)).to(torch.int32).cuda()
    gam, bet = (1 + 0.1 * torch.randn(D, generator=g)).cuda(), (0.1 * torch.randn(D, generator=g)).cuda()
    proj = (torch.randn(D, E, generator=g) * D ** -0.5).cuda()
    f, z = ops.head_fwd(x, rows, gam, bet, proj)
    xr = x.clone().requires_grad_(True)
    zr = F.layer_norm(xr[rows.long()], (D,), gam, bet, 1e-5) @ proj
    fr = zr / zr.norm(dim=-1, keepdim=True)
    assert (z - zr).abs().max() < 2e-5 * max(1.0, zr.abs().max().item())
    assert (f - fr).abs().max() < 2e-6
    df, dz = torch.randn(B, E, generator=g).cuda(), torch.randn(B, E, generator=g).cuda()
    (fr * df).sum().backward(retain_graph=True)
    gbuf = torch.zeros_like(x)
    gb = torch.zeros(B * L, D, device="cuda", dtype=torch.bfloat16)
    ops.head_bwd(df, None, z, x, rows, gam, proj, gbuf, gb)
    assert _rel(gbuf, xr.grad) < 1e-4
    xr.grad = None
    ((fr * df).sum() + (zr * dz).sum()).backward()
    gbuf.zero_()
    ops.head_bwd(df, dz, z, x, rows, gam, proj, gbuf, None)
    assert _rel(gbuf, xr.grad) < 1e-4


def test_decomposed_prompt_fwd_bwd():
    from lpi_b200 import synthetic as S

    fac = {k: v.cuda() for k, v in S.make_prompt_factors(3).items()}
    vis, txt = ops.prompt_fwd(*[fac[k] for k in O.FACTOR_NAMES])
    wv, wt = O.decomposed_prompt(*[fac[k].cpu() for k in O.FACTOR_NAMES])
    assert (vis.cpu() - wv).abs().max() < 1e-6 and (txt.cpu() - wt).abs().max() < 1e-6
    g = torch.Generator().manual_seed(1)
    gv, gt = torch.randn(9, 16, 768, generator=g), torch.randn(9, 16, 512, generator=g)
    fr = {k: fac[k].cpu().clone().requires_grad_(True) for k in O.FACTOR_NAMES}
    rv, rt = O.decomposed_prompt(*[fr[k] for k in O.FACTOR_NAMES])
    ((rv * gv).sum() + (rt * gt).sum()).backward()
    outs = ops.prompt_bwd(*[fac[k] for k in O.FACTOR_NAMES], gv.cuda(), gt.cuda())
    for k, o in zip(O.FACTOR_NAMES, outs):
        assert _rel(o.cpu(), fr[k].grad) < 1e-5, k


@pytest.mark.parametrize("n", [1, 4, 64, 200])
def test_clip_loss_and_contrastive(n):
    g = torch.Generator().manual_seed(n)
    logits = (torch.randn(n, n, generator=g) * 3).cuda()
    loss, d = ops.clip_loss_logits(logits, 1.0)
    lr = logits.clone().requires_grad_(True)
    want = O.clip_loss(lr.cpu())
    want.backward()
    assert abs(float(loss) - float(want)) < 1e-5 * max(1.0, abs(float(want)))
    lrg = torch.autograd.grad(O.clip_loss(lr), lr)[0] if False else None
    l2 = logits.cpu().clone().requires_grad_(True)
    O.clip_loss(l2).backward()
    assert (d.cpu() - l2.grad).abs().max() < 1e-6
    E = 512
    I = F.normalize(torch.randn(n, E, generator=g), dim=-1).cuda()
    T = F.normalize(torch.randn(n, E, generator=g), dim=-1).cuda()
    s = 1 / 0.07
    Ir, Tr = I.cpu().clone().requires_grad_(True), T.cpu().clone().requires_grad_(True)
    ref = O.clip_loss(s * Ir @ Tr.t())
    ref.backward()
    loss, dI, dT, lg = losses.contrastive_fwd_bwd(I, T, s)
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert _rel(dI.cpu(), Ir.grad) < 1e-5 and _rel(dT.cpu(), Tr.grad) < 1e-5
    if n >= 4:       # local-row slices of a global batch (data-parallel ranks)
        r0, nl = n // 4, n // 2
        _, dI2, dT2, _ = losses.contrastive_fwd_bwd(I, T, s, r0, nl)
        assert torch.equal(dI2, dI[r0:r0 + nl]) and torch.equal(dT2, dT[r0:r0 + nl])
    # the one-launch fused kernel against the four-launch form (3 fp32 GEMMs + ClipLoss kernels) and its optional logits
    assert losses.FUSED_INFONCE and n <= losses.FUSED_INFONCE_MAX_N and (lg.cpu() - (s * Ir @ Tr.t()).detach()).abs().max() < 2e-5
    losses.FUSED_INFONCE = False
    try:
        loss_u, dI_u, dT_u, lg_u = losses.contrastive_fwd_bwd(I, T, s)
    finally:
        losses.FUSED_INFONCE = True
    assert abs(float(loss) - float(loss_u)) < 2e-6 * max(1.0, abs(float(loss_u))) and (lg - lg_u).abs().max() < 2e-5
    assert _rel(dI, dI_u) < 1e-5 and _rel(dT, dT_u) < 1e-5
    l0, a0, b0, none = ops.sim_infonce_fwd_bwd(I, T, s, want_grad=False, want_logits=False)
    assert a0 is None and b0 is None and none is None and float(l0) == float(loss)
    w_loss, w_dI, _, _ = ops.sim_infonce_fwd_bwd(I, T, s, weight=0.1)
    assert abs(float(w_loss) - 0.1 * float(loss)) < 1e-6 and _rel(w_dI, 0.1 * dI) < 1e-6


def test_fused_infonce_large_global_batch():
    """n = 1000 (a global batch of 8 x 125, not a multiple of anything convenient), E = 512: loss and both gradients vs fp64 autograd."""
    g = torch.Generator().manual_seed(77)
    n, E, s = 1000, 512, 1 / 0.07
    I = F.normalize(torch.randn(n, E, generator=g), dim=-1)
    T = F.normalize(0.3 * I + torch.randn(n, E, generator=g) / E ** 0.5, dim=-1)
    Ir, Tr = I.double().requires_grad_(True), T.double().requires_grad_(True)
    S_ = s * Ir @ Tr.t()
    lab = torch.arange(n)
    ref = 0.5 * (F.cross_entropy(S_, lab) + F.cross_entropy(S_.t(), lab))
    ref.backward()
    loss, dI, dT, lg = ops.sim_infonce_fwd_bwd(I.cuda(), T.cuda(), s, 125 * 3, 125, want_logits=False)
    assert lg is None and abs(float(loss) - float(ref)) < 1e-5 * float(ref)
    assert _rel(dI.cpu(), Ir.grad[375:500]) < 2e-5 and _rel(dT.cpu(), Tr.grad[375:500]) < 2e-5


def test_alignment_and_task_loss():
    import os

    from lpi_b200 import synthetic as S

    sim = np.loadtxt(os.path.join(os.path.dirname(os.path.abspath(ops.__file__)), "MID", "task_sim_matrix.txt"))
    for t in (1, 2, 4):
        prompts = [O.decomposed_prompt(*[S.make_prompt_factors(s)[k] for k in O.FACTOR_NAMES]) for s in range(t + 1)]
        vis = prompts[-1][0].clone().requires_grad_(True)
        txt = prompts[-1][1].clone().requires_grad_(True)
        want = O.cal_loss(F.normalize(torch.randn(4, 8), dim=-1), F.normalize(torch.randn(4, 8), dim=-1), vis, txt, torch.tensor(0.0),
                          [p[0] for p in prompts[:-1]], [p[1] for p in prompts[:-1]], sim)
        (want["alignment_loss"] + want["task_loss"]).backward()
        Gv = torch.zeros(9, 16, 768, device="cuda")
        Gt = torch.zeros(9, 16, 512, device="cuda")
        al = losses.alignment_fwd_bwd(vis.detach().cuda(), txt.detach().cuda(), Gv, Gt)
        tgt = torch.tensor((sim[:t + 1, :t + 1] > 0.4).astype(np.int32)).cuda()
        vs = torch.stack([p[0].reshape(-1) for p in prompts]).detach().cuda()
        ts = torch.stack([p[1].reshape(-1) for p in prompts]).detach().cuda()
        tl = losses.task_fwd_bwd(vs, ts, tgt, Gv, Gt)
        assert abs(float(al) - float(want["alignment_loss"])) < 1e-4 * max(1.0, abs(float(want["alignment_loss"])))
        assert abs(float(tl) - float(want["task_loss"])) < 1e-5
        assert _rel(Gv.cpu(), vis.grad) < 1e-3 and _rel(Gt.cpu(), txt.grad) < 1e-3


def test_sgd_and_nearest_center():
    g = torch.Generator().manual_seed(2)
    w, gr = torch.randn(5284, generator=g), torch.randn(5284, generator=g)
    wc, buf = w.clone().cuda(), torch.zeros(5284, device="cuda")
    ow, ob = O.sgd_momentum_step(w, gr, None, 0.05)
    ops.sgd_momentum_step(wc, gr.cuda(), buf, 0.05, 0.9, 2e-4, True)
    assert torch.allclose(wc.cpu(), ow, atol=1e-7) and torch.allclose(buf.cpu(), ob, atol=1e-7)
    gr2 = torch.randn(5284, generator=g)
    ow2, ob2 = O.sgd_momentum_step(ow, gr2, ob, 0.03)
    ops.sgd_momentum_step(wc, gr2.cuda(), buf, 0.03, 0.9, 2e-4, False)
    assert torch.allclose(wc.cpu(), ow2, atol=1e-6) and torch.allclose(buf.cpu(), ob2, atol=1e-6)
    f = F.normalize(torch.randn(50, 512, generator=g), dim=-1)
    keys = [F.normalize(torch.randn(5, 512, generator=g), dim=-1) for _ in range(4)]
    keys[2][1] = f[7]
    sel = ops.nearest_center_l1(f.cuda(), torch.stack(keys).cuda())
    assert torch.equal(sel.cpu(), O.nearest_task_l1(f, keys)) and int(sel[7]) == 2


def test_sgemm_strided():
    g = torch.Generator().manual_seed(8)
    a, b = torch.randn(70, 33, generator=g).cuda(), torch.randn(33, 130, generator=g).cuda()
    assert (ops.sgemm(a, b, alpha=0.5) - 0.5 * a @ b).abs().max() < 1e-4
    at = torch.randn(33, 70, generator=g).cuda()
    assert (ops.sgemm(at.t(), b) - at.t() @ b).abs().max() < 1e-4
    bt = torch.randn(130, 33, generator=g).cuda()
    assert (ops.sgemm(a, bt.t()) - a @ bt.t()).abs().max() < 1e-4


def test_factor_fused_assembly_is_bit_identical_to_reconstruct_then_assemble():
    """north_star subsystem 1: DecomposedPrompt.forward (prompts.py:38-57) fused with the token concat (model.py:240-248,
    prompt_learner.py:152-163): the assembly kernels rebuild the prompt rows from the stacked factors of T tasks (per-sample task
    selection) and must equal lpi_prompt_fwd -> table -> assemble bit for bit, forward and backward, with and without a scale."""
    from lpi_b200 import synthetic as S

    g = torch.Generator().manual_seed(41)
    B, n_patch, P, Dv, Dt, T, Lp, r = 6, 196, 16, 768, 512, 3, 9, 4
    facs = [{k: v.cuda() for k, v in S.make_prompt_factors(40 + t).items()} for t in range(T)]
    st = lambda name: torch.stack([f[name] for f in facs]).contiguous()
    tabs = [ops.prompt_fwd(*[f[k] for k in O.FACTOR_NAMES]) for f in facs]
    vt = torch.stack([t[0][0] for t in tabs]).contiguous()            # layer 0 of every task: [T, P, Dv]
    tt = torch.stack([t[1][0] for t in tabs]).contiguous()
    sel = torch.tensor([2, 0, 1, 1, 2, 0], dtype=torch.int32).cuda()
    pe = torch.randn(B * n_patch, Dv, generator=g).cuda()
    cls, pos = torch.randn(Dv, generator=g).cuda(), torch.randn(n_patch + 1, Dv, generator=g).cuda()
    gam, bet = (1 + 0.1 * torch.randn(Dv, generator=g)).cuda(), (0.1 * torch.randn(Dv, generator=g)).cuda()
    for scale in (1.0, 0.5):
        fv = (st("dim_1_share"), st("dim_2_visual"), st("dim_3_visual"), scale)
        ft = (st("dim_1_share"), st("dim_2_textual"), st("dim_3_textual"), scale)
        for s_ in (sel, None):
            x = ops.assemble_vision_factors(pe, cls, pos, fv, s_, gam, bet, B, n_patch, Dv)
            want = ops.assemble_vision(pe, cls, pos, (vt * scale).contiguous(), s_, gam, bet, B, n_patch, P, Dv)
            assert torch.equal(x, want), (scale, s_ is None)
        L = 1 + P + n_patch
        gup = torch.randn(B * L, Dv, generator=g).cuda()
        d = ops.assemble_vision_factors_bwd(gup, fv, sel, gam, B, L, T, Dv)
        assert torch.equal(d, ops.assemble_vision_bwd(gup, (vt * scale).contiguous(), sel, gam, B, L, P, T, Dv))
        V, Lt = 1000, 40
        emb, tpos = torch.randn(V, Dt, generator=g).cuda(), torch.randn(77, Dt, generator=g).cuda()
        tok = torch.randint(0, V, (B, Lt), generator=g).cuda()
        xt = ops.assemble_text_factors(emb, tok, tpos, ft, sel, B, Lt, Dt)
        assert torch.equal(xt, ops.assemble_text(emb, tok, tpos, (tt * scale).contiguous(), sel, B, Lt, P, Dt))


def test_head_fwd_select_equals_head_plus_nearest_center():
    g = torch.Generator().manual_seed(43)
    B, L, D, E, T = 37, 5, 768, 512, 6
    x = torch.randn(B * L, D, generator=g).cuda()
    rows = (torch.arange(B) * L).to(torch.int32).cuda()
    gam, bet = (1 + 0.1 * torch.randn(D, generator=g)).cuda(), (0.1 * torch.randn(D, generator=g)).cuda()
    proj = (torch.randn(D, E, generator=g) * D ** -0.5).cuda()
    f, z = ops.head_fwd(x, rows, gam, bet, proj)
    keys = F.normalize(torch.randn(T, 5, E, generator=g), dim=-1).cuda()
    keys[4, 2] = f[11]                                                 # an exact hit
    keys[1, 0] = keys[3, 4]                                            # a tie between tasks: the first one wins
    f2, z2, sel = ops.head_fwd_select(x, rows, gam, bet, proj, keys.contiguous())
    assert torch.equal(f, f2) and torch.equal(z, z2)
    want = ops.nearest_center_l1(f, keys.contiguous())
    assert torch.equal(sel.long(), want) and int(sel[11]) == 4


@pytest.mark.parametrize("B,L,H,causal,dtype", [(3, 77, 8, True, torch.float16), (2, 213, 12, False, torch.float16),
                                                (4, 39, 8, True, torch.bfloat16), (1, 1, 1, False, torch.bfloat16),
                                                (2, 300, 2, True, torch.float16)])
def test_attention_row_query_fwd_bwd(B, L, H, causal, dtype):
    """lpi_attn_rowq_*: the last block's attention for the one row per sample the head reads (model.py:254-257,
    prompt_learner.py:57-61) == the full attention at those rows; its backward == autograd of a loss that reads those rows only
    (dq zero elsewhere, dk / dv complete); x[rows] gathered and g[rows] scattered by the same launches."""
    g = torch.Generator().manual_seed(L * 13 + H)
    D = H * 64
    qkv = (torch.randn(B * L, 3 * D, generator=g) * 1.5).cuda().to(dtype)
    pos = torch.randint(0, L, (B,), generator=g)
    pos[0] = L - 1
    rows = (torch.arange(B) * L + pos).to(torch.int32).cuda()
    x = torch.randn(B * L, D, generator=g).cuda()
    out_rows, x_rows = ops.attn_rowq_fwd(qkv, rows, B, L, H, causal, x=x)
    assert torch.equal(x_rows, x[rows.long()])
    qr = qkv.float().requires_grad_(True)
    ref, _ = _ref_attention(qr, B, L, H, causal)
    ref_rows = ref[rows.long()]
    tol = 2e-3 if dtype == torch.float16 else 1e-2                    # fp32 arithmetic; only the output is rounded to 16 bits
    assert out_rows.dtype == dtype and (out_rows.float() - ref_rows).abs().max() < tol * ref_rows.abs().max()
    d_rows = torch.randn(B, D, generator=g).cuda().to(dtype)
    ref_rows.backward(d_rows.float())
    g_rows = torch.randn(B, D, generator=g).cuda()
    gfull = torch.zeros(B * L, D, device="cuda")
    dqkv = ops.attn_rowq_bwd(qkv, rows, d_rows, B, L, H, causal, g_rows=g_rows, g=gfull)
    want_g = torch.zeros_like(gfull)
    want_g[rows.long()] = g_rows
    assert torch.equal(gfull, want_g)
    assert dqkv.dtype == dtype and dqkv.shape == qkv.shape
    gt = qr.grad
    scale = gt.abs().max().clamp_min(1e-6)
    assert (dqkv.float() - gt).abs().max() < (4e-3 if dtype == torch.float16 else 2e-2) * scale
    dead = gt == 0
    if dead.any():
        assert dqkv.float()[dead].abs().max() == 0                     # exact zeros where the read rows have no influence
    # linear in d_out: the fp16 gradient scale passes through
    if dtype == torch.float16:
        d_small = (d_rows.float() * 2 ** -10).half()
        a = ops.attn_rowq_bwd(qkv, rows, d_small, B, L, H, causal)
        b2 = ops.attn_rowq_bwd(qkv, rows, (d_small.float() * 1024).half(), B, L, H, causal)
        assert _rel(b2.float(), a.float() * 1024) < 2e-3
