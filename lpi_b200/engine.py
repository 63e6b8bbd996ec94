"""Tower executor: sequences the sm_100a kernels of liblpi_b200.so for the prompted CLIP ViT-B/16 image tower and the
12-layer text transformer, forward and backward (dgrad only -- every CLIP weight is frozen in LPI, sprompt.py:229-237,
so there is no wgrad and GEMM inputs need not be kept).

Reference maths: retrieval/models/clip/model.py:168-259 (ResidualAttentionBlock, Transformer, VisionTransformer),
retrieval/models/clip/prompt_learner.py:52-63 (TextEncoder).  Layout here is batch-major tokens [B*L, D]
(the reference permutes to [L, B, D]; the maths is layout independent).

Precision plan (SURVEY.md section 7 error budget): bf16 GEMM/attention operands, fp32 accumulation, fp32 residual stream,
fp32 LayerNorm statistics, fp32 gradient stream with a bf16 shadow as the A operand of the next dgrad GEMM, fp32 heads.
torch is used for buffers and streams only.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch

from . import ops


@dataclass
class BlockWeights:
    ln1_g: torch.Tensor
    ln1_b: torch.Tensor
    ln2_g: torch.Tensor
    ln2_b: torch.Tensor
    w_in: torch.Tensor          # [3D, D] bf16   (and its transpose for dgrad)
    w_in_t: torch.Tensor        # [D, 3D]
    b_in: torch.Tensor          # [3D] fp32
    w_out: torch.Tensor         # [D, D]
    w_out_t: torch.Tensor
    b_out: torch.Tensor
    w_fc: torch.Tensor          # [4D, D]
    w_fc_t: torch.Tensor        # [D, 4D]
    b_fc: torch.Tensor
    w_proj: torch.Tensor        # [D, 4D]
    w_proj_t: torch.Tensor      # [4D, D]
    b_proj: torch.Tensor


def _bf16(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(dev, torch.float32).to(torch.bfloat16).contiguous()


def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(dev, torch.float32).contiguous()


def _half(t: torch.Tensor, dev, dtype) -> torch.Tensor:
    return t.detach().to(dev, torch.float32).to(dtype).contiguous()


def load_blocks(sd: Dict[str, torch.Tensor], prefix: str, dev, need_grad: bool = True, tf32: bool = False,
                half_dtype=torch.bfloat16) -> List[BlockWeights]:
    """`prefix` = 'visual.transformer.resblocks.' or 'transformer.resblocks.' in a CLIP state_dict.  tf32: keep the GEMM
    weights in fp32 (operands of lpi_gemm_tf32); otherwise they are stored in `half_dtype` (bf16, or fp16 for lpi_gemm_f16)."""
    blocks = []
    i = 0
    while f"{prefix}{i}.ln_1.weight" in sd:
        p = f"{prefix}{i}."
        g = lambda k: sd[p + k]
        def pair(k):
            w = _f32(g(k), dev) if tf32 else _half(g(k), dev, half_dtype)
            return w, (w.t().contiguous() if need_grad else None)
        w_in, w_in_t = pair("attn.in_proj_weight")
        w_out, w_out_t = pair("attn.out_proj.weight")
        w_fc, w_fc_t = pair("mlp.c_fc.weight")
        w_proj, w_proj_t = pair("mlp.c_proj.weight")
        blocks.append(BlockWeights(
            ln1_g=_f32(g("ln_1.weight"), dev), ln1_b=_f32(g("ln_1.bias"), dev), ln2_g=_f32(g("ln_2.weight"), dev),
            ln2_b=_f32(g("ln_2.bias"), dev), w_in=w_in, w_in_t=w_in_t, b_in=_f32(g("attn.in_proj_bias"), dev), w_out=w_out,
            w_out_t=w_out_t, b_out=_f32(g("attn.out_proj.bias"), dev), w_fc=w_fc, w_fc_t=w_fc_t, b_fc=_f32(g("mlp.c_fc.bias"), dev),
            w_proj=w_proj, w_proj_t=w_proj_t, b_proj=_f32(g("mlp.c_proj.bias"), dev)))
        i += 1
    return blocks


@dataclass
class BlockSaved:
    x: torch.Tensor             # block input, fp32 [M, D]
    x1: torch.Tensor            # after the attention residual, fp32
    qkv: torch.Tensor           # bf16 [M, 3D]
    o: torch.Tensor             # attention output, bf16 [M, D]
    lse: torch.Tensor
    z: torch.Tensor             # c_fc pre-activation, bf16 [M, 4D]


@dataclass
class TowerTape:
    B: int
    L: int
    blocks: List[BlockSaved] = field(default_factory=list)
    x_final: Optional[torch.Tensor] = None
    out_rows: Optional[torch.Tensor] = None      # int32 [B]: the last block ran on these rows only (Tower.forward, out_rows)
    injected: Dict[int, bool] = field(default_factory=dict)


class Tower:
    """12 pre-LN residual attention blocks (model.py:187-196) over a [B*L, D] fp32 residual stream."""

    def __init__(self, sd, prefix: str, heads: int, causal: bool, dev, need_grad: bool = True, precision: str = "bf16"):
        if precision not in ("bf16", "tf32", "fp16", "fp32"):
            raise ValueError(f"precision must be 'bf16', 'fp16', 'tf32' or 'fp32', got {precision!r}")
        # 'fp32' = the parity mode of north_star ("1e-5 in fp32"): the same block sequence as 'tf32' (fp32 tensors everywhere) with exact
        # fp32 products on the SIMT pipes (ops.gemm_f32) and fp32 attention (ops.attn_fwd_f32 / attn_bwd_f32) -- a test mode
        self.fp32 = precision == "fp32"
        self.tf32 = precision in ("tf32", "fp32")
        self.half = torch.float16 if precision == "fp16" else torch.bfloat16      # dtype of GEMM / attention operands and shadows
        # fp16 gradient path: the 16-bit gradient stream is carried times 2^10 so that small gradients stay in fp16's normal range
        # (min normal 6.1e-5); every op between two LayerNorm backwards is linear in the gradient, so the factor is exact.
        self.grad_scale = 1024.0 if precision == "fp16" else None
        # the dgrad GEMMs feeding the two LayerNorm backwards hand dy over in 16 bits (half the bytes of those HBM-bound kernels' largest
        # stream); LPI_DH_F32=1 keeps them fp32 for precision studies
        import os
        self.epi_dh = ops.EPI_F32 if os.environ.get("LPI_DH_F32") == "1" else ops.EPI_BF16
        self.blocks = load_blocks(sd, prefix, dev, need_grad, self.tf32, self.half)
        self.heads = heads
        self.causal = causal
        self.width = heads * 64
        self.dev = dev
        # The head reads ONE row per sample of the last block's output (model.py:254-257, prompt_learner.py:57-61), so the last block
        # computes its attention for that query row only (keys / values of every position) and its out_proj / ln_2 / MLP on [B, D]
        # instead of [B*L, D]; the rows that are skipped reach neither the features nor any gradient.  LPI_LAST_BLOCK_FULL=1 (or
        # last_block_rows = False) runs the full block instead; the tf32 / fp32 study modes always do.
        self.last_block_rows = os.environ.get("LPI_LAST_BLOCK_FULL") != "1" and not self.tf32
        # backward: the out_proj dgrad GEMM also produces the attention backward's delta = rowsum(dO o O) in its epilogue
        # (ops.gemm_do_delta) instead of a separate pass over dO and O per block; LPI_FUSED_DELTA=0 runs the separate kernel
        self.fuse_delta = os.environ.get("LPI_FUSED_DELTA", "1") != "0"

    def rows_only(self, L: int) -> bool:
        """Does forward(..., out_rows=rows) return [B, D] (last block on the read rows) for sequences of L tokens?"""
        return self.last_block_rows and not self.tf32 and L <= 512      # the one-row attention kernels are built for L <= 512

    # -------------------------------------------------------------------------------------------- forward
    def forward(self, x: torch.Tensor, B: int, L: int, tape: Optional[TowerTape] = None,
                inject: Optional[dict] = None, out_rows: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x fp32 [B*L, D] (consumed).  `inject` = {'layers': {l,...}, 'table': [T, Lp, P, D] fp32, 'sel': int32[B] or None, 'P': P}
        adds table[sel[b], l] to rows 1..P before block l (opt-in deep-prompt injection, l >= 1).
        out_rows int32 [B] (only honoured when self.last_block_rows): the caller reads just these rows of the output -> the result is
        [B, D], row b = output row out_rows[b]."""
        H = self.heads
        if self.tf32:
            return self._forward_tf32(x, B, L, tape, inject)
        if not self.rows_only(L):
            out_rows = None
        n_blocks = len(self.blocks)
        for li, w in enumerate(self.blocks):
            if inject is not None and li != 0 and li in inject["layers"]:
                ops.inject_prompt_rows(x, inject["table"][:, li].contiguous(), inject["sel"], B, L, inject["P"], self.width)
            _, h = ops.layernorm_fwd(x, w.ln1_g, w.ln1_b, half_dtype=self.half)
            qkv = ops.gemm(h, w.w_in, ops.EPI_BIAS_BF16, bias=w.b_in)
            if out_rows is not None and li == n_blocks - 1:
                # last block, read rows only: one query row per (sample, head) against all keys / values, then row-wise ops on [B, D]
                o, xr = ops.attn_rowq_fwd(qkv, out_rows, B, L, H, self.causal, x=x)
                x1 = ops.gemm(o, w.w_out, ops.EPI_BIAS_RESID_F32, bias=w.b_out, resid=xr)
                _, h2 = ops.layernorm_fwd(x1, w.ln2_g, w.ln2_b, half_dtype=self.half)
                z = torch.empty(B, 4 * self.width, device=x.device, dtype=self.half) if tape is not None else None
                a = ops.gemm(h2, w.w_fc, ops.EPI_BIAS_GELU_BF16, bias=w.b_fc, out2=z)
                x2 = ops.gemm(a, w.w_proj, ops.EPI_BIAS_RESID_F32, bias=w.b_proj, resid=x1)
                if tape is not None:
                    tape.blocks.append(BlockSaved(x=x, x1=x1, qkv=qkv, o=None, lse=None, z=z))
                    tape.out_rows = out_rows
                x = x2
                break
            o, lse = ops.attn_fwd(qkv, B, L, H, self.causal, want_lse=tape is not None)
            if tape is not None:
                x1 = ops.gemm(o, w.w_out, ops.EPI_BIAS_RESID_F32, bias=w.b_out, resid=x)
            else:
                x1 = ops.gemm(o, w.w_out, ops.EPI_BIAS_RESID_F32, bias=w.b_out, resid=x, out=x)      # in place
            _, h2 = ops.layernorm_fwd(x1, w.ln2_g, w.ln2_b, half_dtype=self.half)
            z = torch.empty(x.shape[0], 4 * self.width, device=x.device, dtype=self.half) if tape is not None else None
            a = ops.gemm(h2, w.w_fc, ops.EPI_BIAS_GELU_BF16, bias=w.b_fc, out2=z)
            if tape is not None:
                x2 = ops.gemm(a, w.w_proj, ops.EPI_BIAS_RESID_F32, bias=w.b_proj, resid=x1)
                tape.blocks.append(BlockSaved(x=x, x1=x1, qkv=qkv, o=o, lse=lse, z=z))
            else:
                x2 = ops.gemm(a, w.w_proj, ops.EPI_BIAS_RESID_F32, bias=w.b_proj, resid=x1, out=x1)
            x = x2
        if tape is not None:
            tape.x_final = x
        return x

    def _forward_tf32(self, x, B, L, tape, inject):
        """Same block sequence with fp32 GEMM operands on the TF32 tensor-core path (8x finer operand rounding than bf16);
        attention still runs on bf16 q/k/v but hands its output over in fp32.  precision 'fp32': exact-fp32 SIMT GEMMs and attention."""
        H = self.heads
        mm = ops.gemm_f32 if self.fp32 else ops.gemm_tf32
        for li, w in enumerate(self.blocks):
            if inject is not None and li != 0 and li in inject["layers"]:
                ops.inject_prompt_rows(x, inject["table"][:, li].contiguous(), inject["sel"], B, L, inject["P"], self.width)
            h, _ = ops.layernorm_fwd(x, w.ln1_g, w.ln1_b, want_f32=True, want_bf16=False)
            if self.fp32:
                qkv = mm(h, w.w_in, ops.EPI_BIAS_F32, bias=w.b_in)
                of, lse = ops.attn_fwd_f32(qkv, B, L, H, self.causal)
                o = of
            else:
                qkv = mm(h, w.w_in, ops.EPI_BIAS_BF16, bias=w.b_in)
                o, lse, of = ops.attn_fwd(qkv, B, L, H, self.causal, want_lse=tape is not None, want_f32=True)
            x1 = mm(of, w.w_out, ops.EPI_BIAS_RESID_F32, bias=w.b_out, resid=x, out=None if tape is not None else x)
            h2, _ = ops.layernorm_fwd(x1, w.ln2_g, w.ln2_b, want_f32=True, want_bf16=False)
            z = torch.empty(x.shape[0], 4 * self.width, device=x.device, dtype=torch.float32) if tape is not None else None
            a = mm(h2, w.w_fc, ops.EPI_BIAS_GELU_F32, bias=w.b_fc, out2=z)
            x2 = mm(a, w.w_proj, ops.EPI_BIAS_RESID_F32, bias=w.b_proj, resid=x1, out=None if tape is not None else x1)
            if tape is not None:
                tape.blocks.append(BlockSaved(x=x, x1=x1, qkv=qkv, o=o, lse=lse, z=z))
            x = x2
        if tape is not None:
            tape.x_final = x
        return x

    def _backward_tf32(self, tape, g, inject, inject_grads):
        B, L, H = tape.B, tape.L, self.heads
        mm = ops.gemm_f32 if self.fp32 else ops.gemm_tf32
        for li in range(len(self.blocks) - 1, -1, -1):
            w, s = self.blocks[li], tape.blocks[li]
            dz = mm(g, w.w_proj_t, ops.EPI_DGELU_F32, aux=s.z)
            dh2 = mm(dz, w.w_fc_t, ops.EPI_F32)
            ops.layernorm_bwd(dh2, s.x1, w.ln2_g, g, None, accumulate=True)
            if self.fp32:
                do = mm(g, w.w_out_t, ops.EPI_F32)
                dqkv = ops.attn_bwd_f32(s.qkv, s.o, do, s.lse, B, L, H, self.causal)
            else:
                do = mm(g, w.w_out_t, ops.EPI_BF16)
                dqkv = ops.attn_bwd(s.qkv, s.o, do, s.lse, B, L, H, self.causal, f32=True)
            dh1 = mm(dqkv, w.w_in_t, ops.EPI_F32)
            ops.layernorm_bwd(dh1, s.x, w.ln1_g, g, None, accumulate=True)
            if inject is not None and li != 0 and li in inject["layers"] and inject_grads is not None:
                inject_grads[li] = ops.sum_prompt_rows(g, inject["sel"], B, L, inject["P"], inject["table"].shape[0], self.width)
        return g

    # -------------------------------------------------------------------------------------------- backward
    def backward(self, tape: TowerTape, g: torch.Tensor, g_bf16: torch.Tensor, inject: Optional[dict] = None,
                 inject_grads: Optional[dict] = None) -> torch.Tensor:
        """g fp32 [B*L, D] = d loss / d (tower output), updated in place down to d loss / d (tower input);
        g_bf16 is its 16-bit shadow (must match g on entry; bf16, or fp16 holding grad_scale * g for an fp16 tower).  With `inject`, inject_grads[l] receives
        sum_b g_l[b, 1:P+1] for every injected layer.  When the forward ran its last block on the read rows (tape.out_rows), g and g_bf16
        are [B, D] (gradient of those rows) and the [B*L, D] stream is allocated here: always use the RETURNED tensor."""
        B, L, H = tape.B, tape.L, self.heads
        if self.tf32:
            return self._backward_tf32(tape, g, inject, inject_grads)
        fuse_delta = self.fuse_delta and L <= 256
        delta_pool = torch.zeros(len(self.blocks), B * H * L, device=g.device, dtype=torch.float32) if fuse_delta else None
        for li in range(len(self.blocks) - 1, -1, -1):
            w, s = self.blocks[li], tape.blocks[li]
            dz = ops.gemm(g_bf16, w.w_proj_t, ops.EPI_DGELU_BF16, aux=s.z)
            dh2 = ops.gemm(dz, w.w_fc_t, self.epi_dh)
            ops.layernorm_bwd(dh2, s.x1, w.ln2_g, g, g_bf16, accumulate=True, grad_scale=self.grad_scale)
            rows_only = tape.out_rows is not None and li == len(self.blocks) - 1
            if fuse_delta and not rows_only:
                do = ops.gemm_do_delta(g_bf16, w.w_out_t, s.o, delta_pool[li], L)
                dqkv = ops.attn_bwd(s.qkv, None, do, s.lse, B, L, H, self.causal, delta=delta_pool[li])
                dh1 = ops.gemm(dqkv, w.w_in_t, self.epi_dh)
                ops.layernorm_bwd(dh1, s.x, w.ln1_g, g, g_bf16, accumulate=True, grad_scale=self.grad_scale)
                if inject is not None and li != 0 and li in inject["layers"] and inject_grads is not None:
                    inject_grads[li] = ops.sum_prompt_rows(g, inject["sel"], B, L, inject["P"], inject["table"].shape[0], self.width)
                continue
            do = ops.gemm(g_bf16, w.w_out_t, ops.EPI_BF16)
            if tape.out_rows is not None and li == len(self.blocks) - 1:
                # the last block ran on the read rows: g / g_bf16 are [B, D] up to here; dk / dv of every position (and dq of the read
                # rows) come from the one-row attention backward, which also drops the residual-path gradient into the full stream
                gf = torch.zeros(B * L, self.width, device=g.device, dtype=torch.float32)
                dqkv = ops.attn_rowq_bwd(s.qkv, tape.out_rows, do, B, L, H, self.causal, g_rows=g, g=gf)
                g, g_bf16 = gf, torch.empty(B * L, self.width, device=g.device, dtype=self.half)
            else:
                dqkv = ops.attn_bwd(s.qkv, s.o, do, s.lse, B, L, H, self.causal)
            dh1 = ops.gemm(dqkv, w.w_in_t, self.epi_dh)
            ops.layernorm_bwd(dh1, s.x, w.ln1_g, g, g_bf16, accumulate=True, grad_scale=self.grad_scale)
            if inject is not None and li != 0 and li in inject["layers"] and inject_grads is not None:
                n_tables = inject["table"].shape[0]
                inject_grads[li] = ops.sum_prompt_rows(g, inject["sel"], B, L, inject["P"], n_tables, self.width)
        return g


# ------------------------------------------------------------------------------------------------ image encoder
# fp16 since round 2: against the real reference's fp32 run the image features land at 2.4e-4 (bf16: 1.9e-3), the logits at 1.0e-3 Frobenius
# (2.4e-3) and the visual prompt gradients at 7e-5 (4e-4) at B = 64 (profiles/r2_parity_margins.txt) at the same tensor rate
DEFAULT_VISION_PRECISION = "fp16"


class VisionEngine:
    """VisionTransformer.forward (model.py:227-259) on the kernels: im2col -> patch GEMM -> assemble(+ln_pre) -> tower -> head."""

    def __init__(self, sd: Dict[str, torch.Tensor], dev, prefix: str = "visual.", need_grad: bool = True, precision: Optional[str] = None):
        """precision: 'bf16' or 'fp16' GEMM / attention operands (fp32 accumulation, residual stream, LayerNorm and head either way);
        None = $LPI_VISION_PRECISION or the default below.  fp16 carries 3 more mantissa bits at the same tensor rate and is the
        reference's own GPU dtype (convert_weights, model.py:394-415); its gradient stream is scaled by 2^10 (Tower.grad_scale)."""
        import os

        precision = precision or os.environ.get("LPI_VISION_PRECISION", DEFAULT_VISION_PRECISION)
        if precision not in ("bf16", "fp16", "fp32"):
            raise ValueError(f"vision precision must be 'bf16', 'fp16' or 'fp32' (parity mode), got {precision!r}")
        self.precision = precision
        self.half = {"fp16": torch.float16, "fp32": torch.float32}.get(precision, torch.bfloat16)
        w = sd[prefix + "conv1.weight"]
        self.width, _, self.patch, _ = w.shape
        self.conv_w = _half(w.reshape(self.width, -1), dev, self.half)
        self.cls = _f32(sd[prefix + "class_embedding"], dev)
        self.pos = _f32(sd[prefix + "positional_embedding"], dev)
        self.ln_pre = (_f32(sd[prefix + "ln_pre.weight"], dev), _f32(sd[prefix + "ln_pre.bias"], dev))
        self.ln_post = (_f32(sd[prefix + "ln_post.weight"], dev), _f32(sd[prefix + "ln_post.bias"], dev))
        self.proj = _f32(sd[prefix + "proj"], dev)
        self.tower = Tower(sd, prefix + "transformer.resblocks.", self.width // 64, False, dev, need_grad, precision)
        self.n_patch = self.pos.shape[0] - 1
        self.dev = dev

    def patch_embed(self, images: torch.Tensor) -> torch.Tensor:
        """conv1 (kernel = stride = patch, no bias, model.py:228-231) as im2col + one tensor-core GEMM -> [B * n_patch, D] fp32."""
        patches = ops.im2col_patches(images.contiguous(), self.patch, self.half)
        if self.precision == "fp32":
            return ops.gemm_f32(patches, self.conv_w, ops.EPI_F32)
        return ops.gemm(patches, self.conv_w, ops.EPI_F32)

    def forward(self, images: torch.Tensor, prompt_table: Optional[torch.Tensor] = None, sel: Optional[torch.Tensor] = None,
                tape: Optional[dict] = None, inject_layers: Sequence[int] = (), factors=None, patch_emb: Optional[torch.Tensor] = None,
                centers: Optional[torch.Tensor] = None):
        """images [B,3,R,R] fp32; prompt_table [T, Lp, P, D] fp32 (layer 0 enters the token sequence, model.py:240-248) or None;
        sel int32[B] picks the table per sample (None = table 0).  Returns (L2-normalised features, raw projection z), both [B, E] fp32.

        factors = (dim_1_share [T, Lp, r], dim_2_visual [T, P, r], dim_3_visual [T, D, r], scale): the prompt rows are reconstructed inside
        the assembly kernel (DecomposedPrompt.forward fused with the token concat, prompts.py:38-57 + model.py:240-248) -- no table is
        materialised for the token sequence; prompt_table is then only needed for opt-in deep injection (inject_layers).
        patch_emb: the output of patch_embed(images) when the caller runs several passes over the same images (sprompt.py:336-351 +
        slinet.py:212-220 run the ViT twice per evaluation image).  centers [T, C, E]: the head also returns each sample's task id
        (third return value, int32 [B]; sprompt.py:336-351)."""
        B = images.shape[0]
        D = self.width
        pe = patch_emb if patch_emb is not None else self.patch_embed(images)
        layer0 = None
        if factors is not None:
            P = factors[1].shape[1]
            x = ops.assemble_vision_factors(pe, self.cls, self.pos, factors, sel, self.ln_pre[0], self.ln_pre[1], B, self.n_patch, D)
        else:
            P = 0 if prompt_table is None else prompt_table.shape[2]
            layer0 = None if prompt_table is None else prompt_table[:, 0].contiguous()
            x = ops.assemble_vision(pe, self.cls, self.pos, layer0, sel, self.ln_pre[0], self.ln_pre[1], B, self.n_patch, P, D)
        L = 1 + P + self.n_patch
        inject = None
        if len(inject_layers) > 0 and P > 0:
            if prompt_table is None:
                raise ValueError("deep-prompt injection (inject_layers) needs the reconstructed prompt_table as well as the factors")
            inject = {"layers": set(int(l) for l in inject_layers), "table": prompt_table, "sel": sel, "P": P}
        ttape = TowerTape(B, L) if tape is not None else None
        rows = torch.arange(B, device=x.device, dtype=torch.int32) * L          # the CLS rows: all the head reads (model.py:254)
        x = self.tower.forward(x, B, L, ttape, inject, out_rows=rows)
        if self.tower.rows_only(L):
            rows = torch.arange(B, device=x.device, dtype=torch.int32)          # x is [B, D] already
        task_id = None
        if centers is not None:
            feat, z, task_id = ops.head_fwd_select(x, rows, self.ln_post[0], self.ln_post[1], self.proj, centers)
        else:
            feat, z = ops.head_fwd(x, rows, self.ln_post[0], self.ln_post[1], self.proj)
        if tape is not None:
            n_tables = 0 if P == 0 else (factors[0].shape[0] if factors is not None else prompt_table.shape[0])
            n_layers = 0 if P == 0 else (factors[0].shape[1] if factors is not None else prompt_table.shape[1])
            tape.update(dict(tower=ttape, rows=rows, z=z, x=x, layer0=layer0, factors=factors, sel=sel, P=P, L=L, B=B, inject=inject,
                             n_tables=n_tables, Lp=n_layers))
        return (feat, z, task_id) if centers is not None else (feat, z)

    def select_and_encode(self, images: torch.Tensor, centers: torch.Tensor, prompt_table: Optional[torch.Tensor] = None, factors=None,
                          inject_layers: Sequence[int] = ()):
        """The evaluation path of one image batch (sprompt.py:456-470: get_visual_task_id -> visual_interface) with the patch embedding
        computed ONCE for both ViT passes and the task-id selection produced by the head kernel of the un-prompted pass.
        -> (prompted features [B, E], sel int32 [B], un-prompted features [B, E])"""
        pe = self.patch_embed(images)
        f0, _, sel = self.forward(images, None, None, None, (), None, pe, centers)
        f, _ = self.forward(images, prompt_table, sel, None, inject_layers, factors, pe)
        return f, sel, f0

    def backward(self, tape: dict, dfeat: Optional[torch.Tensor], dz: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """d loss / d prompt_table [T, Lp, P, D] (zero for layers that never entered the encoder).  dfeat = gradient wrt the
        normalised feature, dz = gradient wrt the raw projection (either may be None)."""
        B, L, D, P = tape["B"], tape["L"], self.width, tape["P"]
        dev = tape["x"].device
        n_rows = tape["x"].shape[0]          # B when the last block ran on the read rows only, else B * L
        g = torch.zeros(n_rows, D, device=dev, dtype=torch.float32)
        gb = torch.zeros(n_rows, D, device=dev, dtype=self.tower.half)
        ops.head_bwd(None if dfeat is None else dfeat.contiguous(), None if dz is None else dz.contiguous(), tape["z"], tape["x"],
                     tape["rows"], self.ln_post[0], self.proj, g, gb, grad_scale=self.tower.grad_scale)
        inj_grads = {}
        g = self.tower.backward(tape["tower"], g, gb, tape["inject"], inj_grads)
        if P == 0:
            return None
        G = torch.zeros(tape["n_tables"], tape["Lp"], P, D, device=dev, dtype=torch.float32)
        if tape.get("factors") is not None:
            G[:, 0] = ops.assemble_vision_factors_bwd(g, tape["factors"], tape["sel"], self.ln_pre[0], B, L, tape["n_tables"], D)
        else:
            G[:, 0] = ops.assemble_vision_bwd(g, tape["layer0"], tape["sel"], self.ln_pre[0], B, L, P, tape["n_tables"], D)
        for l, gl in inj_grads.items():
            G[:, l] = gl
        return G


# ------------------------------------------------------------------------------------------------ text encoder
class TextEngine:
    """PromptLearner splice + TextEncoder.forward (prompt_learner.py:133-163, 52-63) on the kernels."""

    def __init__(self, sd: Dict[str, torch.Tensor], dev, need_grad: bool = True, precision: str = "fp16"):
        """precision 'fp16' (default): fp16 GEMM / attention operands with fp32 accumulation -- the 10-bit mantissa the text tower needs
        (with bf16 operands the text-side prompt gradients sit at the 2e-2 parity limit, measured 2.1-2.2e-2)
        at the full kind::f16 tensor rate; it is also the reference's own GPU dtype (convert_weights, model.py:394-415).  The fp16
        gradient stream is scaled by 2^10 (Tower.grad_scale).  'tf32' = fp32 operands on the half-rate TF32 path (same mantissa,
        twice the bytes; the round-1 default before the fp16 kernels existed), 'bf16' for throughput studies."""
        self.emb = _f32(sd["token_embedding.weight"], dev)
        self.pos = _f32(sd["positional_embedding"], dev)
        self.ln_final = (_f32(sd["ln_final.weight"], dev), _f32(sd["ln_final.bias"], dev))
        self.proj = _f32(sd["text_projection"], dev)
        self.width = self.emb.shape[1]
        self.tower = Tower(sd, "transformer.resblocks.", self.width // 64, True, dev, need_grad, precision)
        self.context_length = self.pos.shape[0]
        self.trim_padding = True          # see forward(): skip the positions after the batch's last EOT
        self.zero_pos = torch.zeros_like(self.pos)
        self.dev = dev

    def forward(self, tokens: torch.Tensor, prompt_table: Optional[torch.Tensor] = None, sel: Optional[torch.Tensor] = None,
                tape: Optional[dict] = None, inject_layers: Sequence[int] = (), text_len: Optional[int] = None, factors=None):
        """tokens int64 [B, 77] on the device; prompt_table [T, Lp, P, Dt] fp32 (layer 0 is spliced over positions 1..P) or None.
        factors = (dim_1_share [T, Lp, r], dim_2_textual [T, P, r], dim_3_textual [T, Dt, r], scale): the context rows are reconstructed
        inside the splice kernel instead of being read from a table (see VisionEngine.forward).

        text_len (host int, or the `lpi_text_len` attribute the tokenizer attaches to its output): number of leading token positions
        that contain every caption's EOT.  The mask is causal (model.py:347-353) and the head reads the EOT row only, so positions
        after the last EOT influence neither the features nor any gradient: the tower then runs on [B, text_len] instead of [B, 77]
        (output-exact; COCO-length captions end near position 40).  Unknown -> all positions, no host sync."""
        B, L = tokens.shape
        D = self.width
        if text_len is None:
            text_len = getattr(tokens, "lpi_text_len", None)
        n_prompt = factors[1].shape[1] if factors is not None else (0 if prompt_table is None else prompt_table.shape[2])
        if text_len is not None and self.trim_padding:
            Lt = max(1 + n_prompt, min(int(text_len), L))
            if Lt < L:
                tokens, L = tokens[:, :Lt], Lt
        P = n_prompt
        if factors is not None:
            x = ops.assemble_text_factors(self.emb, tokens.contiguous(), self.pos, factors, sel, B, L, D)
        else:
            layer0 = None if prompt_table is None else prompt_table[:, 0].contiguous()
            x = ops.assemble_text(self.emb, tokens.contiguous(), self.pos, layer0, sel, B, L, P, D)
        inject = None
        if len(inject_layers) > 0 and P > 0:
            if prompt_table is None:
                raise ValueError("deep-prompt injection (inject_layers) needs the reconstructed prompt_table as well as the factors")
            inject = {"layers": set(int(l) for l in inject_layers), "table": prompt_table, "sel": sel, "P": P}
        ttape = TowerTape(B, L) if tape is not None else None
        rows = (torch.arange(B, device=x.device, dtype=torch.int64) * L + tokens.argmax(dim=-1)).to(torch.int32)   # EOT row
        x = self.tower.forward(x, B, L, ttape, inject, out_rows=rows)
        if self.tower.rows_only(L):
            rows = torch.arange(B, device=x.device, dtype=torch.int32)          # x is [B, D] already
        feat, z = ops.head_fwd(x, rows, self.ln_final[0], self.ln_final[1], self.proj)
        if tape is not None:
            n_tables = 0 if P == 0 else (factors[0].shape[0] if factors is not None else prompt_table.shape[0])
            n_layers = 0 if P == 0 else (factors[0].shape[1] if factors is not None else prompt_table.shape[1])
            tape.update(dict(tower=ttape, rows=rows, z=z, x=x, sel=sel, P=P, L=L, B=B, inject=inject, n_tables=n_tables, Lp=n_layers))
        return feat, z

    def backward(self, tape: dict, dfeat: Optional[torch.Tensor], dz: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        B, L, D, P = tape["B"], tape["L"], self.width, tape["P"]
        dev = tape["x"].device
        n_rows = tape["x"].shape[0]          # B when the last block ran on the read rows only, else B * L
        g = torch.zeros(n_rows, D, device=dev, dtype=torch.float32)
        gb = torch.zeros(n_rows, D, device=dev, dtype=self.tower.half)
        ops.head_bwd(None if dfeat is None else dfeat.contiguous(), None if dz is None else dz.contiguous(), tape["z"], tape["x"],
                     tape["rows"], self.ln_final[0], self.proj, g, gb, grad_scale=self.tower.grad_scale)
        inj_grads = {}
        g = self.tower.backward(tape["tower"], g, gb, tape["inject"], inj_grads)
        if P == 0:
            return None
        G = torch.zeros(tape["n_tables"], tape["Lp"], P, D, device=dev, dtype=torch.float32)
        G[:, 0] = ops.sum_prompt_rows(g, tape["sel"], B, L, P, tape["n_tables"], D)
        for l, gl in inj_grads.items():
            G[:, l] = gl
        return G

    # ---- the reference's module boundary: TextEncoder.forward(prompts [B,77,D] already embedded + spliced, tokenized, ...)
    def forward_embedded(self, prompts: torch.Tensor, tokens: torch.Tensor, prompt_table: Optional[torch.Tensor] = None,
                         sel: Optional[torch.Tensor] = None, tape: Optional[dict] = None, inject_layers: Sequence[int] = ()):
        B, L, D = prompts.shape
        rows_all = torch.arange(B * L, device=prompts.device, dtype=torch.int64)
        x = ops.assemble_text(prompts.reshape(B * L, D), rows_all.view(B, L), self.pos, None, None, B, L, 0, D)    # + positional_embedding
        inject = None
        if prompt_table is not None and len(inject_layers) > 0:
            inject = {"layers": set(int(l) for l in inject_layers), "table": prompt_table, "sel": sel, "P": prompt_table.shape[2]}
        ttape = TowerTape(B, L) if tape is not None else None
        rows = (torch.arange(B, device=x.device, dtype=torch.int64) * L + tokens.argmax(dim=-1)).to(torch.int32)
        x = self.tower.forward(x, B, L, ttape, inject, out_rows=rows)
        if self.tower.rows_only(L):
            rows = torch.arange(B, device=x.device, dtype=torch.int32)
        feat, z = ops.head_fwd(x, rows, self.ln_final[0], self.ln_final[1], self.proj)
        if tape is not None:
            tape.update(dict(tower=ttape, rows=rows, z=z, x=x, sel=sel, L=L, B=B, inject=inject,
                             table_shape=None if prompt_table is None else tuple(prompt_table.shape)))
        return feat, z

    def backward_embedded(self, tape: dict, dfeat: Optional[torch.Tensor], dz: Optional[torch.Tensor] = None):
        """-> (d loss / d prompts [B*L, D], d loss / d prompt_table or None)"""
        B, L, D = tape["B"], tape["L"], self.width
        dev = tape["x"].device
        n_rows = tape["x"].shape[0]          # B when the last block ran on the read rows only, else B * L
        g = torch.zeros(n_rows, D, device=dev, dtype=torch.float32)
        gb = torch.zeros(n_rows, D, device=dev, dtype=self.tower.half)
        ops.head_bwd(None if dfeat is None else dfeat.contiguous(), None if dz is None else dz.contiguous(), tape["z"], tape["x"],
                     tape["rows"], self.ln_final[0], self.proj, g, gb, grad_scale=self.tower.grad_scale)
        inj_grads = {}
        g = self.tower.backward(tape["tower"], g, gb, tape["inject"], inj_grads)
        G = None
        if tape["table_shape"] is not None:
            G = torch.zeros(tape["table_shape"], device=dev, dtype=torch.float32)
            for l, gl in inj_grads.items():
                G[:, l] = gl
        return g, G
