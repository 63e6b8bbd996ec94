"""CLIP ViT-B/16 with prompt hooks -- the reference's module surface (retrieval/models/clip/model.py:154-391,
retrieval/models/clip/prompt_learner.py:43-226) over the CUDA engine.

The nn.Modules below hold the parameters under the reference's own names, so `state_dict()` keys, `named_parameters()`
(the freeze policy greps them, sprompt.py:229-237) and `load_state_dict` of an OpenAI CLIP checkpoint all behave as in
the reference; they are created in the reference's order with the reference's initialisers, so a shared torch seed gives the
same random-init model.  No module computes in PyTorch: every forward routes to `engine.VisionEngine` / `engine.TextEngine`
(liblpi_b200.so kernels, bf16 tensor-core GEMMs with fp32 accumulation), built lazily from the current parameter values and
rebuilt after `load_state_dict` / `.to()` / `refresh()`.  CUDA only; there is no CPU path.

Differences from the reference, all deliberate: per-layer deep-prompt injection is a parameter
(`inject_layers`, default () = the reference as shipped, where the branch at model.py:190 is dead code); `CLIP.encode_text`
works (the reference's raises TypeError, C7); weights are never downloaded -- pass a state_dict or keep the random init (C9).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from . import tokenizer as _tok
from ._lib import LpiError
from .autograd import SpliceFn, TextEmbeddedFn, TextEncodeFn, VisionEncodeFn, _as_table
from .engine import TextEngine, VisionEngine


class LayerNorm(nn.LayerNorm):
    """Parameter holder for ln_* (model.py:154-160); the fp32 LayerNorm itself runs in lpi_layernorm_fwd/bwd."""


class QuickGELU(nn.Module):
    """x * sigmoid(1.702 x) (model.py:163-165) -- fused into the c_fc GEMM epilogue; kept for the module tree."""


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None, layer_id=0):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask
        self.layer_id = layer_id


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.width, self.layers, self.heads = width, layers, heads
        self.causal = attn_mask is not None
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask, i) for i in range(layers)])


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int):
        super().__init__()
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(in_channels=3, out_channels=width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self.inject_layers: Tuple[int, ...] = ()
        self._engine: Optional[VisionEngine] = None

    def engine(self) -> VisionEngine:
        dev = self.proj.device
        if dev.type != "cuda":
            raise LpiError("lpi_b200 runs on a CUDA (sm_100a) device only; move the model with .cuda() first")
        if self._engine is None or self._engine.dev != dev:
            sd = {"visual." + k: v for k, v in self.state_dict().items()}
            self._engine = VisionEngine(sd, dev)
        return self._engine

    def refresh(self):
        self._engine = None

    def _load_from_state_dict(self, *a, **k):
        self._engine = None
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def encode(self, x: torch.Tensor, instance_tokens=None):
        """-> (L2-normalised features, raw projection), both [B, output_dim] fp32."""
        table, sel = _as_table(instance_tokens)
        return VisionEncodeFn.apply(self.engine(), x, table, sel, tuple(self.inject_layers))

    def forward(self, x: torch.Tensor, instance_tokens=None):
        """model.py:227-259: images [B,3,R,R], instance_tokens [B, Lp, P, width] or None -> [B, output_dim] (not normalised)."""
        return self.encode(x, instance_tokens)[1]


class CLIP(nn.Module):
    def __init__(self, embed_dim: int, image_resolution: int, vision_layers: int, vision_width: int, vision_patch_size: int,
                 context_length: int, vocab_size: int, transformer_width: int, transformer_heads: int, transformer_layers: int):
        super().__init__()
        if isinstance(vision_layers, (tuple, list)):
            raise LpiError("ModifiedResNet backbones are out of scope (the LPI config uses ViT-B/16 only)")
        self.context_length = context_length
        self.visual = VisionTransformer(input_resolution=image_resolution, patch_size=vision_patch_size, width=vision_width,
                                        layers=vision_layers, heads=vision_width // 64, output_dim=embed_dim)
        self.transformer = Transformer(width=transformer_width, layers=transformer_layers, heads=transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(self.context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.inject_layers: Tuple[int, ...] = ()
        self.text_precision = "fp16"          # 'fp16' (default), 'tf32' or 'bf16' -- see engine.TextEngine
        self._text_engine: Optional[TextEngine] = None
        self.initialize_parameters()

    def initialize_parameters(self):
        """model.py:318-345 -- note the vision blocks keep PyTorch's default init."""
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (self.transformer.width ** -0.5) * ((2 * self.transformer.layers) ** -0.5)
        attn_std = self.transformer.width ** -0.5
        fc_std = (2 * self.transformer.width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    def build_attention_mask(self):
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def text_engine(self) -> TextEngine:
        dev = self.text_projection.device
        if dev.type != "cuda":
            raise LpiError("lpi_b200 runs on a CUDA (sm_100a) device only; move the model with .cuda() first")
        if self._text_engine is None or self._text_engine.dev != dev:
            sd = {k: v for k, v in self.state_dict().items() if not k.startswith("visual.")}
            self._text_engine = TextEngine(sd, dev, precision=self.text_precision)
        return self._text_engine

    def refresh(self):
        self._text_engine = None
        self.visual.refresh()

    def _load_from_state_dict(self, *a, **k):
        self._text_engine = None
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._text_engine = None
        return super()._apply(fn, *a, **k)

    def encode_image(self, image):
        return self.visual(image)

    def encode_text(self, text):
        """Un-prompted text features for token ids [B, 77] (model.py:362-375; fixed)."""
        return TextEncodeFn.apply(self.text_engine(), text, None, None, ())[1]

    def forward(self, image, text):
        i_f, _ = self.visual.encode(image)
        t_f, _ = TextEncodeFn.apply(self.text_engine(), text, None, None, ())
        logit_scale = self.logit_scale.exp()
        logits_per_image = logit_scale * i_f @ t_f.t()
        return logits_per_image, logits_per_image.t()


def build_model(state_dict: dict) -> CLIP:
    """CLIP ViT from an OpenAI-format state_dict (hyper-parameters derived as model.py:418-441)."""
    vision_width = state_dict["visual.conv1.weight"].shape[0]
    vision_layers = len([k for k in state_dict if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    patch = state_dict["visual.conv1.weight"].shape[-1]
    grid = round((state_dict["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    embed_dim = state_dict["text_projection"].shape[1]
    ctx_len = state_dict["positional_embedding"].shape[0]
    vocab = state_dict["token_embedding.weight"].shape[0]
    tw = state_dict["ln_final.weight"].shape[0]
    tl = len(set(k.split(".")[2] for k in state_dict if k.startswith("transformer.resblocks")))
    model = CLIP(embed_dim, patch * grid, vision_layers, vision_width, patch, ctx_len, vocab, tw, tw // 64, tl)
    sd = {k: v for k, v in state_dict.items() if k not in ("input_resolution", "context_length", "vocab_size")}
    model.load_state_dict(sd)
    return model.eval()


def load_clip_to_cpu(args) -> CLIP:
    """prompt_learner.py:10-40 downloads ViT-B/16 from the network; here the weights come from `args['clip_state_dict']`
    (a state_dict or a path to one) or, failing that, a random-init ViT-B/16 (there is no network on the box)."""
    src = args.get("clip_state_dict") if isinstance(args, dict) else None
    if isinstance(src, str):
        src = torch.load(src, map_location="cpu", weights_only=True)      # a state_dict: tensors only
    if src is not None:
        return build_model(src)
    return CLIP(512, 224, 12, 768, 16, 77, 49408, 512, 8, 12).eval()


class TextEncoder(nn.Module):
    def __init__(self, clip_model: CLIP):
        super().__init__()
        self.transformer = clip_model.transformer
        self.positional_embedding = clip_model.positional_embedding
        self.ln_final = clip_model.ln_final
        self.text_projection = clip_model.text_projection
        self.dtype = clip_model.dtype
        object.__setattr__(self, "_clip", clip_model)        # not registered: the engine lives on the CLIP module

    def encode(self, prompts, tokenized_prompts, textual_prompt=None):
        table, sel = _as_table(textual_prompt)
        return TextEmbeddedFn.apply(self._clip.text_engine(), prompts, tokenized_prompts, table, sel, tuple(self._clip.inject_layers))

    def forward(self, prompts, tokenized_prompts, textual_prompt):
        """prompt_learner.py:52-63: prompts [B,77,D] (embedded, context spliced), tokenized [B,77] -> [B,E] (not normalised).
        `textual_prompt` [B,Lp,P,D] is only consumed when inject_layers is non-empty (dead in the reference as shipped)."""
        return self.encode(prompts, tokenized_prompts, textual_prompt)[1]


class PromptLearner(nn.Module):
    def __init__(self, cfg, clip_model: CLIP):
        super().__init__()
        n_ctx = cfg.NCTX
        if cfg.CTXINIT:
            raise LpiError("CTXINIT is not used by the LPI config (prompt_learner.py:82-90)")
        if cfg.CLASS_TOKEN_POSITION != "end":
            raise LpiError("only CLASS_TOKEN_POSITION='end' (the LPI config) is implemented")
        self.clip_model = clip_model
        self.dtype = clip_model.dtype
        ctx_dim = clip_model.ln_final.weight.shape[0]
        ctx_vectors = torch.empty(n_ctx, ctx_dim, dtype=self.dtype)
        nn.init.normal_(ctx_vectors, std=0.02)
        self.prompt_prefix = " ".join(["X"] * n_ctx)
        self.ctx = nn.Parameter(ctx_vectors)
        self.n_ctx = n_ctx
        self.n_cls = None
        self.class_token_position = cfg.CLASS_TOKEN_POSITION
        self._cache = {}

    @property
    def device(self):
        return self.clip_model.token_embedding.weight.device

    def tokenize(self, captions: Sequence[str]) -> torch.Tensor:
        """'X X ... X <caption>.' -> ids [B,77] on the device; each distinct caption is tokenised once and cached
        (the reference re-runs the Python BPE twice per call, prompt_learner.py:130-132)."""
        rows = []
        for c in captions:
            t = self._cache.get(c)
            if t is None:
                t = _tok.tokenize(self.prompt_prefix + " " + c + ".")[0]
                self._cache[c] = t
            rows.append(t)
        host = torch.stack(rows)
        out = host.to(self.device, non_blocking=True)
        # known on the host for free: lets TextEngine.forward skip the positions after the batch's last EOT without a device sync
        out.lpi_text_len = int(host.argmax(dim=-1).max()) + 1
        return out

    def extract_vector(self, captions):
        """prompt_learner.py:118-126: embeddings WITHOUT the context splice (the raw 'X' embeddings stay)."""
        self.n_cls = len(captions)
        tokenized = captions if isinstance(captions, torch.Tensor) else self.tokenize(captions)
        return SpliceFn.apply(self.clip_model.text_engine(), tokenized, None, None), tokenized

    def forward(self, captions, ctx):
        """prompt_learner.py:128-163: captions (list[str], or pre-tokenised ids [B,77]) and ctx [B,P,D] / [P,D] / None
        -> (prompts [B,77,D], tokenized [B,77])."""
        self.n_cls = len(captions)
        tokenized = captions if isinstance(captions, torch.Tensor) else self.tokenize(captions)
        if ctx is None:
            ctx = self.ctx
        if ctx.dim() == 2:
            table, sel = ctx.unsqueeze(0), None
        elif ctx.stride(0) == 0 or ctx.shape[0] == 1:
            table, sel = ctx[0:1], None
        else:
            table, sel = ctx, torch.arange(ctx.shape[0], device=ctx.device, dtype=torch.int32)
        return SpliceFn.apply(self.clip_model.text_engine(), tokenized, table, sel), tokenized


class cfgc(object):
    backbonename = "ViT-B/16"
    NCTX = 16
    CTXINIT = ""
    CSC = False
    CLASS_TOKEN_POSITION = "end"
