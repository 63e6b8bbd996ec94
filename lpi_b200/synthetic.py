"""Deterministic synthetic weights / inputs shared by tests, bench and smoke.

There is no network (the reference downloads CLIP weights at ctor time,
retrieval/models/clip/prompt_learner.py:10-40), so every run uses a seeded random-init
ViT-B/16 CLIP whose tensors live under the reference's own state_dict key names
(retrieval/models/clip/model.py:262-345).  Generated on the CPU with a `torch.Generator`
so the container, the GPU box and the golden fixtures all see bit-identical values.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

VIT_B16 = dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=16,
               context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8, transformer_layers=12)

FACTOR_NAMES = ("dim_1_share", "dim_2_visual", "dim_2_textual", "dim_3_visual", "dim_3_textual")

SOT_TOKEN = 49406
EOT_TOKEN = 49407
X_TOKEN = 343        # BPE id of "x</w>" -- the "X" placeholder of PromptLearner (prompt_learner.py:101)
DOT_TOKEN = 269      # "."


def make_clip_state_dict(seed: int = 0, cfg: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """fp32 CPU tensors keyed like `CLIP.state_dict()`.  Scales follow CLIP.initialize_parameters
    (model.py:318-345) for both towers, with non-trivial LayerNorm affine terms and biases so a
    dropped bias / swapped gamma-beta cannot hide."""
    c = dict(VIT_B16)
    if cfg:
        c.update(cfg)
    g = torch.Generator().manual_seed(1000003 * seed + 17)

    def rn(*shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=g) * std + mean

    sd: Dict[str, torch.Tensor] = {}
    vw, tw, E = c["vision_width"], c["transformer_width"], c["embed_dim"]
    ps = c["vision_patch_size"]
    grid = c["image_resolution"] // ps

    def tower(pfx, width, layers):
        proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
        attn_std = width ** -0.5
        fc_std = (2 * width) ** -0.5
        for i in range(layers):
            p = f"{pfx}resblocks.{i}."
            sd[p + "attn.in_proj_weight"] = rn(3 * width, width, std=attn_std)
            sd[p + "attn.in_proj_bias"] = rn(3 * width, std=0.02)
            sd[p + "attn.out_proj.weight"] = rn(width, width, std=proj_std)
            sd[p + "attn.out_proj.bias"] = rn(width, std=0.02)
            sd[p + "ln_1.weight"] = rn(width, std=0.1, mean=1.0)
            sd[p + "ln_1.bias"] = rn(width, std=0.05)
            sd[p + "mlp.c_fc.weight"] = rn(4 * width, width, std=fc_std)
            sd[p + "mlp.c_fc.bias"] = rn(4 * width, std=0.02)
            sd[p + "mlp.c_proj.weight"] = rn(width, 4 * width, std=proj_std)
            sd[p + "mlp.c_proj.bias"] = rn(width, std=0.02)
            sd[p + "ln_2.weight"] = rn(width, std=0.1, mean=1.0)
            sd[p + "ln_2.bias"] = rn(width, std=0.05)

    sd["visual.conv1.weight"] = rn(vw, 3, ps, ps, std=(3 * ps * ps) ** -0.5)
    sd["visual.class_embedding"] = rn(vw, std=vw ** -0.5)
    sd["visual.positional_embedding"] = rn(grid * grid + 1, vw, std=vw ** -0.5)
    sd["visual.ln_pre.weight"] = rn(vw, std=0.1, mean=1.0)
    sd["visual.ln_pre.bias"] = rn(vw, std=0.05)
    tower("visual.transformer.", vw, c["vision_layers"])
    sd["visual.ln_post.weight"] = rn(vw, std=0.1, mean=1.0)
    sd["visual.ln_post.bias"] = rn(vw, std=0.05)
    sd["visual.proj"] = rn(vw, E, std=vw ** -0.5)
    tower("transformer.", tw, c["transformer_layers"])
    sd["token_embedding.weight"] = rn(c["vocab_size"], tw, std=0.02)
    sd["positional_embedding"] = rn(c["context_length"], tw, std=0.01)
    sd["ln_final.weight"] = rn(tw, std=0.1, mean=1.0)
    sd["ln_final.bias"] = rn(tw, std=0.05)
    sd["text_projection"] = rn(tw, E, std=tw ** -0.5)
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07))
    return sd


def make_prompt_factors(seed: int, layer_num: int = 9, prompt_num: int = 16, dv: int = 768, dt: int = 512,
                        r: int = 4) -> Dict[str, torch.Tensor]:
    """N(0, 0.5^2) factors as DecomposedPrompt.__init__ (prompts.py:21-25)."""
    g = torch.Generator().manual_seed(7919 * seed + 5)
    shapes = dict(dim_1_share=(layer_num, r), dim_2_visual=(prompt_num, r), dim_2_textual=(prompt_num, r),
                  dim_3_visual=(dv, r), dim_3_textual=(dt, r))
    return {k: torch.randn(*s, generator=g) * 0.5 for k, s in shapes.items()}


def make_images(batch: int, seed: int, res: int = 224) -> torch.Tensor:
    g = torch.Generator().manual_seed(1234 + seed)
    return torch.randn(batch, 3, res, res, generator=g)


def make_tokens(batch: int, seed: int, n_ctx: int = 16, context_length: int = 77, min_words: int = 8,
                max_words: int = 20, vocab_lo: int = 320, vocab_hi: int = 40000) -> torch.Tensor:
    """Synthetic pre-tokenised captions shaped like PromptLearner's: SOT, n_ctx placeholder ids,
    8-20 word ids, '.', EOT, zero padding (prompt_learner.py:128-132, clip.py:205-219).
    EOT (49407) is the row maximum so argmax finds it, as in the reference."""
    g = torch.Generator().manual_seed(4321 + seed)
    tok = torch.zeros(batch, context_length, dtype=torch.int64)
    for b in range(batch):
        n = int(torch.randint(min_words, max_words + 1, (1,), generator=g))
        words = torch.randint(vocab_lo, vocab_hi, (n,), generator=g)
        row = [SOT_TOKEN] + [X_TOKEN] * n_ctx + words.tolist() + [DOT_TOKEN, EOT_TOKEN]
        tok[b, : len(row)] = torch.tensor(row)
    return tok


_WORDS: List[str] = (
    "a an the man woman child dog cat bird horse train bus car truck boat plane bike street road field park beach "
    "table chair bench kitchen room plate bowl cup pizza cake sandwich fruit apple banana orange broccoli carrot "
    "red blue green yellow white black brown large small young old tall wooden metal glass standing sitting riding "
    "holding eating playing walking running flying parked looking next near under over behind front with and of in on "
    "two three several many group people person water snow grass sky tree building window door clock sign light "
    "tennis baseball skateboard surfboard frisbee kite umbrella phone laptop keyboard book bed couch toilet sink"
).split()


def make_captions(batch: int, seed: int, min_words: int = 8, max_words: int = 20) -> List[str]:
    """ASCII captions of 8-20 common words (BPE length <= 77 guaranteed)."""
    g = torch.Generator().manual_seed(9876 + seed)
    out = []
    for _ in range(batch):
        n = int(torch.randint(min_words, max_words + 1, (1,), generator=g))
        idx = torch.randint(0, len(_WORDS), (n,), generator=g).tolist()
        out.append(" ".join(_WORDS[i] for i in idx))
    return out


def make_retrieval_set(n_img: int = 1000, caps_per_img: int = 5, dim: int = 512, n_tasks: int = 5, seed: int = 2,
                       signal: float = 0.15):
    """Flickr30K-shaped synthetic eval set: unit image embeddings,
    `caps_per_img` noisy captions each; un-saturated Recall@1/5/10."""
    g = torch.Generator().manual_seed(1234 + seed)
    img = torch.randn(n_img, dim, generator=g)
    img = img / img.norm(dim=-1, keepdim=True)
    txt = signal * img.repeat_interleave(caps_per_img, 0) + torch.randn(n_img * caps_per_img, dim, generator=g) / math.sqrt(dim)
    txt = txt / txt.norm(dim=-1, keepdim=True)
    img2txt = {i: list(range(caps_per_img * i, caps_per_img * (i + 1))) for i in range(n_img)}
    txt2img = {t: t // caps_per_img for t in range(n_img * caps_per_img)}
    cat_i = [i % n_tasks for i in range(n_img)]
    cat_t = [(t // caps_per_img) % n_tasks for t in range(n_img * caps_per_img)]
    return img, txt, img2txt, txt2img, cat_i, cat_t


# ------------------------------------------------------------------------------------------------
# Large-gallery sweep (BASELINE.json configs[4])
# ------------------------------------------------------------------------------------------------
GALLERY_BLOCK = 65536      # rows are generated in fixed blocks keyed on the block index, so any
                           # sharding of the gallery sees bit-identical rows


def gallery_gt(n_queries: int, n_gallery: int) -> torch.Tensor:
    """Ground-truth gallery row of query q: (q * 199999) mod N (int64, host)."""
    return (torch.arange(n_queries, dtype=torch.int64) * 199999) % n_gallery


def _gallery_block(block: int, dim: int, seed: int, device) -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed * 1000003 + 7 * block + 1)
    x = torch.randn(GALLERY_BLOCK, dim, generator=g, device=device, dtype=torch.float32)
    return x / x.norm(dim=-1, keepdim=True)


def make_gallery_shard(n_gallery: int, lo: int, hi: int, n_queries: int, dim: int = 512, seed: int = 5,
                       device="cuda", signal: float = 0.2):
    """Rows [lo, hi) of the synthetic gallery as bf16 plus ALL queries (bf16, identical on every rank).

    gallery row i = normalize(randn) (block-keyed generator); query q = normalize(signal * G[gt(q)] + noise/sqrt(dim)).
    Every rank walks all blocks (cheap: the generator runs at HBM speed) and keeps its own slice, so no
    communication is needed to build the workload.  Values depend on the device type of the generator
    (CPU and CUDA Philox streams differ) but not on the sharding."""
    device = torch.device(device)
    gt = gallery_gt(n_queries, n_gallery).to(device)
    shard = torch.empty(hi - lo, dim, device=device, dtype=torch.bfloat16)
    q_src = torch.empty(n_queries, dim, device=device, dtype=torch.float32)
    n_blocks = (n_gallery + GALLERY_BLOCK - 1) // GALLERY_BLOCK
    for b in range(n_blocks):
        b_lo, b_hi = b * GALLERY_BLOCK, min(n_gallery, (b + 1) * GALLERY_BLOCK)
        x = _gallery_block(b, dim, seed, device)[: b_hi - b_lo]
        s_lo, s_hi = max(lo, b_lo), min(hi, b_hi)
        if s_lo < s_hi:
            shard[s_lo - lo: s_hi - lo] = x[s_lo - b_lo: s_hi - b_lo].to(torch.bfloat16)
        sel = ((gt >= b_lo) & (gt < b_hi)).nonzero().flatten()
        if sel.numel():
            # queries are built from the bf16-rounded gallery row, i.e. from what the scorer sees
            q_src[sel] = x[gt[sel] - b_lo].to(torch.bfloat16).float()
    g = torch.Generator(device=device).manual_seed(seed * 1000003 + 999331)
    noise = torch.randn(n_queries, dim, generator=g, device=device, dtype=torch.float32) / math.sqrt(dim)
    q = signal * q_src + noise
    q = (q / q.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    return shard, q, gt.cpu()
