"""DecomposedPrompt -- same constructor, parameter names and forward() as the reference
(retrieval/models/prompts/prompts.py:4-57), reconstruction and its backward on the prompt kernels."""
from __future__ import annotations

import torch
from torch import nn

from .autograd import DecomposedPromptFn


class DecomposedPrompt(nn.Module):
    """`prompt_depth_vis` / `prompt_depth_text` are really the embedding WIDTHS (768 / 512, slinet.py:46).
    Parameters: dim_1_share [layer_num, r], dim_2_visual / dim_2_textual [prompt_num, r], dim_3_visual [Dv, r],
    dim_3_textual [Dt, r], all N(0, 0.5^2) (prompts.py:21-25).  forward() -> (vis [L,P,Dv], txt [L,P,Dt]) = mean over r."""

    def __init__(self, layer_num, prompt_num, prompt_depth_vis, prompt_depth_text, r=4):
        super().__init__()
        self.d = r
        # same creation order and initialisers as the reference so a shared torch seed gives the same factors
        d1, d2v, d2t = torch.randn(layer_num, r), torch.randn(prompt_num, r), torch.randn(prompt_num, r)
        d3v, d3t = torch.rand(prompt_depth_vis, r), torch.rand(prompt_depth_text, r)
        self.dim_1_share = nn.Parameter(d1)
        self.dim_2_visual = nn.Parameter(d2v)
        self.dim_2_textual = nn.Parameter(d2t)
        self.dim_3_visual = nn.Parameter(d3v)
        self.dim_3_textual = nn.Parameter(d3t)
        for p in (self.dim_1_share, self.dim_2_visual, self.dim_2_textual, self.dim_3_visual, self.dim_3_textual):
            nn.init.normal_(p, std=0.5)
        self.scale = 1

    def interface(self):
        raise NotImplementedError("DecomposedPrompt.interface() is broken in the reference (prompts.py:29-36 reads attributes that "
                                  "do not exist); use forward()")

    def forward(self):
        vis, txt = DecomposedPromptFn.apply(self.dim_1_share, self.dim_2_visual, self.dim_2_textual, self.dim_3_visual,
                                            self.dim_3_textual)
        if self.scale != 1:
            vis, txt = vis * self.scale, txt * self.scale
        return vis, txt
