"""CPU proof, on the oracle restatement of the reference's text encoder (pinned against the real reference by test_oracle_golden.py),
of the property the CUDA text tower relies on: under the causal mask (model.py:347-353) with the feature read at the EOT row
(prompt_learner.py:57-61), the token positions after the batch's last EOT influence neither the text feature nor the gradient
of the spliced context -- so running the tower on [B, text_len] instead of [B, 77] is output-exact."""
import torch

from lpi_b200 import synthetic as S
from oracle import lpi_oracle as O


def test_positions_after_the_last_eot_are_dead(clip_sd):
    tokens = S.make_tokens(3, 5)
    text_len = int(tokens.argmax(-1).max()) + 1
    assert 18 <= text_len < 77
    g = torch.Generator().manual_seed(0)
    ctx = (torch.randn(16, 512, generator=g) * 0.5).requires_grad_(True)
    full = O.text_forward(clip_sd, tokens, ctx)
    (full * torch.linspace(-1, 1, 512)).sum().backward()
    grad_full = ctx.grad.clone()
    ctx.grad = None
    sd_trim = dict(clip_sd)
    sd_trim["positional_embedding"] = clip_sd["positional_embedding"][:text_len]
    trim = O.text_forward(sd_trim, tokens[:, :text_len], ctx)
    (trim * torch.linspace(-1, 1, 512)).sum().backward()
    assert torch.allclose(trim, full, rtol=1e-5, atol=1e-6)
    # fp32 matmuls of different shapes sum in different orders: compare in norm
    assert float((ctx.grad - grad_full).norm() / grad_full.norm()) < 1e-4
    # and a perturbation of a padded position changes nothing at all in the full-length run
    poked = tokens.clone()
    poked[:, text_len:] = 1234                                   # (ids below the EOT id, so argmax still finds the EOT)
    assert torch.allclose(O.text_forward(clip_sd, poked, ctx), full, rtol=1e-5, atol=1e-6)
