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


def test_real_reference_text_encoder_ignores_positions_after_the_eot():
    """The same property on the REAL reference classes (TextEncoder over the reference's Transformer with its causal mask,
    prompt_learner.py:43-63, model.py:347-353), on a small CLIP so the test stays in seconds: perturbing the embedded prompt at the
    positions after each caption's EOT changes neither the feature nor the gradient of the spliced context rows."""
    import pytest

    from oracle import reference_loader as RL

    if not RL.reference_available():
        pytest.skip("reference tree not mounted")
    ns = RL.load_reference()
    torch.manual_seed(5)
    clip = ns.clip_model.CLIP(64, 32, 2, 128, 16, 77, 1000, 128, 2, 3).eval()        # embed 64, text width 128, 2 heads, 3 layers
    enc = ns.prompt_learner.TextEncoder(clip)
    B, L, D = 4, 77, 128
    eot = torch.tensor([20, 33, 27, 40])
    tok = torch.randint(1, 900, (B, L))
    for b in range(B):
        tok[b, eot[b]] = 999                                     # the row maximum marks the EOT (clip.py:205-219)
        tok[b, eot[b] + 1:] = 0
    prompts = torch.randn(B, L, D, requires_grad=True)
    dummy = torch.zeros(B, 9, 16, D)
    w = torch.linspace(-1, 1, 64)
    f0 = enc(prompts, tok, dummy)
    (f0 * w).sum().backward()
    g0 = prompts.grad.clone()
    text_len = int(eot.max()) + 1
    assert float(g0[:, text_len:].abs().max()) == 0.0            # no gradient ever reaches a position after the last EOT
    for b in range(B):
        assert float(g0[b, eot[b] + 1:].abs().max()) == 0.0
    noisy = prompts.detach().clone()
    noisy[:, text_len:] += 10 * torch.randn(B, L - text_len, D)
    f1 = enc(noisy, tok, dummy)
    assert torch.equal(f1, f0)                                   # bit-identical: those positions are masked out of every EOT row
