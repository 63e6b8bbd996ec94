"""CPU checks of the drop-in surface: parameter / state_dict names, same-seed initialisation and tokenisation equal to the
reference's (when /root/reference is mounted -- build container only), config keys, synthetic task data formats."""
import pytest
import torch

from lpi_b200 import data as D, synthetic as S
from lpi_b200.config import default_args
from oracle import reference_loader as RL

needs_ref = pytest.mark.skipif(not RL.reference_available(), reason="reference tree not mounted")


@needs_ref
def test_clip_and_prompt_modules_match_reference_names_and_init():
    from lpi_b200 import clip
    from lpi_b200.prompts import DecomposedPrompt

    ns = RL.load_reference()
    small = (64, 32, 2, 128, 16, 77, 1000, 64, 1, 2)         # a tiny CLIP keeps this test in seconds
    torch.manual_seed(3)
    a = clip.CLIP(*small)
    torch.manual_seed(3)
    b = ns.clip_model.CLIP(*small)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    torch.manual_seed(4)
    p = DecomposedPrompt(9, 16, 768, 512)
    torch.manual_seed(4)
    q = ns.prompts.DecomposedPrompt(9, 16, 768, 512)
    assert [n for n, _ in p.named_parameters()] == [n for n, _ in q.named_parameters()]
    assert all(torch.equal(x, y) for x, y in zip(p.state_dict().values(), q.state_dict().values()))


@needs_ref
def test_tokenizer_matches_reference():
    from lpi_b200 import tokenizer as T

    caps = S.make_captions(32, 11) + ["A man's dog isn't here, it's 42 years-old!! (really?)", "hello   world\n\tx", "café naïve"]
    mine = T.tokenize(["X " * 16 + c + "." for c in caps])
    assert torch.equal(mine, RL.reference_tokenize(caps))
    with pytest.raises(RuntimeError):
        T.tokenize("word " * 100)


@needs_ref
def test_config_covers_reference_keys():
    ref = RL.reference_args()
    mine = default_args()
    missing = [k for k in ref if k not in mine and k not in ("image_root", "annotation_train_root", "annotation_val_root")]
    assert not missing
    for k in ("epochs", "lrate", "weight_decay", "batch_size", "prompt_length", "total_sessions", "NCTX", "prompt_type"):
        assert mine[k] == ref[k], k


def test_synthetic_task_loaders_have_reference_item_formats():
    loaders = D.make_task_loaders(2, n_train=8, n_eval_images=3, caps_per_image=2, batch_size=4, eval_batch_size=2, res=32)
    images, captions, zero, task = next(iter(loaders[1][0]))
    assert images.shape == (4, 3, 32, 32) and len(captions) == 4 and isinstance(captions[0], str) and int(task[0]) == 1
    ds = loaders[1][1].dataset
    assert len(ds.image) == 6 and len(ds.text) == 12 and ds.img2txt[4] == [8, 9] and ds.txt2img[9] == 4
    assert ds.text_cat[:4] == [0, 0, 0, 0] and ds.text_cat[-1] == 1
    img, idx, cat = ds[5]
    assert idx == 5 and cat == 1
