"""TEST INFRASTRUCTURE ONLY -- loads the *real* reference (Kelvin-ywc/LPI, `retrieval/`)
on CPU so that the restatement in `oracle/lpi_oracle.py` can be validated against it and
golden vectors can be generated (tests/golden/make_golden.py).

Nothing under `lpi_b200/` may import this module.  `/root/reference` only exists in the
build container, never on the GPU box: callers must check `reference_available()`.
`__graft_entry__.build()` stages an UNMODIFIED copy of the reference's `retrieval/` package under
`baseline/_ref/retrieval` (git-ignored, so never part of the history, but shipped to the GPU box):
that copy is what `bench.py`'s reference legs time on the box's host cores and -- through stock
torch -- on the B200 itself (`cuda=True`, the "GPU eager" kernel to beat).

The reference has no FFI/test harness of its own; to import it on a CPU-only box we need
(SURVEY.md section 8(c)):
  * stub modules for `ftfy` (simple_tokenizer.py:6) and `timm` (models/vit.py:26-29),
  * `.cuda()` -> identity, `torch.cuda.current_device/device_count` -> 0/1
    (prompt_learner.py:122,132,146-148; sprompt.py:301,442-443,457-458),
  * `load_clip_to_cpu` replaced by a random-init ViT-B/16 CLIP in fp32
    (prompt_learner.py:10-40 would download weights),
  * cwd containing `MID/task_sim_matrix.txt` (slinet.py:171) and writable `logs/`, `res/`.
Nothing in /root/reference is modified.
"""
from __future__ import annotations

import contextlib
import json
import os
import sys
import tempfile
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(_REPO, "baseline", "_ref", "retrieval")


def _find_root() -> str:
    for c in (os.environ.get("LPI_REFERENCE_ROOT"), "/root/reference/retrieval", STAGED_ROOT):
        if c and os.path.isfile(os.path.join(c, "models", "slinet.py")):
            return c
    return "/root/reference/retrieval"


REFERENCE_ROOT = _find_root()

_loaded = {}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "slinet.py"))


def _install_stubs():
    if "ftfy" not in sys.modules:
        try:
            import ftfy  # noqa: F401
        except ImportError:
            m = types.ModuleType("ftfy")
            m.fix_text = lambda s: s  # exact for ASCII captions
            sys.modules["ftfy"] = m
    try:
        import timm  # noqa: F401
        return
    except ImportError:
        pass

    def _mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    import torch.nn as nn

    timm = _mod("timm")
    data = _mod("timm.data")
    for k in ("IMAGENET_DEFAULT_MEAN", "IMAGENET_DEFAULT_STD", "IMAGENET_INCEPTION_MEAN", "IMAGENET_INCEPTION_STD"):
        setattr(data, k, (0.5, 0.5, 0.5))
    models = _mod("timm.models")
    helpers = _mod("timm.models.helpers")
    for k in ("build_model_with_cfg", "resolve_pretrained_cfg", "named_apply", "adapt_input_conv", "checkpoint_seq"):
        setattr(helpers, k, lambda *a, **kw: None)
    layers = _mod("timm.models.layers")
    for k in ("PatchEmbed", "Mlp", "DropPath"):
        setattr(layers, k, type(k, (nn.Module,), {}))
    layers.trunc_normal_ = lambda *a, **kw: None
    layers.lecun_normal_ = lambda *a, **kw: None
    registry = _mod("timm.models.registry")
    registry.register_model = lambda f: f
    timm.data, timm.models = data, models
    models.helpers, models.layers, models.registry = helpers, layers, registry


def scratch_cwd() -> str:
    """A writable directory laid out the way the reference expects its cwd."""
    d = _loaded.get("cwd")
    if d is None:
        d = tempfile.mkdtemp(prefix="lpi_ref_cwd_")
        os.makedirs(os.path.join(d, "logs"), exist_ok=True)
        os.makedirs(os.path.join(d, "res"), exist_ok=True)
        os.symlink(os.path.join(REFERENCE_ROOT, "MID"), os.path.join(d, "MID"))
        _loaded["cwd"] = d
    return d


@contextlib.contextmanager
def in_reference_cwd():
    old = os.getcwd()
    os.chdir(scratch_cwd())
    try:
        yield
    finally:
        os.chdir(old)


def load_reference(cuda: bool = False):
    """Import the reference modules (once) and return a namespace of the ones on the hot path.
    cuda=False (default): CPU oracle -- `.cuda()` and the device queries are patched out.  cuda=True (bench.py's GPU-eager leg on the
    B200 box): nothing is patched, the reference runs through stock torch on the current CUDA device.  One mode per process."""
    if "ns" in _loaded:
        if _loaded.get("cuda") != cuda:
            raise RuntimeError("the reference was already loaded in the other device mode in this process")
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    import torch

    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not cuda:
        # CPU-only patches (prompt_learner.py:122,132,146-148)
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.cuda.current_device = lambda: 0
        torch.cuda.device_count = lambda: 1
    _loaded["cuda"] = cuda
    with in_reference_cwd():
        import models.slinet as slinet
        import models.clip.model as clip_model
        import models.clip.clip as clip_clip
        import models.clip.prompt_learner as prompt_learner
        import models.prompts.prompts as prompts
        import loss.loss as loss
        import methods.sprompt as sprompt

    def _random_clip(args):
        # ViT-B/16 hyper-parameters as build_model would derive them (model.py:418-441);
        # fp32 weights, mirroring clip.load on CPU (clip.py:128-129).
        return clip_model.CLIP(512, 224, 12, 768, 16, 77, 49408, 512, 8, 12).eval()

    slinet.load_clip_to_cpu = _random_clip
    ns = types.SimpleNamespace(
        slinet=slinet, clip_model=clip_model, clip=clip_clip, prompt_learner=prompt_learner,
        prompts=prompts, loss=loss, sprompt=sprompt, root=REFERENCE_ROOT,
    )
    _loaded["ns"] = ns
    return ns


def reference_args() -> dict:
    with open(os.path.join(REFERENCE_ROOT, "configs", "lpi", "coco_lpi.json")) as f:
        args = json.load(f)
    import torch

    args["device"] = [torch.device("cpu")]
    return args


def build_reference_slinet(state_dict=None, prompt_state=None, numtask: int = 1):
    """SliNet(args) from the real reference with deterministic weights loaded.

    `state_dict`: CLIP tensors under the reference's own `clip_model.*` key names.
    `prompt_state`: {task: {factor_name: tensor}} for `prompts.{task}.*`.
    """
    import torch

    ns = load_reference()
    with in_reference_cwd():
        net = ns.slinet.SliNet(reference_args())
    if state_dict is not None:
        missing, unexpected = net.clip_model.load_state_dict(state_dict, strict=True)
        assert not missing and not unexpected
    if prompt_state is not None:
        with torch.no_grad():
            for t, factors in prompt_state.items():
                for k, v in factors.items():
                    getattr(net.prompts[t], k).copy_(v)
    for _ in range(numtask):
        net.update_fc(0)
    return net


def reference_tokenize(captions, n_ctx: int = 16):
    """Token ids exactly as PromptLearner.forward builds them (prompt_learner.py:128-132)."""
    import torch

    ns = load_reference()
    prefix = " ".join(["X"] * n_ctx)
    return torch.cat([ns.clip.tokenize(prefix + " " + c + ".") for c in captions])


def bpe_vocab_path() -> str:
    return os.path.join(REFERENCE_ROOT, "models", "clip", "bpe_simple_vocab_16e6.txt.gz")
