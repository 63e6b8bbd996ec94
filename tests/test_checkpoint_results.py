"""CPU checks of the 'next' rows f3 of SURVEY.md section 8: per-task checkpoint / resume of the learner state and the results
post-processing (the counterpart of the reference's res_handle/reshandle.py, compared with the reference's own output when the
reference tree is mounted)."""
import contextlib
import io
import json
import re

import pytest
import torch

from lpi_b200 import reshandle as RH
from lpi_b200.config import default_args
from oracle import reference_loader as RL

needs_ref = pytest.mark.skipif(not RL.reference_available(), reason="reference tree not mounted")


def _results(n=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    res = {}
    for s in range(n):
        res[s] = {"mscoco": {side: {t: [round(float(x), 3) for x in (torch.rand(3, generator=g) * 100).sort().values]
                                    for t in range(s + 1)} for side in ("i2t", "t2i")}}
    return res


def test_summarize_known_answer():
    res = {0: {"mscoco": {"i2t": {0: [50.0, 70.0, 80.0]}}},
           1: {"mscoco": {"i2t": {0: [40.0, 75.0, 78.0], 1: [10.0, 20.0, 30.0]}}},
           2: {"mscoco": {"i2t": {0: [45.0, 60.0, 79.0], 1: [12.0, 18.0, 33.0], 2: [90.0, 95.0, 99.0]}}}}
    out = RH.summarize(res)
    assert out["per_task"][0]["forgetting"] == [45.0 - 50.0, 60.0 - 75.0, 79.0 - 80.0]
    assert out["per_task"][1]["forgetting"] == [2.0, -2.0, 3.0]
    assert out["per_task"][2]["forgetting"] == [0.0, 0.0, 0.0] and out["per_task"][2]["sessions"] == 1
    assert out["forgetting"] == [(-5.0 + 2.0) / 2, (-15.0 - 2.0) / 2, (-1.0 + 3.0) / 2]
    assert out["avg_recall"][0] == pytest.approx((45.0 + 11.0 + 90.0) / 3)
    assert out["final"][1] == [12.0, 18.0, 33.0]
    w = RH.summarize(res, task_sizes=[1, 1, 2])["weighted_recall"]
    assert w[0] == pytest.approx((45.0 + 11.0 + 2 * 90.0) / 4)
    # JSON round trip turns the integer keys into strings, like the reference's ./res/*.json
    assert RH.summarize(json.loads(json.dumps(res))) == out


@needs_ref
@pytest.mark.parametrize("side", ["i2t", "t2i"])
def test_summarize_matches_reference_reshandle(tmp_path, side):
    import types

    # the reference file is a script (it runs get_res on a hard-coded ./res path at import time): execute only its definitions
    src = open(RL.REFERENCE_ROOT + "/res_handle/reshandle.py").read()
    cut = min(i for i in (src.find("\nfilename ="), src.find("\nget_res("), len(src)) if i > 0)
    ref = types.SimpleNamespace()
    ns = {}
    exec(compile(src[:cut], "reference_reshandle", "exec"), ns)
    ref.get_res = ns["get_res"]
    n = 5
    res = _results(n, seed=3)
    path = tmp_path / "res.json"
    path.write_text(json.dumps(res))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref.get_res(str(path), task_name="mscoco", task_type=side, n=n)
    text = buf.getvalue()
    nums = lambda line: [float(x) for x in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", line)]
    avg_line = next(l for l in text.splitlines() if l.startswith("org average precision"))
    forget_line = next(l for l in text.splitlines() if l.startswith("average forget"))
    out = RH.summarize(res, "mscoco", side)
    a = nums(avg_line.split("precision:")[1])
    assert a == pytest.approx(out["avg_recall"], rel=1e-12)
    f = nums(forget_line.split("forget:", 1)[1])
    assert f[:3] == pytest.approx(out["forgetting"], rel=1e-12, abs=1e-12) and f[3] == pytest.approx(out["avg_forgetting"], abs=1e-12)


def test_checkpoint_roundtrip_restores_prompts_keys_and_results(tmp_path):
    from lpi_b200.sprompt import SPrompts

    torch.manual_seed(0)
    a = SPrompts(default_args(device=[torch.device("cpu")], total_sessions=3))
    a._network.update_fc(0)
    a._network.update_fc(0)
    a.cur_id = 1
    a.all_keys = [torch.randn(5, 512) for _ in range(2)]
    a.textual_all_keys = [torch.randn(5, 512) for _ in range(2)]
    res = _results(2)
    path = a.save_checkpoint(str(tmp_path / "task_1.pt"), res)
    st = torch.load(path, weights_only=False)
    assert sum(v.numel() for k, v in st["trainable"].items() if k.startswith("prompts.")) == 3 * 5284      # 5 284 scalars per task
    assert not any(".clip_model." in k or k.startswith("image_encoder") for k in st["trainable"])          # CLIP itself is not pickled
    torch.manual_seed(1)
    b = SPrompts(default_args(device=[torch.device("cpu")], total_sessions=3))
    assert not torch.equal(b._network.prompts[0].dim_1_share, a._network.prompts[0].dim_1_share)
    got = b.load_checkpoint(path)
    assert got == res and b.cur_id == 1 and b._network.numtask == a._network.numtask == 2
    sa, sb = a._network.state_dict(), b._network.state_dict()
    assert all(torch.equal(sa[k], sb[k]) for k in st["trainable"])
    assert all(torch.equal(x, y) for x, y in zip(a.all_keys + a.textual_all_keys, b.all_keys + b.textual_all_keys))
    (tmp_path / "junk.pt").write_bytes(b"")
    torch.save({"format": "something else"}, str(tmp_path / "other.pt"))
    with pytest.raises(RuntimeError):
        b.load_checkpoint(str(tmp_path / "other.pt"))
