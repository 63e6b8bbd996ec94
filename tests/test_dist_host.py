"""Host-side data-parallel plumbing of the learner on CPU over gloo (world_size 2): `gather_features` (the reference's DP entry point,
retrieval/methods/sprompt.py:38-82), the ragged row gather behind `SPrompts.clustering` in a data-parallel run, and the feature
exchange layout of `lpi_step.train_step` (one concatenated [b, 2E] buffer, rank-major rows)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lpi_b200 import sprompt as SP


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        img = torch.randn(3, 8, generator=g, requires_grad=True)
        txt = torch.randn(3, 8, generator=g, requires_grad=True)
        res = {}
        # --- default mode (gather_with_grad=False, local_loss=False): every slot holds that rank's values, the own slot keeps the graph
        ai, at = SP.gather_features(img, txt, False, False, rank, world)
        res["plain"] = (ai.detach().clone(), at.detach().clone())
        (ai.sum() * 2 + at.sum() * 3).backward()
        res["plain_grad"] = (img.grad.clone(), txt.grad.clone())
        img.grad = txt.grad = None
        # --- local_loss=True: the gathered copies are all detached
        ai, at = SP.gather_features(img, txt, True, False, rank, world)
        res["local_requires_grad"] = bool(ai.requires_grad)
        # --- gather_with_grad=True: gradient flows to the local slice from every rank's loss (sum over ranks of d/d slot)
        ai, at = SP.gather_features(img, txt, False, True, rank, world)
        ((rank + 1.0) * ai.sum()).backward()
        res["gwg_grad"] = img.grad.clone()
        # --- the exchange layout of lpi_step.train_step: concatenated buffer, rank-major
        both = torch.cat([img.detach(), txt.detach()], dim=1).contiguous()
        gathered = torch.empty(world * 3, 16)
        dist.all_gather_into_tensor(gathered, both)
        res["step_layout"] = gathered.clone()
        # --- ragged gather (clustering under DP): rank r contributes 2 + r rows
        rows = torch.full((2 + rank, 4), float(rank)) + torch.arange(2 + rank).view(-1, 1)
        res["ragged"] = SP.gather_rows(rows, dist.group.WORLD)
        res["main"] = SP.is_main_rank(dist.group.WORLD)
        def plain(v):                    # tensors travel as numpy: a child may exit before the parent maps a shared-memory tensor
            if torch.is_tensor(v):
                return v.detach().numpy()
            if isinstance(v, tuple):
                return tuple(plain(x) for x in v)
            return v
        out.put((rank, {k: plain(v) for k, v in res.items()}))
    finally:
        dist.destroy_process_group()


def test_gather_features_and_row_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    def back(v):
        if isinstance(v, tuple):
            return tuple(back(x) for x in v)
        return torch.from_numpy(v) if hasattr(v, "dtype") and hasattr(v, "shape") else v
    got = {r: {k: back(v) for k, v in res.items()} for r, res in (out.get(timeout=90) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    feats = []
    for r in range(2):
        g = torch.Generator().manual_seed(100 + r)
        feats.append((torch.randn(3, 8, generator=g), torch.randn(3, 8, generator=g)))
    want_i = torch.cat([f[0] for f in feats])
    want_t = torch.cat([f[1] for f in feats])
    for r in range(2):
        res = got[r]
        assert torch.equal(res["plain"][0], want_i) and torch.equal(res["plain"][1], want_t)          # rank-major, identical on all ranks
        assert torch.equal(res["plain_grad"][0], torch.full((3, 8), 2.0)) and torch.equal(res["plain_grad"][1], torch.full((3, 8), 3.0))
        assert res["local_requires_grad"] is False
        assert torch.equal(res["gwg_grad"], torch.full((3, 8), 3.0))          # (0 + 1) + (1 + 1): both ranks' losses reach the local slice
        assert torch.equal(res["step_layout"][:, :8], want_i) and torch.equal(res["step_layout"][:, 8:], want_t)
        want_rows = torch.cat([torch.full((2, 4), 0.0) + torch.arange(2).view(-1, 1), torch.full((3, 4), 1.0) + torch.arange(3).view(-1, 1)])
        assert torch.equal(res["ragged"], want_rows)
        assert res["main"] is (r == 0)


def test_single_process_helpers():
    assert SP.is_main_rank(None) is True
    a, b = torch.ones(2, 4), torch.zeros(2, 4)
    ga, gb = SP.gather_features(a, b, world_size=1)
    assert ga is a and gb is b
