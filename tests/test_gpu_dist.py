"""-m gpu multi-GPU parity (needs >= 2 GPUs on the box, otherwise skipped): NCCL gallery-sharded search equals the single-GPU
result bit for bit; the data-parallel training step (sharded batch, all-gathered features, all-reduced 5 284-float gradient)
reproduces the single-process global-batch losses and gradients."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _scorer_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from lpi_b200 import ops, retrieval as R, synthetic as S

        ng, nq = 200_000, 3000
        lo, hi = R.shard_bounds(ng, world, rank, align=256)
        shard, q, gt = S.make_gallery_shard(ng, lo, hi, nq, 512, device=f"cuda:{rank}")
        sc, ix = R.search_topk(q, shard, 10, "bf16", lo, dist.group.WORLD)
        # the same exchange through the C ABI (lpi_comm_*): one all-gather of the packed per-chunk lists + the fused merge / recall kernel
        from lpi_b200.comm import LpiComm
        comm = LpiComm.from_torch_group(dist.group.WORLD)
        nch = ops.sim_topk_chunks(nq, hi - lo)
        buf, sv, iv = R.exchange_buffer(nch, nq, 10, f"cuda:{rank}")
        ops.sim_topk(q, shard, 10, lo, nch, merge=False, out=(sv, iv))
        gptr, gidx = R.gt_csr([[int(g)] for g in gt.tolist()])
        task = torch.zeros(nq, dtype=torch.int32, device=f"cuda:{rank}")
        sc2, ix2, counts = R.merge_recall(buf, gptr.to(sc.device), gidx.to(sc.device), task, 1, comm)
        gsum = comm.all_reduce_sum_(torch.full((4,), float(rank + 1), device=sc.device))
        comm_ok = bool(torch.equal(ix2, ix) and torch.equal(sc2, sc) and float(gsum[0]) == world * (world + 1) / 2 and int(counts[0, 3]) == nq)
        torch.cuda.synchronize()
        comm.close()
        if rank == 0:
            full, q2, _ = S.make_gallery_shard(ng, 0, ng, nq, 512, device="cuda:0")
            assert torch.equal(q, q2) and torch.equal(full[lo:hi], shard)          # sharding-independent workload
            s1, i1 = ops.sim_topk(q2, full, 10)
            out.put(("scorer", bool(torch.equal(ix, i1) and torch.equal(sc, s1) and comm_ok)))
    finally:
        dist.destroy_process_group()


def _train_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from lpi_b200 import lpi_step, ops, synthetic as S
        from lpi_b200.engine import TextEngine, VisionEngine

        sd = S.make_clip_state_dict(0)
        vision, text = VisionEngine(sd, dev), TextEngine(sd, dev)
        B = 8
        images, tokens = S.make_images(B, 3).to(dev), S.make_tokens(B, 3).to(dev)
        fac = {k: v.to(dev) for k, v in S.make_prompt_factors(2).items()}
        prev = [lpi_step.reconstruct({k: v.to(dev) for k, v in S.make_prompt_factors(0).items()})]
        sim = np.loadtxt(os.path.join(os.path.dirname(os.path.abspath(ops.__file__)), "MID", "task_sim_matrix.txt"))
        tgt = torch.tensor((sim[:2, :2] > 0.4).astype(np.int32), device=dev)
        b = B // world
        r = lpi_step.train_step(vision, text, fac, images[rank * b:(rank + 1) * b], tokens[rank * b:(rank + 1) * b], 1 / 0.07, prev, tgt,
                                group=dist.group.WORLD)
        if rank == 0:
            ref = lpi_step.train_step(vision, text, fac, images, tokens, 1 / 0.07, prev, tgt)
            ok = True
            for k in ref["losses"]:
                ok &= abs(float(r["losses"][k]) - float(ref["losses"][k])) < 1e-4 * max(1.0, abs(float(ref["losses"][k])))
            worst = 0.0
            for k in lpi_step.FACTOR_NAMES:
                worst = max(worst, float((r["grads"][k] - ref["grads"][k]).norm() / ref["grads"][k].norm()))
            out.put(("train", bool(ok and worst < 2e-3), worst))
    finally:
        dist.destroy_process_group()


def _learner_worker(rank, world, port, out, tmp):
    """The whole learner data-parallel over 2 ranks: each rank trains on its shard of the task's train set; the K-Means task keys come
    from the gathered features (identical on every rank), results agree, and rank 0 alone writes the checkpoints / result file."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from torch.utils.data import DataLoader, Subset

        from lpi_b200 import data as D, synthetic as S
        from lpi_b200.config import default_args
        from lpi_b200.sprompt import SPrompts

        os.chdir(tmp)
        n_tasks = 2
        args = default_args(clip_state_dict=S.make_clip_state_dict(0), device=[dev], epochs=1, batch_size=4, n_tasks=n_tasks,
                            group=dist.group.WORLD, checkpoint_dir=os.path.join(tmp, "ckpt"))
        learner = SPrompts(args)
        with torch.no_grad():
            for t in range(n_tasks):
                for k, v in S.make_prompt_factors(t).items():
                    getattr(learner._network.prompts[t], k).copy_(v)
        loaders = []
        for tr, te in D.make_task_loaders(n_tasks, 16, 6, 2, 8, 8):
            shard = Subset(tr.dataset, list(range(rank, len(tr.dataset), world)))          # this rank's half of the task's train set
            loaders.append((DataLoader(shard, batch_size=4, shuffle=False), te))
        res = learner.incremental_train(loaders)
        keys = torch.stack(learner.all_keys + learner.textual_all_keys).float()
        fac = torch.cat([getattr(learner._network.prompts[t], k).detach().reshape(-1) for t in range(n_tasks) for k in S.FACTOR_NAMES])
        both_k = [torch.empty_like(keys) for _ in range(world)]
        both_f = [torch.empty_like(fac) for _ in range(world)]
        dist.all_gather(both_k, keys)
        dist.all_gather(both_f, fac)
        all_res = [None] * world
        dist.all_gather_object(all_res, res)
        if rank == 0:
            import glob
            ok = bool(torch.equal(both_k[0], both_k[1]) and torch.equal(both_f[0], both_f[1]) and all_res[0] == all_res[1])
            files = sorted(os.listdir(os.path.join(tmp, "ckpt")))
            ok &= files == ["task_0.pt", "task_1.pt"] and len(glob.glob(os.path.join(tmp, "res", "*.json"))) == 1
            out.put(("learner", ok, files))
    finally:
        dist.destroy_process_group()


def test_two_gpu_learner_keys_results_and_files(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 30600 + os.getpid() % 2000
    procs = [ctx.Process(target=_learner_worker, args=(r, 2, port, out, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=900)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[1], res


@pytest.mark.parametrize("worker", [_scorer_worker, _train_worker])
def test_two_gpu_parity(worker):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[1], res
