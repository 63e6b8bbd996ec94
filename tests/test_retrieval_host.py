"""Host-side logic of the gallery-sharded Recall@K path, on CPU: shard bounds, CSR ground truth, the
result-dict format, and the N > 1 exchange step over gloo (world_size 2).  Compute in these tests is
done by the oracle (tests may use it); the CUDA kernels are covered by the -m gpu tests."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lpi_b200 import retrieval as R
from lpi_b200 import synthetic as S
from oracle import lpi_oracle as O


@pytest.mark.parametrize("n,w,align", [(10, 3, 1), (5_000_000, 8, 256), (1000, 7, 256), (5, 8, 1), (0, 2, 1)])
def test_shard_bounds_partition(n, w, align):
    prev = 0
    sizes = []
    for r in range(w):
        lo, hi = R.shard_bounds(n, w, r, align)
        assert lo == prev and hi >= lo
        if hi < n:
            assert hi % align == 0
        prev = hi
        sizes.append(hi - lo)
    assert prev == n
    assert max(sizes) - min(sizes) <= 2 * align      # last shard carries the ragged tail


def test_gt_csr_and_result_dict():
    ptr, idx = R.gt_csr([[0, 1], [], [7]])
    assert ptr.tolist() == [0, 2, 2, 3] and idx.tolist() == [0, 1, 7]
    ci = torch.tensor([[1, 2, 3, 3], [0, 1, 1, 4]])
    d = R.recall_dict(ci, ci)
    assert d["mscoco"]["i2t"][0] == [100.0 * 1 / 3, 100.0 * 2 / 3, 100.0]
    assert d["mscoco"]["t2i"][1] == [0.0, 25.0, 25.0]
    with pytest.raises(ZeroDivisionError):          # the reference divides by len(ranks_) too (sprompt.py:581)
        R.recall_dict(torch.tensor([[0, 0, 0, 0]]), ci)


def _cpu_merge(all_s, all_i, k):
    w, nq, _ = all_s.shape
    s = all_s.permute(1, 0, 2).reshape(nq, -1).numpy()
    i = all_i.permute(1, 0, 2).reshape(nq, -1).numpy()
    order = np.lexsort((i, -s), axis=1)[:, :k]
    return np.take_along_axis(s, order, 1), np.take_along_axis(i, order, 1)


def _worker(rank, world, port, q, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img, txt, *_ = S.make_retrieval_set(60, 5, 64, 3, seed=9)
        lo, hi = R.shard_bounds(txt.shape[0], world, rank)
        s_local = (img @ txt[lo:hi].t()).numpy()
        k = 10
        sc, ix = O.topk_lowest_index(s_local, k)
        all_s, all_i = R.gather_candidates(torch.from_numpy(sc.astype(np.float32)),
                                           torch.from_numpy((ix + lo).astype(np.int32)), dist.group.WORLD)
        ms, mi = _cpu_merge(all_s, all_i, k)
        if rank == 0:
            out.put((ms, mi))
    finally:
        dist.destroy_process_group()


def test_gallery_sharded_exchange_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, None, out)) for r in range(2)]
    for p in procs:
        p.start()
    ms, mi = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    img, txt, *_ = S.make_retrieval_set(60, 5, 64, 3, seed=9)
    ws, wi = O.topk_lowest_index((img @ txt.t()).numpy(), 10)
    assert np.array_equal(mi, wi)          # sharded + gathered + merged == global top-k, bit-exact
    assert np.array_equal(ms, ws.astype(np.float32))
