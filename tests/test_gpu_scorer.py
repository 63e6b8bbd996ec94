"""-m gpu parity tests of the retrieval scorer: CUDA path (through the C ABI) vs the oracle on the same
seeded inputs.  Bar: top-k indices and Recall@K bit-exact (ties -> lowest index)."""
import os

import numpy as np
import pytest
import torch

from lpi_b200 import ops, retrieval as R, synthetic as S
from oracle import lpi_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle_topk(q, g, k):
    s = (q.float().cpu() @ g.float().cpu().t()).numpy()
    return O.topk_lowest_index(s, k) + (s,)


@pytest.mark.parametrize("nq,ng,dim,chunks", [(1, 1, 64, 1), (128, 256, 64, 1), (100, 1000, 512, 1), (300, 5000, 512, 3),
                                              (1000, 20000, 512, 0), (257, 70001, 512, 0), (37, 9, 512, 1)])
def test_sim_topk_matches_oracle(nq, ng, dim, chunks):
    gen = torch.Generator().manual_seed(nq * 7 + ng)
    q = torch.randn(nq, dim, generator=gen).bfloat16()
    g = torch.randn(ng, dim, generator=gen).bfloat16()
    if ng > 5:
        g[5] = g[3]                         # exact ties: the lower index must win
    k = 10
    v, i = ops.sim_topk(q.cuda(), g.cuda(), k, 0, chunks)
    wv, wi, s = _oracle_topk(q, g, k)
    got_i = i.cpu().numpy()
    kk = min(k, ng)
    assert (got_i[:, kk:] == 0x7FFFFFFF).all()                 # ragged: missing entries are (-inf, INT_MAX)
    bad = 0
    for r in range(nq):
        verdict = O.audit_topk(got_i[r, :kk], wi[r, :kk], q[r], g)
        assert verdict != "bad", (r, got_i[r], wi[r])
        bad += verdict == "near"
    assert bad <= max(1, nq // 200)                             # accumulation-order near-ties are rare
    assert np.allclose(v.cpu().numpy()[:, :kk], wv[:, :kk], rtol=0, atol=2e-5 * max(1.0, np.abs(wv[:, :kk]).max()))


def test_gallery_offset_and_shard_merge_equals_unsharded():
    gen = torch.Generator().manual_seed(3)
    q = torch.randn(500, 512, generator=gen).bfloat16().cuda()
    g = torch.randn(30000, 512, generator=gen).bfloat16().cuda()
    v0, i0 = ops.sim_topk(q, g, 10)
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            lo, hi = R.shard_bounds(g.shape[0], world, r, align=256)
            parts.append(ops.sim_topk(q, g[lo:hi], 10, lo))
        v, i = ops.topk_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
        assert torch.equal(i, i0) and torch.equal(v, v0)        # score of a pair does not depend on the sharding


@pytest.mark.parametrize("order", ["random", "ascending", "descending"])
def test_threshold_seeding_is_exact(order):
    """The seeded two-pass form (thresholds from a pre-pass over the first rows) must return exactly the unseeded lists, including
    exact ties at the threshold and galleries ordered adversarially (best rows last: the seed helps nothing; best rows first: the
    seed equals the final k-th score, so every later candidate sits AT the threshold)."""
    gen = torch.Generator().manual_seed(17)
    q = torch.randn(300, 512, generator=gen).bfloat16().cuda()
    g = torch.randn(40000, 512, generator=gen).bfloat16()
    g[1000:1040] = g[40:80]                       # duplicates of rows inside the seed sample: ties exactly at / above the seed
    g[39000] = g[7]
    if order != "random":
        key = (q[0].float().cpu() @ g.float().t())
        g = g[torch.argsort(key, descending=(order == "descending"))]
    g = g.cuda().contiguous()
    v0, i0 = ops.sim_topk(q, g, 10, 5, seed_rows=0)
    for seed_rows, chunks in ((512, 0), (4096, 3), (10, 1), (40000, 0)):
        v, i = ops.sim_topk(q, g, 10, 5, chunks, seed_rows=seed_rows)
        assert torch.equal(i, i0) and torch.equal(v, v0), (order, seed_rows, chunks)
    # k larger than the seed sample: the seed list is short, its k-th entry is -inf, nothing is filtered
    v, i = ops.sim_topk(q, g, 16, 5, seed_rows=12)
    v1, i1 = ops.sim_topk(q, g, 16, 5, seed_rows=0)
    assert torch.equal(i, i1) and torch.equal(v, v1)


def test_chained_thresholds_over_gallery_chunks_are_exact():
    """Streaming a gallery in chunks with each chunk seeded by the best k-th score of the chunks before it (bench.py e2e leg) and merging the
    lists gives exactly the one-shot result; a chunk's list may legitimately be short."""
    gen = torch.Generator().manual_seed(23)
    q = torch.randn(200, 512, generator=gen).bfloat16().cuda()
    g = torch.randn(30000, 512, generator=gen).bfloat16()
    g[29000:29010] = g[10:20]                     # ties across chunks: the earlier (lower-index) copy must win
    g = g.cuda()
    v0, i0 = ops.sim_topk(q, g, 10, 0, seed_rows=0)
    ls, li, thr = [], [], None
    for c in range(5):
        a, b = c * 6000, (c + 1) * 6000
        s_, i_ = ops.sim_topk(q, g[a:b], 10, a, seed_rows=0, init_thr=thr)
        thr = s_[:, 9].contiguous() if thr is None else torch.maximum(thr, s_[:, 9])
        ls.append(s_); li.append(i_)
    assert bool((li[-1] == 0x7FFFFFFF).any())                   # later chunks keep fewer than k candidates
    v, i = ops.topk_merge(torch.stack(ls), torch.stack(li))
    assert torch.equal(i, i0) and torch.equal(v, v0)


@pytest.mark.parametrize("order", ["random", "ascending", "descending"])
def test_cooperative_thresholds_are_exact(order):
    """Chunks of one launch sharing their per-query thresholds through global memory (lpi_sim_topk_coop_bf16) must give exactly the merged
    lists of the independent chunks -- with exact ties across chunk boundaries and adversarial gallery orders -- while the per-chunk lists
    themselves may be short.  Repeated launches (timing-dependent sharing) agree with each other."""
    gen = torch.Generator().manual_seed(29)
    q = torch.randn(300, 512, generator=gen).bfloat16().cuda()
    g = torch.randn(60000, 512, generator=gen).bfloat16()
    g[45000:45040] = g[100:140]                   # ties between the first and the last chunk: the lower index must win
    g[20000] = g[59999]
    if order != "random":
        key = (q[0].float().cpu() @ g.float().t())
        g = g[torch.argsort(key, descending=(order == "descending"))]
    g = g.cuda().contiguous()
    v0, i0 = ops.sim_topk(q, g, 10, 7, 1, seed_rows=0, coop=False)
    for chunks, seed_rows in ((4, 0), (6, 2048), (3, 16384)):
        for _ in range(3):
            ps, pi = ops.sim_topk(q, g, 10, 7, chunks, merge=False, seed_rows=seed_rows, coop=True)
            v, i = ops.topk_merge(ps, pi)
            assert torch.equal(i, i0) and torch.equal(v, v0), (order, chunks, seed_rows)
    assert bool((pi == 0x7FFFFFFF).any())         # some chunk list is short: its candidates were ruled out by another chunk's threshold
    # init_thr + coop (a streamed gallery whose pieces are themselves chunked)
    thr = v0[:, 9].contiguous() - 0.25
    v, i = ops.sim_topk(q, g, 10, 7, 4, init_thr=thr, coop=True)
    assert torch.equal(i, i0) and torch.equal(v, v0)
    assert torch.equal(thr, v0[:, 9] - 0.25)      # the caller's tensor is not written


def test_fused_merge_recall_equals_separate_kernels():
    """lpi_topk_merge_recall on an exchange buffer laid out like an all-gather of [2, c, Q, k] per rank == topk_merge + recall_counts."""
    gen = torch.Generator().manual_seed(31)
    G, c, nq, k = 3, 2, 777, 10
    sc = torch.randn(G, c, nq, k, generator=gen).sort(dim=-1, descending=True)[0]
    sc[1, 0, :, 3] = sc[0, 1, :, 2]               # equal scores in different parts: the lower index must win
    ix = torch.randperm(G * c * nq * k, generator=gen).to(torch.int32).view(G, c, nq, k) % 5000
    packed = torch.empty(G, 2, c, nq, k, dtype=torch.int32)
    packed[:, 0] = sc.view(torch.int32)
    packed[:, 1] = ix
    gt = [[int(x) for x in torch.randint(0, 5000, (1 + r % 3,), generator=gen)] for r in range(nq)]
    for r in range(0, nq, 5):
        gt[r][0] = int(ix[r % G, r % c, r, r % k])               # plant hits at every position
    ptr, idx = R.gt_csr(gt)
    task = (torch.arange(nq) % 4).to(torch.int32)
    v, i, counts, rank = ops.topk_merge_recall(packed.cuda(), ptr.cuda(), idx.cuda(), task.cuda(), 4, want_rank=True)
    wv, wi = ops.topk_merge(sc.view(G * c, nq, k).cuda(), ix.view(G * c, nq, k).cuda())
    wc, wr = ops.recall_counts(wi, ptr.cuda(), idx.cuda(), task.cuda(), 4, want_rank=True)
    assert torch.equal(v, wv) and torch.equal(i, wi) and torch.equal(counts, wc) and torch.equal(rank, wr)
    assert int(counts[:, 3].sum()) == nq and int(counts[:, 2].sum()) > 0
    # single rank, through the helper that owns the buffer layout
    buf, sv, iv = R.exchange_buffer(c, nq, k, "cuda")
    sv.copy_(sc[0]); iv.copy_(ix[0])
    v1, i1, c1 = R.merge_recall(buf, ptr.cuda(), idx.cuda(), task.cuda(), 4)
    wv1, wi1 = ops.topk_merge(sc[0].cuda(), ix[0].cuda())
    assert torch.equal(v1, wv1) and torch.equal(i1, wi1) and torch.equal(c1, ops.recall_counts(wi1, ptr.cuda(), idx.cuda(), task.cuda(), 4))


def test_lpi_comm_single_rank_roundtrip():
    """lpi_comm_* (the C-ABI exchange steps) with one rank: NCCL is resolved at run time, all-gather / all-reduce are identities, and the
    sharded-search tail accepts the communicator in place of a torch.distributed group."""
    from lpi_b200.comm import LpiComm, unique_id

    uid = unique_id()
    assert len(uid) == 128
    comm = LpiComm(1, 0, uid)
    x = torch.arange(24, dtype=torch.int32, device="cuda").view(2, 3, 4)
    assert torch.equal(comm.all_gather(x), x.unsqueeze(0))
    g = torch.randn(5284, device="cuda")
    want = g.clone()
    assert torch.equal(comm.all_reduce_sum_(g), want)
    gen = torch.Generator().manual_seed(5)
    q = torch.randn(64, 512, generator=gen).bfloat16().cuda()
    gal = torch.randn(5000, 512, generator=gen).bfloat16().cuda()
    buf, sv, iv = R.exchange_buffer(2, 64, 10, "cuda")
    ops.sim_topk(q, gal, 10, 0, 2, merge=False, out=(sv, iv))
    ptr, idx = R.gt_csr([[int(i)] for i in range(64)])
    task = torch.zeros(64, dtype=torch.int32, device="cuda")
    a = R.merge_recall(buf, ptr.cuda(), idx.cuda(), task, 1, comm)
    b = R.merge_recall(buf, ptr.cuda(), idx.cuda(), task, 1, None)
    assert all(torch.equal(x_, y_) for x_, y_ in zip(a, b))
    comm.close()


def test_topk_rows_and_merge_ties():
    gen = torch.Generator().manual_seed(11)
    s = torch.randn(70, 4001, generator=gen)
    s[:, 100] = s[:, 7]
    s[:, 4000] = s[:, 0]
    v, i = ops.topk_rows(s.cuda(), 10)
    wv, wi = O.topk_lowest_index(s.numpy(), 10)
    assert np.array_equal(i.cpu().numpy(), wi) and np.array_equal(v.cpu().numpy(), wv)
    # wide fan-in merge (hierarchical above 25 parts)
    ps = torch.randn(40, 33, 10, generator=gen).sort(dim=-1, descending=True)[0]
    pi = torch.arange(40 * 33 * 10, dtype=torch.int32).view(40, 33, 10)
    mv, mi = ops.topk_merge(ps.cuda(), pi.cuda())
    flat_s = ps.permute(1, 0, 2).reshape(33, -1).numpy()
    flat_i = pi.permute(1, 0, 2).reshape(33, -1).numpy()
    order = np.lexsort((flat_i, -flat_s), axis=1)[:, :10]
    assert np.array_equal(mi.cpu().numpy(), np.take_along_axis(flat_i, order, 1))


@pytest.mark.parametrize("name", ["recall_flickr_seed2.pt", "recall_small_seed5.pt"])
def test_recall_matches_reference_golden(name):
    """Recall@K dict bit-exact against the fixture produced by the REAL reference's itm_eval."""
    g = torch.load(os.path.join(GOLDEN, name), weights_only=False)
    m = g["meta"]
    img, txt, img2txt, txt2img, cat_i, cat_t = S.make_retrieval_set(m["n_img"], m["caps_per_img"], 512, m["n_tasks"],
                                                                    seed=m["seed"], signal=m.get("signal", 0.15))
    got = R.itm_eval_features(img.cuda(), txt.cuda(), txt2img, img2txt, cat_i, cat_t, m["n_tasks"], precision="fp32")
    assert got == g["result"]
    s = (img @ txt.t()).numpy()
    got_dense = R.itm_eval(s, np.ascontiguousarray(s.T), txt2img, img2txt, cat_i, cat_t, m["n_tasks"])
    assert got_dense == g["result"]


def test_fp32_split_scores_match_fp32_dot():
    gen = torch.Generator().manual_seed(21)
    a = torch.randn(64, 512, generator=gen)
    b = torch.randn(300, 512, generator=gen)
    a, b = a / a.norm(dim=-1, keepdim=True), b / b.norm(dim=-1, keepdim=True)
    v, i = R.search_topk(a.cuda(), b.cuda(), 10, precision="fp32")
    s64 = (a.double() @ b.double().t())
    wv, wi = torch.sort(s64, dim=1, descending=True, stable=True)
    assert torch.equal(i.cpu().long(), wi[:, :10])
    # tensor-core fp32 accumulation is not round-to-nearest: scores sit within ~1e-6 of the fp64 dot product
    assert (v.cpu().double() - wv[:, :10]).abs().max() < 3e-6


def test_l2_normalize_and_recall_counts_edge_cases():
    x = torch.randn(33, 512).cuda()
    y, n = ops.l2_normalize(x, want_norm=True)
    assert torch.allclose(y, x / x.norm(dim=-1, keepdim=True), atol=1e-6)
    top = torch.tensor([[3, 1, 2], [0, 0, 0], [9, 9, 9]], dtype=torch.int32).cuda()
    ptr, idx = R.gt_csr([[2], [], [5, 9]])
    counts, rank = ops.recall_counts(top, ptr.cuda(), idx.cuda(), torch.tensor([0, 1, 1], dtype=torch.int32).cuda(), 2,
                                     want_rank=True)
    assert rank.tolist() == [2, 3, 0]
    assert counts.tolist() == [[0, 1, 1, 1], [1, 1, 1, 2]]


def test_large_gallery_property_full_size_subsample():
    """BASELINE-size property check: on a 1M-row gallery, sampled queries' top-10 equal the oracle's, and
    sharding the gallery leaves every list unchanged."""
    ng, nq = 1_000_000, 4096
    shard, q, gt = S.make_gallery_shard(ng, 0, ng, nq, 512, device="cuda")
    v, i = ops.sim_topk(q, shard, 10)
    lo, hi = R.shard_bounds(ng, 2, 1, align=256)
    a = ops.sim_topk(q, shard[:lo], 10, 0)
    b = ops.sim_topk(q, shard[lo:], 10, lo)
    v2, i2 = ops.topk_merge(torch.stack([a[0], b[0]]), torch.stack([a[1], b[1]]))
    assert torch.equal(i, i2) and torch.equal(v, v2)
    sub = torch.arange(0, nq, 128)
    g_host = shard.cpu().float()
    s = (q[sub].cpu().float() @ g_host.t()).numpy()
    _, wi = O.topk_lowest_index(s, 10)
    got = i[sub].cpu().numpy()
    for r in range(len(sub)):
        assert O.audit_topk(got[r], wi[r], q[sub[r]].cpu(), g_host) != "bad"
    # the planted ground truth is found far more often than chance
    hit10 = (i.cpu().long() == gt[:, None]).any(1).float().mean().item()
    assert hit10 > 0.2


def test_full_size_sweep_properties():
    """BASELINE.json configs[4] at FULL size (25 000 queries x 5 000 000 gallery rows, d = 512): the oracle cannot materialise a
    500 GB score matrix, so parity is carried by size-independent properties --
      (a) every list is sorted by (score desc, index asc) and holds valid, distinct gallery indices;
      (b) the reported scores equal an independent fp32 dot product of the returned rows;
      (c) sharding is idempotent: 8 contiguous shards + k-way merge reproduce the unsharded lists bit for bit;
      (d) on a random sample of queries the lists equal the oracle's lowest-index top-k of the dense fp32 scores
          (accumulation-order near-ties audited in fp64, as in the small-size tests);
      (e) Recall@K counted from the lists equals the count-rank definition rank = #{j: s_j > s_gt} on that sample."""
    nq, ng, k = 25_000, 5_000_000, 10
    shard, q, gt = S.make_gallery_shard(ng, 0, ng, nq, 512, device="cuda")
    v, i = ops.sim_topk(q, shard, k)
    torch.cuda.synchronize()
    # (a)
    assert bool((i >= 0).all()) and bool((i < ng).all())
    dv = v[:, 1:] - v[:, :-1]
    assert bool((dv <= 0).all())
    tie = dv == 0
    assert bool((i[:, 1:][tie] > i[:, :-1][tie]).all())
    assert bool((torch.sort(i, dim=1).values[:, 1:] != torch.sort(i, dim=1).values[:, :-1]).all())
    # (b)
    rows = shard[i.long().reshape(-1)].float().view(nq, k, 512)
    dots = torch.einsum("qkd,qd->qk", rows, q.float())
    assert float((dots - v).abs().max()) < 2e-5
    del rows, dots
    # (c)
    parts = []
    for r in range(8):
        lo, hi = R.shard_bounds(ng, 8, r, align=256)
        parts.append(ops.sim_topk(q, shard[lo:hi], k, lo))
    v8, i8 = ops.topk_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(i8, i) and torch.equal(v8, v)
    del parts
    # (d) + (e) on 48 sampled queries against the dense fp32 scores
    sel = torch.randperm(nq, generator=torch.Generator().manual_seed(9))[:48]
    qs = q[sel.cuda()].float()
    dense = torch.empty(len(sel), ng, device="cuda")
    for c0 in range(0, ng, 500_000):
        dense[:, c0:c0 + 500_000] = qs @ shard[c0:c0 + 500_000].float().t()
    dense_np = dense.cpu().numpy()
    _, want = O.topk_lowest_index(dense_np, k)
    got = i[sel.cuda()].cpu().numpy()
    near = 0
    g_cpu = None
    for r in range(len(sel)):
        if (got[r] == want[r]).all():
            continue
        if g_cpu is None:
            g_cpu = shard.cpu()
        verdict = O.audit_topk(got[r], want[r], q[sel[r]].cpu(), g_cpu)
        assert verdict != "bad", (r, got[r], want[r])
        near += 1
    assert near <= 2
    gts = gt[sel]
    s_gt = dense[torch.arange(len(sel)), gts.cuda()]
    rank = (dense > s_gt[:, None]).sum(1) + ((dense == s_gt[:, None]) & (torch.arange(ng, device="cuda")[None, :] < gts.cuda()[:, None])).sum(1)
    for K in (1, 5, 10):
        hit_lists = (torch.as_tensor(got[:, :K]) == gts[:, None]).any(1)
        assert torch.equal(hit_lists, (rank.cpu() < K))
