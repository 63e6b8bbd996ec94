"""Where the per-launch fixed cost of the scorer goes (the 8-GPU strong-scaling limiter): the threshold pre-pass and the main pass timed
separately (CUDA events, 10 launches) for gallery shards of 625 k (= 5 M / 8) ... 5 M rows x 25 000 queries, with the seed variants.
    python tools/scorer_fixed_cost.py"""
import os, sys, ctypes as C, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpi_b200 import ops, synthetic as S
from lpi_b200._lib import call, ptr, stream_ptr

def t_ms(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

dev = torch.device("cuda")
nq, k = 25000, 10
full, q, gt = S.make_gallery_shard(5_000_000, 0, 5_000_000, nq, 512, device=dev)
for ng in (625_000, 1_250_000, 2_500_000, 5_000_000):
    g = full[:ng]
    fl = 2.0 * nq * ng * 512
    nch = ops.sim_topk_chunks(nq, ng)
    row = [f"rows {ng:8d} chunks {nch}"]
    total = t_ms(lambda: ops.sim_topk(q, g, k, 0, merge=False))
    row.append(f"seed+main {total:7.3f} ms ({fl / total / 1e9:6.0f} TF)")
    noseed = t_ms(lambda: ops.sim_topk(q, g, k, 0, merge=False, seed_rows=0))
    row.append(f"main unseeded {noseed:7.3f}")
    # main pass alone with a precomputed seed
    ss = torch.empty(1, nq, k, device=dev); si = torch.empty(1, nq, k, device=dev, dtype=torch.int32)
    call("sim_topk_seed_bf16", ptr(q), ptr(g), nq, 16384, 512, k, ptr(ss), ptr(si), stream_ptr())
    thr = ss[0, :, k - 1].contiguous()
    main = t_ms(lambda: ops.sim_topk(q, g, k, 0, merge=False, init_thr=thr))
    row.append(f"main (given seed) {main:7.3f} ({fl / main / 1e9:6.0f} TF)")
    for sr in (4096, 8192, 16384):
        seed = t_ms(lambda: call("sim_topk_seed_bf16", ptr(q), ptr(g), nq, sr, 512, k, ptr(ss), ptr(si), stream_ptr()))
        row.append(f"seed[{sr}] {seed:6.3f}")
    # exact final thresholds (upper bound of what any seed can give): the true k-th score
    sc, ix = ops.sim_topk(q, g, k, 0)
    best = t_ms(lambda: ops.sim_topk(q, g, k, 0, merge=False, init_thr=sc[:, k - 1].contiguous()))
    row.append(f"main (exact thr) {best:7.3f}")
    for c in (1, 2, 4, 6):
        try:
            tt = t_ms(lambda: ops.sim_topk(q, g, k, 0, n_chunks=c, merge=False, init_thr=thr))
            row.append(f"chunks={c}: {tt:7.3f}")
        except Exception as e:
            row.append(f"chunks={c}: n/a")
    print(" | ".join(row), flush=True)
