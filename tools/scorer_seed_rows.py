"""Scorer step time against the size of the threshold pre-pass (seed_rows), for the shard sizes of 8 / 4 / 2 / 1 GPUs.
    python tools/scorer_seed_rows.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpi_b200 import ops, synthetic as S


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
    row = [f"rows {ng:8d}"]
    for sr in (0, 2048, 4096, 8192, 16384, 32768):
        row.append(f"seed_rows {sr:5d}: {t_ms(lambda: ops.sim_topk(q, g, k, 0, merge=False, seed_rows=sr), n=6 if ng > 2_000_000 else 10):7.3f} ms")
    print(" | ".join(row), flush=True)
