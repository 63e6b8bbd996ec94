"""Scorer main pass at 5 M rows for several gallery-chunk counts (time), used with ncu for the DRAM traffic of each.
    python tools/scorer_chunks_probe.py [n_chunks ...]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpi_b200 import ops, synthetic as S
dev = torch.device("cuda")
nq, k, ng = 25000, 10, int(os.environ.get("NG", "5000000"))
g, q, gt = S.make_gallery_shard(ng, 0, ng, nq, 512, device=dev)
for c in [int(x) for x in sys.argv[1:]] or [3, 6, 12, 24]:
    for _ in range(2):
        ops.sim_topk(q, g, k, 0, c, merge=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.sim_topk(q, g, k, 0, c, merge=False)
    e1.record(); torch.cuda.synchronize()
    print(f"rows {ng} chunks {c:3d}: {e0.elapsed_time(e1) / 5:8.3f} ms (seed + main)", flush=True)
