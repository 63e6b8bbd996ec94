"""Recall@K evaluation of the LPI learner on the B200 (host side of the scorer kernels).

Mirrors `SPrompts._evaluate_retrieval` / `SPrompts.itm_eval`
(/root/reference/retrieval/methods/sprompt.py:433-548, 550-646): same arguments, same result dict
`{'mscoco': {'i2t': {task: [r1, r5, r10]}, 't2i': {...}}}` with Python floats `100.0 * hits / n`.

What changes underneath: the reference builds the dense score matrix on the GPU, copies both
orientations to the host and runs a full `np.argsort` per row.  Here
  * `itm_eval(scores_i2t, scores_t2i, ...)` keeps the dense-matrix signature (drop-in) but ranks on the
    device with a row top-k kernel (Recall@1/5/10 only needs the first 10 positions);
  * `itm_eval_features(image_feats, text_feats, ...)` never forms the score matrix: similarity GEMM with
    the top-k in its epilogue (`lpi_sim_topk_bf16`), optionally with the gallery sharded over the ranks
    of a process group -- one all-gather of the per-shard candidates, then a k-way merge.
Tie rule: (score desc, gallery index asc); the reference's `np.argsort(...)[::-1]` leaves ties
unspecified.  No CPU fallback: every compute call needs liblpi_b200.so and an sm_100a device.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops

TOPK = 10          # Recall@1/5/10 never looks past position 9 (sprompt.py:577-579,605-607)


# ------------------------------------------------------------------------------------------ host logic
def shard_bounds(n: int, world_size: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous balanced row range [lo, hi) of shard `rank`; boundaries are multiples of `align`
    (except the last).  Global gallery index = lo + local index."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    units = (n + align - 1) // align
    base, rem = divmod(units, world_size)
    lo_u = rank * base + min(rank, rem)
    hi_u = lo_u + base + (1 if rank < rem else 0)
    return min(n, lo_u * align), min(n, hi_u * align)


def gt_csr(gt: Sequence[Sequence[int]]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Ground-truth lists -> CSR (ptr[n+1], idx[nnz]) int32 host tensors."""
    ptr = [0]
    flat: List[int] = []
    for g in gt:
        flat.extend(int(x) for x in g)
        ptr.append(len(flat))
    return torch.tensor(ptr, dtype=torch.int32), torch.tensor(flat, dtype=torch.int32)


def _as_int_list(x) -> List[int]:
    if isinstance(x, torch.Tensor):
        return [int(v) for v in x.tolist()]
    return [int(v) for v in x]


def recall_dict(counts_i2t: torch.Tensor, counts_t2i: torch.Tensor) -> Dict:
    """counts[task] = (#rank<1, #rank<5, #rank<10, n) -> the reference's result dict (sprompt.py:638-646).
    A task without items raises ZeroDivisionError exactly like the reference (sprompt.py:581, C19)."""
    def side(c):
        out = {}
        for task, (h1, h5, h10, n) in enumerate(c.tolist()):
            out[task] = [100.0 * h1 / n, 100.0 * h5 / n, 100.0 * h10 / n]
        return out
    return {"mscoco": {"i2t": side(counts_i2t), "t2i": side(counts_t2i)}}


# ------------------------------------------------------------------------------------------ device side
def _counts(topk_idx: torch.Tensor, gt: Sequence[Sequence[int]], category, task_num: int, want_rank=False):
    dev = topk_idx.device
    ptr, idx = gt_csr(gt)
    task = torch.tensor(_as_int_list(category), dtype=torch.int32)
    return ops.recall_counts(topk_idx, ptr.to(dev), idx.to(dev), task.to(dev), task_num, want_rank=want_rank)


def itm_eval(scores_i2t, scores_t2i, txt2img, img2txt, category_i, category_t, task_num: int,
             device: Optional[torch.device] = None) -> Dict:
    """Drop-in for `SPrompts.itm_eval` (sprompt.py:550-646); `task_num` is the learner's `cur_id + 1`.
    scores_*: dense fp32 matrices (numpy or torch, host or device)."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    s_i2t = torch.as_tensor(scores_i2t, dtype=torch.float32).to(dev, non_blocking=True).contiguous()
    s_t2i = torch.as_tensor(scores_t2i, dtype=torch.float32).to(dev, non_blocking=True).contiguous()
    n_i, n_t = s_i2t.shape
    _, top_i = ops.topk_rows(s_i2t, min(TOPK, n_t))
    _, top_t = ops.topk_rows(s_t2i, min(TOPK, n_i))
    c_i = _counts(top_i, [img2txt[i] for i in range(n_i)], category_i, task_num)
    c_t = _counts(top_t, [[txt2img[t]] for t in range(n_t)], category_t, task_num)
    return recall_dict(c_i.cpu(), c_t.cpu())


def prepare_operand(x: torch.Tensor, precision: str, role: int) -> torch.Tensor:
    """fp32/bf16 feature rows -> bf16 scorer operand.  precision 'bf16': round once (the stored-gallery
    format of the large sweeps); 'fp32': exact-product 3-way split over K = 6*dim (role 0 = query side,
    1 = gallery side) so the tensor-core score equals the fp32 dot product to accumulation-order noise."""
    if x.dtype == torch.bfloat16:
        if precision != "bf16":
            raise ops._lib.LpiError("bf16 features can only be scored with precision='bf16'")
        return x.contiguous()
    x = x.to(torch.float32).contiguous()
    if precision == "bf16":
        return ops.split_bf16(x, 1, role)
    if precision == "fp32":
        return ops.split_bf16(x, 6, role)
    raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")


def search_topk(queries: torch.Tensor, gallery_shard: torch.Tensor, k: int = TOPK, precision: str = "bf16",
                gallery_offset: int = 0, group=None, prepared: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k gallery rows for every query by dot product, (score desc, global index asc).

    `gallery_shard` is this rank's contiguous slice of the gallery starting at global row `gallery_offset`;
    with `group` (torch.distributed) the per-shard candidates are all-gathered once ([Q,k] fp32 + int32 per
    rank) and merged on every rank -- the only exchange step of the path."""
    q = queries if prepared else prepare_operand(queries, precision, 0)
    if gallery_shard.shape[0] == 0:      # a rank whose shard is empty still takes part in the exchange
        sc = torch.full((q.shape[0], k), float("-inf"), device=q.device, dtype=torch.float32)
        ix = torch.full((q.shape[0], k), 0x7FFFFFFF, device=q.device, dtype=torch.int32)
    else:
        g = gallery_shard if prepared else prepare_operand(gallery_shard, precision, 1)
        sc, ix = ops.sim_topk(q, g, k, gallery_offset)
    if group is None:
        return sc, ix
    return merge_across_ranks(sc, ix, group)


def gather_candidates(scores: torch.Tensor, idx: torch.Tensor, group) -> Tuple[torch.Tensor, torch.Tensor]:
    """The exchange step: every rank's [Q,k] candidate lists -> [world, Q, k] on every rank
    (NCCL over NVLink on the B200 box; backend-agnostic, so the gloo CPU tests cover it too)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    nq, k = scores.shape
    all_s = torch.empty(world, nq, k, device=scores.device, dtype=torch.float32)
    all_i = torch.empty(world, nq, k, device=scores.device, dtype=torch.int32)
    dist.all_gather_into_tensor(all_s.view(world * nq, k), scores.contiguous(), group=group)
    dist.all_gather_into_tensor(all_i.view(world * nq, k), idx.contiguous(), group=group)
    return all_s, all_i


def merge_across_ranks(scores: torch.Tensor, idx: torch.Tensor, group) -> Tuple[torch.Tensor, torch.Tensor]:
    import torch.distributed as dist

    if dist.get_world_size(group) == 1:
        return scores, idx
    return ops.topk_merge(*gather_candidates(scores, idx, group))


def exchange_buffer(n_chunks: int, n_queries: int, k: int, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """One allocation holding a rank's per-chunk candidate lists the way the exchange wants them: int32 [2, c, Q, k] = the c score
    lists (fp32 bit patterns) followed by the c index lists.  Returns (buffer, scores view fp32 [c,Q,k], idx view int32 [c,Q,k]);
    pass the views to `ops.sim_topk(..., merge=False, out=(scores, idx))`."""
    buf = torch.empty(2, n_chunks, n_queries, k, device=device, dtype=torch.int32)
    return buf, buf[0].view(torch.float32), buf[1]


def merge_recall(buf: torch.Tensor, gt_ptr: torch.Tensor, gt_idx: torch.Tensor, task: torch.Tensor, n_tasks: int, group=None,
                 want_rank: bool = False):
    """Tail of a gallery-sharded search step: ONE all-gather of every rank's exchange buffer (scores and indices of all its chunks
    together, 2 * c * Q * k * 4 bytes per rank) and ONE kernel that merges the world * c lists per query by (score desc, index asc)
    and counts Recall@1/5/10 per task.  -> (scores [Q,k], idx [Q,k], counts [n_tasks,4][, rank [Q]])."""
    packed = buf.unsqueeze(0)
    from .comm import LpiComm

    if isinstance(group, LpiComm):                       # the C-ABI exchange (lpi_comm_allgather) instead of torch.distributed
        if group.n_ranks > 1:
            packed = group.all_gather(buf)
    elif group is not None:
        import torch.distributed as dist

        world = dist.get_world_size(group)
        if world > 1:
            packed = torch.empty(world, *buf.shape, device=buf.device, dtype=buf.dtype)
            dist.all_gather_into_tensor(packed, buf.contiguous(), group=group)
    return ops.topk_merge_recall(packed.contiguous(), gt_ptr, gt_idx, task, n_tasks, want_rank)


def itm_eval_features(image_feats: torch.Tensor, text_feats: torch.Tensor, txt2img, img2txt, category_i, category_t,
                      task_num: int, precision: str = "fp32", group=None) -> Dict:
    """Same result as `itm_eval((I @ T^T), (I @ T^T)^T, ...)` (sprompt.py:509,544-546) without the score
    matrix.  With `group`, every rank passes the FULL query sets and scores its own gallery shard
    (rows `shard_bounds(n, world, rank)` of the other modality)."""
    n_i, n_t = image_feats.shape[0], text_feats.shape[0]
    rank, world = 0, 1
    if group is not None:
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
    out = []
    for q, g, gt, cat in ((image_feats, text_feats, [img2txt[i] for i in range(n_i)], category_i),
                          (text_feats, image_feats, [[txt2img[t]] for t in range(n_t)], category_t)):
        lo, hi = shard_bounds(g.shape[0], world, rank)
        _, top = search_topk(q, g[lo:hi], min(TOPK, g.shape[0]), precision, lo, group if world > 1 else None)
        out.append(_counts(top, gt, cat, task_num).cpu())
    return recall_dict(out[0], out[1])
