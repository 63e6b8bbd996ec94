"""Tensor-level wrappers over the C ABI (include/lpi_b200.h).  torch is used for device memory and
streams only; every FLOP below this file is a hand-written sm_100a kernel in liblpi_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

(EPI_BIAS_BF16, EPI_BIAS_GELU_BF16, EPI_BIAS_RESID_F32, EPI_F32, EPI_ACC_F32, EPI_DGELU_BF16, EPI_BF16, EPI_BIAS_F32,
 EPI_BIAS_GELU_F32, EPI_DGELU_F32) = range(10)

KERNEL_LAUNCHES = 0          # count of lpi kernels launched (bench.py reports it as gpu_launches)


def _count(n=1):
    global KERNEL_LAUNCHES
    KERNEL_LAUNCHES += n


def _chk(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.LpiError(f"{name} must be a CUDA tensor (lpi_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.LpiError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.LpiError(f"{name} must be contiguous")


# ------------------------------------------------------------------------------------------ GEMM
def gemm(a: torch.Tensor, w: torch.Tensor, epi: int, bias: Optional[torch.Tensor] = None,
         resid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None,
         aux: Optional[torch.Tensor] = None, tile_n: int = 0) -> torch.Tensor:
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T).  a, w both bf16 (lpi_gemm_bf16) or both fp16 (lpi_gemm_f16; every 16-bit
    output / out2 / aux is then fp16 too).  See enum lpi_epilogue in include/lpi_b200.h."""
    _lib.require_device()
    h = a.dtype
    if h not in (torch.bfloat16, torch.float16):
        raise _lib.LpiError(f"a must be bf16 or fp16, got {a.dtype}")
    _chk(a, h, "a")
    _chk(w, h, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    f32_out = epi in (EPI_BIAS_RESID_F32, EPI_F32, EPI_ACC_F32, EPI_BIAS_F32)
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32 if f32_out else h)
    _chk(out, torch.float32 if f32_out else h, "out")
    for t, n in ((bias, "bias"), (resid, "resid")):
        if t is not None:
            _chk(t, torch.float32, n)
    for t, n in ((out2, "out2"), (aux, "aux")):
        if t is not None:
            _chk(t, h, n)
    call("gemm_f16" if h == torch.float16 else "gemm_bf16", ptr(a), ptr(w), M, N, K, epi, ptr(bias), ptr(resid), ptr(out), ptr(out2), ptr(aux),
         out.stride(0), tile_n, stream_ptr())
    _count()
    return out


def gemm_do_delta(a: torch.Tensor, w: torch.Tensor, o_saved: torch.Tensor, delta: torch.Tensor, L: int) -> torch.Tensor:
    """out_proj dgrad with the attention backward's delta fused (lpi_gemm_do_delta): -> d_out [M, N] (a's dtype); delta (fp32, ZEROED
    by the caller, [M / L * N / 64 * L]) receives rowsum_head(d_out * o_saved)."""
    _lib.require_device()
    h = a.dtype
    if h not in (torch.bfloat16, torch.float16):
        raise _lib.LpiError(f"a must be bf16 or fp16, got {a.dtype}")
    for t, n in ((a, "a"), (w, "w"), (o_saved, "o_saved")):
        _chk(t, h, n)
    _chk(delta, torch.float32, "delta")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and o_saved.shape == (M, N) and delta.numel() == (M // L) * (N // 64) * L
    out = torch.empty(M, N, device=a.device, dtype=h)
    call("gemm_do_delta", ptr(a), ptr(w), M, N, K, ptr(out), ptr(o_saved), ptr(delta), L, int(h == torch.float16), stream_ptr())
    _count()
    return out


def gemm_tf32(a: torch.Tensor, w: torch.Tensor, epi: int, bias: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
              tile_n: int = 0) -> torch.Tensor:
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T) with fp32 operands on the TF32 tensor-core path (text tower)."""
    _lib.require_device()
    _chk(a, torch.float32, "a")
    _chk(w, torch.float32, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    bf16_out = epi in (EPI_BIAS_BF16, EPI_BF16)
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16 if bf16_out else torch.float32)
    _chk(out, torch.bfloat16 if bf16_out else torch.float32, "out")
    call("gemm_tf32", ptr(a), ptr(w), M, N, K, epi, ptr(bias), ptr(resid), ptr(out), ptr(out2), ptr(aux), out.stride(0), tile_n,
         stream_ptr())
    _count()
    return out


def gemm_f32(a: torch.Tensor, w: torch.Tensor, epi: int, bias: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
             tile_n: int = 0) -> torch.Tensor:
    """fp32 parity mode (north_star "1e-5 in fp32"): out[M,N] = epilogue(a[M,K] @ w[N,K]^T) with EXACT fp32 products on the SIMT pipes
    (lpi_sgemm_bias_f32 + fp32 QuickGELU kernels); same signature as gemm_tf32, fp32 tensors throughout.  A test mode, not a fast one."""
    _lib.require_device()
    _chk(a, torch.float32, "a")
    _chk(w, torch.float32, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    _chk(out, torch.float32, "out")

    def mm(dst, beta, b):
        call("sgemm_bias_f32", ptr(a), ptr(w), ptr(b), ptr(dst), M, N, K, C.c_longlong(K), C.c_longlong(1), C.c_longlong(1), C.c_longlong(K),
             C.c_longlong(dst.stride(0)), C.c_float(1.0), C.c_float(beta), stream_ptr())
        _count()

    if epi in (EPI_F32, EPI_BF16):
        mm(out, 0.0, None)
    elif epi in (EPI_BIAS_F32, EPI_BIAS_BF16):
        mm(out, 0.0, bias)
    elif epi == EPI_BIAS_RESID_F32:
        if out.data_ptr() != resid.data_ptr():
            out.copy_(resid)
        mm(out, 1.0, bias)
    elif epi == EPI_ACC_F32:
        mm(out, 1.0, None)
    elif epi == EPI_BIAS_GELU_F32:
        z = out2 if out2 is not None else torch.empty_like(out)
        mm(z, 0.0, bias)
        call("quick_gelu_f32", ptr(z), ptr(out), C.c_longlong(z.numel()), stream_ptr())
        _count()
    elif epi == EPI_DGELU_F32:
        mm(out, 0.0, None)
        _chk(aux, torch.float32, "aux")
        call("quick_gelu_bwd_f32", ptr(out), ptr(aux), ptr(out), C.c_longlong(out.numel()), stream_ptr())
        _count()
    else:
        raise _lib.LpiError(f"gemm_f32: epilogue {epi} is not part of the fp32 parity mode")
    return out


def attn_fwd_f32(qkv: torch.Tensor, B: int, L: int, H: int, causal: bool):
    """fp32 parity mode: qkv [B*L, 3*H*64] fp32 -> (out [B*L, H*64] fp32, lse [B*H*L] natural log)."""
    _lib.require_device()
    _chk(qkv, torch.float32, "qkv")
    assert qkv.shape == (B * L, 3 * H * 64)
    out = torch.empty(B * L, H * 64, device=qkv.device, dtype=torch.float32)
    lse = torch.empty(B * H * L, device=qkv.device, dtype=torch.float32)
    call("attn_fwd_f32", ptr(qkv), ptr(out), ptr(lse), B, L, H, int(causal), stream_ptr())
    _count()
    return out, lse


def attn_bwd_f32(qkv, out, d_out, lse, B: int, L: int, H: int, causal: bool):
    for t, n in ((qkv, "qkv"), (out, "out"), (d_out, "d_out"), (lse, "lse")):
        _chk(t, torch.float32, n)
    delta = torch.empty(B * H * L, device=qkv.device, dtype=torch.float32)
    dqkv = torch.empty_like(qkv)
    call("attn_bwd_f32", ptr(qkv), ptr(out), ptr(d_out), ptr(lse), ptr(delta), ptr(dqkv), B, L, H, int(causal), stream_ptr())
    _count(2)
    return dqkv


# ------------------------------------------------------------------------------------------ scorer
def sim_topk_chunks(n_queries: int, n_gallery: int) -> int:
    n = C.c_int()
    call("sim_topk_chunks", n_queries, n_gallery, C.byref(n))
    return n.value


import os as _os

SEED_ROWS = int(_os.environ.get("LPI_SEED_ROWS", "16384"))   # gallery rows scored first to seed the per-query thresholds of the main pass
SEED_MIN_GALLERY = 131072    # below this the warm-up is not worth a second launch
SEED_CHUNKED = _os.environ.get("LPI_SEED_CHUNKED", "0") != "0"  # threshold pre-pass in as many work items as fill whole waves (measured: 11.49 vs 11.52 ms at 625 k rows -- the extra merge launch eats the gain; off)
COOP_THRESHOLDS = _os.environ.get("LPI_COOP_THR", "1") != "0"   # chunks of one launch share their per-query thresholds (see sim_topk)
# ... for shards up to this many rows: the sharing removes ~0.6 ms of list warm-up per launch (5 % of a 625 k-row shard, nothing measurable
# at 5 M rows) but lets the clusters drift apart in the gallery stream, which costs L2 hits (5 M rows: 70 GB of DRAM reads with, 37-41 GB
# without; profiles/r2_scorer_traffic.md)
COOP_MAX_ROWS = int(_os.environ.get("LPI_COOP_MAX_ROWS", "2000000"))


def sim_topk(q: torch.Tensor, g: torch.Tensor, k: int = 10, gallery_offset: int = 0, n_chunks: int = 0,
             merge: bool = True, seed_rows: Optional[int] = None, init_thr: Optional[torch.Tensor] = None,
             out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, coop: Optional[bool] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k gallery rows per query by dot product; q [nq,dim], g [ng,dim] bf16.
    Returns (scores fp32, global idx int32), [nq,k] when merged else [n_chunks,nq,k].
    seed_rows: None = automatic (a pre-pass over the first SEED_ROWS rows of large galleries supplies per-query thresholds, which
    leaves the result unchanged and removes most sorted insertions), 0 = off, n = pre-pass over the first n rows.
    init_thr: [nq] fp32, per query a score that at least k rows of the same logical gallery are known to reach (e.g. the running
    maximum of the k-th scores of the chunks of a streamed gallery seen so far); used instead of a pre-pass.  The lists returned then
    hold only candidates that can still enter the union's top-k (possibly fewer than k; the rest is (-inf, INT_MAX)): merge them with
    the lists that produced the threshold.
    out: (scores fp32, idx int32) buffers of shape [n_chunks, nq, k] for the per-chunk lists (e.g. the two halves of one exchange buffer).
    coop: the chunks of the launch share their per-query thresholds through global memory (lpi_sim_topk_coop_bf16): same merged result,
    fewer sorted insertions; the per-chunk lists may then hold fewer than k entries.  None = automatic (on when n_chunks > 1, dim <= 512)."""
    _lib.require_device()
    _chk(q, torch.bfloat16, "q")
    _chk(g, torch.bfloat16, "g")
    nq, dim = q.shape
    ng = g.shape[0]
    if n_chunks <= 0:
        n_chunks = sim_topk_chunks(nq, ng)
    if seed_rows is None:
        seed_rows = SEED_ROWS if ng >= SEED_MIN_GALLERY else 0
    seed_rows = min(int(seed_rows), ng)
    thr_ptr, thr_stride, seed_scores = None, 1, None
    if init_thr is not None:
        _chk(init_thr, torch.float32, "init_thr")
        if tuple(init_thr.shape) != (nq,):
            raise _lib.LpiError(f"init_thr must be [{nq}], got {tuple(init_thr.shape)}")
        thr_ptr, thr_stride = ptr(init_thr), 1
    elif seed_rows >= k:     # k-th largest per-tile maximum of the sample = the seed (-inf if the sample has fewer than k tiles)
        sc_ = sim_topk_chunks(nq, seed_rows) if SEED_CHUNKED else 1       # whole waves of clusters for the pre-pass too
        seed_scores = torch.empty(sc_, nq, k, device=q.device, dtype=torch.float32)
        seed_idx = torch.empty(sc_, nq, k, device=q.device, dtype=torch.int32)
        if sc_ > 1:
            call("sim_topk_seed_chunks_bf16", ptr(q), ptr(g), nq, seed_rows, dim, k, sc_, ptr(seed_scores), ptr(seed_idx), stream_ptr())
            _count()
            ms_, _ = topk_merge(seed_scores, seed_idx)                   # k largest tile maxima of the union of the chunks
            seed_scores = ms_.unsqueeze(0)
        else:
            call("sim_topk_seed_bf16", ptr(q), ptr(g), nq, seed_rows, dim, k, ptr(seed_scores), ptr(seed_idx), stream_ptr())
            _count()
        thr_ptr, thr_stride = C.c_void_p(seed_scores.data_ptr() + 4 * (k - 1)), k
    if out is not None:
        ps, pi = out
        _chk(ps, torch.float32, "out scores")
        _chk(pi, torch.int32, "out idx")
        if tuple(ps.shape) != (n_chunks, nq, k) or tuple(pi.shape) != (n_chunks, nq, k):
            raise _lib.LpiError(f"out buffers must be [{n_chunks}, {nq}, {k}]")
    else:
        ps = torch.empty(n_chunks, nq, k, device=q.device, dtype=torch.float32)
        pi = torch.empty(n_chunks, nq, k, device=q.device, dtype=torch.int32)
    if coop is None:
        coop = COOP_THRESHOLDS and n_chunks > 1 and dim <= 512 and ng <= COOP_MAX_ROWS
    if coop:
        if init_thr is not None:
            shared = init_thr.clone()
        elif seed_scores is not None:
            shared = seed_scores[0, :, k - 1].contiguous()
        else:
            shared = torch.full((nq,), float("-inf"), device=q.device, dtype=torch.float32)
        call("sim_topk_coop_bf16", ptr(q), ptr(g), nq, ng, dim, k, C.c_longlong(gallery_offset), n_chunks, ptr(shared), ptr(ps), ptr(pi),
             stream_ptr())
    else:
        call("sim_topk_bf16", ptr(q), ptr(g), nq, ng, dim, k, C.c_longlong(gallery_offset), n_chunks, thr_ptr, thr_stride, ptr(ps), ptr(pi),
             stream_ptr())
    _count()
    if not merge:
        return ps, pi
    if n_chunks == 1:
        return ps[0], pi[0]
    return topk_merge(ps, pi)


def topk_merge(part_scores: torch.Tensor, part_idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[n_parts,nq,k] partial lists -> [nq,k] by (score desc, idx asc)."""
    _chk(part_scores, torch.float32, "part_scores")
    _chk(part_idx, torch.int32, "part_idx")
    n_parts, nq, k = part_scores.shape
    max_parts = max(2, 256 // k)          # the merge kernel holds <= 256 candidates per query
    while n_parts > max_parts:            # hierarchical merge for very wide fan-in
        groups = [topk_merge(part_scores[i:i + max_parts].contiguous(), part_idx[i:i + max_parts].contiguous())
                  for i in range(0, n_parts, max_parts)]
        part_scores = torch.stack([g[0] for g in groups])
        part_idx = torch.stack([g[1] for g in groups])
        n_parts = part_scores.shape[0]
    os_ = torch.empty(nq, k, device=part_scores.device, dtype=torch.float32)
    oi = torch.empty(nq, k, device=part_scores.device, dtype=torch.int32)
    call("topk_merge", ptr(part_scores), ptr(part_idx), n_parts, nq, k, ptr(os_), ptr(oi), stream_ptr())
    _count()
    return os_, oi


def topk_merge_recall(packed: torch.Tensor, gt_ptr: torch.Tensor, gt_idx: torch.Tensor, task: torch.Tensor, n_tasks: int,
                      want_rank: bool = False):
    """The tail of a (sharded) search step in ONE launch: k-way merge + Recall@K counters.
    packed: int32 [G, 2, c, nq, k] -- per group (rank) the c per-chunk score lists (fp32 bit patterns) followed by the c index lists,
    i.e. exactly what an all-gather of every rank's [2, c, nq, k] exchange buffer produces.
    Returns (scores [nq,k] fp32, idx [nq,k] int32, counts [n_tasks,4] int32[, rank [nq] int32])."""
    _lib.require_device()
    _chk(packed, torch.int32, "packed")
    for t, n in ((gt_ptr, "gt_ptr"), (gt_idx, "gt_idx"), (task, "task")):
        _chk(t, torch.int32, n)
    G, two, c, nq, k = packed.shape
    if two != 2:
        raise _lib.LpiError("packed must be [G, 2, c, nq, k]")
    os_ = torch.empty(nq, k, device=packed.device, dtype=torch.float32)
    oi = torch.empty(nq, k, device=packed.device, dtype=torch.int32)
    counts = torch.empty(n_tasks, 4, device=packed.device, dtype=torch.int32)
    rank = torch.empty(nq, device=packed.device, dtype=torch.int32) if want_rank else None
    idx_base = C.c_void_p(packed.data_ptr() + 4 * c * nq * k)
    call("topk_merge_recall", ptr(packed), idx_base, G * c, c, C.c_longlong(2 * c * nq * k), nq, k, ptr(os_), ptr(oi), ptr(gt_ptr), ptr(gt_idx),
         ptr(task), n_tasks, ptr(counts), ptr(rank), stream_ptr())
    _count()
    return (os_, oi, counts, rank) if want_rank else (os_, oi, counts)


def topk_rows(scores: torch.Tensor, k: int = 10) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k of each row of a dense fp32 matrix; ties -> lowest index."""
    _lib.require_device()
    _chk(scores, torch.float32, "scores")
    n, m = scores.shape
    os_ = torch.empty(n, k, device=scores.device, dtype=torch.float32)
    oi = torch.empty(n, k, device=scores.device, dtype=torch.int32)
    call("topk_rows_f32", ptr(scores), n, m, C.c_longlong(scores.stride(0)), k, ptr(os_), ptr(oi), stream_ptr())
    _count()
    return os_, oi


def recall_counts(topk_idx: torch.Tensor, gt_ptr: torch.Tensor, gt_idx: torch.Tensor, task: torch.Tensor,
                  n_tasks: int, want_rank: bool = False):
    """counts[n_tasks,4] = #{rank<1}, #{rank<5}, #{rank<10}, n per task."""
    _chk(topk_idx, torch.int32, "topk_idx")
    for t, n in ((gt_ptr, "gt_ptr"), (gt_idx, "gt_idx"), (task, "task")):
        _chk(t, torch.int32, n)
    nq, k = topk_idx.shape
    counts = torch.empty(n_tasks, 4, device=topk_idx.device, dtype=torch.int32)
    rank = torch.empty(nq, device=topk_idx.device, dtype=torch.int32) if want_rank else None
    call("recall_counts", ptr(topk_idx), nq, k, ptr(gt_ptr), ptr(gt_idx), ptr(task), n_tasks, ptr(counts), ptr(rank),
         stream_ptr())
    _count()
    return (counts, rank) if want_rank else counts


def split_bf16(x: torch.Tensor, n_terms: int = 1, role: int = 0) -> torch.Tensor:
    """fp32 [n,dim] -> bf16 [n, n_terms*dim] scorer operand (n_terms=6: exact-product hi/mid/lo split)."""
    _lib.require_device()
    _chk(x, torch.float32, "x")
    n, dim = x.shape
    out = torch.empty(n, n_terms * dim, device=x.device, dtype=torch.bfloat16)
    call("split_bf16", ptr(x), n, dim, n_terms, role, ptr(out), stream_ptr())
    _count()
    return out


def l2_normalize(x: torch.Tensor, want_norm: bool = False):
    _lib.require_device()
    _chk(x, torch.float32, "x")
    n, dim = x.shape
    out = torch.empty_like(x)
    nrm = torch.empty(n, device=x.device, dtype=torch.float32) if want_norm else None
    call("l2_normalize", ptr(x), n, dim, ptr(out), ptr(nrm), stream_ptr())
    _count()
    return (out, nrm) if want_norm else out


# ------------------------------------------------------------------------------------------ attention
def attn_fwd(qkv: torch.Tensor, B: int, L: int, H: int, causal: bool, want_lse: bool = True, want_f32: bool = False):
    """qkv [B*L, 3*H*64] bf16 or fp16 -> (out [B*L, H*64] same dtype, lse2 [B*H*L] fp32 or None[, out_f32])."""
    _lib.require_device()
    h = qkv.dtype
    if h not in (torch.bfloat16, torch.float16):
        raise _lib.LpiError(f"qkv must be bf16 or fp16, got {qkv.dtype}")
    _chk(qkv, h, "qkv")
    assert qkv.shape == (B * L, 3 * H * 64)
    out = torch.empty(B * L, H * 64, device=qkv.device, dtype=h)
    lse = torch.empty(B * H * L, device=qkv.device, dtype=torch.float32) if want_lse else None
    if h == torch.float16:
        if want_f32:
            raise _lib.LpiError("attn_fwd: the fp16 path has no fp32 hand-over")
        call("attn_fwd_f16", ptr(qkv), ptr(out), ptr(lse), B, L, H, int(causal), stream_ptr())
        _count()
        return out, lse
    of = torch.empty(B * L, H * 64, device=qkv.device, dtype=torch.float32) if want_f32 else None
    call("attn_fwd", ptr(qkv), ptr(out), ptr(of), ptr(lse), B, L, H, int(causal), stream_ptr())
    _count()
    return (out, lse, of) if want_f32 else (out, lse)


def attn_bwd(qkv, out, d_out, lse, B: int, L: int, H: int, causal: bool, dqkv: Optional[torch.Tensor] = None, f32: bool = False,
             delta: Optional[torch.Tensor] = None):
    """delta given (fp32 [B*H*L], from gemm_do_delta): `out` is not read and no separate delta pass runs."""
    h = qkv.dtype
    for t, n in ((qkv, "qkv"), (d_out, "d_out")) + (((out, "out"),) if delta is None else ()):
        _chk(t, h, n)
    _chk(lse, torch.float32, "lse")
    given = delta is not None
    if given:
        _chk(delta, torch.float32, "delta")
        assert delta.numel() == B * H * L
        out = None
    else:
        delta = torch.empty(B * H * L, device=qkv.device, dtype=torch.float32)
    if h == torch.float16:
        if f32:
            raise _lib.LpiError("attn_bwd: the fp16 path writes fp16 gradients")
        if dqkv is None:
            dqkv = torch.empty_like(qkv)
        call("attn_bwd_f16", ptr(qkv), ptr(out), ptr(d_out), ptr(lse), ptr(delta), ptr(dqkv), B, L, H, int(causal), stream_ptr())
        _count(1 if given else 2)
        return dqkv
    if dqkv is None:
        dqkv = torch.empty_like(qkv, dtype=torch.float32 if f32 else torch.bfloat16)
    call("attn_bwd", ptr(qkv), ptr(out), ptr(d_out), ptr(lse), ptr(delta), None if f32 else ptr(dqkv), ptr(dqkv) if f32 else None, B, L, H,
         int(causal), stream_ptr())
    _count((2 if L <= 256 else 3) - (1 if given else 0))
    return dqkv


def attn_rowq_fwd(qkv: torch.Tensor, rows: torch.Tensor, B: int, L: int, H: int, causal: bool, x: Optional[torch.Tensor] = None):
    """Attention of the one query row per sample the head reads (last block of a tower; model.py:254-257, prompt_learner.py:57-61).
    qkv [B*L, 3*H*64] bf16 / fp16, rows int32 [B] (global row indices) -> out_rows [B, H*64]; with x (fp32 [B*L, H*64]) also x[rows]."""
    _lib.require_device()
    h = qkv.dtype
    if h not in (torch.bfloat16, torch.float16):
        raise _lib.LpiError(f"qkv must be bf16 or fp16, got {qkv.dtype}")
    _chk(qkv, h, "qkv")
    _chk(rows, torch.int32, "rows")
    assert qkv.shape == (B * L, 3 * H * 64) and rows.shape == (B,)
    out = torch.empty(B, H * 64, device=qkv.device, dtype=h)
    x_rows = None
    if x is not None:
        _chk(x, torch.float32, "x")
        assert x.shape == (B * L, H * 64)
        x_rows = torch.empty(B, H * 64, device=qkv.device, dtype=torch.float32)
    call("attn_rowq_fwd", ptr(qkv), ptr(rows), ptr(out), ptr(x), ptr(x_rows), B, L, H, int(causal), int(h == torch.float16), stream_ptr())
    _count()
    return out, x_rows


def attn_rowq_bwd(qkv: torch.Tensor, rows: torch.Tensor, d_out_rows: torch.Tensor, B: int, L: int, H: int, causal: bool,
                  g_rows: Optional[torch.Tensor] = None, g: Optional[torch.Tensor] = None) -> torch.Tensor:
    """-> dqkv [B*L, 3*H*64] (complete: zero where the read rows have no influence); with g_rows / g also g[rows] = g_rows."""
    h = qkv.dtype
    _chk(qkv, h, "qkv")
    _chk(d_out_rows, h, "d_out_rows")
    _chk(rows, torch.int32, "rows")
    assert d_out_rows.shape == (B, H * 64)
    if g is not None:
        _chk(g, torch.float32, "g")
        _chk(g_rows, torch.float32, "g_rows")
        assert g.shape == (B * L, H * 64) and g_rows.shape == (B, H * 64)
    dqkv = torch.empty_like(qkv)
    call("attn_rowq_bwd", ptr(qkv), ptr(rows), ptr(d_out_rows), ptr(dqkv), ptr(g_rows), ptr(g), B, L, H, int(causal),
         int(h == torch.float16), stream_ptr())
    _count()
    return dqkv


# ------------------------------------------------------------------------------------------ LayerNorm / front ends / heads
LN_EPS = 1e-5


def layernorm_fwd(x, gamma, beta, want_f32=False, want_bf16=True, half_dtype=torch.bfloat16):
    """-> (fp32 copy or None, 16-bit shadow (bf16, or fp16 with half_dtype=torch.float16) or None)"""
    _lib.require_device()
    _chk(x, torch.float32, "x")
    M, D = x.shape
    of = torch.empty_like(x) if want_f32 else None
    ob = torch.empty(M, D, device=x.device, dtype=half_dtype) if want_bf16 else None
    call("layernorm_fwd_f16" if half_dtype == torch.float16 else "layernorm_fwd", ptr(x), ptr(gamma), ptr(beta), ptr(of), ptr(ob),
         C.c_longlong(M), D, C.c_float(LN_EPS), stream_ptr())
    _count()
    return of, ob


def layernorm_bwd(dy, x, gamma, g, g_bf16=None, accumulate=True, grad_scale: Optional[float] = None):
    """g = (accumulate ? g : 0) + dLN(dy; x, gamma), in place; optional 16-bit shadow.  dy may be fp32 or the tower's 16-bit type
    (bf16 / fp16, as written by a dgrad GEMM).  With grad_scale (fp16 gradient path): dy is grad_scale * (true dy), g stays true
    scale, the fp16 shadow is grad_scale * g."""
    for t, n in ((x, "x"), (g, "g")):
        _chk(t, torch.float32, n)
    M, D = x.shape
    half = torch.float16 if grad_scale is not None else torch.bfloat16
    if g_bf16 is not None:
        _chk(g_bf16, half, "g16")
    if dy.dtype in (torch.bfloat16, torch.float16):
        _chk(dy, half, "dy")
        call("layernorm_bwd_dy16", ptr(dy), int(half == torch.float16), ptr(x), ptr(gamma), ptr(g), ptr(g_bf16), C.c_longlong(M), D,
             C.c_float(LN_EPS), int(accumulate), C.c_float(grad_scale if grad_scale is not None else 1.0), stream_ptr())
    elif grad_scale is not None:
        _chk(dy, torch.float32, "dy")
        call("layernorm_bwd_f16", ptr(dy), ptr(x), ptr(gamma), ptr(g), ptr(g_bf16), C.c_longlong(M), D, C.c_float(LN_EPS), int(accumulate),
             C.c_float(grad_scale), stream_ptr())
    else:
        _chk(dy, torch.float32, "dy")
        call("layernorm_bwd", ptr(dy), ptr(x), ptr(gamma), ptr(g), ptr(g_bf16), C.c_longlong(M), D, C.c_float(LN_EPS), int(accumulate),
             stream_ptr())
    _count()
    return g


def im2col_patches(images: torch.Tensor, patch: int, half_dtype=torch.bfloat16) -> torch.Tensor:
    """half_dtype: bf16 / fp16 GEMM operand rows, or torch.float32 for the fp32 parity mode."""
    _lib.require_device()
    _chk(images, torch.float32, "images")
    B, ch, R, _ = images.shape
    assert ch == 3
    G = R // patch
    out = torch.empty(B * G * G, 3 * patch * patch, device=images.device, dtype=half_dtype)
    name = {torch.float16: "im2col_patches_f16", torch.float32: "im2col_patches_f32"}.get(half_dtype, "im2col_patches")
    call(name, ptr(images), ptr(out), B, R, patch, stream_ptr())
    _count()
    return out


def assemble_vision(patch_emb, cls, pos, prompt_table, sel, ln_g, ln_b, B, n_patch, P, D):
    x = torch.empty(B * (1 + P + n_patch), D, device=patch_emb.device, dtype=torch.float32)
    call("assemble_vision", ptr(patch_emb), ptr(cls), ptr(pos), ptr(prompt_table), ptr(sel), ptr(ln_g), ptr(ln_b), ptr(x), B, n_patch, P, D,
         C.c_float(LN_EPS), stream_ptr())
    _count()
    return x


def _factor_args(fac):
    """fac = (dim1_share [T, Lp, r], dim2 [T, P, r], dim3 [T, D, r], scale) -> ctypes args of the *_factors entry points."""
    d1, d2, d3, scale = fac
    for t in (d1, d2, d3):
        _chk(t, torch.float32, "factor")
    if d1.dim() != 3 or d2.dim() != 3 or d3.dim() != 3 or not (d1.shape[0] == d2.shape[0] == d3.shape[0]) or not (d1.shape[2] == d2.shape[2] == d3.shape[2]):
        raise _lib.LpiError("factors must be dim1 [T, Lp, r], dim2 [T, P, r], dim3 [T, D, r]")
    return ptr(d1), ptr(d2), ptr(d3), d1.shape[2], d1.shape[1], C.c_float(float(scale))


def assemble_vision_factors(patch_emb, cls, pos, fac, sel, ln_g, ln_b, B, n_patch, D):
    """assemble_vision with the prompt rows reconstructed in the kernel from the DecomposedPrompt factors (no table)."""
    P = fac[1].shape[1]
    x = torch.empty(B * (1 + P + n_patch), D, device=patch_emb.device, dtype=torch.float32)
    call("assemble_vision_factors", ptr(patch_emb), ptr(cls), ptr(pos), *_factor_args(fac), ptr(sel), ptr(ln_g), ptr(ln_b), ptr(x), B, n_patch, P, D,
         C.c_float(LN_EPS), stream_ptr())
    _count()
    return x


def assemble_vision_factors_bwd(g, fac, sel, ln_g, B, L, n_tables, D):
    P = fac[1].shape[1]
    d = torch.empty(n_tables, P, D, device=g.device, dtype=torch.float32)
    call("assemble_vision_factors_bwd", ptr(g), *_factor_args(fac), ptr(sel), ptr(ln_g), ptr(d), B, L, P, n_tables, D, C.c_float(LN_EPS), stream_ptr())
    _count()
    return d


def assemble_text_factors(emb, tokens, pos, fac, sel, B, L, D):
    _lib.require_device()
    _chk(tokens, torch.int64, "tokens")
    P = fac[1].shape[1]
    x = torch.empty(B * L, D, device=emb.device, dtype=torch.float32)
    call("assemble_text_factors", ptr(emb), ptr(tokens), ptr(pos), *_factor_args(fac), ptr(sel), ptr(x), B, L, P, D, stream_ptr())
    _count()
    return x


def assemble_vision_bwd(g, prompt_table, sel, ln_g, B, L, P, n_tables, D):
    d = torch.empty(n_tables, P, D, device=g.device, dtype=torch.float32)
    call("assemble_vision_bwd", ptr(g), ptr(prompt_table), ptr(sel), ptr(ln_g), ptr(d), B, L, P, n_tables, D, C.c_float(LN_EPS), stream_ptr())
    _count()
    return d


def assemble_text(emb, tokens, pos, ctx_table, sel, B, L, P, D):
    _lib.require_device()
    _chk(tokens, torch.int64, "tokens")
    x = torch.empty(B * L, D, device=emb.device, dtype=torch.float32)
    call("assemble_text", ptr(emb), ptr(tokens), ptr(pos), ptr(ctx_table), ptr(sel), ptr(x), B, L, P, D, stream_ptr())
    _count()
    return x


def sum_prompt_rows(g, sel, B, L, P, n_tables, D):
    """d_table[t,p,:] = sum_{b: sel[b]=t} g[b,1+p,:]  (backward of the text splice and of deep-prompt injection)."""
    d = torch.empty(n_tables, P, D, device=g.device, dtype=torch.float32)
    call("assemble_text_bwd", ptr(g), ptr(sel), ptr(d), B, L, P, n_tables, D, stream_ptr())
    _count()
    return d


def inject_prompt_rows(x, prompt, sel, B, L, P, D):
    call("inject_prompt_rows", ptr(x), ptr(prompt), ptr(sel), B, L, P, D, stream_ptr())
    _count()


def head_fwd(x, row_idx, ln_g, ln_b, proj):
    _chk(x, torch.float32, "x")
    _chk(row_idx, torch.int32, "row_idx")
    B = row_idx.shape[0]
    D, E = proj.shape
    z = torch.empty(B, E, device=x.device, dtype=torch.float32)
    f = torch.empty(B, E, device=x.device, dtype=torch.float32)
    call("head_fwd", ptr(x), ptr(row_idx), ptr(ln_g), ptr(ln_b), ptr(proj), ptr(z), ptr(f), B, D, E, C.c_float(LN_EPS), stream_ptr())
    _count(2)
    return f, z


def head_fwd_select(x, row_idx, ln_g, ln_b, proj, centers):
    """head_fwd + task-id selection in its last kernel: -> (feat, z, sel int32 [B]); centers [T, C, E] fp32 (sprompt.py:336-368)."""
    _chk(x, torch.float32, "x")
    _chk(row_idx, torch.int32, "row_idx")
    _chk(centers, torch.float32, "centers")
    B = row_idx.shape[0]
    D, E = proj.shape
    T, Cn, _ = centers.shape
    z = torch.empty(B, E, device=x.device, dtype=torch.float32)
    f = torch.empty(B, E, device=x.device, dtype=torch.float32)
    sel = torch.empty(B, device=x.device, dtype=torch.int32)
    call("head_fwd_select", ptr(x), ptr(row_idx), ptr(ln_g), ptr(ln_b), ptr(proj), ptr(z), ptr(f), ptr(centers), T, Cn, ptr(sel), B, D, E,
         C.c_float(LN_EPS), stream_ptr())
    _count(2)
    return f, z, sel


def head_bwd(dfeat, dz, z, x, row_idx, ln_g, proj, g, g_bf16=None, grad_scale: Optional[float] = None):
    """Rows row_idx of g are assigned; g_bf16 = optional 16-bit shadow (fp16 scaled by grad_scale when grad_scale is given)."""
    for t, n in ((dfeat, "dfeat"), (dz, "dz")):
        if t is not None:
            _chk(t, torch.float32, n)
    B = row_idx.shape[0]
    D, E = proj.shape
    if grad_scale is not None:
        if g_bf16 is not None:
            _chk(g_bf16, torch.float16, "g_f16")
        call("head_bwd_f16", ptr(dfeat), ptr(dz), ptr(z), ptr(x), ptr(row_idx), ptr(ln_g), ptr(proj), ptr(g), ptr(g_bf16), C.c_float(grad_scale),
             B, D, E, C.c_float(LN_EPS), stream_ptr())
    else:
        call("head_bwd", ptr(dfeat), ptr(dz), ptr(z), ptr(x), ptr(row_idx), ptr(ln_g), ptr(proj), ptr(g), ptr(g_bf16), B, D, E,
             C.c_float(LN_EPS), stream_ptr())
    _count(2)


# ------------------------------------------------------------------------------------------ DecomposedPrompt
def prompt_fwd(d1, d2v, d2t, d3v, d3t):
    _lib.require_device()
    for t in (d1, d2v, d2t, d3v, d3t):
        _chk(t, torch.float32, "factor")
    L, r = d1.shape
    P, Dv, Dt = d2v.shape[0], d3v.shape[0], d3t.shape[0]
    vis = torch.empty(L, P, Dv, device=d1.device, dtype=torch.float32)
    txt = torch.empty(L, P, Dt, device=d1.device, dtype=torch.float32)
    call("prompt_fwd", ptr(d1), ptr(d2v), ptr(d2t), ptr(d3v), ptr(d3t), ptr(vis), ptr(txt), L, P, Dv, Dt, r, stream_ptr())
    _count()
    return vis, txt


def prompt_bwd(d1, d2v, d2t, d3v, d3t, g_vis, g_txt):
    _chk(g_vis, torch.float32, "g_vis")
    _chk(g_txt, torch.float32, "g_txt")
    L, r = d1.shape
    P, Dv, Dt = d2v.shape[0], d3v.shape[0], d3t.shape[0]
    ws = torch.empty(2 * L * P * r, device=d1.device, dtype=torch.float32)
    outs = [torch.empty_like(t) for t in (d1, d2v, d2t, d3v, d3t)]
    call("prompt_bwd", ptr(d1), ptr(d2v), ptr(d2t), ptr(d3v), ptr(d3t), ptr(g_vis), ptr(g_txt), ptr(ws), *[ptr(o) for o in outs], L, P, Dv,
         Dt, r, stream_ptr())
    _count(3)
    return outs


# ------------------------------------------------------------------------------------------ losses / optimiser
def sgemm(a: torch.Tensor, b: torch.Tensor, alpha: float = 1.0, out: Optional[torch.Tensor] = None, beta: float = 0.0):
    """out[M,N] = alpha * a[M,K] @ b[K,N] + beta*out for arbitrary-strided 2-D fp32 views (exact fp32 products)."""
    _lib.require_device()
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.is_cuda and b.is_cuda
    M, K = a.shape
    K2, N = b.shape
    assert K == K2
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    call("sgemm_f32", ptr(a), ptr(b), ptr(out), M, N, K, C.c_longlong(a.stride(0)), C.c_longlong(a.stride(1)), C.c_longlong(b.stride(0)),
         C.c_longlong(b.stride(1)), C.c_longlong(out.stride(0)), C.c_float(alpha), C.c_float(beta), stream_ptr())
    _count()
    return out


def clip_loss_logits(logits: torch.Tensor, weight: float = 1.0, want_grad: bool = True):
    """(loss[1], dlogits or None) of weight * 1/2 [CE(S, arange) + CE(S^T, arange)]."""
    _lib.require_device()
    _chk(logits, torch.float32, "logits")
    n = logits.shape[0]
    assert logits.shape == (n, n)
    ws = torch.empty(2 * n, device=logits.device, dtype=torch.float32)
    loss = torch.empty(1, device=logits.device, dtype=torch.float32)
    d = torch.empty_like(logits) if want_grad else None
    call("clip_loss_logits", ptr(logits), n, C.c_float(weight), ptr(ws), ptr(loss), ptr(d), stream_ptr())
    _count(3 if want_grad else 2)
    return loss, d


def sim_infonce_fwd_bwd(img_f: torch.Tensor, txt_f: torch.Tensor, scale: float, row0: int = 0, n_local: Optional[int] = None,
                        weight: float = 1.0, want_grad: bool = True, want_logits: bool = False):
    """ONE launch: loss = weight * ClipLoss(scale * I @ T^T) over the global batch + gradients of the local feature rows
    -> (loss [1], d_img [n_local, E] or None, d_txt [n_local, E] or None, logits [n, n] or None)."""
    _lib.require_device()
    _chk(img_f, torch.float32, "img_f")
    _chk(txt_f, torch.float32, "txt_f")
    n, E = img_f.shape
    assert txt_f.shape == (n, E)
    n_local = n if n_local is None else n_local
    ws = torch.empty(4 * n, device=img_f.device, dtype=torch.float32)
    loss = torch.empty(1, device=img_f.device, dtype=torch.float32)
    logits = torch.empty(n, n, device=img_f.device, dtype=torch.float32) if want_logits else None
    d_img = torch.empty(n_local, E, device=img_f.device, dtype=torch.float32) if want_grad else None
    d_txt = torch.empty(n_local, E, device=img_f.device, dtype=torch.float32) if want_grad else None
    call("sim_infonce_fwd_bwd", ptr(img_f), ptr(txt_f), n, E, C.c_float(scale), C.c_float(weight), row0, n_local, ptr(ws),
         C.c_void_p(ws.data_ptr() + 8 * n), ptr(loss), ptr(logits), ptr(d_img), ptr(d_txt), stream_ptr())
    _count()
    return loss, d_img, d_txt, logits


def row_mean(x: torch.Tensor, scale: float = 1.0):
    _chk(x, torch.float32, "x")
    rows, D = x.shape
    out = torch.empty(rows, device=x.device, dtype=torch.float32)
    call("row_mean", ptr(x), ptr(out), rows, D, C.c_float(scale), stream_ptr())
    _count()
    return out


def add_rowconst(G: torch.Tensor, v: torch.Tensor, alpha: float, accumulate: bool = True):
    rows, D = G.shape
    call("add_rowconst", ptr(G), ptr(v), C.c_longlong(rows), D, C.c_float(alpha), int(accumulate), stream_ptr())
    _count()


def task_loss(X: torch.Tensor, target: torch.Tensor, temperature: float, weight: float, loss_out: torch.Tensor, loss_accumulate: bool,
              grad_last_row: Optional[torch.Tensor], grad_accumulate: bool = True, n_part: int = 64):
    """nt_bxent_loss over rows of X [R,n]; loss_out (+)= weight*loss; grad_last_row (+)= d/dX[R-1]."""
    _chk(X, torch.float32, "X")
    _chk(target, torch.int32, "target")
    R, n = X.shape
    ws = torch.empty(n_part * R * R, device=X.device, dtype=torch.float32)
    coef = torch.empty(R, device=X.device, dtype=torch.float32)
    call("task_loss", ptr(X), R, C.c_longlong(n), ptr(target), C.c_float(temperature), C.c_float(weight), ptr(ws), n_part, ptr(loss_out),
         int(loss_accumulate), ptr(coef), ptr(grad_last_row), int(grad_accumulate), stream_ptr())
    _count(3 if grad_last_row is not None else 2)
    return coef


def sgd_momentum_step(w, g, v, lr, momentum, weight_decay, first_step):
    for t, n in ((w, "w"), (g, "g"), (v, "v")):
        _chk(t, torch.float32, n)
    call("sgd_momentum_step", ptr(w), ptr(g), ptr(v), C.c_longlong(w.numel()), C.c_float(lr), C.c_float(momentum), C.c_float(weight_decay),
         int(first_step), stream_ptr())
    _count()


def nearest_center_l1(feats: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    """feats [B,E], centers [T,C,E] fp32 -> int64 [B] (sprompt.py:336-368)."""
    _lib.require_device()
    _chk(feats, torch.float32, "feats")
    _chk(centers, torch.float32, "centers")
    B, E = feats.shape
    T, Cn, _ = centers.shape
    sel = torch.empty(B, device=feats.device, dtype=torch.int64)
    call("nearest_center_l1", ptr(feats), ptr(centers), B, T, Cn, E, ptr(sel), stream_ptr())
    _count()
    return sel
