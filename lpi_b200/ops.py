"""Tensor-level wrappers over the C ABI (include/lpi_b200.h).  torch is used for device memory and
streams only; every FLOP below this file is a hand-written sm_100a kernel in liblpi_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

EPI_BIAS_BF16, EPI_BIAS_GELU_BF16, EPI_BIAS_RESID_F32, EPI_F32, EPI_ACC_F32, EPI_DGELU_BF16, EPI_BF16, EPI_BIAS_F32 = range(8)

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
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T).  a, w bf16.  See enum lpi_epilogue in include/lpi_b200.h."""
    _lib.require_device()
    _chk(a, torch.bfloat16, "a")
    _chk(w, torch.bfloat16, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    f32_out = epi in (EPI_BIAS_RESID_F32, EPI_F32, EPI_ACC_F32, EPI_BIAS_F32)
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32 if f32_out else torch.bfloat16)
    _chk(out, torch.float32 if f32_out else torch.bfloat16, "out")
    for t, n in ((bias, "bias"), (resid, "resid")):
        if t is not None:
            _chk(t, torch.float32, n)
    for t, n in ((out2, "out2"), (aux, "aux")):
        if t is not None:
            _chk(t, torch.bfloat16, n)
    call("gemm_bf16", ptr(a), ptr(w), M, N, K, epi, ptr(bias), ptr(resid), ptr(out), ptr(out2), ptr(aux),
         out.stride(0), tile_n, stream_ptr())
    _count()
    return out


# ------------------------------------------------------------------------------------------ scorer
def sim_topk_chunks(n_queries: int, n_gallery: int) -> int:
    n = C.c_int()
    call("sim_topk_chunks", n_queries, n_gallery, C.byref(n))
    return n.value


def sim_topk(q: torch.Tensor, g: torch.Tensor, k: int = 10, gallery_offset: int = 0, n_chunks: int = 0,
             merge: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k gallery rows per query by dot product; q [nq,dim], g [ng,dim] bf16.
    Returns (scores fp32, global idx int32), [nq,k] when merged else [n_chunks,nq,k]."""
    _lib.require_device()
    _chk(q, torch.bfloat16, "q")
    _chk(g, torch.bfloat16, "g")
    nq, dim = q.shape
    ng = g.shape[0]
    if n_chunks <= 0:
        n_chunks = sim_topk_chunks(nq, ng)
    ps = torch.empty(n_chunks, nq, k, device=q.device, dtype=torch.float32)
    pi = torch.empty(n_chunks, nq, k, device=q.device, dtype=torch.int32)
    call("sim_topk_bf16", ptr(q), ptr(g), nq, ng, dim, k, C.c_longlong(gallery_offset), n_chunks, ptr(ps), ptr(pi),
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
