// Retrieval-side bandwidth kernels: top-k merge, dense-row top-k, Recall@K bookkeeping, operand preparation.
// Reference: methods/sprompt.py:433-646 (_evaluate_retrieval / itm_eval) -- there a full np.argsort per row on the host.
#include "ptx.cuh"
#include "lpi_internal.h"
#include <math_constants.h>

namespace lpi {

// (score desc, index asc) strict ordering: is a better than b?
__device__ __forceinline__ bool better(float sa, int ia, float sb, int ib) { return sa > sb || (sa == sb && ia < ib); }

// warp-wide argbest over one candidate per lane
__device__ __forceinline__ void warp_best(float& s, int& i, int& owner) {
    owner = lane_id();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float s2 = __shfl_xor_sync(0xffffffffu, s, o);
        int i2 = __shfl_xor_sync(0xffffffffu, i, o);
        int o2 = __shfl_xor_sync(0xffffffffu, owner, o);
        if (better(s2, i2, s, i)) { s = s2; i = i2; owner = o2; }
    }
}

// One warp per query: the n_parts*k candidates are spread over the lanes; k rounds of warp argbest.
// Parts may come in groups (one group per rank of an all-gathered exchange buffer): part p lives at
// (p / parts_per_group) * group_stride + (p % parts_per_group) * nq * k elements from the base pointers.
// RECALL: the same warp also ranks the query's ground truth in the merged list and bumps the per-task Recall@1/5/10 counters
// (sprompt.py:559-619) -- the whole tail of a sharded search step (k-way merge over chunks and ranks + bookkeeping) is one launch.
template <bool RECALL>
__global__ void topk_merge_kernel(const float* __restrict__ ps, const int* __restrict__ pi, int n_parts, int parts_per_group,
                                  long long group_stride, int nq, int k, float* __restrict__ os, int* __restrict__ oi,
                                  const int* __restrict__ gt_ptr, const int* __restrict__ gt_idx, const int* __restrict__ task,
                                  int n_tasks, int* __restrict__ counts, int* __restrict__ rank_out) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    const int lane = lane_id();
    const int total = n_parts * k;
    constexpr int PER = 8;                      // up to 256 candidates per query (e.g. 8 shards x 3 chunks x 10)
    float s[PER];
    int id[PER];
    int g0 = 0, g1 = 0, rank = k;
    if (RECALL) { g0 = gt_ptr[q]; g1 = gt_ptr[q + 1]; }
#pragma unroll
    for (int t = 0; t < PER; ++t) {
        const int c = lane + 32 * t;
        if (c < total) {
            const int part = c / k, j = c - part * k;
            const size_t off = size_t(part / parts_per_group) * group_stride + (size_t(part % parts_per_group) * nq + q) * k + j;
            s[t] = ps[off];
            id[t] = pi[off];
        } else {
            s[t] = -CUDART_INF_F;
            id[t] = 0x7fffffff;
        }
    }
    for (int r = 0; r < k; ++r) {
        float bs = s[0];
        int bi = id[0], bt = 0;
#pragma unroll
        for (int t = 1; t < PER; ++t)
            if (better(s[t], id[t], bs, bi)) { bs = s[t]; bi = id[t]; bt = t; }
        float ws = bs;
        int wi = bi, owner;
        warp_best(ws, wi, owner);
        if (lane == owner) {
#pragma unroll
            for (int t = 0; t < PER; ++t)
                if (t == bt) { s[t] = -CUDART_INF_F; id[t] = 0x7fffffff; }
        }
        if (lane == 0) {
            os[size_t(q) * k + r] = ws;
            oi[size_t(q) * k + r] = wi;
        }
        if (RECALL && rank == k) {              // every lane holds the winner: first position whose index is a ground truth of q
            bool hit = false;
            for (int g = g0 + lane; g < g1; g += 32) hit |= (gt_idx[g] == wi);
            if (__any_sync(0xffffffffu, hit)) rank = r;
        }
    }
    if (RECALL && lane == 0) {
        if (rank_out) rank_out[q] = rank;
        const int t = task ? task[q] : 0;
        if (t >= 0 && t < n_tasks) {
            if (rank < k && rank < 1) atomicAdd(&counts[4 * t + 0], 1);
            if (rank < k && rank < 5) atomicAdd(&counts[4 * t + 1], 1);
            if (rank < k && rank < 10) atomicAdd(&counts[4 * t + 2], 1);
            atomicAdd(&counts[4 * t + 3], 1);
        }
    }
}

// Dense-row top-k: one warp per row; every lane keeps a private sorted top-k (registers/local) over its strided
// slice with a threshold filter, then the 32 lists are merged by k rounds of warp argbest.
template <int KMAX>
__global__ void topk_rows_kernel(const float* __restrict__ sc, int n_rows, int n_cols, long long ld, int k,
                                 float* __restrict__ os, int* __restrict__ oi) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int lane = lane_id();
    const float* p = sc + size_t(row) * ld;
    float ls[KMAX];
    int li[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) { ls[j] = -CUDART_INF_F; li[j] = 0x7fffffff; }
    float thr = -CUDART_INF_F;
    int cnt = 0;
    for (int c = lane; c < n_cols; c += 32) {   // ascending index per lane => strict '>' keeps the lowest index on ties
        const float v = p[c];
        if (v > thr || cnt < k) {
            if (!(v == v)) continue;            // NaN never ranks
            // insert (v, c) keeping (score desc, index asc); fully unrolled so the list stays in registers
            float cs = v;
            int ci = c;
#pragma unroll
            for (int j = 0; j < KMAX; ++j) {
                if (j < k && better(cs, ci, ls[j], li[j])) {
                    float ts = ls[j]; int ti = li[j];
                    ls[j] = cs; li[j] = ci;
                    cs = ts; ci = ti;
                }
            }
            if (cnt < k) ++cnt;
            if (cnt == k) {
#pragma unroll
                for (int j = 0; j < KMAX; ++j)
                    if (j == k - 1) thr = ls[j];
            }
        }
    }
    int head = 0;                                // lists are sorted: only the head of each lane competes
    for (int r = 0; r < k; ++r) {
        float hs = -CUDART_INF_F;
        int hi = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j == head) { hs = ls[j]; hi = li[j]; }
        float ws = hs;
        int wi = hi, owner;
        warp_best(ws, wi, owner);
        if (lane == owner) ++head;
        if (lane == 0) {
            os[size_t(row) * k + r] = ws;
            oi[size_t(row) * k + r] = wi;
        }
    }
}

__global__ void recall_counts_kernel(const int* __restrict__ topk, int nq, int k, const int* __restrict__ gt_ptr,
                                     const int* __restrict__ gt_idx, const int* __restrict__ task, int n_tasks,
                                     int* __restrict__ counts, int* __restrict__ rank_out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int rank = k;
    const int g0 = gt_ptr[q], g1 = gt_ptr[q + 1];
    for (int p = 0; p < k && rank == k; ++p) {
        const int cand = topk[size_t(q) * k + p];
        for (int g = g0; g < g1; ++g)
            if (gt_idx[g] == cand) { rank = p; break; }
    }
    if (rank_out) rank_out[q] = rank;
    const int t = task ? task[q] : 0;
    if (t >= 0 && t < n_tasks) {
        const bool found = rank < k;           // a ground truth outside the kept top-k never counts as a hit
        if (found && rank < 1) atomicAdd(&counts[4 * t + 0], 1);
        if (found && rank < 5) atomicAdd(&counts[4 * t + 1], 1);
        if (found && rank < 10) atomicAdd(&counts[4 * t + 2], 1);
        atomicAdd(&counts[4 * t + 3], 1);
    }
}

// fp32 -> bf16 operand rows; n_terms = 6 lays out the hi/mid/lo split described in lpi_b200.h
__global__ void split_bf16_kernel(const float* __restrict__ x, long long n_elems, int dim, int n_terms, int role,
                                  __nv_bfloat16* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    const long long row = e / dim;
    const int d = int(e - row * dim);
    const float v = x[e];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    if (n_terms == 1) { out[e] = hi; return; }
    const float r1 = v - __bfloat162float(hi);                 // exact in fp32
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(mid);               // exact in fp32
    const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
    __nv_bfloat16* o = out + row * (6LL * dim) + d;
    if (role == 0) { o[0] = hi; o[dim] = hi; o[2 * dim] = mid; o[3 * dim] = mid; o[4 * dim] = hi; o[5 * dim] = lo; }
    else           { o[0] = hi; o[dim] = mid; o[2 * dim] = hi; o[3 * dim] = mid; o[4 * dim] = lo; o[5 * dim] = hi; }
}

// one warp per row
__global__ void l2_normalize_kernel(const float* __restrict__ x, int n, int dim, float* __restrict__ out,
                                    float* __restrict__ norm_out) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int lane = lane_id();
    const float* p = x + size_t(row) * dim;
    float ss = 0.f;
    for (int d = lane; d < dim; d += 32) { float v = p[d]; ss += v * v; }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    if (norm_out && lane == 0) norm_out[row] = nrm;
    for (int d = lane; d < dim; d += 32) out[size_t(row) * dim + d] = p[d] / nrm;
}

}  // namespace lpi

using namespace lpi;

extern "C" int lpi_topk_merge(const float* part_scores, const int* part_idx, int n_parts, int n_queries, int k,
                              float* out_scores, int* out_idx, void* stream) {
    if (n_queries <= 0) return LPI_OK;
    if (k < 1 || n_parts < 1 || long(n_parts) * k > 256)
        return set_error(LPI_ERR_ARG, "topk_merge: n_parts*k=%ld must be in [1,256]", long(n_parts) * k);
    const int threads = 256, wpb = threads / 32;
    topk_merge_kernel<false><<<(n_queries + wpb - 1) / wpb, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        part_scores, part_idx, n_parts, n_parts, 0, n_queries, k, out_scores, out_idx, nullptr, nullptr, nullptr, 0, nullptr, nullptr);
    return check_launch("topk_merge");
}

extern "C" int lpi_topk_merge_recall(const float* part_scores, const int* part_idx, int n_parts, int parts_per_group,
                                     long long group_stride, int n_queries, int k, float* out_scores, int* out_idx,
                                     const int* gt_ptr, const int* gt_idx, const int* task_of_query, int n_tasks, int* counts,
                                     int* rank_out, void* stream) {
    if (n_tasks < 1) return set_error(LPI_ERR_ARG, "topk_merge_recall: n_tasks=%d", n_tasks);
    if (k < 1 || n_parts < 1 || long(n_parts) * k > 256)
        return set_error(LPI_ERR_ARG, "topk_merge_recall: n_parts*k=%ld must be in [1,256]", long(n_parts) * k);
    if (parts_per_group < 1 || n_parts % parts_per_group)
        return set_error(LPI_ERR_ARG, "topk_merge_recall: n_parts=%d is not a multiple of parts_per_group=%d", n_parts, parts_per_group);
    if (!gt_ptr || !gt_idx || !counts) return set_error(LPI_ERR_ARG, "topk_merge_recall: ground truth / counters missing");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(counts, 0, sizeof(int) * 4 * n_tasks, st);
    if (n_queries <= 0) return LPI_OK;
    const int threads = 256, wpb = threads / 32;
    topk_merge_kernel<true><<<(n_queries + wpb - 1) / wpb, threads, 0, st>>>(part_scores, part_idx, n_parts, parts_per_group, group_stride,
                                                                           n_queries, k, out_scores, out_idx, gt_ptr, gt_idx,
                                                                           task_of_query, n_tasks, counts, rank_out);
    return check_launch("topk_merge_recall");
}

extern "C" int lpi_topk_rows_f32(const float* scores, int n_rows, int n_cols, long long ld, int k, float* out_scores,
                                 int* out_idx, void* stream) {
    if (n_rows <= 0) return LPI_OK;
    if (k < 1 || k > 16 || n_cols < 1 || ld < n_cols) return set_error(LPI_ERR_ARG, "topk_rows: bad k=%d n_cols=%d", k, n_cols);
    const int threads = 128, wpb = threads / 32;
    topk_rows_kernel<16><<<(n_rows + wpb - 1) / wpb, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        scores, n_rows, n_cols, ld, k, out_scores, out_idx);
    return check_launch("topk_rows");
}

extern "C" int lpi_recall_counts(const int* topk_idx, int n_queries, int k, const int* gt_ptr, const int* gt_idx,
                                 const int* task_of_query, int n_tasks, int* counts, int* rank_out, void* stream) {
    if (n_tasks < 1) return set_error(LPI_ERR_ARG, "recall_counts: n_tasks=%d", n_tasks);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(counts, 0, sizeof(int) * 4 * n_tasks, st);
    if (n_queries <= 0) return LPI_OK;
    recall_counts_kernel<<<(n_queries + 127) / 128, 128, 0, st>>>(topk_idx, n_queries, k, gt_ptr, gt_idx, task_of_query,
                                                                  n_tasks, counts, rank_out);
    return check_launch("recall_counts");
}

extern "C" int lpi_split_bf16(const float* x, int n, int dim, int n_terms, int role, void* out_bf16, void* stream) {
    if (n <= 0) return LPI_OK;
    if (n_terms != 1 && n_terms != 6) return set_error(LPI_ERR_ARG, "split_bf16: n_terms must be 1 or 6");
    const long long ne = (long long)n * dim;
    split_bf16_kernel<<<unsigned((ne + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ne, dim, n_terms, role, static_cast<__nv_bfloat16*>(out_bf16));
    return check_launch("split_bf16");
}

extern "C" int lpi_l2_normalize(const float* x, int n, int dim, float* out, float* norm_out, void* stream) {
    if (n <= 0) return LPI_OK;
    const int threads = 256, wpb = threads / 32;
    l2_normalize_kernel<<<(n + wpb - 1) / wpb, threads, 0, static_cast<cudaStream_t>(stream)>>>(x, n, dim, out, norm_out);
    return check_launch("l2_normalize");
}
