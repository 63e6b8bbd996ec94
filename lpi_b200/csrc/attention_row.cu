// Attention of the LAST block of a tower, restricted to the one query row per sample that the head reads.
//
// VisionTransformer.forward ends in ln_post(x[:, 0, :]) (retrieval/models/clip/model.py:254-257) and TextEncoder.forward in
// x[arange(B), tokenized.argmax(-1)] (prompt_learner.py:57-61): of the last ResidualAttentionBlock's output (model.py:187-196) only
// ONE row per sample is ever read, forward and backward.  That row's attention needs the keys and values of every position but only
// its own query, and everything after the attention in that block (out_proj, ln_2, c_fc, QuickGELU, c_proj) is row-wise -- so the
// engine runs those on [B, D] instead of [B*L, D] (engine.Tower, `out_rows`).  Results are those of the full block at the read rows;
// the rows that are skipped never reach the features, the loss or any gradient.
//
// One CTA per (sample, head), 256 threads, thread = key position for the score / dP passes (a 128-byte K or V row per thread),
// thread = (dim, key group) for the P.V and dS.K sums.  fp32 arithmetic on 16-bit operands; one launch reads K and V once (42 MB at
// B = 64, L = 213): HBM / latency bound, 13-18 us forward and 20-31 us backward (ncu, cold caches) where the full-sequence kernels take
// 43 / 107, and -- the larger part of the 0.49 ms per step this saves -- out_proj, ln_2 and the MLP of that block run on 64 rows.
#include "ptx.cuh"
#include "lpi_internal.h"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>

namespace lpi {

namespace {

constexpr int HD = 64;
constexpr int ROW_THREADS = 256;
constexpr int MAX_L = 512;                    // keys per thread = MAX_L / ROW_THREADS = 2

template <typename T> __device__ __forceinline__ float2 to_f2(uint32_t u);
template <> __device__ __forceinline__ float2 to_f2<__half>(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
template <> __device__ __forceinline__ float2 to_f2<__nv_bfloat16>(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
template <typename T> __device__ __forceinline__ uint32_t from_f2(float a, float b);
template <> __device__ __forceinline__ uint32_t from_f2<__half>(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
template <> __device__ __forceinline__ uint32_t from_f2<__nv_bfloat16>(float a, float b) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

// dot of a 64-wide 16-bit row in global memory with a 64-float vector in shared memory
template <typename T> __device__ __forceinline__ float dot64(const T* __restrict__ row, const float* __restrict__ v) {
    const uint4* p = reinterpret_cast<const uint4*>(row);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 u = __ldg(p + i);
        const float2 a = to_f2<T>(u.x), b = to_f2<T>(u.y), c = to_f2<T>(u.z), d = to_f2<T>(u.w);
        const float* w = v + i * 8;
        acc += a.x * w[0] + a.y * w[1] + b.x * w[2] + b.y * w[3] + c.x * w[4] + c.y * w[5] + d.x * w[6] + d.y * w[7];
    }
    return acc;
}

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, t) : v + t;
    }
    __syncthreads();                          // red[] may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < ROW_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    return r;
}

// Probabilities of the one query row: sP[k] = softmax_k(q . K[k] / 8), k < nk (sP[k] = 0 for nk <= k < L).  sQ holds the raw query.
template <typename T>
__device__ __forceinline__ void row_softmax(const T* __restrict__ kbase, size_t ld, int L, int nk, const float* sQ, float* sP, float* red) {
    float s[MAX_L / ROW_THREADS];
    float m = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < MAX_L / ROW_THREADS; ++i) {
        const int k = threadIdx.x + i * ROW_THREADS;
        s[i] = -CUDART_INF_F;
        if (k < nk) s[i] = dot64<T>(kbase + size_t(k) * ld, sQ) * 0.125f;
        m = fmaxf(m, s[i]);
    }
    m = block_reduce(m, red, true);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_L / ROW_THREADS; ++i) {
        const int k = threadIdx.x + i * ROW_THREADS;
        s[i] = k < nk ? __expf(s[i] - m) : 0.f;
        sum += s[i];
    }
    sum = block_reduce(sum, red, false);
    const float inv = 1.f / sum;
#pragma unroll
    for (int i = 0; i < MAX_L / ROW_THREADS; ++i) {
        const int k = threadIdx.x + i * ROW_THREADS;
        if (k < L) sP[k] = s[i] * inv;
    }
    __syncthreads();
}

// acc[d] = sum_k w[k] * X[k][d] over k < nk; thread = (pair of dims, key group of 8); result left in sOut[64] after a smem reduction
template <typename T>
__device__ __forceinline__ void weighted_row_sum(const T* __restrict__ xbase, size_t ld, int nk, const float* w, float* part, float* sOut) {
    const int dp = threadIdx.x & 31, kg = threadIdx.x >> 5;      // 32 dim pairs x 8 key groups: a warp reads one 128-byte row
    float a0 = 0.f, a1 = 0.f;
    for (int k = kg; k < nk; k += ROW_THREADS / 32) {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(xbase + size_t(k) * ld) + dp);
        const float2 x = to_f2<T>(u);
        a0 += w[k] * x.x;
        a1 += w[k] * x.y;
    }
    part[kg * HD + 2 * dp] = a0;
    part[kg * HD + 2 * dp + 1] = a1;
    __syncthreads();
    if (threadIdx.x < HD) {
        float r = 0.f;
#pragma unroll
        for (int g = 0; g < ROW_THREADS / 32; ++g) r += part[g * HD + threadIdx.x];
        sOut[threadIdx.x] = r;
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(ROW_THREADS)
attn_rowq_fwd_kernel(const T* __restrict__ qkv, const int* __restrict__ rows, T* __restrict__ out_rows, const float* __restrict__ x,
                     float* __restrict__ x_rows, int B, int L, int H, int causal) {
    __shared__ float sQ[HD], sO[HD], sP[MAX_L], part[(ROW_THREADS / 32) * HD], red[ROW_THREADS / 32];
    pdl_enter();
    const int h = blockIdx.x, b = blockIdx.y;
    const int D = H * HD;
    const size_t ld = size_t(3) * D;
    const int row = rows[b];                                      // global row index in [B*L]
    const int pos = row - b * L;
    const int nk = causal ? pos + 1 : L;
    const T* base = qkv + size_t(b) * L * ld + h * HD;
    if (threadIdx.x < HD) {
        sQ[threadIdx.x] = float(qkv[size_t(row) * ld + h * HD + threadIdx.x]);
        if (x_rows) x_rows[size_t(b) * D + h * HD + threadIdx.x] = x[size_t(row) * D + h * HD + threadIdx.x];
    }
    __syncthreads();
    row_softmax<T>(base + D, ld, L, nk, sQ, sP, red);
    weighted_row_sum<T>(base + 2 * D, ld, nk, sP, part, sO);
    if (threadIdx.x < HD / 2)
        reinterpret_cast<uint32_t*>(out_rows + size_t(b) * D + h * HD)[threadIdx.x] = from_f2<T>(sO[2 * threadIdx.x], sO[2 * threadIdx.x + 1]);
}

// dqkv [B*L, 3D] is written COMPLETELY for this (sample, head): dq is zero except at the query row, dk / dv are zero beyond the causal
// horizon.  g_rows (optional): the residual-path gradient of the read rows is scattered into the full fp32 gradient stream g
// (which the caller zeroed) so that the LayerNorm backward of the block accumulates onto it.
template <typename T>
__global__ void __launch_bounds__(ROW_THREADS)
attn_rowq_bwd_kernel(const T* __restrict__ qkv, const int* __restrict__ rows, const T* __restrict__ dout_rows, T* __restrict__ dqkv,
                     const float* __restrict__ g_rows, float* __restrict__ g, int B, int L, int H, int causal) {
    __shared__ float sQ[HD], sDO[HD], sDQ[HD], sP[MAX_L], sDS[MAX_L], part[(ROW_THREADS / 32) * HD], red[ROW_THREADS / 32];
    pdl_enter();
    const int h = blockIdx.x, b = blockIdx.y;
    const int D = H * HD;
    const size_t ld = size_t(3) * D;
    const int row = rows[b];
    const int pos = row - b * L;
    const int nk = causal ? pos + 1 : L;
    const T* base = qkv + size_t(b) * L * ld + h * HD;
    T* dbase = dqkv + size_t(b) * L * ld + h * HD;
    if (threadIdx.x < HD) {
        sQ[threadIdx.x] = float(qkv[size_t(row) * ld + h * HD + threadIdx.x]);
        sDO[threadIdx.x] = float(dout_rows[size_t(b) * D + h * HD + threadIdx.x]);
        if (g) g[size_t(row) * D + h * HD + threadIdx.x] = g_rows[size_t(b) * D + h * HD + threadIdx.x];
    }
    __syncthreads();
    row_softmax<T>(base + D, ld, L, nk, sQ, sP, red);
    // dP[k] = dO . V[k];  delta = sum_k P[k] dP[k];  dS[k] = P[k] (dP[k] - delta)   (the 1/8 goes to dq and dk)
    float dp[MAX_L / ROW_THREADS];
    float dl = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_L / ROW_THREADS; ++i) {
        const int k = threadIdx.x + i * ROW_THREADS;
        dp[i] = k < nk ? dot64<T>(base + 2 * D + size_t(k) * ld, sDO) : 0.f;
        if (k < nk) dl += sP[k] * dp[i];
    }
    dl = block_reduce(dl, red, false);
#pragma unroll
    for (int i = 0; i < MAX_L / ROW_THREADS; ++i) {
        const int k = threadIdx.x + i * ROW_THREADS;
        if (k < L) sDS[k] = k < nk ? sP[k] * (dp[i] - dl) * 0.125f : 0.f;
    }
    __syncthreads();
    weighted_row_sum<T>(base + D, ld, nk, sDS, part, sDQ);         // dq = sum_k dS[k] K[k] / 8
    // rows of dq | dk | dv: a warp writes one 128-byte row segment per step
    const int dpair = threadIdx.x & 31, kg = threadIdx.x >> 5;
    const float q0 = sQ[2 * dpair], q1 = sQ[2 * dpair + 1], o0 = sDO[2 * dpair], o1 = sDO[2 * dpair + 1];
    for (int k = kg; k < L; k += ROW_THREADS / 32) {
        uint32_t* r = reinterpret_cast<uint32_t*>(dbase + size_t(k) * ld);
        const float ds = sDS[k], p = sP[k];                        // both zero beyond the causal horizon
        r[dpair] = k == pos ? from_f2<T>(sDQ[2 * dpair], sDQ[2 * dpair + 1]) : 0u;
        r[D / 2 + dpair] = from_f2<T>(ds * q0, ds * q1);
        r[D + dpair] = from_f2<T>(p * o0, p * o1);
    }
}

template <typename T>
int rowq_fwd(const void* qkv, const int* rows, void* out_rows, const float* x, float* x_rows, int B, int L, int H, int causal, cudaStream_t st) {
    launch_pdl(attn_rowq_fwd_kernel<T>, dim3(H, B), dim3(ROW_THREADS), 0, st, static_cast<const T*>(qkv), rows, static_cast<T*>(out_rows), x, x_rows, B, L, H, causal);
    return check_launch("attn_rowq_fwd");
}

template <typename T>
int rowq_bwd(const void* qkv, const int* rows, const void* dout_rows, void* dqkv, const float* g_rows, float* g, int B, int L, int H,
             int causal, cudaStream_t st) {
    launch_pdl(attn_rowq_bwd_kernel<T>, dim3(H, B), dim3(ROW_THREADS), 0, st, static_cast<const T*>(qkv), rows, static_cast<const T*>(dout_rows),
               static_cast<T*>(dqkv), g_rows, g, B, L, H, causal);
    return check_launch("attn_rowq_bwd");
}

int check_args(const char* what, const void* qkv, const int* rows, int B, int L, int H) {
    if (!qkv || !rows) return set_error(LPI_ERR_ARG, "%s: null argument", what);
    if (B <= 0 || H <= 0 || L <= 0 || L > MAX_L) return set_error(LPI_ERR_ARG, "%s: need B, H > 0 and 0 < L <= %d (B=%d L=%d H=%d)", what, MAX_L, B, L, H);
    if (B > 65535) return set_error(LPI_ERR_ARG, "%s: B=%d exceeds the grid limit 65535", what, B);
    return LPI_OK;
}

}  // namespace

}  // namespace lpi

extern "C" int lpi_attn_rowq_fwd(const void* qkv, const int* rows, void* out_rows, const float* x, float* x_rows, int B, int L, int H,
                                 int causal, int f16, void* stream) {
    using namespace lpi;
    if (int rc = check_args("attn_rowq_fwd", qkv, rows, B, L, H)) return rc;
    if (!out_rows || ((x == nullptr) != (x_rows == nullptr))) return set_error(LPI_ERR_ARG, "attn_rowq_fwd: out_rows missing, or x / x_rows not given together");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return f16 ? rowq_fwd<__half>(qkv, rows, out_rows, x, x_rows, B, L, H, causal, st)
               : rowq_fwd<__nv_bfloat16>(qkv, rows, out_rows, x, x_rows, B, L, H, causal, st);
}

extern "C" int lpi_attn_rowq_bwd(const void* qkv, const int* rows, const void* dout_rows, void* dqkv, const float* g_rows, float* g, int B,
                                 int L, int H, int causal, int f16, void* stream) {
    using namespace lpi;
    if (int rc = check_args("attn_rowq_bwd", qkv, rows, B, L, H)) return rc;
    if (!dout_rows || !dqkv || ((g == nullptr) != (g_rows == nullptr)))
        return set_error(LPI_ERR_ARG, "attn_rowq_bwd: dout_rows / dqkv missing, or g / g_rows not given together");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return f16 ? rowq_bwd<__half>(qkv, rows, dout_rows, dqkv, g_rows, g, B, L, H, causal, st)
               : rowq_bwd<__nv_bfloat16>(qkv, rows, dout_rows, dqkv, g_rows, g, B, L, H, causal, st);
}
