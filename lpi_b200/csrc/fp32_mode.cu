// fp32 parity mode of the towers (BASELINE.json north_star: "embeddings, logits and loss within ... 1e-5 in fp32").
// The reference run un-converted is plain fp32 (clip.py:128-129; retrieval/models/clip/model.py:154-196): to be comparable at 1e-5 the
// towers need exact fp32 products, which no tensor-core path gives (bf16 / fp16 / TF32 operands round at 2^-9 .. 2^-11, and the
// tensor-core accumulator truncates).  This file holds the SIMT pieces that complete such a path next to lpi_sgemm_bias_f32
// (loss.cu), the fp32 LayerNorm / head kernels (elementwise.cu) and lpi_im2col_patches_f32: attention forward / backward with fp32
// FFMA products and expf, and QuickGELU / its derivative as elementwise kernels.  It is a TEST mode (a few TFLOP/s), selected with
// precision="fp32" on the engines; the throughput path stays on tcgen05.
#include "ptx.cuh"
#include "lpi_internal.h"
#include <math_constants.h>

namespace lpi {

constexpr int HD = 64;                         // head width (both towers)

// One CTA per (sample, head); K and V of the head staged in shared memory; one warp per query row, lanes over the head width
// (2 columns each); online softmax in fp32 with expf.  qkv [B*L, 3*H*64] (q | k | v, heads contiguous), out [B*L, H*64],
// lse [B*H*L] = natural-log sum-exp of the scaled scores (kept for the backward).
__global__ void __launch_bounds__(256)
attn_fwd_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ lse, int B, int L, int H, int causal) {
    extern __shared__ float sm[];              // K[L][64] | V[L][64]
    float* sK = sm;
    float* sV = sm + size_t(L) * HD;
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int D = H * HD, ld = 3 * D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < L * HD; i += blockDim.x) {
        const int j = i / HD, d = i % HD;
        const float* row = qkv + (size_t(b) * L + j) * ld + h * HD + d;
        sK[i] = row[D];
        sV[i] = row[2 * D];
    }
    __syncthreads();
    const float scale = 0.125f;                // 1 / sqrt(64)
    for (int i = warp; i < L; i += nw) {
        const float* q = qkv + (size_t(b) * L + i) * ld + h * HD;
        const float q0 = q[lane] * scale, q1 = q[lane + 32] * scale;      // nn.MultiheadAttention scales q before the product
        float m = -CUDART_INF_F, l = 0.f, o0 = 0.f, o1 = 0.f;
        const int jend = causal ? i + 1 : L;
        for (int j = 0; j < jend; ++j) {
            const float s = warp_sum(q0 * sK[j * HD + lane] + q1 * sK[j * HD + lane + 32]);
            const float mn = fmaxf(m, s);
            const float corr = expf(m - mn), p = expf(s - mn);
            l = l * corr + p;
            o0 = o0 * corr + p * sV[j * HD + lane];
            o1 = o1 * corr + p * sV[j * HD + lane + 32];
            m = mn;
        }
        const float inv = 1.f / l;
        float* o = out + (size_t(b) * L + i) * D + h * HD;
        o[lane] = o0 * inv;
        o[lane + 32] = o1 * inv;
        if (lane == 0) lse[(size_t(b) * H + h) * L + i] = m + logf(l);
    }
}

// Backward, query side: dq[i] = sum_j dS[i,j] k[j] / 8 with dS = P * (dO v^T - delta), delta[i] = dO[i] . O[i]  (K, V staged).
__global__ void __launch_bounds__(256)
attn_bwd_dq_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ out, const float* __restrict__ dout,
                       const float* __restrict__ lse, float* __restrict__ delta, float* __restrict__ dqkv, int B, int L, int H, int causal) {
    extern __shared__ float sm[];
    float* sK = sm;
    float* sV = sm + size_t(L) * HD;
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int D = H * HD, ld = 3 * D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < L * HD; i += blockDim.x) {
        const int j = i / HD, d = i % HD;
        const float* row = qkv + (size_t(b) * L + j) * ld + h * HD + d;
        sK[i] = row[D];
        sV[i] = row[2 * D];
    }
    __syncthreads();
    const float scale = 0.125f;
    for (int i = warp; i < L; i += nw) {
        const size_t tok = size_t(b) * L + i;
        const float* q = qkv + tok * ld + h * HD;
        const float q0 = q[lane] * scale, q1 = q[lane + 32] * scale;
        const float do0 = dout[tok * D + h * HD + lane], do1 = dout[tok * D + h * HD + lane + 32];
        const float dl = warp_sum(do0 * out[tok * D + h * HD + lane] + do1 * out[tok * D + h * HD + lane + 32]);
        const float ls = lse[(size_t(b) * H + h) * L + i];
        if (lane == 0) delta[(size_t(b) * H + h) * L + i] = dl;
        float a0 = 0.f, a1 = 0.f;
        const int jend = causal ? i + 1 : L;
        for (int j = 0; j < jend; ++j) {
            const float s = warp_sum(q0 * sK[j * HD + lane] + q1 * sK[j * HD + lane + 32]);
            const float dp = warp_sum(do0 * sV[j * HD + lane] + do1 * sV[j * HD + lane + 32]);
            const float ds = expf(s - ls) * (dp - dl);
            a0 = fmaf(ds, sK[j * HD + lane], a0);
            a1 = fmaf(ds, sK[j * HD + lane + 32], a1);
        }
        dqkv[tok * ld + h * HD + lane] = a0 * scale;
        dqkv[tok * ld + h * HD + lane + 32] = a1 * scale;
    }
}

// Backward, key / value side: one warp per key row j; Q (pre-scaled) and dO of the head staged.
//   dv[j] = sum_i P[i,j] dO[i],   dk[j] = sum_i dS[i,j] q[i] / 8
__global__ void __launch_bounds__(256)
attn_bwd_dkv_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, const float* __restrict__ lse,
                        const float* __restrict__ delta, float* __restrict__ dqkv, int B, int L, int H, int causal) {
    extern __shared__ float sm[];
    float* sQ = sm;                             // q / 8
    float* sO = sm + size_t(L) * HD;            // dO
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int D = H * HD, ld = 3 * D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < L * HD; i += blockDim.x) {
        const int r = i / HD, d = i % HD;
        sQ[i] = qkv[(size_t(b) * L + r) * ld + h * HD + d] * 0.125f;
        sO[i] = dout[(size_t(b) * L + r) * D + h * HD + d];
    }
    __syncthreads();
    const float* lse_h = lse + (size_t(b) * H + h) * L;
    const float* del_h = delta + (size_t(b) * H + h) * L;
    for (int j = warp; j < L; j += nw) {
        const size_t tok = size_t(b) * L + j;
        const float k0 = qkv[tok * ld + D + h * HD + lane], k1 = qkv[tok * ld + D + h * HD + lane + 32];
        const float v0 = qkv[tok * ld + 2 * D + h * HD + lane], v1 = qkv[tok * ld + 2 * D + h * HD + lane + 32];
        float dk0 = 0.f, dk1 = 0.f, dv0 = 0.f, dv1 = 0.f;
        for (int i = causal ? j : 0; i < L; ++i) {
            const float s = warp_sum(sQ[i * HD + lane] * k0 + sQ[i * HD + lane + 32] * k1);
            const float dp = warp_sum(sO[i * HD + lane] * v0 + sO[i * HD + lane + 32] * v1);
            const float p = expf(s - lse_h[i]);
            const float ds = p * (dp - del_h[i]);
            dv0 = fmaf(p, sO[i * HD + lane], dv0);
            dv1 = fmaf(p, sO[i * HD + lane + 32], dv1);
            dk0 = fmaf(ds, sQ[i * HD + lane], dk0);             // sQ already carries the 1/8
            dk1 = fmaf(ds, sQ[i * HD + lane + 32], dk1);
        }
        dqkv[tok * ld + D + h * HD + lane] = dk0;
        dqkv[tok * ld + D + h * HD + lane + 32] = dk1;
        dqkv[tok * ld + 2 * D + h * HD + lane] = dv0;
        dqkv[tok * ld + 2 * D + h * HD + lane + 32] = dv1;
    }
}

// QuickGELU z * sigmoid(1.702 z) (model.py:163-165) and dy * its derivative, fp32 with expf and IEEE division
__global__ void quick_gelu_f32_kernel(const float* __restrict__ z, float* __restrict__ out, long n) {
    const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) { const float x = z[i]; out[i] = x * (1.f / (1.f + expf(-1.702f * x))); }
}
__global__ void quick_gelu_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ z, float* __restrict__ out, long n) {
    const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        const float x = z[i];
        const float s = 1.f / (1.f + expf(-1.702f * x));
        out[i] = dy[i] * (s * (1.f + 1.702f * x * (1.f - s)));
    }
}

static int attn_smem(int L) { return int(2 * size_t(L) * HD * sizeof(float)); }

template <typename K>
static int set_smem(K kern, int bytes, const char* what) {
    if (bytes > 220 * 1024) return set_error(LPI_ERR_UNSUPPORTED, "%s: sequence too long for the fp32 parity kernels (%d B of shared memory)", what, bytes);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return set_error(LPI_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return LPI_OK;
}

}  // namespace lpi

using namespace lpi;

extern "C" int lpi_attn_fwd_f32(const float* qkv, float* out, float* lse, int B, int L, int H, int causal, void* stream) {
    if (B <= 0) return LPI_OK;
    if (L < 1 || H < 1) return set_error(LPI_ERR_ARG, "attn_fwd_f32: bad shape L=%d H=%d", L, H);
    if (int rc = set_smem(attn_fwd_f32_kernel, attn_smem(L), "attn_fwd_f32")) return rc;
    attn_fwd_f32_kernel<<<B * H, 256, attn_smem(L), static_cast<cudaStream_t>(stream)>>>(qkv, out, lse, B, L, H, causal);
    return check_launch("attn_fwd_f32");
}

extern "C" int lpi_attn_bwd_f32(const float* qkv, const float* out, const float* d_out, const float* lse, float* delta_ws, float* dqkv,
                                int B, int L, int H, int causal, void* stream) {
    if (B <= 0) return LPI_OK;
    if (L < 1 || H < 1) return set_error(LPI_ERR_ARG, "attn_bwd_f32: bad shape L=%d H=%d", L, H);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (int rc = set_smem(attn_bwd_dq_f32_kernel, attn_smem(L), "attn_bwd_f32")) return rc;
    if (int rc = set_smem(attn_bwd_dkv_f32_kernel, attn_smem(L), "attn_bwd_f32")) return rc;
    attn_bwd_dq_f32_kernel<<<B * H, 256, attn_smem(L), st>>>(qkv, out, d_out, lse, delta_ws, dqkv, B, L, H, causal);
    attn_bwd_dkv_f32_kernel<<<B * H, 256, attn_smem(L), st>>>(qkv, d_out, lse, delta_ws, dqkv, B, L, H, causal);
    return check_launch("attn_bwd_f32");
}

extern "C" int lpi_quick_gelu_f32(const float* z, float* out, long long n, void* stream) {
    if (n <= 0) return LPI_OK;
    quick_gelu_f32_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(z, out, n);
    return check_launch("quick_gelu_f32");
}

extern "C" int lpi_quick_gelu_bwd_f32(const float* dy, const float* z, float* out, long long n, void* stream) {
    if (n <= 0) return LPI_OK;
    quick_gelu_bwd_f32_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, z, out, n);
    return check_launch("quick_gelu_bwd_f32");
}
