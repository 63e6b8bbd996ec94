// Bandwidth-bound kernels of the prompted CLIP towers: LayerNorm fwd/bwd, patch extraction, token assembly
// (CLS / prompt rows / patches, + positional embedding, + ln_pre), text embedding gather + prompt splice,
// and the encoder heads (ln_post / ln_final -> projection -> L2 normalise) with their backward.
// Reference: retrieval/models/clip/model.py:154-160 (LayerNorm in fp32), :227-259 (VisionTransformer.forward),
// retrieval/models/clip/prompt_learner.py:52-63,128-163 (TextEncoder / PromptLearner), retrieval/models/slinet.py:122,133.
// All rows are [tokens, D] fp32 (residual stream) or bf16 (GEMM operands); one warp per row, 16-byte accesses.
#include <cuda_fp16.h>
#include <math_constants.h>
#include "ptx.cuh"
#include "lpi_internal.h"

namespace lpi {


struct RowStats { float mean, rstd; };

// mean / rstd of one row held as v[n] per lane (n = D/32 values, strided by 32 float4 groups)
template <int NV>   // NV float4 per lane
__device__ __forceinline__ RowStats row_stats(const float4 (&v)[NV], int D, float eps) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += a * a + b * b + c * c + d * d;
    }
    const float var = warp_sum(q) / D;       // biased variance, two-pass (as torch native_layer_norm)
    return {mean, rsqrtf(var + eps)};
}

// 16-bit shadows are bf16 (vision tower) or fp16 (F16: text tower, see gemm.cu)
template <bool F16>
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    if (F16) {
        const __half2 v = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<const uint32_t*>(&v);
    }
    return pack_bf16x2(lo, hi);
}

template <int NV, bool F16 = false>
__device__ __forceinline__ void ln_apply_store(const float4 (&v)[NV], RowStats st, const float* gamma, const float* beta, int lane,
                                               float* out_f32, __nv_bfloat16* out_bf16) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
        float4 y;
        y.x = (v[i].x - st.mean) * st.rstd * g.x + b.x;
        y.y = (v[i].y - st.mean) * st.rstd * g.y + b.y;
        y.z = (v[i].z - st.mean) * st.rstd * g.z + b.z;
        y.w = (v[i].w - st.mean) * st.rstd * g.w + b.w;
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + c) = y;
        if (out_bf16) *reinterpret_cast<uint2*>(out_bf16 + c) = make_uint2(pack_h2<F16>(y.x, y.y), pack_h2<F16>(y.z, y.w));
    }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
template <int NV, bool F16 = false>
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, long M, int D, float eps) {
    pdl_enter();
    const long row = (long(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(x + row * D + (i * 32 + lane) * 4);
    const RowStats st = row_stats<NV>(v, D, eps);
    ln_apply_store<NV, F16>(v, st, gamma, beta, lane, out_f32 ? out_f32 + row * D : nullptr, out_bf16 ? out_bf16 + row * D : nullptr);
}

// dx = rstd * (gdy - mean(gdy) - xhat * mean(gdy * xhat)),  gdy = gamma * dy;   g = (accumulate ? g : 0) + dx
template <int NV>
__device__ __forceinline__ void ln_bwd_row(const float4 (&xv)[NV], const float4 (&dyv)[NV], const float* gamma, int D, float eps, int lane,
                                           float4 (&dx)[NV]) {
    const RowStats st = row_stats<NV>(xv, D, eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4);
        float4 a;
        a.x = g.x * dyv[i].x; a.y = g.y * dyv[i].y; a.z = g.z * dyv[i].z; a.w = g.w * dyv[i].w;
        float4 h;
        h.x = (xv[i].x - st.mean) * st.rstd; h.y = (xv[i].y - st.mean) * st.rstd;
        h.z = (xv[i].z - st.mean) * st.rstd; h.w = (xv[i].w - st.mean) * st.rstd;
        s1 += a.x + a.y + a.z + a.w;
        s2 += a.x * h.x + a.y * h.y + a.z * h.z + a.w * h.w;
        dx[i] = a;
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        dx[i].x = st.rstd * (dx[i].x - s1 - (xv[i].x - st.mean) * st.rstd * s2);
        dx[i].y = st.rstd * (dx[i].y - s1 - (xv[i].y - st.mean) * st.rstd * s2);
        dx[i].z = st.rstd * (dx[i].z - s1 - (xv[i].z - st.mean) * st.rstd * s2);
        dx[i].w = st.rstd * (dx[i].w - s1 - (xv[i].w - st.mean) * st.rstd * s2);
    }
}

// F16: the fp16 gradient path carries gradients multiplied by `shadow_scale` (a power of two) so small values stay in fp16's
// normal range: dy arrives scaled (dy_scale = 1 / shadow_scale brings it back), g stays true-scale fp32, the shadow is scaled again.
// DY16: dy arrives in the tower's 16-bit type (bf16, or fp16 when F16) straight from the dgrad GEMM's epilogue -- half the bytes of
// the largest stream this HBM-bound kernel reads.
template <int NV, bool F16 = false, bool DY16 = false>
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                                     float* __restrict__ g, __nv_bfloat16* __restrict__ g_bf16, long M, int D, float eps, int accumulate,
                                     float dy_scale = 1.0f, float shadow_scale = 1.0f) {
    pdl_enter();
    const long row = (long(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    float4 xv[NV], dyv[NV], dx[NV], gv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {                   // every load of the row is issued before the first reduction: ONE exposed memory latency
        const long o = row * D + (i * 32 + lane) * 4;
        xv[i] = *reinterpret_cast<const float4*>(x + o);
        if (DY16) {
            const uint2 w = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + o);
            const float2 a = F16 ? __half22float2(*reinterpret_cast<const __half2*>(&w.x)) : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.x));
            const float2 b = F16 ? __half22float2(*reinterpret_cast<const __half2*>(&w.y)) : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.y));
            dyv[i] = make_float4(a.x, a.y, b.x, b.y);
        } else {
            dyv[i] = *reinterpret_cast<const float4*>(dy + o);
        }
        gv[i] = accumulate ? *reinterpret_cast<const float4*>(g + o) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (F16) {
#pragma unroll
        for (int i = 0; i < NV; ++i) { dyv[i].x *= dy_scale; dyv[i].y *= dy_scale; dyv[i].z *= dy_scale; dyv[i].w *= dy_scale; }
    }
    ln_bwd_row<NV>(xv, dyv, gamma, D, eps, lane, dx);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const long o = row * D + (i * 32 + lane) * 4;
        float4 r = dx[i];
        r.x += gv[i].x; r.y += gv[i].y; r.z += gv[i].z; r.w += gv[i].w;
        *reinterpret_cast<float4*>(g + o) = r;
        if (g_bf16) {
            if (F16) { r.x *= shadow_scale; r.y *= shadow_scale; r.z *= shadow_scale; r.w *= shadow_scale; }
            *reinterpret_cast<uint2*>(g_bf16 + o) = make_uint2(pack_h2<F16>(r.x, r.y), pack_h2<F16>(r.z, r.w));
        }
    }
}

// ------------------------------------------------------------------------------------------------ vision front end
// images [B,3,R,R] fp32 -> patch rows [B*G*G, 3*P*P] bf16, column = c*P*P + i*P + j (conv1.weight.view(D,-1) order)
// (fp32 parity mode: the same gather without the 16-bit rounding)
__global__ void im2col_f32_kernel(const float* __restrict__ img, float* __restrict__ out, int B, int R, int P) {
    const int G = R / P, K = 3 * P * P;
    const long n = long(B) * G * G * K / 4;
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long e = t * 4;
    const int col = int(e % K);
    const long prow = e / K;
    const int gx = int(prow % G), gy = int((prow / G) % G), b = int(prow / (G * G));
    const int c = col / (P * P), i = (col / P) % P, j = col % P;
    *reinterpret_cast<float4*>(out + e) = *reinterpret_cast<const float4*>(img + ((long(b) * 3 + c) * R + gy * P + i) * R + gx * P + j);
}

template <bool F16>
__global__ void im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int R, int P) {
    const int G = R / P, K = 3 * P * P;
    const long n = long(B) * G * G * K / 4;          // 4 consecutive j per thread (P % 4 == 0)
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long e = t * 4;
    const int col = int(e % K);
    const long prow = e / K;
    const int gx = int(prow % G), gy = int((prow / G) % G), b = int(prow / (G * G));
    const int c = col / (P * P), i = (col / P) % P, j = col % P;
    const float4 v = *reinterpret_cast<const float4*>(img + ((long(b) * 3 + c) * R + gy * P + i) * R + gx * P + j);
    *reinterpret_cast<uint2*>(out + e) = make_uint2(pack_h2<F16>(v.x, v.y), pack_h2<F16>(v.z, v.w));
}

// DecomposedPrompt row straight from the factors (prompts.py:38-57, layer 0): y[c] = scale * (1/r) sum_k (a[k] b[k]) c3[c, k] -- the same
// association and order as prompt_fwd_kernel, so the fused assembly is bit-identical to reconstruct-then-assemble.
struct PromptFactors {
    const float* d1;      // [T, Lp, r] layer factor (dim_1_share) of every selectable task; layer 0 enters the token sequence
    const float* d2;      // [T, P, r]  prompt factor of this modality
    const float* d3;      // [T, D, r]  width factor of this modality
    int r, Lp;
    float scale;          // DecomposedPrompt.scale (1 in the reference)
};
__device__ __forceinline__ float prompt_value(const PromptFactors& f, int t, int p, int P, int D, int c) {
    const float* a = f.d1 + long(t) * f.Lp * f.r;
    const float* b = f.d2 + (long(t) * P + p) * f.r;
    const float* c3 = f.d3 + (long(t) * D + c) * f.r;
    float s = 0.f;
    for (int k = 0; k < f.r; ++k) s = fmaf(a[k] * b[k], c3[k], s);
    s *= 1.f / f.r;
    return f.scale == 1.f ? s : s * f.scale;
}

// Token assembly + ln_pre (model.py:235-250).  Row l of sample b:
//   l = 0          : class_embedding + pos[0]
//   1 <= l <= P    : prompt_table[sel[b]][l-1]              (NO positional term)
//   l > P          : patch_emb[b, l-1-P] + pos[l-P]
//   (prompt rows: from the materialised table, or -- fac.d1 != nullptr -- reconstructed here from the tri-factor decomposition)
template <int NV>
__global__ void assemble_vision_kernel(const float* __restrict__ patch_emb, const float* __restrict__ cls, const float* __restrict__ pos,
                                       const float* __restrict__ prompt_table, const int* __restrict__ sel, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float* __restrict__ x_out, int B, int n_patch, int P, int D, float eps,
                                       PromptFactors fac) {
    const int L = 1 + P + n_patch;
    const long row = (long(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= long(B) * L) return;
    const int lane = threadIdx.x & 31;
    const int b = int(row / L), l = int(row % L);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (l == 0) {
            const float4 a = *reinterpret_cast<const float4*>(cls + c), p = *reinterpret_cast<const float4*>(pos + c);
            v[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
        } else if (l <= P) {
            const int t = sel ? sel[b] : 0;
            if (fac.d1) v[i] = make_float4(prompt_value(fac, t, l - 1, P, D, c), prompt_value(fac, t, l - 1, P, D, c + 1),
                                           prompt_value(fac, t, l - 1, P, D, c + 2), prompt_value(fac, t, l - 1, P, D, c + 3));
            else v[i] = *reinterpret_cast<const float4*>(prompt_table + (long(t) * P + (l - 1)) * D + c);
        } else {
            const float4 a = *reinterpret_cast<const float4*>(patch_emb + (long(b) * n_patch + (l - 1 - P)) * D + c);
            const float4 p = *reinterpret_cast<const float4*>(pos + long(l - P) * D + c);
            v[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
        }
    }
    const RowStats st = row_stats<NV>(v, D, eps);
    ln_apply_store<NV>(v, st, gamma, beta, lane, x_out + row * D, nullptr);
}

// d prompt_table[t, p, :] = sum over {b : sel[b] == t} of LNbwd(g[b, 1+p, :]; x = prompt_table[t, p, :])
// grid = (P, T); each warp walks a strided slice of the batch, block-level reduction in shared memory (deterministic).
template <int NV>
__global__ void __launch_bounds__(512)
assemble_vision_bwd_kernel(const float* __restrict__ g, const float* __restrict__ prompt_table, const int* __restrict__ sel,
                                           const float* __restrict__ gamma, float* __restrict__ d_prompt, int B, int L, int P, int D, float eps,
                                           PromptFactors fac) {
    extern __shared__ float red[];                   // [warps][D]
    const int p = blockIdx.x, t = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float4 xv[NV], acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (fac.d1) xv[i] = make_float4(prompt_value(fac, t, p, P, D, c), prompt_value(fac, t, p, P, D, c + 1), prompt_value(fac, t, p, P, D, c + 2),
                                        prompt_value(fac, t, p, P, D, c + 3));
        else xv[i] = *reinterpret_cast<const float4*>(prompt_table + (long(t) * P + p) * D + c);
        acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int b = warp; b < B; b += nw) {
        if (sel && sel[b] != t) continue;
        if (!sel && t != 0) continue;
        float4 dyv[NV], dx[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) dyv[i] = *reinterpret_cast<const float4*>(g + (long(b) * L + 1 + p) * D + (i * 32 + lane) * 4);
        ln_bwd_row<NV>(xv, dyv, gamma, D, eps, lane, dx);
#pragma unroll
        for (int i = 0; i < NV; ++i) { acc[i].x += dx[i].x; acc[i].y += dx[i].y; acc[i].z += dx[i].z; acc[i].w += dx[i].w; }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(red + warp * D + (i * 32 + lane) * 4) = acc[i];
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nw; ++w) s += red[w * D + c];
        d_prompt[(long(t) * P + p) * D + c] = s;
    }
}

// ------------------------------------------------------------------------------------------------ text front end
// x[b, l] = (l in [1, P] and ctx given ? ctx_table[sel[b]][l-1] : token_embedding[tok[b, l]]) + pos[l]
__global__ void assemble_text_kernel(const float* __restrict__ emb, const long long* __restrict__ tok, const float* __restrict__ pos,
                                     const float* __restrict__ ctx_table, const int* __restrict__ sel, float* __restrict__ x_out, int B, int L,
                                     int P, int D, PromptFactors fac) {
    const long row = (long(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= long(B) * L) return;
    const int lane = threadIdx.x & 31;
    const int b = int(row / L), l = int(row % L);
    if (fac.d1 && l >= 1 && l <= P) {                // context row reconstructed from the factors (+ positional term, prompt_learner.py:53)
        const int t = sel ? sel[b] : 0;
        for (int c = lane * 4; c < D; c += 128) {
            const float4 p = *reinterpret_cast<const float4*>(pos + long(l) * D + c);
            *reinterpret_cast<float4*>(x_out + row * D + c) =
                make_float4(prompt_value(fac, t, l - 1, P, D, c) + p.x, prompt_value(fac, t, l - 1, P, D, c + 1) + p.y,
                            prompt_value(fac, t, l - 1, P, D, c + 2) + p.z, prompt_value(fac, t, l - 1, P, D, c + 3) + p.w);
        }
        return;
    }
    const float* src;
    if (ctx_table && l >= 1 && l <= P) src = ctx_table + (long(sel ? sel[b] : 0) * P + (l - 1)) * D;
    else src = emb + long(tok[row]) * D;
    for (int c = lane * 4; c < D; c += 128) {
        const float4 a = *reinterpret_cast<const float4*>(src + c), p = *reinterpret_cast<const float4*>(pos + long(l) * D + c);
        *reinterpret_cast<float4*>(x_out + row * D + c) = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    }
}

// d ctx_table[t, p, :] = sum over {b : sel[b] == t} g[b, 1+p, :]        grid = (P, T, ceil(D / 128)), block = 128 columns x 8 batch slices
__global__ void assemble_text_bwd_kernel(const float* __restrict__ g, const int* __restrict__ sel, float* __restrict__ d_ctx, int B, int L, int P,
                                         int D) {
    __shared__ float part[8][128];
    const int p = blockIdx.x, t = blockIdx.y;
    const int c = blockIdx.z * 128 + threadIdx.x, sl = threadIdx.y;
    float s = 0.f;
    if (c < D)
        for (int b = sl; b < B; b += 8)                  // fixed order per slice, slices combined in order below: deterministic
            if ((sel ? sel[b] : 0) == t) s += g[(long(b) * L + 1 + p) * D + c];
    part[sl][threadIdx.x] = s;
    __syncthreads();
    if (sl == 0 && c < D) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) v += part[j][threadIdx.x];
        d_ctx[(long(t) * P + p) * D + c] = v;
    }
}

// Deep-prompt injection (the *intended* semantics of model.py:190-193, opt-in): x[b, 1+p, :] += prompt[sel[b], p, :]
__global__ void inject_prompt_rows_kernel(float* __restrict__ x, const float* __restrict__ prompt, const int* __restrict__ sel, int B, int L, int P,
                                          int D) {
    const long e = (long(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (e >= long(B) * P * D) return;
    const int c = int(e % D), p = int((e / D) % P), b = int(e / (long(D) * P));
    const float4 a = *reinterpret_cast<const float4*>(prompt + (long(sel ? sel[b] : 0) * P + p) * D + c);
    float4* dst = reinterpret_cast<float4*>(x + (long(b) * L + 1 + p) * D + c);
    float4 v = *dst;
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    *dst = v;
}

// ------------------------------------------------------------------------------------------------ encoder heads
// feat[b] = normalize( LN(x[row_idx[b]]) @ proj ),  proj [D, E] row-major (model.py:254-257, prompt_learner.py:57-61, slinet.py:122,133)
// grid = (ceil(B/HB), E/32): a block normalises HB rows into smem (cheap, recomputed per column slice) and produces a 32-wide
// slice of the projection for them with the D-long reduction split over its 8 warps (partials combined through smem); the L2
// normalisation over E follows in head_norm_kernel.
constexpr int HB = 8;
// KC = projection rows whose loads are in flight together.  8: with 32 the kernel measured 40-58 us against 15-22 us (ncu, cold caches),
// so the wider batch is not used
template <int KC>
__global__ void __launch_bounds__(256)
head_fwd_kernel(const float* __restrict__ x, const int* __restrict__ row_idx, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ proj, float* __restrict__ z_out, int B, int D, int E, float eps) {
    extern __shared__ float sm[];                    // y[HB][D] | part[8][HB][32]
    float* y = sm;
    float* part = sm + HB * D;
    const int b0 = blockIdx.x * HB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (b0 + warp < B) {                             // 8 warps: one row each
        const float* xr = x + long(row_idx[b0 + warp]) * D;
        float s = 0.f;
        for (int c = lane; c < D; c += 32) s += xr[c];
        const float mean = warp_sum(s) / D;
        float q = 0.f;
        for (int c = lane; c < D; c += 32) { const float d = xr[c] - mean; q += d * d; }
        const float rstd = rsqrtf(warp_sum(q) / D + eps);
        for (int c = lane; c < D; c += 32) y[warp * D + c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
    } else {
        for (int c = lane; c < D; c += 32) y[warp * D + c] = 0.f;
    }
    __syncthreads();
    const int e = blockIdx.y * 32 + lane;
    float acc[HB];
#pragma unroll
    for (int h = 0; h < HB; ++h) acc[h] = 0.f;
    if (e < E) {
        const int kspan = D / 8;                     // D % 64 == 0 for every supported width
        for (int k0 = warp * kspan; k0 < (warp + 1) * kspan; k0 += KC) {
            float w[KC];
#pragma unroll
            for (int i = 0; i < KC; ++i) w[i] = __ldg(proj + long(k0 + i) * E + e);
#pragma unroll
            for (int i = 0; i < KC; ++i)
#pragma unroll
                for (int h = 0; h < HB; ++h) acc[h] = fmaf(y[h * D + k0 + i], w[i], acc[h]);
        }
    }
#pragma unroll
    for (int h = 0; h < HB; ++h) part[(warp * HB + h) * 32 + lane] = acc[h];
    __syncthreads();
    {                                                // warp w finishes sample w
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += part[(w * HB + warp) * 32 + lane];
        if (e < E && b0 + warp < B) z_out[long(b0 + warp) * E + e] = v;
    }
}

// feat[b] = z[b] / ||z[b]||   (one warp per sample)
__global__ void head_norm_kernel(const float* __restrict__ z, float* __restrict__ feat, int B, int E) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    float ss = 0.f;
    for (int e = lane; e < E; e += 32) { const float v = z[long(b) * E + e]; ss += v * v; }
    const float inv = 1.f / sqrtf(warp_sum(ss));
    for (int e = lane; e < E; e += 32) feat[long(b) * E + e] = z[long(b) * E + e] * inv;
}

// feat[b] = z[b] / ||z[b]|| AND the task-id of sample b in the same warp (sprompt.py:336-368: argmin over tasks of the min L1 distance to
// the task's centres, first occurrence on ties) -- the un-prompted pass of the evaluation ends in its selection, no extra launch.
// Same summation order as nearest_center_kernel (loss.cu), so the selection is bit-identical to the two-kernel form.
__global__ void head_norm_select_kernel(const float* __restrict__ z, float* __restrict__ feat, const float* __restrict__ centers, int T, int C,
                                        int* __restrict__ sel, int B, int E) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    float ss = 0.f;
    for (int e = lane; e < E; e += 32) { const float v = z[long(b) * E + e]; ss += v * v; }
    const float inv = 1.f / sqrtf(warp_sum(ss));
    for (int e = lane; e < E; e += 32) feat[long(b) * E + e] = z[long(b) * E + e] * inv;
    float best = CUDART_INF_F;
    int best_t = 0;
    for (int t = 0; t < T; ++t) {
        float tmin = CUDART_INF_F;
        for (int c = 0; c < C; ++c) {
            const float* k = centers + (long(t) * C + c) * E;
            float s = 0.f;
            for (int d = lane; d < E; d += 32) s += fabsf(z[long(b) * E + d] * inv - k[d]);
            s = warp_sum(s);
            tmin = fminf(tmin, s);
        }
        if (tmin < best) { best = tmin; best_t = t; }
    }
    if (lane == 0) sel[b] = best_t;
}

// Backward of the head in two kernels (the first version ran everything for one sample in one block: 64 blocks, each streaming
// the whole 1.5 MB projection through a serial loop -- 251 us for 25 MFLOP):
//   head_bwd_dy_kernel : (dfeat -> dz via the L2-norm backward) + dz_direct, then dy = dz @ proj^T for HBB samples x 64 rows of proj
//                        per block; dy is parked in g[row_idx[b], :] (those rows are assigned by this op anyway)
//   head_bwd_ln_kernel : one warp per sample: LN backward of the parked dy in place, plus the bf16 shadow
// dfeat = gradient wrt the normalised feature, dz_direct = gradient wrt the raw projection (what VisionTransformer.forward /
// TextEncoder.forward return in the reference); either may be NULL.
constexpr int HBB = 8;
__global__ void __launch_bounds__(256)
head_bwd_dy_kernel(const float* __restrict__ dfeat, const float* __restrict__ dz_direct, const float* __restrict__ z,
                   const int* __restrict__ row_idx, const float* __restrict__ proj, float* __restrict__ g, int B, int D, int E) {
    extern __shared__ float sm[];                    // dz[HBB][E]
    const int b0 = blockIdx.x * HBB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (b0 + warp < B) {                             // 8 warps: one sample each.  f = z*inv ; dz = (df - f (f.df)) * inv
        const int b = b0 + warp;
        const float* zr = z + long(b) * E;
        const float* dfr = dfeat ? dfeat + long(b) * E : nullptr;
        float a = 0.f, c = 0.f;
        for (int e = lane; e < E; e += 32) { a += zr[e] * zr[e]; c += dfr ? zr[e] * dfr[e] : 0.f; }
        a = warp_sum(a); c = warp_sum(c);
        const float inv = 1.f / sqrtf(a);
        for (int e = lane; e < E; e += 32) {
            float v = dfr ? (dfr[e] - zr[e] * inv * (c * inv)) * inv : 0.f;
            if (dz_direct) v += dz_direct[long(b) * E + e];
            sm[warp * E + e] = v;
        }
    } else {
        for (int e = lane; e < E; e += 32) sm[warp * E + e] = 0.f;
    }
    __syncthreads();
    const int k_end = min(D, (blockIdx.y + 1) * 64);
    if (E == 512) {
        // the usual width: the 16 loads of a projection row are issued together, and the next row's before this row's reductions
        float w[16], wn[16];
        int k = blockIdx.y * 64 + warp;
        if (k < k_end) {
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = __ldg(proj + long(k) * E + lane + 32 * i);
        }
        for (; k < k_end; k += 8) {
            const bool more = k + 8 < k_end;
            if (more) {
#pragma unroll
                for (int i = 0; i < 16; ++i) wn[i] = __ldg(proj + long(k + 8) * E + lane + 32 * i);
            }
            float acc[HBB];
#pragma unroll
            for (int h = 0; h < HBB; ++h) acc[h] = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i)
#pragma unroll
                for (int h = 0; h < HBB; ++h) acc[h] = fmaf(sm[h * E + lane + 32 * i], w[i], acc[h]);
#pragma unroll
            for (int h = 0; h < HBB; ++h) {
                const float v = warp_sum(acc[h]);
                if (lane == 0 && b0 + h < B) g[long(row_idx[b0 + h]) * D + k] = v;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = wn[i];
        }
        return;
    }
    for (int k = blockIdx.y * 64 + warp; k < k_end; k += 8) {          // dy[b, k] = sum_e dz[b, e] proj[k, e]
        float acc[HBB];
#pragma unroll
        for (int h = 0; h < HBB; ++h) acc[h] = 0.f;
        for (int e = lane; e < E; e += 32) {
            const float w = __ldg(proj + long(k) * E + e);
#pragma unroll
            for (int h = 0; h < HBB; ++h) acc[h] = fmaf(sm[h * E + e], w, acc[h]);
        }
#pragma unroll
        for (int h = 0; h < HBB; ++h) {
            const float v = warp_sum(acc[h]);
            if (lane == 0 && b0 + h < B) g[long(row_idx[b0 + h]) * D + k] = v;
        }
    }
}

template <int NV, bool F16 = false>
__global__ void head_bwd_ln_kernel(const float* __restrict__ x, const int* __restrict__ row_idx, const float* __restrict__ gamma,
                                   float* __restrict__ g, __nv_bfloat16* __restrict__ g_bf16, int B, int D, float eps, float shadow_scale = 1.0f) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    const long row = row_idx[b];
    float4 xv[NV], dyv[NV], dx[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const long o = row * D + (i * 32 + lane) * 4;
        xv[i] = *reinterpret_cast<const float4*>(x + o);
        dyv[i] = *reinterpret_cast<const float4*>(g + o);
    }
    ln_bwd_row<NV>(xv, dyv, gamma, D, eps, lane, dx);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const long o = row * D + (i * 32 + lane) * 4;
        *reinterpret_cast<float4*>(g + o) = dx[i];
        if (g_bf16) {
            const float m = F16 ? shadow_scale : 1.0f;
            *reinterpret_cast<uint2*>(g_bf16 + o) = make_uint2(pack_h2<F16>(dx[i].x * m, dx[i].y * m), pack_h2<F16>(dx[i].z * m, dx[i].w * m));
        }
    }
}

template <typename F>
static int dispatch_nv(int D, F f) {
    switch (D) {
        case 512: return f(std::integral_constant<int, 4>{});
        case 768: return f(std::integral_constant<int, 6>{});
        case 1024: return f(std::integral_constant<int, 8>{});
        case 256: return f(std::integral_constant<int, 2>{});
        case 128: return f(std::integral_constant<int, 1>{});
    }
    return set_error(LPI_ERR_UNSUPPORTED, "row width D=%d not supported (128, 256, 512, 768, 1024)", D);
}

}  // namespace lpi

using namespace lpi;

static inline unsigned warp_grid(long rows, int threads) { return unsigned((rows * 32 + threads - 1) / threads); }

extern "C" int lpi_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* out_f32, void* out_bf16, long long M, int D,
                                 float eps, void* stream) {
    if (M <= 0) return LPI_OK;
    if (!out_f32 && !out_bf16) return set_error(LPI_ERR_ARG, "layernorm_fwd: no output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = dispatch_nv(D, [&](auto nv) {
        launch_pdl(layernorm_fwd_kernel<decltype(nv)::value>, dim3(warp_grid(M, 256)), dim3(256), 0, st, x, gamma, beta, out_f32,
                   static_cast<__nv_bfloat16*>(out_bf16), long(M), D, eps);
        return 0;
    });
    return rc ? rc : check_launch("layernorm_fwd");
}

extern "C" int lpi_layernorm_fwd_f16(const float* x, const float* gamma, const float* beta, float* out_f32, void* out_f16, long long M, int D,
                                     float eps, void* stream) {
    if (M <= 0) return LPI_OK;
    if (!out_f32 && !out_f16) return set_error(LPI_ERR_ARG, "layernorm_fwd_f16: no output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rc = dispatch_nv(D, [&](auto nv) {
        launch_pdl(layernorm_fwd_kernel<decltype(nv)::value, true>, dim3(warp_grid(M, 256)), dim3(256), 0, st, x, gamma, beta, out_f32,
                   static_cast<__nv_bfloat16*>(out_f16), long(M), D, eps);
        return 0;
    });
    return rc ? rc : check_launch("layernorm_fwd_f16");
}

extern "C" int lpi_layernorm_bwd_f16(const float* dy_scaled, const float* x, const float* gamma, float* g, void* g_f16, long long M, int D,
                                     float eps, int accumulate, float grad_scale, void* stream) {
    if (M <= 0) return LPI_OK;
    if (!(grad_scale > 0.f)) return set_error(LPI_ERR_ARG, "layernorm_bwd_f16: grad_scale must be positive");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rc = dispatch_nv(D, [&](auto nv) {
        launch_pdl(layernorm_bwd_kernel<decltype(nv)::value, true>, dim3(warp_grid(M, 128)), dim3(128), 0, st, dy_scaled, x, gamma, g,
                   static_cast<__nv_bfloat16*>(g_f16), long(M), D, eps, accumulate, 1.0f / grad_scale, grad_scale);
        return 0;
    });
    return rc ? rc : check_launch("layernorm_bwd_f16");
}

extern "C" int lpi_layernorm_bwd_dy16(const void* dy16, int is_f16, const float* x, const float* gamma, float* g, void* g16, long long M, int D,
                                      float eps, int accumulate, float grad_scale, void* stream) {
    if (M <= 0) return LPI_OK;
    if (is_f16 && !(grad_scale > 0.f)) return set_error(LPI_ERR_ARG, "layernorm_bwd_dy16: grad_scale must be positive");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float* dy = static_cast<const float*>(dy16);       // reinterpreted inside the kernel
    auto gh = static_cast<__nv_bfloat16*>(g16);
    const int rc = dispatch_nv(D, [&](auto nv) {
        constexpr int NV = decltype(nv)::value;
        if (is_f16) launch_pdl(layernorm_bwd_kernel<NV, true, true>, dim3(warp_grid(M, 128)), dim3(128), 0, st, dy, x, gamma, g, gh, long(M), D, eps, accumulate, 1.0f / grad_scale, grad_scale);
        else launch_pdl(layernorm_bwd_kernel<NV, false, true>, dim3(warp_grid(M, 128)), dim3(128), 0, st, dy, x, gamma, g, gh, long(M), D, eps, accumulate, 1.0f, 1.0f);
        return 0;
    });
    return rc ? rc : check_launch("layernorm_bwd_dy16");
}

extern "C" int lpi_layernorm_bwd(const float* dy, const float* x, const float* gamma, float* g, void* g_bf16, long long M, int D, float eps,
                                 int accumulate, void* stream) {
    if (M <= 0) return LPI_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = dispatch_nv(D, [&](auto nv) {
        launch_pdl(layernorm_bwd_kernel<decltype(nv)::value>, dim3(warp_grid(M, 128)), dim3(128), 0, st, dy, x, gamma, g,
                   static_cast<__nv_bfloat16*>(g_bf16), long(M), D, eps, accumulate, 1.0f, 1.0f);
        return 0;
    });
    return rc ? rc : check_launch("layernorm_bwd");
}

static int im2col_entry(const float* images, void* out16, bool f16, int B, int resolution, int patch, void* stream) {
    if (B <= 0) return LPI_OK;
    if (patch % 4 || resolution % patch) return set_error(LPI_ERR_ARG, "im2col: bad resolution %d / patch %d", resolution, patch);
    const int G = resolution / patch;
    const long n = long(B) * G * G * 3 * patch * patch / 4;
    auto* out = static_cast<__nv_bfloat16*>(out16);
    if (f16) im2col_kernel<true><<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(images, out, B, resolution, patch);
    else im2col_kernel<false><<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(images, out, B, resolution, patch);
    return check_launch("im2col");
}

extern "C" int lpi_im2col_patches(const float* images, void* out_bf16, int B, int resolution, int patch, void* stream) {
    return im2col_entry(images, out_bf16, false, B, resolution, patch, stream);
}

extern "C" int lpi_im2col_patches_f32(const float* images, float* out_f32, int B, int resolution, int patch, void* stream) {
    if (B <= 0) return LPI_OK;
    if (patch % 4 || resolution % patch) return set_error(LPI_ERR_ARG, "im2col: bad resolution %d / patch %d", resolution, patch);
    const int G = resolution / patch;
    const long n = long(B) * G * G * 3 * patch * patch / 4;
    im2col_f32_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(images, out_f32, B, resolution, patch);
    return check_launch("im2col_f32");
}

extern "C" int lpi_im2col_patches_f16(const float* images, void* out_f16, int B, int resolution, int patch, void* stream) {
    return im2col_entry(images, out_f16, true, B, resolution, patch, stream);
}

static int assemble_vision_impl(const float* patch_emb, const float* cls, const float* pos, const float* prompt_table, PromptFactors fac,
                                const int* sel, const float* ln_gamma, const float* ln_beta, float* x_out, int B, int n_patch, int P, int D,
                                float eps, void* stream) {
    if (B <= 0) return LPI_OK;
    if (P > 0 && !prompt_table && !fac.d1) return set_error(LPI_ERR_ARG, "assemble_vision: P=%d but neither a prompt table nor factors", P);
    const long rows = long(B) * (1 + P + n_patch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = dispatch_nv(D, [&](auto nv) {
        assemble_vision_kernel<decltype(nv)::value><<<warp_grid(rows, 256), 256, 0, st>>>(patch_emb, cls, pos, prompt_table, sel, ln_gamma,
                                                                                            ln_beta, x_out, B, n_patch, P, D, eps, fac);
        return 0;
    });
    return rc ? rc : check_launch("assemble_vision");
}

static int check_factors(const float* d1, const float* d2, const float* d3, int r, int Lp, const char* what) {
    if (!d1 || !d2 || !d3) return set_error(LPI_ERR_ARG, "%s: null factor", what);
    if (r < 1 || r > 8 || Lp < 1) return set_error(LPI_ERR_ARG, "%s: bad rank r=%d / layer count %d", what, r, Lp);
    return LPI_OK;
}

extern "C" int lpi_assemble_vision(const float* patch_emb, const float* cls, const float* pos, const float* prompt_table, const int* sel,
                                   const float* ln_gamma, const float* ln_beta, float* x_out, int B, int n_patch, int P, int D, float eps,
                                   void* stream) {
    return assemble_vision_impl(patch_emb, cls, pos, prompt_table, PromptFactors{nullptr, nullptr, nullptr, 0, 0, 1.f}, sel, ln_gamma, ln_beta,
                                x_out, B, n_patch, P, D, eps, stream);
}

extern "C" int lpi_assemble_vision_factors(const float* patch_emb, const float* cls, const float* pos, const float* dim1_share,
                                           const float* dim2_vis, const float* dim3_vis, int r, int n_layers, float scale, const int* sel,
                                           const float* ln_gamma, const float* ln_beta, float* x_out, int B, int n_patch, int P, int D,
                                           float eps, void* stream) {
    if (int rc = check_factors(dim1_share, dim2_vis, dim3_vis, r, n_layers, "assemble_vision_factors")) return rc;
    return assemble_vision_impl(patch_emb, cls, pos, nullptr, PromptFactors{dim1_share, dim2_vis, dim3_vis, r, n_layers, scale}, sel, ln_gamma,
                                ln_beta, x_out, B, n_patch, P, D, eps, stream);
}

static int assemble_vision_bwd_impl(const float* g, const float* prompt_table, PromptFactors fac, const int* sel, const float* ln_gamma,
                                    float* d_prompt, int B, int L, int P, int n_tables, int D, float eps, void* stream) {
    if (B <= 0 || P <= 0) return LPI_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int threads = 512;
    int rc = dispatch_nv(D, [&](auto nv) {
        assemble_vision_bwd_kernel<decltype(nv)::value><<<dim3(P, n_tables), threads, (threads / 32) * D * sizeof(float), st>>>(
            g, prompt_table, sel, ln_gamma, d_prompt, B, L, P, D, eps, fac);
        return 0;
    });
    return rc ? rc : check_launch("assemble_vision_bwd");
}

extern "C" int lpi_assemble_vision_bwd(const float* g, const float* prompt_table, const int* sel, const float* ln_gamma, float* d_prompt,
                                       int B, int L, int P, int n_tables, int D, float eps, void* stream) {
    return assemble_vision_bwd_impl(g, prompt_table, PromptFactors{nullptr, nullptr, nullptr, 0, 0, 1.f}, sel, ln_gamma, d_prompt, B, L, P,
                                    n_tables, D, eps, stream);
}

extern "C" int lpi_assemble_vision_factors_bwd(const float* g, const float* dim1_share, const float* dim2_vis, const float* dim3_vis, int r,
                                               int n_layers, float scale, const int* sel, const float* ln_gamma, float* d_prompt, int B, int L,
                                               int P, int n_tables, int D, float eps, void* stream) {
    if (int rc = check_factors(dim1_share, dim2_vis, dim3_vis, r, n_layers, "assemble_vision_factors_bwd")) return rc;
    return assemble_vision_bwd_impl(g, nullptr, PromptFactors{dim1_share, dim2_vis, dim3_vis, r, n_layers, scale}, sel, ln_gamma, d_prompt, B,
                                    L, P, n_tables, D, eps, stream);
}

static int assemble_text_impl(const float* token_embedding, const long long* tokens, const float* pos, const float* ctx_table,
                              PromptFactors fac, const int* sel, float* x_out, int B, int L, int P, int D, void* stream) {
    if (B <= 0) return LPI_OK;
    if (D % 128) return set_error(LPI_ERR_ARG, "assemble_text: D=%d must be a multiple of 128", D);
    assemble_text_kernel<<<warp_grid(long(B) * L, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(token_embedding, tokens, pos, ctx_table,
                                                                                                      sel, x_out, B, L, P, D, fac);
    return check_launch("assemble_text");
}

extern "C" int lpi_assemble_text(const float* token_embedding, const long long* tokens, const float* pos, const float* ctx_table,
                                 const int* sel, float* x_out, int B, int L, int P, int D, void* stream) {
    return assemble_text_impl(token_embedding, tokens, pos, ctx_table, PromptFactors{nullptr, nullptr, nullptr, 0, 0, 1.f}, sel, x_out, B, L, P,
                              D, stream);
}

extern "C" int lpi_assemble_text_factors(const float* token_embedding, const long long* tokens, const float* pos, const float* dim1_share,
                                         const float* dim2_txt, const float* dim3_txt, int r, int n_layers, float scale, const int* sel,
                                         float* x_out, int B, int L, int P, int D, void* stream) {
    if (int rc = check_factors(dim1_share, dim2_txt, dim3_txt, r, n_layers, "assemble_text_factors")) return rc;
    return assemble_text_impl(token_embedding, tokens, pos, nullptr, PromptFactors{dim1_share, dim2_txt, dim3_txt, r, n_layers, scale}, sel,
                              x_out, B, L, P, D, stream);
}

extern "C" int lpi_assemble_text_bwd(const float* g, const int* sel, float* d_ctx, int B, int L, int P, int n_tables, int D, void* stream) {
    if (B <= 0 || P <= 0) return LPI_OK;
    assemble_text_bwd_kernel<<<dim3(P, n_tables, (D + 127) / 128), dim3(128, 8), 0, static_cast<cudaStream_t>(stream)>>>(g, sel, d_ctx, B, L, P, D);
    return check_launch("assemble_text_bwd");
}

extern "C" int lpi_inject_prompt_rows(float* x, const float* prompt, const int* sel, int B, int L, int P, int D, void* stream) {
    if (B <= 0 || P <= 0) return LPI_OK;
    if (D % 4) return set_error(LPI_ERR_ARG, "inject_prompt_rows: D=%d must be a multiple of 4", D);
    const long n = long(B) * P * D / 4;
    inject_prompt_rows_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, prompt, sel, B, L, P, D);
    return check_launch("inject_prompt_rows");
}

extern "C" int lpi_head_fwd(const float* x, const int* row_idx, const float* ln_gamma, const float* ln_beta, const float* proj, float* z_out,
                            float* feat_out, int B, int D, int E, float eps, void* stream) {
    if (B <= 0) return LPI_OK;
    if (D % 64) return set_error(LPI_ERR_ARG, "head_fwd: D=%d must be a multiple of 64", D);
    const int smem = (HB * D + 8 * HB * 32) * sizeof(float);
    if (smem > 48 * 1024) return set_error(LPI_ERR_UNSUPPORTED, "head_fwd: D=%d too wide", D);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    head_fwd_kernel<8><<<dim3((B + HB - 1) / HB, (E + 31) / 32), 256, smem, st>>>(x, row_idx, ln_gamma, ln_beta, proj, z_out, B, D, E, eps);
    head_norm_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(z_out, feat_out, B, E);
    return check_launch("head_fwd");
}

extern "C" int lpi_head_fwd_select(const float* x, const int* row_idx, const float* ln_gamma, const float* ln_beta, const float* proj,
                                   float* z_out, float* feat_out, const float* centers, int n_tasks, int n_centers, int* sel_out, int B, int D,
                                   int E, float eps, void* stream) {
    if (B <= 0) return LPI_OK;
    if (D % 64) return set_error(LPI_ERR_ARG, "head_fwd_select: D=%d must be a multiple of 64", D);
    if (!centers || !sel_out || n_tasks < 1 || n_centers < 1) return set_error(LPI_ERR_ARG, "head_fwd_select: no centres / selection buffer");
    const int smem = (HB * D + 8 * HB * 32) * sizeof(float);
    if (smem > 48 * 1024) return set_error(LPI_ERR_UNSUPPORTED, "head_fwd_select: D=%d too wide", D);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    head_fwd_kernel<8><<<dim3((B + HB - 1) / HB, (E + 31) / 32), 256, smem, st>>>(x, row_idx, ln_gamma, ln_beta, proj, z_out, B, D, E, eps);
    head_norm_select_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(z_out, feat_out, centers, n_tasks, n_centers, sel_out, B, E);
    return check_launch("head_fwd_select");
}

static int head_bwd_impl(const float* dfeat, const float* dz_direct, const float* z, const float* x, const int* row_idx, const float* ln_gamma,
                         const float* proj, float* g, void* g_bf16, int B, int D, int E, float eps, bool f16, float shadow_scale, void* stream) {
    if (B <= 0) return LPI_OK;
    if (!dfeat && !dz_direct) return set_error(LPI_ERR_ARG, "head_bwd: no upstream gradient");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int smem = HBB * E * sizeof(float);
    if (smem > 48 * 1024) return set_error(LPI_ERR_UNSUPPORTED, "head_bwd: E=%d too wide", E);
    head_bwd_dy_kernel<<<dim3((B + HBB - 1) / HBB, (D + 63) / 64), 256, smem, st>>>(dfeat, dz_direct, z, row_idx, proj, g, B, D, E);
    const int rc = dispatch_nv(D, [&](auto nv) {
        constexpr int NV = decltype(nv)::value;
        auto gb = static_cast<__nv_bfloat16*>(g_bf16);
        if (f16) head_bwd_ln_kernel<NV, true><<<(B * 32 + 127) / 128, 128, 0, st>>>(x, row_idx, ln_gamma, g, gb, B, D, eps, shadow_scale);
        else head_bwd_ln_kernel<NV, false><<<(B * 32 + 127) / 128, 128, 0, st>>>(x, row_idx, ln_gamma, g, gb, B, D, eps, 1.0f);
        return 0;
    });
    return rc ? rc : check_launch("head_bwd");
}

extern "C" int lpi_head_bwd(const float* dfeat, const float* dz_direct, const float* z, const float* x, const int* row_idx,
                            const float* ln_gamma, const float* proj, float* g, void* g_bf16, int B, int D, int E, float eps, void* stream) {
    return head_bwd_impl(dfeat, dz_direct, z, x, row_idx, ln_gamma, proj, g, g_bf16, B, D, E, eps, false, 1.0f, stream);
}

extern "C" int lpi_head_bwd_f16(const float* dfeat, const float* dz_direct, const float* z, const float* x, const int* row_idx,
                                const float* ln_gamma, const float* proj, float* g, void* g_f16, float grad_scale, int B, int D, int E, float eps,
                                void* stream) {
    return head_bwd_impl(dfeat, dz_direct, z, x, row_idx, ln_gamma, proj, g, g_f16, B, D, E, eps, true, grad_scale, stream);
}
