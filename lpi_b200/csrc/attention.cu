// Fused (flash-style) multi-head attention, forward and backward, for the prompted CLIP towers.
//
// Reference: nn.MultiheadAttention inside ResidualAttentionBlock.attention
// (retrieval/models/clip/model.py:172,183-185): softmax(q k^T / sqrt(64) + mask) v per head, heads are
// contiguous 64-wide column slices of the packed in_proj output; mask = none (vision, L = 213 / 197) or
// additive causal -inf above the diagonal (text, L = 77, model.py:347-353); dropout 0.
//
// Layout: qkv [B*L, 3*D] bf16 (row = b*L + l; columns q | k | v, D = H*64), out [B*L, D] bf16.
// The score matrix never leaves registers: S tiles are mma.sync m16n8k16 accumulators, softmax is online
// over 64-key blocks, P is re-packed in registers as the A operand of P.V.  The forward stores the per-row
// log-sum-exp (log2 domain) so the backward recomputes P exactly as the forward saw it.
// Sequences are short (L <= 256) and attention is ~4% of the tower FLOPs; whole K/V
// of one head live in shared memory, so there is no K/V streaming pipeline to manage.
#include <cuda_fp16.h>
#include "ptx.cuh"
#include "lpi_internal.h"

namespace lpi {

constexpr int DH = 64;          // head width (both towers)
constexpr int QB = 64;          // rows per CTA (4 warps x 16)
constexpr int ATT_THREADS = 128;

__device__ __forceinline__ uint32_t sw_off(int row, int chunk) {       // 128-byte rows, 16-byte chunks XOR-swizzled
    return uint32_t(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// rows [row0, row0+nrows) of one head slice (64 bf16 = 8 chunks per row) -> swizzled smem tile; rows >= L are zero
__device__ __forceinline__ void load_head_rows(uint32_t smem_tile, const __nv_bfloat16* base, long ld, int row0, int nrows, int L) {
    for (int i = threadIdx.x; i < nrows * 8; i += ATT_THREADS) {
        const int r = i >> 3, c = i & 7;
        const uint32_t dst = smem_tile + sw_off(r, c);
        if (row0 + r < L) cp_async16(dst, base + long(row0 + r) * ld + c * 8);
        else asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
    }
}

// A fragments (16 rows x 64 cols = 4 k-steps) of the tile rows [r0, r0+16)
__device__ __forceinline__ void load_a_frags(uint32_t tile, int r0, uint32_t (&a)[4][4]) {
    const int lane = threadIdx.x & 31;
    const int row = r0 + (lane & 15), half = lane >> 4;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(tile + sw_off(row, ks * 2 + half), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}

// acc[8][4] (16 x 64) = A(16 x 64) . T[rows t0..t0+63][0..63]^T   (T row-major [n][k]: "K pattern")
// `steps` (1..4, warp-uniform) = number of 16-row groups of T that hold valid rows; the rest are skipped (their accumulators stay 0)
__device__ __forceinline__ void mma_a_tT(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile, int t0, int steps = 4) {
    const int lane = threadIdx.x & 31;
    const int m = lane >> 3, r = lane & 7;
#pragma unroll
    for (int np = 0; np < 4; ++np) {            // pairs of 8-wide n tiles
        if (np >= steps) break;
        const int row = t0 + np * 16 + (m >> 1) * 8 + r;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(tile + sw_off(row, ks * 2 + (m & 1)), b0, b1, b2, b3);
            mma_bf16(acc[2 * np], a[ks], b0, b1);
            mma_bf16(acc[2 * np + 1], a[ks], b2, b3);
        }
    }
}

// acc[8][4] (16 x 64) += P(16 x 64, bf16 A fragments p[4][4]) . T[rows t0..t0+63][0..63]   (T row-major [k][n]: "V pattern")
__device__ __forceinline__ void mma_p_t(float (&acc)[8][4], const uint32_t (&p)[4][4], uint32_t tile, int t0, int steps = 4) {
    const int lane = threadIdx.x & 31;
    const int m = lane >> 3, r = lane & 7;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {            // 16 rows of T per k-step
        if (ks >= steps) break;
        const int row = t0 + ks * 16 + (m & 1) * 8 + r;
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(tile + sw_off(row, np * 2 + (m >> 1)), b0, b1, b2, b3);
            mma_bf16(acc[2 * np], p[ks], b0, b1);
            mma_bf16(acc[2 * np + 1], p[ks], b2, b3);
        }
    }
}

__device__ __forceinline__ void acc_to_a(const float (&c)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        a[t][0] = pack_bf16x2(c[2 * t][0], c[2 * t][1]);
        a[t][1] = pack_bf16x2(c[2 * t][2], c[2 * t][3]);
        a[t][2] = pack_bf16x2(c[2 * t + 1][0], c[2 * t + 1][1]);
        a[t][3] = pack_bf16x2(c[2 * t + 1][2], c[2 * t + 1][3]);
    }
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// 16 x 64 accumulator tile -> bf16 rows of a [.., ld] matrix (rows >= L skipped)
__device__ __forceinline__ void store_tile_bf16(const float (&c)[8][4], __nv_bfloat16* base, long ld, int row0, int L, float mul) {
    const int lane = threadIdx.x & 31;
    const int r = lane >> 2, cq = (lane & 3) * 2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int row = row0 + r + h * 8;
        if (row >= L) continue;
        __nv_bfloat16* p = base + long(row) * ld + cq;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
            *reinterpret_cast<uint32_t*>(p + nt * 8) = pack_bf16x2(c[nt][2 * h] * mul, c[nt][2 * h + 1] * mul);
    }
}

// same tile as fp32 rows (TF32 consumers: the text tower's out_proj / in_proj dgrad GEMMs read fp32 operands)
__device__ __forceinline__ void store_tile_f32(const float (&c)[8][4], float* base, long ld, int row0, int L, float mul) {
    const int lane = threadIdx.x & 31;
    const int r = lane >> 2, cq = (lane & 3) * 2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int row = row0 + r + h * 8;
        if (row >= L) continue;
        float* p = base + long(row) * ld + cq;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<float2*>(p + nt * 8) = make_float2(c[nt][2 * h] * mul, c[nt][2 * h + 1] * mul);
    }
}

// ------------------------------------------------------------------------------------------------ forward
template <bool CAUSAL>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ out_f32,
                float* __restrict__ lse2, int L, int H, float scale_log2) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int qc = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int D = H * DH;
    const long ld = 3L * D;
    const int q0 = qc * QB;
    const int kmax = CAUSAL ? min(L, q0 + QB) : L;          // keys this CTA can ever look at
    const int nkb = (kmax + 63) / 64;
    const uint32_t sQ = smem_u32(smem), sK = sQ + QB * 128, sV = sK + nkb * 64 * 128;
    const __nv_bfloat16* base = qkv + long(b) * L * ld + h * DH;
    load_head_rows(sQ, base, ld, q0, QB, L);
    load_head_rows(sK, base + D, ld, 0, nkb * 64, kmax);
    load_head_rows(sV, base + 2 * D, ld, 0, nkb * 64, kmax);
    cp_async_wait_all();
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (q0 + warp * 16 >= L) return;                        // this warp's 16 rows are all padding (L = 213: 2 of 16 warps)
    const int r_lo = q0 + warp * 16 + (lane >> 2);          // this thread's two rows: r_lo, r_lo + 8
    uint32_t qa[4][4];
    load_a_frags(sQ, warp * 16, qa);
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
    const int my_nkb = CAUSAL ? min(nkb, (q0 + warp * 16 + 15) / 64 + 1) : nkb;
    for (int kb = 0; kb < my_nkb; ++kb) {
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        const int steps = min(4, (kmax - kb * 64 + 15) >> 4);   // 16-key groups with at least one valid key
        mma_a_tT(s, qa, sK, kb * 64, steps);
        float bm[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = kb * 64 + nt * 8 + (lane & 3) * 2 + (e & 1);
                const int row = r_lo + (e >> 1) * 8;
                float v = s[nt][e] * scale_log2;
                if (col >= L || (CAUSAL && col > row)) v = -INFINITY;
                s[nt][e] = v;
                bm[e >> 1] = fmaxf(bm[e >> 1], v);
            }
        float corr[2], mu[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const float mn = fmaxf(m[hh], quad_max(bm[hh]));
            mu[hh] = (mn == -INFINITY) ? 0.f : mn;          // fully masked so far: exp2(-inf - 0) = 0, no NaN
            corr[hh] = exp2f(m[hh] - mu[hh]);
            m[hh] = mn;
            l[hh] *= corr[hh];
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = exp2f(s[nt][e] - mu[e >> 1]);
                s[nt][e] = p;
                l[e >> 1] += p;
            }
            o[nt][0] *= corr[0]; o[nt][1] *= corr[0]; o[nt][2] *= corr[1]; o[nt][3] *= corr[1];
        }
        uint32_t pa[4][4];
        acc_to_a(s, pa);
        mma_p_t(o, pa, sV, kb * 64, steps);
    }
    l[0] = quad_sum(l[0]);
    l[1] = quad_sum(l[1]);
    const float inv0 = l[0] > 0.f ? 1.f / l[0] : 0.f, inv1 = l[1] > 0.f ? 1.f / l[1] : 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= inv0; o[nt][1] *= inv0; o[nt][2] *= inv1; o[nt][3] *= inv1; }
    store_tile_bf16(o, out + long(b) * L * D + h * DH, D, q0 + warp * 16, L, 1.f);
    if (out_f32) store_tile_f32(o, out_f32 + long(b) * L * D + h * DH, D, q0 + warp * 16, L, 1.f);
    if (lse2 && (lane & 3) == 0) {
        float* lp = lse2 + (long(b) * H + h) * L;
        if (r_lo < L) lp[r_lo] = m[0] + log2f(l[0]);
        if (r_lo + 8 < L) lp[r_lo + 8] = m[1] + log2f(l[1]);
    }
}

// ------------------------------------------------------------------------------------------------ backward
// delta[b,h,l] = sum_d dO[b,l,h,d] * O[b,l,h,d]   (one warp per token row; 16-byte loads, 8 lanes per head, 3 shuffles per head)
template <bool F16>
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, float* __restrict__ delta,
                                  int B, int L, int H) {
    pdl_enter();
    const long row = (long(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= long(B) * L) return;
    const int lane = threadIdx.x & 31;
    const int b = int(row / L), l = int(row - long(b) * L);
    const uint4* po = reinterpret_cast<const uint4*>(o + row * H * DH);
    const uint4* pd = reinterpret_cast<const uint4*>(d_o + row * H * DH);
    for (int chunk = lane; chunk < H * 8; chunk += 32) {      // H * 8 is a multiple of 8, so the 8 lanes of a head stay together
        const uint4 a = __ldg(po + chunk), c = __ldg(pd + chunk);
        const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wc[4] = {c.x, c.y, c.z, c.w};
        float v = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 fa = F16 ? __half22float2(*reinterpret_cast<const __half2*>(&wa[e])) : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wa[e]));
            const float2 fc = F16 ? __half22float2(*reinterpret_cast<const __half2*>(&wc[e])) : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wc[e]));
            v = fmaf(fa.x, fc.x, v);
            v = fmaf(fa.y, fc.y, v);
        }
        const unsigned mask = __activemask();
        v += __shfl_xor_sync(mask, v, 1);
        v += __shfl_xor_sync(mask, v, 2);
        v += __shfl_xor_sync(mask, v, 4);
        if ((lane & 7) == 0) delta[(long(b) * H + (chunk >> 3)) * L + l] = v;
    }
}

// dQ: CTA = 64 query rows of one (b, h); recompute P from the stored LSE, dS = P o (dP - delta), dQ = scale * dS K
template <bool CAUSAL>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ d_o, const float* __restrict__ lse2,
                   const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dqkv_f32, int L, int H, float scale,
                   float scale_log2) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int qc = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int D = H * DH;
    const long ld = 3L * D;
    const int q0 = qc * QB;
    const int kmax = CAUSAL ? min(L, q0 + QB) : L;
    const int nkb = (kmax + 63) / 64;
    const uint32_t sQ = smem_u32(smem), sdO = sQ + QB * 128, sK = sdO + QB * 128, sV = sK + nkb * 64 * 128;
    const __nv_bfloat16* base = qkv + long(b) * L * ld + h * DH;
    load_head_rows(sQ, base, ld, q0, QB, L);
    load_head_rows(sdO, d_o + long(b) * L * D + h * DH, D, q0, QB, L);
    load_head_rows(sK, base + D, ld, 0, nkb * 64, kmax);
    load_head_rows(sV, base + 2 * D, ld, 0, nkb * 64, kmax);
    cp_async_wait_all();
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (q0 + warp * 16 >= L) return;
    const int r_lo = q0 + warp * 16 + (lane >> 2);
    uint32_t qa[4][4], da[4][4];
    load_a_frags(sQ, warp * 16, qa);
    load_a_frags(sdO, warp * 16, da);
    const float* lp = lse2 + (long(b) * H + h) * L;
    const float* dp = delta + (long(b) * H + h) * L;
    float lse_r[2], del_r[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int row = r_lo + hh * 8;
        lse_r[hh] = row < L ? lp[row] : INFINITY;           // padded rows: P = exp2(-inf) = 0
        del_r[hh] = row < L ? dp[row] : 0.f;
    }
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    const int my_nkb = CAUSAL ? min(nkb, (q0 + warp * 16 + 15) / 64 + 1) : nkb;
    for (int kb = 0; kb < my_nkb; ++kb) {
        float s[8][4], dpv[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
            dpv[i][0] = dpv[i][1] = dpv[i][2] = dpv[i][3] = 0.f;
        }
        const int steps = min(4, (kmax - kb * 64 + 15) >> 4);
        mma_a_tT(s, qa, sK, kb * 64, steps);
        mma_a_tT(dpv, da, sV, kb * 64, steps);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = kb * 64 + nt * 8 + (lane & 3) * 2 + (e & 1);
                const int row = r_lo + (e >> 1) * 8;
                float p = exp2f(s[nt][e] * scale_log2 - lse_r[e >> 1]);
                if (col >= L || (CAUSAL && col > row)) p = 0.f;
                s[nt][e] = p * (dpv[nt][e] - del_r[e >> 1]);
            }
        uint32_t dsa[4][4];
        acc_to_a(s, dsa);
        mma_p_t(dq, dsa, sK, kb * 64, steps);
    }
    if (dqkv_f32) store_tile_f32(dq, dqkv_f32 + long(b) * L * ld + h * DH, ld, q0 + warp * 16, L, scale);
    else store_tile_bf16(dq, dqkv + long(b) * L * ld + h * DH, ld, q0 + warp * 16, L, scale);
}

// dK, dV: CTA = 64 key rows of one (b, h); everything is computed transposed so each warp owns 16 keys.
template <bool CAUSAL>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ d_o, const float* __restrict__ lse2,
                    const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dqkv_f32, int L, int H, float scale,
                    float scale_log2) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int kc = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int D = H * DH;
    const long ld = 3L * D;
    const int k0 = kc * QB;
    const int q_begin = CAUSAL ? (k0 / 64) * 64 : 0;        // queries before the key chunk never attend to it
    const int nqb = (L - q_begin + 63) / 64;
    const uint32_t sK = smem_u32(smem), sV = sK + QB * 128, sQ = sV + QB * 128, sdO = sQ + nqb * 64 * 128;
    float* s_lse = reinterpret_cast<float*>(smem + (2 * QB + 2 * nqb * 64) * 128);
    float* s_del = s_lse + nqb * 64;
    const __nv_bfloat16* base = qkv + long(b) * L * ld + h * DH;
    load_head_rows(sK, base + D, ld, k0, QB, L);
    load_head_rows(sV, base + 2 * D, ld, k0, QB, L);
    load_head_rows(sQ, base + long(q_begin) * ld, ld, 0, nqb * 64, L - q_begin);
    load_head_rows(sdO, d_o + (long(b) * L + q_begin) * D + h * DH, D, 0, nqb * 64, L - q_begin);
    for (int i = threadIdx.x; i < nqb * 64; i += ATT_THREADS) {
        const int row = q_begin + i;
        s_lse[i] = row < L ? lse2[(long(b) * H + h) * L + row] : INFINITY;
        s_del[i] = row < L ? delta[(long(b) * H + h) * L + row] : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k0 + warp * 16 >= L) return;                        // all 16 keys of this warp are padding
    const int j_lo = k0 + warp * 16 + (lane >> 2);          // this thread's two keys: j_lo, j_lo + 8
    uint32_t ka[4][4], va[4][4];
    load_a_frags(sK, warp * 16, ka);
    load_a_frags(sV, warp * 16, va);
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    for (int qb = 0; qb < nqb; ++qb) {
        float st[8][4], dpt[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
            dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
        }
        const int steps = min(4, (L - q_begin - qb * 64 + 15) >> 4);    // 16-query groups with at least one valid query
        mma_a_tT(st, ka, sQ, qb * 64, steps);   // S^T tile: 16 keys x 64 queries
        mma_a_tT(dpt, va, sdO, qb * 64, steps); // dP^T tile
        uint32_t pa[4][4], dsa[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int qi = qb * 64 + nt * 8 + (lane & 3) * 2 + (e & 1);      // local query index
                const int row = q_begin + qi;                                     // global query
                const int key = j_lo + (e >> 1) * 8;
                float p = exp2f(st[nt][e] * scale_log2 - s_lse[qi]);
                if (key >= L || row >= L || (CAUSAL && key > row)) p = 0.f;
                st[nt][e] = p;
                dpt[nt][e] = p * (dpt[nt][e] - s_del[qi]);
            }
        acc_to_a(st, pa);
        acc_to_a(dpt, dsa);
        mma_p_t(dv, pa, sdO, qb * 64, steps);   // dV += P^T dO
        mma_p_t(dk, dsa, sQ, qb * 64, steps);   // dK += dS^T Q
    }
    if (dqkv_f32) {
        float* dbase = dqkv_f32 + long(b) * L * ld + h * DH;
        store_tile_f32(dk, dbase + D, ld, k0 + warp * 16, L, scale);
        store_tile_f32(dv, dbase + 2 * D, ld, k0 + warp * 16, L, 1.f);
    } else {
        __nv_bfloat16* dbase = dqkv + long(b) * L * ld + h * DH;
        store_tile_bf16(dk, dbase + D, ld, k0 + warp * 16, L, scale);
        store_tile_bf16(dv, dbase + 2 * D, ld, k0 + warp * 16, L, 1.f);
    }
}

static int attn_smem_fwd(int L) { return (QB + 2 * ((L + 63) / 64) * 64) * 128; }
static int attn_smem_dq(int L) { return (2 * QB + 2 * ((L + 63) / 64) * 64) * 128; }
static int attn_smem_dkv(int L) { const int n = ((L + 63) / 64) * 64; return (2 * QB + 2 * n) * 128 + 2 * n * 4; }

template <typename K>
static int set_smem(K kern, int bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return set_error(LPI_ERR_CUDA, "attention: cudaFuncSetAttribute(%d): %s", bytes, cudaGetErrorString(e));
    return 0;
}

}  // namespace lpi

using namespace lpi;

static int attn_check(int B, int L, int H) {
    if (B <= 0 || L <= 0 || H <= 0) return set_error(LPI_ERR_ARG, "attention: empty problem B=%d L=%d H=%d", B, L, H);
    if (L > 512) return set_error(LPI_ERR_ARG, "attention: L=%d > 512 (whole K/V of a head must fit in shared memory)", L);
    if (B > 65535 || H > 65535) return set_error(LPI_ERR_ARG, "attention: B or H exceeds the grid limit");
    return 0;
}

extern "C" int lpi_attn_fwd(const void* qkv, void* out, float* out_f32, float* lse2, int B, int L, int H, int causal, void* stream) {
    if (int rc = attn_check(B, L, H)) return rc;
    if (attn_tc_enabled(L)) return attn_fwd_tc(qkv, out, out_f32, lse2, B, L, H, causal, false, static_cast<cudaStream_t>(stream));
    const float scale_log2 = 0.125f * 1.4426950408889634f;      // 1/sqrt(64) * log2(e)
    const dim3 grid((L + QB - 1) / QB, H, B);
    const int smem = attn_smem_fwd(L);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto q = static_cast<const __nv_bfloat16*>(qkv);
    auto o = static_cast<__nv_bfloat16*>(out);
    if (causal) {
        if (int rc = set_smem(attn_fwd_kernel<true>, smem)) return rc;
        attn_fwd_kernel<true><<<grid, ATT_THREADS, smem, st>>>(q, o, out_f32, lse2, L, H, scale_log2);
    } else {
        if (int rc = set_smem(attn_fwd_kernel<false>, smem)) return rc;
        attn_fwd_kernel<false><<<grid, ATT_THREADS, smem, st>>>(q, o, out_f32, lse2, L, H, scale_log2);
    }
    return check_launch("attn_fwd");
}

extern "C" int lpi_attn_bwd(const void* qkv, const void* out, const void* d_out, const float* lse2, float* delta_ws, void* dqkv,
                            float* dqkv_f32, int B, int L, int H, int causal, void* stream) {
    if (int rc = attn_check(B, L, H)) return rc;
    if (!dqkv && !dqkv_f32) return set_error(LPI_ERR_ARG, "attn_bwd: no output");
    const float scale = 0.125f, scale_log2 = 0.125f * 1.4426950408889634f;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto q = static_cast<const __nv_bfloat16*>(qkv);
    auto o = static_cast<const __nv_bfloat16*>(out);
    auto d_o = static_cast<const __nv_bfloat16*>(d_out);
    auto dq = static_cast<__nv_bfloat16*>(dqkv);
    const long rows = long(B) * L;
    // out == NULL: delta_ws already holds rowsum(dO o O) (lpi_gemm_do_delta computed it in the epilogue of the out_proj dgrad)
    if (o) launch_pdl(attn_delta_kernel<false>, dim3(unsigned((rows * 32 + 255) / 256)), dim3(256), 0, st, o, d_o, delta_ws, B, L, H);
    if (attn_tc_enabled(L)) return attn_bwd_tc(qkv, d_out, lse2, delta_ws, dqkv, dqkv_f32, B, L, H, causal, false, st);
    const dim3 grid((L + QB - 1) / QB, H, B);
    if (causal) {
        if (int rc = set_smem(attn_bwd_dq_kernel<true>, attn_smem_dq(L))) return rc;
        if (int rc = set_smem(attn_bwd_dkv_kernel<true>, attn_smem_dkv(L))) return rc;
        attn_bwd_dq_kernel<true><<<grid, ATT_THREADS, attn_smem_dq(L), st>>>(q, d_o, lse2, delta_ws, dq, dqkv_f32, L, H, scale, scale_log2);
        attn_bwd_dkv_kernel<true><<<grid, ATT_THREADS, attn_smem_dkv(L), st>>>(q, d_o, lse2, delta_ws, dq, dqkv_f32, L, H, scale, scale_log2);
    } else {
        if (int rc = set_smem(attn_bwd_dq_kernel<false>, attn_smem_dq(L))) return rc;
        if (int rc = set_smem(attn_bwd_dkv_kernel<false>, attn_smem_dkv(L))) return rc;
        attn_bwd_dq_kernel<false><<<grid, ATT_THREADS, attn_smem_dq(L), st>>>(q, d_o, lse2, delta_ws, dq, dqkv_f32, L, H, scale, scale_log2);
        attn_bwd_dkv_kernel<false><<<grid, ATT_THREADS, attn_smem_dkv(L), st>>>(q, d_o, lse2, delta_ws, dq, dqkv_f32, L, H, scale, scale_log2);
    }
    return check_launch("attn_bwd");
}

// fp16 storage (q / k / v / P / dS / outputs), tcgen05 kernels only (L <= 256): the text tower's precision class
extern "C" int lpi_attn_fwd_f16(const void* qkv, void* out, float* lse2, int B, int L, int H, int causal, void* stream) {
    if (int rc = attn_check(B, L, H)) return rc;
    if (L > 256) return set_error(LPI_ERR_UNSUPPORTED, "attn_fwd_f16: L=%d > 256", L);
    return attn_fwd_tc(qkv, out, nullptr, lse2, B, L, H, causal, true, static_cast<cudaStream_t>(stream));
}

extern "C" int lpi_attn_bwd_f16(const void* qkv, const void* out, const void* d_out, const float* lse2, float* delta_ws, void* dqkv, int B, int L,
                                int H, int causal, void* stream) {
    if (int rc = attn_check(B, L, H)) return rc;
    if (L > 256) return set_error(LPI_ERR_UNSUPPORTED, "attn_bwd_f16: L=%d > 256", L);
    if (!dqkv) return set_error(LPI_ERR_ARG, "attn_bwd_f16: no output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long rows = long(B) * L;
    if (out)      // NULL: delta_ws already holds rowsum(dO o O) (lpi_gemm_do_delta)
        launch_pdl(attn_delta_kernel<true>, dim3(unsigned((rows * 32 + 255) / 256)), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(out),
                   static_cast<const __nv_bfloat16*>(d_out), delta_ws, B, L, H);
    return attn_bwd_tc(qkv, d_out, lse2, delta_ws, dqkv, nullptr, B, L, H, causal, true, st);
}
