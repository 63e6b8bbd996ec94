// tcgen05 / TMEM attention for the prompted CLIP towers (sequences of at most 256 tokens, head width 64).
//
// Reference: nn.MultiheadAttention inside ResidualAttentionBlock.attention (retrieval/models/clip/model.py:172,183-185):
// softmax(q k^T / sqrt(64) + mask) v per head; mask = none (vision, L = 213 / 197) or causal (text, L = 77,
// model.py:347-353).  Layout as in attention.cu: qkv [B*L, 3*D] bf16, columns q | k | v, heads = 64-wide slices.
//
// The legacy kernels in attention.cu (mma.sync, register-resident scores) ran at ~130 TFLOP/s and cost 28 % of the training
// step for 4 % of its FLOPs.  Here a whole head fits on chip, so there is no K/V streaming loop at all:
//
//   forward, one CTA per (128 query rows, head, sample), two CTAs per SM:
//     TMA   : Q tile [128 x 64], K and V [kpad x 64] (3-D tensor map over [B, L, 3D]: rows past L arrive as zeros)
//     MMA 1 : S[128 x kpad] = Q K^T           tcgen05.mma M=128 N=kpad K=16 x4, fp32 in TMEM columns [0, kpad)
//     warps : thread <-> query row (TMEM lane): pass 1 row max, pass 2 p = 2^(s c - m) -> bf16 P written to smem in the
//             SWIZZLE_128B K-major image (it overwrites the dead Q / K tiles), one mbarrier per 64-key block
//     MMA 2 : O[128 x 64] += P_blk V_blk      V is consumed as an MN-major B operand straight from its TMA image;
//             O aliases S columns [0, 64) (block 0 of S has been read by every row before the first P block is published)
//     warps : O / rowsum -> bf16 -> per-warp smem transpose -> 16-byte coalesced global stores; log2-domain LSE per row
//
//   backward, one CTA per (head, sample): see attn_bwd_tc_kernel below.
#include "ptx.cuh"
#include "lpi_internal.h"
#include <stdlib.h>

namespace lpi {

constexpr int TC_BM = 128;                 // query rows per tile = TMEM lanes
constexpr int TC_TILE = TC_BM * 128;       // 16 KB: 128 rows x 128 B (64 bf16)
constexpr int TC_MAXL = 256;

__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
// MN-major SWIZZLE_128B operand (PTX ISA "canonical layouts", MN-major / 128B swizzle): 64 contiguous M/N elements per 128-byte
// row, one row per K index, 8-row atoms SBO = 1024 B apart, the next 64 M/N elements LBO bytes away.
__device__ __forceinline__ uint64_t make_desc_mnmajor_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    return make_smem_desc(saddr, lbo_bytes, 1024, 2);
}
// runtime-N instruction descriptor (make_idesc is constexpr but N is only known at launch)
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) { return make_idesc(kFmtBF16, M, N, a_mn, b_mn); }

struct AttnFwdArgs {
    __nv_bfloat16* out;
    float* out_f32;
    float* lse2;
    int L, H, kv_rows;          // kv_rows = TMA box rows of the K / V loads = round_up(L, 16)
    float scale_log2;
};

constexpr int FWD_THREADS = 160;
constexpr int FWD_BAR_OFF = 6 * TC_TILE;                    // Q | K (2 tiles) | V (2 tiles) | P block 3
constexpr int FWD_SMEM = FWD_BAR_OFF + 128 + 1024;          // + barriers + 1024-byte alignment slack

template <bool CAUSAL>
__global__ void __launch_bounds__(FWD_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnFwdArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sQ = base, sK = base + TC_TILE, sV = base + 3 * TC_TILE;
    auto sP = [&](int blk) { return blk < 3 ? base + uint32_t(blk) * TC_TILE : base + 5u * TC_TILE; };   // blocks 0..2 overwrite Q | K
    const uint32_t bar = base + FWD_BAR_OFF;
    const uint32_t bar_qk = bar, bar_v = bar + 8, bar_s = bar + 16, bar_o = bar + 24;
    auto bar_p = [&](int c) { return bar + 32u + 8u * c; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + FWD_BAR_OFF + 64);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int L = p.L, D = p.H * 64;
    const int q0 = qt * TC_BM;
    const int nk = CAUSAL ? min(L, q0 + TC_BM) : L;        // keys this tile can attend to
    const int kpad = (nk + 15) & ~15;                      // MMA N (S) and K extent (P V)
    const int n_sub = (kpad + 31) >> 5;                    // 32-column sub-blocks of S / P

    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmKV);
            mbar_init(bar_qk, 1);
            mbar_init(bar_v, 1);
            mbar_init(bar_s, 1);
            mbar_init(bar_o, 1);
            for (int c = 0; c < 4; ++c) mbar_init(bar_p(c), 128);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<256>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            const uint32_t kv_bytes = uint32_t(p.kv_rows) * 128u;
            mbar_arrive_expect_tx(bar_qk, TC_TILE + kv_bytes);
            tma_load_3d(sQ, &tmQ, bar_qk, h * 64, q0, b);
            tma_load_3d(sK, &tmKV, bar_qk, D + h * 64, 0, b);
            mbar_arrive_expect_tx(bar_v, kv_bytes);
            tma_load_3d(sV, &tmKV, bar_v, 2 * D + h * 64, 0, b);
            mbar_wait(bar_qk, 0);
            tc_fence_after();
            const uint32_t idesc_s = idesc_bf16(TC_BM, kpad, 0, 0);
            const uint64_t dq = make_desc_kmajor_sw128(sQ), dk = make_desc_kmajor_sw128(sK);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
            umma_commit(bar_s);
            mbar_wait(bar_v, 0);
            const uint32_t idesc_o = idesc_bf16(TC_BM, 64, 0, 1);       // B = V is MN-major
            const int n_blk = (kpad + 63) >> 6;
            for (int c = 0; c < n_blk; ++c) {
                mbar_wait(bar_p(c), 0);
                tc_fence_after();
                const int ksteps = min(4, (kpad - 64 * c) >> 4);
                const uint64_t dp = make_desc_kmajor_sw128(sP(c));
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t dv = make_desc_mnmajor_sw128(sV + uint32_t(64 * c + 16 * ks) * 128u, 8192);
                    umma_f16_ss(tmem, dp + 2 * ks, dv, idesc_o, (c | ks) != 0);
                }
            }
            umma_commit(bar_o);
        }
    } else {
        const int r = warp * 32 + lane;                    // row inside the tile = TMEM lane
        const int row = q0 + r;
        const uint32_t t_lane = tmem + (uint32_t(warp * 32) << 16);
        const int lim = CAUSAL ? min(nk, row + 1) : nk;    // this row attends to columns [0, lim)
        // sub-blocks this warp has to evaluate (warp-uniform); later ones are all-masked for every row of the warp
        const int n_sub_w = CAUSAL ? min(n_sub, (min(nk, q0 + warp * 32 + 32) + 31) >> 5) : n_sub;
        mbar_wait(bar_s, 0);
        tc_fence_after();
        float mx = -INFINITY;
        for (int sb = 0; sb < n_sub_w; ++sb) {
            uint32_t v[32];
            LPI_TMEM_LD_X32(t_lane + uint32_t(sb * 32), v);
            tmem_ld_wait();
            if (sb * 32 + 32 <= lim) {
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (sb * 32 + j < lim) mx = fmaxf(mx, __uint_as_float(v[j]));
            }
        }
        const float sc = p.scale_log2;
        const float m2 = mx * sc;                          // lim >= 1, so mx is finite
        float sum = 0.f;
        for (int sb = 0; sb < n_sub; ++sb) {
            uint32_t pk[16];
            if (sb < n_sub_w) {
                uint32_t v[32];
                LPI_TMEM_LD_X32(t_lane + uint32_t(sb * 32), v);
                tmem_ld_wait();
                if (sb * 32 + 32 <= lim) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), sc, -m2));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sc, -m2));
                        sum += p0 + p1;
                        pk[j] = pack_bf16x2(p0, p1);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), sc, -m2));
                        float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sc, -m2));
                        if (sb * 32 + 2 * j >= lim) p0 = 0.f;
                        if (sb * 32 + 2 * j + 1 >= lim) p1 = 0.f;
                        sum += p0 + p1;
                        pk[j] = pack_bf16x2(p0, p1);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[j] = 0u;
            }
            const int blk = sb >> 1;
            const uint32_t rowaddr = sP(blk) + uint32_t(r) * 128u;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                st_shared_v4(rowaddr + (uint32_t(((sb & 1) * 4 + q) ^ (r & 7)) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            if ((sb & 1) || sb == n_sub - 1) {             // 64-key block complete
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(bar_p(blk));
            }
        }
        mbar_wait(bar_o, 0);
        tc_fence_after();
        uint32_t o[64];
        LPI_TMEM_LD_X64(t_lane, o);
        tmem_ld_wait();
        const float inv = 1.0f / sum;
        if (p.lse2 && row < L) p.lse2[(size_t(b) * p.H + h) * L + row] = m2 + log2f(sum);
        if (p.out_f32 && row < L) {
            float4* dst = reinterpret_cast<float4*>(p.out_f32 + (size_t(b) * L + row) * D + h * 64);
#pragma unroll
            for (int q = 0; q < 16; ++q)
                dst[q] = make_float4(__uint_as_float(o[4 * q]) * inv, __uint_as_float(o[4 * q + 1]) * inv,
                                     __uint_as_float(o[4 * q + 2]) * inv, __uint_as_float(o[4 * q + 3]) * inv);
        }
        // bf16 rows -> this warp's 4 KB of the (dead) P block 0 -> coalesced 16-byte stores, 8 lanes per row
        const uint32_t stg = sP(0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                w[e] = pack_bf16x2(__uint_as_float(o[8 * q + 2 * e]) * inv, __uint_as_float(o[8 * q + 2 * e + 1]) * inv);
            st_shared_v4(stg + uint32_t(r) * 128u + (uint32_t(q ^ (r & 7)) << 4), w[0], w[1], w[2], w[3]);
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int rr = warp * 32 + it * 4 + (lane >> 3), c = lane & 7;
            const uint4 v = ld_shared_v4(stg + uint32_t(rr) * 128u + (uint32_t(c ^ (rr & 7)) << 4));
            if (q0 + rr < L) *reinterpret_cast<uint4*>(p.out + (size_t(b) * L + q0 + rr) * D + h * 64 + c * 8) = v;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<256>(tmem);
    }
}

// ------------------------------------------------------------------------------------------------ backward
// One CTA per (head, sample); the whole head (Q, K, V, dO <= 256 rows each) is resident in shared memory and the five
// products run on tcgen05 with every accumulator in TMEM (all 512 columns):
//
//   for key half j (128 keys), for query tile i (128 rows):
//     S  = Q_i K_j^T, dP = dO_i V_j^T                                  -> TMEM [0,128) and [128,256)
//     threads (row <-> lane, two warps per lane quadrant splitting the 128 columns):
//         P = 2^(S c - lse), dS = P o (dP - delta)                      -> bf16, smem, SWIZZLE_128B rows of 64 keys
//     dQ_i += dS K_j      (A = dS K-major,  B = K_j MN-major)           -> TMEM [256 + 64 i, +64)
//     dV_j += P^T dO_i    (A = P  MN-major, B = dO_i MN-major)          -> TMEM [448, 512)
//     dK_j += dS^T Q_i    (A = dS MN-major, B = Q_i MN-major)           -> TMEM [384, 448)
//
// The same smem image of P / dS serves as a K-major operand (rows = queries) and as an MN-major operand (rows = keys), so
// nothing is ever transposed.  exp2 is evaluated once per score (the legacy dq + dkv kernels recomputed it twice).
struct AttnBwdArgs {
    const float* lse2;
    const float* delta;
    __nv_bfloat16* dqkv;
    float* dqkv_f32;
    int L, H, rows;             // rows = TMA box rows = round_up(L, 16)
    float scale, scale_log2;
};

constexpr int BWD_THREADS = 288;                            // 8 compute warps + 1 control warp
constexpr int BWD_BAR_OFF = 12 * TC_TILE;                   // Q, dO, K, V (2 tiles each) | P (2 key blocks) | dS (2 key blocks)
constexpr int BWD_SMEM = BWD_BAR_OFF + 128 + 1024;

template <bool F32>
__device__ __forceinline__ void bwd_store_row32(const AttnBwdArgs& p, const uint32_t (&v)[32], size_t elem_off, float mul) {
    if (F32) {
        float4* dst = reinterpret_cast<float4*>(p.dqkv_f32 + elem_off);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            dst[q] = make_float4(__uint_as_float(v[4 * q]) * mul, __uint_as_float(v[4 * q + 1]) * mul, __uint_as_float(v[4 * q + 2]) * mul,
                                 __uint_as_float(v[4 * q + 3]) * mul);
    } else {
        uint4* dst = reinterpret_cast<uint4*>(p.dqkv + elem_off);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            dst[q] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * q]) * mul, __uint_as_float(v[8 * q + 1]) * mul),
                                pack_bf16x2(__uint_as_float(v[8 * q + 2]) * mul, __uint_as_float(v[8 * q + 3]) * mul),
                                pack_bf16x2(__uint_as_float(v[8 * q + 4]) * mul, __uint_as_float(v[8 * q + 5]) * mul),
                                pack_bf16x2(__uint_as_float(v[8 * q + 6]) * mul, __uint_as_float(v[8 * q + 7]) * mul));
    }
}

template <bool CAUSAL, bool F32>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const AttnBwdArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sQ = base, sdO = base + 2 * TC_TILE, sK = base + 4 * TC_TILE, sV = base + 6 * TC_TILE;
    const uint32_t sP = base + 8 * TC_TILE, sdS = base + 10 * TC_TILE;
    const uint32_t bar = base + BWD_BAR_OFF;
    const uint32_t bar_ld = bar, bar_sdp = bar + 8, bar_pds = bar + 16;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + BWD_BAR_OFF + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, b = blockIdx.y;
    const int L = p.L, D = p.H * 64;
    const int n_t = (L + TC_BM - 1) / TC_BM;               // query tiles = key halves (1 or 2)
    constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DQ = 256, COL_DK = 384, COL_DV = 448;

    if (warp == 8) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQKV);
            tma_prefetch_desc(&tmDO);
            mbar_init(bar_ld, 1);
            mbar_init(bar_sdp, 1);
            mbar_init(bar_pds, 256);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            const uint32_t bytes = uint32_t(p.rows) * 128u;
            mbar_arrive_expect_tx(bar_ld, 4 * bytes);
            tma_load_3d(sQ, &tmQKV, bar_ld, h * 64, 0, b);
            tma_load_3d(sK, &tmQKV, bar_ld, D + h * 64, 0, b);
            tma_load_3d(sV, &tmQKV, bar_ld, 2 * D + h * 64, 0, b);
            tma_load_3d(sdO, &tmDO, bar_ld, h * 64, 0, b);
            mbar_wait(bar_ld, 0);
            tc_fence_after();
            const uint32_t idesc_out_k = idesc_bf16(TC_BM, 64, 0, 1);     // dQ: A K-major, B MN-major
            const uint32_t idesc_out_mn = idesc_bf16(TC_BM, 64, 1, 1);    // dK, dV: both MN-major
            int blk = 0;
            for (int j = 0; j < n_t; ++j) {
                const int kp = (min(TC_BM, L - TC_BM * j) + 15) & ~15;
                const uint32_t idesc_s = idesc_bf16(TC_BM, kp, 0, 0);
                for (int i = 0; i < n_t; ++i, ++blk) {
                    const uint64_t dq = make_desc_kmajor_sw128(sQ + i * TC_TILE), dk = make_desc_kmajor_sw128(sK + j * TC_TILE);
                    const uint64_t ddo = make_desc_kmajor_sw128(sdO + i * TC_TILE), dv = make_desc_kmajor_sw128(sV + j * TC_TILE);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss(tmem + COL_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss(tmem + COL_DP, ddo + 2 * k, dv + 2 * k, idesc_s, k != 0);
                    umma_commit(bar_sdp);
                    mbar_wait(bar_pds, blk & 1);
                    tc_fence_after();
                    // dQ_i += dS K_j   (K extent = kp keys)
                    for (int ks = 0; ks < (kp >> 4); ++ks) {
                        const uint64_t a = make_desc_kmajor_sw128(sdS + (ks >> 2) * TC_TILE) + 2 * (ks & 3);
                        const uint64_t bb = make_desc_mnmajor_sw128(sK + j * TC_TILE + ks * 2048, 8192);
                        umma_f16_ss(tmem + COL_DQ + 64 * i, a, bb, idesc_out_k, (j | ks) != 0);
                    }
                    // dV_j += P^T dO_i, dK_j += dS^T Q_i   (K extent = 128 queries of tile i; rows past L are zero)
                    const int qsteps = (min(TC_BM, L - TC_BM * i) + 15) >> 4;
                    for (int ks = 0; ks < qsteps; ++ks) {
                        const uint64_t a = make_desc_mnmajor_sw128(sP + ks * 2048, TC_TILE);
                        const uint64_t bb = make_desc_mnmajor_sw128(sdO + i * TC_TILE + ks * 2048, 8192);
                        umma_f16_ss(tmem + COL_DV, a, bb, idesc_out_mn, (i | ks) != 0);
                    }
                    for (int ks = 0; ks < qsteps; ++ks) {
                        const uint64_t a = make_desc_mnmajor_sw128(sdS + ks * 2048, TC_TILE);
                        const uint64_t bb = make_desc_mnmajor_sw128(sQ + i * TC_TILE + ks * 2048, 8192);
                        umma_f16_ss(tmem + COL_DK, a, bb, idesc_out_mn, (i | ks) != 0);
                    }
                }
            }
            umma_commit(bar_sdp);                           // everything retired -> final drain
        }
    } else {
        const int quad = warp & 3, half = warp >> 2;
        const int r = quad * 32 + lane;                     // TMEM lane
        const uint32_t t_lane = tmem + (uint32_t(quad * 32) << 16);
        const float sc = p.scale_log2;
        const size_t ld = 3 * size_t(D);
        auto drain_kv = [&](int j) {                        // dK_j, dV_j: lane = key
            const int key = TC_BM * j + r;
            uint32_t v[32];
            LPI_TMEM_LD_X32(t_lane + COL_DK + 32 * half, v);
            tmem_ld_wait();
            if (key < L) bwd_store_row32<F32>(p, v, (size_t(b) * L + key) * ld + D + h * 64 + 32 * half, p.scale);
            LPI_TMEM_LD_X32(t_lane + COL_DV + 32 * half, v);
            tmem_ld_wait();
            if (key < L) bwd_store_row32<F32>(p, v, (size_t(b) * L + key) * ld + 2 * D + h * 64 + 32 * half, 1.0f);
        };
        int blk = 0;
        for (int j = 0; j < n_t; ++j) {
            const int kp = (min(TC_BM, L - TC_BM * j) + 15) & ~15;
            for (int i = 0; i < n_t; ++i, ++blk) {
                const int row = TC_BM * i + r;
                float lse_r = INFINITY, del_r = 0.f;        // padded rows: P = 2^(-inf) = 0
                if (row < L) {
                    lse_r = p.lse2[(size_t(b) * p.H + h) * L + row];
                    del_r = p.delta[(size_t(b) * p.H + h) * L + row];
                }
                // valid key columns of this block: [0, lim); padded query rows are masked outright (their S may be garbage)
                const int lim = row < L ? (CAUSAL ? min(L, row + 1) : L) - TC_BM * j : 0;
                mbar_wait(bar_sdp, blk & 1);
                tc_fence_after();
                if (i == 0 && j > 0) {
                    drain_kv(j - 1);                        // complete since the commit above covers every earlier MMA
                    tc_fence_before();
                }
#pragma unroll 1
                for (int s = 0; s < 2; ++s) {
                    const int c0 = 64 * half + 32 * s;      // column inside the 128-key block
                    if (c0 >= kp) break;
                    uint32_t sv[32], dv[32];
                    LPI_TMEM_LD_X32(t_lane + COL_S + c0, sv);
                    LPI_TMEM_LD_X32(t_lane + COL_DP + c0, dv);
                    tmem_ld_wait();
                    uint32_t pk[16], dk[16];
                    const bool full = c0 + 32 <= lim;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * e]), sc, -lse_r));
                        float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * e + 1]), sc, -lse_r));
                        if (!full) {
                            if (c0 + 2 * e >= lim) p0 = 0.f;
                            if (c0 + 2 * e + 1 >= lim) p1 = 0.f;
                        }
                        const float d0 = p0 * (__uint_as_float(dv[2 * e]) - del_r);
                        const float d1 = p1 * (__uint_as_float(dv[2 * e + 1]) - del_r);
                        pk[e] = pack_bf16x2(p0, p1);
                        dk[e] = pack_bf16x2(d0, d1);
                    }
                    const uint32_t off = uint32_t(half) * TC_TILE + uint32_t(r) * 128u;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t ch = uint32_t((s * 4 + q) ^ (r & 7)) << 4;
                        st_shared_v4(sP + off + ch, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                        st_shared_v4(sdS + off + ch, dk[4 * q], dk[4 * q + 1], dk[4 * q + 2], dk[4 * q + 3]);
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(bar_pds);
            }
        }
        mbar_wait(bar_sdp, blk & 1);
        tc_fence_after();
        drain_kv(n_t - 1);
        for (int i = 0; i < n_t; ++i) {                     // dQ_i: lane = query
            const int row = TC_BM * i + r;
            uint32_t v[32];
            LPI_TMEM_LD_X32(t_lane + COL_DQ + 64 * i + 32 * half, v);
            tmem_ld_wait();
            if (row < L) bwd_store_row32<F32>(p, v, (size_t(b) * L + row) * ld + h * 64 + 32 * half, p.scale);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

// 3-D view [B, L, cols] of a row-major [B*L, cols] bf16 matrix, box = [1, box_rows, 64 columns], SWIZZLE_128B;
// rows past L (and past B) are zero-filled, so a tile never sees the next sample's tokens.
static int make_tmap_rows3d(CUtensorMap* m, const void* ptr, int B, int L, int cols, int box_rows) {
    if (int rc = ensure_tma_encoder()) return rc;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (cols % 8)) return set_error(LPI_ERR_ARG, "attention: operand must be 16-byte aligned");
    static PFN_encodeTiled enc = nullptr;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
            return set_error(LPI_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
        enc = reinterpret_cast<PFN_encodeTiled>(fn);
    }
    cuuint64_t dims[3] = {cuuint64_t(cols), cuuint64_t(L), cuuint64_t(B)};
    cuuint64_t strides[2] = {cuuint64_t(cols) * 2, cuuint64_t(L) * cols * 2};
    cuuint32_t box[3] = {64, cuuint32_t(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(LPI_ERR_CUDA, "attention: cuTensorMapEncodeTiled failed: %d", int(r));
    return 0;
}

bool attn_tc_enabled(int L) {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LPI_ATTN_TC");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1 && L <= TC_MAXL;
}

int attn_fwd_tc(const void* qkv, void* out, float* out_f32, float* lse2, int B, int L, int H, int causal, cudaStream_t st) {
    const int D = H * 64;
    const int kv_rows = (L + 15) & ~15;
    CUtensorMap tmQ, tmKV;
    if (int rc = make_tmap_rows3d(&tmQ, qkv, B, L, 3 * D, TC_BM)) return rc;
    if (int rc = make_tmap_rows3d(&tmKV, qkv, B, L, 3 * D, kv_rows)) return rc;
    AttnFwdArgs a{static_cast<__nv_bfloat16*>(out), out_f32, lse2, L, H, kv_rows, 0.125f * 1.4426950408889634f};
    const dim3 grid((L + TC_BM - 1) / TC_BM, H, B);
    static bool configured = false;
    if (!configured) {
        cudaError_t e1 = cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
        cudaError_t e2 = cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
        if (e1 != cudaSuccess || e2 != cudaSuccess) return set_error(LPI_ERR_CUDA, "attn_fwd_tc: cudaFuncSetAttribute failed");
        configured = true;
    }
    if (causal) attn_fwd_tc_kernel<true><<<grid, FWD_THREADS, FWD_SMEM, st>>>(tmQ, tmKV, a);
    else attn_fwd_tc_kernel<false><<<grid, FWD_THREADS, FWD_SMEM, st>>>(tmQ, tmKV, a);
    return check_launch("attn_fwd_tc");
}

int attn_bwd_tc(const void* qkv, const void* d_out, const float* lse2, const float* delta, void* dqkv, float* dqkv_f32, int B, int L, int H,
                int causal, cudaStream_t st) {
    const int D = H * 64;
    const int rows = (L + 15) & ~15;
    CUtensorMap tmQKV, tmDO;
    if (int rc = make_tmap_rows3d(&tmQKV, qkv, B, L, 3 * D, rows)) return rc;
    if (int rc = make_tmap_rows3d(&tmDO, d_out, B, L, D, rows)) return rc;
    AttnBwdArgs a{lse2, delta, static_cast<__nv_bfloat16*>(dqkv), dqkv_f32, L, H, rows, 0.125f, 0.125f * 1.4426950408889634f};
    const dim3 grid(H, B);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaSuccess;
        auto set = [&](const void* k) { if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM); };
        set(reinterpret_cast<const void*>(attn_bwd_tc_kernel<true, true>));
        set(reinterpret_cast<const void*>(attn_bwd_tc_kernel<true, false>));
        set(reinterpret_cast<const void*>(attn_bwd_tc_kernel<false, true>));
        set(reinterpret_cast<const void*>(attn_bwd_tc_kernel<false, false>));
        if (e != cudaSuccess) return set_error(LPI_ERR_CUDA, "attn_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    const bool f32 = dqkv_f32 != nullptr;
    if (causal) {
        if (f32) attn_bwd_tc_kernel<true, true><<<grid, BWD_THREADS, BWD_SMEM, st>>>(tmQKV, tmDO, a);
        else attn_bwd_tc_kernel<true, false><<<grid, BWD_THREADS, BWD_SMEM, st>>>(tmQKV, tmDO, a);
    } else {
        if (f32) attn_bwd_tc_kernel<false, true><<<grid, BWD_THREADS, BWD_SMEM, st>>>(tmQKV, tmDO, a);
        else attn_bwd_tc_kernel<false, false><<<grid, BWD_THREADS, BWD_SMEM, st>>>(tmQKV, tmDO, a);
    }
    return check_launch("attn_bwd_tc");
}

}  // namespace lpi
