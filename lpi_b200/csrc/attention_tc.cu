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
#include <cuda_fp16.h>
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
template <bool F16>
__device__ __forceinline__ uint32_t idesc_h(int M, int N, int a_mn, int b_mn) { return make_idesc(F16 ? kFmtF16 : kFmtBF16, M, N, a_mn, b_mn); }
// 16-bit storage type of q / k / v / P / dS / outputs: bf16 (vision tower) or fp16 (F16: text tower, 10-bit mantissa)
template <bool F16>
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    if (F16) {
        const __half2 v = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<const uint32_t*>(&v);
    }
    return pack_bf16x2(lo, hi);
}

// Debug timeline: with a trace buffer installed (lpi_debug_attn_trace), the first and the last CTA of the grid record clock64()
// at their pipeline events; 64 slots per CTA.  Null in production: one predictable branch per event.
static unsigned long long* g_attn_trace = nullptr;
#define LPI_TRACE(p, slot)                                                                                              \
    do {                                                                                                                \
        if ((p).trace) {                                                                                                \
            const unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);                        \
            const unsigned tot = gridDim.x * gridDim.y * gridDim.z;                                                     \
            if (lin == 0 || lin == tot - 1) (p).trace[(lin ? 64 : 0) + (slot)] = clock64();                             \
        }                                                                                                               \
    } while (0)

struct AttnFwdArgs {
    __nv_bfloat16* out;
    float* out_f32;
    float* lse2;
    int L, H, kv_rows;          // kv_rows = TMA box rows of the K / V loads = round_up(L, 16)
    float scale_log2;
    unsigned long long* trace;  // debug timeline (clock64 per event) or null
};

constexpr int FWD_THREADS = 160;
constexpr int FWD_BAR_OFF = 6 * TC_TILE;                    // Q | K (2 tiles) | V (2 tiles) | P block 3
constexpr int FWD_SMEM = FWD_BAR_OFF + 128 + 1024;          // + barriers + 1024-byte alignment slack

template <bool CAUSAL, bool F16>
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
            LPI_TRACE(p, 0);
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
    pdl_launch_dependents();
    pdl_wait();                                            // barriers / TMEM above overlap the previous kernel's tail (PDL)

    if (warp == 4) {
        if (lane == 0) {
            const uint32_t kv_bytes = uint32_t(p.kv_rows) * 128u;
            LPI_TRACE(p, 1);
            mbar_arrive_expect_tx(bar_qk, TC_TILE + kv_bytes);
            tma_load_3d(sQ, &tmQ, bar_qk, h * 64, q0, b);
            tma_load_3d(sK, &tmKV, bar_qk, D + h * 64, 0, b);
            mbar_arrive_expect_tx(bar_v, kv_bytes);
            tma_load_3d(sV, &tmKV, bar_v, 2 * D + h * 64, 0, b);
            mbar_wait(bar_qk, 0);
            LPI_TRACE(p, 2);
            tc_fence_after();
            const uint32_t idesc_s = idesc_h<F16>(TC_BM, kpad, 0, 0);
            const uint64_t dq = make_desc_kmajor_sw128(sQ), dk = make_desc_kmajor_sw128(sK);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
            umma_commit(bar_s);
            mbar_wait(bar_v, 0);
            LPI_TRACE(p, 3);
            const uint32_t idesc_o = idesc_h<F16>(TC_BM, 64, 0, 1);       // B = V is MN-major
            const int n_blk = (kpad + 63) >> 6;
            for (int c = 0; c < n_blk; ++c) {
                mbar_wait(bar_p(c), 0);
                LPI_TRACE(p, 4 + c);
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
        if (threadIdx.x == 0) LPI_TRACE(p, 8);
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
        if (threadIdx.x == 0) LPI_TRACE(p, 9);
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
                        pk[j] = pack_h2<F16>(p0, p1);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), sc, -m2));
                        float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sc, -m2));
                        if (sb * 32 + 2 * j >= lim) p0 = 0.f;
                        if (sb * 32 + 2 * j + 1 >= lim) p1 = 0.f;
                        sum += p0 + p1;
                        pk[j] = pack_h2<F16>(p0, p1);
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
        if (threadIdx.x == 0) LPI_TRACE(p, 10);
        mbar_wait(bar_o, 0);
        if (threadIdx.x == 0) LPI_TRACE(p, 11);
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
                w[e] = pack_h2<F16>(__uint_as_float(o[8 * q + 2 * e]) * inv, __uint_as_float(o[8 * q + 2 * e + 1]) * inv);
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
    if (threadIdx.x == 0) LPI_TRACE(p, 12);
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<256>(tmem);
        if (lane == 0) LPI_TRACE(p, 13);
    }
}

// ------------------------------------------------------------------------------------------------ backward
// One CTA per (head, sample); the whole head (Q, K, V, dO <= 256 rows each) is resident in shared memory and the five
// products run on tcgen05 with every accumulator in TMEM (all 512 columns).  Work is cut into blocks of 128 queries x 64 keys
// (n = jb * n_t + i: key block jb outer, query tile i inner) and software-pipelined over double-buffered S / dP accumulators
// and double-buffered P / dS smem slots, so the tensor pipe computes S, dP of block n + 1 and the five output products of
// block n - 1 while the compute warps run exp2 on block n:
//
//     S  = Q_i K_jb^T, dP = dO_i V_jb^T                               -> TMEM [128 u, +64) and [128 u + 64, +64), u = n & 1
//     compute warps (row <-> lane, two warps per lane quadrant splitting the 64 key columns):
//         P = 2^(S c - lse), dS = P o (dP - delta)                     -> bf16, smem slot u, SWIZZLE_128B rows of 64 keys
//     dQ_i  += dS K_jb     (M = 128; A = dS K-major,  B = K_jb MN-major)   -> TMEM [256 + 64 i, +64)
//     dV_jb += P^T dO_i    (M = 64;  A = P  MN-major, B = dO_i MN-major)   -> TMEM [448, 512), lanes 16 q' + (0..15) of each quadrant q'
//     dK_jb += dS^T Q_i    (M = 64;  A = dS MN-major, B = Q_i MN-major)    -> TMEM [384, 448)
//
// The same smem image of P / dS serves as a K-major operand (rows = queries) and as an MN-major operand (rows = keys), so
// nothing is ever transposed; exp2 is evaluated once per score (the legacy dq + dkv kernels recomputed it twice).  Four drain
// warps copy dK_jb / dV_jb out of TMEM as soon as the last query tile of a key block retires, off the compute warps' path;
// bf16 gradients leave through a swizzled staging tile and TMA stores (rows past L are clipped by the tensor map).
//
// What the first capture showed (profiles/r1_attention_tc.md): the tensor pipe was busy 19 % of the time because ONE thread
// issues every MMA and each issue cost ~110 cycles of dependent descriptor arithmetic; descriptors are therefore kept as
// 32-bit low words (the high word is the same constant for every SWIZZLE_128B operand) that advance by +2 (32 B, K-major k-step)
// or +128 (2048 B = 16 rows, MN-major k-step), and the issue loops are fully unrolled.
struct AttnBwdArgs {
    const float* lse2;
    const float* delta;
    __nv_bfloat16* dqkv;
    float* dqkv_f32;
    int L, H, rows;             // rows = TMA box rows = round_up(L, 16)
    float scale, scale_log2;
    unsigned long long* trace;  // debug timeline (clock64 per event) or null
};

constexpr int BWD_THREADS = 416;                            // warps 0-7 compute, 8-11 drain (one per lane quadrant), 12 control
constexpr int BWD_STAGE_OFF = 12 * TC_TILE;                 // P slots (2 x 16 KB) | dS slots (2 x 16 KB) | Q, dO, K, V (2 tiles each)
constexpr int BWD_BAR_OFF = BWD_STAGE_OFF + TC_TILE;        // | dK / dV staging (2 x 8 KB)
constexpr int BWD_SMEM = BWD_BAR_OFF + 128 + 1024;

constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);          // SBO = 1024 B, descriptor version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo_k(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ void umma_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
}
// smem tile -> global through the tensor map (rows / columns outside the tensor are clipped); bulk async-group completion
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void store_row_f32(float* dst, const uint32_t (&v)[N], float mul) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int q = 0; q < N / 4; ++q)
        d4[q] = make_float4(__uint_as_float(v[4 * q]) * mul, __uint_as_float(v[4 * q + 1]) * mul, __uint_as_float(v[4 * q + 2]) * mul,
                            __uint_as_float(v[4 * q + 3]) * mul);
}
// 32 fp32 values of one row -> bf16 -> chunks [chunk0, chunk0 + 4) of the row's 128-byte line in a SWIZZLE_128B staging tile
template <bool F16>
__device__ __forceinline__ void stage_row32_h(uint32_t tile, int row, int chunk0, const uint32_t (&v)[32], float mul) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
        st_shared_v4(tile + uint32_t(row) * 128u + (uint32_t((chunk0 + q) ^ (row & 7)) << 4),
                     pack_h2<F16>(__uint_as_float(v[8 * q]) * mul, __uint_as_float(v[8 * q + 1]) * mul),
                     pack_h2<F16>(__uint_as_float(v[8 * q + 2]) * mul, __uint_as_float(v[8 * q + 3]) * mul),
                     pack_h2<F16>(__uint_as_float(v[8 * q + 4]) * mul, __uint_as_float(v[8 * q + 5]) * mul),
                     pack_h2<F16>(__uint_as_float(v[8 * q + 6]) * mul, __uint_as_float(v[8 * q + 7]) * mul));
}

template <bool CAUSAL, bool F32, bool F16>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmOut,
                   const AttnBwdArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sP = base, sdS = base + 2 * TC_TILE;                        // slot u at + u * TC_TILE
    const uint32_t sQ = base + 4 * TC_TILE, sdO = base + 6 * TC_TILE, sK = base + 8 * TC_TILE, sV = base + 10 * TC_TILE;
    const uint32_t sStage = base + BWD_STAGE_OFF;                              // dK tile | dV tile, 64 rows x 128 B each
    const uint32_t bar = base + BWD_BAR_OFF;
    const uint32_t bar_ld0 = bar, bar_ld1 = bar + 8, bar_drained = bar + 16;
    auto bar_sdp = [&](int u) { return bar + 24u + 8u * u; };
    auto bar_pds = [&](int u) { return bar + 40u + 8u * u; };
    auto bar_out = [&](int u) { return bar + 56u + 8u * u; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + BWD_BAR_OFF + 80);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, b = blockIdx.y;
    const int L = p.L, D = p.H * 64;
    const int n_t = (L + TC_BM - 1) / TC_BM;               // query tiles (1 or 2)
    const int n_kb = (L + 63) >> 6;                        // 64-key blocks (1..4)
    const int n_blocks = n_kb * n_t;
    constexpr uint32_t COL_DQ = 256, COL_DK = 384, COL_DV = 448;
    constexpr int CTRL_WARP = 12;

    if (warp == CTRL_WARP) {
        if (lane == 0) {
            LPI_TRACE(p, 0);
            tma_prefetch_desc(&tmQKV);
            tma_prefetch_desc(&tmDO);
            if (!F32) tma_prefetch_desc(&tmOut);
            mbar_init(bar_ld0, 1);
            mbar_init(bar_ld1, 1);
            mbar_init(bar_drained, 128);
            for (int u = 0; u < 2; ++u) {
                mbar_init(bar_sdp(u), 1);
                mbar_init(bar_pds(u), 256);
                mbar_init(bar_out(u), 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
        if (lane == 0) {
            // PDL: everything above overlaps the tail of the previous kernel; its outputs (qkv, dO) are read from here on
            pdl_launch_dependents();
            pdl_wait();
            const uint32_t bytes = uint32_t(p.rows) * 128u;
            mbar_arrive_expect_tx(bar_ld0, 2 * bytes);
            tma_load_3d(sQ, &tmQKV, bar_ld0, h * 64, 0, b);
            tma_load_3d(sK, &tmQKV, bar_ld0, D + h * 64, 0, b);
            mbar_arrive_expect_tx(bar_ld1, 2 * bytes);
            tma_load_3d(sdO, &tmDO, bar_ld1, h * 64, 0, b);
            tma_load_3d(sV, &tmQKV, bar_ld1, 2 * D + h * 64, 0, b);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                                            // every thread, before its first global-memory access (lse / delta / outputs)

    if (warp == CTRL_WARP) {
        if (lane == 0) {
            LPI_TRACE(p, 1);
            const uint32_t idesc_dq = idesc_h<F16>(TC_BM, 64, 0, 1);        // A K-major, B MN-major
            const uint32_t idesc_kv = idesc_h<F16>(64, 64, 1, 1);           // M = 64: both MN-major
            const uint32_t q_lo = desc_lo_k(sQ), do_lo = desc_lo_k(sdO), k_lo = desc_lo_k(sK), v_lo = desc_lo_k(sV);
            const uint32_t qmn_lo = desc_lo_mn(sQ, 8192), domn_lo = desc_lo_mn(sdO, 8192), kmn_lo = desc_lo_mn(sK, 8192);
            const uint32_t ds_lo = desc_lo_k(sdS), dsmn_lo = desc_lo_mn(sdS, 8192), pmn_lo = desc_lo_mn(sP, 8192);
            constexpr uint32_t TILE16 = TC_TILE >> 4, HALF16 = TC_TILE >> 5;     // descriptor-unit (16 B) offsets of a tile / a 64-row half
            auto issue_sdp = [&](int n) {
                const int jb = n / n_t, i = n - jb * n_t, u = n & 1;
                const int kp = min(64, (L - 64 * jb + 15) & ~15);
                const uint32_t idesc_s = idesc_h<F16>(TC_BM, kp, 0, 0);
                const uint32_t qa = q_lo + i * TILE16, ka = k_lo + jb * HALF16, da = do_lo + i * TILE16, va = v_lo + jb * HALF16;
                const uint32_t ts = tmem + 128 * u;
                if (n == 0) { mbar_wait(bar_ld0, 0); LPI_TRACE(p, 2); tc_fence_after(); }
                if (n == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_lo(ts, qa + 2 * k, ka + 2 * k, idesc_s, k != 0);
                    mbar_wait(bar_ld1, 0); LPI_TRACE(p, 3); tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_lo(ts + 64, da + 2 * k, va + 2 * k, idesc_s, k != 0);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                       // two independent accumulation chains, interleaved
                        umma_bf16_lo(ts, qa + 2 * k, ka + 2 * k, idesc_s, k != 0);
                        umma_bf16_lo(ts + 64, da + 2 * k, va + 2 * k, idesc_s, k != 0);
                    }
                }
                umma_commit(bar_sdp(u));
            };
            issue_sdp(0);
            for (int n = 0; n < n_blocks; ++n) {
                const int jb = n / n_t, i = n - jb * n_t, u = n & 1;
                const int kp = min(64, (L - 64 * jb + 15) & ~15);
                if (n + 1 < n_blocks) issue_sdp(n + 1);         // its TMEM buffer was released with bar_pds of block n - 1
                mbar_wait(bar_pds(u), (n >> 1) & 1);
                LPI_TRACE(p, 8 + n);
                tc_fence_after();
                {                                               // dQ_i += dS K_jb
                    const uint32_t a = ds_lo + u * TILE16, bb = kmn_lo + jb * HALF16, td = tmem + COL_DQ + 64 * i;
                    const int ksteps = kp >> 4;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        if (ks < ksteps) umma_bf16_lo(td, a + 2 * ks, bb + 128 * ks, idesc_dq, (jb | ks) != 0);
                }
                if (i == 0 && jb > 0) {                         // dK / dV of the previous key block must have left TMEM
                    mbar_wait(bar_drained, (jb - 1) & 1);
                    LPI_TRACE(p, 16 + jb);
                    tc_fence_after();
                }
                {                                               // dV_jb += P^T dO_i ; dK_jb += dS^T Q_i (interleaved chains)
                    const int qsteps = (min(TC_BM, L - TC_BM * i) + 15) >> 4;      // rows past L are zero in P / dS
                    const uint32_t pa = pmn_lo + u * TILE16, dob = domn_lo + i * TILE16, dsa = dsmn_lo + u * TILE16, qb = qmn_lo + i * TILE16;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        if (ks < qsteps) {
                            umma_bf16_lo(tmem + COL_DV, pa + 128 * ks, dob + 128 * ks, idesc_kv, (i | ks) != 0);
                            umma_bf16_lo(tmem + COL_DK, dsa + 128 * ks, qb + 128 * ks, idesc_kv, (i | ks) != 0);
                        }
                }
                umma_commit(bar_out(u));
            }
        }
    } else if (warp < 8) {
        // ------------------------------------------------------------ compute warps: P, dS of every block; final dQ drain
        const int quad = warp & 3, half = warp >> 2;
        const int r = quad * 32 + lane;                     // TMEM lane = query row inside the tile
        const uint32_t t_lane = tmem + (uint32_t(quad * 32) << 16);
        const float sc = p.scale_log2;
        float lse_r[2] = {INFINITY, INFINITY}, del_r[2] = {0.f, 0.f};          // padded rows: P = 0
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = TC_BM * i + r;
            if (row < L) {
                lse_r[i] = p.lse2[(size_t(b) * p.H + h) * L + row];
                del_r[i] = p.delta[(size_t(b) * p.H + h) * L + row];
            }
        }
        const int c0 = 32 * half;                           // this warp's columns inside the 64-key block
        for (int n = 0; n < n_blocks; ++n) {
            const int jb = n / n_t, i = n - jb * n_t, u = n & 1;
            const int kp = min(64, (L - 64 * jb + 15) & ~15);
            const int row = TC_BM * i + r;
            // valid key columns of this block: [0, lim); padded query rows are masked outright
            const int lim = row < L ? (CAUSAL ? min(L, row + 1) : L) - 64 * jb : 0;
            const float lse_i = i ? lse_r[1] : lse_r[0], del_i = i ? del_r[1] : del_r[0];
            mbar_wait(bar_sdp(u), (n >> 1) & 1);
            if (threadIdx.x == 0) LPI_TRACE(p, 24 + n);
            if (n >= 2) mbar_wait(bar_out(u), ((n - 2) >> 1) & 1);             // slot u no longer read by the products of block n - 2
            if (threadIdx.x == 0) LPI_TRACE(p, 32 + n);
            tc_fence_after();
            if (c0 < kp) {
                uint32_t sv[32], dv[32];
                LPI_TMEM_LD_X32(t_lane + 128 * u + c0, sv);
                LPI_TMEM_LD_X32(t_lane + 128 * u + 64 + c0, dv);
                tmem_ld_wait();
                uint32_t pk[16], dk[16];
                const bool full = c0 + 32 <= lim;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * e]), sc, -lse_i));
                    float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * e + 1]), sc, -lse_i));
                    if (!full) {
                        if (c0 + 2 * e >= lim) p0 = 0.f;
                        if (c0 + 2 * e + 1 >= lim) p1 = 0.f;
                    }
                    const float d0 = p0 * (__uint_as_float(dv[2 * e]) - del_i);
                    const float d1 = p1 * (__uint_as_float(dv[2 * e + 1]) - del_i);
                    pk[e] = pack_h2<F16>(p0, p1);
                    dk[e] = pack_h2<F16>(d0, d1);
                }
                const uint32_t off = uint32_t(u) * TC_TILE + uint32_t(r) * 128u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t ch = uint32_t((half * 4 + q) ^ (r & 7)) << 4;
                    st_shared_v4(sP + off + ch, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                    st_shared_v4(sdS + off + ch, dk[4 * q], dk[4 * q + 1], dk[4 * q + 2], dk[4 * q + 3]);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_pds(u));
            if (threadIdx.x == 0) LPI_TRACE(p, 40 + n);
        }
        mbar_wait(bar_out((n_blocks - 1) & 1), ((n_blocks - 1) >> 1) & 1);     // the last commit covers every product
        if (threadIdx.x == 0) LPI_TRACE(p, 48);
        tc_fence_after();
        const size_t ld = 3 * size_t(D);
        for (int i = 0; i < n_t; ++i) {                     // dQ_i: lane = query
            const int row = TC_BM * i + r;
            uint32_t v[32];
            LPI_TMEM_LD_X32(t_lane + COL_DQ + 64 * i + 32 * half, v);
            tmem_ld_wait();
            if (F32) {
                if (row < L) store_row_f32<32>(p.dqkv_f32 + (size_t(b) * L + row) * ld + h * 64 + 32 * half, v, p.scale);
            } else {
                stage_row32_h<F16>(sP + i * TC_TILE, r, 4 * half, v, p.scale);   // the P slots are dead: every product has retired
            }
        }
        if (!F32) {
            fence_proxy_async_smem();
            named_bar_sync(2, 256);
            if (threadIdx.x == 0) {
                for (int i = 0; i < n_t; ++i) {
                    tma_store_3d(&tmOut, sP + i * TC_TILE, h * 64, TC_BM * i, b);
                    if (TC_BM * i + 64 < L) tma_store_3d(&tmOut, sP + i * TC_TILE + TC_TILE / 2, h * 64, TC_BM * i + 64, b);
                }
                tma_store_commit();
                tma_store_wait_all();
            }
        }
    } else if (warp < CTRL_WARP) {
        // ------------------------------------------------------------ drain warps 8..11: dK_jb, dV_jb (M = 64 accumulators keep
        // row m in lane 32 (m / 16) + m % 16, so each lane quadrant holds 16 keys in its lanes 0..15)
        const int quad = warp & 3;
        const uint32_t t_src = tmem + (uint32_t(quad * 32) << 16);
        const int krow = 16 * quad + lane;                  // key inside the 64-key block (lanes 0..15 only)
        const size_t ld = 3 * size_t(D);
        for (int jb = 0; jb < n_kb; ++jb) {
            const int n_last = jb * n_t + n_t - 1;
            if (!F32 && jb > 0) {                           // the previous block's TMA stores must have read the staging tiles
                if (warp == 8 && lane == 0) tma_store_wait_read();
                named_bar_sync(1, 128);
            }
            mbar_wait(bar_out(n_last & 1), (n_last >> 1) & 1);
            tc_fence_after();
            const int key = 64 * jb + krow;
#pragma unroll
            for (int part = 0; part < 4; ++part) {          // dK cols 0-31, 32-63, dV cols 0-31, 32-63
                const bool is_v = part >= 2;
                uint32_t v[32];
                LPI_TMEM_LD_X32(t_src + (is_v ? COL_DV : COL_DK) + 32 * (part & 1), v);
                tmem_ld_wait();
                if (part == 3) {
                    tc_fence_before();
                    mbar_arrive(bar_drained);               // values are in registers: the accumulators may be overwritten
                }
                if (lane < 16) {
                    if (F32) {
                        if (key < L)
                            store_row_f32<32>(p.dqkv_f32 + (size_t(b) * L + key) * ld + (is_v ? 2 : 1) * size_t(D) + h * 64 + 32 * (part & 1), v,
                                              is_v ? 1.0f : p.scale);
                    } else {
                        stage_row32_h<F16>(sStage + (is_v ? TC_TILE / 2 : 0), krow, 4 * (part & 1), v, is_v ? 1.0f : p.scale);
                    }
                }
            }
            if (!F32) {
                fence_proxy_async_smem();
                named_bar_sync(1, 128);
                if (warp == 8 && lane == 0) {
                    tma_store_3d(&tmOut, sStage, D + h * 64, 64 * jb, b);
                    tma_store_3d(&tmOut, sStage + TC_TILE / 2, 2 * D + h * 64, 64 * jb, b);
                    tma_store_commit();
                }
            }
        }
        if (!F32 && warp == 8 && lane == 0) tma_store_wait_all();
    }
    if (threadIdx.x == 0) LPI_TRACE(p, 49);
    if (warp == 8 && lane == 0) LPI_TRACE(p, 50);
    tc_fence_before();
    __syncthreads();
    if (warp == CTRL_WARP) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
        if (lane == 0) LPI_TRACE(p, 51);
    }
}

// 3-D view [B, L, cols] of a row-major [B*L, cols] bf16 matrix, box = [1, box_rows, 64 columns], SWIZZLE_128B;
// rows past L (and past B) are zero-filled, so a tile never sees the next sample's tokens.
static int make_tmap_rows3d(CUtensorMap* m, const void* ptr, int B, int L, int cols, int box_rows, bool f16 = false) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (cols % 8)) return set_error(LPI_ERR_ARG, "attention: operand must be 16-byte aligned");
    cuuint64_t dims[3] = {cuuint64_t(cols), cuuint64_t(L), cuuint64_t(B)};
    cuuint64_t strides[2] = {cuuint64_t(cols) * 2, cuuint64_t(L) * cols * 2};
    cuuint32_t box[3] = {64, cuuint32_t(box_rows), 1};
    return make_tmap_cached(m, ptr, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dims, strides, box);
}

bool attn_tc_enabled(int L) {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LPI_ATTN_TC");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1 && L <= TC_MAXL;
}

int attn_fwd_tc(const void* qkv, void* out, float* out_f32, float* lse2, int B, int L, int H, int causal, bool f16, cudaStream_t st) {
    const int D = H * 64;
    const int kv_rows = (L + 15) & ~15;
    CUtensorMap tmQ, tmKV;
    if (int rc = make_tmap_rows3d(&tmQ, qkv, B, L, 3 * D, TC_BM, f16)) return rc;
    if (int rc = make_tmap_rows3d(&tmKV, qkv, B, L, 3 * D, kv_rows, f16)) return rc;
    AttnFwdArgs a{static_cast<__nv_bfloat16*>(out), out_f32, lse2, L, H, kv_rows, 0.125f * 1.4426950408889634f, g_attn_trace};
    const dim3 grid((L + TC_BM - 1) / TC_BM, H, B);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaSuccess;
        auto set = [&](const void* k) { if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM); };
        set(reinterpret_cast<const void*>(attn_fwd_tc_kernel<true, false>));
        set(reinterpret_cast<const void*>(attn_fwd_tc_kernel<false, false>));
        set(reinterpret_cast<const void*>(attn_fwd_tc_kernel<true, true>));
        set(reinterpret_cast<const void*>(attn_fwd_tc_kernel<false, true>));
        if (e != cudaSuccess) return set_error(LPI_ERR_CUDA, "attn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    if (causal) {
        if (f16) launch_pdl(attn_fwd_tc_kernel<true, true>, grid, dim3(FWD_THREADS), FWD_SMEM, st, tmQ, tmKV, a);
        else launch_pdl(attn_fwd_tc_kernel<true, false>, grid, dim3(FWD_THREADS), FWD_SMEM, st, tmQ, tmKV, a);
    } else {
        if (f16) launch_pdl(attn_fwd_tc_kernel<false, true>, grid, dim3(FWD_THREADS), FWD_SMEM, st, tmQ, tmKV, a);
        else launch_pdl(attn_fwd_tc_kernel<false, false>, grid, dim3(FWD_THREADS), FWD_SMEM, st, tmQ, tmKV, a);
    }
    return check_launch("attn_fwd_tc");
}

int attn_bwd_tc(const void* qkv, const void* d_out, const float* lse2, const float* delta, void* dqkv, float* dqkv_f32, int B, int L, int H,
                int causal, bool f16, cudaStream_t st) {
    const int D = H * 64;
    const int rows = (L + 15) & ~15;
    if (f16 && dqkv_f32) return set_error(LPI_ERR_ARG, "attn_bwd: the fp16 path writes fp16 gradients only");
    CUtensorMap tmQKV, tmDO, tmOut;
    if (int rc = make_tmap_rows3d(&tmQKV, qkv, B, L, 3 * D, rows, f16)) return rc;
    if (int rc = make_tmap_rows3d(&tmDO, d_out, B, L, D, rows, f16)) return rc;
    if (dqkv_f32) tmOut = tmQKV;                            // unused by the fp32 variant (direct stores)
    else if (int rc = make_tmap_rows3d(&tmOut, dqkv, B, L, 3 * D, 64, f16)) return rc;
    AttnBwdArgs a{lse2, delta, static_cast<__nv_bfloat16*>(dqkv), dqkv_f32, L, H, rows, 0.125f, 0.125f * 1.4426950408889634f, g_attn_trace};
    const dim3 grid(H, B);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaSuccess;
        auto set = [&](const void* k) { if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM); };
        set(reinterpret_cast<const void*>(attn_bwd_tc_kernel<true, true, false>));
        set(reinterpret_cast<const void*>(attn_bwd_tc_kernel<true, false, false>));
        set(reinterpret_cast<const void*>(attn_bwd_tc_kernel<false, true, false>));
        set(reinterpret_cast<const void*>(attn_bwd_tc_kernel<false, false, false>));
        set(reinterpret_cast<const void*>(attn_bwd_tc_kernel<true, false, true>));
        set(reinterpret_cast<const void*>(attn_bwd_tc_kernel<false, false, true>));
        if (e != cudaSuccess) return set_error(LPI_ERR_CUDA, "attn_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    const bool f32 = dqkv_f32 != nullptr;
    if (causal) {
        if (f16) launch_pdl(attn_bwd_tc_kernel<true, false, true>, grid, dim3(BWD_THREADS), BWD_SMEM, st, tmQKV, tmDO, tmOut, a);
        else if (f32) launch_pdl(attn_bwd_tc_kernel<true, true, false>, grid, dim3(BWD_THREADS), BWD_SMEM, st, tmQKV, tmDO, tmOut, a);
        else launch_pdl(attn_bwd_tc_kernel<true, false, false>, grid, dim3(BWD_THREADS), BWD_SMEM, st, tmQKV, tmDO, tmOut, a);
    } else {
        if (f16) launch_pdl(attn_bwd_tc_kernel<false, false, true>, grid, dim3(BWD_THREADS), BWD_SMEM, st, tmQKV, tmDO, tmOut, a);
        else if (f32) launch_pdl(attn_bwd_tc_kernel<false, true, false>, grid, dim3(BWD_THREADS), BWD_SMEM, st, tmQKV, tmDO, tmOut, a);
        else launch_pdl(attn_bwd_tc_kernel<false, false, false>, grid, dim3(BWD_THREADS), BWD_SMEM, st, tmQKV, tmDO, tmOut, a);
    }
    return check_launch("attn_bwd_tc");
}

}  // namespace lpi

// debug aid (not part of the public header): install a device buffer of >= 128 uint64 that the attention kernels fill with
// clock64() stamps of their pipeline events (first and last CTA of the grid); NULL switches tracing off.
extern "C" int lpi_debug_attn_trace(void* device_buf) {
    lpi::g_attn_trace = static_cast<unsigned long long*>(device_buf);
    return 0;
}
