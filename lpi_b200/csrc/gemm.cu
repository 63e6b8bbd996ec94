// bf16 tensor-core GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (both operands K-major, fp32 accumulate)
//
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> 4-6 stage smem ring -> tcgen05.mma (M=128, N=BN, K=16)
//   -> double-buffered fp32 accumulators in TMEM -> tcgen05.ld -> fused epilogue.
//
// One persistent CTA per SM, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2..5 = epilogue (one TMEM lane quadrant each; thread <-> output row).
//
// Two users share the pipeline:
//   * MODE_GEMM  -- the encoder linears of the prompted CLIP towers and their dgrads
//                   (reference: nn.MultiheadAttention in_proj/out_proj + mlp.c_fc/c_proj,
//                   retrieval/models/clip/model.py:168-196) with bias / QuickGELU / residual /
//                   dQuickGELU fused into the epilogue;
//   * MODE_TOPK  -- the retrieval scorer (reference: `image_feats @ text_feats.t()` then a full
//                   np.argsort per row, retrieval/methods/sprompt.py:509,559-567,597-599): the score
//                   tile never leaves TMEM/registers; each epilogue thread keeps the running top-k
//                   of its query row ordered by (score desc, gallery index asc).
#include "ptx.cuh"
#include "lpi_internal.h"
#include <string.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>

namespace lpi {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int GEMM_THREADS = 192;
constexpr int TOPK_MAX = 16;
#ifndef LPI_RASTER_N
#define LPI_RASTER_N 1       // pair GEMM tile order: 1 = N fastest (A block fetched once, hit in L2 by its other N tiles), 0 = M fastest (A/B builds)
#endif
#ifndef LPI_H2_ACT
#define LPI_H2_ACT 1         // fp16 towers: (d)QuickGELU on packed halves (0 = fp32 arithmetic as in the bf16 tower; A/B builds)
#endif
#ifndef LPI_F16_PRECISE_ACT
#define LPI_F16_PRECISE_ACT 0   // fp16 towers: 1 = ex2 + rcp sigmoid in fp32 (2 MUFU ops per element) for precision studies; a build-time switch so
#endif                          // that the shipped epilogues carry one activation path, not a runtime branch per element
#ifndef LPI_AUX_LDCS
#define LPI_AUX_LDCS 0      // 1 = evict-first loads of the saved pre-activation (measured slower: 75.7 vs 72.5 us on the dGELU GEMM)
#endif
#ifndef LPI_EPI_DIRECT
#define LPI_EPI_DIRECT 0     // pair GEMM epilogue: 1 = accumulator rows straight to global memory with 256-bit accesses (no smem staging)
#endif
#ifndef LPI_EPI16
#define LPI_EPI16 1          // 16-bit staging of 16-bit outputs in the pair GEMM epilogue (0 = the fp32 transpose for every epilogue)
#endif

enum { MODE_GEMM = 0, MODE_TOPK = 1 };

struct GemmArgs {
    int M, N, K;
    int epi;
    const float* bias;             // [N] fp32 or null
    const float* resid;            // [M, ldo] fp32 (EPI_RESID) -- may alias out_f32
    float* out_f32;                // [M, ldo]
    __nv_bfloat16* out_bf16;       // [M, ldo]
    __nv_bfloat16* out2_bf16;      // [M, ldo] second output (pre-activation) or null
    const __nv_bfloat16* aux_bf16; // [M, ldo] (EPI_DGELU: saved pre-activation)
    float* out2_f32;               // [M, ldo] fp32 pre-activation (EPI_BIAS_GELU_F32)
    const float* aux_f32;          // [M, ldo] fp32 pre-activation (EPI_DGELU_F32)
    int ldo;
    int probe;                     // only in -DLPI_DEBUG_PROBE builds (tools/gemm_probe.py): 1 = the pair GEMM skips its A-tile loads (WRONG results, timing study)
    float* delta;                  // EPI_BF16 of the out_proj dgrad: delta[b, h, l] += rowsum over head h of out(fp32) * aux (aux = the saved attention
    int delta_L;                   //   output O [M, N], N = 64 H, row = b * delta_L + l); delta [M / L * H * L] is zeroed by the caller (lpi_gemm_do_delta)
    int clc;                       // pair GEMM: 1 = one cluster per tile in the grid, resident clusters steal the pending ones (cluster launch control)
    int precise_act;               // fp16 outputs: 1 = ex2 + rcp sigmoid (2 MUFU ops), 0 = tanh.approx (1 MUFU op, |err| <= 2^-12)
    // MODE_TOPK
    int k;                         // top-k (<= TOPK_MAX)
    int n_chunks;                  // gallery split per query tile (load balance)
    int tiles_per_chunk;
    long long gallery_offset;      // global index of gallery row 0 of this shard
    int seed_mode;                 // 1: keep the top-k of the per-TILE maxima only (threshold pre-pass: one candidate per 256 gallery rows)
    int init_thr_stride;           // element stride of init_thr (lets the k-th column of a [M, k] list be used in place)
    const float* init_thr;         // [M] or null: per-query score every kept candidate must reach (seeded by a pre-pass over a gallery sample)
    float* shared_thr;             // [M] or null: thresholds SHARED by the work items of one launch (cooperative mode, see coop_* below)
    float* topk_scores;            // [n_chunks, M, k]
    int* topk_idx;                 // [n_chunks, M, k]
};

template <int BN>
struct Cfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int LIST_OFF = BAR_OFF + 256;
    static constexpr int EPI_STAGE_BYTES = 4 * 32 * 32 * 4;                  // MODE_GEMM: per-warp transpose tiles (reuse the list region)
    static constexpr int SMEM_BYTES = LIST_OFF + (BM * TOPK_MAX * 8 > EPI_STAGE_BYTES ? BM * TOPK_MAX * 8 : EPI_STAGE_BYTES) + 1024;   // +1024 alignment slack
};

// The sequence of (m0, n0) tiles a CTA walks is a pure function of blockIdx, so each warp role
// re-derives it independently -- no inter-warp broadcast of scheduling decisions.
template <int MODE, int BN>
struct TileWalk {
    int num_m, num_n;
    // GEMM
    int tile, total;
    // TOPK
    int item, n_items, n_cur, n_end, q_tile, chunk;
    int tiles_per_chunk;
    __device__ TileWalk(const GemmArgs& p) {
        num_m = (p.M + BM - 1) / BM;
        num_n = (p.N + BN - 1) / BN;
        if (MODE == MODE_GEMM) {
            tile = blockIdx.x;
            total = num_m * num_n;
        } else {
            tiles_per_chunk = p.tiles_per_chunk;
            n_items = num_m * p.n_chunks;
            item = blockIdx.x;
            begin_item();
        }
    }
    __device__ void begin_item() {
        if (item < n_items) {
            q_tile = item % num_m;          // chunk-major: concurrent CTAs stream the same gallery region (L2 reuse)
            chunk = item / num_m;
            n_cur = chunk * tiles_per_chunk;
            n_end = min(num_n, n_cur + tiles_per_chunk);
        }
    }
    // returns false when the CTA is out of work; `first`/`last` flag the ends of a top-k work item
    __device__ bool next(int& m0, int& n0, bool& first, bool& last) {
        if (MODE == MODE_GEMM) {
            if (tile >= total) return false;
            m0 = (tile % num_m) * BM;       // M fastest: neighbouring CTAs share the weight (B) tile
            n0 = (tile / num_m) * BN;
            first = last = true;
            tile += gridDim.x;
            return true;
        } else {
            while (item < n_items && n_cur >= n_end) {   // empty chunk (cannot happen with sane args) or item done
                item += gridDim.x;
                begin_item();
            }
            if (item >= n_items) return false;
            m0 = q_tile * BM;
            n0 = n_cur * BN;
            first = (n_cur == chunk * tiles_per_chunk);
            ++n_cur;
            last = (n_cur >= n_end);
            if (last) {
                item += gridDim.x;
                begin_item();
            }
            return true;
        }
    }
};

// QuickGELU z * sigmoid(1.702 z) (model.py:163-165) and its derivative.  The plain `/` compiles to the IEEE division
// sequence (FCHK + a slow-path CALL per element), which made the GELU epilogues 4x slower than the MMA loop, so:
//   bf16 outputs : sigmoid(x) = 0.5 tanh.approx(0.5 x) + 0.5 -- ONE MUFU op per element, |error| <= 2^-11 (output rounding is 2^-9)
//   fp32 outputs : 1 / (1 + exp(-x)) with ex2.approx + rcp.approx (2 MUFU ops, ~2 ulp)
template <bool PRECISE>
__device__ __forceinline__ float sigmoid_fast(float x) {
    if (PRECISE) return __fdividef(1.0f, 1.0f + __expf(-x));
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
}
// fp16 towers: the same two functions on PACKED halves -- the 16-bit epilogues are instruction-bound (the dQuickGELU epilogue executed 4x the
// instructions of the plain one and held the tensor pipe at 50 %, profiles/r2_gemm_vs_cublas.md), and both their input (the fp16
// pre-activation, stored for the backward either way) and their output are fp16, so fp16 arithmetic costs ~1.5 ulp(fp16) on top of the
// output rounding: ONE MUFU op and 3-5 HFMA2-class instructions per TWO elements.
__device__ __forceinline__ uint32_t h2_sigmoid_1702(uint32_t z2) {          // sigmoid(1.702 z) = 0.5 tanh(0.851 z) + 0.5
    const __half2 z = *reinterpret_cast<const __half2*>(&z2);
    const __half2 a = __hmul2(z, __float2half2_rn(0.851f));
    uint32_t t;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&a)));
    const __half2 s = __hfma2(*reinterpret_cast<const __half2*>(&t), __float2half2_rn(0.5f), __float2half2_rn(0.5f));
    return *reinterpret_cast<const uint32_t*>(&s);
}
__device__ __forceinline__ uint32_t h2_quick_gelu(uint32_t z2) {
    const uint32_t s2 = h2_sigmoid_1702(z2);
    const __half2 r = __hmul2(*reinterpret_cast<const __half2*>(&z2), *reinterpret_cast<const __half2*>(&s2));
    return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ float2 h2_quick_gelu_grad(uint32_t z2) {          // s + 1.702 z s (1 - s)
    const uint32_t s2 = h2_sigmoid_1702(z2);
    const __half2 s = *reinterpret_cast<const __half2*>(&s2);
    const __half2 u = __hmul2(*reinterpret_cast<const __half2*>(&z2), __float2half2_rn(1.702f));
    const __half2 q = __hfma2(__hneg2(s), s, s);                              // s (1 - s)
    return __half22float2(__hfma2(u, q, s));
}
template <bool PRECISE = false>
__device__ __forceinline__ float quick_gelu(float z) { return z * sigmoid_fast<PRECISE>(1.702f * z); }
template <bool PRECISE = false>
__device__ __forceinline__ float quick_gelu_grad(float z) {
    const float s = sigmoid_fast<PRECISE>(1.702f * z);
    return s * (1.0f + 1.702f * z * (1.0f - s));
}

// Fused epilogue of one 32-row x 32-column accumulator block owned by one warp.
// tcgen05.ld hands every thread one ROW (32 consecutive columns); writing that straight to global memory makes each warp
// store instruction touch 32 different rows in 16-byte pieces (and the residual / pre-activation reads likewise): 32 LSU
// wavefronts per instruction, which ran the GELU / dGELU / residual epilogues 3-4x slower than the MMA main loop.
// So the block is transposed through a per-warp 4 KB smem tile (128-byte rows, 16-byte chunks XOR-swizzled by row & 7 so
// both directions are bank-conflict free) using 16-byte accesses only, and all global traffic is issued with 8 lanes
// covering one row: every instruction moves 4 rows x 128 B (fp32) or 4 x 64 B (bf16) in full sectors.
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// 16-bit outputs / saved pre-activations are bf16 (vision tower) or fp16 (F16 = true: the text tower, whose operands need the
// 10-bit mantissa -- same precision class as TF32 at the full kind::f16 rate and half the operand bytes)
template <bool F16>
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    if (F16) {
        const __half2 v = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<const uint32_t*>(&v);
    }
    return pack_bf16x2(lo, hi);
}
template <bool F16>
__device__ __forceinline__ float2 unpack_h2(uint32_t w) {
    if (F16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}

template <int EPI, bool F16 = false>
__device__ __forceinline__ void epilogue_block(const GemmArgs& p, const float4* __restrict__ st, int row0, int col0, int lane,
                                               const uint2* pre_aux = nullptr, const float4* pre_resid = nullptr) {
    constexpr bool kBias = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_F32 ||
                            EPI == EPI_BIAS_GELU_F32);
    constexpr bool kF32Out = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_F32 || EPI == EPI_ACC_F32 || EPI == EPI_BIAS_F32 ||
                              EPI == EPI_BIAS_GELU_F32 || EPI == EPI_DGELU_F32);
    const int rows = min(32, p.M - row0);                      // warp-uniform
    if (rows <= 0) return;
    const int rsub = lane >> 3, c = lane & 7;                  // this lane: rows rsub, rsub+4, ..., columns 4c .. 4c+3
    const int col = col0 + c * 4;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kBias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + rsub;
        v[i] = f4_add(st[r * 8 + (c ^ (r & 7))], b4);
    }
    const size_t base = size_t(row0 + rsub) * p.ldo + col;     // + i*4 rows
    if (kF32Out) {
        if (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_ACC_F32 || EPI == EPI_DGELU_F32) {
            const float* src = (EPI == EPI_BIAS_RESID_F32) ? p.resid : (EPI == EPI_ACC_F32 ? p.out_f32 : p.aux_f32);
            float4 e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (EPI == EPI_BIAS_RESID_F32 && pre_resid) e[i] = pre_resid[i];      // prefetched by the caller (see prefetch_resid_block)
                else e[i] = (i * 4 + rsub < rows) ? *reinterpret_cast<const float4*>(src + base + size_t(i * 4) * p.ldo) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (EPI == EPI_DGELU_F32) {
                    v[i].x *= quick_gelu_grad<true>(e[i].x); v[i].y *= quick_gelu_grad<true>(e[i].y);
                    v[i].z *= quick_gelu_grad<true>(e[i].z); v[i].w *= quick_gelu_grad<true>(e[i].w);
                } else {
                    v[i] = f4_add(v[i], e[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i * 4 + rsub >= rows) continue;
            const size_t off = base + size_t(i * 4) * p.ldo;
            if (EPI == EPI_BIAS_GELU_F32) {
                if (p.out2_f32) *reinterpret_cast<float4*>(p.out2_f32 + off) = v[i];
                v[i] = make_float4(quick_gelu<true>(v[i].x), quick_gelu<true>(v[i].y), quick_gelu<true>(v[i].z), quick_gelu<true>(v[i].w));
            }
            *reinterpret_cast<float4*>(p.out_f32 + off) = v[i];
            if ((EPI == EPI_BIAS_RESID_F32 || EPI == EPI_ACC_F32) && p.out_bf16)
                *reinterpret_cast<uint2*>(p.out_bf16 + off) = make_uint2(pack_h2<F16>(v[i].x, v[i].y), pack_h2<F16>(v[i].z, v[i].w));
        }
    } else {
        if (EPI == EPI_DGELU_BF16) {
            uint2 e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (pre_aux) e[i] = pre_aux[i];              // prefetched by the caller while the main loop was still running
                else e[i] = (i * 4 + rsub < rows) ? *reinterpret_cast<const uint2*>(p.aux_bf16 + base + size_t(i * 4) * p.ldo) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (LPI_H2_ACT && F16 && !bool(LPI_F16_PRECISE_ACT)) {                  // packed-half derivative (see h2_quick_gelu_grad)
                    const float2 d0 = h2_quick_gelu_grad(e[i].x), d1 = h2_quick_gelu_grad(e[i].y);
                    v[i].x *= d0.x; v[i].y *= d0.y; v[i].z *= d1.x; v[i].w *= d1.y;
                    continue;
                }
                const float2 z0 = unpack_h2<F16>(e[i].x);
                const float2 z1 = unpack_h2<F16>(e[i].y);
                if (F16 && bool(LPI_F16_PRECISE_ACT)) {
                    v[i].x *= quick_gelu_grad<true>(z0.x); v[i].y *= quick_gelu_grad<true>(z0.y);
                    v[i].z *= quick_gelu_grad<true>(z1.x); v[i].w *= quick_gelu_grad<true>(z1.y);
                } else {
                    v[i].x *= quick_gelu_grad<false>(z0.x); v[i].y *= quick_gelu_grad<false>(z0.y);
                    v[i].z *= quick_gelu_grad<false>(z1.x); v[i].w *= quick_gelu_grad<false>(z1.y);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i * 4 + rsub >= rows) continue;
            const size_t off = base + size_t(i * 4) * p.ldo;
            if (EPI == EPI_BIAS_GELU_BF16) {
                if (p.out2_bf16) *reinterpret_cast<uint2*>(p.out2_bf16 + off) = make_uint2(pack_h2<F16>(v[i].x, v[i].y), pack_h2<F16>(v[i].z, v[i].w));
                if (F16 && bool(LPI_F16_PRECISE_ACT)) v[i] = make_float4(quick_gelu<true>(v[i].x), quick_gelu<true>(v[i].y), quick_gelu<true>(v[i].z), quick_gelu<true>(v[i].w));
                else v[i] = make_float4(quick_gelu<false>(v[i].x), quick_gelu<false>(v[i].y), quick_gelu<false>(v[i].z), quick_gelu<false>(v[i].w));
            }
            *reinterpret_cast<uint2*>(p.out_bf16 + off) = make_uint2(pack_h2<F16>(v[i].x, v[i].y), pack_h2<F16>(v[i].z, v[i].w));
        }
    }
}

// 16-bit-output epilogues without an aux / residual operand (EPI_BIAS_BF16, EPI_BF16, EPI_BIAS_GELU_BF16): bias and activation are applied
// in the accumulator layout (thread = row; the bias read is a warp-wide broadcast) and the block is transposed through shared memory as
// 16-bit values -- 2 KB per 32x32 block instead of 4 KB of fp32, i.e. half the smem traffic the probe in DESIGN.md section 8 points at.
// Row r of the block occupies 64 B = four 16-byte chunks, chunk q stored at slot q ^ ((r >> 1) & 3): both directions are conflict-free.
template <int EPI, bool F16>
__device__ __forceinline__ void epilogue_block16(const GemmArgs& p, const uint32_t (&r)[32], uint4* __restrict__ st, int row0, int col0, int lane) {
    const int rows = min(32, p.M - row0);                      // warp-uniform
    if (rows <= 0) return;
    constexpr bool kBias = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16);
    float v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kBias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j));
        v[4 * j] = __uint_as_float(r[4 * j]) + b4.x;
        v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
        v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z;
        v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
    }
    const int sw = (lane >> 1) & 3;
    // `streaming`: the saved pre-activation is read again a whole forward and half a backward later -- it goes out with an evict-first
    // hint so that it does not push the activation the next GEMM is about to read out of the 126 MB L2
    auto stage_and_store = [&](__nv_bfloat16* dst, bool streaming = false) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            st[lane * 4 + (q ^ sw)] = make_uint4(pack_h2<F16>(v[8 * q], v[8 * q + 1]), pack_h2<F16>(v[8 * q + 2], v[8 * q + 3]),
                                                 pack_h2<F16>(v[8 * q + 4], v[8 * q + 5]), pack_h2<F16>(v[8 * q + 6], v[8 * q + 7]));
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int rr = it * 8 + (lane >> 2), q = lane & 3;
            const uint4 w = st[rr * 4 + (q ^ ((rr >> 1) & 3))];
            if (rr < rows) {
                uint4* g = reinterpret_cast<uint4*>(dst + size_t(row0 + rr) * p.ldo + col0 + q * 8);
                if (streaming) __stcs(g, w); else *g = w;
            }
        }
        __syncwarp();
    };
    if (LPI_H2_ACT && EPI == EPI_BIAS_GELU_BF16 && F16 && !bool(LPI_F16_PRECISE_ACT)) {
        // fp16 tower: pack the pre-activation once, store it, then the activation on the packed halves
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = pack_h2<true>(v[2 * j], v[2 * j + 1]);
        auto stage_words = [&](__nv_bfloat16* dst, bool streaming = false) {
#pragma unroll
            for (int q = 0; q < 4; ++q) st[lane * 4 + (q ^ sw)] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int rr = it * 8 + (lane >> 2), q = lane & 3;
                const uint4 x = st[rr * 4 + (q ^ ((rr >> 1) & 3))];
                if (rr < rows) {
                    uint4* g = reinterpret_cast<uint4*>(dst + size_t(row0 + rr) * p.ldo + col0 + q * 8);
                    if (streaming) __stcs(g, x); else *g = x;
                }
            }
            __syncwarp();
        };
        if (p.out2_bf16) stage_words(p.out2_bf16, true);
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = h2_quick_gelu(w[j]);
        stage_words(p.out_bf16);
        return;
    }
    if (EPI == EPI_BIAS_GELU_BF16) {
        if (p.out2_bf16) stage_and_store(p.out2_bf16, true);    // pre-activation, kept for the backward
        if (F16 && bool(LPI_F16_PRECISE_ACT)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = quick_gelu<true>(v[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = quick_gelu<false>(v[j]);
        }
    }
    stage_and_store(p.out_bf16);
}

// The saved pre-activation read by the dGELU epilogue was written a whole forward pass earlier, i.e. it comes from DRAM: loaded
// inside epilogue_block each 32-column block paid ~1.5 k cycles of exposed latency (4 blocks per tile > the 6 k-cycle main loop of a
// K = 768 tile).  The pair kernel therefore issues these loads two blocks ahead -- the first two before it even waits for the
// accumulator -- into registers that are indexed with compile-time constants only.
__device__ __forceinline__ void prefetch_aux_block(const GemmArgs& p, int row0, int col0, int lane, uint2 (&e)[8]) {
    const int rows = min(32, p.M - row0);
    const int rsub = lane >> 3, c = lane & 7;
    const size_t base = size_t(row0 + rsub) * p.ldo + col0 + c * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#if LPI_AUX_LDCS
        e[i] = (i * 4 + rsub < rows) ? __ldcs(reinterpret_cast<const uint2*>(p.aux_bf16 + base + size_t(i * 4) * p.ldo)) : make_uint2(0u, 0u);   // read once: evict-first
#else
        e[i] = (i * 4 + rsub < rows) ? __ldg(reinterpret_cast<const uint2*>(p.aux_bf16 + base + size_t(i * 4) * p.ldo)) : make_uint2(0u, 0u);
#endif
}

// Same idea for the fp32 residual of EPI_BIAS_RESID_F32 (out-proj / c_proj forward): the residual stream was written one or two kernels
// earlier (L2 at best), and loading it inside epilogue_block exposed that latency once per 32-column block.  (The residual never aliases
// a row another CTA writes, and this thread's own store of a block comes after its load of the same block, so in-place use stays legal.)
__device__ __forceinline__ void prefetch_resid_block(const GemmArgs& p, int row0, int col0, int lane, float4 (&e)[8]) {
    const int rows = min(32, p.M - row0);
    const int rsub = lane >> 3, c = lane & 7;
    const size_t base = size_t(row0 + rsub) * p.ldo + col0 + c * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        e[i] = (i * 4 + rsub < rows) ? *reinterpret_cast<const float4*>(p.resid + base + size_t(i * 4) * p.ldo) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------------ direct epilogue (LPI_EPI_DIRECT)
// The smem-staged epilogues above compete with the MMA main loop for shared-memory bandwidth: a streaming 256 x 256 CTA-pair tile
// already fills and reads ~128 B/clk, and with the epilogue removed the K = 768 GEMMs run at 1.36-1.54 PFLOP/s instead of 0.89-1.18
// (profiles/r2_gemm_vs_cublas.md).  sm_100 has 256-bit global accesses: in the accumulator layout (thread = row) a thread's 32 columns are
// 64 contiguous bytes (16-bit) or 128 (fp32), i.e. 2 or 4 full 32-byte sectors, so rows can go to and come from global memory directly
// -- no staging, no __syncwarp, every sector written whole.
__device__ __forceinline__ void ldg256(const void* ptr, uint32_t* a) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]) : "l"(ptr));
}
__device__ __forceinline__ void stg256(void* ptr, const uint32_t* a) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(ptr), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]) : "memory");
}
// saved pre-activation (16-bit, EPI_DGELU_BF16) / fp32 residual (EPI_BIAS_RESID_F32) of one row x 32 columns, issued ahead of use
__device__ __forceinline__ void prefetch_aux_direct(const GemmArgs& p, int row, int col0, bool valid, uint32_t (&e)[16]) {
    if (valid) {
        const __nv_bfloat16* src = p.aux_bf16 + size_t(row) * p.ldo + col0;
        ldg256(src, e);
        ldg256(src + 16, e + 8);
    }
}
__device__ __forceinline__ void prefetch_resid_direct(const GemmArgs& p, int row, int col0, bool valid, uint32_t (&e)[32]) {
    if (valid) {
        const float* src = p.resid + size_t(row) * p.ldo + col0;
#pragma unroll
        for (int j = 0; j < 4; ++j) ldg256(src + 8 * j, e + 8 * j);
    }
}

template <int EPI, bool F16>
__device__ __forceinline__ void epilogue_direct(const GemmArgs& p, const uint32_t (&r)[32], int row, int col0, bool valid, const uint32_t* pre) {
    constexpr bool kBias = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_F32 ||
                            EPI == EPI_BIAS_GELU_F32);
    constexpr bool kF32Out = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_F32 || EPI == EPI_ACC_F32 || EPI == EPI_BIAS_F32 ||
                              EPI == EPI_BIAS_GELU_F32 || EPI == EPI_DGELU_F32);
    if (!valid) return;
    float v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kBias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j));      // same address in every lane: one broadcast
        v[4 * j] = __uint_as_float(r[4 * j]) + b4.x;
        v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
        v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z;
        v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
    }
    const size_t off = size_t(row) * p.ldo + col0;
    auto store16 = [&](__nv_bfloat16* dst) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = pack_h2<F16>(v[2 * j], v[2 * j + 1]);
        stg256(dst + off, w);
        stg256(dst + off + 16, w + 8);
    };
    auto store32 = [&](float* dst) {
        uint32_t w[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(v[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) stg256(dst + off + 8 * j, w + 8 * j);
    };
    if (!kF32Out) {
        if (EPI == EPI_DGELU_BF16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (LPI_H2_ACT && F16 && !bool(LPI_F16_PRECISE_ACT)) {
                    const float2 d = h2_quick_gelu_grad(pre[j]);
                    v[2 * j] *= d.x; v[2 * j + 1] *= d.y;
                } else {
                    const float2 z = unpack_h2<F16>(pre[j]);
                    if (F16 && bool(LPI_F16_PRECISE_ACT)) { v[2 * j] *= quick_gelu_grad<true>(z.x); v[2 * j + 1] *= quick_gelu_grad<true>(z.y); }
                    else { v[2 * j] *= quick_gelu_grad<false>(z.x); v[2 * j + 1] *= quick_gelu_grad<false>(z.y); }
                }
            }
        }
        if (EPI == EPI_BIAS_GELU_BF16) {
            if (p.out2_bf16) store16(p.out2_bf16);                  // pre-activation, kept for the backward
            if (F16 && bool(LPI_F16_PRECISE_ACT)) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = quick_gelu<true>(v[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = quick_gelu<false>(v[j]);
            }
        }
        store16(p.out_bf16);
    } else {
        if (EPI == EPI_BIAS_RESID_F32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(pre[j]);
        } else if (EPI == EPI_ACC_F32 || EPI == EPI_DGELU_F32) {
            const float* src = (EPI == EPI_ACC_F32 ? p.out_f32 : p.aux_f32) + off;
            uint32_t e[32];
#pragma unroll
            for (int j = 0; j < 4; ++j) ldg256(src + 8 * j, e + 8 * j);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (EPI == EPI_ACC_F32) v[j] += __uint_as_float(e[j]);
                else v[j] *= quick_gelu_grad<true>(__uint_as_float(e[j]));
            }
        }
        if (EPI == EPI_BIAS_GELU_F32) {
            if (p.out2_f32) store32(p.out2_f32);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = quick_gelu<true>(v[j]);
        }
        store32(p.out_f32);
        if ((EPI == EPI_BIAS_RESID_F32 || EPI == EPI_ACC_F32) && p.out_bf16) store16(p.out_bf16);
    }
}

// ------------------------------------------------------------------------------------------------ running top-k (MODE_TOPK)
// One epilogue thread owns one query row; its sorted list (score desc, gallery index asc) lives in shared memory,
// column-major [j][row].  Insertion is rare after warm-up (~k ln(N/k) per row over the whole gallery) and stays out of line,
// so the hot path is: 64 accumulator columns -> max tree -> one compare.  The accumulator registers are only ever indexed
// with compile-time constants (a dynamic r[j] forces the whole array into local memory: 512 B of STL per thread per chunk,
// which showed up as 60 % L1 throughput in the first capture).
struct TopkState {
    float* sc;      // [TOPK_MAX][BM]
    int* id;
    int row;        // r_local
    int k;
    int cnt;
    float thr;
};
struct TopkCT { int cnt; float thr; };      // returned by value so both stay in registers across the out-of-line call

__device__ __noinline__ TopkCT topk_insert(float* sc, int* id, int row, int k, int cnt, float thr_in, float v, int gidx) {
    int pos = cnt < k ? cnt : k - 1;
    while (pos > 0 && sc[(pos - 1) * BM + row] < v) {            // strict: an equal score with a larger index never displaces
        sc[pos * BM + row] = sc[(pos - 1) * BM + row];
        id[pos * BM + row] = id[(pos - 1) * BM + row];
        --pos;
    }
    sc[pos * BM + row] = v;
    id[pos * BM + row] = gidx;
    if (cnt < k) ++cnt;
    return {cnt, cnt == k ? sc[(k - 1) * BM + row] : thr_in};    // a seeded threshold stays in force until the list is full
}

// largest float below x (x finite or -inf): "v > float_prev(s)" == "v >= s", so a seed taken from k sampled gallery rows admits them again
__device__ __forceinline__ float float_prev(float x) {
    if (!(x > -INFINITY)) return -INFINITY;
    const uint32_t b = __float_as_uint(x);
    if (x > 0.f) return __uint_as_float(b - 1);
    if (x == 0.f) return __uint_as_float(0x80000001u);
    return __uint_as_float(b + 1);
}

// Cooperative thresholds.  A gallery shard is cut into chunks for load balance, and every (query tile, chunk) item used to warm its
// top-k lists up alone: with the exact final thresholds handed in the 625 k-row shard of the 8-GPU run takes 10.7 ms, with the seeded
// ones 12.2 ms (profiles/r2_scorer_fixed_cost.txt) -- the difference is sorted insertions of candidates that another chunk's list
// already rules out.  So the items of one launch share one fp32 per query row in global memory: a row publishes its own k-th score
// whenever that improves (k gallery rows reach it, so it bounds the final k-th score from below) and polls the shared value once per
// gallery tile.  Whatever an item then drops is below a score k rows of the union reach, so the MERGED result is unchanged; the
// per-chunk lists themselves depend on timing and may hold fewer than k entries (rest: -inf / INT_MAX), as with init_thr.
__device__ __forceinline__ void coop_publish(float* addr, float v) {      // atomic max on fp32 through the integer ordering
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ float coop_poll(const float* addr) { return __ldcg(addr); }      // L2: never a stale L1 line

// maximum of one 64-column accumulator block (ragged tail masked) -- the whole epilogue of the threshold pre-pass
template <int NCOL>
__device__ __forceinline__ float block_max(uint32_t (&r)[NCOL], int col0, int n_valid) {
    if (col0 + NCOL > n_valid) {
#pragma unroll
        for (int j = 0; j < NCOL; ++j)
            if (col0 + j >= n_valid) r[j] = 0xff800000u;
    }
    float m[4] = {__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3])};
#pragma unroll
    for (int j = 4; j < NCOL; j += 4) {
        m[0] = fmaxf(m[0], __uint_as_float(r[j]));
        m[1] = fmaxf(m[1], __uint_as_float(r[j + 1]));
        m[2] = fmaxf(m[2], __uint_as_float(r[j + 2]));
        m[3] = fmaxf(m[3], __uint_as_float(r[j + 3]));
    }
    return fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
}

// The check is hierarchical: one compare on the block maximum (hot path), then one per 16-column group, then per element.  During the
// warm-up of a list some row of the warp passes in most blocks, and the whole warp pays for the slow path, so its cost matters:
// at 8 GPUs (625 k rows per rank) the flat 64-element scan cost 3.4 ms of a 15.5 ms launch (profiles/r1_scorer_seed.md).
template <int NCOL>
__device__ __forceinline__ void topk_consume(TopkState& st, uint32_t (&r)[NCOL], int col0, int n_valid, int gidx0) {
    static_assert(NCOL % 16 == 0, "group scan");
    if (col0 + NCOL > n_valid) {                  // ragged gallery tail: rows past N were zero-filled by TMA
#pragma unroll
        for (int j = 0; j < NCOL; ++j)
            if (col0 + j >= n_valid) r[j] = 0xff800000u;   // -inf
    }
    float gmax[NCOL / 16];
#pragma unroll
    for (int g = 0; g < NCOL / 16; ++g) {         // independent max chains, one per group
        float m0 = __uint_as_float(r[16 * g]), m1 = __uint_as_float(r[16 * g + 1]);
#pragma unroll
        for (int j = 2; j < 16; j += 2) {
            m0 = fmaxf(m0, __uint_as_float(r[16 * g + j]));
            m1 = fmaxf(m1, __uint_as_float(r[16 * g + j + 1]));
        }
        gmax[g] = fmaxf(m0, m1);
    }
    float bmax = gmax[0];
#pragma unroll
    for (int g = 1; g < NCOL / 16; ++g) bmax = fmaxf(bmax, gmax[g]);
    if (bmax > st.thr) {
#pragma unroll
        for (int g = 0; g < NCOL / 16; ++g) {
            if (gmax[g] > st.thr) {
#pragma unroll
                for (int j = 16 * g; j < 16 * g + 16; ++j) {
                    const float v = __uint_as_float(r[j]);
                    if (v > st.thr) {
                        const TopkCT ct = topk_insert(st.sc, st.id, st.row, st.k, st.cnt, st.thr, v, gidx0 + col0 + j);
                        st.cnt = ct.cnt;
                        st.thr = ct.thr;
                    }
                }
            }
        }
    }
}

// TF32 = true: fp32 operands in memory (32 elements = 128 B per swizzle row), tcgen05.mma kind::tf32 (K = 8 per instruction)
template <int MODE, int BN, int EPI, bool TF32 = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + C::BAR_OFF;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + C::BAR_OFF + 8 * (2 * C::STAGES + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int BKE = TF32 ? 32 : BK;          // elements per 128-byte smem row
    const int num_k = p.K / BKE;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < C::STAGES; ++s) {
                mbar_init(full_bar(s), 1);
                mbar_init(empty_bar(s), 1);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(tfull_bar(a), 1);
                mbar_init(tempty_bar(a), 4);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<C::TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            TileWalk<MODE, BN> walk(p);
            int stage = 0;
            uint32_t phase = 0;
            int m0, n0;
            bool first, last;
            while (walk.next(m0, n0, first, last)) {
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                    tma_load_2d(sa, &tmA, full_bar(stage), kb * BKE, m0);
                    tma_load_2d(sa + A_STAGE_BYTES, &tmB, full_bar(stage), kb * BKE, n0);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (single thread)
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(TF32 ? kFmtTF32 : kFmtBF16, BM, BN, 0, 0);
            TileWalk<MODE, BN> walk(p);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            int m0, n0;
            bool first, last;
            while (walk.next(m0, n0, first, last)) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                    const uint64_t da = make_desc_kmajor_sw128(sa);
                    const uint64_t db = make_desc_kmajor_sw128(sa + A_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) { // +32 B per K step inside the 128 B swizzle row (16 bf16 or 8 tf32)
                        if (TF32) umma_tf32_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        else umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(empty_bar(stage));           // smem slot reusable once these MMAs retire
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar(acc));                 // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps (2..5)
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may touch
        const int r_local = quad * 32 + lane;      // accumulator row owned by this thread
        TileWalk<MODE, BN> walk(p);
        int acc = 0;
        uint32_t acc_phase = 0;
        int m0, n0;
        bool first, last;
        TopkState tk{reinterpret_cast<float*>(smem_gen + C::LIST_OFF), reinterpret_cast<int*>(smem_gen + C::LIST_OFF + BM * TOPK_MAX * 4),
                     r_local, p.k, 0, -INFINITY};
        while (walk.next(m0, n0, first, last)) {
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const int row = m0 + r_local;
            if (MODE == MODE_TOPK && first) { tk.thr = (p.init_thr && row < p.M) ? float_prev(p.init_thr[size_t(row) * p.init_thr_stride]) : -INFINITY; tk.cnt = 0; }
            if (MODE == MODE_GEMM) {
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + uint32_t(acc * BN + c * 32) + (uint32_t(quad * 32) << 16);
                    LPI_TMEM_LD_X32(taddr, r);
                    tmem_ld_wait();
                    float4* stg = reinterpret_cast<float4*>(smem_gen + C::LIST_OFF) + (warp - 2) * 32 * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j)          // own row, 16-byte chunk j -> swizzled slot
                        stg[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                                       __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                    __syncwarp();
                    epilogue_block<EPI>(p, stg, m0 + quad * 32, n0 + c * 32, lane);
                    __syncwarp();
                }
            } else {
                float tmax = -INFINITY;
#pragma unroll 1
                for (int c = 0; c < BN / 64; ++c) {
                    uint32_t r[64];
                    const uint32_t taddr = tmem_base + uint32_t(acc * BN + c * 64) + (uint32_t(quad * 32) << 16);
                    LPI_TMEM_LD_X64(taddr, r);
                    tmem_ld_wait();
                    if (p.seed_mode) tmax = fmaxf(tmax, block_max<64>(r, n0 + c * 64, p.N));
                    else topk_consume<64>(tk, r, n0 + c * 64, p.N, int(p.gallery_offset));
                }
                if (p.seed_mode && tmax > tk.thr) {          // one candidate per tile: its maximum
                    const TopkCT ct = topk_insert(tk.sc, tk.id, tk.row, tk.k, tk.cnt, tk.thr, tmax, n0);
                    tk.cnt = ct.cnt;
                    tk.thr = ct.thr;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            if (MODE == MODE_TOPK && last) {
                // item (q_tile, chunk) finished: m0 identifies q_tile; the chunk is n0's
                const int chunk = (n0 / BN) / p.tiles_per_chunk;
                if (row < p.M) {
                    float* os = p.topk_scores + (size_t(chunk) * p.M + row) * p.k;
                    int* oi = p.topk_idx + (size_t(chunk) * p.M + row) * p.k;
                    for (int j = 0; j < p.k; ++j) {
                        os[j] = j < tk.cnt ? tk.sc[j * BM + r_local] : -INFINITY;
                        oi[j] = j < tk.cnt ? tk.id[j * BM + r_local] : 0x7fffffff;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

// ================================================================================================================
// CTA-pair variant (cta_group::2): two CTAs on the two SMs of a TPC execute ONE tcgen05.mma with M = 256, N = 256.
// Each CTA holds its own 128 rows of A and HALF of the B tile (128 of the 256 N rows), so per CTA the L2 -> smem traffic of
// a 128 x 256 x K output block drops from (128 + 256) K to (128 + 128) K elements -- the 1-CTA kernel is bound by that
// traffic (profiles/r1_scorer_v1_ncu.md: ~9 TB/s of TMA at 729 TFLOP/s).  In MODE_TOPK the query tile additionally stays
// RESIDENT in shared memory for the whole gallery chunk (128 KB for d = 512), leaving only the half gallery tile to stream:
// 128 KB per 33.5 MFLOP per CTA = 262 FLOP/B.
//
// Protocol (CUTLASS-style 2SM pipeline): both CTAs issue their own TMA loads with .cta_group::2 so the bytes complete on
// the LEADER's (cluster rank 0) full barrier; the leader's single MMA thread issues the MMAs, whose commits are multicast to
// the empty / accumulator-full barriers of BOTH CTAs; every CTA's epilogue warps drain their own TMEM lanes (their 128 rows)
// and release the accumulator on the leader's barrier (remote arrive for the peer).
// ================================================================================================================
// BN_ = cluster tile width: 256, or 192 / 128 for the GEMM shapes whose tile count would otherwise leave a mostly empty last wave
// (N = 768 at B = 64: 54 x 3 tiles of 256 = 2.19 waves of 74 clusters; 54 x 4 tiles of 192 = 2.92)
template <int MODE, int BN_ = 256>
struct PairCfg {
    static constexpr int BN = BN_;
    static constexpr int A_BYTES = BM * 128;                 // one k-block of A: 128 rows x 128 B
    static constexpr int BH_BYTES = (BN / 2) * 128;          // half B tile: BN/2 rows x 128 B
    static constexpr int A_RESIDENT_KB = 8;                  // MODE_TOPK: up to 8 k-blocks (dim <= 512) stay resident
    static constexpr int STAGE_BYTES = (MODE == MODE_GEMM) ? A_BYTES + BH_BYTES : BH_BYTES;
    static constexpr int STAGES = (MODE == MODE_GEMM) ? 6 : 5;
    static constexpr int RING_OFF = (MODE == MODE_GEMM) ? 0 : A_RESIDENT_KB * A_BYTES;
    static constexpr int BAR_OFF = RING_OFF + STAGES * STAGE_BYTES;
    static constexpr int LIST_OFF = BAR_OFF + 256;
    // MODE_GEMM runs EIGHT epilogue warps (two per TMEM lane quadrant, each draining half of the 256 accumulator columns): with four,
    // the GELU / dGELU / residual epilogues of the K = 768 GEMMs took ~2x their 48-MMA main loop (95 us for a 46 us GEMM)
    static constexpr int EPI_WARPS = (MODE == MODE_GEMM) ? 8 : 4;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int CLC_SLOTS = 3;                      // response ring of the dynamic (cluster-launch-control) tile scheduler
    static constexpr int TAIL = (MODE == MODE_GEMM) ? EPI_WARPS * 32 * 32 * 4 : BM * TOPK_MAX * 8;
    static constexpr int SMEM_BYTES = LIST_OFF + TAIL + 1024;
    static constexpr int TMEM_COLS = 512;
};

enum { OP_BF16 = 0, OP_TF32 = 1, OP_F16 = 2 };      // operand type of the CTA-pair GEMM

// CL = CTAs per cluster: 2 (one CTA pair), or 4 in MODE_GEMM = TWO pairs working on vertically adjacent 256-row blocks of the SAME N tile.
// The two pairs need the same weight (B) rows, so each of the four CTAs fetches only a QUARTER of the B tile and multicasts it to its
// counterpart in the other pair: per CTA and k-block 16 KB (A) + 8 KB (B) leave L2 instead of 16 + 16.  The encoder GEMMs are bound by
// exactly that traffic: their time follows (operand bytes read from L2 + 2 x epilogue bytes) / 12.5 TB/s on every shape
// (profiles/r2_gemm_vs_cublas.md).  Protocol changes against CL = 2: a stage is free when the MMAs of BOTH pairs have read it (empty
// barriers count 2 commits, multicast to all four CTAs); every pair leader still expects the bytes of its own pair's stage.
template <int MODE, int EPI, int OP, int BN_ = 256, int CL = 2>
__global__ void __launch_bounds__(PairCfg<MODE, BN_>::THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs p) {
    static_assert(CL == 2 || (CL == 4 && MODE == MODE_GEMM), "4-CTA clusters exist for the GEMM mode only");
    using C = PairCfg<MODE, BN_>;
    constexpr int NPAIR = CL / 2;
    constexpr bool TF32 = OP == OP_TF32;
    constexpr bool F16 = OP == OP_F16;
    constexpr int BN = C::BN;
    constexpr int BKE = TF32 ? 32 : BK;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // identical in both CTAs
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + C::BAR_OFF;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
    const uint32_t afull_bar = bar_base + 8u * (2 * C::STAGES + 4);       // MODE_TOPK: resident query tile landed / may be overwritten
    const uint32_t aempty_bar = bar_base + 8u * (2 * C::STAGES + 5);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + C::BAR_OFF + 8 * (2 * C::STAGES + 6));
    // dynamic tile scheduling (p.clc, MODE_GEMM with CTA pairs): a ring of CLC_SLOTS responses, each with a full barrier in every CTA
    // (16 response bytes) and an empty barrier that lives in the leader CTA (all 19 consumers of the cluster arrive on it)
    auto clc_full = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 7 + s); };
    auto clc_empty = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 7 + C::CLC_SLOTS + s); };
    auto clc_resp = [&](int s) { return bar_base + 208u + 16u * s; };
    static_assert(8 * (2 * C::STAGES + 7 + 2 * C::CLC_SLOTS) <= 208 && 208 + 16 * C::CLC_SLOTS <= 256, "barrier region is 256 bytes");
    constexpr int CLC_CONSUMERS = 2 * (1 + C::EPI_WARPS) + 1;     // TMA thread + epilogue warps of both CTAs, the leader's MMA thread
    const bool clc = (MODE == MODE_GEMM) && (CL == 2) && p.clc != 0;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const uint32_t rank = crank & 1;                 // rank inside the CTA pair
    const uint32_t pair = crank >> 1;                // pair inside the cluster (always 0 for CL = 2)
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const uint16_t pair_mask = uint16_t(3u << (2 * pair)), all_mask = uint16_t((1u << CL) - 1);
    const int num_k = p.K / BKE;
    const int num_m = (p.M + BM - 1) / BM, num_mp = (num_m + 1) / 2, num_n = (p.N + BN - 1) / BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < C::STAGES; ++s) {
                mbar_init(full_bar(s), 1);       // leader's own arrive.expect_tx (bytes of both CTAs)
                mbar_init(empty_bar(s), NPAIR);  // multicast commit of every pair whose MMAs read (a multicast copy of) this stage
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(tfull_bar(a), 1);      // multicast commit
                mbar_init(tempty_bar(a), 2 * C::EPI_WARPS);     // every epilogue warp of both CTAs (only the leader's copy is used)
            }
            mbar_init(afull_bar, 1);
            mbar_init(aempty_bar, 1);
            if (MODE == MODE_GEMM) {
                for (int s2 = 0; s2 < C::CLC_SLOTS; ++s2) {
                    mbar_init(clc_full(s2), 1);                 // the scheduler's arrive.expect_tx (16 response bytes)
                    mbar_init(clc_empty(s2), CLC_CONSUMERS);
                }
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc_pair<C::TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                          // peer barriers initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (MODE == MODE_GEMM) {
        // PDL (lpi::launch_pdl): the set-up above ran while the previous kernel of the stream was finishing; its outputs (A, resid, aux)
        // are touched only after this point.  The dependents of THIS kernel may be scheduled as its CTAs retire.
        pdl_launch_dependents();
        pdl_wait();
    }

    // ---- work decomposition: a pure function of (cluster_id, rank), re-derived by every role
    // GEMM : cluster tile t -> (nt = t % num_n, mp = t / num_n), N fastest: the clusters running at the same time cover all N tiles of
    //        a few row blocks, so a block of the LARGE operand (A: up to 84 MB of activations) is fetched from HBM once and hit in L2 by
    //        its other N tiles; the weights (B, <= 4.7 MB) stay L2-resident either way.  (M fastest re-streamed A once per N tile:
    //        210 MB of DRAM reads for the 84 MB operand of the K = 3072 dgrad, profiles/r2_gemm_vs_cublas.md.)  This CTA owns rows
    //        (2 mp + rank) * 128
    // TOPK : item i -> (qp = i % num_mp, chunk = i / num_mp); gallery tiles [chunk * tpc, min(num_n, (chunk + 1) * tpc))
    //        (CL = 4: cluster tile t -> (nt = t % num_n, mq = t / num_n), pair q of the cluster owns row-block pair mp = 2 mq + q)
    const int total = (MODE == MODE_GEMM) ? ((num_mp + NPAIR - 1) / NPAIR) * num_n : num_mp * p.n_chunks;
    // TOPK item order stays chunk-major (concurrent clusters stream the same gallery region) with cooperative thresholds too: the
    // chunk-fastest order, which lets the chunks of one query tile warm each other up from the first wave on, measured 7 % SLOWER
    // (625 k rows: 12.10 vs 11.34 ms; the exact-threshold run slows down as well, i.e. it is the lost L2 locality of the gallery stream)
    constexpr bool topk_qmajor = false;
    // Tile sequence of this cluster.  Static mode: t = cluster_id, + n_clusters, ...  Dynamic mode (clc): the grid holds one cluster per
    // tile; a cluster starts with its own tile and then takes over the tiles of clusters that have not been launched yet, in launch
    // order, until none is left -- late starters (SMs still busy with the previous kernel or with the other tower's stream) simply
    // take fewer tiles instead of holding the whole kernel back.  Every role walks the same response ring.
    const uint32_t clc_empty_leader_base = mapa_u32(clc_empty(0), 0);
    auto next_tile = [&](int t, int& slot, uint32_t& ph, bool release) -> int {
        if (!clc) return t + n_clusters;
        mbar_wait(clc_full(slot), ph);
        const int x = clc_decode(clc_resp(slot));
        fence_proxy_async_smem();                    // this read is ordered before the asynchronous rewrite of the slot
        if (release) mbar_arrive_cluster(clc_empty_leader_base + 8u * slot);
        if (++slot == C::CLC_SLOTS) { slot = 0; ph ^= 1; }
        return x < 0 ? total : x / CL;
    };

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (one thread per CTA)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, aphase = 0;
            int cslot = 0, qslot = 0;
            uint32_t cph = 0, qph = 0;
            const uint32_t clc_full_peer = mapa_u32(clc_full(0), 1);
            // dynamic mode: the leader's TMA thread is the tile scheduler.  ONE cancellation request per tile, issued when the last ring
            // of k-blocks of the current tile is about to be requested: late enough that a cluster never hoards tiles it will only reach
            // much later (with 2.9 tiles per cluster, three requests in flight per cluster cost up to one tile time of imbalance), early
            // enough that the response (a round trip to the work distributor) is back when the loads of this tile have been issued.
            auto clc_request = [&]() {
                mbar_wait(clc_empty(qslot), qph ^ 1);           // every consumer of the cluster has read the slot's previous response
                mbar_arrive_expect_tx(clc_full(qslot), 16);
                mbar_arrive_expect_tx_cluster(clc_full_peer + 8u * qslot, 16);
                clc_try_cancel_multicast(clc_resp(qslot), clc_full(qslot));
                if (++qslot == C::CLC_SLOTS) { qslot = 0; qph ^= 1; }
            };
            const int kb_request = num_k > C::STAGES ? num_k - C::STAGES : 0;
            for (int t = cluster_id; t < total; t = next_tile(t, cslot, cph, true)) {
                const int mp = (MODE == MODE_GEMM) ? (CL == 4 ? NPAIR * (t / num_n) + int(pair) : (LPI_RASTER_N ? t / num_n : t % num_mp)) : (topk_qmajor ? t / p.n_chunks : t % num_mp);
                const int second = (MODE == MODE_GEMM) ? ((CL == 4 || LPI_RASTER_N) ? t % num_n : t / num_mp) : (topk_qmajor ? t % p.n_chunks : t / num_mp);
                const int m0 = (2 * mp + int(rank)) * BM;
                int n_begin, n_end;
                if (MODE == MODE_GEMM) { n_begin = second; n_end = second + 1; }
                else {
                    n_begin = second * p.tiles_per_chunk;
                    n_end = min(num_n, n_begin + p.tiles_per_chunk);
                    // resident query tile: all k-blocks once per item, after the previous item's MMAs retired
                    mbar_wait(aempty_bar, aphase ^ 1);
                    if (leader) mbar_arrive_expect_tx(afull_bar, 2u * num_k * C::A_BYTES);
                    for (int kb = 0; kb < num_k; ++kb) tma_load_2d_pair(smem_base + kb * C::A_BYTES, &tmA, afull_bar, kb * BKE, m0);
                    aphase ^= 1;
                }
                for (int nt = n_begin; nt < n_end; ++nt) {
                    const int n0 = nt * BN + int(rank) * (BN / 2);  // this CTA streams its half of the B tile
                    for (int kb = 0; kb < num_k; ++kb) {
                        if (MODE == MODE_GEMM && clc && leader && kb == kb_request) clc_request();
                        mbar_wait(empty_bar(stage), phase ^ 1);
#ifdef LPI_DEBUG_PROBE
                        const bool skip_a = (MODE == MODE_GEMM) && p.probe == 1;
#else
                        constexpr bool skip_a = false;
#endif
                        if (leader) mbar_arrive_expect_tx(full_bar(stage), skip_a ? 2u * C::BH_BYTES : 2u * C::STAGE_BYTES);
                        const uint32_t sa = smem_base + C::RING_OFF + stage * C::STAGE_BYTES;
                        if (MODE == MODE_GEMM) {
                            if (!skip_a) tma_load_2d_pair(sa, &tmA, full_bar(stage), kb * BKE, m0);
                            if (CL == 4) {
                                // this CTA's quarter of the B tile (half of the half its pair-rank holds), to itself and to the CTA of the
                                // same pair-rank in the other pair
                                tma_load_2d_pair_mc(sa + C::A_BYTES + pair * (C::BH_BYTES / 2), &tmB, full_bar(stage), kb * BKE,
                                                    n0 + int(pair) * (BN / 4), uint16_t(5u << rank));
                            } else {
                                tma_load_2d_pair(sa + C::A_BYTES, &tmB, full_bar(stage), kb * BKE, n0);
                            }
                        } else {
                            tma_load_2d_pair(sa, &tmB, full_bar(stage), kb * BKE, n0);
                        }
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA, single thread)
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc(TF32 ? kFmtTF32 : (F16 ? kFmtF16 : kFmtBF16), 2 * BM, BN, 0, 0);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, aphase = 0;
            int cslot = 0;
            uint32_t cph = 0;
            for (int t = cluster_id; t < total; t = next_tile(t, cslot, cph, true)) {
                const int second = (MODE == MODE_GEMM) ? ((CL == 4 || LPI_RASTER_N) ? t % num_n : t / num_mp) : (topk_qmajor ? t % p.n_chunks : t / num_mp);
                int n_begin, n_end;
                if (MODE == MODE_GEMM) { n_begin = second; n_end = second + 1; }
                else {
                    n_begin = second * p.tiles_per_chunk;
                    n_end = min(num_n, n_begin + p.tiles_per_chunk);
                    mbar_wait(afull_bar, aphase);
                    aphase ^= 1;
                    tc_fence_after();
                }
                for (int nt = n_begin; nt < n_end; ++nt) {
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * 256;
                    for (int kb = 0; kb < num_k; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + C::RING_OFF + stage * C::STAGE_BYTES;
                        const uint64_t da = make_desc_kmajor_sw128(MODE == MODE_GEMM ? sa : smem_base + kb * C::A_BYTES);
                        const uint64_t db = make_desc_kmajor_sw128(MODE == MODE_GEMM ? sa + C::A_BYTES : sa);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            if (TF32) umma_tf32_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                            else umma_f16_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        }
                        umma_commit_pair(empty_bar(stage), all_mask);        // the stage is reusable in every CTA it was multicast to
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit_pair(tfull_bar(acc), pair_mask);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                if (MODE == MODE_TOPK) umma_commit_pair(aempty_bar, pair_mask);     // every MMA reading the resident tile has retired
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps (2..) of BOTH CTAs: own 128 rows;
        // MODE_GEMM: warps 2..5 drain columns [0, 128), warps 6..9 columns [128, 256) of the same lane quadrants
        const int quad = warp & 3;
        const int col_half = (warp - 2) >> 2;                  // 0 for the first four epilogue warps
        const int r_local = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        TopkState tk{reinterpret_cast<float*>(smem_gen + C::LIST_OFF), reinterpret_cast<int*>(smem_gen + C::LIST_OFF + BM * TOPK_MAX * 4),
                     r_local, p.k, 0, -INFINITY};
        const uint32_t tempty_leader0 = mapa_u32(tempty_bar(0), 2 * pair), tempty_leader1 = mapa_u32(tempty_bar(1), 2 * pair);
        int cslot = 0;
        uint32_t cph = 0;
        // dynamic mode: every lane reads the response, lane 0 releases the slot once the whole warp has
        auto next_tile_warp = [&](int t) -> int {
            if (!clc) return t + n_clusters;
            const int nt = next_tile(t, cslot, cph, false);
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(clc_empty_leader_base + 8u * ((cslot + C::CLC_SLOTS - 1) % C::CLC_SLOTS));
            return nt;
        };
        for (int t = cluster_id; t < total; t = next_tile_warp(t)) {
            const int mp = (MODE == MODE_GEMM) ? (CL == 4 ? NPAIR * (t / num_n) + int(pair) : (LPI_RASTER_N ? t / num_n : t % num_mp)) : (topk_qmajor ? t / p.n_chunks : t % num_mp);
                const int second = (MODE == MODE_GEMM) ? ((CL == 4 || LPI_RASTER_N) ? t % num_n : t / num_mp) : (topk_qmajor ? t % p.n_chunks : t / num_mp);
            const int m0 = (2 * mp + int(rank)) * BM;
            const int row = m0 + r_local;
            int n_begin, n_end;
            if (MODE == MODE_GEMM) { n_begin = second; n_end = second + 1; }
            else { n_begin = second * p.tiles_per_chunk; n_end = min(num_n, n_begin + p.tiles_per_chunk); }
            tk.thr = (MODE == MODE_TOPK && p.init_thr && row < p.M) ? float_prev(p.init_thr[size_t(row) * p.init_thr_stride]) : -INFINITY;
            tk.cnt = 0;
            const bool coop = (MODE == MODE_TOPK) && p.shared_thr != nullptr && !p.seed_mode && row < p.M;
            float coop_floor = tk.thr, coop_pub = -INFINITY;   // best outside bound seen / own k-th score last published
            for (int nt = n_begin; nt < n_end; ++nt) {
                const int n0 = nt * BN;
                constexpr int NCH = BN / 64;                   // 32-column blocks per epilogue warp
                float coop_in = -INFINITY;
                if (coop) coop_in = coop_poll(p.shared_thr + row);     // in flight across the accumulator wait
                constexpr bool kPrefetchAux = (MODE == MODE_GEMM) && (EPI == EPI_DGELU_BF16) && !LPI_EPI_DIRECT;
                constexpr bool kPrefetchResid = (MODE == MODE_GEMM) && (EPI == EPI_BIAS_RESID_F32) && !LPI_EPI_DIRECT;
                // direct epilogue: this thread's row of the saved pre-activation / residual, two 32-column blocks ahead (the first two are
                // requested before the accumulator wait)
                constexpr int NPRE = (EPI == EPI_BIAS_RESID_F32) ? 32 : 16;
                uint32_t d0[NPRE], d1[NPRE];
                const bool valid = row < p.M;
                auto prefetch = [&](int c, uint32_t (&e)[NPRE]) {
                    if constexpr (EPI == EPI_DGELU_BF16) prefetch_aux_direct(p, row, n0 + c * 32, valid, e);
                    if constexpr (EPI == EPI_BIAS_RESID_F32) prefetch_resid_direct(p, row, n0 + c * 32, valid, e);
                };
                if (MODE == MODE_GEMM && LPI_EPI_DIRECT) {
                    prefetch(col_half * NCH + 0, d0);
                    if (NCH > 1) prefetch(col_half * NCH + 1, d1);
                }
                uint2 pre0[8], pre1[8];
                float4 rs0[8], rs1[8];
                if (kPrefetchAux) {
                    prefetch_aux_block(p, m0 + quad * 32, n0 + (col_half * NCH + 0) * 32, lane, pre0);
                    prefetch_aux_block(p, m0 + quad * 32, n0 + (col_half * NCH + 1) * 32, lane, pre1);
                }
                if (kPrefetchResid) {
                    prefetch_resid_block(p, m0 + quad * 32, n0 + (col_half * NCH + 0) * 32, lane, rs0);
                    prefetch_resid_block(p, m0 + quad * 32, n0 + (col_half * NCH + 1) * 32, lane, rs1);
                }
                // out_proj dgrad with the attention backward's delta fused (p.delta): this thread's row of the saved attention output O for
                // every 32-column block the warp drains, requested before the accumulator wait (one 64-byte segment per block)
                constexpr bool kDelta = (MODE == MODE_GEMM) && (EPI == EPI_BF16) && !TF32;
                uint32_t od[kDelta ? NCH : 1][16];
                if (kDelta && p.delta != nullptr) {
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) prefetch_aux_direct(p, row, n0 + (col_half * NCH + cc) * 32, valid, od[kDelta ? cc : 0]);
                }
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                if (MODE == MODE_GEMM && LPI_EPI_DIRECT) {
                    // thread = output row; 32-column blocks straight to global memory (see epilogue_direct)
                    const int cb = col_half * NCH;
                    auto block = [&](int c, const uint32_t* pre) {
                        uint32_t r[32];
                        const uint32_t taddr = tmem_base + uint32_t(acc * 256 + c * 32) + (uint32_t(quad * 32) << 16);
                        LPI_TMEM_LD_X32(taddr, r);
                        tmem_ld_wait();
                        epilogue_direct<EPI, F16>(p, r, row, n0 + c * 32, valid, pre);
                    };
                    block(cb + 0, d0);
                    if (NCH > 2) prefetch(cb + 2, d0);
                    if (NCH > 1) block(cb + 1, d1);
                    if (NCH > 3) prefetch(cb + 3, d1);
                    if (NCH > 2) block(cb + 2, d0);
                    if (NCH > 3) block(cb + 3, d1);
                } else if (MODE == MODE_GEMM) {
                    auto do_block = [&](int c, const uint2* pre, const float4* pre_rs = nullptr, int dl = 0) {
#if defined(LPI_DEBUG_PROBE) && LPI_DEBUG_PROBE == 2
                        return;                                // timing study: no epilogue at all (results WRONG)
#endif
                        uint32_t r[32];
                        const uint32_t taddr = tmem_base + uint32_t(acc * 256 + c * 32) + (uint32_t(quad * 32) << 16);
                        LPI_TMEM_LD_X32(taddr, r);
                        tmem_ld_wait();
#if defined(LPI_DEBUG_PROBE) && LPI_DEBUG_PROBE == 3
                        if (r[0] == 0x7fc01234u && r[31] == 0x7fc04321u) p.out_f32[0] = 1.f;      // timing study: TMEM drain only (results WRONG)
                        return;
#endif
                        if (kDelta && p.delta != nullptr && valid) {
                            // delta[b, h, l] += sum over these 32 columns of dO (fp32 accumulator) * O: two addends per (row, head) on a
                            // zeroed buffer, so the result does not depend on their order
                            const uint32_t* o = od[kDelta ? dl : 0];
                            float dsum = 0.f;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float2 of = F16 ? __half22float2(*reinterpret_cast<const __half2*>(&o[j]))
                                                      : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o[j]));
                                dsum = fmaf(__uint_as_float(r[2 * j]), of.x, dsum);
                                dsum = fmaf(__uint_as_float(r[2 * j + 1]), of.y, dsum);
                            }
                            const int bb = row / p.delta_L, ll = row - bb * p.delta_L, hh = (n0 + c * 32) >> 6;
                            atomicAdd(p.delta + (size_t(bb) * (p.N >> 6) + hh) * p.delta_L + ll, dsum);
                        }
                        float4* stg = reinterpret_cast<float4*>(smem_gen + C::LIST_OFF) + (warp - 2) * 32 * 8;
                        constexpr bool k16 = (EPI == EPI_BIAS_BF16 || EPI == EPI_BF16 || EPI == EPI_BIAS_GELU_BF16) && LPI_EPI16;
                        if (k16) {
                            epilogue_block16<EPI, F16>(p, r, reinterpret_cast<uint4*>(stg), m0 + quad * 32, n0 + c * 32, lane);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                stg[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                                               __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                            __syncwarp();
                            epilogue_block<EPI, F16>(p, stg, m0 + quad * 32, n0 + c * 32, lane, pre, pre_rs);
                            __syncwarp();
                        }
                    };
                    if (kPrefetchAux) {
                        static_assert(NCH >= 2 && NCH <= 4, "the aux prefetch schedule below is written for 2-4 blocks per warp");
                        const int cb = col_half * NCH;
                        do_block(cb + 0, pre0);
                        if (NCH > 2) prefetch_aux_block(p, m0 + quad * 32, n0 + (cb + 2) * 32, lane, pre0);
                        do_block(cb + 1, pre1);
                        if (NCH > 3) prefetch_aux_block(p, m0 + quad * 32, n0 + (cb + 3) * 32, lane, pre1);
                        if (NCH > 2) do_block(cb + 2, pre0);
                        if (NCH > 3) do_block(cb + 3, pre1);
                    } else if (kPrefetchResid) {
                        const int cb = col_half * NCH;
                        do_block(cb + 0, nullptr, rs0);
                        if (NCH > 2) prefetch_resid_block(p, m0 + quad * 32, n0 + (cb + 2) * 32, lane, rs0);
                        do_block(cb + 1, nullptr, rs1);
                        if (NCH > 3) prefetch_resid_block(p, m0 + quad * 32, n0 + (cb + 3) * 32, lane, rs1);
                        if (NCH > 2) do_block(cb + 2, nullptr, rs0);
                        if (NCH > 3) do_block(cb + 3, nullptr, rs1);
                    } else {
                        if (kDelta) {                  // unrolled: the prefetched O segments are indexed statically
#pragma unroll
                            for (int cc = 0; cc < NCH; ++cc) do_block(col_half * NCH + cc, nullptr, nullptr, cc);
                        } else {
#pragma unroll 1
                            for (int c = col_half * NCH; c < (col_half + 1) * NCH; ++c) do_block(c, nullptr);
                        }
                    }
                } else {
                    float tmax = -INFINITY;
                    if (coop) {
                        coop_floor = fmaxf(coop_floor, float_prev(coop_in));    // "v > float_prev(s)" == "v >= s": ties with the bound stay in
                        tk.thr = fmaxf(tk.thr, coop_floor);
                    }
                    const float thr_before = tk.thr;
#pragma unroll 1
                    for (int c = 0; c < BN / 64; ++c) {
                        uint32_t r[64];
                        const uint32_t taddr = tmem_base + uint32_t(acc * 256 + c * 64) + (uint32_t(quad * 32) << 16);
                        LPI_TMEM_LD_X64(taddr, r);
                        tmem_ld_wait();
                        if (p.seed_mode) tmax = fmaxf(tmax, block_max<64>(r, n0 + c * 64, p.N));
                        else topk_consume<64>(tk, r, n0 + c * 64, p.N, int(p.gallery_offset));
                    }
                    if (p.seed_mode && tmax > tk.thr) {      // one candidate per tile: its maximum
                        const TopkCT ct = topk_insert(tk.sc, tk.id, tk.row, tk.k, tk.cnt, tk.thr, tmax, n0);
                        tk.cnt = ct.cnt;
                        tk.thr = ct.thr;
                    }
                    if (coop && tk.cnt == tk.k && tk.thr != thr_before) {
                        // an insertion into a full list leaves thr = the list's own k-th score: publish it if it improved, and keep
                        // filtering with the better of it and the outside bound
                        if (tk.thr > coop_pub) { coop_pub = tk.thr; coop_publish(p.shared_thr + row, tk.thr); }
                        tk.thr = fmaxf(tk.thr, coop_floor);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc == 0 ? tempty_leader0 : tempty_leader1);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (MODE == MODE_TOPK && row < p.M) {
                float* os = p.topk_scores + (size_t(second) * p.M + row) * p.k;
                int* oi = p.topk_idx + (size_t(second) * p.M + row) * p.k;
                for (int j = 0; j < p.k; ++j) {
                    os[j] = j < tk.cnt ? tk.sc[j * BM + r_local] : -INFINITY;
                    oi[j] = j < tk.cnt ? tk.id[j * BM + r_local] : 0x7fffffff;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                          // nobody frees TMEM / exits while the peer may still touch this CTA
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair<C::TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------ host side
static PFN_encodeTiled g_encode = nullptr;

int ensure_tma_encoder() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
        return set_error(LPI_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    return 0;
}

PFN_encodeTiled tma_encoder() { return g_encode; }      // valid after ensure_tma_encoder()

// 2D row-major [rows, cols] tensor map with a [box_rows, box_cols] box, SWIZZLE_128B.
// A tensor map is a pure function of (base pointer, element type, extents, strides, box): the encoded descriptors are kept in a small
// per-thread direct-mapped cache, so the eager path (one Python thread issuing ~400 launches per training step, weights and the
// caching allocator's activation buffers at the same addresses every step) stops paying two or three cuTensorMapEncodeTiled calls per
// kernel; graph replays never reach this code.  SWIZZLE_128B, no interleave, 256 B L2 promotion, zero fill -- as every kernel here uses.
int make_tmap_cached(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                     const cuuint32_t* box) {
    if (int rc = ensure_tma_encoder()) return rc;
    struct Entry {
        const void* ptr;
        cuuint64_t dims[3], strides[2];
        cuuint32_t box[3];
        int dt, rank;
        bool valid;
        CUtensorMap map;
    };
    constexpr int kSlots = 1024;
    static thread_local Entry* cache = nullptr;
    if (!cache) cache = new Entry[kSlots]();
    Entry key{};
    key.ptr = ptr; key.dt = int(dt); key.rank = rank;
    for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides[i];
    uint64_t h = reinterpret_cast<uintptr_t>(ptr) * 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 3; ++i) h = (h ^ key.dims[i] ^ (uint64_t(key.box[i]) << 40)) * 0xFF51AFD7ED558CCDull;
    h ^= key.strides[0] * 31 + key.strides[1] * 131 + uint64_t(key.dt) * 7 + uint64_t(rank);
    Entry& e = cache[(h >> 17) & (kSlots - 1)];
    if (e.valid && e.ptr == key.ptr && e.dt == key.dt && e.rank == key.rank && !memcmp(e.dims, key.dims, sizeof(key.dims)) &&
        !memcmp(e.strides, key.strides, sizeof(key.strides)) && !memcmp(e.box, key.box, sizeof(key.box))) {
        *m = e.map;
        return 0;
    }
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, dt, cuuint32_t(rank), const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(LPI_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d", int(r));
    key.valid = true;
    key.map = *m;
    e = key;
    return 0;
}

int make_tmap_2d(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int elem_bytes, uint64_t rows, uint64_t cols,
                 uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld_elems * elem_bytes) & 15))
        return set_error(LPI_ERR_ARG, "TMA operand must be 16-byte aligned (ptr=%p ld=%llu)", ptr, (unsigned long long)ld_elems);
    if (box_cols * elem_bytes != 128 || box_rows > 256)
        return set_error(LPI_ERR_ARG, "bad TMA box %ux%u", box_rows, box_cols);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * elem_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    return make_tmap_cached(m, ptr, dt, 2, dims, strides, box);
}

static int g_num_sms = 0;
int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <int MODE, int BN, int EPI, bool TF32 = false>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, int grid, cudaStream_t st) {
    auto kern = gemm_tn_kernel<MODE, BN, EPI, TF32>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES);
        if (e != cudaSuccess) return set_error(LPI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    kern<<<grid, GEMM_THREADS, Cfg<BN>::SMEM_BYTES, st>>>(tmA, tmB, a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(LPI_ERR_CUDA, "gemm launch: %s", cudaGetErrorString(e));
    return 0;
}

template <int MODE, int EPI, int OP, int BN = 256, int CL = 2>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, int n_clusters, cudaStream_t st) {
    auto kern = gemm_pair_kernel<MODE, EPI, OP, BN, CL>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<MODE, BN>::SMEM_BYTES);
        if (e != cudaSuccess) return set_error(LPI_ERR_CUDA, "cudaFuncSetAttribute(pair): %s", cudaGetErrorString(e));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL * n_clusters);
    cfg.blockDim = dim3(PairCfg<MODE, BN>::THREADS);
    cfg.dynamicSmemBytes = PairCfg<MODE, BN>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // see lpi_internal.h (launch_pdl) and the kernel's pdl_wait()
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (MODE == MODE_GEMM && pdl_enabled()) ? 2 : 1;
    if (CL > 2) {
        // the kernel is persistent (one CTA per SM): a grid larger than what can be co-resident would run in two rounds.  4-CTA clusters
        // must sit inside one GPC, so fewer than sms / 4 of them may fit (GPCs with a TPC count that is not a multiple of two)
        static int max_clusters = 0;
        if (!max_clusters) {
            int n = 0;
            cudaLaunchConfig_t probe = cfg;
            probe.gridDim = dim3(CL * (num_sms() / CL));
            if (cudaOccupancyMaxActiveClusters(&n, kern, &probe) != cudaSuccess || n < 1) n = num_sms() / CL;
            max_clusters = n;
            if (getenv("LPI_GEMM_VERBOSE")) fprintf(stderr, "[lpi_b200] %d-CTA clusters co-resident: %d (of %d SMs)\n", CL, n, num_sms());
        }
        if (n_clusters > max_clusters) cfg.gridDim = dim3(CL * max_clusters);
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, a);
    if (e != cudaSuccess) return set_error(LPI_ERR_CUDA, "pair gemm launch: %s", cudaGetErrorString(e));
    return 0;
}

template <int OP, int BN, int CL>
static int launch_pair_epi_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, int n_clusters, cudaStream_t st) {
    constexpr bool TF32 = OP == OP_TF32;
    constexpr int OPH = TF32 ? OP_BF16 : OP;          // the 16-bit-only epilogues are never instantiated for TF32 operands
    switch (a.epi) {
        case EPI_BIAS_BF16: return launch_pair<MODE_GEMM, EPI_BIAS_BF16, OP, BN, CL>(tmA, tmB, a, n_clusters, st);
        case EPI_BIAS_RESID_F32: return launch_pair<MODE_GEMM, EPI_BIAS_RESID_F32, OP, BN, CL>(tmA, tmB, a, n_clusters, st);
        case EPI_F32: return launch_pair<MODE_GEMM, EPI_F32, OP, BN, CL>(tmA, tmB, a, n_clusters, st);
        case EPI_BF16: return launch_pair<MODE_GEMM, EPI_BF16, OP, BN, CL>(tmA, tmB, a, n_clusters, st);
        case EPI_BIAS_GELU_BF16: if (!TF32) return launch_pair<MODE_GEMM, EPI_BIAS_GELU_BF16, OPH, BN, CL>(tmA, tmB, a, n_clusters, st); break;
        case EPI_DGELU_BF16: if (!TF32) return launch_pair<MODE_GEMM, EPI_DGELU_BF16, OPH, BN, CL>(tmA, tmB, a, n_clusters, st); break;
        case EPI_BIAS_F32: if (!TF32) return launch_pair<MODE_GEMM, EPI_BIAS_F32, OPH, BN, CL>(tmA, tmB, a, n_clusters, st); break;
        case EPI_ACC_F32: if (!TF32) return launch_pair<MODE_GEMM, EPI_ACC_F32, OPH, BN, CL>(tmA, tmB, a, n_clusters, st); break;
        case EPI_BIAS_GELU_F32: if (TF32 && BN == 256 && CL == 2) return launch_pair<MODE_GEMM, EPI_BIAS_GELU_F32, OP_TF32, 256, 2>(tmA, tmB, a, n_clusters, st); break;
        case EPI_DGELU_F32: if (TF32 && BN == 256 && CL == 2) return launch_pair<MODE_GEMM, EPI_DGELU_F32, OP_TF32, 256, 2>(tmA, tmB, a, n_clusters, st); break;
    }
    return set_error(LPI_ERR_ARG, "epilogue %d is not available for this operand type / tile width %d / cluster %d", a.epi, BN, CL);
}

template <int OP>
static int launch_pair_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, int n_clusters, int bn, int cl, cudaStream_t st) {
    if (OP != OP_TF32) {                              // narrower cluster tiles and 4-CTA clusters exist for the 16-bit operand types only
        constexpr int O16 = OP == OP_TF32 ? OP_BF16 : OP;
        if (cl == 4) {
            if (bn == 256) return launch_pair_epi_bn<O16, 256, 4>(tmA, tmB, a, n_clusters, st);
            if (bn == 192) return launch_pair_epi_bn<O16, 192, 4>(tmA, tmB, a, n_clusters, st);
            if (bn == 128) return launch_pair_epi_bn<O16, 128, 4>(tmA, tmB, a, n_clusters, st);
        }
        if (bn == 192) return launch_pair_epi_bn<O16, 192, 2>(tmA, tmB, a, n_clusters, st);
        if (bn == 128) return launch_pair_epi_bn<O16, 128, 2>(tmA, tmB, a, n_clusters, st);
    }
    return launch_pair_epi_bn<OP, 256, 2>(tmA, tmB, a, n_clusters, st);
}

template <int BN>
static int launch_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, int grid, cudaStream_t st) {
    switch (a.epi) {
        case EPI_BIAS_BF16: return launch<MODE_GEMM, BN, EPI_BIAS_BF16>(tmA, tmB, a, grid, st);
        case EPI_BIAS_GELU_BF16: return launch<MODE_GEMM, BN, EPI_BIAS_GELU_BF16>(tmA, tmB, a, grid, st);
        case EPI_BIAS_RESID_F32: return launch<MODE_GEMM, BN, EPI_BIAS_RESID_F32>(tmA, tmB, a, grid, st);
        case EPI_F32: return launch<MODE_GEMM, BN, EPI_F32>(tmA, tmB, a, grid, st);
        case EPI_BIAS_F32: return launch<MODE_GEMM, BN, EPI_BIAS_F32>(tmA, tmB, a, grid, st);
        case EPI_ACC_F32: return launch<MODE_GEMM, BN, EPI_ACC_F32>(tmA, tmB, a, grid, st);
        case EPI_DGELU_BF16: return launch<MODE_GEMM, BN, EPI_DGELU_BF16>(tmA, tmB, a, grid, st);
        case EPI_BF16: return launch<MODE_GEMM, BN, EPI_BF16>(tmA, tmB, a, grid, st);
    }
    return set_error(LPI_ERR_ARG, "unknown epilogue %d", a.epi);
}

template <int BN>
static int launch_epi_tf32(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, int grid, cudaStream_t st) {
    switch (a.epi) {
        case EPI_BIAS_BF16: return launch<MODE_GEMM, BN, EPI_BIAS_BF16, true>(tmA, tmB, a, grid, st);
        case EPI_BIAS_RESID_F32: return launch<MODE_GEMM, BN, EPI_BIAS_RESID_F32, true>(tmA, tmB, a, grid, st);
        case EPI_F32: return launch<MODE_GEMM, BN, EPI_F32, true>(tmA, tmB, a, grid, st);
        case EPI_BF16: return launch<MODE_GEMM, BN, EPI_BF16, true>(tmA, tmB, a, grid, st);
        case EPI_BIAS_GELU_F32: return launch<MODE_GEMM, BN, EPI_BIAS_GELU_F32, true>(tmA, tmB, a, grid, st);
        case EPI_DGELU_F32: return launch<MODE_GEMM, BN, EPI_DGELU_F32, true>(tmA, tmB, a, grid, st);
    }
    return set_error(LPI_ERR_ARG, "epilogue %d is not available for TF32 operands", a.epi);
}

}  // namespace lpi

using namespace lpi;

static thread_local float* g_delta_out;
static thread_local int g_delta_L;

static int gemm_entry(int op, const void* A, const void* B, int M, int N, int K, int epi, const void* bias, const void* resid,
                      void* out, void* out2, const void* aux, int ldo, int tile_n, void* stream) {
    const bool tf32 = op == OP_TF32;
    const int bke = tf32 ? 32 : BK;
    if (M <= 0 || N <= 0 || K <= 0) return set_error(LPI_ERR_ARG, "gemm: empty problem %dx%dx%d", M, N, K);
    if (K % bke) return set_error(LPI_ERR_ARG, "gemm: K=%d must be a multiple of %d", K, bke);
    if (N % 128) return set_error(LPI_ERR_ARG, "gemm: N=%d must be a multiple of 128", N);
    if (ldo < N || (ldo % 8)) return set_error(LPI_ERR_ARG, "gemm: bad ldo=%d", ldo);
    const bool need_bias = (epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16 || epi == EPI_BIAS_RESID_F32 || epi == EPI_BIAS_F32 ||
                            epi == EPI_BIAS_GELU_F32);
    if (need_bias && !bias) return set_error(LPI_ERR_ARG, "gemm: epilogue %d needs a bias", epi);
    if (epi == EPI_BIAS_RESID_F32 && !resid) return set_error(LPI_ERR_ARG, "gemm: residual epilogue needs resid");
    if ((epi == EPI_DGELU_BF16 || epi == EPI_DGELU_F32) && !aux)
        return set_error(LPI_ERR_ARG, "gemm: dgelu epilogue needs the saved pre-activation");
    if (!out) return set_error(LPI_ERR_ARG, "gemm: null output");
    int bn = tile_n;
    const int sms = num_sms();
    // tile_n: 0 = automatic; 128 / 256 = 1-CTA 128 x tile_n tiles; 512 (= 1256) / 1192 / 1128 = CTA-pair 256 x {256, 192, 128} cluster tiles
    int pair_bn = 0, cl = 2;
    // 4-CTA clusters (B tile multicast between two CTA pairs) are OPT-IN (LPI_GEMM_CLUSTER4=1, or tile_n 2xxx): correct on every shape
    // and ~10 % faster per SM, but only 33 such clusters (132 of 148 SMs) can be co-resident on a B200 -- 4-CTA clusters must sit
    // inside one GPC -- which cancels the gain: qkv 43.4 vs 42.0 us, K = 3072 dgrad 64.3 vs 56.0 us (profiles/r2_gemm_cluster4.txt)
    static int cluster4 = -1;
    if (cluster4 < 0) {
        const char* e = getenv("LPI_GEMM_CLUSTER4");
        cluster4 = (e && e[0] == '1') ? 1 : 0;
    }
    const long mp_all = ((M + BM - 1) / BM + 1) / 2;
    if (bn == 0) {
        // CTA-pair tiles whenever N allows: they measured at or above the 1-CTA tiles on every encoder shape
        // (profiles/r2_gemm_microbench.txt).  Cluster tile width and cluster size are the combination that wastes least of the last wave
        // (full waves of 74 pair clusters / 37 four-CTA clusters) x a per-tile efficiency: a narrower tile streams more operand bytes per
        // FLOP, a 4-CTA cluster (B multicast to two pairs) 25 % fewer.
        double best = 0.0;
        const int cand[3] = {256, 192, 128};
        const double weight[3] = {1.0, 0.96, 0.88};
        for (int c4 = 0; c4 < 2; ++c4) {
            if (c4 && (tf32 || !cluster4 || mp_all < 4)) continue;
            for (int i = 0; i < 3; ++i) {
                if (N % cand[i] || (tf32 && cand[i] != 256)) continue;
                const long tiles = (c4 ? (mp_all + 1) / 2 : mp_all) * (N / cand[i]), ncl = c4 ? sms / 4 : sms / 2;
                const double eff = double(tiles) / double(((tiles + ncl - 1) / ncl) * ncl) * weight[i] * (c4 ? 1.10 : 1.0);
                if (eff > best + 1e-9) { best = eff; pair_bn = cand[i]; cl = c4 ? 4 : 2; }
            }
        }
        if (!pair_bn) bn = 128;
    } else if (bn == 512 || bn == 1256) pair_bn = 256;
    else if (bn == 1192) pair_bn = 192;
    else if (bn == 1128) pair_bn = 128;
    else if (bn == 2256 || bn == 2192 || bn == 2128) { pair_bn = bn - 2000; cl = 4; }
    else if (bn != 128 && bn != 256) return set_error(LPI_ERR_ARG, "gemm: tile_n must be 0, 128, 256 (1-CTA), 512 / 1192 / 1128 (CTA pair) or 2256 / 2192 / 2128 (two pairs, 4-CTA cluster)");
    if (cl == 4 && tf32) return set_error(LPI_ERR_ARG, "gemm_tf32: 4-CTA clusters are built for the 16-bit operand types only");
    const bool pair = pair_bn != 0;
    if (pair && tf32 && pair_bn != 256) return set_error(LPI_ERR_ARG, "gemm_tf32: only the 256-wide CTA-pair tile is built");
    if (N % (pair ? pair_bn : bn)) return set_error(LPI_ERR_ARG, "gemm: N=%d not a multiple of the tile width (tile_n=%d)", N, tile_n);
    CUtensorMap tmA, tmB;
    const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (op == OP_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
    const int eb = tf32 ? 4 : 2;
    if (int rc = make_tmap_2d(&tmA, A, dt, eb, M, K, K, BM, bke)) return rc;
    if (int rc = make_tmap_2d(&tmB, B, dt, eb, N, K, K, pair ? pair_bn / cl : bn, bke)) return rc;      // box = what ONE CTA fetches of the B tile
    GemmArgs a{};
    a.M = M; a.N = N; a.K = K; a.epi = epi; a.ldo = ldo;
    {
        a.precise_act = LPI_F16_PRECISE_ACT;
#ifdef LPI_DEBUG_PROBE
        // timing-study build only (never the shipped library): a stray environment variable must not be able to corrupt results
        static int probe = -1;
        if (probe < 0) {
            const char* e = getenv("LPI_GEMM_PROBE");
            probe = e ? atoi(e) : 0;
            if (probe) fprintf(stderr, "[lpi_b200] LPI_GEMM_PROBE=%d: GEMM RESULTS ARE WRONG (timing study build)\n", probe);
        }
        a.probe = probe;
#endif
    }
    a.bias = static_cast<const float*>(bias);
    a.resid = static_cast<const float*>(resid);
    a.delta = (epi == EPI_BF16) ? g_delta_out : nullptr;
    a.delta_L = g_delta_L;
    if (epi == EPI_DGELU_F32) a.aux_f32 = static_cast<const float*>(aux);
    else a.aux_bf16 = static_cast<const __nv_bfloat16*>(aux);
    if (epi == EPI_BIAS_GELU_F32) a.out2_f32 = static_cast<float*>(out2);
    else a.out2_bf16 = static_cast<__nv_bfloat16*>(out2);
    if (epi == EPI_BIAS_RESID_F32 || epi == EPI_F32 || epi == EPI_ACC_F32 || epi == EPI_BIAS_F32 || epi == EPI_BIAS_GELU_F32 ||
        epi == EPI_DGELU_F32) {
        a.out_f32 = static_cast<float*>(out);
        a.out_bf16 = (epi == EPI_BIAS_RESID_F32 || epi == EPI_ACC_F32) ? static_cast<__nv_bfloat16*>(out2) : nullptr;
        a.out2_bf16 = nullptr;
    } else {
        a.out_bf16 = static_cast<__nv_bfloat16*>(out);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pair) {
        const long ctiles = (cl == 4 ? (mp_all + 1) / 2 : mp_all) * (N / pair_bn);
        const long max_cl = sms / cl;
        // dynamic tile scheduling (cluster launch control): the grid holds one cluster per tile and the resident clusters take over the
        // tiles of the pending ones.  Static round-robin over min(tiles, 74) persistent clusters with LPI_GEMM_CLC=0 (and for 4-CTA
        // clusters / TF32 operands).
        // OPT-IN (LPI_GEMM_CLC=1): per GEMM it is as fast as the static schedule once a cluster keeps only ONE request in flight, issued
        // late in its tile (three in flight per cluster let clusters hoard tiles: +2-4 us on the 2.9-tiles-per-cluster shapes), but the
        // training step gains nothing from it (8.37 vs 8.35 ms; vision tower alone 7.32 vs 7.38 ms), so the static schedule stays.
        // (Capping the text tower's GEMMs to 8-37 persistent pairs so that they hold fewer SMs beside the vision tower was also measured:
        // 8.50-8.95 ms against 8.48 ms without a cap.)
        static int use_clc = -1;
        if (use_clc < 0) {
            const char* e = getenv("LPI_GEMM_CLC");
            use_clc = (e && e[0] == '1') ? 1 : 0;
        }
        a.clc = (use_clc && cl == 2 && !tf32 && ctiles > max_cl) ? 1 : 0;
        const int n_clusters = int(a.clc ? ctiles : (ctiles < max_cl ? ctiles : max_cl));
        return tf32 ? launch_pair_epi<OP_TF32>(tmA, tmB, a, n_clusters, pair_bn, cl, st)
                    : (op == OP_F16 ? launch_pair_epi<OP_F16>(tmA, tmB, a, n_clusters, pair_bn, cl, st)
                                    : launch_pair_epi<OP_BF16>(tmA, tmB, a, n_clusters, pair_bn, cl, st));
    }
    if (op == OP_F16) return set_error(LPI_ERR_UNSUPPORTED, "gemm_f16: CTA-pair tiles only (tile_n 0 / 512 / 1192 / 1128)");
    const long tiles = long((M + BM - 1) / BM) * (N / bn);
    const int grid = int(tiles < sms ? tiles : sms);
    if (tf32) return bn == 256 ? launch_epi_tf32<256>(tmA, tmB, a, grid, st) : launch_epi_tf32<128>(tmA, tmB, a, grid, st);
    return bn == 256 ? launch_epi<256>(tmA, tmB, a, grid, st) : launch_epi<128>(tmA, tmB, a, grid, st);
}

// out_proj dgrad (EPI_BF16) with the attention backward's delta fused into the epilogue -- see include/lpi_b200.h
extern "C" int lpi_gemm_do_delta(const void* A, const void* Wt, int M, int N, int K, void* out, const void* o_saved, float* delta, int L, int f16,
                                 void* stream) {
    if (!o_saved || !delta || L <= 0 || (M % L) || (N % 64)) return set_error(LPI_ERR_ARG, "gemm_do_delta: need O, delta, M %% L == 0, N %% 64 == 0 (M=%d L=%d N=%d)", M, L, N);
    g_delta_out = delta;
    g_delta_L = L;
    const int rc = gemm_entry(f16 ? OP_F16 : OP_BF16, A, Wt, M, N, K, EPI_BF16, nullptr, nullptr, out, nullptr, o_saved, N, 0, stream);
    g_delta_out = nullptr;
    g_delta_L = 0;
    return rc;
}

extern "C" int lpi_gemm_bf16(const void* A, const void* B, int M, int N, int K, int epi, const void* bias, const void* resid, void* out,
                             void* out2, const void* aux, int ldo, int tile_n, void* stream) {
    if (epi == EPI_BIAS_GELU_F32 || epi == EPI_DGELU_F32)
        return set_error(LPI_ERR_ARG, "gemm_bf16: epilogue %d is only available for TF32 operands", epi);
    return gemm_entry(OP_BF16, A, B, M, N, K, epi, bias, resid, out, out2, aux, ldo, tile_n, stream);
}

extern "C" int lpi_gemm_tf32(const void* A, const void* B, int M, int N, int K, int epi, const void* bias, const void* resid, void* out,
                             void* out2, const void* aux, int ldo, int tile_n, void* stream) {
    return gemm_entry(OP_TF32, A, B, M, N, K, epi, bias, resid, out, out2, aux, ldo, tile_n, stream);
}

extern "C" int lpi_gemm_f16(const void* A, const void* B, int M, int N, int K, int epi, const void* bias, const void* resid, void* out,
                            void* out2, const void* aux, int ldo, int tile_n, void* stream) {
    if (epi == EPI_BIAS_GELU_F32 || epi == EPI_DGELU_F32)
        return set_error(LPI_ERR_ARG, "gemm_f16: epilogue %d is only available for TF32 operands", epi);
    if (tile_n == 128 || tile_n == 256) return set_error(LPI_ERR_ARG, "gemm_f16: only the CTA-pair tiles (tile_n 0 / 512 / 1192 / 1128 / 2256 / 2192 / 2128) are built");
    return gemm_entry(OP_F16, A, B, M, N, K, epi, bias, resid, out, out2, aux, ldo, tile_n, stream);
}

static bool scorer_pair_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LPI_SCORER_PAIR");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

extern "C" int lpi_sim_topk_chunks(int n_queries, int n_gallery, int* n_chunks_out) {
    // Work items = query tiles (tile PAIRS for the CTA-pair kernel) x gallery chunks; pick the chunk count that fills
    // whole waves of SMs (clusters).
    const bool pair = scorer_pair_enabled();
    const int sms = pair ? num_sms() / 2 : num_sms();
    const int qt1 = (n_queries + BM - 1) / BM;
    const int qt = pair ? (qt1 + 1) / 2 : qt1;
    const int nt = (n_gallery + 255) / 256;
    int best = 1;
    double best_eff = 0;
    for (int c = 1; c <= 64 && c <= nt; ++c) {
        const int tpc = (nt + c - 1) / c;
        if ((nt + tpc - 1) / tpc != c) continue;          // would leave an empty chunk
        long items = long(qt) * c;
        long waves = (items + sms - 1) / sms;
        double eff = double(items) / double(waves * sms);
        // each chunk restarts the top-k warm-up, so only accept more chunks for a real gain
        if (eff > best_eff + 0.02) { best_eff = eff; best = c; }
    }
    *n_chunks_out = best;
    return 0;
}

static int sim_topk_impl(const void* Q, const void* G, int n_queries, int n_gallery, int dim, int k, long long gallery_offset, int n_chunks,
                         const float* init_thr, int init_thr_stride, int seed_mode, float* part_scores, int* part_idx, void* stream,
                         float* shared_thr = nullptr) {
    if (n_queries <= 0 || n_gallery <= 0) return set_error(LPI_ERR_ARG, "sim_topk: empty problem");
    if (dim % BK) return set_error(LPI_ERR_ARG, "sim_topk: dim=%d must be a multiple of %d", dim, BK);
    if (k < 1 || k > TOPK_MAX) return set_error(LPI_ERR_ARG, "sim_topk: k=%d out of range [1,%d]", k, TOPK_MAX);
    if (n_chunks < 1) return set_error(LPI_ERR_ARG, "sim_topk: n_chunks=%d", n_chunks);
    if (gallery_offset + n_gallery > 0x7fffffffLL) return set_error(LPI_ERR_ARG, "sim_topk: gallery index exceeds int32");
    const int nt = (n_gallery + 255) / 256;
    if (n_chunks > nt) return set_error(LPI_ERR_ARG, "sim_topk: n_chunks=%d > gallery tiles=%d", n_chunks, nt);
    const bool pair = scorer_pair_enabled() && dim <= PairCfg<MODE_TOPK>::A_RESIDENT_KB * BK;   // query tile must fit resident
    CUtensorMap tmA, tmB;
    if (int rc = make_tmap_2d(&tmA, Q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, n_queries, dim, dim, BM, BK)) return rc;
    if (int rc = make_tmap_2d(&tmB, G, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, n_gallery, dim, dim, pair ? 128 : 256, BK)) return rc;
    GemmArgs a{};
    a.M = n_queries; a.N = n_gallery; a.K = dim;
    a.k = k; a.n_chunks = n_chunks;
    a.tiles_per_chunk = (nt + n_chunks - 1) / n_chunks;
    if (long(a.tiles_per_chunk) * (n_chunks - 1) >= nt)
        return set_error(LPI_ERR_ARG, "sim_topk: n_chunks=%d leaves an empty chunk for %d tiles", n_chunks, nt);
    a.gallery_offset = gallery_offset;
    a.seed_mode = seed_mode;
    a.init_thr = init_thr;
    a.init_thr_stride = init_thr_stride > 0 ? init_thr_stride : 1;
    a.shared_thr = shared_thr;
    a.topk_scores = part_scores;
    a.topk_idx = part_idx;
    const int sms = num_sms();
    if (pair) {
        const long items = long(((n_queries + BM - 1) / BM + 1) / 2) * n_chunks;
        const int n_clusters = int(items < sms / 2 ? items : sms / 2);
        return launch_pair<MODE_TOPK, EPI_F32, OP_BF16>(tmA, tmB, a, n_clusters, static_cast<cudaStream_t>(stream));
    }
    const long items = long((n_queries + BM - 1) / BM) * n_chunks;
    const int grid = int(items < sms ? items : sms);
    return launch<MODE_TOPK, 256, EPI_F32>(tmA, tmB, a, grid, static_cast<cudaStream_t>(stream));
}

extern "C" int lpi_sim_topk_bf16(const void* Q, const void* G, int n_queries, int n_gallery, int dim, int k,
                                 long long gallery_offset, int n_chunks, const float* init_thr, int init_thr_stride, float* part_scores,
                                 int* part_idx, void* stream) {
    return sim_topk_impl(Q, G, n_queries, n_gallery, dim, k, gallery_offset, n_chunks, init_thr, init_thr_stride, 0, part_scores, part_idx,
                         stream);
}

// Cooperative variant: `shared_thr` [n_queries] fp32 is read AND written -- on entry a score that at least k rows of the logical gallery
// are known to reach per query (a seed / the k-th scores of earlier chunks, or -inf), on exit the best k-th score any chunk of this
// launch reached.  The work items of the launch exchange their thresholds through it (see coop_publish): same merged result, fewer
// sorted insertions; the per-chunk lists may hold fewer than k entries.  Needs the CTA-pair kernel (dim <= 512).
extern "C" int lpi_sim_topk_coop_bf16(const void* Q, const void* G, int n_queries, int n_gallery, int dim, int k,
                                      long long gallery_offset, int n_chunks, float* shared_thr, float* part_scores, int* part_idx,
                                      void* stream) {
    if (!shared_thr) return set_error(LPI_ERR_ARG, "sim_topk_coop: shared_thr is null");
    return sim_topk_impl(Q, G, n_queries, n_gallery, dim, k, gallery_offset, n_chunks, shared_thr, 1, 0, part_scores, part_idx, stream,
                         shared_thr);
}

// Threshold pre-pass: seed_scores[q, 0..k) = the k largest per-tile (256 gallery rows) maxima of query q over the first n_rows rows.
// The k-th of them is reached by k distinct gallery rows, so `seed_scores + (k - 1)` with stride k is a valid init_thr for
// lpi_sim_topk_bf16 over any gallery containing these rows; one candidate per tile keeps this pass MMA-bound (no warm-up insertions).
extern "C" int lpi_sim_topk_seed_bf16(const void* Q, const void* G, int n_queries, int n_rows, int dim, int k, float* seed_scores,
                                      int* seed_idx_ws, void* stream) {
    return sim_topk_impl(Q, G, n_queries, n_rows, dim, k, 0, 1, nullptr, 1, 1, seed_scores, seed_idx_ws, stream);
}

// The same pre-pass cut into n_chunks work items per query tile (seed_scores / seed_idx_ws [n_chunks, n_queries, k]): 98 query-tile pairs
// on 74 clusters are 1.3 waves, 98 x 3 shorter items are 3.97.  Merge the chunk lists (lpi_topk_merge) and take the k-th entry.
extern "C" int lpi_sim_topk_seed_chunks_bf16(const void* Q, const void* G, int n_queries, int n_rows, int dim, int k, int n_chunks,
                                             float* seed_scores, int* seed_idx_ws, void* stream) {
    return sim_topk_impl(Q, G, n_queries, n_rows, dim, k, 0, n_chunks, nullptr, 1, 1, seed_scores, seed_idx_ws, stream);
}
