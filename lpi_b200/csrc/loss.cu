// Loss-side kernels of the LPI learner: the symmetric contrastive (InfoNCE) loss with its backward, the prompt alignment
// and task losses, the fused SGD(momentum, weight decay) step and the L1 nearest-centre task selector.
// Reference: retrieval/loss/loss.py:6-33 (nt_bxent_loss), :75-87 (ClipLoss.forward); retrieval/models/slinet.py:137-183
// (cal_loss / cal_task_loss); retrieval/methods/sprompt.py:253-254 (SGD + cosine), :336-368 (task-id selection).
// Everything here is small (B x B logits with B <= a few thousand, 9 x 9 alignment logits, <= 12 x 12 task matrix), so the
// matrix products are plain fp32 SIMT tiles: exact fp32 products, no tensor-core rounding in the loss path.
#include "ptx.cuh"
#include "lpi_internal.h"
#include <math_constants.h>
#include <cooperative_groups.h>

namespace lpi {

// ------------------------------------------------------------------------------------------------ generic small fp32 GEMM
// C[m,n] = alpha * sum_k A[m*a_m + k*a_k] * B[k*b_k + n*b_n] (+ beta * C[m,n]);  64x64 tile, 16-deep, 256 threads x (4x4)
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C, int M, int N, int K, long a_m, long a_k,
             long b_k, long b_n, long ldc, float alpha, float beta, const float* __restrict__ bias = nullptr) {
    __shared__ float sA[16][64 + 4];
    __shared__ float sB[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const bool a_k_fast = (a_k == 1);
            const int mm = a_k_fast ? i / 16 : i % 64, kk = a_k_fast ? i % 16 : i / 64;
            const int gm = m0 + mm, gk = k0 + kk;
            sA[kk][mm] = (gm < M && gk < K) ? A[gm * a_m + gk * a_k] : 0.f;
            const bool b_k_fast = (b_k == 1);
            const int nn = b_k_fast ? i / 16 : i % 64, kb = b_k_fast ? i % 16 : i / 64;
            const int gn = n0 + nn, gkb = k0 + kb;
            sB[kb][nn] = (gn < N && gkb < K) ? Bm[gkb * b_k + gn * b_n] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
            if (gm < M && gn < N) {
                float v = alpha * acc[i][j];
                if (bias) v += bias[gn];
                if (beta != 0.f) v += beta * C[gm * ldc + gn];
                C[gm * ldc + gn] = v;
            }
        }
}

// ------------------------------------------------------------------------------------------------ ClipLoss on logits
// lse[0..n) = row LSE, lse[n..2n) = column LSE.  Rows: one warp per row.  Columns: one thread per column (coalesced).
__global__ void clip_lse_kernel(const float* __restrict__ S, int n, float* __restrict__ lse) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw < n) {
        const float* r = S + long(gw) * n;
        float m = -CUDART_INF_F;
        for (int j = lane; j < n; j += 32) m = fmaxf(m, r[j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < n; j += 32) s += expf(r[j] - m);
        s = warp_sum(s);
        if (lane == 0) lse[gw] = m + logf(s);
    }
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) {
        float m = -CUDART_INF_F;
        for (int i = 0; i < n; ++i) m = fmaxf(m, S[long(i) * n + c]);
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += expf(S[long(i) * n + c] - m);
        lse[n + c] = m + logf(s);
    }
}

// loss = 1/(2n) sum_i (lse_row[i] + lse_col[i] - 2 S[i,i])      (single block, deterministic tree)
__global__ void clip_loss_reduce_kernel(const float* __restrict__ S, const float* __restrict__ lse, int n, float* __restrict__ loss, float weight) {
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += lse[i] + lse[n + i] - 2.f * S[long(i) * n + i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) *loss = weight * v / (2.f * n);
    }
}

// dS[i,j] = weight * (softmax_row + softmax_col - 2 delta_ij) / (2n)
__global__ void clip_dlogits_kernel(const float* __restrict__ S, const float* __restrict__ lse, int n, float* __restrict__ dS, float weight) {
    const long e = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= long(n) * n) return;
    const int i = int(e / n), j = int(e % n);
    const float s = S[e];
    float g = expf(s - lse[i]) + expf(s - lse[n + j]);
    if (i == j) g -= 2.f;
    dS[e] = weight * g / (2.f * n);
}

// ------------------------------------------------------------------------------------------------ fused similarity + InfoNCE, forward and backward
// north_star subsystem 3 (reference: logits = exp(logit_scale) * I @ T^T, slinet.py:138-141; ClipLoss.forward, loss.py:75-87; autograd):
// ONE cooperative launch computes the n x n scaled similarities, both log-sum-exps, the loss and the gradients of the local feature rows
//     dI[i] = scale * sum_j G[i,j] T[j],  dT[j] = scale * sum_i G[i,j] I[i],   G = weight * (softmax_row + softmax_col - 2 Id) / (2n)
// without the logits ever having to exist in memory (they are written only when the caller asks for them).  One warp owns one row of I
// (or of T): its feature row sits in registers, the lanes split the width, every score is one warp-reduced fp32 dot product (exact fp32
// products: the loss path keeps no tensor-core rounding), recomputed in the backward phase instead of stored.  Phases are separated by
// grid-wide barriers; all reductions run in a fixed order (deterministic).  Work: ~5 n^2 E MACs -- microseconds at the global batches of
// BASELINE configs[2] (64 ... 512), a few ms at n = 4096.
namespace cg = cooperative_groups;

// The opposite feature matrix is walked in tiles of TJ rows staged in shared memory by the whole block (coalesced float4 loads, all in
// flight at once): a warp that fetched each row itself paid one L2 round trip per score (130 us at n = 64 for microseconds of math).
template <int NV>          // feature width E = 32 * NV' with NV' <= NV
__global__ void __launch_bounds__(256)
sim_infonce_kernel(const float* __restrict__ img, const float* __restrict__ txt, int n, int E, float scale, float weight, int row0, int n_local,
                   float* __restrict__ lse, float* __restrict__ terms, float* __restrict__ loss, float* __restrict__ logits,
                   float* __restrict__ d_img, float* __restrict__ d_txt, int TJ) {
    extern __shared__ float tile[];            // [TJ][E]
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nv = E >> 5;
    auto stage = [&](const float* Bm, int j0) {        // rows [j0, j0 + TJ) of Bm -> tile (rows past n: zeros)
        __syncthreads();
        const int n4 = TJ * E / 4;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            const int r = (i * 4) / E;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j0 + r < n) v = *reinterpret_cast<const float4*>(Bm + size_t(j0) * E + size_t(i) * 4);
            reinterpret_cast<float4*>(tile)[i] = v;
        }
        __syncthreads();
    };
    // ---- phase 1: row log-sum-exps (side 0: rows of I against all of T) and column ones (side 1: rows of T against all of I);
    //      a block works on 8 consecutive rows of ONE side, so all its warps walk the same opposite matrix
    const int bps = (n + 7) >> 3;
    for (int bt = blockIdx.x; bt < 2 * bps; bt += gridDim.x) {
        const bool col = bt >= bps;
        const int idx = (bt - (col ? bps : 0)) * 8 + warp;
        const bool valid = idx < n;
        const float* A = (col ? txt : img) + size_t(valid ? idx : 0) * E;
        const float* Bm = col ? img : txt;
        float a[NV];
#pragma unroll
        for (int t = 0; t < NV; ++t) a[t] = (t < nv && valid) ? A[lane + 32 * t] : 0.f;
        float m = -CUDART_INF_F, l = 0.f, diag = 0.f;
        for (int j0 = 0; j0 < n; j0 += TJ) {
            stage(Bm, j0);
            const int jn = min(TJ, n - j0);
            if (valid) {
                for (int jj = 0; jj < jn; ++jj) {
                    const float* b = tile + jj * E;
                    float p = 0.f;
#pragma unroll
                    for (int t = 0; t < NV; ++t) if (t < nv) p = fmaf(a[t], b[lane + 32 * t], p);
                    const float sc = scale * warp_sum(p);
                    const float mn = fmaxf(m, sc);
                    l = l * expf(m - mn) + expf(sc - mn);
                    m = mn;
                    if (j0 + jj == idx) diag = sc;
                    if (!col && logits && lane == 0) logits[size_t(idx) * n + j0 + jj] = sc;
                }
            }
        }
        if (valid && lane == 0) {
            const float v = m + logf(l);
            lse[(col ? n : 0) + idx] = v;
            terms[(col ? n : 0) + idx] = v - diag;
        }
    }
    grid.sync();
    // ---- phase 2: loss = weight / (2n) * sum_i [(row_lse_i - S_ii) + (col_lse_i - S_ii)], one block, fixed order
    if (blockIdx.x == 0) {
        __shared__ float red[8];
        float s = 0.f;
        for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) s += terms[i];
        s = warp_sum(s);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float v = 0.f;
            for (int w = 0; w < (blockDim.x >> 5); ++w) v += red[w];
            *loss = weight * v / (2.f * n);
        }
    }
    // ---- phase 3: gradients of the local rows (side 0: dI rows, side 1: dT rows); scores are recomputed from the staged tiles
    if (d_img == nullptr && d_txt == nullptr) return;
    const float w2 = weight / (2.f * n);
    const int bpl = (n_local + 7) >> 3;
    for (int bt = blockIdx.x; bt < 2 * bpl; bt += gridDim.x) {
        const bool col = bt >= bpl;
        const int loc = (bt - (col ? bpl : 0)) * 8 + warp;
        float* out = col ? d_txt : d_img;
        const bool valid = loc < n_local && out != nullptr;
        const int idx = row0 + (loc < n_local ? loc : 0);
        const float* A = (col ? txt : img) + size_t(idx) * E;
        const float* Bm = col ? img : txt;
        const float own = lse[(col ? n : 0) + idx];              // LSE of this row (row side) / column (column side)
        const float* other = col ? lse : lse + n;                // LSEs of the opposite side, indexed by j
        float a[NV], acc[NV];
#pragma unroll
        for (int t = 0; t < NV; ++t) { a[t] = (t < nv && valid) ? A[lane + 32 * t] : 0.f; acc[t] = 0.f; }
        for (int j0 = 0; j0 < n; j0 += TJ) {
            stage(Bm, j0);
            const int jn = min(TJ, n - j0);
            if (valid) {
                for (int jj = 0; jj < jn; ++jj) {
                    const float* b = tile + jj * E;
                    float bv[NV], p = 0.f;
#pragma unroll
                    for (int t = 0; t < NV; ++t) { bv[t] = t < nv ? b[lane + 32 * t] : 0.f; p = fmaf(a[t], bv[t], p); }
                    const float sc = scale * warp_sum(p);
                    float g = expf(sc - own) + expf(sc - other[j0 + jj]);
                    if (j0 + jj == idx) g -= 2.f;
                    g *= w2;
#pragma unroll
                    for (int t = 0; t < NV; ++t) acc[t] = fmaf(g, bv[t], acc[t]);
                }
            }
        }
        if (valid) {
            float* o = out + size_t(loc) * E;
#pragma unroll
            for (int t = 0; t < NV; ++t) if (t < nv) o[lane + 32 * t] = scale * acc[t];
        }
    }
}

// ------------------------------------------------------------------------------------------------ prompt-side helpers
// out[row] = scale * mean_d x[row, d]                     (alignment loss: mean over the width, / temperature)
__global__ void row_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int D, float scale) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s += x[long(row) * D + d];
    s = warp_sum(s);
    if (lane == 0) out[row] = scale * s / D;
}

// G[row, d] (+)= alpha * v[row]                            (backward of row_mean)
__global__ void add_rowconst_kernel(float* __restrict__ G, const float* __restrict__ v, long rows, int D, float alpha, int accumulate) {
    const long e = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= rows * D) return;
    const float a = alpha * v[e / D];
    G[e] = accumulate ? G[e] + a : a;
}

// Gram partials of X [R, n] (R <= 16): part[blk][i*R + j] = sum over this block's slice of X[i,:] . X[j,:]   (j <= i)
__global__ void __launch_bounds__(256) gram_partial_kernel(const float* __restrict__ X, int R, long n, float* __restrict__ part) {
    __shared__ float red[8];
    const long per = (n + gridDim.x - 1) / gridDim.x;
    const long lo = blockIdx.x * per, hi = min(n, lo + per);
    for (int i = 0; i < R; ++i)
        for (int j = 0; j <= i; ++j) {
            float s = 0.f;
            for (long e = lo + threadIdx.x; e < hi; e += 256) s = fmaf(X[i * n + e], X[j * n + e], s);
            s = warp_sum(s);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
            __syncthreads();
            if (threadIdx.x == 0) {
                float t = 0.f;
                for (int w = 0; w < 8; ++w) t += red[w];
                part[long(blockIdx.x) * R * R + i * R + j] = t;
            }
        }
}

// nt_bxent_loss (loss.py:6-33) from the Gram matrix, plus d loss / d cos[t, j] folded into per-row coefficients for the
// LAST row t = R-1 (the only trainable one):  dX_t = sum_j coef[j] * X_j.     Single block.
__global__ void task_loss_kernel(const float* __restrict__ part, int n_part, int R, const int* __restrict__ target /* [R,R] */, float temperature,
                                 float weight, float* __restrict__ loss_out, int loss_accumulate, float* __restrict__ coef /* [R] */) {
    __shared__ float G[16 * 16];
    __shared__ float ell_sum_pos[16], ell_sum_neg[16], dl[16 * 16];
    for (int e = threadIdx.x; e < R * R; e += blockDim.x) {
        const int i = e / R, j = e % R;
        const int a = i >= j ? i : j, b = i >= j ? j : i;
        float s = 0.f;
        for (int p = 0; p < n_part; ++p) s += part[long(p) * R * R + a * R + b];
        G[e] = s;
    }
    __syncthreads();
    if (threadIdx.x < R) {
        const int i = threadIdx.x;
        float sp = 0.f, sn = 0.f;
        int np = 0;
        for (int j = 0; j < R; ++j) np += target[i * R + j] != 0;
        for (int j = 0; j < R; ++j) {
            const float ni = sqrtf(G[i * R + i]), nj = sqrtf(G[j * R + j]);
            const float c = (i == j) ? CUDART_INF_F : G[i * R + j] / fmaxf(ni * nj, 1e-8f);
            const float z = (i == j) ? 1.f : 1.f / (1.f + expf(-c / temperature));
            const float y = target[i * R + j] != 0 ? 1.f : 0.f;
            // binary_cross_entropy_with_logits applied to the already-sigmoided z (reference quirk, loss.py:21)
            const float ell = fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z)));
            if (y != 0.f) sp += ell; else sn += ell;
            // d ell / d c = (1 - y - sigmoid(-z)) * z (1 - z) / temperature ; row weight 1/npos or 1/nneg, and 1/R for the mean
            const float dz = 1.f - y - 1.f / (1.f + expf(z));
            const float w = (y != 0.f) ? 1.f / np : 1.f / (R - np);
            dl[i * R + j] = (i == j) ? 0.f : dz * z * (1.f - z) / temperature * w / R;
        }
        ell_sum_pos[i] = sp / np;
        ell_sum_neg[i] = sn / (R - np);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int i = 0; i < R; ++i) tot += ell_sum_pos[i] + ell_sum_neg[i];
        const float v = weight * tot / R;
        *loss_out = loss_accumulate ? *loss_out + v : v;
        // gradient wrt row t: cos[t,j] appears at (t,j) and (j,t)
        const int t = R - 1;
        const float nt = sqrtf(G[t * R + t]);
        float self = 0.f;
        for (int j = 0; j < R; ++j) {
            if (j == t) continue;
            const float nj = sqrtf(G[j * R + j]);
            const float gc = weight * (dl[t * R + j] + dl[j * R + t]);
            const float c = G[t * R + j] / fmaxf(nt * nj, 1e-8f);
            coef[j] = gc / fmaxf(nt * nj, 1e-8f);           // d cos / dX_t = X_j / (|Xt||Xj|) - cos * X_t / |Xt|^2
            self -= gc * c / (nt * nt);
        }
        coef[t] = self;
    }
}

// out[e] (+)= sum_r coef[r] * X[r, e]
__global__ void axpy_rows_kernel(const float* __restrict__ X, const float* __restrict__ coef, int R, long n, float* __restrict__ out, int accumulate) {
    const long e = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float s = accumulate ? out[e] : 0.f;
    for (int r = 0; r < R; ++r) s = fmaf(coef[r], X[r * n + e], s);
    out[e] = s;
}

// ------------------------------------------------------------------------------------------------ optimiser / task-id
// torch.optim.SGD(momentum, weight_decay), no dampening / nesterov (sprompt.py:253): g += wd*w; v = first ? g : m*v + g; w -= lr*v
__global__ void sgd_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ v, long n, float lr, float momentum, float wd,
                           int first_step) {
    const long e = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const float gg = g[e] + wd * w[e];
    const float vv = first_step ? gg : momentum * v[e] + gg;
    v[e] = vv;
    w[e] -= lr * vv;
}

// sel[b] = argmin_t min_c sum_d |f[b,d] - centers[t,c,d]|  (first occurrence on ties, torch.min semantics); warp per sample
__global__ void nearest_center_kernel(const float* __restrict__ f, const float* __restrict__ centers, int B, int T, int C, int E, long long* __restrict__ sel) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    float best = CUDART_INF_F;
    int best_t = 0;
    for (int t = 0; t < T; ++t) {
        float tmin = CUDART_INF_F;
        for (int c = 0; c < C; ++c) {
            const float* k = centers + (long(t) * C + c) * E;
            float s = 0.f;
            for (int d = lane; d < E; d += 32) s += fabsf(f[long(b) * E + d] - k[d]);
            s = warp_sum(s);
            tmin = fminf(tmin, s);
        }
        if (tmin < best) { best = tmin; best_t = t; }
    }
    if (lane == 0) sel[b] = best_t;
}

}  // namespace lpi

using namespace lpi;

extern "C" int lpi_sgemm_f32(const float* A, const float* B, float* C, int M, int N, int K, long long a_m, long long a_k, long long b_k,
                             long long b_n, long long ldc, float alpha, float beta, void* stream) {
    if (M <= 0 || N <= 0) return LPI_OK;
    sgemm_kernel<<<dim3((N + 63) / 64, (M + 63) / 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(A, B, C, M, N, K, a_m, a_k, b_k, b_n, ldc,
                                                                                                     alpha, beta);
    return check_launch("sgemm_f32");
}

// C = alpha * A @ B + bias[n] + beta * C: the linear layers of the fp32 parity mode of the towers (exact fp32 products, FFMA)
extern "C" int lpi_sgemm_bias_f32(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, long long a_m,
                                  long long a_k, long long b_k, long long b_n, long long ldc, float alpha, float beta, void* stream) {
    if (M <= 0 || N <= 0) return LPI_OK;
    sgemm_kernel<<<dim3((N + 63) / 64, (M + 63) / 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(A, B, C, M, N, K, a_m, a_k, b_k, b_n, ldc,
                                                                                                     alpha, beta, bias);
    return check_launch("sgemm_bias_f32");
}

extern "C" int lpi_clip_loss_logits(const float* logits, int n, float weight, float* lse_ws /* 2n */, float* loss_out, float* dlogits,
                                    void* stream) {
    if (n <= 0) return set_error(LPI_ERR_ARG, "clip_loss: empty logits");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    clip_lse_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(logits, n, lse_ws);
    clip_loss_reduce_kernel<<<1, 256, 0, st>>>(logits, lse_ws, n, loss_out, weight);
    if (dlogits) {
        const long ne = long(n) * n;
        clip_dlogits_kernel<<<unsigned((ne + 255) / 256), 256, 0, st>>>(logits, lse_ws, n, dlogits, weight);
    }
    return check_launch("clip_loss_logits");
}

extern "C" int lpi_sim_infonce_fwd_bwd(const float* img_f, const float* txt_f, int n, int E, float scale, float weight, int row0, int n_local,
                                       float* lse_ws /* 2n */, float* terms_ws /* 2n */, float* loss_out, float* logits_out /* n*n or NULL */,
                                       float* d_img /* [n_local, E] or NULL */, float* d_txt /* [n_local, E] or NULL */, void* stream) {
    if (n <= 0) return set_error(LPI_ERR_ARG, "sim_infonce: empty batch");
    if (E % 32 || E < 32 || E > 1024) return set_error(LPI_ERR_ARG, "sim_infonce: E=%d must be a multiple of 32 in [32, 1024]", E);
    if (row0 < 0 || n_local < 0 || row0 + n_local > n) return set_error(LPI_ERR_ARG, "sim_infonce: local rows [%d, %d) outside [0, %d)", row0, row0 + n_local, n);
    if (!lse_ws || !terms_ws || !loss_out) return set_error(LPI_ERR_ARG, "sim_infonce: workspace / loss pointer missing");
    auto kern = (E <= 512) ? sim_infonce_kernel<16> : sim_infonce_kernel<32>;
    int TJ = 16384 / E;                                          // 64 KB of staged rows per block
    if (TJ > 32) TJ = 32;
    const int smem = TJ * E * int(sizeof(float));
    static int max_blocks[2] = {0, 0};
    const int slot = E <= 512 ? 0 : 1;
    if (!max_blocks[slot]) {
        int per_sm = 0, dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) != cudaSuccess)
            return set_error(LPI_ERR_CUDA, "sim_infonce: cudaFuncSetAttribute failed");
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 64 * 1024) != cudaSuccess || per_sm < 1)
            return set_error(LPI_ERR_CUDA, "sim_infonce: occupancy query failed");
        max_blocks[slot] = per_sm * sms;                         // a cooperative grid must be co-resident
    }
    int grid = 2 * ((n + 7) / 8);                                // one block per 8 rows of one side
    if (grid > max_blocks[slot]) grid = max_blocks[slot];
    if (grid < 1) grid = 1;
    void* args[] = {&img_f, &txt_f, &n, &E, &scale, &weight, &row0, &n_local, &lse_ws, &terms_ws, &loss_out, &logits_out, &d_img, &d_txt, &TJ};
    cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(256), args, size_t(smem), static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return set_error(LPI_ERR_CUDA, "sim_infonce: cooperative launch: %s", cudaGetErrorString(e));
    return LPI_OK;
}

extern "C" int lpi_row_mean(const float* x, float* out, int rows, int D, float scale, void* stream) {
    if (rows <= 0) return LPI_OK;
    row_mean_kernel<<<(rows * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, rows, D, scale);
    return check_launch("row_mean");
}

extern "C" int lpi_add_rowconst(float* G, const float* v, long long rows, int D, float alpha, int accumulate, void* stream) {
    if (rows <= 0) return LPI_OK;
    add_rowconst_kernel<<<unsigned((rows * D + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(G, v, rows, D, alpha, accumulate);
    return check_launch("add_rowconst");
}

extern "C" int lpi_task_loss(const float* X, int R, long long n, const int* target, float temperature, float weight, float* part_ws, int n_part,
                             float* loss_out, int loss_accumulate, float* coef_out, float* grad_last_row, int grad_accumulate, void* stream) {
    if (R < 2 || R > 16) return set_error(LPI_ERR_ARG, "task_loss: R=%d rows must be in [2,16]", R);
    if (n_part < 1) return set_error(LPI_ERR_ARG, "task_loss: n_part=%d", n_part);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    gram_partial_kernel<<<n_part, 256, 0, st>>>(X, R, n, part_ws);
    task_loss_kernel<<<1, 32, 0, st>>>(part_ws, n_part, R, target, temperature, weight, loss_out, loss_accumulate, coef_out);
    if (grad_last_row) axpy_rows_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(X, coef_out, R, n, grad_last_row, grad_accumulate);
    return check_launch("task_loss");
}

extern "C" int lpi_sgd_momentum_step(float* w, const float* g, float* v, long long n, float lr, float momentum, float weight_decay,
                                     int first_step, void* stream) {
    if (n <= 0) return LPI_OK;
    sgd_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(w, g, v, n, lr, momentum, weight_decay, first_step);
    return check_launch("sgd_momentum_step");
}

extern "C" int lpi_nearest_center_l1(const float* feats, const float* centers, int B, int n_tasks, int n_centers, int E, long long* sel_out,
                                     void* stream) {
    if (B <= 0) return LPI_OK;
    if (n_tasks < 1 || n_centers < 1) return set_error(LPI_ERR_ARG, "nearest_center: no centres");
    nearest_center_kernel<<<(B * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(feats, centers, B, n_tasks, n_centers, E, sel_out);
    return check_launch("nearest_center_l1");
}
