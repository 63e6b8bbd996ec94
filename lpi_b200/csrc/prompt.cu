// DecomposedPrompt: rank-r tri-factor reconstruction of the [layer, prompt, width] prompt tensors, forward and backward.
// Reference: retrieval/models/prompts/prompts.py:38-57 -- Y[l,p,d] = mean_k a[l,k] b[p,k] c[d,k] with the layer factor `a`
// (dim_1_share) shared between the visual and textual tensors; the reference materialises [L,P,D,r] temporaries, here the
// contraction happens in registers.  Backward:
//   da[l,k] = 1/r sum_{p,d} G[l,p,d] b[p,k] c[d,k]  (summed over both modalities),  db[p,k] = 1/r sum_{l,d} G a c,
//   dc[d,k] = 1/r sum_{l,p} G a b.
#include "ptx.cuh"
#include "lpi_internal.h"

namespace lpi {

constexpr int RMAX = 8;

struct PromptArgs {
    const float* a;       // [L, r] shared
    const float* b[2];    // [P, r] visual / textual
    const float* c[2];    // [D_m, r]
    float* y[2];          // [L, P, D_m]
    int L, P, r;
    int D[2];
};

// grid = (L*P, 2): one block per (layer, prompt) row and modality
__global__ void prompt_fwd_kernel(PromptArgs p) {
    const int m = blockIdx.y, row = blockIdx.x;
    const int l = row / p.P, pp = row % p.P;
    float ab[RMAX];
    for (int k = 0; k < p.r; ++k) ab[k] = p.a[l * p.r + k] * p.b[m][pp * p.r + k];
    const float inv_r = 1.f / p.r;
    const int D = p.D[m];
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < p.r; ++k) s = fmaf(ab[k], p.c[m][d * p.r + k], s);   // same association as the reference: (a*b)*c
        p.y[m][long(row) * D + d] = s * inv_r;
    }
}

struct PromptBwdArgs {
    const float* a;
    const float* b[2];
    const float* c[2];
    const float* G[2];    // [L, P, D_m] upstream gradient (dense)
    float* t1;            // workspace [2, L*P, r]
    float* da;            // [L, r]
    float* db[2];         // [P, r]
    float* dc[2];         // [D_m, r]
    int L, P, r;
    int D[2];
};

// K1: t1[m][row][k] = sum_d G[m][row][d] * c[m][d][k]          grid = (L*P, 2), 128 threads
__global__ void prompt_bwd_t1_kernel(PromptBwdArgs p) {
    __shared__ float red[4][RMAX];
    const int m = blockIdx.y, row = blockIdx.x;
    const int D = p.D[m];
    float acc[RMAX];
#pragma unroll
    for (int k = 0; k < RMAX; ++k) acc[k] = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float g = p.G[m][long(row) * D + d];
#pragma unroll
        for (int k = 0; k < RMAX; ++k)
            if (k < p.r) acc[k] = fmaf(g, p.c[m][d * p.r + k], acc[k]);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < RMAX; ++k) {
        const float v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < p.r)
        p.t1[(long(m) * p.L * p.P + row) * p.r + threadIdx.x] =
            red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
}

// K2: dc[m][d][k] = 1/r sum_{l,p} G[m][l,p,d] a[l,k] b[m][p,k]        grid = (ceil(D/128), 2), thread per d
__global__ void prompt_bwd_dc_kernel(PromptBwdArgs p) {
    extern __shared__ float ab[];                    // [L*P][r]
    const int m = blockIdx.y;
    const int rows = p.L * p.P;
    for (int i = threadIdx.x; i < rows * p.r; i += blockDim.x) {
        const int row = i / p.r, k = i % p.r;
        ab[i] = p.a[(row / p.P) * p.r + k] * p.b[m][(row % p.P) * p.r + k];
    }
    __syncthreads();
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    const int D = p.D[m];
    if (d >= D) return;
    float acc[RMAX];
#pragma unroll
    for (int k = 0; k < RMAX; ++k) acc[k] = 0.f;
    for (int row0 = 0; row0 < rows; row0 += 8) {     // 8 independent loads in flight (the serial version paid 144 DRAM latencies)
        float g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = row0 + j < rows ? p.G[m][long(row0 + j) * D + d] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (row0 + j >= rows) break;
#pragma unroll
            for (int k = 0; k < RMAX; ++k)
                if (k < p.r) acc[k] = fmaf(g[j], ab[(row0 + j) * p.r + k], acc[k]);
        }
    }
    const float inv_r = 1.f / p.r;
    for (int k = 0; k < p.r; ++k) p.dc[m][d * p.r + k] = acc[k] * inv_r;
}

// K3: da, db from t1 (single block)
__global__ void prompt_bwd_ab_kernel(PromptBwdArgs p) {
    const float inv_r = 1.f / p.r;
    const int LP = p.L * p.P;
    for (int i = threadIdx.x; i < p.L * p.r; i += blockDim.x) {
        const int l = i / p.r, k = i % p.r;
        float s = 0.f;
        for (int m = 0; m < 2; ++m)
            for (int pp = 0; pp < p.P; ++pp) s = fmaf(p.t1[(long(m) * LP + l * p.P + pp) * p.r + k], p.b[m][pp * p.r + k], s);
        p.da[i] = s * inv_r;
    }
    for (int m = 0; m < 2; ++m)
        for (int i = threadIdx.x; i < p.P * p.r; i += blockDim.x) {
            const int pp = i / p.r, k = i % p.r;
            float s = 0.f;
            for (int l = 0; l < p.L; ++l) s = fmaf(p.t1[(long(m) * LP + l * p.P + pp) * p.r + k], p.a[l * p.r + k], s);
            p.db[m][i] = s * inv_r;
        }
}

}  // namespace lpi

using namespace lpi;

extern "C" int lpi_prompt_fwd(const float* dim1_share, const float* dim2_vis, const float* dim2_txt, const float* dim3_vis, const float* dim3_txt,
                              float* vis_out, float* txt_out, int L, int P, int Dv, int Dt, int r, void* stream) {
    if (r < 1 || r > RMAX) return set_error(LPI_ERR_ARG, "prompt_fwd: rank r=%d out of range [1,%d]", r, RMAX);
    if (L <= 0 || P <= 0) return set_error(LPI_ERR_ARG, "prompt_fwd: empty prompt tensor");
    PromptArgs a{};
    a.a = dim1_share; a.b[0] = dim2_vis; a.b[1] = dim2_txt; a.c[0] = dim3_vis; a.c[1] = dim3_txt;
    a.y[0] = vis_out; a.y[1] = txt_out; a.L = L; a.P = P; a.r = r; a.D[0] = Dv; a.D[1] = Dt;
    prompt_fwd_kernel<<<dim3(L * P, 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("prompt_fwd");
}

extern "C" int lpi_prompt_bwd(const float* dim1_share, const float* dim2_vis, const float* dim2_txt, const float* dim3_vis, const float* dim3_txt,
                              const float* g_vis, const float* g_txt, float* ws /* 2*L*P*r floats */, float* d_dim1, float* d_dim2_vis,
                              float* d_dim2_txt, float* d_dim3_vis, float* d_dim3_txt, int L, int P, int Dv, int Dt, int r, void* stream) {
    if (r < 1 || r > RMAX) return set_error(LPI_ERR_ARG, "prompt_bwd: rank r=%d out of range [1,%d]", r, RMAX);
    PromptBwdArgs a{};
    a.a = dim1_share; a.b[0] = dim2_vis; a.b[1] = dim2_txt; a.c[0] = dim3_vis; a.c[1] = dim3_txt;
    a.G[0] = g_vis; a.G[1] = g_txt; a.t1 = ws; a.da = d_dim1; a.db[0] = d_dim2_vis; a.db[1] = d_dim2_txt;
    a.dc[0] = d_dim3_vis; a.dc[1] = d_dim3_txt; a.L = L; a.P = P; a.r = r; a.D[0] = Dv; a.D[1] = Dt;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    prompt_bwd_t1_kernel<<<dim3(L * P, 2), 128, 0, st>>>(a);
    const int dmax = Dv > Dt ? Dv : Dt;
    prompt_bwd_dc_kernel<<<dim3((dmax + 127) / 128, 2), 128, L * P * r * sizeof(float), st>>>(a);
    prompt_bwd_ab_kernel<<<1, 256, 0, st>>>(a);
    return check_launch("prompt_bwd");
}
