// Internal declarations shared by the .cu translation units of liblpi_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include "../../include/lpi_b200.h"

namespace lpi {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int set_error(int code, const char* fmt, ...);   // records lpi_last_error(), returns `code`
int check_launch(const char* what);              // cudaGetLastError -> error code
int num_sms();
int ensure_tma_encoder();
PFN_encodeTiled tma_encoder();                    // cuTensorMapEncodeTiled entry point (after ensure_tma_encoder() returned 0)
int make_tmap_cached(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                     const cuuint32_t* box);       // SWIZZLE_128B tiled map of rank 2 or 3, memoised per thread
int make_tmap_2d(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int elem_bytes, uint64_t rows, uint64_t cols,
                 uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols);

// Programmatic dependent launch: the tower kernels (GEMM, attention, LayerNorm) are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization so that a kernel's launch latency and prologue overlap the tail of its predecessor
// in the stream (ptx.cuh: pdl_wait / pdl_launch_dependents).  LPI_PDL=0 launches them plainly (the device-side instructions are then
// no-ops).  Works eagerly and under stream capture (the edge becomes a programmatic dependency of the CUDA graph).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// tcgen05 attention (attention_tc.cu); L <= 256
bool attn_tc_enabled(int L);
int attn_fwd_tc(const void* qkv, void* out, float* out_f32, float* lse2, int B, int L, int H, int causal, bool f16, cudaStream_t st);
int attn_bwd_tc(const void* qkv, const void* d_out, const float* lse2, const float* delta, void* dqkv, float* dqkv_f32, int B, int L, int H,
                int causal, bool f16, cudaStream_t st);

}  // namespace lpi
