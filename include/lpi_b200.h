/* lpi_b200 -- C ABI of the B200-native (sm_100a) LPI continual image-text retrieval hot path.
 *
 * The reference (Kelvin-ywc/LPI, retrieval/) has NO plugin / operator / FFI layer on this path: every op is a
 * PyTorch/ATen call made from Python nn.Modules.  This header is therefore the
 * boundary a maintainer would bind from those modules (ctypes stub in INTEGRATION.md); each entry point cites
 * the reference lines whose arithmetic it replaces (paths relative to /root/reference/retrieval).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library keeps no tensor state;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return 0 on success, a negative LPI_ERR_* otherwise; lpi_last_error() gives the message
 *     (thread-local); no C++ exception crosses the ABI;
 *   - row-major, batch-major token layout: activations are [B*L, D] (the reference uses [L,B,D]).
 */
#ifndef LPI_B200_H
#define LPI_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define LPI_OK 0
#define LPI_ERR_ARG (-1)
#define LPI_ERR_CUDA (-2)
#define LPI_ERR_UNSUPPORTED (-3)

const char* lpi_last_error(void);
int lpi_version(void);
/* Fails (LPI_ERR_UNSUPPORTED) unless the current device is compute capability 10.x. */
int lpi_device_check(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core GEMM  D[M,N] = A[M,K] . B[N,K]^T   (bf16 operands, K-major, fp32 accumulate in TMEM)
 * replaces: F.linear inside nn.MultiheadAttention (in_proj/out_proj) and mlp.c_fc/c_proj,
 *           models/clip/model.py:172-181,183-196, and their autograd dgrads (weights are frozen, no wgrad).
 * K % 64 == 0, N % 128 == 0, ldo % 8 == 0; M arbitrary.  tile_n: 0 = auto, 128 or 256.
 * ------------------------------------------------------------------------------------------------ */
enum lpi_epilogue {
    EPI_BIAS_BF16 = 0,      /* out(bf16)  = acc + bias                                   (QKV projection)      */
    EPI_BIAS_GELU_BF16 = 1, /* out(bf16)  = QuickGELU(acc + bias); out2(bf16) = acc+bias (c_fc, model.py:163)  */
    EPI_BIAS_RESID_F32 = 2, /* out(fp32)  = resid + acc + bias; out2(bf16, optional) = same  (out_proj, c_proj)*/
    EPI_F32 = 3,            /* out(fp32)  = acc                                                                 */
    EPI_ACC_F32 = 4,        /* out(fp32) += acc; out2(bf16, optional) = same             (dgrad into a stream)  */
    EPI_DGELU_BF16 = 5,     /* out(bf16)  = acc * QuickGELU'(aux)                        (c_proj dgrad)         */
    EPI_BF16 = 6,           /* out(bf16)  = acc                                          (out_proj dgrad)       */
    EPI_BIAS_F32 = 7        /* out(fp32)  = acc + bias                                   (patch embedding)      */
};
int lpi_gemm_bf16(const void* A, const void* B, int M, int N, int K, int epi, const void* bias_f32,
                  const void* resid_f32, void* out, void* out2, const void* aux_bf16, int ldo, int tile_n,
                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * Retrieval scorer: similarity GEMM with the top-k kept in the epilogue (score matrix never written).
 * replaces: `score_matrix_t2i = (image_feats @ text_feats.t()).t()` + D2H + per-row np.argsort,
 *           methods/sprompt.py:509,544,559-567,597-599.
 * Q [n_queries, dim] bf16, G [n_gallery, dim] bf16 (this rank's gallery shard), dim % 64 == 0, k <= 16.
 * Writes n_chunks partial lists: part_scores/part_idx [n_chunks, n_queries, k], each sorted by
 * (score desc, global gallery index asc); missing entries are (-inf, INT_MAX).
 * The fp32 accumulation order of a (query, item) score is fixed (K ascending in 16-wide steps), so a score
 * does not depend on tile position, chunking or sharding.
 * ------------------------------------------------------------------------------------------------ */
int lpi_sim_topk_chunks(int n_queries, int n_gallery, int* n_chunks_out);
int lpi_sim_topk_bf16(const void* Q, const void* G, int n_queries, int n_gallery, int dim, int k,
                      long long gallery_offset, int n_chunks, float* part_scores, int* part_idx, void* stream);
/* k-way merge of n_parts partial lists (chunks and/or all-gathered shards) -> [n_queries, k]. */
int lpi_topk_merge(const float* part_scores, const int* part_idx, int n_parts, int n_queries, int k,
                   float* out_scores, int* out_idx, void* stream);
/* Top-k of each row of a dense fp32 score matrix (drop-in `itm_eval(scores_i2t, scores_t2i, ...)`,
 * methods/sprompt.py:550-567,594-599); ties -> lowest index. */
int lpi_topk_rows_f32(const float* scores, int n_rows, int n_cols, long long ld, int k, float* out_scores,
                      int* out_idx, void* stream);
/* Recall@1/5/10 bookkeeping (methods/sprompt.py:559-623): rank[q] = first position p < k with
 * topk_idx[q,p] in GT(q) (GT as CSR gt_ptr/gt_idx), else k.  counts[task] += {rank<1, rank<5, rank<10, 1}. */
int lpi_recall_counts(const int* topk_idx, int n_queries, int k, const int* gt_ptr, const int* gt_idx,
                      const int* task_of_query, int n_tasks, int* counts /* [n_tasks,4] zeroed by callee */,
                      int* rank_out /* [n_queries] or NULL */, void* stream);
/* fp32 features -> bf16 operand rows for the scorer.  n_terms = 1: plain round-to-nearest bf16 [n, dim].
 * n_terms = 6: exact-product split x = hi+mid+lo (3 x bf16) laid out [n, 6*dim] as
 *   role 0 (query):   hi hi mid mid hi lo      role 1 (gallery): hi mid hi mid lo hi
 * so one bf16 GEMM over K = 6*dim reproduces the fp32 dot product to ~2^-23 (all products exact in fp32). */
int lpi_split_bf16(const float* x, int n, int dim, int n_terms, int role, void* out_bf16, void* stream);
/* x[n,dim] fp32 -> x/||x||_2 (no epsilon, models/slinet.py:122,133); in place allowed. */
int lpi_l2_normalize(const float* x, int n, int dim, float* out, float* norm_out /* [n] or NULL */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LPI_B200_H */
