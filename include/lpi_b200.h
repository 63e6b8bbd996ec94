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
    EPI_BIAS_F32 = 7,       /* out(fp32)  = acc + bias                                   (patch embedding)      */
    EPI_BIAS_GELU_F32 = 8,  /* out(fp32)  = QuickGELU(acc + bias); out2(fp32) = acc+bias  (c_fc, TF32 towers only)  */
    EPI_DGELU_F32 = 9       /* out(fp32)  = acc * QuickGELU'(aux fp32)                    (c_proj dgrad, TF32 only)  */
};
int lpi_gemm_bf16(const void* A, const void* B, int M, int N, int K, int epi, const void* bias_f32,
                  const void* resid_f32, void* out, void* out2, const void* aux_bf16, int ldo, int tile_n,
                  void* stream);
/* Same pipeline with fp32 operands in memory and tcgen05.mma kind::tf32 (10-bit mantissa, half the bf16 rate): the text tower
 * runs on it because its gradients sit at the 2e-2 parity limit with bf16 operands (SURVEY.md section 7 error budget).
 * A [M,K] fp32, B [N,K] fp32, K % 32 == 0.  Epilogues: BIAS_BF16, BIAS_RESID_F32, F32, BF16, BIAS_GELU_F32, DGELU_F32. */
int lpi_gemm_tf32(const void* A, const void* B, int M, int N, int K, int epi, const void* bias_f32, const void* resid_f32,
                  void* out, void* out2, const void* aux, int ldo, int tile_n, void* stream);
/* Same pipeline with fp16 operands (tcgen05.mma kind::f16, fp16 inputs): the 10-bit mantissa of TF32 at the full bf16 rate and half
 * the operand bytes -- the reference itself runs its CLIP weights in fp16 on the GPU (models/clip/model.py:394-415,522).  Every
 * "bf16" output / out2 / aux of the epilogue table is fp16 here.  N % 256 == 0 (CTA-pair tiles), K % 64 == 0.
 * Epilogues: BIAS_BF16, BIAS_GELU_BF16, BIAS_RESID_F32, F32, ACC_F32, DGELU_BF16, BF16, BIAS_F32. */
int lpi_gemm_f16(const void* A, const void* B, int M, int N, int K, int epi, const void* bias_f32, const void* resid_f32,
                 void* out, void* out2, const void* aux_f16, int ldo, int tile_n, void* stream);

/* out_proj dgrad fused with the attention backward's delta (autograd of nn.MultiheadAttention, models/clip/model.py:172,183-185):
 *   out[M, N] (16-bit) = A[M, K] . Wt[N, K]^T                      -- d loss / d (attention output), what lpi_gemm_* EPI_BF16 computes
 *   delta[b, h, l]    += sum_{d < 64} acc_fp32[b*L + l, 64 h + d] * o_saved[b*L + l, 64 h + d]
 * i.e. the row sums the softmax backward needs, taken from the fp32 accumulator in the epilogue (thread = row; two atomic adds per
 * (row, head) onto the caller-ZEROED delta [M/L, N/64, L], so the sum does not depend on their order).  Replaces one pass over dO and O
 * per block.  M % L == 0, N % 64 == 0; f16: 0 = bf16 operands / outputs, 1 = fp16. */
int lpi_gemm_do_delta(const void* A, const void* Wt, int M, int N, int K, void* out, const void* o_saved, float* delta, int L, int f16,
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
 * init_thr (optional, element stride init_thr_stride): per-query score s such that only items scoring >= s are kept.  Seeding it
 * with the k-th best score over ANY subset of the same gallery (e.g. the k-th column of a previous call on its first rows) is exact
 * -- at least k items reach it, so the true top-k all pass -- and removes most of the list warm-up (sorted insertions), which is a
 * fixed cost per launch and therefore what limits strong scaling over small shards.
 * ------------------------------------------------------------------------------------------------ */
int lpi_sim_topk_chunks(int n_queries, int n_gallery, int* n_chunks_out);
int lpi_sim_topk_bf16(const void* Q, const void* G, int n_queries, int n_gallery, int dim, int k,
                      long long gallery_offset, int n_chunks, const float* init_thr, int init_thr_stride,
                      float* part_scores, int* part_idx, void* stream);
/* Cooperative thresholds: shared_thr [n_queries] fp32 is read and written -- in: a score >= k gallery rows are known to reach per query
 * (or -inf); out: the best k-th score any chunk of the launch reached.  Same merged top-k as lpi_sim_topk_bf16, fewer sorted insertions
 * (the chunks of one launch exchange their k-th scores through it); per-chunk lists may hold < k entries (rest: -inf / INT_MAX). */
int lpi_sim_topk_coop_bf16(const void* Q, const void* G, int n_queries, int n_gallery, int dim, int k, long long gallery_offset,
                           int n_chunks, float* shared_thr, float* part_scores, int* part_idx, void* stream);
/* Threshold pre-pass for lpi_sim_topk_bf16: seed_scores [n_queries, k] = the k largest per-tile (256 rows) maxima over the first
 * n_rows gallery rows (seed_idx_ws [n_queries, k] is scratch).  `seed_scores + (k - 1)` with init_thr_stride = k is then a valid
 * init_thr: k distinct rows reach it.  One candidate per tile keeps the pass MMA-bound. */
int lpi_sim_topk_seed_bf16(const void* Q, const void* G, int n_queries, int n_rows, int dim, int k, float* seed_scores,
                           int* seed_idx_ws, void* stream);
/* the pre-pass in n_chunks work items per query tile (whole waves of clusters): outputs [n_chunks, n_queries, k]; merge, then take the k-th */
int lpi_sim_topk_seed_chunks_bf16(const void* Q, const void* G, int n_queries, int n_rows, int dim, int k, int n_chunks,
                                  float* seed_scores, int* seed_idx_ws, void* stream);
/* k-way merge of n_parts partial lists (chunks and/or all-gathered shards) -> [n_queries, k]. */
int lpi_topk_merge(const float* part_scores, const int* part_idx, int n_parts, int n_queries, int k,
                   float* out_scores, int* out_idx, void* stream);
/* The whole tail of a gallery-sharded search step in one launch: k-way merge of chunk / rank lists + Recall@K bookkeeping
 * (sprompt.py:559-619).  Parts may be grouped (one group per rank of an all-gathered exchange buffer): part p is read at
 * (p / parts_per_group) * group_stride + (p % parts_per_group) * n_queries * k elements from part_scores / part_idx.
 * counts[n_tasks,4] (zeroed here) = #{rank<1}, #{rank<5}, #{rank<10}, n per task; rank_out[n_queries] optional (k = not in the list). */
int lpi_topk_merge_recall(const float* part_scores, const int* part_idx, int n_parts, int parts_per_group, long long group_stride,
                          int n_queries, int k, float* out_scores, int* out_idx, const int* gt_ptr, const int* gt_idx,
                          const int* task_of_query, int n_tasks, int* counts, int* rank_out, void* stream);
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

/* ------------------------------------------------------------------------------------------------
 * Fused multi-head attention (flash-style; scores never leave registers).
 * replaces: nn.MultiheadAttention inside ResidualAttentionBlock.attention, models/clip/model.py:172,183-185
 *           (+ the causal mask built at model.py:347-353) and its autograd backward.
 * qkv [B*L, 3*H*64] bf16 (row = b*L + l; columns q | k | v), out / d_out [B*L, H*64] bf16, lse2 [B*H*L] fp32
 * (log2-domain log-sum-exp saved by the forward for the backward), delta_ws [B*H*L] fp32 scratch, dqkv like qkv.
 * lpi_attn_bwd / lpi_attn_bwd_f16 with out == NULL: delta_ws is an INPUT holding delta[b,h,l] = sum_d d_out * out, as lpi_gemm_do_delta
 * leaves it (no separate delta pass).
 * ------------------------------------------------------------------------------------------------ */
int lpi_attn_fwd(const void* qkv, void* out, float* out_f32 /* optional fp32 copy of out */, float* lse2, int B, int L, int H,
                 int causal, void* stream);
int lpi_attn_bwd(const void* qkv, const void* out, const void* d_out, const float* lse2, float* delta_ws, void* dqkv,
                 float* dqkv_f32 /* if non-NULL the gradient is written here in fp32 instead of dqkv */, int B, int L, int H,
                 int causal, void* stream);
/* fp16 storage for qkv / out / d_out / dqkv (and P, dS inside the kernels); L <= 256.  The backward is linear in d_out, so a
 * caller that scales d_out by a power of two (fp16 gradient scaling) gets dqkv scaled by the same factor. */
int lpi_attn_fwd_f16(const void* qkv, void* out, float* lse2, int B, int L, int H, int causal, void* stream);
int lpi_attn_bwd_f16(const void* qkv, const void* out, const void* d_out, const float* lse2, float* delta_ws, void* dqkv, int B, int L,
                     int H, int causal, void* stream);
/* Attention of a tower's LAST block for the one query row per sample that the head reads: ln_post(x[:, 0, :]) in
 * VisionTransformer.forward (models/clip/model.py:254-257), x[arange(B), tokenized.argmax(-1)] in TextEncoder.forward
 * (models/clip/prompt_learner.py:57-61).  rows int32 [B] = global row (b*L + position) of each sample's read row; keys / values of
 * all L positions (up to the row's own position when causal), L <= 512.  f16: 0 = bf16 storage, 1 = fp16.
 * fwd: out_rows [B, H*64] = softmax(q K^T / 8) V of those rows; optionally x_rows [B, H*64] = x[rows] (fp32 residual rows gathered
 *      by the same launch; x and x_rows are given together or both NULL).
 * bwd: dqkv [B*L, 3*H*64] is written completely (dq zero outside the read rows, dk / dv zero beyond the causal horizon), linear in
 *      dout_rows [B, H*64]; optionally g[rows] = g_rows (fp32 [B, H*64] scattered into the caller-zeroed [B*L, H*64] stream). */
int lpi_attn_rowq_fwd(const void* qkv, const int* rows, void* out_rows, const float* x, float* x_rows, int B, int L, int H, int causal,
                      int f16, void* stream);
int lpi_attn_rowq_bwd(const void* qkv, const int* rows, const void* dout_rows, void* dqkv, const float* g_rows, float* g, int B, int L,
                      int H, int causal, int f16, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm in fp32 (models/clip/model.py:154-160; eps inside the sqrt, biased variance).
 * fwd: x [M,D] fp32 -> out_f32 and/or out_bf16 (either may be NULL).
 * bwd: g = (accumulate ? g : 0) + dLN(dy; x, gamma); optional bf16 shadow of g (A operand of the next dgrad GEMM).
 * D in {128, 256, 512, 768, 1024}.
 * ------------------------------------------------------------------------------------------------ */
int lpi_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* out_f32, void* out_bf16, long long M, int D,
                      float eps, void* stream);
int lpi_layernorm_bwd(const float* dy, const float* x, const float* gamma, float* g, void* g_bf16, long long M, int D, float eps,
                      int accumulate, void* stream);
/* fp16 shadows for the fp16 GEMM path.  The fp16 gradient stream is carried multiplied by grad_scale (a power of two, so the
 * scaling is exact) to keep small gradients inside fp16's normal range: dy_scaled = grad_scale * dy, g stays true-scale fp32,
 * g_f16 = fp16(grad_scale * g). */
int lpi_layernorm_fwd_f16(const float* x, const float* gamma, const float* beta, float* out_f32, void* out_f16, long long M, int D,
                          float eps, void* stream);
int lpi_layernorm_bwd_f16(const float* dy_scaled, const float* x, const float* gamma, float* g, void* g_f16, long long M, int D,
                          float eps, int accumulate, float grad_scale, void* stream);
/* dy in the tower's 16-bit type, as written by the dgrad GEMM's EPI_BF16 epilogue (bf16: is_f16 = 0, grad_scale ignored; fp16:
 * is_f16 = 1, dy carries grad_scale): halves the largest stream of this HBM-bound kernel. */
int lpi_layernorm_bwd_dy16(const void* dy16, int is_f16, const float* x, const float* gamma, float* g, void* g16, long long M, int D,
                           float eps, int accumulate, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Vision front end (VisionTransformer.forward, models/clip/model.py:227-250).
 * im2col: images [B,3,R,R] fp32 -> patch rows [B*(R/P)^2, 3*P*P] bf16 in conv1.weight.view(D,-1) column order, so that
 *         conv1 (kernel = stride = P, no bias) is one lpi_gemm_bf16.
 * assemble: row 0 = class_embedding + pos[0]; rows 1..P = prompt_table[sel[b]] (NO positional term, model.py:240-248);
 *           rows P+1.. = patch_emb + pos[1..]; then ln_pre.  P = 0 / prompt_table NULL = un-prompted CLIP (extract_vector).
 *           prompt_table [n_tables, P, D] fp32, sel [B] int32 or NULL (= table 0 for every sample).
 * assemble_bwd: d_prompt[t,p,:] = sum_{b: sel[b]=t} dLN(g[b,1+p,:]; prompt_table[t,p,:], ln_pre.weight)   (overwrites d_prompt)
 * ------------------------------------------------------------------------------------------------ */
int lpi_im2col_patches(const float* images, void* out_bf16, int B, int resolution, int patch, void* stream);
int lpi_im2col_patches_f16(const float* images, void* out_f16, int B, int resolution, int patch, void* stream);   /* fp16 vision tower */
int lpi_assemble_vision(const float* patch_emb, const float* cls, const float* pos, const float* prompt_table, const int* sel,
                        const float* ln_gamma, const float* ln_beta, float* x_out, int B, int n_patch, int P, int D, float eps,
                        void* stream);
/* The same assembly with the prompt rows reconstructed in the kernel from the DecomposedPrompt factors (prompts.py:38-57 fused with
 * model.py:240-248: no [L,P,D] table is materialised for the token sequence).  dim1_share [T, n_layers, r], dim2 [T, P, r], dim3 [T, D, r]
 * for the T selectable tasks (sel[b] picks one; NULL = task 0); layer 0 enters the sequence; bit-identical to lpi_prompt_fwd + assemble. */
int lpi_assemble_vision_factors(const float* patch_emb, const float* cls, const float* pos, const float* dim1_share, const float* dim2_vis,
                                const float* dim3_vis, int r, int n_layers, float scale, const int* sel, const float* ln_gamma,
                                const float* ln_beta, float* x_out, int B, int n_patch, int P, int D, float eps, void* stream);
int lpi_assemble_vision_factors_bwd(const float* g, const float* dim1_share, const float* dim2_vis, const float* dim3_vis, int r, int n_layers,
                                    float scale, const int* sel, const float* ln_gamma, float* d_prompt, int B, int L, int P, int n_tables,
                                    int D, float eps, void* stream);
int lpi_assemble_vision_bwd(const float* g, const float* prompt_table, const int* sel, const float* ln_gamma, float* d_prompt,
                            int B, int L, int P, int n_tables, int D, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Text front end (PromptLearner.forward splice + TextEncoder positional add,
 * models/clip/prompt_learner.py:133-163, 52-53): x[b,l] = (1 <= l <= P ? ctx_table[sel[b]][l-1] : token_embedding[tok[b,l]]) + pos[l].
 * ctx_table NULL = extract_vector path (raw "X" embeddings, prompt_learner.py:118-126).  tokens int64 [B,L].
 * bwd: d_ctx[t,p,:] = sum_{b: sel[b]=t} g[b,1+p,:]   (overwrites d_ctx)
 * ------------------------------------------------------------------------------------------------ */
int lpi_assemble_text(const float* token_embedding, const long long* tokens, const float* pos, const float* ctx_table,
                      const int* sel, float* x_out, int B, int L, int P, int D, void* stream);
/* text splice with the context rows reconstructed from the factors (prompts.py:38-57 fused with prompt_learner.py:152-163, 53) */
int lpi_assemble_text_factors(const float* token_embedding, const long long* tokens, const float* pos, const float* dim1_share,
                              const float* dim2_txt, const float* dim3_txt, int r, int n_layers, float scale, const int* sel, float* x_out,
                              int B, int L, int P, int D, void* stream);
int lpi_assemble_text_bwd(const float* g, const int* sel, float* d_ctx, int B, int L, int P, int n_tables, int D, void* stream);
/* Opt-in deep-prompt injection (the intended semantics of the dead branch at models/clip/model.py:190-193):
 * x[b, 1+p, :] += prompt[sel[b], p, :] before block `layer`; its backward is lpi_assemble_text_bwd on the block-input gradient. */
int lpi_inject_prompt_rows(float* x, const float* prompt, const int* sel, int B, int L, int P, int D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Encoder heads: feat[b] = normalize(LN(x[row_idx[b]]) @ proj), proj [D,E] fp32 row-major
 * (ln_post + proj, model.py:254-257; ln_final + EOT gather + text_projection, prompt_learner.py:57-61; L2 norm slinet.py:122,133).
 * z_out [B,E] = un-normalised features (kept for the backward).
 * bwd: g[row_idx[b], :] = d/dx given dfeat (gradient wrt the normalised feature) and/or dz_direct (gradient wrt the raw
 *      projection z); either may be NULL.  Rows are ASSIGNED; the caller zeroes g first.
 * ------------------------------------------------------------------------------------------------ */
int lpi_head_fwd(const float* x, const int* row_idx, const float* ln_gamma, const float* ln_beta, const float* proj, float* z_out,
                 float* feat_out, int B, int D, int E, float eps, void* stream);
/* head_fwd that also ends in the task-id selection of every sample (sprompt.py:336-368): sel_out[b] = argmin_t min_c |feat[b] - centers[t,c]|_1,
 * centers [n_tasks, n_centers, E]; the un-prompted pass of the evaluation needs no separate nearest-centre launch. */
int lpi_head_fwd_select(const float* x, const int* row_idx, const float* ln_gamma, const float* ln_beta, const float* proj, float* z_out,
                        float* feat_out, const float* centers, int n_tasks, int n_centers, int* sel_out, int B, int D, int E, float eps,
                        void* stream);
int lpi_head_bwd(const float* dfeat, const float* dz_direct, const float* z, const float* x, const int* row_idx, const float* ln_gamma,
                 const float* proj, float* g, void* g_bf16, int B, int D, int E, float eps, void* stream);
/* same with g_f16 = fp16(grad_scale * g) as the shadow (fp16 gradient path, see lpi_layernorm_bwd_f16) */
int lpi_head_bwd_f16(const float* dfeat, const float* dz_direct, const float* z, const float* x, const int* row_idx, const float* ln_gamma,
                     const float* proj, float* g, void* g_f16, float grad_scale, int B, int D, int E, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DecomposedPrompt (models/prompts/prompts.py:38-57): Y[l,p,d] = mean_k dim1[l,k] dim2[p,k] dim3[d,k], dim1 shared by both
 * modalities.  bwd consumes dense upstream gradients g_vis [L,P,Dv], g_txt [L,P,Dt]; ws = 2*L*P*r floats.
 * ------------------------------------------------------------------------------------------------ */
int lpi_prompt_fwd(const float* dim1_share, const float* dim2_vis, const float* dim2_txt, const float* dim3_vis, const float* dim3_txt,
                   float* vis_out, float* txt_out, int L, int P, int Dv, int Dt, int r, void* stream);
int lpi_prompt_bwd(const float* dim1_share, const float* dim2_vis, const float* dim2_txt, const float* dim3_vis, const float* dim3_txt,
                   const float* g_vis, const float* g_txt, float* ws, float* d_dim1, float* d_dim2_vis, float* d_dim2_txt,
                   float* d_dim3_vis, float* d_dim3_txt, int L, int P, int Dv, int Dt, int r, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Losses (models/slinet.py:137-183, loss/loss.py:6-33,75-87) and optimiser (methods/sprompt.py:253-254).
 * sgemm: C[m,n] = alpha * sum_k A[m*a_m + k*a_k] * B[k*b_k + n*b_n] + beta * C  -- exact fp32 products for the B x B logits,
 *        d logits -> d features, the 9 x 9 alignment logits.
 * clip_loss_logits: loss = weight * 1/2 [CE(S, arange) + CE(S^T, arange)]; dlogits (optional) = d loss / d S.  lse_ws = 2n floats.
 * task_loss: nt_bxent_loss over the rows of X [R, n] (flattened prompts of tasks 0..R-1, R <= 16) including the reference's
 *        double sigmoid; loss_out (+)= weight * loss; grad_last_row [n] (+)= d/dX[R-1] (the only trainable row).
 *        part_ws = n_part*R*R floats, coef_out = R floats.
 * ------------------------------------------------------------------------------------------------ */
int lpi_sgemm_f32(const float* A, const float* B, float* C, int M, int N, int K, long long a_m, long long a_k, long long b_k,
                  long long b_n, long long ldc, float alpha, float beta, void* stream);
int lpi_clip_loss_logits(const float* logits, int n, float weight, float* lse_ws, float* loss_out, float* dlogits, void* stream);
/* Fused similarity + InfoNCE, forward AND backward, in ONE (cooperative) launch (slinet.py:138-141 + loss.py:75-87 + autograd):
 * img_f / txt_f [n, E] fp32 L2-normalised features of the GLOBAL batch; loss_out = weight * ClipLoss(scale * I T^T);
 * d_img / d_txt [n_local, E] = gradients w.r.t. rows [row0, row0 + n_local) of I / T (the rows this rank's towers produced; NULL = skip);
 * logits_out [n, n] only if non-NULL -- the score matrix is otherwise never written; lse_ws, terms_ws: 2n floats of scratch each. */
int lpi_sim_infonce_fwd_bwd(const float* img_f, const float* txt_f, int n, int E, float scale, float weight, int row0, int n_local,
                            float* lse_ws, float* terms_ws, float* loss_out, float* logits_out, float* d_img, float* d_txt, void* stream);
int lpi_row_mean(const float* x, float* out, int rows, int D, float scale, void* stream);
int lpi_add_rowconst(float* G, const float* v, long long rows, int D, float alpha, int accumulate, void* stream);
int lpi_task_loss(const float* X, int R, long long n, const int* target, float temperature, float weight, float* part_ws, int n_part,
                  float* loss_out, int loss_accumulate, float* coef_out, float* grad_last_row, int grad_accumulate, void* stream);
int lpi_sgd_momentum_step(float* w, const float* g, float* v, long long n, float lr, float momentum, float weight_decay,
                          int first_step, void* stream);
/* sel[b] = argmin_t min_c sum_d |f[b,d] - centers[t,c,d]| (methods/sprompt.py:336-368); centers [T,C,E]; sel int64 [B]. */
int lpi_nearest_center_l1(const float* feats, const float* centers, int B, int n_tasks, int n_centers, int E, long long* sel_out,
                          void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32 parity mode of the towers (north_star: "1e-5 in fp32"; reference = the un-converted fp32 model, clip.py:128-129,
 * models/clip/model.py:154-196).  Exact fp32 products on the SIMT pipes -- a TEST mode selected with precision="fp32" on the engines:
 *   linear layers       lpi_sgemm_bias_f32  (C = alpha A B + bias[n] + beta C; arbitrary strides like lpi_sgemm_f32)
 *   patch extraction    lpi_im2col_patches_f32
 *   attention           lpi_attn_fwd_f32 / lpi_attn_bwd_f32 on qkv [B*L, 3*H*64] fp32; lse = natural-log sum-exp [B*H*L];
 *                       delta_ws [B*H*L] scratch; dqkv [B*L, 3*H*64] is overwritten
 *   QuickGELU           lpi_quick_gelu_f32 / lpi_quick_gelu_bwd_f32 (out = dy * QuickGELU'(z)), model.py:163-165
 * LayerNorm, token assembly, heads, prompt and loss kernels are fp32 already.
 * ------------------------------------------------------------------------------------------------ */
int lpi_sgemm_bias_f32(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, long long a_m, long long a_k,
                       long long b_k, long long b_n, long long ldc, float alpha, float beta, void* stream);
int lpi_im2col_patches_f32(const float* images, float* out_f32, int B, int resolution, int patch, void* stream);
int lpi_attn_fwd_f32(const float* qkv, float* out, float* lse, int B, int L, int H, int causal, void* stream);
int lpi_attn_bwd_f32(const float* qkv, const float* out, const float* d_out, const float* lse, float* delta_ws, float* dqkv, int B, int L,
                     int H, int causal, void* stream);
int lpi_quick_gelu_f32(const float* z, float* out, long long n, void* stream);
int lpi_quick_gelu_bwd_f32(const float* dy, const float* z, float* out, long long n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Exchange steps of the multi-GPU path (reference anchor: gather_features, methods/sprompt.py:38-82).
 * One process per GPU; NCCL is resolved at run time (a copy already loaded into the process, $LPI_NCCL_LIB, or the system
 * libnccl.so.2).  The opaque communicator is the only state the library keeps.
 *   lpi_comm_unique_id : rank 0 fills lpi_comm_unique_id_bytes() bytes, the host ships them to the other ranks
 *   lpi_comm_init      : collective; binds the CURRENT CUDA device of each process
 *   lpi_comm_allgather : recv[n_ranks * bytes_per_rank] <- every rank's send[bytes_per_rank], rank-major: the ONE exchange of a sharded
 *                        search step (each rank's [2, c, Q, k] candidate buffer, see lpi_topk_merge_recall) or of a data-parallel
 *                        training step (the [b, 2E] image|text feature rows, see lpi_sim_infonce_fwd_bwd)
 *   lpi_comm_allreduce_sum_f32 : the flat 5 284-float prompt gradient of a data-parallel step
 * The Python layer uses torch.distributed for the same two collectives (same NCCL underneath); lpi_b200/comm.py wraps these entry points.
 * ------------------------------------------------------------------------------------------------ */
int lpi_comm_unique_id_bytes(void);
int lpi_comm_unique_id(void* id_out);
int lpi_comm_init(void** comm_out, int n_ranks, int rank, const void* unique_id);
int lpi_comm_allgather(void* comm, const void* send, void* recv, long long bytes_per_rank, void* stream);
int lpi_comm_allreduce_sum_f32(void* comm, const float* send, float* recv, long long n, void* stream);
int lpi_comm_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* LPI_B200_H */
