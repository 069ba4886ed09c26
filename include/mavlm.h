/* mavlm.h -- C ABI of the B200 (sm_100a) visual-memory path.
 *
 * Drop-in boundary for the hot path of 1023604540/Memory-Augmented-VLM:
 *   mm_projector -> get_2dPool (bilinear 27x27->14x14) -> temporal PE -> chunked recurrent
 *   memory (formation + evolution) -> memory fuser -> token assembly.
 * The reference has no FFI of its own: its boundary is a set of Python nn.Module call sites
 * (llava/model/llava_arch.py:302, 495, 511, 530-537, 546, 550-551).  Each entry point below
 * names the reference call site it replaces; the Python host in
 * `memory-augmented-vlm_b200/` keeps the reference module signatures / state_dict keys and
 * binds these symbols with ctypes.
 *
 * Conventions
 *   - plain C: pointers are CUDA device pointers unless named host_*; sizes are element counts.
 *   - `stream` is a cudaStream_t passed as void* (the caller's current stream); nothing in this
 *     library synchronises, allocates persistent device memory, throws, or calls exit().
 *   - return value: 0 on success, negative MAVLM_E_* otherwise; mavlm_last_error_string() gives
 *     the message for the calling thread.
 *   - dtype: MAVLM_F32 runs the exact fp32 SIMT tier (parity <= 1e-5 vs the fp64 oracle),
 *     MAVLM_BF16 runs the TMA + tcgen05/TMEM tier (bf16 operands, fp32 accumulate/softmax/LN);
 *     MAVLM_F16 (the reference inference loader's default dtype, builder.py:27) runs the same tier with fp16
 *     operands / outputs (forward entry points only).
 *   - there is no CPU fallback and no other-architecture path: a non-sm_100 device is an error.
 */
#ifndef MAVLM_H_
#define MAVLM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAVLM_VERSION 100 /* 0.1.0 */
#define MAVLM_API __attribute__((visibility("default")))

enum { MAVLM_F32 = 0, MAVLM_BF16 = 1, MAVLM_F16 = 2 };
enum { MAVLM_ACT_NONE = 0, MAVLM_ACT_GELU_ERF = 1, MAVLM_ACT_RELU = 2 };
enum { MAVLM_POOL_BILINEAR = 0, MAVLM_POOL_AVERAGE = 1, MAVLM_POOL_MAX = 2 };
enum {
  MAVLM_OK = 0,
  MAVLM_E_INVALID = -1,  /* bad argument (shape / alignment / dtype)            */
  MAVLM_E_CUDA = -2,     /* CUDA runtime / driver error (launch failure, ...)    */
  MAVLM_E_ARCH = -3,     /* device is not sm_100 (B200)                          */
  MAVLM_E_INDEX = -4,    /* index out of range (PE frame index), host-validated  */
  MAVLM_E_WORKSPACE = -5 /* workspace too small                                   */
};

MAVLM_API int mavlm_version(void);
MAVLM_API const char* mavlm_last_error_string(void);
/* Number of CUDA kernels this library has launched in this process (monotonic; for accounting). */
MAVLM_API unsigned long long mavlm_launch_count(void);
/* Verifies that `device` is compute capability 10.x; MAVLM_E_ARCH otherwise. */
MAVLM_API int mavlm_check_device(int device);

/* ---- a2 + a3: get_2dPool (llava_arch.py:277-297) fused with TemporalPositionalEncoding
 * (position_encoding.py:57-64).  x [F, side*side, D] -> y [F, out_side*out_side, D];
 * bilinear, align_corners=False.  pe_table (fp32 [max_frames, D]) and frame_idx (int64 [F], device)
 * may both be NULL for pooling alone.  Indices are validated by the caller on the host
 * (position_encoding.py:73-76 raises ValueError there).  mode average/max are the reference's two
 * other (shape-incompatible downstream) modes with kernel = stride. */
MAVLM_API int mavlm_pool_pe_fwd(const void* x, void* y, const float* pe_table, const int64_t* frame_idx, int frames, int side,
                      int out_side, int stride, int dim, int mode, int dtype, void* stream);

/* ---- a1 (second projector layer) + a3 fused: C = A W^T + bias + pe_table[frame_idx[row / tokens_per_frame]]
 * (builder.py:47 followed by position_encoding.py:57-64 when the spatial pool runs before W2): the PE add rides in
 * the GEMM epilogue instead of a separate read-modify-write pass over [F, 196, D].  bf16 / fp16 tier. */
MAVLM_API int mavlm_gemm_bias_pe_fwd(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias,
                                     const float* pe_table, const int64_t* frame_idx, int tokens_per_frame, void* C,
                                     int64_t ldc, int M, int N, int K, int dtype, void* stream);

/* ---- a3 alone: x [T, N, C] + pe_table[frame_idx[t]] cast to dtype (position_encoding.py:57-64). */
MAVLM_API int mavlm_add_pe_fwd(const void* x, void* y, const float* pe_table, const int64_t* frame_idx, int frames, int tokens,
                     int dim, int dtype, void* stream);

/* ---- nn.Linear (+activation / +residual) used by mm_projector (builder.py:41-48), q/k/v/dense
 * projections (MemoryController.py:23,37-39), the RMT MLP (:63-67) and the fuser (llava_arch.py:132-136):
 *   C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]) (+ resid[M,N]) (+ addvec[N])
 * A, W, resid in `dtype`; bias/addvec in `dtype`; C in out_dtype (MAVLM_F32 allowed with bf16 inputs:
 * used for the pre-LayerNorm sum).  ld* are row strides in elements.  bias/resid/addvec may be NULL.
 */
MAVLM_API int mavlm_gemm_bias_act_fwd(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* resid,
                            int64_t ldr, const void* addvec, void* C, int64_t ldc, int M, int N, int K, int act,
                            int dtype, int out_dtype, void* stream);

/* ---- co-scheduled nn.Linear calls ("tail fill").  The recurrence (MemoryController.py:118-158) is a serial chain of
 * GEMMs with 1568 rows, each of which leaves a third of the tile slots of its last wave empty, while other nn.Linear
 * calls of the same step wait for nothing on that chain: the frame-side k_proj / v_proj of LATER chunks
 * (MemoryController.py:49-50) and the memory_fuser of FINISHED states (llava_arch.py:545-546).  mavlm_gemm_fill_fwd
 * runs one critical-path nn.Linear (`primary`, whole) and, in the SM time its last wave leaves idle, tiles
 * [fill_begin, ...) of a second one (`filler`); mavlm_gemm_tiles_fwd runs a tile range of one problem on its own (the
 * rest of a filler before its first consumer).  Tiles are 256 x 256 outputs in a fixed rasterised order
 * (mavlm_gemm_num_tiles of them); results are bit-identical to mavlm_gemm_bias_act_fwd with the same tile.
 * bf16 / fp16 tier only.  Same epilogue semantics as mavlm_gemm_bias_act_fwd / mavlm_gemm_bias_pe_fwd. */
typedef struct mavlm_gemm_desc {
  const void* A;        int64_t lda;   /* [M, K] */
  const void* W;        int64_t ldw;   /* [N, K] (nn.Linear weight) */
  const void* bias;                    /* [N] or NULL */
  const void* resid;    int64_t ldr;   /* [M, N] or NULL */
  const void* addvec;                  /* [N] or NULL */
  const float* pe_table; const int64_t* frame_idx; int32_t tokens_per_frame;   /* fused temporal PE or NULL */
  void* C;              int64_t ldc;   /* [M, N] */
  int32_t M, N, K, act, out_dtype;
} mavlm_gemm_desc;
MAVLM_API int mavlm_gemm_num_tiles(const mavlm_gemm_desc* g);
MAVLM_API int mavlm_gemm_tiles_fwd(const mavlm_gemm_desc* g, int tile_begin, int tile_end, int dtype, void* stream);
/* *fill_done_end receives the first filler tile NOT computed (== fill_begin when nothing fitted). */
MAVLM_API int mavlm_gemm_fill_fwd(const mavlm_gemm_desc* primary, const mavlm_gemm_desc* filler, int fill_begin,
                                  int fill_avail_end, int* fill_done_end, int dtype, void* stream);

/* ---- tensor.to(dtype) between the two tiers (fp32 <-> bf16), contiguous n elements. */
MAVLM_API int mavlm_cast_fwd(const void* x, void* y, int64_t n, int src_dtype, int dst_dtype, void* stream);

/* ---- nn.LayerNorm over the last dim (MemoryController.py:24,28): y = (x-mean)/sqrt(var+eps)*g+b.
 * x in x_dtype (MAVLM_F32 pre-LN sums or `dtype`), gamma/beta and y in dtype.  x 16-byte aligned (rows are fetched with
 * cp.async.bulk).  gamma / beta are module PARAMETERS: the kernel stages them before it waits for the launch ahead of
 * it on the stream (programmatic dependent launch), so they must not be the output of one of this library's own
 * launches still in flight on `stream` (tensors written by torch kernels or copies are fine: those complete first). */
MAVLM_API int mavlm_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, int rows, int dim, float eps,
                        int x_dtype, int dtype, void* stream);

/* ---- multi-head softmax(q k^T * scale) v of Attention.forward (MemoryController.py:42-54), flash-style,
 * probs never materialised in the bf16 tier.  Q [B, Lq, H*dh] (row stride ldq), K/V [B, Lk, H*dh]
 * (row strides ldk/ldv, batch strides in elements), O [B, Lq, H*dh].  lse (fp32 [B,H,Lq], natural log) may be
 * NULL.  col_scores (fp32 [B, Lk]) if non-NULL receives sum over heads and queries of the normalised
 * probabilities (MemoryController.py:135; fp32 tier only -- the bf16 / fp16 tier gets them from mavlm_xattn_colsum).
 * workspace: see mavlm_xattn_workspace_bytes. */
MAVLM_API size_t mavlm_xattn_workspace_bytes(int batch, int heads, int lq, int lk, int head_dim, int dtype);
MAVLM_API int mavlm_xattn_fwd(const void* Q, int64_t ldq, int64_t q_batch_stride, const void* K, int64_t ldk,
                    int64_t k_batch_stride, const void* V, int64_t ldv, int64_t v_batch_stride, void* O, int64_t ldo,
                    int64_t o_batch_stride, float* lse, float* col_scores, int batch, int heads, int lq, int lk,
                    int head_dim, float scale, int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* ---- frame scores of TransformerProjector.forward (MemoryController.py:135-139: attn_probs.sum(heads).sum(queries)) in
 * the bf16 / fp16 tier, where the fused attention never forms the probabilities: a second pass over K with the LSE that
 * mavlm_xattn_fwd saved.  col_scores fp32 [B, Lk] = sum over heads h and queries q of exp(q.k * scale - lse[b,h,q]); every
 * row of it sums to heads * lq.  One batched tcgen05 GEMM (keys x queries per head) with an exp2 + row-sum epilogue:
 * half the MMA work of the attention call itself. */
MAVLM_API int mavlm_xattn_colsum(const void* Q, int64_t ldq, int64_t q_batch_stride, const void* K, int64_t ldk,
                                 int64_t k_batch_stride, const float* lse, float* col_scores, int batch, int heads, int lq,
                                 int lk, int head_dim, float scale, int dtype, void* stream);

/* ---- a14 + a15: type embeddings + token assembly (llava_arch.py:548-554, 620-629, 708-731).
 * Writes seq [10 + n_mem_rows + 1 + 9 + n_fine*tokens + 1, D]:
 *   rows of embed_table at prompt_mem_ids | mem (+type_emb[0]) | newline | prompt_frm rows |
 *   frames[fine_idx] (+type_emb[1]) | newline.
 * mem may be NULL when the fuser GEMM already wrote that segment in place (addvec = type_emb[0]).
 * prompt ids / fine_idx are device int64.  drop_frames != 0 stops after the memory newline. */
MAVLM_API int mavlm_assemble_fwd(void* seq, const void* mem, int64_t n_mem_rows, const void* frames, const int64_t* fine_idx,
                       int n_fine, int tokens, const void* type_emb, const void* newline, const void* embed_table,
                       const int64_t* prompt_mem_ids, int n_prompt_mem, const int64_t* prompt_frm_ids,
                       int n_prompt_frm, int dim, int drop_frames, int dtype, void* stream);

/* ---- text / vision splice + padding (llava_arch.py:745-878), the consumer of the assembled sequence: every output
 * row copies one source row.  row_src (device int64 [n_rows]): >= 0 -> embed_table[row_src] (embed_tokens of a text
 * token), -1 -> zeros (padding), <= -2 -> feats[-(row_src + 2)] (a row of the video token sequence).  embed_table has
 * n_table_rows rows and feats n_feat_rows rows, both of `dtype`; a source row outside its table is written as zeros
 * (never read): the host validates token ids first and raises like the reference's embedding lookup. */
MAVLM_API int mavlm_gather_rows_fwd(void* out, int64_t ld_out, const void* embed_table, int64_t n_table_rows,
                                    const void* feats, int64_t n_feat_rows, const int64_t* row_src, int64_t n_rows,
                                    int dim, int dtype, void* stream);

/* ---- frame pre-processing, the producer side of the path: decoded uint8 frames [F, H, W, 3] (the
 * `.pt` video tensor of extract_video_frames/video_reader_tmp.py:87, fed at train.py:1239) ->
 * SigLipImageProcessor.preprocess (siglip_encoder.py:47-67): PIL bicubic resize to `size`, rescale 1/255, normalize,
 * channels-first -> pixel_values [F, 3, out_h, out_w] (f32, or bf16 = the tower's `.to(dtype)`, siglip_encoder.py:585).
 * Bit-exact with Pillow's two-pass 22-bit fixed-point resampler (uint8 intermediate, horizontal pass first, a pass
 * skipped when that dimension already matches) and transformers' rescale / normalize arithmetic.
 *
 * mavlm_resize_coeffs (host, no CUDA): Pillow's precompute_coeffs + normalize_coeffs_8bpc for one axis.
 *   bounds int32 [out, 2] = (first input index, tap count), kk int32 [out, ksize_capacity].  Returns ksize (the
 *   number of columns needed) when bounds / kk are NULL or on success, negative on error.
 * mavlm_frames_preprocess_fwd: the tables are DEVICE pointers (NULL for an axis whose size already matches);
 *   tmp: uint8 [F, in_h, out_w, 3] workspace (needed when the width changes); resized_u8: optional uint8
 *   [F, out_h, out_w, 3] copy of the resized frames (NULL to skip); mean3 / std3: HOST float[3]. */
MAVLM_API int mavlm_resize_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* kk, int ksize_capacity);
MAVLM_API int mavlm_frames_preprocess_fwd(const uint8_t* frames, int n_frames, int in_h, int in_w, uint8_t* tmp,
                                          uint8_t* resized_u8, void* pixel_values, int out_h, int out_w,
                                          const int32_t* bounds_h, const int32_t* kk_h, int ksize_h,
                                          const int32_t* bounds_v, const int32_t* kk_v, int ksize_v, double rescale,
                                          const float* mean3, const float* std3, int out_dtype, void* stream);

/* ======================= legacy memories and scene segmentation =======================
 * The Flash-VStream-style memories (memory_module/compress_functions.py, memory_builder.py:41-190) and the scene
 * segmentation (memory_module/segment.py) that sit beside the recurrent memory in the reference tree. */

/* Streaming compression of a video x [n_frames, row_elems = P*D] to `keep` frames:
 *   mode 0 drop_feature   (compress_functions.py:20-56)    coins: 1 = drop the right member of the most similar pair
 *   mode 1 merge_feature  (:59-91)
 *   mode 2 k_drop_feature (:176-218)                       coins: 1 = drop the row index of the flat argmax
 *   mode 3 k_merge_feature(:221-264)
 * One launch per streamed frame, all decisions on the device.  coins: DEVICE uint8 [n_frames - keep] (the host
 * draws them like the reference: one random.randint(0, 1) per frame; NULL for the merge modes).
 * out [keep, row_elems]; out_sim fp32: [keep - 1] adjacent similarities (modes 0, 1) or [keep, keep] (modes 2, 3),
 * may be NULL; decisions int32 [n_frames - keep, 2]: (idx, idx) / (idx, idx + 1) / (left, right) per frame, from
 * which the host rebuilds the reference's step_indices.  Needs n_frames > keep (the reference returns shorter
 * videos unchanged), 1 <= keep <= 64. */
MAVLM_API size_t mavlm_stream_compress_workspace_bytes(int64_t row_elems, int keep, int mode, int dtype);
MAVLM_API int mavlm_stream_compress_fwd(const void* x, int64_t n_frames, int64_t row_elems, int keep, int mode,
                                        const uint8_t* coins, void* out, float* out_sim, int32_t* decisions,
                                        void* workspace, size_t workspace_bytes, int dtype, void* stream);
/* The same for a batch of independent videos in one launch sequence (grid.z = video; one launch per frame index up to
 * the longest video): x_ptrs DEVICE array of `batch` base pointers, n_frames HOST array, coins [batch, coin_stride],
 * decisions [batch, coin_stride, 2] with coin_stride >= longest - keep, out [batch, keep, row_elems], out_sim
 * [batch, keep - 1] (at least 1) or [batch, keep, keep].  Frames and decisions are those of `batch` single calls
 * (similarities agree to fp32 rounding: the number of partial sums per row pair depends on the batch). */
MAVLM_API size_t mavlm_stream_compress_batched_workspace_bytes(int batch, int64_t row_elems, int keep, int mode, int dtype);
MAVLM_API int mavlm_stream_compress_batched_fwd(const void* const* x_ptrs, const int64_t* n_frames, int batch,
                                                int64_t row_elems, int keep, int mode, const uint8_t* coins,
                                                int64_t coin_stride, void* out, float* out_sim, int32_t* decisions,
                                                void* workspace, size_t workspace_bytes, int dtype, void* stream);
/* out[t, :] = mean over the tokens of frame t: the `features.mean(dim=1)` feeding segment() (segment.py:266) and the
 * scheduler (llava_arch.py:528).  x [frames, tokens, dim]. */
MAVLM_API int mavlm_frame_mean_fwd(const void* x, void* out, int frames, int tokens, int dim, int dtype, void* stream);
/* sim[t] = torch.cosine_similarity(x[t], x[t + 1], eps) for t < rows - 1 (segment.py:30, 69, 228;
 * compress_functions.py:30), rounded through the storage dtype.  Rows of row_elems elements, ld apart. */
MAVLM_API size_t mavlm_adjacent_cosine_workspace_bytes(int64_t rows, int64_t row_elems, int dtype);
MAVLM_API int mavlm_adjacent_cosine_fwd(const void* x, int64_t rows, int64_t row_elems, int64_t ld, float eps, float* sim,
                                        void* workspace, size_t workspace_bytes, int dtype, void* stream);
/* cal_depth_score (segment.py:3-25) / cal_left_depth_score (:210-223) on fp32 similarities; bit-exact. */
MAVLM_API int mavlm_depth_scores_fwd(const float* sim, float* depth, int n, int left_only, void* stream);
/* compress_spatial_features (memory_builder.py:72-99): avg_pool2d with window = stride on [frames, side*side, dim]
 * channels-last tokens -> [frames, o*o, dim], o = (side - window) / window + 1 (floor mode). */
MAVLM_API int mavlm_avg_pool_fwd(const void* x, void* out, int frames, int side, int window, int dim, int dtype,
                                 void* stream);
/* One iteration of (weighted_)kmeans_feature (compress_functions.py:94-173) over whole frames: labels (first
 * nearest centroid) and per-cluster weight sums against `cent`, new centroids for the non-empty clusters,
 * diff[k] = |cent_k - new_k| (0 for an empty cluster, which the caller re-seeds like the reference and measures
 * with mavlm_row_distance_fwd).  weights: fp32 [n_frames] or NULL (plain k-means); dist: optional fp32
 * [n_frames, clusters]; new_cent NULL: distances / labels / weight sums only (the key-frame search,
 * memory_builder.py:157-158). */
MAVLM_API size_t mavlm_kmeans_workspace_bytes(int64_t n_frames, int64_t row_elems, int clusters, int dtype);
MAVLM_API int mavlm_kmeans_iter_fwd(const void* x, const float* weights, const void* cent, void* new_cent,
                                    int32_t* labels, float* wsum, float* diff, float* dist, int64_t n_frames,
                                    int64_t row_elems, int clusters, void* workspace, size_t workspace_bytes, int dtype,
                                    void* stream);
MAVLM_API int mavlm_row_distance_fwd(const void* a, const void* b, float* dist, int rows, int64_t row_elems,
                                     void* workspace, size_t workspace_bytes, int dtype, void* stream);
/* Turing-memory update (memory_builder.py:52-64): w = ratio * softmax(scores * scale) over the n new tokens of each
 * row (columns [n, n_pad) are zero-filled so that w can feed the tensor-core GEMM), and, when mem is given,
 * mem_scaled = mem * (1 - sum_j w_ij); the caller accumulates w @ new onto it with mavlm_gemm_ex. */
MAVLM_API int mavlm_ntm_softmax_fwd(const float* scores, int64_t ld_scores, int64_t rows, int n, float scale,
                                    float ratio, void* w, int64_t ld_w, int n_pad, const void* mem, void* mem_scaled,
                                    int dim, int dtype, void* stream);

/* ======================= backward pass (training: BPTT through the memory, fuser) =======================
 * The reference trains this path with PyTorch autograd (train.py:1694-1728 unfreezes recurrent_memory_transformer,
 * memory_fuser, token_type_embedding; frame features are detached, llava_arch.py:302).  These entry points are
 * what the autograd.Functions of the host modules call. */

/* General GEMM: C[M,N] (+)= alpha * op(A) op(B).  trans_a = 0: A is [M,K]; 1: A is [K,M].  trans_b = 1: B is [N,K]
 * (nn.Linear weight layout); 0: B is [K,N].  Batched over outer*inner problems; strides (elements) =
 * {a_outer, a_inner, b_outer, b_inner, c_outer, c_inner} or NULL.  dgrad: dX = dY W (trans_b = 0);
 * wgrad: dW (+)= dY^T X (trans_a = 1, trans_b = 0, accumulate over chunks).  MAVLM_BF16 runs on the tcgen05
 * kernel (transposed operands are consumed as MN-major UMMA operands; alpha must be 1; out_dtype may be
 * MAVLM_F32), MAVLM_F16 likewise for the inference layouts (trans_a = 0), MAVLM_F32 on the SIMT tier. */
MAVLM_API int mavlm_gemm_ex(const void* A, int64_t lda, int trans_a, const void* B, int64_t ldb, int trans_b, void* C,
                            int64_t ldc, int M, int N, int K, float alpha, int accumulate, int outer, int inner,
                            const int64_t* host_strides, int dtype, int out_dtype, void* stream);
/* out[n] (+)= sum_m x[m,n]  (fp32 out): bias / type-embedding / newline gradients. */
MAVLM_API int mavlm_colsum(const void* x, int64_t ld, float* out, int M, int N, int accumulate, int dtype, void* stream);
/* LayerNorm backward from the saved fp32 pre-LN sum: dpre (fp32), dgamma/dbeta (fp32, ACCUMULATED into). */
MAVLM_API int mavlm_layernorm_bwd(const float* pre, const void* gamma, const void* dy, float* dpre, float* dgamma,
                                  float* dbeta, int rows, int dim, float eps, int dtype, void* stream);
/* Unfused activation (training keeps the GELU pre-activation) and its backward (ref = output for ReLU,
 * pre-activation for GELU). */
MAVLM_API int mavlm_act_fwd(const void* x, void* y, int64_t n, int act, int dtype, void* stream);
MAVLM_API int mavlm_act_bwd(const void* dy, const void* ref, void* dx, int64_t n, int act, int dtype, void* stream);
/* Attention backward (MemoryController.py:51-54): from q, k, v, o, dO and the forward's LSE to dQ, dK, dV in the
 * layouts of q, k, v. */
MAVLM_API size_t mavlm_xattn_bwd_workspace_bytes(int batch, int heads, int lq, int lk, int head_dim, int dtype);
MAVLM_API int mavlm_xattn_bwd(const void* Q, int64_t ldq, int64_t qb, const void* K, int64_t ldk, int64_t kb,
                              const void* V, int64_t ldv, int64_t vb, const void* O, int64_t ldo, int64_t ob,
                              const void* dO, int64_t lddo, int64_t dob, const float* lse, void* dQ, int64_t lddq,
                              int64_t dqb, void* dK, int64_t lddk, int64_t dkb, void* dV, int64_t lddv, int64_t dvb,
                              int batch, int heads, int lq, int lk, int head_dim, float scale, int dtype,
                              void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAVLM_H_ */
