// extern "C" dispatch for the dense ops: picks the tcgen05 tier (bf16) or the exact SIMT tier (fp32).
#include "common.cuh"

namespace mavlm {
int gemm_fp32(const float* A, long long lda, const float* W, long long ldw, const float* bias, const float* resid,
              long long ldr, const float* addvec, float* C, long long ldc, int M, int N, int K, int act,
              cudaStream_t st);
int gemm_bf16_tc(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw,
                 const __nv_bfloat16* bias, const __nv_bfloat16* resid, long long ldr, const __nv_bfloat16* addvec,
                 void* C, long long ldc, int M, int N, int K, int act, int out_f32, cudaStream_t st, int half = 0,
                 const float* pe_table = nullptr, const long long* pe_idx = nullptr, int pe_tokens = 0);
void gemm_tc_force_bn(int bn);
int gemm_fill_num_tiles(const mavlm_gemm_desc* d);
int gemm_fill_range(const mavlm_gemm_desc* d, int t0, int t1, int half, cudaStream_t st);
int gemm_fill_fwd(const mavlm_gemm_desc* prim, const mavlm_gemm_desc* fill, int fill_begin, int fill_avail_end,
                  int* fill_done_end, int half, cudaStream_t st);
int xattn_fp32(const float* Q, long long ldq, long long qb, const float* K, long long ldk, long long kb, const float* V,
               long long ldv, long long vb, float* O, long long ldo, long long ob, float* lse, float* col_scores,
               int batch, int heads, int lq, int lk, int dh, float scale, void* ws, size_t ws_bytes, cudaStream_t st);
int xattn_bf16_tc(const __nv_bfloat16* Q, long long ldq, long long qb, const __nv_bfloat16* K, long long ldk,
                  long long kb, const __nv_bfloat16* V, long long ldv, long long vb, __nv_bfloat16* O, long long ldo,
                  long long ob, float* lse, int batch, int heads, int lq, int lk, int dh, float scale, void* ws,
                  size_t ws_bytes, cudaStream_t st, int half = 0);
size_t xattn_bf16_workspace_bytes(int batch, int heads, int lq, int lk, int dh);
void attn_force_groups(int n);
void attn_pair_force_groups(int n);
void attn_use_pair_kernel(bool on);
void attn_pair_set_trace(unsigned long long* buf);
void attn_tc_set_trace(unsigned long long* buf);
int xattn_colsum_tc(const __nv_bfloat16* Q, long long ldq, long long qb, const __nv_bfloat16* K, long long ldk,
                    long long kb, const float* lse, float* out, int batch, int heads, int lq, int lk, int dh, float scale,
                    int half, cudaStream_t st);
}  // namespace mavlm

using namespace mavlm;

extern "C" {

int mavlm_gemm_bias_act_fwd(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* resid,
                            int64_t ldr, const void* addvec, void* C, int64_t ldc, int M, int N, int K, int act,
                            int dtype, int out_dtype, void* stream) {
  MAVLM_REQUIRE(M >= 0 && N >= 0 && K > 0, MAVLM_E_INVALID, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  MAVLM_REQUIRE(act >= MAVLM_ACT_NONE && act <= MAVLM_ACT_RELU, MAVLM_E_INVALID, "gemm: bad activation %d", act);
  MAVLM_REQUIRE(A != nullptr && W != nullptr && C != nullptr, MAVLM_E_INVALID, "gemm: NULL operand");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == MAVLM_F32) {
    MAVLM_REQUIRE(out_dtype == MAVLM_F32, MAVLM_E_INVALID, "gemm: fp32 inputs require fp32 output");
    return gemm_fp32(static_cast<const float*>(A), lda, static_cast<const float*>(W), ldw,
                     static_cast<const float*>(bias), static_cast<const float*>(resid), ldr,
                     static_cast<const float*>(addvec), static_cast<float*>(C), ldc, M, N, K, act, st);
  }
  MAVLM_REQUIRE(dtype == MAVLM_BF16 || dtype == MAVLM_F16, MAVLM_E_INVALID, "gemm: bad dtype %d", dtype);
  MAVLM_REQUIRE(out_dtype == dtype || out_dtype == MAVLM_F32, MAVLM_E_INVALID, "gemm: bad out_dtype %d", out_dtype);
  return gemm_bf16_tc(static_cast<const __nv_bfloat16*>(A), lda, static_cast<const __nv_bfloat16*>(W), ldw,
                      static_cast<const __nv_bfloat16*>(bias), static_cast<const __nv_bfloat16*>(resid), ldr,
                      static_cast<const __nv_bfloat16*>(addvec), C, ldc, M, N, K, act, out_dtype == MAVLM_F32, st,
                      dtype == MAVLM_F16 ? 1 : 0);
}

int mavlm_gemm_bias_pe_fwd(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const float* pe_table,
                           const int64_t* frame_idx, int tokens_per_frame, void* C, int64_t ldc, int M, int N, int K,
                           int dtype, void* stream) {
  MAVLM_REQUIRE(M >= 0 && N >= 0 && K > 0, MAVLM_E_INVALID, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  MAVLM_REQUIRE(A != nullptr && W != nullptr && C != nullptr && pe_table != nullptr && frame_idx != nullptr,
                MAVLM_E_INVALID, "gemm + PE: NULL operand");
  MAVLM_REQUIRE(dtype == MAVLM_BF16 || dtype == MAVLM_F16, MAVLM_E_INVALID,
                "gemm + PE is a tensor-core tier fusion (bf16 / fp16); fp32 uses gemm_bias_act_fwd + add_pe_fwd");
  static_assert(sizeof(long long) == sizeof(int64_t), "index type");
  return gemm_bf16_tc(static_cast<const __nv_bfloat16*>(A), lda, static_cast<const __nv_bfloat16*>(W), ldw,
                      static_cast<const __nv_bfloat16*>(bias), nullptr, 0, nullptr, C, ldc, M, N, K, MAVLM_ACT_NONE, 0,
                      static_cast<cudaStream_t>(stream), dtype == MAVLM_F16 ? 1 : 0, pe_table,
                      reinterpret_cast<const long long*>(frame_idx), tokens_per_frame);
}

int mavlm_gemm_num_tiles(const mavlm_gemm_desc* g) {
  MAVLM_REQUIRE(g != nullptr && g->M > 0 && g->N > 0, MAVLM_E_INVALID, "gemm_num_tiles: bad problem");
  return gemm_fill_num_tiles(g);
}

int mavlm_gemm_tiles_fwd(const mavlm_gemm_desc* g, int tile_begin, int tile_end, int dtype, void* stream) {
  MAVLM_REQUIRE(g != nullptr, MAVLM_E_INVALID, "gemm_tiles: NULL problem");
  MAVLM_REQUIRE(dtype == MAVLM_BF16 || dtype == MAVLM_F16, MAVLM_E_INVALID, "gemm_tiles: tensor-core tier only (bf16 / fp16)");
  MAVLM_REQUIRE(g->out_dtype == dtype || g->out_dtype == MAVLM_F32, MAVLM_E_INVALID, "gemm_tiles: bad out_dtype %d", g->out_dtype);
  return gemm_fill_range(g, tile_begin, tile_end, dtype == MAVLM_F16 ? 1 : 0, static_cast<cudaStream_t>(stream));
}

int mavlm_gemm_fill_fwd(const mavlm_gemm_desc* primary, const mavlm_gemm_desc* filler, int fill_begin, int fill_avail_end,
                        int* fill_done_end, int dtype, void* stream) {
  MAVLM_REQUIRE(primary != nullptr, MAVLM_E_INVALID, "gemm_fill: NULL primary");
  MAVLM_REQUIRE(dtype == MAVLM_BF16 || dtype == MAVLM_F16, MAVLM_E_INVALID, "gemm_fill: tensor-core tier only (bf16 / fp16)");
  MAVLM_REQUIRE(primary->out_dtype == dtype || primary->out_dtype == MAVLM_F32, MAVLM_E_INVALID, "gemm_fill: bad out_dtype");
  if (filler != nullptr)
    MAVLM_REQUIRE(filler->out_dtype == dtype || filler->out_dtype == MAVLM_F32, MAVLM_E_INVALID, "gemm_fill: bad filler out_dtype");
  return gemm_fill_fwd(primary, filler, fill_begin, fill_avail_end, fill_done_end, dtype == MAVLM_F16 ? 1 : 0,
                       static_cast<cudaStream_t>(stream));
}

size_t mavlm_xattn_workspace_bytes(int batch, int heads, int lq, int lk, int head_dim, int dtype) {
  if (dtype == MAVLM_F32) return static_cast<size_t>(batch) * heads * lq * static_cast<size_t>(lk) * sizeof(float);
  if (batch <= 0 || lq <= 0 || lk <= 0) return 0;
  return xattn_bf16_workspace_bytes(batch, heads, lq, lk, head_dim);
}

int mavlm_xattn_fwd(const void* Q, int64_t ldq, int64_t q_batch_stride, const void* K, int64_t ldk,
                    int64_t k_batch_stride, const void* V, int64_t ldv, int64_t v_batch_stride, void* O, int64_t ldo,
                    int64_t o_batch_stride, float* lse, float* col_scores, int batch, int heads, int lq, int lk,
                    int head_dim, float scale, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  MAVLM_REQUIRE(batch >= 0 && heads > 0 && lq >= 0 && lk >= 0 && head_dim > 0, MAVLM_E_INVALID, "xattn: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == MAVLM_F32)
    return xattn_fp32(static_cast<const float*>(Q), ldq, q_batch_stride, static_cast<const float*>(K), ldk,
                      k_batch_stride, static_cast<const float*>(V), ldv, v_batch_stride, static_cast<float*>(O), ldo,
                      o_batch_stride, lse, col_scores, batch, heads, lq, lk, head_dim, scale, workspace,
                      workspace_bytes, st);
  MAVLM_REQUIRE(dtype == MAVLM_BF16 || dtype == MAVLM_F16, MAVLM_E_INVALID, "xattn: bad dtype %d", dtype);
  MAVLM_REQUIRE(col_scores == nullptr, MAVLM_E_INVALID,
                "xattn: col_scores (frame scores, MemoryController.py:135) is produced by the fp32 tier only");
  return xattn_bf16_tc(static_cast<const __nv_bfloat16*>(Q), ldq, q_batch_stride,
                       static_cast<const __nv_bfloat16*>(K), ldk, k_batch_stride,
                       static_cast<const __nv_bfloat16*>(V), ldv, v_batch_stride, static_cast<__nv_bfloat16*>(O), ldo,
                       o_batch_stride, lse, batch, heads, lq, lk, head_dim, scale, workspace, workspace_bytes, st,
                       dtype == MAVLM_F16 ? 1 : 0);
}

int mavlm_xattn_colsum(const void* Q, int64_t ldq, int64_t q_batch_stride, const void* K, int64_t ldk,
                       int64_t k_batch_stride, const float* lse, float* col_scores, int batch, int heads, int lq, int lk,
                       int head_dim, float scale, int dtype, void* stream) {
  MAVLM_REQUIRE(batch >= 0 && heads > 0 && lq >= 0 && lk >= 0 && head_dim > 0, MAVLM_E_INVALID, "xattn_colsum: bad shape");
  MAVLM_REQUIRE(dtype == MAVLM_BF16 || dtype == MAVLM_F16, MAVLM_E_INVALID,
                "xattn_colsum is the tensor-core tier's second pass; the fp32 tier returns col_scores from mavlm_xattn_fwd");
  if (batch == 0 || lk == 0) return MAVLM_OK;
  if (lq == 0) {
    MAVLM_CUDA_OK(cudaMemsetAsync(col_scores, 0, static_cast<size_t>(batch) * lk * sizeof(float), static_cast<cudaStream_t>(stream)));
    return MAVLM_OK;
  }
  return xattn_colsum_tc(static_cast<const __nv_bfloat16*>(Q), ldq, q_batch_stride, static_cast<const __nv_bfloat16*>(K),
                         ldk, k_batch_stride, lse, col_scores, batch, heads, lq, lk, head_dim, scale,
                         dtype == MAVLM_F16 ? 1 : 0, static_cast<cudaStream_t>(stream));
}

/* development knob (not part of the reference-facing surface): force the GEMM tile (0 = heuristic).
   -1 = heuristic over single-CTA tiles only; 64/128/192/256 = single-CTA 128 x BN tiles; 1128/1192/1256 = CTA-pair 256 x BN tiles (K-major operands) */
MAVLM_API int mavlm_debug_force_gemm_bn(int bn) {
  MAVLM_REQUIRE(bn == -1 || bn == 0 || bn == 64 || bn == 128 || bn == 192 || bn == 256 || bn == 1128 || bn == 1192 || bn == 1256,
                MAVLM_E_INVALID, "bad BN %d", bn);
  gemm_tc_force_bn(bn);
  return MAVLM_OK;
}

/* development knob: bit 4 (16) turns programmatic dependent launch off (A/B timing of the launch overlap); bit 6 (64)
   runs head_dim 448 attention on the CTA-pair kernel (attn_pair.cu: correct, measured slower, kept for A/B).  No flag
   changes what a kernel computes or stores. */
MAVLM_API int mavlm_debug_set_flags(int flags) {
  pdl_force_off((flags & 16) != 0);
  attn_use_pair_kernel((flags & 64) != 0);
  return MAVLM_OK;
}

/* development knob: event trace of the first CTA pair of the head_dim 448 attention kernel: a DEVICE buffer of
   2 CTAs x 3 roles x 256 uint64 ((clock64 << 8) | event code), zero-filled by the caller; NULL switches it off */
MAVLM_API int mavlm_debug_attn_trace(void* device_buffer) {
  attn_pair_set_trace(static_cast<unsigned long long*>(device_buffer));
  return MAVLM_OK;
}

/* development knob: event trace of the single-CTA attention kernel (attn_tc.cu): a DEVICE buffer of
   grid x 2 roles x 64 uint64 ((clock64 << 8) | event code), zero-filled by the caller; NULL switches it off */
MAVLM_API int mavlm_debug_attn_tc_trace(void* device_buffer) {
  attn_tc_set_trace(static_cast<unsigned long long*>(device_buffer));
  return MAVLM_OK;
}

/* development knob: force the number of attention CTA groups (0 = as many as fit on the SMs) */
MAVLM_API int mavlm_debug_force_attn_groups(int n) {
  attn_force_groups(n);
  attn_pair_force_groups(n);
  return MAVLM_OK;
}

}  // extern "C"
