// Exact-fp32 SIMT tier (sm_100a CUDA cores): tiled FMA GEMM with the same epilogues as the
// tensor-core tier, row softmax and column sums.  This tier exists because the fp32 parity bar
// (<= 1e-5 vs the fp64 oracle, BASELINE config 1) cannot be met with bf16/tf32 tensor-core
// operands; it is also what attention uses in fp32 (scores materialised in a workspace, which
// is how the frame-score output K9 is produced).
#include "common.cuh"

namespace mavlm {

struct SimtGemmParams {
  const float* A; long long lda; long long a_batch;      // A[m][k], k contiguous
  const float* B; long long ldb; long long b_batch;      // B_NK: B[n][k] (k contiguous); B_KN: B[k][n] (n contiguous)
  float* C; long long ldc; long long c_batch;
  const float* bias; const float* resid; long long ldr; const float* addvec;
  int M, N, K; int act; float alpha; int accumulate;   // accumulate: C += result (gradient accumulation)
  int inner_batch;            // blockIdx.z = outer * inner_batch + inner
  long long a_batch2, b_batch2, c_batch2;  // strides of the inner batch index (heads)
};

constexpr int SBM = 128, SBN = 64, SBK = 16, STHREADS = 256;

// A_KM: A stored [k][m] (m contiguous) instead of [m][k];  B_KN: B stored [k][n] instead of [n][k].
template <bool A_KM, bool B_KN>
__global__ void __launch_bounds__(STHREADS) simt_gemm_kernel(SimtGemmParams p) {
  __shared__ __align__(16) float As[SBK][SBM + 4];
  __shared__ __align__(16) float Bs[SBK][SBN + 4];
  const int outer = blockIdx.z / p.inner_batch, inner = blockIdx.z % p.inner_batch;
  const float* A = p.A + outer * p.a_batch + inner * p.a_batch2;
  const float* B = p.B + outer * p.b_batch + inner * p.b_batch2;
  float* C = p.C + outer * p.c_batch + inner * p.c_batch2;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads, each 8 (m) x 4 (n)
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += SBK) {
    if (!A_KM) {
      // A tile: 128 rows x 16 k -> 512 float4, two per thread, stored transposed
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int q = tid + it * STHREADS;
        const int r = q / 4, kk = (q % 4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int gm = m0 + r, gk = k0 + kk;
        if (gm < p.M) {
          if (gk + 3 < p.K && (reinterpret_cast<uintptr_t>(A + gm * p.lda + gk) & 15) == 0) {
            v = *reinterpret_cast<const float4*>(A + gm * p.lda + gk);
          } else {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
            for (int e = 0; e < 4; ++e)
              if (gk + e < p.K) t[e] = A[gm * p.lda + gk + e];
            v = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
        As[kk + 0][r] = v.x; As[kk + 1][r] = v.y; As[kk + 2][r] = v.z; As[kk + 3][r] = v.w;
      }
    } else {
      // A stored [k][m]: 16 k x 128 m -> 512 float4 along m, two per thread
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int q = tid + it * STHREADS;
        const int kk = q / 32, c = (q % 32) * 4;
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        const int gk = k0 + kk, gm = m0 + c;
        if (gk < p.K) {
          if (gm + 3 < p.M && (reinterpret_cast<uintptr_t>(A + gk * p.lda + gm) & 15) == 0) {
            const float4 v = *reinterpret_cast<const float4*>(A + gk * p.lda + gm);
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
          } else {
            for (int e = 0; e < 4; ++e)
              if (gm + e < p.M) t[e] = A[gk * p.lda + gm + e];
          }
        }
        *reinterpret_cast<float4*>(&As[kk][c]) = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
    if (!B_KN) {
      const int r = tid / 4, kk = (tid % 4) * 4;  // 64 rows x 16 k
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int gn = n0 + r, gk = k0 + kk;
      if (gn < p.N) {
        if (gk + 3 < p.K && (reinterpret_cast<uintptr_t>(B + gn * p.ldb + gk) & 15) == 0) {
          v = *reinterpret_cast<const float4*>(B + gn * p.ldb + gk);
        } else {
          float t[4] = {0.f, 0.f, 0.f, 0.f};
          for (int e = 0; e < 4; ++e)
            if (gk + e < p.K) t[e] = B[gn * p.ldb + gk + e];
          v = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      Bs[kk + 0][r] = v.x; Bs[kk + 1][r] = v.y; Bs[kk + 2][r] = v.z; Bs[kk + 3][r] = v.w;
    } else {
      const int kk = tid / 16, c = (tid % 16) * 4;  // 16 k x 64 n
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      const int gk = k0 + kk, gn = n0 + c;
      if (gk < p.K) {
        if (gn + 3 < p.N && (reinterpret_cast<uintptr_t>(B + gk * p.ldb + gn) & 15) == 0) {
          float4 v = *reinterpret_cast<const float4*>(B + gk * p.ldb + gn);
          t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
        } else {
          for (int e = 0; e < 4; ++e)
            if (gn + e < p.N) t[e] = B[gk * p.ldb + gn + e];
        }
      }
      *reinterpret_cast<float4*>(&Bs[kk][c]) = make_float4(t[0], t[1], t[2], t[3]);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      float a[8], b[4];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty * 8 + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (p.bias) v += p.bias[gn];
      if (p.act == MAVLM_ACT_GELU_ERF) v = gelu_erf_f(v);
      else if (p.act == MAVLM_ACT_RELU) v = fmaxf(v, 0.f);
      if (p.resid) v += p.resid[gm * p.ldr + gn];
      if (p.addvec) v += p.addvec[gn];
      if (p.accumulate) v += C[gm * p.ldc + gn];
      C[gm * p.ldc + gn] = v;
    }
  }
}

int simt_gemm_launch(const SimtGemmParams& p, bool b_kn, int batches, cudaStream_t st, bool a_km = false) {
  if (p.M == 0 || p.N == 0 || batches == 0) return MAVLM_OK;
  // unaligned rows (e.g. head_dim 2 in the tiny golden fixtures) take the scalar load path in-kernel
  dim3 grid(ceil_div(p.N, SBN), ceil_div(p.M, SBM), batches);
  MAVLM_REQUIRE(grid.y <= 65535 && grid.z <= 65535, MAVLM_E_INVALID, "fp32 gemm: grid too large");
  if (a_km) {
    if (b_kn) simt_gemm_kernel<true, true><<<grid, STHREADS, 0, st>>>(p);
    else simt_gemm_kernel<true, false><<<grid, STHREADS, 0, st>>>(p);
  } else {
    if (b_kn) simt_gemm_kernel<false, true><<<grid, STHREADS, 0, st>>>(p);
    else simt_gemm_kernel<false, false><<<grid, STHREADS, 0, st>>>(p);
  }
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

// in-place softmax over rows of length n (fp32)
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ s, int n) {
  __shared__ float red[32];
  float* r = s + static_cast<long long>(blockIdx.x) * n;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, r[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float e = expf(r[i] - m);
    r[i] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) sum += red[w];
  const float inv = 1.f / sum;
  for (int i = threadIdx.x; i < n; i += blockDim.x) r[i] *= inv;
}

// lse[row] from already-normalised rows is not recoverable; the fp32 tier computes it separately
__global__ void __launch_bounds__(256) lse_rows_kernel(const float* __restrict__ s, float* __restrict__ lse, int n) {
  __shared__ float red[32];
  const float* r = s + static_cast<long long>(blockIdx.x) * n;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, r[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sum += expf(r[i] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    sum = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) sum += red[w];
    lse[blockIdx.x] = m + logf(sum);
  }
}

// out[b][col] += sum over a slab of rows of probs[b][rows][col]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ probs, float* __restrict__ out,
                                                     long long rows_per_batch, int n, int rows_per_block) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n) return;
  const long long b = blockIdx.z;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(r0 + rows_per_block, rows_per_batch);
  const float* p = probs + (b * rows_per_batch) * n + col;
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += p[r * n];
  atomicAdd(out + b * n + col, s);
}

int xattn_fp32(const float* Q, long long ldq, long long qb, const float* K, long long ldk, long long kb, const float* V,
               long long ldv, long long vb, float* O, long long ldo, long long ob, float* lse, float* col_scores,
               int batch, int heads, int lq, int lk, int dh, float scale, void* ws, size_t ws_bytes, cudaStream_t st) {
  const size_t need = static_cast<size_t>(batch) * heads * lq * static_cast<size_t>(lk) * sizeof(float);
  MAVLM_REQUIRE(ws != nullptr && ws_bytes >= need, MAVLM_E_WORKSPACE,
                "xattn fp32: workspace of %zu bytes needed, %zu given", need, ws_bytes);
  if (batch == 0 || lq == 0) return MAVLM_OK;
  MAVLM_REQUIRE(lk > 0, MAVLM_E_INVALID, "xattn: empty key set");
  float* S = static_cast<float*>(ws);
  const long long hs = static_cast<long long>(lq) * lk;  // scores stride per head
  SimtGemmParams p{};
  p.A = Q; p.lda = ldq; p.a_batch = qb; p.a_batch2 = dh;
  p.B = K; p.ldb = ldk; p.b_batch = kb; p.b_batch2 = dh;
  p.C = S; p.ldc = lk; p.c_batch = hs * heads; p.c_batch2 = hs;
  p.M = lq; p.N = lk; p.K = dh; p.act = MAVLM_ACT_NONE; p.alpha = scale; p.inner_batch = heads;
  int rc = simt_gemm_launch(p, false, batch * heads, st);
  if (rc) return rc;
  const long long rows = static_cast<long long>(batch) * heads * lq;
  if (lse != nullptr) {
    lse_rows_kernel<<<static_cast<unsigned>(rows), 256, 0, st>>>(S, lse, lk);
    MAVLM_LAUNCH_OK();
  }
  softmax_rows_kernel<<<static_cast<unsigned>(rows), 256, 0, st>>>(S, lk);
  MAVLM_LAUNCH_OK();
  if (col_scores != nullptr) {
    MAVLM_CUDA_OK(cudaMemsetAsync(col_scores, 0, static_cast<size_t>(batch) * lk * sizeof(float), st));
    const int rpb = 128;
    dim3 grid(ceil_div(lk, 256), ceil_div(heads * lq, rpb), batch);
    colsum_kernel<<<grid, 256, 0, st>>>(S, col_scores, static_cast<long long>(heads) * lq, lk, rpb);
    MAVLM_LAUNCH_OK();
  }
  SimtGemmParams g{};
  g.A = S; g.lda = lk; g.a_batch = hs * heads; g.a_batch2 = hs;
  g.B = V; g.ldb = ldv; g.b_batch = vb; g.b_batch2 = dh;
  g.C = O; g.ldc = ldo; g.c_batch = ob; g.c_batch2 = dh;
  g.M = lq; g.N = dh; g.K = lk; g.act = MAVLM_ACT_NONE; g.alpha = 1.f; g.inner_batch = heads;
  return simt_gemm_launch(g, true, batch * heads, st);
}

// General fp32 GEMM of the backward pass: C[M,N] (+)= alpha * op(A) op(B), with (outer, inner) batching.
int gemm_ex_fp32(const float* A, long long lda, int trans_a, const float* B, long long ldb, int trans_b, float* C,
                 long long ldc, int M, int N, int K, float alpha, int accumulate, int outer, int inner,
                 const long long* strides /* a_o, a_i, b_o, b_i, c_o, c_i */, cudaStream_t st) {
  SimtGemmParams p{};
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.M = M; p.N = N; p.K = K; p.act = MAVLM_ACT_NONE; p.alpha = alpha; p.accumulate = accumulate;
  p.inner_batch = inner < 1 ? 1 : inner;
  if (strides != nullptr) {
    p.a_batch = strides[0]; p.a_batch2 = strides[1];
    p.b_batch = strides[2]; p.b_batch2 = strides[3];
    p.c_batch = strides[4]; p.c_batch2 = strides[5];
  }
  // trans_b = 1: B is [N,K] (nn.Linear weight layout, the kernel's native B_NK); 0: B is [K,N]
  return simt_gemm_launch(p, trans_b == 0, (outer < 1 ? 1 : outer) * p.inner_batch, st, trans_a != 0);
}

int gemm_fp32(const float* A, long long lda, const float* W, long long ldw, const float* bias, const float* resid,
              long long ldr, const float* addvec, float* C, long long ldc, int M, int N, int K, int act,
              cudaStream_t st) {
  SimtGemmParams p{};
  p.A = A; p.lda = lda; p.B = W; p.ldb = ldw; p.C = C; p.ldc = ldc;
  p.bias = bias; p.resid = resid; p.ldr = ldr; p.addvec = addvec;
  p.M = M; p.N = N; p.K = K; p.act = act; p.alpha = 1.f; p.inner_batch = 1;
  return simt_gemm_launch(p, false, 1, st);
}

}  // namespace mavlm
