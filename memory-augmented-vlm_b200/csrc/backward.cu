// Backward pass of the memory path (training config: BPTT through the recurrent memory, fuser,
// type embeddings; frame features are detached, llava_arch.py:302,481).  This file holds the
// non-GEMM pieces (column sums for bias grads, LayerNorm backward, activation backward, the
// elementwise steps of attention backward) and the C ABI; GEMMs go through gemm_ex_fp32 (SIMT tier).
#include "common.cuh"

namespace mavlm {

int gemm_ex_fp32(const float* A, long long lda, int trans_a, const float* B, long long ldb, int trans_b, float* C,
                 long long ldc, int M, int N, int K, float alpha, int accumulate, int outer, int inner,
                 const long long* strides, cudaStream_t st);
int attn_bwd_scores_gemm(int mode, const __nv_bfloat16* A, long long lda, long long a_batch, const __nv_bfloat16* B,
                         long long ldb, long long b_batch, __nv_bfloat16* out, const float* vec, int batch, int heads, int M,
                         int N, int dh, float scale, cudaStream_t st);
int gemm_ex_bf16(const __nv_bfloat16* A, long long lda, int trans_a, const __nv_bfloat16* B, long long ldb, int trans_b,
                 void* C, long long ldc, int M, int N, int K, int accumulate, int out_f32, int outer, int inner,
                 const long long* s6, cudaStream_t st, int half = 0);

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// out[n] (+)= sum_m x[m, n]   (bias / embedding gradients), fp32 accumulation, one atomic per block column
template <typename T>
__global__ void __launch_bounds__(256) colsum_t_kernel(const T* __restrict__ x, long long ld, float* __restrict__ out,
                                                       int M, int N, int rows_per_block) {
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(r0 + rows_per_block, M);
  float s = 0.f;
  if (col < N)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) s += ldf(x + static_cast<long long>(r) * ld + col);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(out + col, t);
  }
}

// bf16, N % 8 == 0, 16-byte aligned rows: a warp reads 512 contiguous bytes of a row (8 columns per lane); the eight
// warps of a block stride the rows.  (The scalar kernel above moves 64 bytes per warp load: 1.7 ms of a 50 ms training
// step went into bias-gradient sums.)
__global__ void __launch_bounds__(256) colsum_bf16_vec_kernel(const __nv_bfloat16* __restrict__ x, long long ld,
                                                              float* __restrict__ out, int M, int N, int rows_per_block) {
  __shared__ float red[8][256 + 8];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(r0 + rows_per_block, M);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < N) {
#pragma unroll 4
    for (int r = r0 + wy; r < r1; r += 8) {
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(x + static_cast<long long>(r) * ld + col));
      const uint32_t* h = reinterpret_cast<const uint32_t*>(&t);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h[e]));
        s[2 * e] += f.x;
        s[2 * e + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[wy][lane * 8 + e] = s[e];
  __syncthreads();
  const int c = threadIdx.x;  // one column of the block's 256 per thread
  if (blockIdx.x * 256 + c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][c];
    atomicAdd(out + blockIdx.x * 256 + c, t);
  }
}

__device__ __forceinline__ float block_sum_b(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// LayerNorm backward.  x = pre-LN sum (fp32), y = xhat*g + b.  One CTA walks LN_ROWS consecutive rows and
// keeps the dgamma / dbeta partial sums of its columns in registers (one atomic per column per CTA).
constexpr int LN_ROWS = 8;
template <typename T, int THREADS, int CACHE>
__global__ void __launch_bounds__(THREADS) layernorm_bwd_kernel(const float* __restrict__ x, const T* __restrict__ gamma,
                                                                const T* __restrict__ dy, float* __restrict__ dx,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                int rows, int dim, float eps) {
  __shared__ float red[32];
  float dg[CACHE][4], db[CACHE][4], g[CACHE][4];
#pragma unroll
  for (int c = 0; c < CACHE; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      dg[c][k] = 0.f;
      db[c][k] = 0.f;
      const int i = (c * THREADS + threadIdx.x) * 4 + k;
      g[c][k] = i < dim ? ldf(gamma + i) : 0.f;
    }
  const int r0 = blockIdx.x * LN_ROWS, r1 = min(r0 + LN_ROWS, rows);
  for (int r = r0; r < r1; ++r) {
    const float* xr = x + static_cast<long long>(r) * dim;
    const T* dyr = dy + static_cast<long long>(r) * dim;
    float v[CACHE][4], d[CACHE][4];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CACHE; ++c) {
      const int i = (c * THREADS + threadIdx.x) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[c][k] = (i + k < dim) ? xr[i + k] : 0.f;
        d[c][k] = (i + k < dim) ? ldf(dyr + i + k) : 0.f;
        s += v[c][k];
      }
    }
    const float mean = block_sum_b(s, red) / dim;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CACHE; ++c) {
      const int i = (c * THREADS + threadIdx.x) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i + k < dim) {
          const float t = v[c][k] - mean;
          q += t * t;
        }
    }
    const float rstd = rsqrtf(block_sum_b(q, red) / dim + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CACHE; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[c][k] = (v[c][k] - mean) * rstd;  // xhat (0 outside dim because d == 0 there is not enough: mask below)
        const float gd = g[c][k] * d[c][k];
        s1 += gd;
        s2 += gd * v[c][k];
      }
    s1 = block_sum_b(s1, red) / dim;
    s2 = block_sum_b(s2, red) / dim;
    float* dxr = dx + static_cast<long long>(r) * dim;
#pragma unroll
    for (int c = 0; c < CACHE; ++c) {
      const int i = (c * THREADS + threadIdx.x) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i + k < dim) {
          dxr[i + k] = rstd * (g[c][k] * d[c][k] - s1 - v[c][k] * s2);
          dg[c][k] += d[c][k] * v[c][k];
          db[c][k] += d[c][k];
        }
    }
  }
#pragma unroll
  for (int c = 0; c < CACHE; ++c) {
    const int i = (c * THREADS + threadIdx.x) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i + k < dim) {
        atomicAdd(dgamma + i + k, dg[c][k]);
        atomicAdd(dbeta + i + k, db[c][k]);
      }
  }
}

// LayerNorm backward, vectorised and persistent (dim % 4 == 0, 16-byte aligned rows): the kernel above walks 8 rows per
// CTA with scalar loads (2-byte loads of dy), four block reductions of two barriers each per row and one atomic per
// column per 8 rows -- 344 us for the 12 544 x 3584 rows of a batch-8 training step = 1.3 TB/s.  Here a CTA strides
// the rows (grid = 2 x SMs), reads x as float4 and dy as 8 / 16-byte vectors, PREFETCHES the next row's vectors into
// registers before the current row's reductions (the loads stay in flight across the barriers), needs three barriers
// per row (mean; variance; the two dot products together -- each reduction has its own shared-memory slot, so no
// barrier guards the slot's reuse) and keeps its dgamma / dbeta partial sums in registers to the end (one atomic per
// column per CTA: 296 instead of 1568 per column).
template <typename T>
struct DyVec;
template <>
struct DyVec<float> {
  using V = float4;
  __device__ static __forceinline__ void unpack(const float4& t, float (&d)[4]) { d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w; }
};
template <>
struct DyVec<__nv_bfloat16> {
  using V = uint2;
  __device__ static __forceinline__ void unpack(const uint2& t, float (&d)[4]) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y;
  }
};

template <int NW>
__device__ __forceinline__ float block_sum_1(float v, float* slot) {  // slot[NW], one barrier
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < NW; ++i) t += slot[i];
  return t;
}

template <typename T, int THREADS, int CACHE>
__global__ void __launch_bounds__(THREADS) layernorm_bwd_vec_kernel(const float* __restrict__ x, const T* __restrict__ gamma,
                                                                    const T* __restrict__ dy, float* __restrict__ dx,
                                                                    float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                    int rows, int dim, float eps) {
  constexpr int NW = THREADS / 32;
  using DV = typename DyVec<T>::V;
  __shared__ float red[4][NW];
  float dg[CACHE][4], db[CACHE][4], g[CACHE][4];
#pragma unroll
  for (int c = 0; c < CACHE; ++c) {
    const int i = (c * THREADS + threadIdx.x) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      dg[c][k] = 0.f;
      db[c][k] = 0.f;
      g[c][k] = i < dim ? ldf(gamma + i + k) : 0.f;
    }
  }
  float4 xn[CACHE];
  DV dn[CACHE];
  auto fetch = [&](long long r) {
#pragma unroll
    for (int c = 0; c < CACHE; ++c) {
      const int i = (c * THREADS + threadIdx.x) * 4;
      if (i < dim) {
        xn[c] = *reinterpret_cast<const float4*>(x + r * dim + i);
        dn[c] = *reinterpret_cast<const DV*>(dy + r * dim + i);
      }
    }
  };
  long long r = blockIdx.x;
  if (r < rows) fetch(r);
  const float inv_dim = 1.f / static_cast<float>(dim);
  for (; r < rows; r += gridDim.x) {
    float v[CACHE][4], d[CACHE][4];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CACHE; ++c) {
      const int i = (c * THREADS + threadIdx.x) * 4;
      if (i < dim) {
        v[c][0] = xn[c].x; v[c][1] = xn[c].y; v[c][2] = xn[c].z; v[c][3] = xn[c].w;
        DyVec<T>::unpack(dn[c], d[c]);
        s += (v[c][0] + v[c][1]) + (v[c][2] + v[c][3]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[c][k] = 0.f; d[c][k] = 0.f; }
      }
    }
    const long long rn = r + gridDim.x;
    if (rn < rows) fetch(rn);  // in flight during this row's three reductions
    const float mean = block_sum_1<NW>(s, red[0]) * inv_dim;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CACHE; ++c) {
      const int i = (c * THREADS + threadIdx.x) * 4;
      if (i < dim) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float t = v[c][k] - mean;
          q += t * t;
        }
      }
    }
    const float rstd = rsqrtf(block_sum_1<NW>(q, red[1]) * inv_dim + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CACHE; ++c) {
      const int i = (c * THREADS + threadIdx.x) * 4;
      if (i < dim) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[c][k] = (v[c][k] - mean) * rstd;  // xhat
          const float gd = g[c][k] * d[c][k];
          s1 += gd;
          s2 += gd * v[c][k];
        }
      }
    }
    {  // both dot products behind ONE barrier
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if ((threadIdx.x & 31) == 0) {
        red[2][threadIdx.x >> 5] = s1;
        red[3][threadIdx.x >> 5] = s2;
      }
      __syncthreads();
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        t1 += red[2][i];
        t2 += red[3][i];
      }
      s1 = t1 * inv_dim;
      s2 = t2 * inv_dim;
    }
    float* dxr = dx + r * dim;
#pragma unroll
    for (int c = 0; c < CACHE; ++c) {
      const int i = (c * THREADS + threadIdx.x) * 4;
      if (i < dim) {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          o[k] = rstd * (g[c][k] * d[c][k] - s1 - v[c][k] * s2);
          dg[c][k] += d[c][k] * v[c][k];
          db[c][k] += d[c][k];
        }
        *reinterpret_cast<float4*>(dxr + i) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CACHE; ++c) {
    const int i = (c * THREADS + threadIdx.x) * 4;
    if (i < dim) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        atomicAdd(dgamma + i + k, dg[c][k]);
        atomicAdd(dbeta + i + k, db[c][k]);
      }
    }
  }
}

// activations: forward (training keeps GELU unfused so that the pre-activation survives) and backward
template <typename T>
__global__ void __launch_bounds__(256) act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, int act) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = ldf(x + i);
    stf(y + i, act == MAVLM_ACT_GELU_ERF ? gelu_erf_f(v) : (act == MAVLM_ACT_RELU ? fmaxf(v, 0.f) : v));
  }
}
// ref = the activation OUTPUT for ReLU (mask y > 0), the PRE-activation for GELU
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ ref,
                                                      T* __restrict__ dx, long long n, int act) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float g = ldf(dy + i), r = ldf(ref + i);
    float o = g;
    if (act == MAVLM_ACT_RELU) {
      o = r > 0.f ? g : 0.f;
    } else if (act == MAVLM_ACT_GELU_ERF) {
      const float cdf = 0.5f * (1.f + erff(r * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * __expf(-0.5f * r * r);
      o = g * (cdf + r * pdf);
    }
    stf(dx + i, o);
  }
}

// bf16, n % 8 == 0, 16-byte aligned: 8 elements per thread and iteration (the scalar kernels above move 64 bytes per
// warp load: 0.59 of the copy bandwidth on the 12 544 x 14 336 MLP activations of a training step).  Same arithmetic
// per element as the scalar kernels.
__device__ __forceinline__ float act_fwd_one(float v, int act) {
  return act == MAVLM_ACT_GELU_ERF ? gelu_erf_f(v) : (act == MAVLM_ACT_RELU ? fmaxf(v, 0.f) : v);
}
__device__ __forceinline__ float act_bwd_one(float g, float r, int act) {
  if (act == MAVLM_ACT_RELU) return r > 0.f ? g : 0.f;
  if (act == MAVLM_ACT_GELU_ERF) {
    const float cdf = 0.5f * (1.f + erff(r * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * r * r);
    return g * (cdf + r * pdf);
  }
  return g;
}
__device__ __forceinline__ void unpack8_bf16(const uint4& t, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8_bf16(const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return t;
}
__global__ void __launch_bounds__(256) act_fwd_bf16_vec_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                               long long nvec, int act) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v[8];
    unpack8_bf16(x[i], v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = act_fwd_one(v[k], act);
    y[i] = pack8_bf16(v);
  }
}
__global__ void __launch_bounds__(256) act_bwd_bf16_vec_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ ref,
                                                               uint4* __restrict__ dx, long long nvec, int act) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float g[8], r[8];
    unpack8_bf16(dy[i], g);
    unpack8_bf16(ref[i], r);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = act_bwd_one(g[k], r[k], act);
    dx[i] = pack8_bf16(g);
  }
}

// ---- attention backward, fp32 tier: elementwise steps over the materialised [B*H*Lq, Lk] score workspace ----
// s[row, :] = exp(s[row, :] - lse[row])
__global__ void __launch_bounds__(256) probs_from_lse_kernel(float* __restrict__ s, const float* __restrict__ lse, int n) {
  float* r = s + static_cast<long long>(blockIdx.x) * n;
  const float l = lse[blockIdx.x];
  for (int i = threadIdx.x; i < n; i += blockDim.x) r[i] = __expf(r[i] - l);
}
// D[b,h,q] = sum_c dO[b,q,h*dh+c] * O[b,q,h*dh+c]
__global__ void __launch_bounds__(128) rowdot_kernel(const float* __restrict__ dO, long long ldd, long long dob,
                                                     const float* __restrict__ O, long long ldo, long long ob,
                                                     float* __restrict__ D, int heads, int lq, int dh) {
  __shared__ float red[32];
  const int q = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const float* a = dO + b * dob + static_cast<long long>(q) * ldd + h * dh;
  const float* c = O + b * ob + static_cast<long long>(q) * ldo + h * dh;
  float s = 0.f;
  for (int i = threadIdx.x; i < dh; i += blockDim.x) s += a[i] * c[i];
  s = block_sum_b(s, red);
  if (threadIdx.x == 0) D[(static_cast<long long>(b) * heads + h) * lq + q] = s;
}
// dp[row, :] = p[row, :] * (dp[row, :] - D[row]) * scale      (dS, in place over dP)
__global__ void __launch_bounds__(256) ds_kernel(const float* __restrict__ p, float* __restrict__ dp,
                                                 const float* __restrict__ D, int n, float scale) {
  const float* pr = p + static_cast<long long>(blockIdx.x) * n;
  float* dr = dp + static_cast<long long>(blockIdx.x) * n;
  const float d = D[blockIdx.x];
  for (int i = threadIdx.x; i < n; i += blockDim.x) dr[i] = pr[i] * (dr[i] - d) * scale;
}

// ---- attention backward, bf16 tier: scores / dP stay fp32 in the workspace, P and dS are stored as bf16 GEMM operands
// p[row, :] = bf16(exp(s[row, :] * scale - lse[row]))
__global__ void __launch_bounds__(256) probs_bf16_kernel(const float* __restrict__ s, const float* __restrict__ lse,
                                                         __nv_bfloat16* __restrict__ p, int n, float scale) {
  const float* r = s + static_cast<long long>(blockIdx.x) * n;
  __nv_bfloat16* o = p + static_cast<long long>(blockIdx.x) * n;
  const float l = lse[blockIdx.x];
  for (int i = threadIdx.x * 2; i < n; i += blockDim.x * 2) {
    const float2 v = *reinterpret_cast<const float2*>(r + i);
    *reinterpret_cast<__nv_bfloat162*>(o + i) = __floats2bfloat162_rn(__expf(v.x * scale - l), __expf(v.y * scale - l));
  }
}
// ds[row, :] = bf16(p[row, :] * (dp[row, :] - D[row]) * scale), written over p
__global__ void __launch_bounds__(256) ds_bf16_kernel(__nv_bfloat16* __restrict__ p, const float* __restrict__ dp,
                                                      const float* __restrict__ D, int n, float scale) {
  __nv_bfloat16* pr = p + static_cast<long long>(blockIdx.x) * n;
  const float* dr = dp + static_cast<long long>(blockIdx.x) * n;
  const float d = D[blockIdx.x];
  for (int i = threadIdx.x * 2; i < n; i += blockDim.x * 2) {
    const float2 pv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pr + i));
    const float2 g = *reinterpret_cast<const float2*>(dr + i);
    *reinterpret_cast<__nv_bfloat162*>(pr + i) = __floats2bfloat162_rn(pv.x * (g.x - d) * scale, pv.y * (g.y - d) * scale);
  }
}
__global__ void __launch_bounds__(128) rowdot_bf16_kernel(const __nv_bfloat16* __restrict__ dO, long long ldd,
                                                          long long dob, const __nv_bfloat16* __restrict__ O,
                                                          long long ldo, long long ob, float* __restrict__ D, int heads,
                                                          int lq, int dh) {
  __shared__ float red[32];
  const int q = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const __nv_bfloat16* a = dO + b * dob + static_cast<long long>(q) * ldd + h * dh;
  const __nv_bfloat16* c = O + b * ob + static_cast<long long>(q) * ldo + h * dh;
  float s = 0.f;
  for (int i = threadIdx.x; i < dh; i += blockDim.x) s += __bfloat162float(a[i]) * __bfloat162float(c[i]);
  s = block_sum_b(s, red);
  if (threadIdx.x == 0) D[(static_cast<long long>(b) * heads + h) * lq + q] = s;
}

static int ew_grid(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  return static_cast<int>(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace mavlm

using namespace mavlm;

extern "C" {

int mavlm_gemm_ex(const void* A, int64_t lda, int trans_a, const void* B, int64_t ldb, int trans_b, void* C, int64_t ldc,
                  int M, int N, int K, float alpha, int accumulate, int outer, int inner, const int64_t* strides,
                  int dtype, int out_dtype, void* stream) {
  MAVLM_REQUIRE(M >= 0 && N >= 0 && K > 0, MAVLM_E_INVALID, "gemm_ex: bad shape");
  long long st6[6] = {0, 0, 0, 0, 0, 0};
  if (strides != nullptr)
    for (int i = 0; i < 6; ++i) st6[i] = strides[i];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == MAVLM_F32) {
    MAVLM_REQUIRE(out_dtype == MAVLM_F32, MAVLM_E_INVALID, "gemm_ex: fp32 inputs require fp32 output");
    return gemm_ex_fp32(static_cast<const float*>(A), lda, trans_a, static_cast<const float*>(B), ldb, trans_b,
                        static_cast<float*>(C), ldc, M, N, K, alpha, accumulate, outer, inner, st6, st);
  }
  MAVLM_REQUIRE(dtype == MAVLM_BF16 || dtype == MAVLM_F16, MAVLM_E_INVALID, "gemm_ex: bad dtype %d", dtype);
  MAVLM_REQUIRE(alpha == 1.f, MAVLM_E_INVALID, "gemm_ex: the tensor-core tier has no alpha scaling (got %f)", alpha);
  MAVLM_REQUIRE(out_dtype == MAVLM_F32 || out_dtype == dtype, MAVLM_E_INVALID, "gemm_ex: output must be fp32 or the input dtype");
  MAVLM_REQUIRE(dtype == MAVLM_BF16 || trans_a == 0, MAVLM_E_INVALID,
                "gemm_ex: fp16 is the inference dtype (trans_a = 1 is a training layout)");
  return gemm_ex_bf16(static_cast<const __nv_bfloat16*>(A), lda, trans_a, static_cast<const __nv_bfloat16*>(B), ldb,
                      trans_b, C, ldc, M, N, K, accumulate, out_dtype == MAVLM_F32, outer, inner, st6, st,
                      dtype == MAVLM_F16 ? 1 : 0);
}

int mavlm_colsum(const void* x, int64_t ld, float* out, int M, int N, int accumulate, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype == MAVLM_F32 || dtype == MAVLM_BF16, MAVLM_E_INVALID, "colsum: bad dtype");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N == 0) return MAVLM_OK;
  if (!accumulate) MAVLM_CUDA_OK(cudaMemsetAsync(out, 0, static_cast<size_t>(N) * sizeof(float), st));
  if (M == 0) return MAVLM_OK;
  if (dtype == MAVLM_BF16 && N % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // enough row blocks to fill the SMs a few times over
    const int col_blocks = ceil_div(N, 256);
    int rpb_v = 512;
    while (rpb_v > 64 && static_cast<long long>(col_blocks) * ceil_div(M, rpb_v) < 4ll * sm_count()) rpb_v >>= 1;
    colsum_bf16_vec_kernel<<<dim3(col_blocks, ceil_div(M, rpb_v)), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ld, out,
                                                                                 M, N, rpb_v);
    MAVLM_LAUNCH_OK();
    return MAVLM_OK;
  }
  const int rpb = 256;
  dim3 grid(ceil_div(N, 32), ceil_div(M, rpb));
  dim3 block(32, 8);
  if (dtype == MAVLM_F32)
    colsum_t_kernel<float><<<grid, block, 0, st>>>(static_cast<const float*>(x), ld, out, M, N, rpb);
  else
    colsum_t_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ld, out, M, N, rpb);
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

int mavlm_layernorm_bwd(const float* pre, const void* gamma, const void* dy, float* dpre, float* dgamma, float* dbeta,
                        int rows, int dim, float eps, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype == MAVLM_F32 || dtype == MAVLM_BF16, MAVLM_E_INVALID, "layernorm_bwd: bad dtype");
  MAVLM_REQUIRE(dim > 0 && dim <= 4096, MAVLM_E_INVALID, "layernorm_bwd: dim %d must be <= 4096", dim);
  if (rows == 0) return MAVLM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool f32 = dtype == MAVLM_F32;
  const uintptr_t dy_mask = f32 ? 15 : 7;
  if (dim % 4 == 0 && dim > 512 && (reinterpret_cast<uintptr_t>(pre) & 15) == 0 && (reinterpret_cast<uintptr_t>(dpre) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dy) & dy_mask) == 0) {
    const int pgrid = rows < 2 * sm_count() ? rows : 2 * sm_count();
    if (f32)
      layernorm_bwd_vec_kernel<float, 256, 4><<<pgrid, 256, 0, st>>>(pre, static_cast<const float*>(gamma),
                                                                      static_cast<const float*>(dy), dpre, dgamma, dbeta,
                                                                      rows, dim, eps);
    else
      layernorm_bwd_vec_kernel<__nv_bfloat16, 256, 4><<<pgrid, 256, 0, st>>>(
          pre, static_cast<const __nv_bfloat16*>(gamma), static_cast<const __nv_bfloat16*>(dy), dpre, dgamma, dbeta, rows,
          dim, eps);
    MAVLM_LAUNCH_OK();
    return MAVLM_OK;
  }
  const int grid = ceil_div(rows, LN_ROWS);
#define MAVLM_LNB(T, TH, CA)                                                                                       \
  layernorm_bwd_kernel<T, TH, CA><<<grid, TH, 0, st>>>(pre, static_cast<const T*>(gamma), static_cast<const T*>(dy), \
                                                       dpre, dgamma, dbeta, rows, dim, eps)
  if (dim <= 512) {
    if (dtype == MAVLM_F32) MAVLM_LNB(float, 128, 1);
    else MAVLM_LNB(__nv_bfloat16, 128, 1);
  } else {
    if (dtype == MAVLM_F32) MAVLM_LNB(float, 256, 4);
    else MAVLM_LNB(__nv_bfloat16, 256, 4);
  }
#undef MAVLM_LNB
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

int mavlm_act_fwd(const void* x, void* y, int64_t n, int act, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype == MAVLM_F32 || dtype == MAVLM_BF16, MAVLM_E_INVALID, "act_fwd: bad dtype");
  if (n == 0) return MAVLM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == MAVLM_BF16 && n % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0)
    act_fwd_bf16_vec_kernel<<<ew_grid(n / 8), 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<uint4*>(y), n / 8, act);
  else if (dtype == MAVLM_F32)
    act_fwd_kernel<float><<<ew_grid(n), 256, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(y), n, act);
  else
    act_fwd_kernel<__nv_bfloat16><<<ew_grid(n), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                              static_cast<__nv_bfloat16*>(y), n, act);
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

int mavlm_act_bwd(const void* dy, const void* ref, void* dx, int64_t n, int act, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype == MAVLM_F32 || dtype == MAVLM_BF16, MAVLM_E_INVALID, "act_bwd: bad dtype");
  if (n == 0) return MAVLM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == MAVLM_BF16 && n % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(ref) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0)
    act_bwd_bf16_vec_kernel<<<ew_grid(n / 8), 256, 0, st>>>(static_cast<const uint4*>(dy), static_cast<const uint4*>(ref),
                                                            static_cast<uint4*>(dx), n / 8, act);
  else if (dtype == MAVLM_F32)
    act_bwd_kernel<float><<<ew_grid(n), 256, 0, st>>>(static_cast<const float*>(dy), static_cast<const float*>(ref),
                                                      static_cast<float*>(dx), n, act);
  else
    act_bwd_kernel<__nv_bfloat16><<<ew_grid(n), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dy),
                                                              static_cast<const __nv_bfloat16*>(ref),
                                                              static_cast<__nv_bfloat16*>(dx), n, act);
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

size_t mavlm_xattn_bwd_workspace_bytes(int batch, int heads, int lq, int lk, int head_dim, int dtype) {
  (void)head_dim;
  const size_t rows = static_cast<size_t>(batch) * heads * lq;
  if (dtype == MAVLM_BF16)  // P / dS bf16 (the fp32 scores and dP live only in TMEM accumulators), D fp32
    return rows * static_cast<size_t>(lk) * 2 + rows * sizeof(float) + 16;
  return (2 * rows * static_cast<size_t>(lk) + rows) * sizeof(float);  // P, dP/dS, D
}

int mavlm_xattn_bwd(const void* Q, int64_t ldq, int64_t qb, const void* K, int64_t ldk, int64_t kb, const void* V,
                    int64_t ldv, int64_t vb, const void* O, int64_t ldo, int64_t ob, const void* dO, int64_t lddo,
                    int64_t dob, const float* lse, void* dQ, int64_t lddq, int64_t dqb, void* dK, int64_t lddk,
                    int64_t dkb, void* dV, int64_t lddv, int64_t dvb, int batch, int heads, int lq, int lk, int head_dim,
                    float scale, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  MAVLM_REQUIRE(dtype == MAVLM_F32 || dtype == MAVLM_BF16, MAVLM_E_INVALID, "xattn_bwd: bad dtype %d", dtype);
  const size_t need = mavlm_xattn_bwd_workspace_bytes(batch, heads, lq, lk, head_dim, dtype);
  MAVLM_REQUIRE(workspace != nullptr && workspace_bytes >= need, MAVLM_E_WORKSPACE,
                "xattn_bwd: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
  if (batch == 0 || lq == 0 || lk == 0) return MAVLM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long rows = static_cast<long long>(batch) * heads * lq;
  const long long hs = static_cast<long long>(lq) * lk;
  if (dtype == MAVLM_BF16) {
    // Tensor-core backward: five batched tcgen05 GEMMs (transposed operands as MN-major UMMA operands); the two
    // elementwise steps ride in GEMM epilogues -- scores -> P (exp with the forward's LSE) and dP -> dS (in place over P)
    // are formed from the fp32 accumulators, only P and dS (bf16 GEMM operands) ever reach memory.
    MAVLM_REQUIRE(lk % 8 == 0 && head_dim % 8 == 0, MAVLM_E_INVALID, "xattn_bwd bf16: lk and head_dim must be multiples of 8");
    __nv_bfloat16* Pb = static_cast<__nv_bfloat16*>(workspace);                 // P, then dS   [B*H, Lq, Lk]
    float* Dv = reinterpret_cast<float*>(Pb + rows * lk);                       // D = rowsum(dO * O)   [B*H, Lq]
    const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(Q);
    const __nv_bfloat16* k = static_cast<const __nv_bfloat16*>(K);
    const __nv_bfloat16* v = static_cast<const __nv_bfloat16*>(V);
    const __nv_bfloat16* o = static_cast<const __nv_bfloat16*>(O);
    const __nv_bfloat16* go = static_cast<const __nv_bfloat16*>(dO);
    const int dh = head_dim;
    int rc;
    // P = exp(Q K^T * scale - lse)
    if ((rc = attn_bwd_scores_gemm(2, q, ldq, qb, k, ldk, kb, Pb, lse, batch, heads, lq, lk, dh, scale, st))) return rc;
    {  // dV = P^T dO
      const long long s6[6] = {hs * heads, hs, dob, dh, dvb, dh};
      if ((rc = gemm_ex_bf16(Pb, lk, 1, go, lddo, 0, dV, lddv, lk, dh, lq, 0, 0, batch, heads, s6, st))) return rc;
    }
    rowdot_bf16_kernel<<<dim3(lq, heads, batch), 128, 0, st>>>(go, lddo, dob, o, ldo, ob, Dv, heads, lq, dh);
    MAVLM_LAUNCH_OK();
    // dS = P * (dO V^T - D) * scale, over P
    if ((rc = attn_bwd_scores_gemm(3, go, lddo, dob, v, ldv, vb, Pb, Dv, batch, heads, lq, lk, dh, scale, st))) return rc;
    {  // dQ = dS K
      const long long s6[6] = {hs * heads, hs, kb, dh, dqb, dh};
      if ((rc = gemm_ex_bf16(Pb, lk, 0, k, ldk, 0, dQ, lddq, lq, dh, lk, 0, 0, batch, heads, s6, st))) return rc;
    }
    {  // dK = dS^T Q
      const long long s6[6] = {hs * heads, hs, qb, dh, dkb, dh};
      if ((rc = gemm_ex_bf16(Pb, lk, 1, q, ldq, 0, dK, lddk, lk, dh, lq, 0, 0, batch, heads, s6, st))) return rc;
    }
    return MAVLM_OK;
  }
  float* P = static_cast<float*>(workspace);
  float* dP = P + rows * lk;
  float* D = dP + rows * lk;
  const float* q = static_cast<const float*>(Q);
  const float* k = static_cast<const float*>(K);
  const float* v = static_cast<const float*>(V);
  const float* o = static_cast<const float*>(O);
  const float* go = static_cast<const float*>(dO);
  const int dh = head_dim, bh = batch * heads;
  int rc;
  // S = scale * Q K^T  ->  P = exp(S - lse)
  {
    const long long s6[6] = {qb, dh, kb, dh, hs * heads, hs};
    if ((rc = gemm_ex_fp32(q, ldq, 0, k, ldk, 1, P, lk, lq, lk, dh, scale, 0, batch, heads, s6, st))) return rc;
  }
  probs_from_lse_kernel<<<static_cast<unsigned>(rows), 256, 0, st>>>(P, lse, lk);
  MAVLM_LAUNCH_OK();
  // dV = P^T dO
  {
    const long long s6[6] = {hs * heads, hs, dob, dh, dvb, dh};
    if ((rc = gemm_ex_fp32(P, lk, 1, go, lddo, 0, static_cast<float*>(dV), lddv, lk, dh, lq, 1.f, 0, batch, heads, s6,
                           st)))
      return rc;
  }
  // dP = dO V^T
  {
    const long long s6[6] = {dob, dh, vb, dh, hs * heads, hs};
    if ((rc = gemm_ex_fp32(go, lddo, 0, v, ldv, 1, dP, lk, lq, lk, dh, 1.f, 0, batch, heads, s6, st))) return rc;
  }
  rowdot_kernel<<<dim3(lq, heads, batch), 128, 0, st>>>(go, lddo, dob, o, ldo, ob, D, heads, lq, dh);
  MAVLM_LAUNCH_OK();
  ds_kernel<<<static_cast<unsigned>(rows), 256, 0, st>>>(P, dP, D, lk, scale);
  MAVLM_LAUNCH_OK();
  // dQ = dS K ; dK = dS^T Q
  {
    const long long s6[6] = {hs * heads, hs, kb, dh, dqb, dh};
    if ((rc = gemm_ex_fp32(dP, lk, 0, k, ldk, 0, static_cast<float*>(dQ), lddq, lq, dh, lk, 1.f, 0, batch, heads, s6,
                           st)))
      return rc;
  }
  {
    const long long s6[6] = {hs * heads, hs, qb, dh, dkb, dh};
    if ((rc = gemm_ex_fp32(dP, lk, 1, q, ldq, 0, static_cast<float*>(dK), lddk, lk, dh, lq, 1.f, 0, batch, heads, s6,
                           st)))
      return rc;
  }
  (void)bh;
  return MAVLM_OK;
}

}  // extern "C"
