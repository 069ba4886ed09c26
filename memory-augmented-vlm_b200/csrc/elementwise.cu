// Bandwidth-bound kernels of the visual-memory path (sm_100a): fused spatial pool + temporal PE,
// standalone PE add, LayerNorm, token assembly.  All are 16-byte vectorised along the hidden
// dimension, coalesced, fp32 internal math with a single rounding to the storage dtype.
#include "vec.cuh"

namespace mavlm {


// ---------------------------------------------------------------------------------------
// K2 + K3: channels-last spatial pool (+ PE).  One thread per (frame, out token, 16-byte
// channel group); the 4 bilinear taps are 4 independent 16-byte loads.
// Tap arithmetic follows ATen upsample_bilinear2d (align_corners=False): scale = in/out (fp32),
// src = max(scale*(o+0.5)-0.5, 0), i0 = (int)src, i1 = i0 + (i0 < in-1), l1 = src - i0.
// ---------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(256) pool_pe_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                      const float* __restrict__ pe, const int64_t* __restrict__ fidx,
                                                      int frames, int side, int out_side, int stride, int dim) {
  constexpr int V = Vec<T>::N;
  pdl_trigger();
  pdl_wait();
  const int groups = dim / V;
  const long long total = static_cast<long long>(frames) * out_side * out_side * groups;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    long long t = i / groups;
    const int ox = static_cast<int>(t % out_side);
    t /= out_side;
    const int oy = static_cast<int>(t % out_side);
    const int f = static_cast<int>(t / out_side);
    const T* xf = x + (static_cast<long long>(f) * side * side) * dim + g * V;
    float acc[V];
    if (MODE == MAVLM_POOL_BILINEAR) {
      const float scale = static_cast<float>(side) / static_cast<float>(out_side);
      float sy = fmaxf(scale * (oy + 0.5f) - 0.5f, 0.f);
      float sx = fmaxf(scale * (ox + 0.5f) - 0.5f, 0.f);
      int y0 = min(static_cast<int>(sy), side - 1), x0 = min(static_cast<int>(sx), side - 1);
      int y1 = y0 + (y0 < side - 1), x1 = x0 + (x0 < side - 1);
      float ly1 = sy - y0, lx1 = sx - x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
      float v00[V], v01[V], v10[V], v11[V];
      Vec<T>::load_plain(xf + (static_cast<long long>(y0) * side + x0) * dim, v00);
      Vec<T>::load_plain(xf + (static_cast<long long>(y0) * side + x1) * dim, v01);
      Vec<T>::load_plain(xf + (static_cast<long long>(y1) * side + x0) * dim, v10);
      Vec<T>::load_plain(xf + (static_cast<long long>(y1) * side + x1) * dim, v11);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] = ly0 * (lx0 * v00[k] + lx1 * v01[k]) + ly1 * (lx0 * v10[k] + lx1 * v11[k]);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] = (MODE == MAVLM_POOL_MAX) ? -INFINITY : 0.f;
      for (int dy = 0; dy < stride; ++dy)
        for (int dx = 0; dx < stride; ++dx) {
          float v[V];
          Vec<T>::load_plain(xf + (static_cast<long long>(oy * stride + dy) * side + (ox * stride + dx)) * dim, v);
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k] = (MODE == MAVLM_POOL_MAX) ? fmaxf(acc[k], v[k]) : acc[k] + v[k];
        }
      if (MODE == MAVLM_POOL_AVERAGE) {
        const float inv = 1.f / static_cast<float>(stride * stride);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] *= inv;
      }
    }
    if (pe != nullptr) {
      const float* p = pe + fidx[f] * dim + g * V;
#pragma unroll
      for (int k = 0; k < V; k += 4) {
        float4 q = __ldg(reinterpret_cast<const float4*>(p + k));
        acc[k] += q.x; acc[k + 1] += q.y; acc[k + 2] += q.z; acc[k + 3] += q.w;
      }
    }
    Vec<T>::store(y + ((static_cast<long long>(f) * out_side + oy) * out_side + ox) * dim + g * V, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) add_pe_kernel(const T* x, T* y,  // y may alias x (in place)
                                                     const float* __restrict__ pe, const int64_t* __restrict__ fidx,
                                                     int frames, int tokens, int dim) {
  constexpr int V = Vec<T>::N;
  pdl_trigger();
  pdl_wait();
  const int groups = dim / V;
  const long long total = static_cast<long long>(frames) * tokens * groups;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    const long long row = i / groups;
    const int f = static_cast<int>(row / tokens);
    float v[V];
    Vec<T>::load_plain(x + row * dim + g * V, v);
    const float* p = pe + fidx[f] * dim + g * V;
#pragma unroll
    for (int k = 0; k < V; k += 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(p + k));
      v[k] += q.x; v[k + 1] += q.y; v[k + 2] += q.z; v[k + 3] += q.w;
    }
    Vec<T>::store(y + row * dim + g * V, v);
  }
}

// ---------------------------------------------------------------------------------------
// LayerNorm: persistent CTAs (one per SM, 12 warps), one WARP per row, rows staged through shared memory by the
// bulk-copy engine.  Every warp owns a one-row buffer and an mbarrier: lane 0 requests the row with ONE
// cp.async.bulk (14 KB for D = 3584 fp32), the warp pulls it into registers as 16-byte vectors and immediately
// requests its NEXT row into the same buffer, so a row's statistics / normalisation / stores overlap the next row's
// fetch and ~170 KB per SM stay in flight for the whole kernel (the register-only version had its loads in flight only
// during the load phase of each 4-row CTA and re-staged gamma / beta per CTA: 0.35-0.63 of the copy bandwidth even at
// 50 k rows).  Exact two-pass statistics in fp32 (eps = 1e-12 in the RMT makes the one-pass E[x^2]-E[x]^2 form unsafe).
// gamma / beta are module parameters: staged once per CTA BEFORE griddepcontrol.wait (under the previous kernel's tail).
// Rows are dealt round-robin over (warp, CTA) so that a single-wave launch (1568 rows of a memory state) spreads evenly.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int LN_WARPS = 12;

// global -> shared bulk copy completing on an mbarrier; `after` is an unused operand that orders the request behind
// the value's computation (the warp reduction that consumed every lane's reads of the buffer being overwritten)
__device__ __forceinline__ void ln_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, float after) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];  // %4"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "f"(after)
               : "memory");
}

template <typename TI, typename TO, int CACHE>  // CACHE = 4-float vectors per lane: dim <= 128 * CACHE
__global__ void __launch_bounds__(32 * LN_WARPS, 1) layernorm_kernel(const TI* __restrict__ x,
                                                                  const TO* __restrict__ gamma,
                                                                  const TO* __restrict__ beta, TO* __restrict__ y,
                                                                  int rows, int dim, float eps) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  TO* sg = reinterpret_cast<TO*>(ln_smem);
  TO* sb = sg + dim;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t row_bytes = static_cast<uint32_t>(dim) * sizeof(TI);
  uint8_t* bufs = ln_smem + 2 * static_cast<size_t>(dim) * sizeof(TO);
  const TI* buf = reinterpret_cast<const TI*>(bufs + static_cast<size_t>(warp) * row_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(bufs + static_cast<size_t>(LN_WARPS) * row_bytes) + warp;
  pdl_trigger();
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  {
    constexpr int EV = 16 / sizeof(TO);
    const int nv = dim / EV;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      reinterpret_cast<uint4*>(sg)[i] = __ldg(reinterpret_cast<const uint4*>(gamma) + i);
      reinterpret_cast<uint4*>(sb)[i] = __ldg(reinterpret_cast<const uint4*>(beta) + i);
    }
  }
  __syncthreads();
  pdl_wait();  // x is the previous kernel's output
  const long long stride = static_cast<long long>(gridDim.x) * LN_WARPS;
  long long row = static_cast<long long>(warp) * gridDim.x + blockIdx.x;
  if (row < rows && lane == 0) {
    mbar_expect_tx(bar, row_bytes);
    ln_bulk_load(const_cast<TI*>(buf), x + row * dim, row_bytes, bar, 0.f);
  }
  uint32_t phase = 0;
  for (; row < rows; row += stride) {
    mbar_wait(bar, phase);
    phase ^= 1;
    float v[CACHE][4];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CACHE; ++c) {
      const int i = (c * 32 + lane) * 4;
      if (i < dim) {
        if constexpr (sizeof(TI) == 4) {
          const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(buf) + i);
          v[c][0] = t.x; v[c][1] = t.y; v[c][2] = t.z; v[c][3] = t.w;
        } else {
          const uint2 t = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(buf) + i);
          const float2 a = Elem16<TO>::unpack2(t.x);   // a 16-bit input has the output's type
          const float2 b = Elem16<TO>::unpack2(t.y);
          v[c][0] = a.x; v[c][1] = a.y; v[c][2] = b.x; v[c][3] = b.y;
        }
        s += (v[c][0] + v[c][1]) + (v[c][2] + v[c][3]);
      } else {
        v[c][0] = v[c][1] = v[c][2] = v[c][3] = 0.f;
      }
    }
    const float mean = warp_sum(s) / static_cast<float>(dim);
    // every lane's reads of the buffer fed `mean`: the buffer is free, the next row goes into it now
    const long long next = row + stride;
    if (next < rows && lane == 0) {
      mbar_expect_tx(bar, row_bytes);
      ln_bulk_load(const_cast<TI*>(buf), x + next * dim, row_bytes, bar, mean);
    }
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CACHE; ++c) {
      const int i = (c * 32 + lane) * 4;
      if (i < dim) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float d = v[c][k] - mean;
          q += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(dim) + eps);
    TO* yr = y + row * dim;
#pragma unroll
    for (int c = 0; c < CACHE; ++c) {
      const int i = (c * 32 + lane) * 4;
      if (i < dim) {
        float g[4], b[4], o[4];
        if constexpr (sizeof(TO) == 4) {
          const float4 tg = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sg) + i);
          const float4 tb = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sb) + i);
          g[0] = tg.x; g[1] = tg.y; g[2] = tg.z; g[3] = tg.w;
          b[0] = tb.x; b[1] = tb.y; b[2] = tb.z; b[3] = tb.w;
        } else {
          const uint2 tg = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(sg) + i);
          const uint2 tb = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(sb) + i);
          const float2 g0 = Elem16<TO>::unpack2(tg.x);
          const float2 g1 = Elem16<TO>::unpack2(tg.y);
          const float2 b0 = Elem16<TO>::unpack2(tb.x);
          const float2 b1 = Elem16<TO>::unpack2(tb.y);
          g[0] = g0.x; g[1] = g0.y; g[2] = g1.x; g[3] = g1.y;
          b[0] = b0.x; b[1] = b0.y; b[2] = b1.x; b[3] = b1.y;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = (v[c][k] - mean) * rstd * g[k] + b[k];
        if constexpr (sizeof(TO) == 4) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(yr) + i) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
          uint2 t;
          t.x = Elem16<TO>::pack2(o[0], o[1]);
          t.y = Elem16<TO>::pack2(o[2], o[3]);
          *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(yr) + i) = t;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// K12 + K13: token assembly.  One CTA per output row that this launch writes; the row kind is decoded from its index.
// When the memory-token rows are already in place (mem == nullptr: the fuser GEMM's epilogue stored them) the grid
// does not cover them.  A thread issues all of its (up to 4) 16-byte row loads before the first store.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) assemble_kernel(T* __restrict__ seq, const T* __restrict__ mem, long long n_mem,
                                                       const T* __restrict__ frames,
                                                       const int64_t* __restrict__ fine_idx, int n_fine, int tokens,
                                                       const T* __restrict__ type_emb, const T* __restrict__ newline,
                                                       const T* __restrict__ table,
                                                       const int64_t* __restrict__ pm_ids, int n_pm,
                                                       const int64_t* __restrict__ pf_ids, int n_pf, int dim) {
  constexpr int V = Vec<T>::N;
  constexpr int U = 4;
  pdl_trigger();
  pdl_wait();
  long long out_row = blockIdx.x;
  if (mem == nullptr && out_row >= n_pm) out_row += n_mem;  // rows [n_pm, n_pm + n_mem) are not part of this grid
  long long r = out_row;
  const T* src = nullptr;
  const T* add = nullptr;
  const long long n_fine_rows = static_cast<long long>(n_fine) * tokens;
  if (r < n_pm) {
    src = table + pm_ids[r] * dim;
  } else if ((r -= n_pm) < n_mem) {
    src = mem + r * dim;
    add = type_emb;
  } else if ((r -= n_mem) < 1) {
    src = newline;
  } else if ((r -= 1) < n_pf) {
    src = table + pf_ids[r] * dim;
  } else if ((r -= n_pf) < n_fine_rows) {
    const long long f = r / tokens, t = r % tokens;
    src = frames + (fine_idx[f] * tokens + t) * dim;
    add = type_emb + dim;
  } else {
    src = newline;
  }
  T* dst = seq + out_row * dim;
  const int nvec = dim / V;
  for (int base = threadIdx.x; base < nvec; base += blockDim.x * U) {
    float v[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = base + u * blockDim.x;
      if (j < nvec) Vec<T>::load_plain(src + j * V, v[u]);
    }
    if (add != nullptr) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = base + u * blockDim.x;
        if (j < nvec) {
          float a[V];
          Vec<T>::load(add + j * V, a);
#pragma unroll
          for (int k = 0; k < V; ++k) v[u][k] += a[k];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = base + u * blockDim.x;
      if (j < nvec) Vec<T>::store(dst + j * V, v[u]);
    }
  }
}

// dtype conversion (fp32 <-> bf16 / fp16), 8 elements per thread
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      float v[8];
      if constexpr (sizeof(TI) == 4) {
        const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + i);
        const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + i + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
        Vec<TI>::load(x + i, v);
      }
      if constexpr (sizeof(TO) == 4) {
        float* o = reinterpret_cast<float*>(y) + i;
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      } else {
        Vec<TO>::store(y + i, v);
      }
    } else {
      for (long long k = i; k < n; ++k) {
        float f;
        if constexpr (sizeof(TI) == 4) f = reinterpret_cast<const float*>(x)[k];
        else f = Elem16<TI>::to_float(x[k]);
        if constexpr (sizeof(TO) == 4) reinterpret_cast<float*>(y)[k] = f;
        else y[k] = Elem16<TO>::from_float(f);
      }
    }
  }
}

// Text / vision splice (llava_arch.py:745-878): every output row copies one source row.
// src >= 0: embed_table[src];  src == -1: zero (padding);  src <= -2: feats[-(src + 2)].
template <typename T>
__global__ void __launch_bounds__(128) gather_rows_kernel(T* __restrict__ out, long long ld_out,
                                                          const T* __restrict__ table, const T* __restrict__ feats,
                                                          const int64_t* __restrict__ src, int dim,
                                                          long long n_table_rows, long long n_feat_rows) {
  constexpr int V = Vec<T>::N;
  const long long r = blockIdx.x;
  const int64_t sidx = src[r];
  // a source row outside its table is never dereferenced: the row is written as zeros (the host layer validates
  // the ids before the launch and raises like nn.Embedding would)
  const T* sp = nullptr;
  if (sidx >= 0) {
    if (sidx < n_table_rows) sp = table + sidx * dim;
  } else if (sidx <= -2) {
    const long long f = -(sidx + 2);
    if (f < n_feat_rows) sp = feats + f * dim;
  }
  T* dst = out + r * ld_out;
  for (int i = threadIdx.x * V; i < dim; i += blockDim.x * V) {
    float v[V];
    if (sp != nullptr) {
      Vec<T>::load(sp + i, v);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = 0.f;
    }
    Vec<T>::store(dst + i, v);
  }
}


template <typename T>
static int launch_pool(const void* x, void* y, const float* pe, const int64_t* fidx, int frames, int side,
                       int out_side, int stride, int dim, int mode, cudaStream_t st) {
  const long long total = static_cast<long long>(frames) * out_side * out_side * (dim / Vec<T>::N);
  const int grid = grid_for(total, 256);
  const T* xp = static_cast<const T*>(x);
  T* yp = static_cast<T*>(y);
  if (mode == MAVLM_POOL_BILINEAR) {
    LaunchCfg lc;
    make_launch(lc, dim3(grid), dim3(256), 0, st, 1, 8);
    MAVLM_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, pool_pe_kernel<T, MAVLM_POOL_BILINEAR>, xp, yp, pe, fidx, frames, side,
                                     out_side, stride, dim));
  } else if (mode == MAVLM_POOL_AVERAGE)
    pool_pe_kernel<T, MAVLM_POOL_AVERAGE><<<grid, 256, 0, st>>>(xp, yp, pe, fidx, frames, side, out_side, stride, dim);
  else
    pool_pe_kernel<T, MAVLM_POOL_MAX><<<grid, 256, 0, st>>>(xp, yp, pe, fidx, frames, side, out_side, stride, dim);
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

int layernorm_launch(const void* x, const void* gamma, const void* beta, void* y, int rows, int dim, float eps,
                     int x_dtype, int dtype, cudaStream_t st) {
  MAVLM_REQUIRE(dim % 4 == 0 && dim > 0 && dim <= 128 * 32, MAVLM_E_INVALID,
                "layernorm: dim %d must be a multiple of 4 and <= 4096", dim);
  MAVLM_REQUIRE(x_dtype == MAVLM_F32 || x_dtype == dtype, MAVLM_E_INVALID, "layernorm: x_dtype must be f32 or dtype");
  MAVLM_REQUIRE(dtype == MAVLM_F32 || dim % 8 == 0, MAVLM_E_INVALID, "layernorm: 16-bit dim %d must be a multiple of 8", dim);
  MAVLM_REQUIRE((reinterpret_cast<uintptr_t>(gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(beta) & 15) == 0,
                MAVLM_E_INVALID, "layernorm: gamma / beta must be 16-byte aligned");
  MAVLM_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, MAVLM_E_INVALID, "layernorm: x must be 16-byte aligned");
  if (rows == 0) return MAVLM_OK;
  LaunchCfg lc;
#define MAVLM_LN(TI, TO, CA)                                                                                       \
  do {                                                                                                             \
    const size_t smem = 2 * static_cast<size_t>(dim) * sizeof(TO) + LN_WARPS * static_cast<size_t>(dim) * sizeof(TI) + \
                        LN_WARPS * sizeof(uint64_t);                                                              \
    static bool configured_dev[64] = {}; /* per instantiation and device */                                       \
    int dev_id = 0;                                                                                                \
    MAVLM_CUDA_OK(cudaGetDevice(&dev_id));                                                                         \
    if (!configured_dev[dev_id & 63]) {                                                                            \
      MAVLM_CUDA_OK(cudaFuncSetAttribute(layernorm_kernel<TI, TO, CA>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         232448));                                                                 \
      configured_dev[dev_id & 63] = true;                                                                          \
    }                                                                                                              \
    make_launch(lc, dim3(rows < sm_count() ? rows : sm_count()), dim3(32 * LN_WARPS), smem, st, 1, 4);             \
    MAVLM_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, layernorm_kernel<TI, TO, CA>, static_cast<const TI*>(x),             \
                                     static_cast<const TO*>(gamma), static_cast<const TO*>(beta),                 \
                                     static_cast<TO*>(y), rows, dim, eps));                                       \
  } while (0)
#define MAVLM_LN_DT(CA)                                                         \
  do {                                                                          \
    if (dtype == MAVLM_F32) MAVLM_LN(float, float, CA);                         \
    else if (dtype == MAVLM_BF16 && x_dtype == MAVLM_F32) MAVLM_LN(float, __nv_bfloat16, CA); \
    else if (dtype == MAVLM_BF16) MAVLM_LN(__nv_bfloat16, __nv_bfloat16, CA);   \
    else if (x_dtype == MAVLM_F32) MAVLM_LN(float, __half, CA);                 \
    else MAVLM_LN(__half, __half, CA);                                          \
  } while (0)
  if (dim <= 128 * 8) MAVLM_LN_DT(8);
  else if (dim <= 128 * 28) MAVLM_LN_DT(28);
  else MAVLM_LN_DT(32);
#undef MAVLM_LN_DT
#undef MAVLM_LN
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

}  // namespace mavlm

using namespace mavlm;

extern "C" {

int mavlm_pool_pe_fwd(const void* x, void* y, const float* pe_table, const int64_t* frame_idx, int frames, int side,
                      int out_side, int stride, int dim, int mode, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype_ok(dtype), MAVLM_E_INVALID, "pool_pe: bad dtype %d", dtype);
  MAVLM_REQUIRE(mode >= MAVLM_POOL_BILINEAR && mode <= MAVLM_POOL_MAX, MAVLM_E_INVALID,
                "Unexpected mm_spatial_pool_mode: %d", mode);
  const int vec = dtype == MAVLM_F32 ? 4 : 8;
  MAVLM_REQUIRE(dim > 0 && dim % vec == 0, MAVLM_E_INVALID, "pool_pe: dim %d must be a multiple of %d", dim, vec);
  MAVLM_REQUIRE(side > 0 && out_side > 0 && stride > 0, MAVLM_E_INVALID, "pool_pe: bad geometry");
  MAVLM_REQUIRE((pe_table == nullptr) == (frame_idx == nullptr), MAVLM_E_INVALID,
                "pool_pe: pe_table and frame_idx must both be given or both be NULL");
  if (mode != MAVLM_POOL_BILINEAR)
    MAVLM_REQUIRE(out_side * stride <= side, MAVLM_E_INVALID, "pool_pe: window exceeds the input");
  if (frames == 0) return MAVLM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MAVLM_DISPATCH_DTYPE(dtype, return launch_pool<T>(x, y, pe_table, frame_idx, frames, side, out_side, stride, dim, mode, st));
}

int mavlm_add_pe_fwd(const void* x, void* y, const float* pe_table, const int64_t* frame_idx, int frames, int tokens,
                     int dim, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype_ok(dtype), MAVLM_E_INVALID, "add_pe: bad dtype %d", dtype);
  const int vec = dtype == MAVLM_F32 ? 4 : 8;
  MAVLM_REQUIRE(dim > 0 && dim % vec == 0, MAVLM_E_INVALID, "add_pe: dim %d must be a multiple of %d", dim, vec);
  MAVLM_REQUIRE(pe_table != nullptr && frame_idx != nullptr, MAVLM_E_INVALID, "add_pe: NULL table / indices");
  if (frames == 0 || tokens == 0) return MAVLM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(frames) * tokens * (dim / vec);
  const int grid = grid_for(total, 256);
  LaunchCfg lc;
  make_launch(lc, dim3(grid), dim3(256), 0, st, 1, 8);
  MAVLM_DISPATCH_DTYPE(dtype, MAVLM_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, add_pe_kernel<T>, static_cast<const T*>(x),
                                                              static_cast<T*>(y), pe_table, frame_idx, frames, tokens,
                                                              dim)));
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

int mavlm_gather_rows_fwd(void* out, int64_t ld_out, const void* embed_table, int64_t n_table_rows, const void* feats,
                          int64_t n_feat_rows, const int64_t* row_src, int64_t n_rows, int dim, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype_ok(dtype), MAVLM_E_INVALID, "gather_rows: bad dtype %d", dtype);
  const int vec = dtype == MAVLM_F32 ? 4 : 8;
  MAVLM_REQUIRE(dim > 0 && dim % vec == 0 && ld_out % vec == 0, MAVLM_E_INVALID,
                "gather_rows: dim %d / ld_out must be multiples of %d", dim, vec);
  MAVLM_REQUIRE(n_table_rows >= 0 && n_feat_rows >= 0, MAVLM_E_INVALID, "gather_rows: negative table size");
  if (n_rows == 0) return MAVLM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MAVLM_DISPATCH_DTYPE(dtype, (gather_rows_kernel<T><<<static_cast<unsigned>(n_rows), 128, 0, st>>>(
                                  static_cast<T*>(out), ld_out, static_cast<const T*>(embed_table),
                                  static_cast<const T*>(feats), row_src, dim, static_cast<long long>(n_table_rows),
                                  static_cast<long long>(n_feat_rows))));
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

int mavlm_cast_fwd(const void* x, void* y, int64_t n, int src_dtype, int dst_dtype, void* stream) {
  MAVLM_REQUIRE(dtype_ok(src_dtype) && dtype_ok(dst_dtype) && src_dtype != dst_dtype &&
                    (src_dtype == MAVLM_F32 || dst_dtype == MAVLM_F32),
                MAVLM_E_INVALID, "cast: dtypes must be f32 <-> bf16 / f16");
  MAVLM_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                MAVLM_E_INVALID, "cast: pointers must be 16-byte aligned");
  if (n == 0) return MAVLM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = grid_for((n + 7) / 8, 256);
  if (src_dtype == MAVLM_F32) {
    if (dst_dtype == MAVLM_BF16)
      cast_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const float*>(x),
                                                               static_cast<__nv_bfloat16*>(y), n);
    else
      cast_kernel<float, __half><<<grid, 256, 0, st>>>(static_cast<const float*>(x), static_cast<__half*>(y), n);
  } else if (src_dtype == MAVLM_BF16) {
    cast_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                             static_cast<float*>(y), n);
  } else {
    cast_kernel<__half, float><<<grid, 256, 0, st>>>(static_cast<const __half*>(x), static_cast<float*>(y), n);
  }
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

int mavlm_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, int rows, int dim, float eps,
                        int x_dtype, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype_ok(dtype), MAVLM_E_INVALID, "layernorm: bad dtype %d", dtype);
  return layernorm_launch(x, gamma, beta, y, rows, dim, eps, x_dtype, dtype, static_cast<cudaStream_t>(stream));
}

int mavlm_assemble_fwd(void* seq, const void* mem, int64_t n_mem_rows, const void* frames, const int64_t* fine_idx,
                       int n_fine, int tokens, const void* type_emb, const void* newline, const void* embed_table,
                       const int64_t* prompt_mem_ids, int n_prompt_mem, const int64_t* prompt_frm_ids,
                       int n_prompt_frm, int dim, int drop_frames, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype_ok(dtype), MAVLM_E_INVALID, "assemble: bad dtype %d", dtype);
  const int vec = dtype == MAVLM_F32 ? 4 : 8;
  MAVLM_REQUIRE(dim > 0 && dim % vec == 0, MAVLM_E_INVALID, "assemble: dim %d must be a multiple of %d", dim, vec);
  long long rows = n_prompt_mem + n_mem_rows + 1;
  if (!drop_frames) rows += n_prompt_frm + static_cast<long long>(n_fine) * tokens + 1;
  if (mem == nullptr) rows -= n_mem_rows;  // already in place: the grid skips them
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MAVLM_DISPATCH_DTYPE(dtype, (assemble_kernel<T><<<static_cast<unsigned>(rows), 128, 0, st>>>(
                                  static_cast<T*>(seq), static_cast<const T*>(mem), n_mem_rows,
                                  static_cast<const T*>(frames), fine_idx, n_fine, tokens,
                                  static_cast<const T*>(type_emb), static_cast<const T*>(newline),
                                  static_cast<const T*>(embed_table), prompt_mem_ids, n_prompt_mem, prompt_frm_ids,
                                  n_prompt_frm, dim)));
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

}  // extern "C"
