// 16-byte vector access in the storage dtype, dtype dispatch and grid sizing shared by the bandwidth-bound
// kernels (elementwise.cu, legacy_memory.cu).
#pragma once
#include "common.cuh"

namespace mavlm {

template <typename T>
struct Vec;  // 16-byte vector of T <-> floats
template <>
struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float (&v)[4]) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static void load_plain(const float* p, float (&v)[4]) {  // data of the previous kernel (see ld_dep_u4)
    const uint4 t = ld_dep_u4(p);
    v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
  }
  __device__ static void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ static void load_plain(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = ld_dep_u4(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]);
    t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

template <>
struct Vec<__half> {
  static constexpr int N = 8;
  __device__ static void unpack(const uint4& t, float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ static void load(const __half* p, float (&v)[8]) { unpack(__ldg(reinterpret_cast<const uint4*>(p)), v); }
  __device__ static void load_plain(const __half* p, float (&v)[8]) { unpack(ld_dep_u4(p), v); }
  __device__ static void store(__half* p, const float (&v)[8]) {
    uint4 t;
    t.x = Elem16<__half>::pack2(v[0], v[1]); t.y = Elem16<__half>::pack2(v[2], v[3]);
    t.z = Elem16<__half>::pack2(v[4], v[5]); t.w = Elem16<__half>::pack2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// run `expr` with T bound to the storage type of `dtype`
#define MAVLM_DISPATCH_DTYPE(dtype, ...)                         \
  do {                                                           \
    if ((dtype) == MAVLM_F32) { using T = float; __VA_ARGS__; }  \
    else if ((dtype) == MAVLM_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { using T = __half; __VA_ARGS__; }                      \
  } while (0)
static inline bool dtype_ok(int dtype) { return dtype == MAVLM_F32 || dtype == MAVLM_BF16 || dtype == MAVLM_F16; }

static inline int grid_for(long long total_threads, int block) {
  long long b = (total_threads + block - 1) / block;
  const long long cap = static_cast<long long>(sm_count()) * 32;  // grid-stride beyond 32 CTAs/SM
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace mavlm
