// Frame pre-processing feeding the visual-memory path: decoded uint8 frames [F, H, W, 3]
// -> SigLipImageProcessor.preprocess (siglip_encoder.py:47-67, called from train.py:1239 on the `.pt` video
// tensor of extract_video_frames/video_reader_tmp.py:87) -> pixel_values [F, 3, 384, 384].
//
// The reference runs this per frame on the host through PIL: Image.resize(BICUBIC) is Pillow's separable
// two-pass resampler (src/libImaging/Resample.c) in 22-bit fixed point with a uint8 intermediate image,
// horizontal pass first, a pass being skipped when that dimension already has the target size; then
// transformers' rescale (float64(u8) * 1/255 -> float32) and normalize ((x - mean) / std in float32).
// Here: the coefficient tables are built by mavlm_resize_coeffs (host, the same double arithmetic), the
// horizontal pass is one kernel (uint8 -> uint8) and the vertical pass + rescale + normalize + HWC->CHW
// (+ cast to the tower dtype) a second one.  Integer / byte work, bit-exact with PIL; HBM-bound:
// algorithmic bytes per frame = 3*(H*W + H*384) for pass 1, 3*H*384 + 3*384*384*out_bytes for pass 2.
#include <cmath>

#include "common.cuh"

namespace mavlm {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

static inline double rs_bicubic(double x) {  // Resample.c bicubic_filter, a = -0.5
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

__device__ __forceinline__ int clip8(int acc) {
  const int v = acc >> RS_PRECISION_BITS;  // arithmetic shift, like the C reference
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: in [rows][W][3] -> out [rows][OW][3].  One CTA per input row: the row is staged in shared
// memory with 16-byte loads (every byte of the frame crosses HBM -> SM exactly once, coalesced), each thread
// filters output pixels out of shared memory, and the output row leaves through shared memory as 16-byte stores.
constexpr int RS_H_THREADS = 128;
__global__ void __launch_bounds__(RS_H_THREADS) resize_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                                const int* __restrict__ bounds,
                                                                const int* __restrict__ kk, int ksize, int W, int OW) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  const long long row = blockIdx.x;
  const int in_bytes = W * 3, out_bytes = OW * 3;
  const uint8_t* src = in + row * in_bytes;
  uint8_t* dst = out + row * out_bytes;
  // stage the input row; `lead` bytes of slack so that the 16-byte loads are aligned in global AND shared memory
  const int lead = static_cast<int>(reinterpret_cast<uintptr_t>(src) & 15);
  uint8_t* srow = rs_smem;                              // holds src[-lead .. in_bytes) rounded up to 16
  const int nvec = (lead + in_bytes + 15) >> 4;
  const uint4* gsrc = reinterpret_cast<const uint4*>(src - lead);
  for (int i = threadIdx.x; i < nvec; i += RS_H_THREADS) {
    // the first / last vector may straddle the neighbouring rows (same allocation) except at the very ends
    const bool edge = (i == 0 && lead != 0 && row == 0) || (i == nvec - 1 && row == gridDim.x - 1 && ((lead + in_bytes) & 15));
    if (!edge) {
      reinterpret_cast<uint4*>(srow)[i] = __ldg(gsrc + i);
    } else {
      for (int b = 0; b < 16; ++b) {
        const int off = i * 16 + b - lead;
        srow[i * 16 + b] = (off >= 0 && off < in_bytes) ? src[off] : 0;
      }
    }
  }
  uint8_t* orow = rs_smem + ((lead + in_bytes + 15) & ~15) + 16;
  __syncthreads();
  const uint8_t* px = srow + lead;
  for (int ox = threadIdx.x; ox < OW; ox += RS_H_THREADS) {
    const int xmin = __ldg(bounds + 2 * ox), cnt = __ldg(bounds + 2 * ox + 1);
    const int* k = kk + static_cast<long long>(ox) * ksize;
    const uint8_t* p = px + xmin * 3;
    int a0 = 1 << (RS_PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int x = 0; x < cnt; ++x) {
      const int c = __ldg(k + x);
      a0 += p[3 * x] * c;
      a1 += p[3 * x + 1] * c;
      a2 += p[3 * x + 2] * c;
    }
    orow[3 * ox] = static_cast<uint8_t>(clip8(a0));
    orow[3 * ox + 1] = static_cast<uint8_t>(clip8(a1));
    orow[3 * ox + 2] = static_cast<uint8_t>(clip8(a2));
  }
  __syncthreads();
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (out_bytes & 15) == 0) {
    for (int i = threadIdx.x; i < (out_bytes >> 4); i += RS_H_THREADS)
      reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(orow)[i];
  } else {
    for (int i = threadIdx.x; i < out_bytes; i += RS_H_THREADS) dst[i] = orow[i];
  }
}

// vertical pass (ksize == 0: identity) + rescale + normalize + HWC -> CHW; in [F][H][OW][3].  One thread per
// XV consecutive output pixels of a row: a tap is 3*XV contiguous bytes (32-bit loads when aligned), the three
// channel planes are written as XV-element vectors.
constexpr int RS_XV = 4;
template <typename TO>
__global__ void __launch_bounds__(256) resize_v_norm_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ u8_out,
                                                            TO* __restrict__ out, const int* __restrict__ bounds,
                                                            const int* __restrict__ kk, int ksize, int frames, int H,
                                                            int OH, int OW, double rescale, float m0, float m1, float m2,
                                                            float s0, float s1, float s2) {
  const int groups = (OW + RS_XV - 1) / RS_XV;
  const long long total = static_cast<long long>(frames) * OH * groups;
  const bool aligned = (OW % RS_XV == 0) && ((reinterpret_cast<uintptr_t>(in) & 3) == 0);
  const long long pitch = static_cast<long long>(OW) * 3;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    const long long t = i / groups;
    const int oy = static_cast<int>(t % OH);
    const int f = static_cast<int>(t / OH);
    const int ox0 = g * RS_XV;
    const int nx = min(RS_XV, OW - ox0);
    int acc[3 * RS_XV];
    int ymin = oy, cnt = 1;
    const int* k = nullptr;
    if (ksize > 0) {
      ymin = __ldg(bounds + 2 * oy);
      cnt = __ldg(bounds + 2 * oy + 1);
      k = kk + static_cast<long long>(oy) * ksize;
#pragma unroll
      for (int j = 0; j < 3 * RS_XV; ++j) acc[j] = 1 << (RS_PRECISION_BITS - 1);
    } else {
#pragma unroll
      for (int j = 0; j < 3 * RS_XV; ++j) acc[j] = 0;
    }
    const uint8_t* p = in + ((static_cast<long long>(f) * H + ymin) * OW + ox0) * 3;
    for (int y = 0; y < cnt; ++y) {
      const int c = ksize > 0 ? __ldg(k + y) : 1;
      if (aligned) {
        const uint32_t* pw = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
        for (int w = 0; w < 3 * RS_XV / 4; ++w) {
          const uint32_t v = __ldg(pw + w);
          acc[4 * w] += static_cast<int>(v & 0xff) * c;
          acc[4 * w + 1] += static_cast<int>((v >> 8) & 0xff) * c;
          acc[4 * w + 2] += static_cast<int>((v >> 16) & 0xff) * c;
          acc[4 * w + 3] += static_cast<int>(v >> 24) * c;
        }
      } else {
        for (int j = 0; j < 3 * nx; ++j) acc[j] += p[j] * c;
      }
      p += pitch;
    }
    int v[3 * RS_XV];
#pragma unroll
    for (int j = 0; j < 3 * RS_XV; ++j) v[j] = ksize > 0 ? clip8(acc[j]) : acc[j];
    const long long pix = (static_cast<long long>(f) * OH + oy) * OW + ox0;
    if (u8_out != nullptr) {
      uint8_t* o = u8_out + pix * 3;
      for (int j = 0; j < 3 * nx; ++j) o[j] = static_cast<uint8_t>(v[j]);
    }
    if (out != nullptr) {
      // transformers rescale: float64(u8) * scale -> float32; normalize: (x - mean) / std in float32 (IEEE division)
      const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
      const long long plane = static_cast<long long>(OH) * OW;
      TO* o = out + static_cast<long long>(f) * 3 * plane + static_cast<long long>(oy) * OW + ox0;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        TO r[RS_XV];
#pragma unroll
        for (int x = 0; x < RS_XV; ++x)
          r[x] = static_cast<TO>(__fdiv_rn(static_cast<float>(static_cast<double>(v[3 * x + ch]) * rescale) - mean[ch], sd[ch]));
        TO* oc = o + ch * plane;
        if (nx == RS_XV && (reinterpret_cast<uintptr_t>(oc) & (sizeof(TO) * RS_XV - 1)) == 0) {
          if (sizeof(TO) == 4) *reinterpret_cast<uint4*>(oc) = *reinterpret_cast<const uint4*>(r);
          else *reinterpret_cast<uint2*>(oc) = *reinterpret_cast<const uint2*>(r);
        } else {
          for (int x = 0; x < nx; ++x) oc[x] = r[x];
        }
      }
    }
  }
}

static int rs_grid(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

}  // namespace mavlm

using namespace mavlm;

extern "C" {

int mavlm_resize_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* kk, int ksize_capacity) {
  MAVLM_REQUIRE(in_size > 0 && out_size > 0, MAVLM_E_INVALID, "resize_coeffs: sizes must be positive");
  // Pillow Resample.c precompute_coeffs (box = whole axis) + normalize_coeffs_8bpc, bicubic
  const double scale = static_cast<double>(static_cast<float>(in_size) - 0.0f) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  if (bounds == nullptr || kk == nullptr) return ksize;  // size query
  MAVLM_REQUIRE(ksize_capacity >= ksize, MAVLM_E_INVALID, "resize_coeffs: kk needs %d columns, %d given", ksize,
                ksize_capacity);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double k[512];
    MAVLM_REQUIRE(xmax <= 512, MAVLM_E_INVALID, "resize_coeffs: filter too wide (%d taps)", xmax);
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      const double w = rs_bicubic((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    int32_t* row = kk + static_cast<long long>(xx) * ksize_capacity;
    for (int x = 0; x < ksize_capacity; ++x) row[x] = 0;
    for (int x = 0; x < xmax; ++x) {
      double w = k[x];
      if (ww != 0.0) w /= ww;
      row[x] = w < 0 ? static_cast<int>(-0.5 + w * (1 << RS_PRECISION_BITS))
                     : static_cast<int>(0.5 + w * (1 << RS_PRECISION_BITS));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return ksize;
}

int mavlm_frames_preprocess_fwd(const uint8_t* frames, int n_frames, int in_h, int in_w, uint8_t* tmp,
                                uint8_t* resized_u8, void* pixel_values, int out_h, int out_w,
                                const int32_t* bounds_h, const int32_t* kk_h, int ksize_h, const int32_t* bounds_v,
                                const int32_t* kk_v, int ksize_v, double rescale, const float* mean3,
                                const float* std3, int out_dtype, void* stream) {
  MAVLM_REQUIRE(out_dtype == MAVLM_F32 || out_dtype == MAVLM_BF16 || out_dtype == MAVLM_F16, MAVLM_E_INVALID,
                "preprocess: bad dtype %d", out_dtype);
  MAVLM_REQUIRE(n_frames >= 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, MAVLM_E_INVALID,
                "preprocess: bad geometry");
  MAVLM_REQUIRE(mean3 != nullptr && std3 != nullptr, MAVLM_E_INVALID, "preprocess: NULL mean / std");
  MAVLM_REQUIRE((in_w == out_w) == (bounds_h == nullptr) && (in_w == out_w) == (kk_h == nullptr), MAVLM_E_INVALID,
                "preprocess: horizontal tables must be given exactly when the width changes");
  MAVLM_REQUIRE((in_h == out_h) == (bounds_v == nullptr) && (in_h == out_h) == (kk_v == nullptr), MAVLM_E_INVALID,
                "preprocess: vertical tables must be given exactly when the height changes");
  if (n_frames == 0) return MAVLM_OK;
  MAVLM_REQUIRE(in_w == out_w || tmp != nullptr, MAVLM_E_WORKSPACE,
                "preprocess: the horizontal pass needs a [F, H, out_w, 3] uint8 workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* vin = frames;
  if (in_w != out_w) {
    const long long rows = static_cast<long long>(n_frames) * in_h;
    MAVLM_REQUIRE(rows < (1ll << 31), MAVLM_E_INVALID, "preprocess: too many rows (%lld)", rows);
    const size_t smem = static_cast<size_t>(((in_w * 3 + 15 + 15) & ~15) + 16 + ((out_w * 3 + 15) & ~15));
    MAVLM_REQUIRE(smem <= 48 * 1024, MAVLM_E_INVALID, "preprocess: rows of %d pixels are too wide for the staging buffer",
                  in_w);
    resize_h_kernel<<<static_cast<unsigned>(rows), RS_H_THREADS, smem, st>>>(frames, tmp, bounds_h, kk_h, ksize_h, in_w,
                                                                              out_w);
    MAVLM_LAUNCH_OK();
    vin = tmp;
  }
  const long long total = static_cast<long long>(n_frames) * out_h * ((out_w + RS_XV - 1) / RS_XV);
  const int kv = in_h != out_h ? ksize_v : 0;
  if (out_dtype == MAVLM_F32)
    resize_v_norm_kernel<float><<<rs_grid(total), 256, 0, st>>>(vin, resized_u8, static_cast<float*>(pixel_values),
                                                                bounds_v, kk_v, kv, n_frames, in_h, out_h, out_w, rescale,
                                                                mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  else if (out_dtype == MAVLM_F16)
    resize_v_norm_kernel<__half><<<rs_grid(total), 256, 0, st>>>(
        vin, resized_u8, static_cast<__half*>(pixel_values), bounds_v, kk_v, kv, n_frames, in_h, out_h, out_w, rescale,
        mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  else
    resize_v_norm_kernel<__nv_bfloat16><<<rs_grid(total), 256, 0, st>>>(
        vin, resized_u8, static_cast<__nv_bfloat16*>(pixel_values), bounds_v, kk_v, kv, n_frames, in_h, out_h, out_w,
        rescale, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

}  // extern "C"
