// Host-side runtime glue of libmavlm: thread-local error string, device check, cached SM count,
// TMA descriptor encoding through the driver entry point (no link-time libcuda dependency, so the
// library also loads on a CPU-only box for the symbol tests).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace mavlm {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in `%s`", static_cast<int>(e), cudaGetErrorString(e), file, line, what);
  return MAVLM_E_CUDA;
}

static bool g_pdl_off = false;
void pdl_force_off(bool off) { g_pdl_off = off; }
bool pdl_enabled(int family) {
  static int env = -1, mask = 0xff;
  if (env < 0) {
    const char* e = getenv("MAVLM_PDL");
    env = (e != nullptr && e[0] == '0') ? 0 : 1;
    const char* m = getenv("MAVLM_PDL_MASK");
    if (m != nullptr) mask = atoi(m);
  }
  return env == 1 && !g_pdl_off && (family == 0 || (mask & family) != 0);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  return make_tmap(out, base, 2, 128, rank, dims, strides_bytes, box);
}

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int swizzle_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  MAVLM_REQUIRE((elem_bytes == 2 || elem_bytes == 4) && (swizzle_bytes == 128 || swizzle_bytes == 64) &&
                    static_cast<int>(box[0]) * elem_bytes <= swizzle_bytes,
                MAVLM_E_INVALID, "bad TMA element size / swizzle (%d, %d, box %u)", elem_bytes, swizzle_bytes, box[0]);
  MAVLM_REQUIRE(fn != nullptr, MAVLM_E_CUDA, "cuTensorMapEncodeTiled not available from the CUDA driver");
  MAVLM_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, MAVLM_E_INVALID,
                "TMA source pointer %p is not 16-byte aligned", base);
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      MAVLM_REQUIRE((gstr[i - 1] & 15) == 0, MAVLM_E_INVALID, "TMA stride %llu is not a multiple of 16 bytes",
                    static_cast<unsigned long long>(gstr[i - 1]));
    }
  }
  CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MAVLM_REQUIRE(r == CUDA_SUCCESS, MAVLM_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return MAVLM_OK;
}

}  // namespace mavlm

extern "C" {

int mavlm_version(void) { return MAVLM_VERSION; }

const char* mavlm_last_error_string(void) { return mavlm::g_err; }

unsigned long long mavlm_launch_count(void) { return mavlm::g_launches; }

int mavlm_check_device(int device) {
  int major = 0, minor = 0;
  MAVLM_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  MAVLM_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  MAVLM_REQUIRE(major == 10, MAVLM_E_ARCH,
                "device %d is sm_%d%d; this library is built for sm_100a (B200) only and has no fallback", device,
                major, minor);
  return MAVLM_OK;
}

}  // extern "C"
