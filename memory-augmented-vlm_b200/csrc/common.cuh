// Shared device/host helpers for the mavlm sm_100a kernels: error plumbing, PTX wrappers for
// mbarrier / TMA / tcgen05 (TMEM alloc, MMA, commit, ld/st), UMMA descriptors.
// Everything here is written for sm_100a only; there is no other-arch path.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mavlm.h"

namespace mavlm {

// ---------------------------------------------------------------------------------------
// host-side error plumbing (C-ABI never throws; see include/mavlm.h)
// ---------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define MAVLM_CUDA_OK(expr)                                                        \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) return ::mavlm::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define MAVLM_REQUIRE(cond, code, ...)      \
  do {                                      \
    if (!(cond)) {                          \
      ::mavlm::set_error(__VA_ARGS__);      \
      return (code);                        \
    }                                       \
  } while (0)

extern unsigned long long g_launches;  // kernels launched by this library (bench.py's gpu_launches)
#define MAVLM_LAUNCH_OK()                  \
  do {                                     \
    ++::mavlm::g_launches;                 \
    MAVLM_CUDA_OK(cudaGetLastError());     \
  } while (0)

int sm_count();  // cached multiProcessorCount of the current device

// Encodes a bf16 tensor map (TMA descriptor). rank 2..4, dims/strides innermost first
// (strides in bytes for dims 1..rank-1), 128B swizzle. Returns 0 / negative error.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box);
// General form: elem_bytes 2 (bf16) or 4 (fp32); swizzle_bytes 128 or 64 (the box's inner extent in bytes
// must not exceed it).
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int swizzle_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch (PDL): the hot-path kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, call griddepcontrol.launch_dependents at their start and
// griddepcontrol.wait before their first access to global memory.  The next kernel's CTAs then become resident
// as soon as SMs free up and run their prologue (barrier init, TMEM allocation, descriptor prefetch) under the
// tail of the previous kernel; the wait returns once the previous grid has completed and flushed.  The
// recurrence is ~45 dependent launches per 64 frames, so the launch-to-launch gap matters.
// MAVLM_PDL=0 in the environment (read once) or mavlm_debug_set_flags bit 4 disables it.
bool pdl_enabled(int family = 0);  // family: 1 gemm, 2 attention, 4 layernorm, 8 pool / pe (MAVLM_PDL_MASK, default all)
void pdl_force_off(bool off);

struct LaunchCfg {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
};
// cluster = 0 / 1: no cluster attribute
inline void make_launch(LaunchCfg& lc, dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, int pdl) {
  lc.cfg = cudaLaunchConfig_t{};
  lc.cfg.gridDim = grid;
  lc.cfg.blockDim = block;
  lc.cfg.dynamicSmemBytes = smem;
  lc.cfg.stream = st;
  unsigned n = 0;
  if (cluster > 1) {
    lc.attr[n].id = cudaLaunchAttributeClusterDimension;
    lc.attr[n].val.clusterDim.x = static_cast<unsigned>(cluster);
    lc.attr[n].val.clusterDim.y = 1;
    lc.attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl && pdl_enabled(pdl)) {
    lc.attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    lc.attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  lc.cfg.attrs = lc.attr;
  lc.cfg.numAttrs = n;
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------
// device-side PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// PDL (see make_launch): no-ops in a kernel launched without the attribute
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Loads of data written by the PREVIOUS kernel in a PDL kernel: they must stay behind griddepcontrol.wait.
// nvcc hoists ld.global.nc (__ldg / const __restrict__) loads above an asm volatile with a "memory" clobber
// (the LayerNorm kernel read its input before the wait that way), so these are volatile asm themselves:
// volatile asm statements keep their relative order.
__device__ __forceinline__ uint4 ld_dep_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ld_dep_u2(const void* p) {
  uint2 v;
  asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (launch failure -> error code on the
// host), never as a hung GPU.  The bound is wall-clock (~2 s of SM cycles), checked rarely.
#ifndef MAVLM_MBAR_TIMEOUT_CYCLES
#define MAVLM_MBAR_TIMEOUT_CYCLES 4000000000ll
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FF) == 0 && clock64() - t0 > MAVLM_MBAR_TIMEOUT_CYCLES) {
      printf("mavlm: mbarrier timeout block(%d,%d,%d) thread %d bar@%u parity %u\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// ---- TMA ----
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// TMA store: shared memory box -> global (bulk async group of the issuing thread); out-of-range rows /
// columns of the box are clipped by the hardware.
__device__ __forceinline__ void tma_store_2d(const void* smem_src, const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* smem_src, const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // at most N groups still reading their shared-memory source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all tcgen05.mma issued so far by this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- CTA pairs (cluster of 2, tcgen05 cta_group::2) ----
// A pair = the two CTAs of a (2,1,1) cluster on the two SMs of a TPC.  One thread of the rank-0 CTA
// ("leader") issues M=256 MMAs that read A/B halves from BOTH CTAs' shared memory (same CTA-relative
// offsets) and write 128 accumulator rows into each CTA's own TMEM.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Possibly remote arrive.  Default (.release.cta) semantics on purpose: what the waiter consumes is TMEM /
// shared-memory state ordered by tcgen05 fences, not this thread's global stores -- .release.cluster compiles
// to MEMBAR.ALL.GPU + ERRBAR, i.e. a wait for every outstanding global store of the epilogue (12 % of the
// samples of the pair GEMM).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; the transaction bytes are credited to the mbarrier at
// `bar_cluster_addr` (the leader's), the data lands in the issuing CTA's own shared memory.
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {  // warp w of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {  // warp w of BOTH CTAs
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once all MMAs issued so far have completed) on the mbarrier at this CTA-relative offset in
// every CTA of `cta_mask`.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
      "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
      "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors (bf16, 128B swizzle) ----
// Shared-memory matrix descriptor, sm_100 format (version 1): start address >> 4 in [0,14),
// LBO >> 4 in [16,30), SBO >> 4 in [32,46), version in [46,48), swizzle mode in [61,64) (2 = 128B).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// K-major operand tile [rows][64 bf16] (128 B per row, 8-row swizzle atoms of 1024 B).
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr) { return umma_smem_desc(smem_addr, 16, 1024); }
// MN-major operand tile [k rows][64 bf16 of MN] (128 B per k row): atoms of 8 k rows = 1024 B;
// lbo = byte stride between consecutive 64-wide MN groups.
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t smem_addr, uint32_t mn_group_stride_bytes) {
  return umma_smem_desc(smem_addr, mn_group_stride_bytes, 1024);
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32, M x N tile.
// `fmt`: operand format of A and B, 1 = BF16 (default), 0 = F16 (Elem16<T>::kUmmaFormat).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major, uint32_t fmt = 1) {
  return (1u << 4)                                  // D format: F32
         | (fmt << 7)                               // A format
         | (fmt << 10)                              // B format
         | (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// ---- inter-CTA hand-off through global memory (in-kernel fix-up of split work) ----
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_flag_gpu(const unsigned int* p) {  // bounded like mbar_wait
  if (ld_acquire_gpu(p) != 0u) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu(p) == 0u) {
    __nanosleep(64);
    if (clock64() - t0 > MAVLM_MBAR_TIMEOUT_CYCLES) {
      printf("mavlm: partial-result flag timeout block %d\n", blockIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ float ld_cg_f32(const float* p) {  // L2 load: data written by another CTA of this grid
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// Ampere-style asynchronous copy, 16 bytes, L2 only (data written by another CTA of this grid)
__device__ __forceinline__ void cp_async_cg16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 16-bit storage types of the tensor-core tier: conversions and the UMMA operand-format code
template <typename T>
struct Elem16;
template <>
struct Elem16<__nv_bfloat16> {
  static constexpr uint32_t kUmmaFormat = 1;  // kind::f16 A/B format BF16
  __device__ static __forceinline__ float2 unpack2(uint32_t v) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
  }
  __device__ static __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
  }
  __device__ static __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16(v); }
};
template <>
struct Elem16<__half> {
  static constexpr uint32_t kUmmaFormat = 0;  // F16
  __device__ static __forceinline__ float2 unpack2(uint32_t v) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
  }
  __device__ static __forceinline__ uint32_t pack2(float lo, float hi) {
    __half2 t = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
  }
  __device__ static __forceinline__ float to_float(__half v) { return __half2float(v); }
  __device__ static __forceinline__ __half from_float(float v) { return __float2half(v); }
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// GELU(erf) for the bf16 tensor-core epilogues:  erfc(z) = 2^(-Q(z)),  Q a degree-6 polynomial through the
// origin fitted (erfc-weighted least squares on [0, 4.2]) to -log2(erfc(z)); written in |x| = sqrt(2) z.
// |erf error| <= 3.5e-7, GELU abs error <= 7e-7 over all x (fp32 evaluation; tools/fit_gelu.py reproduces the
// coefficients and the bound) -- 4 orders below the bf16 output's half-ulp.  One MUFU (ex2) + 10 FMA-pipe
// instructions per element; libdevice erff is ~30 with a branch and the K = 1152 projector GEMM's epilogue
// (128 x 256 outputs per tile against only 18 k-blocks of MMAs) is bound by instruction issue / MUFU rate.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float ax = fminf(fabsf(x), 5.939697f);   // erfc(4.2) = 2.9e-9: erf == 1 in fp32 beyond
  float q = fmaf(-1.975574536e-05f, ax, 6.616290625e-04f);
  q = fmaf(q, ax, -7.758132054e-03f);
  q = fmaf(q, ax, 5.296287755e-02f);
  q = fmaf(q, ax, 4.590668649e-01f);
  q = fmaf(q, ax, 1.151119014e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-q * ax));
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), 1.0f - e, hx);          // 0.5 x (1 + sign(x) erf(|x| / sqrt 2))
}
#endif  // __CUDACC__

}  // namespace mavlm
