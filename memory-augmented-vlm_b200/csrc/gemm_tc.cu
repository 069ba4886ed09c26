// bf16 GEMM on the 5th-gen tensor cores (sm_100a):  C = act(A * W^T + bias) (+ resid) (+ addvec)
//   A [M,K] row-major (K-major operand), W [N,K] row-major (nn.Linear weight, K-major operand).
// Persistent, warp-specialised CTA (320 threads, one CTA per SM; CG = 2: a CTA pair per 256-row tile):
//   warp 0  : TMA producer  (cp.async.bulk.tensor, 128B swizzle, STAGES-deep smem ring, mbarrier tx)
//   warp 1  : MMA issuer    (one elected lane issues tcgen05.mma cta_group::1, M=128 x N=BN x K=16;
//                            accumulators double-buffered in TMEM so the epilogue of tile i overlaps
//                            the main loop of tile i+1; tcgen05.commit releases smem stages)
//   warps 2-9: epilogue     (tcgen05.ld 32x32b -> registers -> bias / GELU(erf) / ReLU / residual ->
//                            bf16 or fp32 rows, 16-byte vector stores); two warps per TMEM lane quadrant,
//                            each taking half of the tile's columns (the erf epilogue of the K=1152
//                            projector GEMM was epilogue-bound with four warps: 55 % tensor-active)
// Tiles are rasterised in groups of M-tiles (sized on the host so that a group's A rows stay in L2), m-fastest
// inside a group and sweeping N, so the tiles in flight share a few A row-blocks and W slabs; plain
// m-fastest order re-read A once per N-slab wave (3.2 GB of DRAM reads for the 12544x14336x3584 GEMM).
#include "common.cuh"

namespace mavlm {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;                         // two warps per TMEM lane quadrant, splitting the columns
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;   // 320

struct GemmTcParams {
  int M, N, K;
  const void* bias;      // bias / resid / addvec: 16-bit storage type T of the kernel (bf16 or fp16)
  const void* resid;
  long long ldr;
  const void* addvec;
  // fused temporal PE (position_encoding.py:57-64): C[row] += pe_table[frame_idx[row / pe_tokens]] (fp32 table)
  const float* pe_table;
  const long long* pe_idx;
  int pe_tokens;
  int half;        // 1: operands / outputs are fp16 (kernels instantiated with T = __half)
  void* C;
  long long ldc;
  int act;
  int out_f32;
  int m_tiles, n_tiles;
  int group_m;     // M-tiles per rasterisation group
  // general / batched mode (backward pass): problem index = outer * inner + inner_idx
  int batches, inner;
  long long c_outer, c_inner;
  int accumulate;  // C += result
  int tma_store;   // the epilogue stages C rows in shared memory and writes them with TMA stores (tmC)
  // tile range of this launch: [tile_begin, tile_end) of the m_tiles * n_tiles * batches tiles (rasterised order)
  int tile_begin, tile_end;
  // tail fill (problem 2 of a launch only): the workers that would idle in the primary's last wave -- workers
  // [fill_first_idle, workers) -- walk this problem's tiles tile_begin + (worker - fill_first_idle) + k * fill_n_idle
  int fill_first_idle, fill_n_idle;
  // EPI = 1 (probability column sums, MemoryController.py:135): rows = keys, columns = queries of one (batch, head)
  // problem; the epilogue adds sum_q exp2(c * cs_scale_log2 - lse[q] * log2 e) of its columns to cs_out[b][key]
  const float* cs_lse;   // [batches][N] natural-log LSE of the attention forward
  float* cs_out;         // [outer][cs_out_stride] (summed over the inner = head problems and all query tiles)
  float cs_scale_log2;
  long long cs_out_stride;
};

// CG = CTAs per MMA (tcgen05 cta_group): 1 = one CTA owns a 128 x BN tile; 2 = a CTA pair owns a 256 x BN
// tile, each CTA staging its own 128 rows of A and HALF of the W slab (BN/2 rows), so the shared-memory
// bytes read per MMA cycle drop by a third and a stage is 32 KB instead of 48 KB (deeper ring).
// bf16 / fp16 TMA stores use 32 x 64-column boxes when every epilogue warp converts an even number of 32-column chunks
__host__ __device__ constexpr bool gemm_wide_store(int bn) { return ((bn / 32) / 2) % 2 == 0; }

template <int BN, int CG = 1>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_ROWS = BN / CG;                // W rows staged by one CTA
  static constexpr int B_BYTES = B_ROWS * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = CG == 2 ? ((BN >= 256) ? 6 : (BN >= 192 ? 6 : 8))
                                        : ((BN >= 256) ? 4 : (BN >= 192 ? 4 : (BN >= 128 ? 6 : 8)));
  // per epilogue warp: a 4 KB staging tile for TMA stores (bf16: two 32 x 64 B buffers, fp32: one 32 x 128 B)
  static constexpr int OUT_STAGE_BYTES = 4096;
  static constexpr int ACC_STRIDE = (BN <= 64) ? 64 : (BN <= 128 ? 128 : 256);
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  // ring | output staging | barriers | tmem ptr ; +1024 for manual alignment
  static constexpr int SMEM_BYTES =
      STAGES * STAGE_BYTES + GEMM_EPI_WARPS * OUT_STAGE_BYTES + (2 * STAGES + 4) * 8 + 16 + 1024;
};

// Tile rasterisation: N is swept inside groups of `group` M-tiles (m-fastest inside a group), so the tiles in
// flight share `group` A row-blocks and ~workers/group W slabs out of L2, and W is re-read once per group.
// The host sizes the group so that a group's A rows stay L2-resident through the sweep (gemm_group_m).
__device__ __forceinline__ void gemm_tile_coords(int tile, int m_tiles, int n_tiles, int group, int& m_blk,
                                                 int& n_blk) {
  const int GROUP = group;
  const int per_group = GROUP * n_tiles;
  const int g = tile / per_group, r = tile - g * per_group;
  const int first_m = g * GROUP;
  const int gm = min(GROUP, m_tiles - first_m);
  m_blk = first_m + r % gm;
  n_blk = r / gm;
}

// A_MN / B_MN: the operand is stored with its M (resp. N) index contiguous ([K, M] / [K, N] row-major: the
// transposed operands of dgrad / wgrad / attention backward).  Such a tile is loaded as 64x64 boxes
// [64 k-rows x 64 m] and consumed as an MN-major UMMA operand (8-k-row atoms of 1 KB, 64-wide M groups 8 KB apart).
template <int BN, bool A_MN, bool B_MN, int CG, typename T = __nv_bfloat16, int EPI = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ GemmTcParams p,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
               const __grid_constant__ CUtensorMap tmC2, const __grid_constant__ GemmTcParams p2, const int has2) {
  static_assert(CG == 1 || CG == 2, "one CTA or a CTA pair per tile");
  static_assert(CG == 1 || !B_MN || (BN / CG) % 64 == 0, "an MN-major W half must be whole 64-column boxes");
  using Cfg = GemmCfg<BN, CG>;
  constexpr int TILE_M = GEMM_BM * CG;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sOut = smem + STAGES * Cfg::STAGE_BYTES;  // 1024-byte aligned (every stage size is a multiple of 1 KB)
  uint64_t* full = reinterpret_cast<uint64_t*>(sOut + GEMM_EPI_WARPS * Cfg::OUT_STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A launch walks up to two problems with the same tile shape: the primary (all workers, round robin over its tile
  // range) and, for TAIL FILL, a second problem whose tiles go only to the workers that would idle in the primary's
  // last wave (GemmTcParams::fill_*).  Every role runs the same two loops, so the smem / TMEM pipelines just continue.
#define MAVLM_GEMM_PROBLEM(pi)                                                                            \
  const GemmTcParams& pp = (pi) ? p2 : p;                                                                 \
  const CUtensorMap* ta = (pi) ? &tmA2 : &tmA;                                                            \
  const CUtensorMap* tb = (pi) ? &tmB2 : &tmB;                                                            \
  const CUtensorMap* tc = (pi) ? &tmC2 : &tmC;                                                            \
  (void)ta; (void)tb; (void)tc;                                                                           \
  if ((pi) && worker < pp.fill_first_idle) break;                                                         \
  const int t_first = pp.tile_begin + ((pi) ? worker - pp.fill_first_idle : worker);                      \
  const int t_step = (pi) ? pp.fill_n_idle : workers;                                                     \
  const int tiles_per_batch = pp.m_tiles * pp.n_tiles;                                                    \
  const int kblocks = (pp.K + GEMM_BK - 1) / GEMM_BK;                                                     \
  (void)tiles_per_batch; (void)kblocks
  // CTA pair: rank 0 is the leader (issues the MMAs, owns the full / tempty barriers both CTAs signal)
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int worker = CG == 2 ? blockIdx.x / 2 : blockIdx.x;       // persistent tile walker (CTA or pair)
  const int workers = CG == 2 ? gridDim.x / 2 : gridDim.x;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.tma_store) prefetch_tmap(&tmC);
    if (has2) {
      prefetch_tmap(&tmA2);
      prefetch_tmap(&tmB2);
      if (p2.tma_store) prefetch_tmap(&tmC2);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], GEMM_EPI_WARPS * CG);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  // griddepcontrol.wait: everything above overlapped the previous kernel's tail; memory written by that kernel is
  // touched only behind the wait (each role calls it before its first such access)

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      pdl_wait();
      for (int pi = 0; pi <= has2; ++pi) {
      MAVLM_GEMM_PROBLEM(pi);
      for (int tile = t_first; tile < pp.tile_end; tile += t_step) {
        int m_blk, n_blk;
        const int batch = tile / tiles_per_batch;
        gemm_tile_coords(tile - batch * tiles_per_batch, pp.m_tiles, pp.n_tiles, pp.group_m, m_blk, n_blk);
        const int m0 = m_blk * TILE_M + static_cast<int>(rank) * GEMM_BM;
        const int n0 = n_blk * BN + static_cast<int>(rank) * Cfg::B_ROWS;
        const int bi = batch % pp.inner, bo = batch / pp.inner;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a_dst = sA + stage * Cfg::A_BYTES;
          uint8_t* b_dst = sB + stage * Cfg::B_BYTES;
          if (CG == 2) {
            // both CTAs' boxes are credited to the leader's barrier, which expects the pair's bytes
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
            const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
            if (!A_MN) {
              tma_load_4d_pair(a_dst, ta, bar, kb * GEMM_BK, m0, bi, bo);
            } else {
#pragma unroll
              for (int i = 0; i < GEMM_BM / 64; ++i)
                tma_load_4d_pair(a_dst + i * 8192, ta, bar, m0 + 64 * i, kb * GEMM_BK, bi, bo);
            }
            if (!B_MN) {
              tma_load_4d_pair(b_dst, tb, bar, kb * GEMM_BK, n0, bi, bo);
            } else {
#pragma unroll
              for (int i = 0; i < Cfg::B_ROWS / 64; ++i)
                tma_load_4d_pair(b_dst + i * 8192, tb, bar, n0 + 64 * i, kb * GEMM_BK, bi, bo);
            }
          } else {
            mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
            if (!A_MN) {
              tma_load_4d(a_dst, ta, &full[stage], kb * GEMM_BK, m0, bi, bo);
            } else {
#pragma unroll
              for (int i = 0; i < GEMM_BM / 64; ++i)
                tma_load_4d(a_dst + i * 8192, ta, &full[stage], m0 + 64 * i, kb * GEMM_BK, bi, bo);
            }
            if (!B_MN) {
              tma_load_4d(b_dst, tb, &full[stage], kb * GEMM_BK, n0, bi, bo);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_4d(b_dst + i * 8192, tb, &full[stage], n0 + 64 * i, kb * GEMM_BK, bi, bo);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {  // (touches shared memory / TMEM only: no wait needed)
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN, A_MN ? 1 : 0, B_MN ? 1 : 0, Elem16<T>::kUmmaFormat);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int pi = 0; pi <= has2; ++pi) {
      MAVLM_GEMM_PROBLEM(pi);
      for (int tile = t_first; tile < pp.tile_end; tile += t_step) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES), b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
          const uint64_t a_desc = A_MN ? umma_desc_mnmajor(a_addr, 8192) : umma_desc_kmajor(a_addr);
          const uint64_t b_desc = B_MN ? umma_desc_mnmajor(b_addr, 8192) : umma_desc_kmajor(b_addr);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {  // K-major: +32 B inside the 128 B swizzle row; MN-major: +16 k-rows = 2 KB
            if (CG == 2)
              umma_bf16_pair(d_tmem, a_desc + (A_MN ? 128 : 2) * k, b_desc + (B_MN ? 128 : 2) * k, idesc, (kb | k) != 0);
            else
              umma_bf16(d_tmem, a_desc + (A_MN ? 128 : 2) * k, b_desc + (B_MN ? 128 : 2) * k, idesc, (kb | k) != 0);
          }
          if (CG == 2) umma_commit_pair(&empty[stage], 0b11);  // frees the stage in both CTAs
          else umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (CG == 2) umma_commit_pair(&tfull[acc], 0b11);
        else umma_commit(&tfull[acc]);
        if ((acc ^= 1) == 0) acc_phase ^= 1;
      }
      }
    }
  } else {
    pdl_wait();                                 // bias / resid / C of an accumulating GEMM come from earlier kernels
    const int q = warp & 3;                     // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;           // which half of the tile's columns this warp converts
    const int row_in_tile = q * 32 + lane;
    constexpr int CHUNKS = BN / 32, CPH = CHUNKS / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t tempty_leader = CG == 2 ? mapa_u32(smem_u32(&tempty[0]), 0) : 0u;
    uint8_t* out_stage = sOut + (warp - 2) * Cfg::OUT_STAGE_BYTES;
    for (int pi = 0; pi <= has2; ++pi) {
      MAVLM_GEMM_PROBLEM(pi);
      for (int tile = t_first; tile < pp.tile_end; tile += t_step) {
      int m_blk, n_blk;
      const int batch = tile / tiles_per_batch;
      gemm_tile_coords(tile - batch * tiles_per_batch, pp.m_tiles, pp.n_tiles, pp.group_m, m_blk, n_blk);
      const int m0 = m_blk * TILE_M + static_cast<int>(rank) * GEMM_BM, n0 = n_blk * BN;
      const long long c_off = (batch / pp.inner) * pp.c_outer + (batch % pp.inner) * pp.c_inner;
      const int row = m0 + row_in_tile;
      const bool row_ok = row < pp.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::ACC_STRIDE;

      float cs_acc = 0.f;  // EPI == 1: this thread's row (key) sum over the tile's columns (queries)
      // one 32-column chunk of this thread's row: + bias -> activation -> (+ resid) (+ addvec) -> store
      auto emit = [&](const uint32_t (&r)[32], int c) {
        const int nc = n0 + c * 32;
        if (nc >= pp.N) return;                      // warp-uniform
        if (!row_ok && !pp.tma_store) return;        // TMA-store mode keeps the whole warp in the protocol
        const bool full_chunk = nc + 32 <= pp.N;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if constexpr (EPI == 1) {
          // this thread's key against 32 queries: normalised probabilities from the forward's LSE, summed over queries
          if (row_ok) {
            const float* l = pp.cs_lse + static_cast<long long>(batch) * pp.N + nc;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (full_chunk || nc + j < pp.N) {
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v[j], pp.cs_scale_log2, -1.44269504088896340736f * __ldg(l + j))));
                cs_acc += e;
              }
            }
          }
          return;
        }
        if constexpr (EPI == 2) {
          // attention backward, scores -> probabilities: P = exp2(s * scale * log2 e - lse[row] * log2 e), stored in T
          // (the unfused form wrote the fp32 scores and re-read them in a separate pass)
          const float l2 = row_ok ? -1.44269504088896340736f * __ldg(pp.cs_lse + static_cast<long long>(batch) * pp.M + row) : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v[j], pp.cs_scale_log2, l2)));
            v[j] = e;
          }
        }
        if constexpr (EPI == 3) {
          // attention backward, dP -> dS IN PLACE over P: dS = P * (dP - D[row]) * scale  (cs_lse holds D = rowsum(dO * O),
          // cs_scale_log2 the plain softmax scale; C holds P on entry and dS on exit -- same thread, same addresses)
          if (row_ok) {
            const float dsum = __ldg(pp.cs_lse + static_cast<long long>(batch) * pp.M + row);
            const T* pr = static_cast<const T*>(pp.C) + c_off + row * pp.ldc + nc;
            if (full_chunk) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 t = ld_dep_u4(reinterpret_cast<const uint4*>(pr) + g);
                const uint32_t* h = reinterpret_cast<const uint32_t*>(&t);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = Elem16<T>::unpack2(h[e]);
                  v[g * 8 + 2 * e] = f.x * (v[g * 8 + 2 * e] - dsum) * pp.cs_scale_log2;
                  v[g * 8 + 2 * e + 1] = f.y * (v[g * 8 + 2 * e + 1] - dsum) * pp.cs_scale_log2;
                }
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (nc + j < pp.N) v[j] = Elem16<T>::to_float(pr[j]) * (v[j] - dsum) * pp.cs_scale_log2;
            }
          }
        }
        // v += src[0..32).  `dep`: src was written by the previous kernel (resid) -> ordered load, see ld_dep_u4
        auto add_vec32 = [&](const T* src, bool dep) {
          if (full_chunk) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 t = dep ? ld_dep_u4(reinterpret_cast<const uint4*>(src) + g)
                                  : __ldg(reinterpret_cast<const uint4*>(src) + g);
              const uint32_t* h = reinterpret_cast<const uint32_t*>(&t);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = Elem16<T>::unpack2(h[e]);
                v[g * 8 + 2 * e] += f.x;
                v[g * 8 + 2 * e + 1] += f.y;
              }
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (nc + j < pp.N) v[j] += Elem16<T>::to_float(src[j]);
          }
        };
        if (pp.bias != nullptr) add_vec32(static_cast<const T*>(pp.bias) + nc, false);
        if (pp.act == MAVLM_ACT_GELU_ERF) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
        } else if (pp.act == MAVLM_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (pp.resid != nullptr && row_ok) add_vec32(static_cast<const T*>(pp.resid) + row * pp.ldr + nc, true);
        if (pp.addvec != nullptr) add_vec32(static_cast<const T*>(pp.addvec) + nc, false);
        if (pp.pe_table != nullptr && row_ok) {  // the frame's PE row (fp32, L2-resident: 196 rows share it)
          const float* pe = pp.pe_table + __ldg(pp.pe_idx + row / pp.pe_tokens) * pp.N + nc;
          if (full_chunk) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(pe) + g);
              v[4 * g] += t.x; v[4 * g + 1] += t.y; v[4 * g + 2] += t.z; v[4 * g + 3] += t.w;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (nc + j < pp.N) v[j] += pe[j];
          }
        }
        if (pp.tma_store) {
          // Coalesced output: the warp's 32 rows x 32 columns go to a swizzled shared-memory tile (thread = row,
          // conflict-free 16-byte writes) and leave as ONE TMA store, which also clips the M / N tails.  Direct
          // register stores (32 rows x 16 B per instruction, half sectors) cost 30 % of the K = 1152 projector GEMM.
          if (pp.out_f32) {
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            uint8_t* dst = out_stage + lane * 128;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<float4*>(dst + ((g ^ (lane & 7)) << 4)) =
                  make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(out_stage, tc, nc, m0 + q * 32);
              bulk_commit();
            }
          } else if constexpr (gemm_wide_store(BN)) {
            // an even number of chunks per warp: two chunks share one 32 x 128 B tile (128B swizzle) and leave as ONE
            // TMA store of whole 128-byte lines -- half as many stores as 32 x 64 B tiles, which bound the epilogue of
            // the short-K projector GEMM (64 stores per 128 x 256 tile and CTA)
            const int sub = c & 1;                                   // this warp's first chunk index is even
            const bool closes = sub == 1 || nc + 32 >= pp.N;         // the odd chunk would be past N: store now
            if (sub == 0) {
              if (lane == 0) bulk_wait_read<0>();                    // the previous store has read the tile
              __syncwarp();
            }
            uint8_t* dst = out_stage + lane * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 t;
              t.x = Elem16<T>::pack2(v[8 * g], v[8 * g + 1]);
              t.y = Elem16<T>::pack2(v[8 * g + 2], v[8 * g + 3]);
              t.z = Elem16<T>::pack2(v[8 * g + 4], v[8 * g + 5]);
              t.w = Elem16<T>::pack2(v[8 * g + 6], v[8 * g + 7]);
              *reinterpret_cast<uint4*>(dst + (((4 * sub + g) ^ (lane & 7)) << 4)) = t;  // 128B swizzle
            }
            if (closes) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(out_stage, tc, nc - 32 * sub, m0 + q * 32);
                bulk_commit();
              }
            }
          } else {
            uint8_t* buf = out_stage + (c & 1) * 2048;
            if (lane == 0) bulk_wait_read<1>();  // the store issued two chunks ago has read this buffer
            __syncwarp();
            uint8_t* dst = buf + lane * 64;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 t;
              t.x = Elem16<T>::pack2(v[8 * g], v[8 * g + 1]);
              t.y = Elem16<T>::pack2(v[8 * g + 2], v[8 * g + 3]);
              t.z = Elem16<T>::pack2(v[8 * g + 4], v[8 * g + 5]);
              t.w = Elem16<T>::pack2(v[8 * g + 6], v[8 * g + 7]);
              *reinterpret_cast<uint4*>(dst + ((g ^ ((lane >> 1) & 3)) << 4)) = t;  // 64B swizzle
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(buf, tc, nc, m0 + q * 32);
              bulk_commit();
            }
          }
          return;
        }
        if (pp.out_f32) {
          float* cp = static_cast<float*>(pp.C) + c_off + row * pp.ldc + nc;
          if (pp.accumulate) {
            for (int j = 0; j < 32; ++j)
              if (nc + j < pp.N) v[j] += cp[j];
          }
          if (full_chunk) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              reinterpret_cast<float4*>(cp)[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (nc + j < pp.N) cp[j] = v[j];
          }
        } else {
          T* cp = static_cast<T*>(pp.C) + c_off + row * pp.ldc + nc;
          if (pp.accumulate) {
            for (int j = 0; j < 32; ++j)
              if (nc + j < pp.N) v[j] += Elem16<T>::to_float(cp[j]);
          }
          if (full_chunk) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 t;
              t.x = Elem16<T>::pack2(v[8 * g], v[8 * g + 1]);
              t.y = Elem16<T>::pack2(v[8 * g + 2], v[8 * g + 3]);
              t.z = Elem16<T>::pack2(v[8 * g + 4], v[8 * g + 5]);
              t.w = Elem16<T>::pack2(v[8 * g + 6], v[8 * g + 7]);
              reinterpret_cast<uint4*>(cp)[g] = t;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (nc + j < pp.N) cp[j] = Elem16<T>::from_float(v[j]);
          }
        }
      };
      // the accumulator buffer goes back to the MMA warp as soon as this warp's last TMEM load has landed
      // in registers, before the chunk is converted and stored
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(tempty_leader + acc * 8);  // the leader's MMA thread waits for both CTAs
          else mbar_arrive(&tempty[acc]);
        }
      };

      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      // software pipeline over the chunks: the TMEM load of chunk c+1 is in flight while chunk c is converted
      // (tcgen05.ld latency was the top stall of the epilogue: 19 % of all samples in the K = 1152 projector GEMM)
      constexpr int C0 = 0, C1 = CPH;
      const int cb = half * CPH;
      uint32_t ra[32], rb[32];
      tmem_ld32(t_row + (cb + C0) * 32, ra);
#pragma unroll 1
      for (int i = C0; i < C1; i += 2) {  // two chunks per trip: the register buffers alternate statically
        tmem_ld_wait();
        if (i + 1 < C1) tmem_ld32(t_row + (cb + i + 1) * 32, rb);
        else release_acc();
        emit(ra, cb + i);
        if (i + 1 < C1) {
          tmem_ld_wait();
          if (i + 2 < C1) tmem_ld32(t_row + (cb + i + 2) * 32, ra);
          else release_acc();
          emit(rb, cb + i + 1);
        }
      }
      if constexpr (EPI == 1) {
        if (row_ok) atomicAdd(pp.cs_out + (batch / pp.inner) * pp.cs_out_stride + row, cs_acc);
      }
      if ((acc ^= 1) == 0) acc_phase ^= 1;
    }
    }
    if (lane == 0) bulk_wait_all();  // every TMA store of this warp has completed (no-op without TMA stores)
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // no CTA of the pair exits (or frees TMEM) while the other may still touch it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// m / n tile counts and the rasterisation group of a problem for a tile shape (depends on the problem only, so a
// problem that is walked over several launches -- tail fill -- sees one consistent tile numbering)
static void gemm_set_tiling(GemmTcParams& p, int bn, int cg) {
  p.m_tiles = ceil_div(p.M, GEMM_BM * cg);
  p.n_tiles = ceil_div(p.N, bn);
  // a group's A rows (group * TILE_M * K bf16) should stay in L2 (126 MB, shared with the W slabs in flight and
  // the streaming C writes) while N is swept: ~32 MB.  With the old fixed 16 x 128 rows the 12544 x 14336 x
  // 3584 K/V projection read 1.06 GB from DRAM for 0.19 GB of operands (W re-read once per group).
  const long long tile_bytes = static_cast<long long>(GEMM_BM) * cg * p.K * 2;
  long long g = (40ll << 20) / (tile_bytes > 0 ? tile_bytes : 1);
  if (tile_bytes * p.m_tiles <= (64ll << 20)) g = p.m_tiles;  // all of A fits: one group, W is read exactly once
  if (g < 2) g = 2;
  if (g > p.m_tiles) g = p.m_tiles;
  p.group_m = static_cast<int>(g);
  if (p.batches < 1) p.batches = 1;
  if (p.inner < 1) p.inner = 1;
}

// One launch.  `p` must carry its tiling and tile range; `p2` (may be null) is the tail-fill problem with its range
// and fill_* fields.  `full_grid`: launch every worker even if the primary has fewer tiles (the idle ones fill).
template <int BN, bool A_MN, bool B_MN, int CG, typename T = __nv_bfloat16, int EPI = 0>
static int launch_gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmTcParams& p,
                          cudaStream_t st, const CUtensorMap* tmA2 = nullptr, const CUtensorMap* tmB2 = nullptr,
                          const CUtensorMap* tmC2 = nullptr, const GemmTcParams* p2 = nullptr, int workers_forced = 0) {
  using Cfg = GemmCfg<BN, CG>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "gemm smem budget exceeded");
  static bool configured_dev[64] = {};  // the attribute is per device (one process may drive several GPUs)
  int dev_id = 0;
  MAVLM_CUDA_OK(cudaGetDevice(&dev_id));
  bool& configured = configured_dev[dev_id & 63];
  if (!configured) {
    MAVLM_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN, CG, T, EPI>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const long long tiles = static_cast<long long>(p.tile_end) - p.tile_begin;
  const int workers_max = sm_count() / CG;
  int workers = static_cast<int>(tiles < workers_max ? tiles : workers_max);
  if (workers_forced > 0) workers = workers_forced;
  if (workers < 1) return MAVLM_OK;
  LaunchCfg lc;
  make_launch(lc, dim3(workers * CG), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, st, CG, 1);
  const int has2 = p2 != nullptr ? 1 : 0;
  MAVLM_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, gemm_tc_kernel<BN, A_MN, B_MN, CG, T, EPI>, tmA, tmB, tmC, p, has2 ? *tmA2 : tmA,
                                   has2 ? *tmB2 : tmB, has2 ? *tmC2 : tmC, has2 ? *p2 : p, has2));
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

template <bool A_MN, bool B_MN>
static int dispatch_bn(int bn, int cg, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                       const GemmTcParams& p, cudaStream_t st) {
  if constexpr (!A_MN && !B_MN) {
    if (p.half) {  // fp16 operands: forward (K-major) GEMMs only
      if (cg == 2) {
        switch (bn) {
          case 256: return launch_gemm_tc<256, false, false, 2, __half>(tmA, tmB, tmC, p, st);
          case 192: return launch_gemm_tc<192, false, false, 2, __half>(tmA, tmB, tmC, p, st);
          default:  return launch_gemm_tc<128, false, false, 2, __half>(tmA, tmB, tmC, p, st);
        }
      }
      switch (bn) {
        case 256: return launch_gemm_tc<256, false, false, 1, __half>(tmA, tmB, tmC, p, st);
        case 192: return launch_gemm_tc<192, false, false, 1, __half>(tmA, tmB, tmC, p, st);
        case 128: return launch_gemm_tc<128, false, false, 1, __half>(tmA, tmB, tmC, p, st);
        default:  return launch_gemm_tc<64, false, false, 1, __half>(tmA, tmB, tmC, p, st);
      }
    }
    if (cg == 2) {
      switch (bn) {
        case 256: return launch_gemm_tc<256, false, false, 2>(tmA, tmB, tmC, p, st);
        case 192: return launch_gemm_tc<192, false, false, 2>(tmA, tmB, tmC, p, st);
        default:  return launch_gemm_tc<128, false, false, 2>(tmA, tmB, tmC, p, st);
      }
    }
  }
  if constexpr (!A_MN && B_MN) {
    if (p.half) {  // fp16 with an MN-major B operand ([K, N] row-major: P @ V, w @ new in the inference-only modules)
      if (cg == 2) {
        if (bn >= 192) return launch_gemm_tc<256, false, true, 2, __half>(tmA, tmB, tmC, p, st);
        return launch_gemm_tc<128, false, true, 2, __half>(tmA, tmB, tmC, p, st);
      }
      switch (bn) {
        case 256: return launch_gemm_tc<256, false, true, 1, __half>(tmA, tmB, tmC, p, st);
        case 192: return launch_gemm_tc<192, false, true, 1, __half>(tmA, tmB, tmC, p, st);
        case 128: return launch_gemm_tc<128, false, true, 1, __half>(tmA, tmB, tmC, p, st);
        default:  return launch_gemm_tc<64, false, true, 1, __half>(tmA, tmB, tmC, p, st);
      }
    }
  }
  if constexpr (A_MN) {
    MAVLM_REQUIRE(!p.half, MAVLM_E_INVALID, "fp16 gemm: a transposed A operand is a training layout; fp16 is inference-only");
  }
  if constexpr (A_MN || B_MN) {
    if (cg == 2) {  // transposed operands (dgrad / wgrad): CTA-pair tiles of 256 x 256 or 256 x 128
      if (bn >= 192) return launch_gemm_tc<256, A_MN, B_MN, 2>(tmA, tmB, tmC, p, st);
      return launch_gemm_tc<128, A_MN, B_MN, 2>(tmA, tmB, tmC, p, st);
    }
  }
  switch (bn) {
    case 256: return launch_gemm_tc<256, A_MN, B_MN, 1>(tmA, tmB, tmC, p, st);
    case 192: return launch_gemm_tc<192, A_MN, B_MN, 1>(tmA, tmB, tmC, p, st);
    case 128: return launch_gemm_tc<128, A_MN, B_MN, 1>(tmA, tmB, tmC, p, st);
    default:  return launch_gemm_tc<64, A_MN, B_MN, 1>(tmA, tmB, tmC, p, st);
  }
}

// Pick the tile: minimise (waves x tile width / tile efficiency) over single-CTA 128 x BN tiles on 148 SMs and
// CTA-pair 256 x BN tiles on 74 pairs.  Wide tiles re-use A better (smem bytes per MMA cycle drop), pair tiles
// halve the W bytes each SM stages, narrow tiles quantise better.
// Encoding of the choice (also the debug override): BN + 1000 * (CG - 1).
static int g_force_bn = 0;
int gemm_tc_pick_tile(int M, int N, int batches, bool pair_ok) {
  if (g_force_bn == -1) pair_ok = false;  // debug: heuristic restricted to single-CTA tiles
  else if (g_force_bn) return (g_force_bn >= 1000 && !pair_ok) ? g_force_bn - 1000 : g_force_bn;
  const int sms = sm_count();
  const int cand[4] = {256, 192, 128, 64};
  const float eff1[4] = {1.00f, 0.97f, 0.90f, 0.62f};
  const float eff2[4] = {1.08f, 1.04f, 0.96f, 0.f};
  int best = 128;
  float best_cost = 1e30f;
  for (int i = 0; i < 4; ++i) {
    const int nt = ceil_div(N, cand[i]) * batches;
    const float c1 = static_cast<float>(ceil_div(ceil_div(M, GEMM_BM) * nt, sms)) * cand[i] / eff1[i];
    if (c1 < best_cost - 1e-3f) {
      best_cost = c1;
      best = cand[i];
    }
    if (pair_ok && eff2[i] > 0.f) {
      const float c2 = static_cast<float>(ceil_div(ceil_div(M, 2 * GEMM_BM) * nt, sms / 2)) * cand[i] / eff2[i];
      if (c2 < best_cost - 1e-3f) {
        best_cost = c2;
        best = 1000 + cand[i];
      }
    }
  }
  return best;
}

// Operand description for the general entry: `rows` x `cols` is the STORED 2-D shape of one problem's operand
// (K-major operand: rows = M or N, cols = K; MN-major operand: rows = K, cols = M or N).
static int make_operand_map(CUtensorMap* tm, const void* base, long long ld, int rows, int cols, bool mn_major,
                            int tile_rows, int inner, int outer, long long s_inner, long long s_outer) {
  const uint64_t dims[4] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(rows), static_cast<uint64_t>(inner),
                            static_cast<uint64_t>(outer)};
  // unused batch dims keep a valid (16-byte multiple) stride
  const uint64_t si = inner > 1 ? static_cast<uint64_t>(s_inner) * 2 : static_cast<uint64_t>(ld) * 2;
  const uint64_t so = outer > 1 ? static_cast<uint64_t>(s_outer) * 2 : static_cast<uint64_t>(ld) * 2;
  const uint64_t str[3] = {static_cast<uint64_t>(ld) * 2, si, so};
  const uint32_t box[4] = {64, mn_major ? 64u : static_cast<uint32_t>(tile_rows), 1, 1};
  return make_tmap_bf16(tm, base, 4, dims, str, box);
}

// General tensor-core GEMM:  C[M,N] (+)= op(A) op(B) (+ epilogue), batched over outer x inner problems.
int gemm_tc_general(const __nv_bfloat16* A, long long lda, bool a_mn, const __nv_bfloat16* B, long long ldb, bool b_mn,
                    GemmTcParams p, int outer, int inner, const long long* s6, cudaStream_t st) {
  if (p.M == 0 || p.N == 0) return MAVLM_OK;
  MAVLM_REQUIRE(p.K > 0 && lda % 8 == 0 && ldb % 8 == 0, MAVLM_E_INVALID,
                "bf16 gemm: lda (%lld) and ldb (%lld) must be multiples of 8", lda, ldb);
  MAVLM_REQUIRE((a_mn ? p.M : p.K) % 8 == 0 && (b_mn ? p.N : p.K) % 8 == 0, MAVLM_E_INVALID,
                "bf16 gemm: the contiguous extent of each operand must be a multiple of 8 (M=%d N=%d K=%d)", p.M, p.N,
                p.K);
  const int cvec = p.out_f32 ? 4 : 8;
  MAVLM_REQUIRE(p.ldc % cvec == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0, MAVLM_E_INVALID,
                "bf16 gemm: C must be 16-byte aligned with ldc %% %d == 0", cvec);
  if (p.resid != nullptr)
    MAVLM_REQUIRE(p.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(p.resid) & 15) == 0, MAVLM_E_INVALID,
                  "bf16 gemm: resid must be 16-byte aligned with ldr %% 8 == 0");
  if (outer < 1) outer = 1;
  if (inner < 1) inner = 1;
  const long long z6[6] = {0, 0, 0, 0, 0, 0};
  if (s6 == nullptr) s6 = z6;
  p.batches = outer * inner;
  p.inner = inner;
  p.c_outer = s6[4];
  p.c_inner = s6[5];
  const bool pair_ok = p.batches == 1;
  const int choice = gemm_tc_pick_tile(p.M, p.N, p.batches, pair_ok);
  const int cg = choice >= 1000 ? 2 : 1;
  int bn = choice % 1000;
  // CTA-pair kernels with a transposed operand exist for 256 x 256 and 256 x 128 tiles only (dispatch_bn): the tiling
  // and the output boxes below must describe the tile the kernel really walks
  if (cg == 2 && (a_mn || b_mn)) bn = bn >= 192 ? 256 : 128;
  else if (cg == 2 && bn < 128) bn = 128;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_operand_map(&tmA, A, lda, a_mn ? p.K : p.M, a_mn ? p.M : p.K, a_mn, GEMM_BM, inner, outer, s6[1],
                             s6[0])))
    return rc;
  if ((rc = make_operand_map(&tmB, B, ldb, b_mn ? p.K : p.N, b_mn ? p.N : p.K, b_mn, bn / cg, inner, outer, s6[3],
                             s6[2])))
    return rc;
  // forward GEMMs (one problem, plain store) write C with TMA stores; batched / accumulating ones store directly
  CUtensorMap tmC = tmA;
  p.tma_store = (p.batches == 1 && !p.accumulate) ? 1 : 0;
  if (p.tma_store) {
    const int eb = p.out_f32 ? 4 : 2;
    const uint64_t dims[2] = {static_cast<uint64_t>(p.N), static_cast<uint64_t>(p.M)};
    const uint64_t str[1] = {static_cast<uint64_t>(p.ldc) * eb};
    const bool wide = !p.out_f32 && gemm_wide_store(bn);
    const uint32_t box[2] = {wide ? 64u : 32u, 32};
    if ((rc = make_tmap(&tmC, p.C, eb, (p.out_f32 || wide) ? 128 : 64, 2, dims, str, box))) return rc;
  }
  gemm_set_tiling(p, bn, cg);
  p.tile_begin = 0;
  p.tile_end = p.m_tiles * p.n_tiles * p.batches;
  if (a_mn)
    return b_mn ? dispatch_bn<true, true>(bn, cg, tmA, tmB, tmC, p, st)
                : dispatch_bn<true, false>(bn, cg, tmA, tmB, tmC, p, st);
  return b_mn ? dispatch_bn<false, true>(bn, cg, tmA, tmB, tmC, p, st)
              : dispatch_bn<false, false>(bn, cg, tmA, tmB, tmC, p, st);
}

int gemm_bf16_tc(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw,
                 const __nv_bfloat16* bias, const __nv_bfloat16* resid, long long ldr, const __nv_bfloat16* addvec,
                 void* C, long long ldc, int M, int N, int K, int act, int out_f32, cudaStream_t st, int half,
                 const float* pe_table, const long long* pe_idx, int pe_tokens) {
  MAVLM_REQUIRE(K > 0 && K % 8 == 0, MAVLM_E_INVALID, "bf16 gemm: K (%d) must be a multiple of 8", K);
  GemmTcParams p{};
  p.half = half;
  if (pe_table != nullptr) {
    MAVLM_REQUIRE(pe_idx != nullptr && pe_tokens > 0 && N % 4 == 0 && (reinterpret_cast<uintptr_t>(pe_table) & 15) == 0,
                  MAVLM_E_INVALID, "gemm + PE: needs frame indices, tokens per frame > 0, N %% 4 == 0, aligned table");
    p.pe_table = pe_table; p.pe_idx = pe_idx; p.pe_tokens = pe_tokens;
  }
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.resid = resid; p.ldr = ldr; p.addvec = addvec;
  p.C = C; p.ldc = ldc; p.act = act; p.out_f32 = out_f32;
  return gemm_tc_general(A, lda, false, W, ldw, false, p, 1, 1, nullptr, st);
}

// ---------------------------------------------------------------------------------------------------------------
// Tail fill: co-scheduling an off-critical-path GEMM into the idle tail of a critical-path one.
//
// The recurrence is a serial chain of GEMMs with only 1568 rows: 7 x 14 = 98 CTA-pair tiles (256 x 256) on 74 pairs
// is 1.32 waves, so 50 of the 148 tile slots of such a launch compute nothing (the 1568 x 3584 x 3584 launches ran at
// 0.55 of the rate of the large GEMMs).  The step also contains GEMM work that nothing on the chain waits for yet: the
// frame-side K/V projection of LATER chunks and the fuser MLP of FINISHED states.  A launch therefore takes TWO
// problems of the same tile shape (CTA pair, 256 x 256): every worker walks the primary's tiles round-robin; the
// workers that would idle in its last wave then take tiles [begin, end) of the fill problem -- as many per idle worker
// as fit into one primary tile's time (K_primary / K_fill) -- so the chain's latency is unchanged and the fill work
// costs nothing.  The host keeps a cursor per fill problem and finishes whatever is left with a plain range launch
// before the first consumer.  Results are bit-identical to separate launches (same tiles, same k order).
// ---------------------------------------------------------------------------------------------------------------
constexpr int FILL_BN = 256;
constexpr int FILL_CG = 2;

static int fill_prepare(const mavlm_gemm_desc& d, int half, GemmTcParams& p, CUtensorMap& tmA, CUtensorMap& tmB,
                        CUtensorMap& tmC) {
  MAVLM_REQUIRE(d.A != nullptr && d.W != nullptr && d.C != nullptr, MAVLM_E_INVALID, "gemm fill: NULL operand");
  MAVLM_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0 && d.K % 8 == 0 && d.lda % 8 == 0 && d.ldw % 8 == 0, MAVLM_E_INVALID,
                "gemm fill: bad shape / strides (M=%d N=%d K=%d)", d.M, d.N, d.K);
  MAVLM_REQUIRE(d.act >= MAVLM_ACT_NONE && d.act <= MAVLM_ACT_RELU, MAVLM_E_INVALID, "gemm fill: bad activation %d", d.act);
  const int out_f32 = d.out_dtype == MAVLM_F32 ? 1 : 0;
  const int cvec = out_f32 ? 4 : 8;
  MAVLM_REQUIRE(d.ldc % cvec == 0 && (reinterpret_cast<uintptr_t>(d.C) & 15) == 0, MAVLM_E_INVALID,
                "gemm fill: C must be 16-byte aligned with ldc %% %d == 0", cvec);
  if (d.resid != nullptr)
    MAVLM_REQUIRE(d.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(d.resid) & 15) == 0, MAVLM_E_INVALID,
                  "gemm fill: resid must be 16-byte aligned with ldr %% 8 == 0");
  p = GemmTcParams{};
  p.half = half;
  p.M = d.M; p.N = d.N; p.K = d.K;
  p.bias = d.bias; p.resid = d.resid; p.ldr = d.ldr; p.addvec = d.addvec;
  if (d.pe_table != nullptr) {
    MAVLM_REQUIRE(d.frame_idx != nullptr && d.tokens_per_frame > 0 && d.N % 4 == 0, MAVLM_E_INVALID, "gemm fill: bad PE fusion");
    p.pe_table = d.pe_table; p.pe_idx = reinterpret_cast<const long long*>(d.frame_idx); p.pe_tokens = d.tokens_per_frame;
  }
  p.C = d.C; p.ldc = d.ldc; p.act = d.act; p.out_f32 = out_f32;
  p.batches = 1; p.inner = 1;
  p.tma_store = 1;
  gemm_set_tiling(p, FILL_BN, FILL_CG);
  int rc;
  if ((rc = make_operand_map(&tmA, d.A, d.lda, d.M, d.K, false, GEMM_BM, 1, 1, 0, 0))) return rc;
  if ((rc = make_operand_map(&tmB, d.W, d.ldw, d.N, d.K, false, FILL_BN / FILL_CG, 1, 1, 0, 0))) return rc;
  const int eb = out_f32 ? 4 : 2;
  const uint64_t dims[2] = {static_cast<uint64_t>(d.N), static_cast<uint64_t>(d.M)};
  const uint64_t str[1] = {static_cast<uint64_t>(d.ldc) * eb};
  const bool wide = !out_f32 && gemm_wide_store(FILL_BN);
  const uint32_t box[2] = {wide ? 64u : 32u, 32};
  return make_tmap(&tmC, d.C, eb, (out_f32 || wide) ? 128 : 64, 2, dims, str, box);
}

int gemm_fill_num_tiles(const mavlm_gemm_desc* d) {
  return ceil_div(d->M, GEMM_BM * FILL_CG) * ceil_div(d->N, FILL_BN);
}

// tiles [t0, t1) of one problem as an ordinary launch (the flush of a fill problem, or a whole GEMM with t0 = 0)
int gemm_fill_range(const mavlm_gemm_desc* d, int t0, int t1, int half, cudaStream_t st) {
  GemmTcParams p;
  CUtensorMap tmA, tmB, tmC;
  int rc;
  if ((rc = fill_prepare(*d, half, p, tmA, tmB, tmC))) return rc;
  const int total = p.m_tiles * p.n_tiles;
  MAVLM_REQUIRE(0 <= t0 && t0 <= t1 && t1 <= total, MAVLM_E_INVALID, "gemm fill: tile range [%d, %d) outside [0, %d)", t0, t1, total);
  if (t0 == t1) return MAVLM_OK;
  p.tile_begin = t0;
  p.tile_end = t1;
  if (half) return launch_gemm_tc<FILL_BN, false, false, FILL_CG, __half>(tmA, tmB, tmC, p, st);
  return launch_gemm_tc<FILL_BN, false, false, FILL_CG>(tmA, tmB, tmC, p, st);
}

// primary (whole problem) + as many tiles of `fill` from `fill_begin` (at most up to fill_avail_end) as the primary's
// last wave leaves room for; *fill_done_end receives the first fill tile NOT done
int gemm_fill_fwd(const mavlm_gemm_desc* prim, const mavlm_gemm_desc* fill, int fill_begin, int fill_avail_end,
                  int* fill_done_end, int half, cudaStream_t st) {
  GemmTcParams p, p2;
  CUtensorMap tmA, tmB, tmC, tmA2, tmB2, tmC2;
  int rc;
  if ((rc = fill_prepare(*prim, half, p, tmA, tmB, tmC))) return rc;
  p.tile_begin = 0;
  p.tile_end = p.m_tiles * p.n_tiles;
  if (fill_done_end != nullptr) *fill_done_end = fill_begin;
  const int workers = sm_count() / FILL_CG;
  int n2 = 0, first_idle = 0, n_idle = 0;
  if (fill != nullptr && fill_avail_end > fill_begin) {
    if ((rc = fill_prepare(*fill, half, p2, tmA2, tmB2, tmC2))) return rc;
    const int total2 = p2.m_tiles * p2.n_tiles;
    MAVLM_REQUIRE(fill_begin >= 0 && fill_avail_end <= total2, MAVLM_E_INVALID, "gemm fill: range [%d, %d) outside [0, %d)",
                  fill_begin, fill_avail_end, total2);
    const int r = p.tile_end % workers;
    first_idle = r;
    n_idle = r == 0 ? 0 : workers - r;
    // fill tiles per idle worker: what fits into the time of one primary tile (same tile shape: time ~ K)
    int reps = 0;
    if (p2.K <= p.K) reps = p.K / p2.K;
    else if (p2.K * 4 <= p.K * 5) reps = 1;
    n2 = n_idle * reps;
    if (n2 > fill_avail_end - fill_begin) n2 = fill_avail_end - fill_begin;
  }
  if (n2 > 0) {
    p2.tile_begin = fill_begin;
    p2.tile_end = fill_begin + n2;
    p2.fill_first_idle = first_idle;
    p2.fill_n_idle = n_idle;
    if (fill_done_end != nullptr) *fill_done_end = fill_begin + n2;
    if (half)
      return launch_gemm_tc<FILL_BN, false, false, FILL_CG, __half>(tmA, tmB, tmC, p, st, &tmA2, &tmB2, &tmC2, &p2, workers);
    return launch_gemm_tc<FILL_BN, false, false, FILL_CG>(tmA, tmB, tmC, p, st, &tmA2, &tmB2, &tmC2, &p2, workers);
  }
  if (half) return launch_gemm_tc<FILL_BN, false, false, FILL_CG, __half>(tmA, tmB, tmC, p, st);
  return launch_gemm_tc<FILL_BN, false, false, FILL_CG>(tmA, tmB, tmC, p, st);
}

// Frame scores in the tensor-core tier (MemoryController.py:135-139): out[b][key] = sum over heads and queries of the
// NORMALISED attention probabilities, from a second pass over K with the forward's LSE -- a batched (b, h) GEMM
// S^T = K_h Q_h^T (rows = keys on the TMEM lanes, columns = queries) whose epilogue turns every accumulator into
// exp2(s * scale * log2 e - lse[q] * log2 e) and adds the row sums to out.  Half the MMA work of one attention call, at
// GEMM efficiency (N = 256 score tiles: the O accumulator that forces 64-key tiles on the attention kernel is absent).
int xattn_colsum_tc(const __nv_bfloat16* Q, long long ldq, long long qb, const __nv_bfloat16* K, long long ldk,
                    long long kb, const float* lse, float* out, int batch, int heads, int lq, int lk, int dh, float scale,
                    int half, cudaStream_t st) {
  MAVLM_REQUIRE(lq > 0 && lk > 0 && dh > 0 && dh % 8 == 0 && ldq % 8 == 0 && ldk % 8 == 0 && qb % 8 == 0 && kb % 8 == 0,
                MAVLM_E_INVALID, "xattn_colsum: head_dim and strides must be multiples of 8");
  MAVLM_REQUIRE(lse != nullptr && out != nullptr && scale > 0.f, MAVLM_E_INVALID, "xattn_colsum: needs the forward's LSE");
  MAVLM_CUDA_OK(cudaMemsetAsync(out, 0, static_cast<size_t>(batch) * lk * sizeof(float), st));
  GemmTcParams p{};
  p.half = half;
  p.M = lk; p.N = lq; p.K = dh;
  p.C = out; p.ldc = 4; p.out_f32 = 1;
  p.batches = batch * heads;
  p.inner = heads;
  p.cs_lse = lse; p.cs_out = out; p.cs_scale_log2 = scale * 1.44269504088896340736f; p.cs_out_stride = lk;
  constexpr int BN = 256;
  gemm_set_tiling(p, BN, 1);
  p.tile_begin = 0;
  p.tile_end = p.m_tiles * p.n_tiles * p.batches;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_operand_map(&tmA, K, ldk, lk, dh, false, GEMM_BM, heads, batch, dh, kb))) return rc;
  if ((rc = make_operand_map(&tmB, Q, ldq, lq, dh, false, BN, heads, batch, dh, qb))) return rc;
  if (half) return launch_gemm_tc<BN, false, false, 1, __half, 1>(tmA, tmB, tmA, p, st);
  return launch_gemm_tc<BN, false, false, 1, __nv_bfloat16, 1>(tmA, tmB, tmA, p, st);
}

// Attention backward with the elementwise passes fused into GEMM epilogues (bf16):
//   mode 2:  P[b,h] = exp(Q_h K_h^T * scale - lse)                      (bf16 out; was: fp32 scores + a pass over them)
//   mode 3:  dS[b,h] = P * (dO_h V_h^T - D) * scale, in place over P    (was: fp32 dP + a pass over P and dP)
// A [rows_a x dh] and B [rows_b x dh] are K-major column slices of [B, L, H*dh] tensors (batch stride, head offset dh);
// out [batch*heads][M][N] bf16 contiguous; vec [batch*heads][M] fp32 (LSE, natural log, or D).
int attn_bwd_scores_gemm(int mode, const __nv_bfloat16* A, long long lda, long long a_batch, const __nv_bfloat16* B,
                         long long ldb, long long b_batch, __nv_bfloat16* out, const float* vec, int batch, int heads, int M,
                         int N, int dh, float scale, cudaStream_t st) {
  MAVLM_REQUIRE(mode == 2 || mode == 3, MAVLM_E_INVALID, "attn_bwd_scores_gemm: bad mode");
  MAVLM_REQUIRE(dh % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && a_batch % 8 == 0 && b_batch % 8 == 0 && N % 8 == 0,
                MAVLM_E_INVALID, "attention backward: head_dim, key count and strides must be multiples of 8");
  GemmTcParams p{};
  p.M = M; p.N = N; p.K = dh;
  p.C = out; p.ldc = N; p.out_f32 = 0;
  p.batches = batch * heads;
  p.inner = heads;
  p.c_outer = static_cast<long long>(heads) * M * N;
  p.c_inner = static_cast<long long>(M) * N;
  p.cs_lse = vec;
  p.cs_scale_log2 = mode == 2 ? scale * 1.44269504088896340736f : scale;
  constexpr int BN = 256;
  gemm_set_tiling(p, BN, 1);
  p.tile_begin = 0;
  p.tile_end = p.m_tiles * p.n_tiles * p.batches;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_operand_map(&tmA, A, lda, M, dh, false, GEMM_BM, heads, batch, dh, a_batch))) return rc;
  if ((rc = make_operand_map(&tmB, B, ldb, N, dh, false, BN, heads, batch, dh, b_batch))) return rc;
  if (mode == 2) return launch_gemm_tc<BN, false, false, 1, __nv_bfloat16, 2>(tmA, tmB, tmA, p, st);
  return launch_gemm_tc<BN, false, false, 1, __nv_bfloat16, 3>(tmA, tmB, tmA, p, st);
}

// Backward-pass entry (mavlm_gemm_ex, bf16): trans_a = 1 -> A stored [K,M]; trans_b = 0 -> B stored [K,N].
int gemm_ex_bf16(const __nv_bfloat16* A, long long lda, int trans_a, const __nv_bfloat16* B, long long ldb, int trans_b,
                 void* C, long long ldc, int M, int N, int K, int accumulate, int out_f32, int outer, int inner,
                 const long long* s6, cudaStream_t st, int half) {
  GemmTcParams p{};
  p.half = half;
  p.M = M; p.N = N; p.K = K;
  p.C = C; p.ldc = ldc; p.act = MAVLM_ACT_NONE; p.out_f32 = out_f32; p.accumulate = accumulate;
  return gemm_tc_general(A, lda, trans_a != 0, B, ldb, trans_b == 0, p, outer, inner, s6, st);
}

void gemm_tc_force_bn(int bn) { g_force_bn = bn; }

}  // namespace mavlm
