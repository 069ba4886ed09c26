// bf16 GEMM on the 5th-gen tensor cores (sm_100a):  C = act(A * W^T + bias) (+ resid) (+ addvec)
//   A [M,K] row-major (K-major operand), W [N,K] row-major (nn.Linear weight, K-major operand).
// Persistent, warp-specialised CTA (192 threads, one per SM):
//   warp 0  : TMA producer  (cp.async.bulk.tensor, 128B swizzle, STAGES-deep smem ring, mbarrier tx)
//   warp 1  : MMA issuer    (one elected lane issues tcgen05.mma cta_group::1, M=128 x N=BN x K=16;
//                            accumulators double-buffered in TMEM so the epilogue of tile i overlaps
//                            the main loop of tile i+1; tcgen05.commit releases smem stages)
//   warps 2-9: epilogue     (tcgen05.ld 32x32b -> registers -> bias / GELU(erf) / ReLU / residual ->
//                            bf16 or fp32 rows, 16-byte vector stores); two warps per TMEM lane quadrant,
//                            each taking half of the tile's columns (the erf epilogue of the K=1152
//                            projector GEMM was epilogue-bound with four warps: 55 % tensor-active)
// Tiles are rasterised in groups of 16 M-tiles, m-fastest inside a group and sweeping N, so the ~148
// tiles in flight share ~16 A row-blocks and ~9 W slabs (L2-resident working set of ~30 MB); plain
// m-fastest order re-read A once per N-slab wave (3.2 GB of DRAM reads for the 12544x14336x3584 GEMM).
#include "common.cuh"

namespace mavlm {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;                         // two warps per TMEM lane quadrant, splitting the columns
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;   // 320
constexpr int GEMM_GROUP_M = 16;                         // tile rasterisation: sweep N inside groups of 16 M-tiles

struct GemmTcParams {
  int M, N, K;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* resid;
  long long ldr;
  const __nv_bfloat16* addvec;
  void* C;
  long long ldc;
  int act;
  int out_f32;
  int m_tiles, n_tiles;
  // general / batched mode (backward pass): problem index = outer * inner + inner_idx
  int batches, inner;
  long long c_outer, c_inner;
  int accumulate;  // C += result
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 192 ? 5 : (BN >= 128 ? 6 : 8));
  static constexpr int ACC_STRIDE = (BN <= 64) ? 64 : (BN <= 128 ? 128 : 256);
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  // ring | bias[2][BN] | addv[2][BN] | barriers | tmem ptr ; +1024 for manual alignment
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 4 * BN * 4 + (2 * STAGES + 4) * 8 + 16 + 1024;
};

__device__ __forceinline__ void gemm_tile_coords(int tile, int m_tiles, int n_tiles, int& m_blk, int& n_blk) {
  const int per_group = GEMM_GROUP_M * n_tiles;
  const int g = tile / per_group, r = tile - g * per_group;
  const int first_m = g * GEMM_GROUP_M;
  const int gm = min(GEMM_GROUP_M, m_tiles - first_m);
  m_blk = first_m + r % gm;
  n_blk = r / gm;
}

// A_MN / B_MN: the operand is stored with its M (resp. N) index contiguous ([K, M] / [K, N] row-major: the
// transposed operands of dgrad / wgrad / attention backward).  Such a tile is loaded as 64x64 boxes
// [64 k-rows x 64 m] and consumed as an MN-major UMMA operand (8-k-row atoms of 1 KB, 64-wide M groups 8 KB apart).
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmTcParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  float* bias_s = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  float* addv_s = bias_s + 2 * BN;
  uint64_t* full = reinterpret_cast<uint64_t*>(addv_s + 2 * BN);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_batch = p.m_tiles * p.n_tiles;
  const int num_tiles = tiles_per_batch * p.batches;
  const int kblocks = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        const int batch = tile / tiles_per_batch;
        gemm_tile_coords(tile - batch * tiles_per_batch, p.m_tiles, p.n_tiles, m_blk, n_blk);
        const int m0 = m_blk * GEMM_BM, n0 = n_blk * BN;
        const int bi = batch % p.inner, bo = batch / p.inner;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
          uint8_t* a_dst = sA + stage * Cfg::A_BYTES;
          uint8_t* b_dst = sB + stage * Cfg::B_BYTES;
          if (!A_MN) {
            tma_load_4d(a_dst, &tmA, &full[stage], kb * GEMM_BK, m0, bi, bo);
          } else {
#pragma unroll
            for (int i = 0; i < GEMM_BM / 64; ++i)
              tma_load_4d(a_dst + i * 8192, &tmA, &full[stage], m0 + 64 * i, kb * GEMM_BK, bi, bo);
          }
          if (!B_MN) {
            tma_load_4d(b_dst, &tmB, &full[stage], kb * GEMM_BK, n0, bi, bo);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_4d(b_dst + i * 8192, &tmB, &full[stage], n0 + 64 * i, kb * GEMM_BK, bi, bo);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
          for (int k = 0; k < GEMM_BK / 16; ++k)  // K-major: +32 B inside the 128 B swizzle row; MN-major: +16 k-rows = 2 KB
            umma_bf16(d_tmem, a_desc + (A_MN ? 128 : 2) * k, b_desc + (B_MN ? 128 : 2) * k, idesc, (kb | k) != 0);
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        if ((acc ^= 1) == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;                     // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;           // which half of the tile's columns this warp converts
    const int et = (warp - 2) * 32 + lane;      // 0..255 within the epilogue group
    const int row_in_tile = q * 32 + lane;
    constexpr int CHUNKS = BN / 32, CPH = CHUNKS / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      const int batch = tile / tiles_per_batch;
      gemm_tile_coords(tile - batch * tiles_per_batch, p.m_tiles, p.n_tiles, m_blk, n_blk);
      const int m0 = m_blk * GEMM_BM, n0 = n_blk * BN;
      const long long c_off = (batch / p.inner) * p.c_outer + (batch % p.inner) * p.c_inner;
      float* bs = bias_s + acc * BN;
      float* as = addv_s + acc * BN;
      for (int j = et; j < BN; j += 32 * GEMM_EPI_WARPS) {
        const int n = n0 + j;
        bs[j] = (p.bias != nullptr && n < p.N) ? __bfloat162float(p.bias[n]) : 0.f;
        as[j] = (p.addvec != nullptr && n < p.N) ? __bfloat162float(p.addvec[n]) : 0.f;
      }
      named_bar_sync(1, 32 * GEMM_EPI_WARPS);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + row_in_tile;
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int c = half * CPH; c < (half + 1) * CPH; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::ACC_STRIDE + c * 32, r);
        tmem_ld_wait();
        const int nc = n0 + c * 32;
        if (row_ok && nc < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bs[c * 32 + j];
          if (p.act == MAVLM_ACT_GELU_ERF) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
          } else if (p.act == MAVLM_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          const bool full_chunk = nc + 32 <= p.N;
          if (p.resid != nullptr) {
            const __nv_bfloat16* rp = p.resid + row * p.ldr + nc;
            if (full_chunk) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 t = __ldg(reinterpret_cast<const uint4*>(rp) + g);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(h[e]);
                  v[g * 8 + 2 * e] += f.x;
                  v[g * 8 + 2 * e + 1] += f.y;
                }
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (nc + j < p.N) v[j] += __bfloat162float(rp[j]);
            }
          }
          if (p.addvec != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += as[c * 32 + j];
          }
          if (p.out_f32) {
            float* cp = static_cast<float*>(p.C) + c_off + row * p.ldc + nc;
            if (p.accumulate) {
              for (int j = 0; j < 32; ++j)
                if (nc + j < p.N) v[j] += cp[j];
            }
            if (full_chunk) {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                reinterpret_cast<float4*>(cp)[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
            } else {
              for (int j = 0; j < 32; ++j)
                if (nc + j < p.N) cp[j] = v[j];
            }
          } else {
            __nv_bfloat16* cp = static_cast<__nv_bfloat16*>(p.C) + c_off + row * p.ldc + nc;
            if (p.accumulate) {
              for (int j = 0; j < 32; ++j)
                if (nc + j < p.N) v[j] += __bfloat162float(cp[j]);
            }
            if (full_chunk) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint4 t;
                t.x = pack_bf16x2(v[8 * g], v[8 * g + 1]);
                t.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
                t.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
                t.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
                reinterpret_cast<uint4*>(cp)[g] = t;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (nc + j < p.N) cp[j] = __float2bfloat16(v[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if ((acc ^= 1) == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmTcParams p, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    MAVLM_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
    configured = true;
  }
  p.m_tiles = ceil_div(p.M, GEMM_BM);
  p.n_tiles = ceil_div(p.N, BN);
  if (p.batches < 1) p.batches = 1;
  if (p.inner < 1) p.inner = 1;
  const long long tiles = static_cast<long long>(p.m_tiles) * p.n_tiles * p.batches;
  const int grid = static_cast<int>(tiles < sm_count() ? tiles : sm_count());
  gemm_tc_kernel<BN, A_MN, B_MN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, p);
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

template <bool A_MN, bool B_MN>
static int dispatch_bn(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmTcParams& p, cudaStream_t st) {
  switch (bn) {
    case 256: return launch_gemm_tc<256, A_MN, B_MN>(tmA, tmB, p, st);
    case 192: return launch_gemm_tc<192, A_MN, B_MN>(tmA, tmB, p, st);
    case 128: return launch_gemm_tc<128, A_MN, B_MN>(tmA, tmB, p, st);
    default:  return launch_gemm_tc<64, A_MN, B_MN>(tmA, tmB, p, st);
  }
}

// Pick the N tile: minimise (waves x tile width / tile efficiency); wide tiles re-use A better
// (smem read B/clk per MMA cycle drops), narrow tiles quantise better on 148 SMs.
static int g_force_bn = 0;
int gemm_tc_pick_bn(int M, int N) {
  if (g_force_bn) return g_force_bn;
  const int sms = sm_count();
  const int cand[4] = {256, 192, 128, 64};
  const float eff[4] = {1.00f, 0.97f, 0.90f, 0.62f};
  int best = 128;
  float best_cost = 1e30f;
  const int mt = ceil_div(M, GEMM_BM);
  for (int i = 0; i < 4; ++i) {
    const int tiles = mt * ceil_div(N, cand[i]);
    const int waves = ceil_div(tiles, sms);
    const float cost = static_cast<float>(waves) * cand[i] / eff[i];
    if (cost < best_cost - 1e-3f) {
      best_cost = cost;
      best = cand[i];
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
  const int bn = gemm_tc_pick_bn(p.M * p.batches, p.N);
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_operand_map(&tmA, A, lda, a_mn ? p.K : p.M, a_mn ? p.M : p.K, a_mn, GEMM_BM, inner, outer, s6[1],
                             s6[0])))
    return rc;
  if ((rc = make_operand_map(&tmB, B, ldb, b_mn ? p.K : p.N, b_mn ? p.N : p.K, b_mn, bn, inner, outer, s6[3], s6[2])))
    return rc;
  if (a_mn) return b_mn ? dispatch_bn<true, true>(bn, tmA, tmB, p, st) : dispatch_bn<true, false>(bn, tmA, tmB, p, st);
  return b_mn ? dispatch_bn<false, true>(bn, tmA, tmB, p, st) : dispatch_bn<false, false>(bn, tmA, tmB, p, st);
}

int gemm_bf16_tc(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw,
                 const __nv_bfloat16* bias, const __nv_bfloat16* resid, long long ldr, const __nv_bfloat16* addvec,
                 void* C, long long ldc, int M, int N, int K, int act, int out_f32, cudaStream_t st) {
  MAVLM_REQUIRE(K > 0 && K % 8 == 0, MAVLM_E_INVALID, "bf16 gemm: K (%d) must be a multiple of 8", K);
  GemmTcParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.resid = resid; p.ldr = ldr; p.addvec = addvec;
  p.C = C; p.ldc = ldc; p.act = act; p.out_f32 = out_f32;
  return gemm_tc_general(A, lda, false, W, ldw, false, p, 1, 1, nullptr, st);
}

// Backward-pass entry (mavlm_gemm_ex, bf16): trans_a = 1 -> A stored [K,M]; trans_b = 0 -> B stored [K,N].
int gemm_ex_bf16(const __nv_bfloat16* A, long long lda, int trans_a, const __nv_bfloat16* B, long long ldb, int trans_b,
                 void* C, long long ldc, int M, int N, int K, int accumulate, int out_f32, int outer, int inner,
                 const long long* s6, cudaStream_t st) {
  GemmTcParams p{};
  p.M = M; p.N = N; p.K = K;
  p.C = C; p.ldc = ldc; p.act = MAVLM_ACT_NONE; p.out_f32 = out_f32; p.accumulate = accumulate;
  return gemm_tc_general(A, lda, trans_a != 0, B, ldb, trans_b == 0, p, outer, inner, s6, st);
}

void gemm_tc_force_bn(int bn) { g_force_bn = bn; }

}  // namespace mavlm
