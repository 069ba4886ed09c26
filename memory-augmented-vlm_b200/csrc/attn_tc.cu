// Fused multi-head cross-attention  O = softmax(Q K^T * scale) V  on tcgen05 / TMEM (sm_100a), bf16
// operands, fp32 scores / softmax / accumulation, probabilities never written to HBM.
// Replaces Attention.forward's matmul -> softmax -> matmul (MemoryController.py:51-54), which
// materialises probs [1,8,1568,6272].
//
// One CTA = 128 query rows of one (batch, head).  head_dim DH in {128, 448} (448 = OV-7B; the
// 0.5B model's 112 is zero-padded to 128 by the host-side weight packing).  320 threads:
//   warp 0   : TMA producer. Q tile resident in smem as DH/64 K-major slices [128 x 64]; K and V
//              stream through one ring of 16 KB "pair slots" = two adjacent [64 keys x 64 dh] slices
//              (128B swizzle); a 64-key block of K or V is ceil(DH/128) pair slots.
//   warp 1   : MMA issuer.  S[128 x 64 keys] = sum over dh slices Q_s K_s^T   (N=64, TMEM cols DH..DH+63)
//              O[128 x DH] += P[128 x 64 keys] V: V slices are MN-major B operands consumed straight
//              from their row-major layout, two slices per MMA (N=128) so the A operand (P) is re-read
//              from smem half as often (N=64 MMAs are smem-bandwidth bound: 6 KB per 32 cycles).
//              Issue order QK(j+1) before PV(j): the tensor pipe computes the next scores while the
//              softmax warps turn S(j) into P(j).
//   warps 2-9: softmax, two threads per query row (32 of the 64 key columns each; the exp2 work is
//              MUFU-bound, two warps per scheduler hide each other's latency).  tcgen05.ld S ->
//              registers (frees S immediately), online softmax in the log2 domain (ex2.approx.ftz on
//              fma(raw, scale*log2e, -m)) with LAZY rescaling of O (O in TMEM is only rescaled when a
//              row max grows by more than 2^8), P -> bf16 -> 128B-swizzled smem (double buffered).
// TMEM budget at DH=448: 448 (O) + 64 (S) = 512 columns, which is why the key block is 64.
#include "common.cuh"

namespace mavlm {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 64;
constexpr int ATT_SM_WARPS = 8;
constexpr int ATT_THREADS = 64 + 32 * ATT_SM_WARPS;  // 320
constexpr int ATT_SLICE_BYTES = ATT_BKV * 64 * 2;    // 8 KB
constexpr int ATT_SLOT_BYTES = 2 * ATT_SLICE_BYTES;  // 16 KB pair slot
constexpr int ATT_P_BYTES = ATT_BQ * ATT_BKV * 2;    // 16 KB

struct AttnTcParams {
  int lq, lk, kv_blocks;
  float scale_log2;
  __nv_bfloat16* O;
  long long ldo, o_batch;
  float* lse;
  int heads;
};

template <int DH>
struct AttnCfg {
  static constexpr int NS = DH / 64;          // 64-wide dh slices
  static constexpr int NPS = (NS + 1) / 2;    // pair slots per 64-key block
  static constexpr int Q_BYTES = ATT_BQ * DH * 2;
  static constexpr int RING = (DH == 448) ? 5 : 6;
  static constexpr int TMEM_COLS = (DH + 64 <= 256) ? 256 : 512;
  static constexpr int S_COL = DH;
  static constexpr int NBARS = 2 * RING + 7;
  static constexpr int XCHG_BYTES = 2 * 2 * ATT_BQ * 4;
  static constexpr int SMEM_BYTES = Q_BYTES + 2 * ATT_P_BYTES + RING * ATT_SLOT_BYTES + XCHG_BYTES + NBARS * 8 + 16;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DH>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, AttnTcParams p) {
  using Cfg = AttnCfg<DH>;
  constexpr int NS = Cfg::NS;
  constexpr int NPS = Cfg::NPS;
  constexpr int RING = Cfg::RING;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + Cfg::Q_BYTES;
  uint8_t* sKV = sP + 2 * ATT_P_BYTES;
  float* xchg = reinterpret_cast<float*>(sKV + RING * ATT_SLOT_BYTES);  // [2 parity][2 half][128 rows]
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xchg) + Cfg::XCHG_BYTES);
  uint64_t* kv_empty = kv_full + RING;
  uint64_t* q_full = kv_empty + RING;
  uint64_t* s_full = q_full + 1;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_full = s_free + 1;  // [2]
  uint64_t* o_done = p_full + 2;  // [2]: PV(j) commits to o_done[j & 1] so that a parity wait is never ambiguous
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ, h = blockIdx.y, b = blockIdx.z;
  const int J = p.kv_blocks;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {  // 128B-swizzle atoms need 1024-byte aligned tiles
      printf("mavlm: attention smem base not 1024-byte aligned\n");
      __trap();
    }
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    for (int s = 0; s < RING; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, ATT_SM_WARPS);
    mbar_init(&p_full[0], ATT_SM_WARPS);
    mbar_init(&p_full[1], ATT_SM_WARPS);
    mbar_init(&o_done[0], 1);
    mbar_init(&o_done[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
      for (int s = 0; s < NS; ++s) tma_load_3d(sQ + s * (ATT_BQ * 128), &tmQ, q_full, h * DH + 64 * s, q0, b);
      int stage = 0;
      uint32_t phase = 0;
      auto load_block = [&](const CUtensorMap* tm, int jj) {
        for (int ps = 0; ps < NPS; ++ps) {
          const int nsl = (NS - 2 * ps) >= 2 ? 2 : 1;
          mbar_wait(&kv_empty[stage], phase ^ 1);
          mbar_expect_tx(&kv_full[stage], nsl * ATT_SLICE_BYTES);
          for (int e = 0; e < nsl; ++e)
            tma_load_3d(sKV + stage * ATT_SLOT_BYTES + e * ATT_SLICE_BYTES, tm, &kv_full[stage],
                        h * DH + 64 * (2 * ps + e), jj * ATT_BKV, b);
          if (++stage == RING) { stage = 0; phase ^= 1; }
        }
      };
      load_block(&tmK, 0);
      for (int j = 1; j < J; ++j) {
        load_block(&tmK, j);
        load_block(&tmV, j - 1);
      }
      load_block(&tmV, J - 1);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BQ, ATT_BKV, 0, 0);
      constexpr uint32_t idesc_pv128 = umma_idesc_bf16(ATT_BQ, 128, 0, 1);  // B = two V slices, MN-major
      constexpr uint32_t idesc_pv64 = umma_idesc_bf16(ATT_BQ, 64, 0, 1);
      const uint32_t s_tmem = tmem_base + Cfg::S_COL;
      int stage = 0;
      uint32_t phase = 0;
      auto issue_pv = [&](int jj) {
        mbar_wait(&p_full[jj & 1], (jj >> 1) & 1);
        tc_fence_after();
        const uint64_t p_desc = umma_desc_kmajor(smem_u32(sP + (jj & 1) * ATT_P_BYTES));
        for (int ps = 0; ps < NPS; ++ps) {
          const bool pair = (NS - 2 * ps) >= 2;
          mbar_wait(&kv_full[stage], phase);
          tc_fence_after();
          // 64-wide dh groups are ATT_SLICE_BYTES apart (LBO); 8-key groups 1024 B apart (SBO)
          const uint64_t v_desc = umma_desc_mnmajor(smem_u32(sKV + stage * ATT_SLOT_BYTES), ATT_SLICE_BYTES);
#pragma unroll
          for (int k = 0; k < ATT_BKV / 16; ++k)  // 16 keys per MMA: +32 B in P rows, +2 k-atoms (2 KB) in V
            umma_bf16(tmem_base + ps * 128, p_desc + 2 * k, v_desc + 128 * k, pair ? idesc_pv128 : idesc_pv64,
                      (jj | k) != 0);
          umma_commit(&kv_empty[stage]);
          if (++stage == RING) { stage = 0; phase ^= 1; }
        }
        umma_commit(&o_done[jj & 1]);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      for (int j = 0; j < J; ++j) {
        if (j > 0) {
          mbar_wait(s_free, (j - 1) & 1);
          tc_fence_after();
        }
        for (int ps = 0; ps < NPS; ++ps) {
          const int nsl = (NS - 2 * ps) >= 2 ? 2 : 1;
          mbar_wait(&kv_full[stage], phase);
          tc_fence_after();
          for (int e = 0; e < nsl; ++e) {
            const int s = 2 * ps + e;
            const uint64_t q_desc = umma_desc_kmajor(smem_u32(sQ + s * (ATT_BQ * 128)));
            const uint64_t k_desc = umma_desc_kmajor(smem_u32(sKV + stage * ATT_SLOT_BYTES + e * ATT_SLICE_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(s_tmem, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, (s | k) != 0);
          }
          umma_commit(&kv_empty[stage]);
          if (++stage == RING) { stage = 0; phase ^= 1; }
        }
        umma_commit(s_full);
        if (j > 0) issue_pv(j - 1);
      }
      issue_pv(J - 1);
    }
  } else {
    const int qd = warp & 3;             // TMEM lane quadrant
    const int half = (warp - 2) >> 2;    // which 32 of the block's 64 key columns
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    constexpr int OCH = DH / 64;         // 32-column O chunks per half
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < J; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(tmem_base + lane_off + Cfg::S_COL + 32 * half, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      float s[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(r[i]);
      if (j == J - 1) {
        const int valid = p.lk - j * ATT_BKV - 32 * half;  // keys beyond lk were zero-filled by TMA
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= valid) s[i] = -INFINITY;
      }
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < 32; ++i) mx = fmaxf(mx, s[i]);
      float* xb = xchg + (j & 1) * (2 * ATT_BQ);
      xb[half * ATT_BQ + row] = mx;
      named_bar_sync(1, 32 * ATT_SM_WARPS);
      mx = fmaxf(mx, xb[(half ^ 1) * ATT_BQ + row]) * p.scale_log2;  // scale > 0: max commutes with scaling
      if (j == 0) {
        m_used = mx;
      } else {
        const bool need = mx > m_used + 8.f;
        if (__any_sync(0xffffffffu, need)) {  // identical in both warps of a quadrant (same rows, same mx)
          // PV(j-1) finished => O is quiescent until P(j) is published.  s_full(j) implies PV(j-2) and
          // older are complete, so o_done[(j-1)&1] is at most one completion behind: parity is exact.
          mbar_wait(&o_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
          const float alpha = need ? ex2_approx(m_used - mx) : 1.f;
#pragma unroll 1
          for (int c = half * OCH; c < (half + 1) * OCH; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem_base + lane_off + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tmem_base + lane_off + c * 32, o);
          }
          tmem_st_wait();
          tc_fence_before();
          if (need) {
            l *= alpha;
            m_used = mx;
          }
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        s[i] = ex2_approx(fmaf(s[i], p.scale_log2, -m_used));
        sum += s[i];
      }
      l += sum;
      uint8_t* prow = sP + (j & 1) * ATT_P_BYTES + row * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 t;
        t.x = pack_bf16x2(s[8 * c], s[8 * c + 1]);
        t.y = pack_bf16x2(s[8 * c + 2], s[8 * c + 3]);
        t.z = pack_bf16x2(s[8 * c + 4], s[8 * c + 5]);
        t.w = pack_bf16x2(s[8 * c + 6], s[8 * c + 7]);
        *reinterpret_cast<uint4*>(prow + (((4 * half + c) ^ (row & 7)) << 4)) = t;  // 128B swizzle (K-major A)
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
    }
    // row sum = the two halves' partial sums (both used the same m_used)
    named_bar_sync(1, 32 * ATT_SM_WARPS);
    xchg[half * ATT_BQ + row] = l;
    named_bar_sync(1, 32 * ATT_SM_WARPS);
    l += xchg[(half ^ 1) * ATT_BQ + row];
    mbar_wait(&o_done[(J - 1) & 1], ((J - 1) >> 1) & 1);  // PV(J-3) known complete (s_full(J-1)): exact as above
    tc_fence_after();
    const float inv = 1.f / l;
    const int q = q0 + row;
    __nv_bfloat16* orow = p.O + b * p.o_batch + static_cast<long long>(q) * p.ldo + h * DH;
#pragma unroll 1
    for (int c = half * OCH; c < (half + 1) * OCH; ++c) {
      uint32_t o[32];
      tmem_ld32(tmem_base + lane_off + c * 32, o);
      tmem_ld_wait();
      if (q < p.lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 t;
          t.x = pack_bf16x2(__uint_as_float(o[8 * g]) * inv, __uint_as_float(o[8 * g + 1]) * inv);
          t.y = pack_bf16x2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv);
          t.z = pack_bf16x2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv);
          t.w = pack_bf16x2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv);
          reinterpret_cast<uint4*>(orow + c * 32)[g] = t;
        }
      }
    }
    if (half == 0 && p.lse != nullptr && q < p.lq)
      p.lse[(static_cast<long long>(b) * p.heads + h) * p.lq + q] = (m_used + log2f(l)) * 0.69314718055994530942f;
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int DH>
static int launch_attn(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const AttnTcParams& p,
                       int batch, cudaStream_t st) {
  using Cfg = AttnCfg<DH>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "attention smem budget exceeded");
  static bool configured = false;
  if (!configured) {
    MAVLM_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
    configured = true;
  }
  dim3 grid(ceil_div(p.lq, ATT_BQ), p.heads, batch);
  attn_tc_kernel<DH><<<grid, ATT_THREADS, Cfg::SMEM_BYTES, st>>>(tmQ, tmK, tmV, p);
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

int xattn_bf16_tc(const __nv_bfloat16* Q, long long ldq, long long qb, const __nv_bfloat16* K, long long ldk,
                  long long kb, const __nv_bfloat16* V, long long ldv, long long vb, __nv_bfloat16* O, long long ldo,
                  long long ob, float* lse, int batch, int heads, int lq, int lk, int dh, float scale,
                  cudaStream_t st) {
  if (batch == 0 || lq == 0) return MAVLM_OK;
  MAVLM_REQUIRE(dh == 128 || dh == 448, MAVLM_E_INVALID,
                "bf16 xattn: head_dim %d not supported by the tcgen05 kernel (128 or 448; 112 is padded to 128 by "
                "the host packing)", dh);
  MAVLM_REQUIRE(lk > 0, MAVLM_E_INVALID, "xattn: empty key set");
  MAVLM_REQUIRE(scale > 0.f, MAVLM_E_INVALID, "xattn: scale must be positive");
  MAVLM_REQUIRE(ldo % 8 == 0 && ob % 8 == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, MAVLM_E_INVALID,
                "bf16 xattn: O must be 16-byte aligned with ldo %% 8 == 0");
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t cols = static_cast<uint64_t>(heads) * dh;
  auto mk = [&](CUtensorMap* tm, const void* base, long long ld, long long bs, int rows, uint32_t box_rows) {
    const uint64_t dims[3] = {cols, static_cast<uint64_t>(rows), static_cast<uint64_t>(batch)};
    // a batch of one may be described with any batch stride; keep it a valid multiple of 16 bytes
    const uint64_t bstride = batch > 1 ? static_cast<uint64_t>(bs) * 2 : static_cast<uint64_t>(ld) * 2 * rows;
    const uint64_t str[2] = {static_cast<uint64_t>(ld) * 2, bstride};
    const uint32_t box[3] = {64, box_rows, 1};
    return make_tmap_bf16(tm, base, 3, dims, str, box);
  };
  int rc;
  if ((rc = mk(&tmQ, Q, ldq, qb, lq, ATT_BQ))) return rc;
  if ((rc = mk(&tmK, K, ldk, kb, lk, ATT_BKV))) return rc;
  if ((rc = mk(&tmV, V, ldv, vb, lk, ATT_BKV))) return rc;
  AttnTcParams p{};
  p.lq = lq; p.lk = lk; p.kv_blocks = ceil_div(lk, ATT_BKV);
  p.scale_log2 = scale * 1.44269504088896340736f;
  p.O = O; p.ldo = ldo; p.o_batch = ob; p.lse = lse; p.heads = heads;
  return dh == 448 ? launch_attn<448>(tmQ, tmK, tmV, p, batch, st) : launch_attn<128>(tmQ, tmK, tmV, p, batch, st);
}

}  // namespace mavlm
