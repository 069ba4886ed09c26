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
//
// Scheduling: persistent and balanced ("stream-K" over key blocks), in GROUPS of CTAs that share K/V.
// One OV-7B video is only 13 q tiles x 8 heads = 104 (b,h,q-tile) items for 148 SMs: an item-per-CTA grid
// leaves 30 % of the SMs idle.  A group = `gs` CTAs, one per q tile of a q-tile group (gs = 13 here), that walk
// the SAME (batch, head, key block) sequence in lockstep-ish order, so a K/V block is pulled from HBM once and
// served to the whole group out of L2 (a per-CTA stream-K range destroys exactly this sharing: measured 40 %
// slower per key block).  Group g owns the contiguous range [g*U/G, (g+1)*U/G) of U = B*H*ngq*J
// (b, h, q-group, key block) units.  A CTA walks its group's range as segments (item, j0..j1): a segment that
// covers its whole item writes O / LSE directly.  An item cut across groups is finished INSIDE the kernel: every
// part that does not start at key block 0 is the FIRST segment of its group -- it writes unnormalised fp32
// (O, m, l) to its CTA's workspace slot and raises a flag; the part that starts at key block 0 is the LAST
// segment of its group -- it waits for the flags of the later parts (long set: they were computed first),
// combines them with its own accumulator using the usual log-sum-exp weights and writes the final rows.
// (A separate merge kernel cost 20 us per attention call at B = 1: a launch plus a second pass over all partials.)
//
// Why the tensor pipe stays near 55 % here (ncu, B = 8): with head_dim 448 the O accumulator takes 448 of the 512
// TMEM columns, so a score tile can only be 64 keys wide, and an SS-mode MMA with N = 64 runs at half rate
// (the 4 KB A-operand fetch takes 64 cycles; the N=64 GEMM tile shows the same 0.5x).  QK^T therefore costs twice
// its math time; a CTA-pair (cta_group::2) version was built and measured at parity, i.e. it is the A fetch and
// not the K/V bytes that bounds the kernel.
#include "common.cuh"

namespace mavlm {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 64;
constexpr int ATT_SM_WARPS = 8;
constexpr int ATT_THREADS = 64 + 32 * ATT_SM_WARPS;  // 320
constexpr int ATT_SLICE_BYTES = ATT_BKV * 64 * 2;    // 8 KB
constexpr int ATT_SLOT_BYTES = 2 * ATT_SLICE_BYTES;  // 16 KB pair slot
constexpr int ATT_P_BYTES = ATT_BQ * ATT_BKV * 2;    // 16 KB
constexpr int MERGE_MAX_PARTS = 14;                  // an item is never cut into more than 1 + 14 parts

struct AttnTcParams {
  int lq, lk, kv_blocks;
  float scale_log2;
  void* O;  // 16-bit storage type T of the kernel (bf16 or fp16)
  long long ldo, o_batch;
  float* lse;
  int heads, qtiles, items;
  int gs, ngq, groups;  // CTAs per group (q tiles of one q-group), q-groups per (b,h), number of groups
  long long units;      // batch * heads * ngq * kv_blocks  (group-level units)
  float* ws;            // partial slots: [grid] x { O fp32 [DH/32][32][128], m [128], l [128] }
  unsigned int* flags;  // [grid] "this CTA's partial slot is complete" (zeroed by the host before the launch)
  unsigned long long* trace;  // development: per CTA 2 roles x 64 event stamps (NULL in production)
};

// development trace: role 0 = MMA issuer, role 1 = lane 0 of the first softmax warp; (clock64 << 8) | code
#define AT_TR(role, code)                                                                                        \
  do {                                                                                                           \
    if (p.trace != nullptr && tr_n < 64) {                                                                       \
      p.trace[(blockIdx.x * 2 + (role)) * 64 + tr_n] = (static_cast<unsigned long long>(clock64()) << 8) | (code); \
      ++tr_n;                                                                                                    \
    }                                                                                                            \
  } while (0)

template <int DH>
__host__ __device__ constexpr long long attn_slot_floats() { return static_cast<long long>(ATT_BQ) * DH + 2 * ATT_BQ; }

// unit range of group g
__device__ __forceinline__ void attn_group_range(const AttnTcParams& p, int g, long long& u0, long long& u1) {
  u0 = static_cast<long long>(g) * p.units / p.groups;
  u1 = static_cast<long long>(g + 1) * p.units / p.groups;
}

template <int DH>
struct AttnCfg {
  static constexpr int NS = DH / 64;          // 64-wide dh slices
  static constexpr int NPS = (NS + 1) / 2;    // pair slots per 64-key block
  static constexpr int Q_BYTES = ATT_BQ * DH * 2;
  static constexpr int RING = (DH == 448) ? 5 : 6;
  static constexpr int TMEM_COLS = (DH + 64 <= 256) ? 256 : 512;
  static constexpr int S_COL = DH;
  static constexpr int NBARS = 2 * RING + 9;
  static constexpr int XCHG_BYTES = 2 * 2 * ATT_BQ * 4;
  static constexpr int SMEM_BYTES = Q_BYTES + 2 * ATT_P_BYTES + RING * ATT_SLOT_BYTES + XCHG_BYTES + NBARS * 8 + 16;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DH, typename T = __nv_bfloat16>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, AttnTcParams p) {
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
  uint64_t* o_done = p_full + 2;  // [2]: PV(n) commits to o_done[n & 1] so that a parity wait is never ambiguous
  uint64_t* q_free = o_done + 2;  // all QK MMAs of a segment done: Q smem may be overwritten
  uint64_t* o_free = q_free + 1;  // epilogue of a segment has read O out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int J = p.kv_blocks;
  const int grp = blockIdx.x / p.gs, grp_r = blockIdx.x % p.gs;  // group, and this CTA's q tile inside a q-group
  long long u_begin, u_end;
  attn_group_range(p, grp, u_begin, u_end);

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {  // 128B-swizzle atoms need 1024-byte aligned tiles
      printf("mavlm: attention smem base not 1024-byte aligned\n");
      __trap();
    }
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmO);
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
    mbar_init(q_free, 1);
    mbar_init(o_free, ATT_SM_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int b = 0, h = 0;
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
      int seg = 0;
      for (long long u = u_begin; u < u_end; ++seg) {
        const int item = static_cast<int>(u / J), j0 = static_cast<int>(u - static_cast<long long>(item) * J);
        const int j1 = static_cast<int>(min(static_cast<long long>(J), j0 + (u_end - u)));
        const int qt = (item % p.ngq) * p.gs + grp_r;  // may be >= qtiles: TMA zero-fills, nothing is stored
        h = (item / p.ngq) % p.heads;
        b = item / (p.ngq * p.heads);
        load_block(&tmK, j0);
        mbar_wait(q_free, (seg & 1) ^ 1);  // previous segment's QK MMAs no longer read sQ
        mbar_expect_tx(q_full, Cfg::Q_BYTES);
        for (int s = 0; s < NS; ++s)
          tma_load_3d(sQ + s * (ATT_BQ * 128), &tmQ, q_full, h * DH + 64 * s, qt * ATT_BQ, b);
        for (int j = j0 + 1; j < j1; ++j) {
          load_block(&tmK, j);
          load_block(&tmV, j - 1);
        }
        load_block(&tmV, j1 - 1);
        u += j1 - j0;
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BQ, ATT_BKV, 0, 0, Elem16<T>::kUmmaFormat);
      constexpr uint32_t idesc_pv128 = umma_idesc_bf16(ATT_BQ, 128, 0, 1, Elem16<T>::kUmmaFormat);  // B = two V slices, MN-major
      constexpr uint32_t idesc_pv64 = umma_idesc_bf16(ATT_BQ, 64, 0, 1, Elem16<T>::kUmmaFormat);
      const uint32_t s_tmem = tmem_base + Cfg::S_COL;
      int stage = 0;
      uint32_t phase = 0;
      int n = 0;  // key blocks processed by this CTA so far (all segments): indexes the P / barrier parities
      auto issue_pv = [&](int nn, bool first_of_segment) {
        mbar_wait(&p_full[nn & 1], (nn >> 1) & 1);
        tc_fence_after();
        const uint64_t p_desc = umma_desc_kmajor(smem_u32(sP + (nn & 1) * ATT_P_BYTES));
        for (int ps = 0; ps < NPS; ++ps) {
          const bool pair = (NS - 2 * ps) >= 2;
          mbar_wait(&kv_full[stage], phase);
          tc_fence_after();
          // 64-wide dh groups are ATT_SLICE_BYTES apart (LBO); 8-key groups 1024 B apart (SBO)
          const uint64_t v_desc = umma_desc_mnmajor(smem_u32(sKV + stage * ATT_SLOT_BYTES), ATT_SLICE_BYTES);
#pragma unroll
          for (int k = 0; k < ATT_BKV / 16; ++k)  // 16 keys per MMA: +32 B in P rows, +2 k-atoms (2 KB) in V
            umma_bf16(tmem_base + ps * 128, p_desc + 2 * k, v_desc + 128 * k, pair ? idesc_pv128 : idesc_pv64,
                      !first_of_segment || k != 0);
          umma_commit(&kv_empty[stage]);
          if (++stage == RING) { stage = 0; phase ^= 1; }
        }
        umma_commit(&o_done[nn & 1]);
      };
      int seg = 0;
      int tr_n = 0;
      AT_TR(0, 1);
      for (long long u = u_begin; u < u_end; ++seg) {
        const int item = static_cast<int>(u / J), j0 = static_cast<int>(u - static_cast<long long>(item) * J);
        const int j1 = static_cast<int>(min(static_cast<long long>(J), j0 + (u_end - u)));
        mbar_wait(q_full, seg & 1);
        tc_fence_after();
        AT_TR(0, 20);
        for (int j = j0; j < j1; ++j, ++n) {
          if (n > 0) {
            mbar_wait(s_free, (n - 1) & 1);
            tc_fence_after();
          }
          for (int ps = 0; ps < NPS; ++ps) {
            const int nsl = (NS - 2 * ps) >= 2 ? 2 : 1;
            mbar_wait(&kv_full[stage], phase);
            tc_fence_after();
            for (int e = 0; e < nsl; ++e) {
              const int s = 2 * ps + e;
              const uint64_t q_desc = umma_desc_kmajor(smem_u32(sQ + s * (ATT_BQ * 128)));
              const uint64_t k_desc =
                  umma_desc_kmajor(smem_u32(sKV + stage * ATT_SLOT_BYTES + e * ATT_SLICE_BYTES));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(s_tmem, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, (s | k) != 0);
            }
            umma_commit(&kv_empty[stage]);
            if (++stage == RING) { stage = 0; phase ^= 1; }
          }
          umma_commit(s_full);
          if (j == j1 - 1) umma_commit(q_free);
          if (j > j0) {
            if (j == j0 + 1) {  // first PV of the segment overwrites O: the previous epilogue must be done with it
              mbar_wait(o_free, (seg & 1) ^ 1);
              tc_fence_after();
            }
            issue_pv(n - 1, j == j0 + 1);
          }
        }
        if (j1 - j0 == 1) {
          mbar_wait(o_free, (seg & 1) ^ 1);
          tc_fence_after();
        }
        issue_pv(n - 1, j1 - j0 == 1);
        AT_TR(0, 21);
        u += j1 - j0;
      }
    }
  } else {
    const int qd = warp & 3;             // TMEM lane quadrant
    const int half = (warp - 2) >> 2;    // which 32 of the block's 64 key columns
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    constexpr int OCH = DH / 64;         // 32-column O chunks per half
    int n = 0;  // key blocks processed by this CTA so far (all segments)
    int tr_n = (warp == 2 && lane == 0) ? 0 : 64;
    AT_TR(1, 1);
    for (long long u = u_begin; u < u_end;) {
      const int item = static_cast<int>(u / J), j0 = static_cast<int>(u - static_cast<long long>(item) * J);
      const int j1 = static_cast<int>(min(static_cast<long long>(J), j0 + (u_end - u)));
      const int qt = (item % p.ngq) * p.gs + grp_r, h = (item / p.ngq) % p.heads, b = item / (p.ngq * p.heads);
      float m_used = -INFINITY, l = 0.f;
      AT_TR(1, 10);
      for (int j = j0; j < j1; ++j, ++n) {
        mbar_wait(s_full, n & 1);
        tc_fence_after();
        if (j == j0) AT_TR(1, 11);
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
        float* xb = xchg + (n & 1) * (2 * ATT_BQ);
        xb[half * ATT_BQ + row] = mx;
        named_bar_sync(1, 32 * ATT_SM_WARPS);
        mx = fmaxf(mx, xb[(half ^ 1) * ATT_BQ + row]) * p.scale_log2;  // scale > 0: max commutes with scaling
        if (j == j0) {
          m_used = mx;
        } else {
          const bool need = mx > m_used + 8.f;
          if (__any_sync(0xffffffffu, need)) {  // identical in both warps of a quadrant (same rows, same mx)
            // PV(n-1) finished => O is quiescent until P(n) is published.  s_full(n) implies PV(n-2) and
            // older are complete, so o_done[(n-1)&1] is at most one completion behind: parity is exact.
            mbar_wait(&o_done[(n - 1) & 1], ((n - 1) >> 1) & 1);
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
        uint8_t* prow = sP + (n & 1) * ATT_P_BYTES + row * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 t;
          t.x = Elem16<T>::pack2(s[8 * c], s[8 * c + 1]);
          t.y = Elem16<T>::pack2(s[8 * c + 2], s[8 * c + 3]);
          t.z = Elem16<T>::pack2(s[8 * c + 4], s[8 * c + 5]);
          t.w = Elem16<T>::pack2(s[8 * c + 6], s[8 * c + 7]);
          *reinterpret_cast<uint4*>(prow + (((4 * half + c) ^ (row & 7)) << 4)) = t;  // 128B swizzle (K-major A)
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[n & 1]);
      }
      // ---- segment epilogue.  row sum = the two halves' partial sums (both used the same m_used)
      // (exchange through the buffer the NEXT key block will not use, so a fast thread cannot overwrite it)
      AT_TR(1, 12);
      float* xl = xchg + ((n & 1) ^ 1) * (2 * ATT_BQ);
      named_bar_sync(1, 32 * ATT_SM_WARPS);
      xl[half * ATT_BQ + row] = l;
      named_bar_sync(1, 32 * ATT_SM_WARPS);
      l += xl[(half ^ 1) * ATT_BQ + row];
      mbar_wait(&o_done[(n - 1) & 1], ((n - 1) >> 1) & 1);  // PV(n-3) known complete (s_full(n-1)): exact as above
      tc_fence_after();
      AT_TR(1, 13);
      const int q = qt * ATT_BQ + row;
      if (j0 > 0) {
        // ---- later part of an item that starts in an earlier group: unnormalised fp32 O + (m, l) into this CTA's
        // workspace slot, [chunk][column][row] so that both this write and the merging read are coalesced
        float* wsb = p.ws + static_cast<long long>(blockIdx.x) * attn_slot_floats<DH>();
#pragma unroll 1
        for (int c = half * OCH; c < (half + 1) * OCH; ++c) {
          uint32_t o[32];
          tmem_ld32(tmem_base + lane_off + c * 32, o);
          tmem_ld_wait();
          float* wc = wsb + static_cast<long long>(c) * 32 * ATT_BQ + row;
#pragma unroll
          for (int i = 0; i < 32; ++i) wc[i * ATT_BQ] = __uint_as_float(o[i]);
        }
        if (half == 0) {
          wsb[static_cast<long long>(ATT_BQ) * DH + row] = m_used;
          wsb[static_cast<long long>(ATT_BQ) * DH + ATT_BQ + row] = l;
        }
        __threadfence();
        named_bar_sync(1, 32 * ATT_SM_WARPS);
        if (warp == 2 && lane == 0) st_release_gpu(p.flags + blockIdx.x, 1u);
        AT_TR(1, 14);
      } else {
        // ---- the part that starts at key block 0 owns the item's result.  If the item continues in later groups
        // (j1 < J) their parts were those groups' first segments: wait for them and fold them in.
        float w_own = 1.f;
        int np = 0;
        int g_part[MERGE_MAX_PARTS];
        float w_part[MERGE_MAX_PARTS];
        if (j1 < J) {
          const long long item_end = (static_cast<long long>(item) + 1) * J;
          float m_all = m_used;
          for (int g2 = grp + 1; g2 < p.groups && np < MERGE_MAX_PARTS; ++g2) {
            long long v0, v1;
            attn_group_range(p, g2, v0, v1);
            if (v0 >= item_end) break;
            g_part[np++] = g2;
          }
          if (warp == 2 && lane == 0)
            for (int i = 0; i < np; ++i) wait_flag_gpu(p.flags + static_cast<long long>(g_part[i]) * p.gs + grp_r);
          named_bar_sync(1, 32 * ATT_SM_WARPS);
          __threadfence();
          AT_TR(1, 15);
          for (int i = 0; i < np; ++i) {
            const float* sl = p.ws + (static_cast<long long>(g_part[i]) * p.gs + grp_r) * attn_slot_floats<DH>();
            w_part[i] = ld_cg_f32(sl + static_cast<long long>(ATT_BQ) * DH + row);  // m of that part
            m_all = fmaxf(m_all, w_part[i]);
          }
          w_own = ex2_approx(m_used - m_all);
          l *= w_own;
          for (int i = 0; i < np; ++i) {
            const float* sl = p.ws + (static_cast<long long>(g_part[i]) * p.gs + grp_r) * attn_slot_floats<DH>();
            w_part[i] = ex2_approx(w_part[i] - m_all);
            l += w_part[i] * ld_cg_f32(sl + static_cast<long long>(ATT_BQ) * DH + ATT_BQ + row);
          }
          m_used = m_all;
        }
        const float inv = 1.f / l;
        const float own_scale = w_own * inv;
        // Output rows leave through shared memory and TMA stores (which clip rows >= lq): the warp's 32 rows x 32
        // columns go to a 64B-swizzled 2 KB tile (thread = row, conflict-free 16-byte writes), double-buffered in
        // this warp's 4 KB of the two P buffers -- every P V has completed (o_done above) and no warp writes P again
        // before the next key block's named barrier, which every warp reaches after its last wait_read below.
        // Direct register stores (32 rows x 16 B per instruction) took ~14 k cycles per segment.
        uint8_t* out_stage = sP + (warp - 2) * 4096;
        // The other parts' partial O tiles (this warp's 32 rows x 32 columns of one chunk = 32 pieces of 128 bytes)
        // are prefetched with cp.async into a ring of PF 4 KB buffers per warp: a merging segment is its CTA's last one
        // (the item continues in the next group, so this CTA's unit range ends here) and every MMA has completed, so
        // the Q tile and the K/V ring are free.  Plain L2 loads (32 in flight per thread and chunk) added ~11 k cycles.
        constexpr int PF_Q = Cfg::Q_BYTES / (ATT_SM_WARPS * 4096);
        constexpr int PF_R = (RING * ATT_SLOT_BYTES) / (ATT_SM_WARPS * 4096);
        constexpr int PF = PF_Q + PF_R;
        auto pf_buf = [&](int r) -> uint8_t* {
          return r < PF_Q ? sQ + ((warp - 2) * PF_Q + r) * 4096 : sKV + ((warp - 2) * PF_R + (r - PF_Q)) * 4096;
        };
        const int pf_items = OCH * np;
        auto pf_issue = [&](int t) {
          const int c = half * OCH + t / np, k = t - (t / np) * np;
          const float* src = p.ws + (static_cast<long long>(g_part[k]) * p.gs + grp_r) * attn_slot_floats<DH>() +
                             static_cast<long long>(c) * 32 * ATT_BQ + qd * 32 + (lane & 7) * 4;
          uint8_t* dst = pf_buf(t % PF) + (lane & 7) * 16;
#pragma unroll
          for (int tt = 0; tt < 8; ++tt) {
            const int i = tt * 4 + (lane >> 3);
            cp_async_cg16(dst + i * 128, src + i * ATT_BQ);
          }
        };
        if (np > 0) {
          for (int t = 0; t < PF; ++t) {
            if (t < pf_items) pf_issue(t);
            cp_async_commit();
          }
        }
        int pf_t = 0;
#pragma unroll 1
        for (int c = half * OCH; c < (half + 1) * OCH; ++c) {
          uint32_t o[32];
          tmem_ld32(tmem_base + lane_off + c * 32, o);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(o[i]) * own_scale;
          for (int k = 0; k < np; ++k, ++pf_t) {
            cp_async_wait<PF - 1>();  // item pf_t has landed (one group is committed per item, empty ones included)
            __syncwarp();
            const float* sl = reinterpret_cast<const float*>(pf_buf(pf_t % PF)) + lane;
            const float wk = w_part[k] * inv;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaf(wk, sl[i * 32], v[i]);
            __syncwarp();
            if (pf_t + PF < pf_items) pf_issue(pf_t + PF);
            cp_async_commit();
          }
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
            tma_store_3d(buf, &tmO, h * DH + c * 32, qt * ATT_BQ + qd * 32, b);
            bulk_commit();
          }
        }
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        if (half == 0 && p.lse != nullptr && q < p.lq)
          p.lse[(static_cast<long long>(b) * p.heads + h) * p.lq + q] = (m_used + log2f(l)) * 0.69314718055994530942f;
      }
      tc_fence_before();
      __syncwarp();
      AT_TR(1, 16);
      if (lane == 0) mbar_arrive(o_free);
      u += j1 - j0;
    }
    if (lane == 0) bulk_wait_all();  // every TMA store of this warp has completed
    AT_TR(1, 2);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// Group geometry: q tiles are split into ngq q-groups of gs <= 16 tiles; as many groups as fit on the SMs.
struct AttnGeom {
  int qtiles, ngq, gs, groups;
  long long units;
};
static int g_force_groups = 0;
void attn_force_groups(int n) { g_force_groups = n; }
static AttnGeom attn_geometry(int batch, int heads, int lq, int lk) {
  AttnGeom g;
  g.qtiles = ceil_div(lq, ATT_BQ);
  g.ngq = ceil_div(g.qtiles, 16);
  g.gs = ceil_div(g.qtiles, g.ngq);
  g.units = static_cast<long long>(batch) * heads * g.ngq * ceil_div(lk, ATT_BKV);
  // CO-RESIDENCY: parts of an item cut across groups are merged by spin-waiting on flags raised by other CTAs of the
  // same grid, so every CTA must be resident at once.  One CTA fits per SM (shared memory), hence groups * gs <= SMs --
  // also for the debug override.  This holds while the kernel has the GPU's SMs to itself, which is how the library
  // is used (one stream per device; the wait is bounded and traps instead of hanging if that is ever violated).
  const long long resident = sm_count() / g.gs;
  long long groups = resident;
  if (g_force_groups > 0) groups = g_force_groups < resident ? g_force_groups : resident;
  // an item (one (b,h,q-group), J units) must not be cut into more than MERGE_MAX_PARTS parts
  const long long bh = static_cast<long long>(batch) * heads * g.ngq;
  if (groups > bh * (MERGE_MAX_PARTS - 1)) groups = bh * (MERGE_MAX_PARTS - 1);
  if (groups < 1) groups = 1;
  if (groups > g.units) groups = g.units;
  g.groups = static_cast<int>(groups);
  return g;
}

template <int DH, typename T = __nv_bfloat16>
static int launch_attn(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmO,
                       const AttnTcParams& p, cudaStream_t st) {
  using Cfg = AttnCfg<DH>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "attention smem budget exceeded");
  static bool configured_dev[64] = {};  // the attribute is per device (one process may drive several GPUs)
  int dev_id = 0;
  MAVLM_CUDA_OK(cudaGetDevice(&dev_id));
  bool& configured = configured_dev[dev_id & 63];
  if (!configured) {
    MAVLM_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel<DH, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
    configured = true;
  }
  if (p.units % p.groups != 0 || (p.units / p.groups) % p.kv_blocks != 0)  // some item is split across groups
    MAVLM_CUDA_OK(cudaMemsetAsync(p.flags, 0, static_cast<size_t>(p.groups) * p.gs * sizeof(unsigned int), st));
  LaunchCfg lc;  // (after a memset node the PDL attribute is inert: the edge is then a full dependency)
  make_launch(lc, dim3(p.groups * p.gs), dim3(ATT_THREADS), Cfg::SMEM_BYTES, st, 1, 2);
  MAVLM_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, attn_tc_kernel<DH, T>, tmQ, tmK, tmV, tmO, p));
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

// head_dim 448 can run on the CTA-pair kernel (attn_pair.cu; mavlm_debug_set_flags bit 6).  OFF by default: it is
// correct (tests/test_gpu_parity.py) but 1.5-1.9x SLOWER than this kernel -- the P tile it ships between the two CTAs
// (32 KB per 128-key block) moves at ~8 B/cycle over DSMEM stores (profiles/r2_attn_pair_trace.md)
size_t xattn_pair_workspace_bytes(int batch, int heads, int lq, int lk);
int xattn_bf16_pair(const __nv_bfloat16* Q, long long ldq, long long qb, const __nv_bfloat16* K, long long ldk, long long kb,
                    const __nv_bfloat16* V, long long ldv, long long vb, __nv_bfloat16* O, long long ldo, long long ob,
                    float* lse, int batch, int heads, int lq, int lk, float scale, void* ws, size_t ws_bytes,
                    cudaStream_t st, int half);
static bool g_use_pair = false;
void attn_use_pair_kernel(bool on) { g_use_pair = on; }
static unsigned long long* g_attn_trace = nullptr;
void attn_tc_set_trace(unsigned long long* buf) { g_attn_trace = buf; }

size_t xattn_bf16_workspace_bytes(int batch, int heads, int lq, int lk, int dh) {
  const AttnGeom g = attn_geometry(batch, heads, lq, lk);
  const long long slot = static_cast<long long>(ATT_BQ) * dh + 2 * ATT_BQ;
  const size_t ctas = static_cast<size_t>(g.groups) * g.gs;
  size_t need = ctas * slot * sizeof(float) + ctas * sizeof(unsigned int);  // one partial slot + one flag per CTA
  if (dh == 448) {
    const size_t pair = xattn_pair_workspace_bytes(batch, heads, lq, lk);
    if (pair > need) need = pair;
  }
  return need;
}

int xattn_bf16_tc(const __nv_bfloat16* Q, long long ldq, long long qb, const __nv_bfloat16* K, long long ldk,
                  long long kb, const __nv_bfloat16* V, long long ldv, long long vb, __nv_bfloat16* O, long long ldo,
                  long long ob, float* lse, int batch, int heads, int lq, int lk, int dh, float scale, void* ws,
                  size_t ws_bytes, cudaStream_t st, int half) {
  if (batch == 0 || lq == 0) return MAVLM_OK;
  if (dh == 448 && g_use_pair)
    return xattn_bf16_pair(Q, ldq, qb, K, ldk, kb, V, ldv, vb, O, ldo, ob, lse, batch, heads, lq, lk, scale, ws, ws_bytes, st,
                           half);
  MAVLM_REQUIRE(dh == 128 || dh == 448, MAVLM_E_INVALID,
                "bf16 xattn: head_dim %d not supported by the tcgen05 kernel (128 or 448; 112 is padded to 128 by "
                "the host packing)", dh);
  MAVLM_REQUIRE(lk > 0, MAVLM_E_INVALID, "xattn: empty key set");
  MAVLM_REQUIRE(scale > 0.f, MAVLM_E_INVALID, "xattn: scale must be positive");
  MAVLM_REQUIRE(ldo % 8 == 0 && ob % 8 == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, MAVLM_E_INVALID,
                "bf16 xattn: O must be 16-byte aligned with ldo %% 8 == 0");
  CUtensorMap tmQ, tmK, tmV, tmO;
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
  {  // O leaves through TMA stores of 32 x 32 tiles (64B swizzle); rows >= lq are clipped by the hardware
    const uint64_t dims[3] = {cols, static_cast<uint64_t>(lq), static_cast<uint64_t>(batch)};
    const uint64_t bstride = batch > 1 ? static_cast<uint64_t>(ob) * 2 : static_cast<uint64_t>(ldo) * 2 * lq;
    const uint64_t str[2] = {static_cast<uint64_t>(ldo) * 2, bstride};
    const uint32_t box[3] = {32, 32, 1};
    if ((rc = make_tmap(&tmO, O, 2, 64, 3, dims, str, box))) return rc;
  }
  const size_t need = xattn_bf16_workspace_bytes(batch, heads, lq, lk, dh);
  MAVLM_REQUIRE(ws != nullptr && ws_bytes >= need && (reinterpret_cast<uintptr_t>(ws) & 15) == 0, MAVLM_E_WORKSPACE,
                "bf16 xattn: 16-byte aligned workspace of %zu bytes needed, %zu given", need, ws_bytes);
  AttnTcParams p{};
  p.lq = lq; p.lk = lk; p.kv_blocks = ceil_div(lk, ATT_BKV);
  p.scale_log2 = scale * 1.44269504088896340736f;
  p.O = O; p.ldo = ldo; p.o_batch = ob; p.lse = lse; p.heads = heads;
  const AttnGeom geo = attn_geometry(batch, heads, lq, lk);
  p.qtiles = geo.qtiles; p.ngq = geo.ngq; p.gs = geo.gs; p.groups = geo.groups; p.units = geo.units;
  p.items = batch * heads * p.qtiles;
  p.ws = static_cast<float*>(ws);
  p.trace = g_attn_trace;
  p.flags = reinterpret_cast<unsigned int*>(p.ws + static_cast<long long>(p.groups) * p.gs * (static_cast<long long>(ATT_BQ) * dh + 2 * ATT_BQ));
  if (half)
    return dh == 448 ? launch_attn<448, __half>(tmQ, tmK, tmV, tmO, p, st)
                     : launch_attn<128, __half>(tmQ, tmK, tmV, tmO, p, st);
  return dh == 448 ? launch_attn<448>(tmQ, tmK, tmV, tmO, p, st) : launch_attn<128>(tmQ, tmK, tmV, tmO, p, st);
}

}  // namespace mavlm
