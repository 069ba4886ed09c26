// Fused multi-head cross-attention for head_dim 448 (OV-7B) on a CTA PAIR:  O = softmax(Q K^T * scale) V.
// Same contract as attn_tc.cu (Attention.forward, MemoryController.py:51-54; probabilities never leave the chip).
//
// Why a pair.  With head_dim 448 the O accumulator takes 448 of a CTA's 512 TMEM columns, so the single-CTA kernel can
// only hold a 64-key score tile, and an SS-mode tcgen05.mma with N = 64 is bound by its 4 KB A-operand fetch (64 cycles
// for 32 cycles of math): QK^T runs at half rate and the kernel tops out near 0.45 of the tensor peak
// (profiles/r1_attention_analysis.md).  Here the two CTAs of a (2,1,1) cluster own the SAME 128 query rows and split O
// by COLUMNS -- rank 0 holds O[:, 0:256], rank 1 O[:, 256:448] -- which leaves room for two 128-key score tiles
// (256 + 2 x 128 = 512 columns), so every QK^T MMA has N = 128 (math = operand fetch).  The QK^T work is not
// duplicated: the 128-key blocks ALTERNATE between the CTAs.  The owner of a block computes S = Q K^T for it, runs the
// softmax, keeps P (bf16) in its own TMEM -- aliased onto the score tile, as the A operand of its own P V MMAs (TS mode:
// no shared-memory A fetch) -- and ships the same P tile into the PEER's shared memory over DSMEM, where it is the SS
// A operand of the peer's P V MMAs for that block.  Each CTA therefore issues, per two key blocks: one QK^T (128 x 128
// x 448) and two P V (128 x {256|192} x 128): 3840 / 3328 tensor cycles per 256 keys against 5376 for the single-CTA
// kernel.
//
// One running softmax reference per query row is shared by the pair: the owner of block g folds its row maxima into
// the reference (lazily: only when a maximum exceeds it by more than 2^8), sends the new reference along with P, and
// both CTAs rescale their O halves (and their partial row sums) when it changed -- the online softmax is a serial chain
// over key blocks either way; here the chain hops between the two CTAs once per block (DSMEM store + remote
// mbarrier arrive), and only the row maxima are on it: the reference is published before the exponentials are taken.
// Row sums are accumulated per CTA over its own blocks and exchanged once per segment.
//
// Per CTA, 320 threads: warp 0 = TMA producer (Q resident: 7 K-major slices; K and V stream through a ring of five
// 16 KB slices [128 keys x 64 dh]), warp 1 = MMA issuer (one elected lane), warps 2-9 = softmax / correction /
// epilogue (two threads per query row, 64 of the block's 128 key columns each).
// Shared memory: Q 112 KB + incoming P 32 KB + ring 80 KB + 2.7 KB of exchange buffers and barriers = 226.7 KB.
//
// Scheduling, split-KV balancing and the in-kernel merge of items cut across groups are those of attn_tc.cu, with a
// group = `gs` PAIRS (one per q tile of a q-group) walking the same (batch, head, key block) sequence so that K / V are
// pulled from HBM once per group and served from L2.
#include "common.cuh"

namespace mavlm {

constexpr int PA_DH = 448;
constexpr int PA_NS = PA_DH / 64;                    // 7 dh slices
constexpr int PA_BQ = 128;
constexpr int PA_BK = 128;                           // keys per block
constexpr int PA_SM_WARPS = 8;
constexpr int PA_THREADS = 64 + 32 * PA_SM_WARPS;    // 320
constexpr int PA_SLICE = PA_BK * 64 * 2;             // 16 KB: [128 rows x 64 elements], 128B swizzle
constexpr int PA_RING = 5;
constexpr int PA_Q_BYTES = PA_NS * PA_SLICE;         // 112 KB
constexpr int PA_PIN_BYTES = 2 * PA_SLICE;           // 32 KB: P tile [128 x 128 keys] as two K-major halves
constexpr int PA_XCH_FLOATS = 2 * 2 * PA_BQ;         // [parity][half][row]  (also [cta][half][row] for the row sums)
constexpr int PA_NBARS = 2 * PA_RING + 2 + 2 + 2 + 2 + 1 + 1 + 1 + 1 + 4 + 1 + 1;   // 28
constexpr int PA_SMEM_BYTES = PA_Q_BYTES + PA_PIN_BYTES + PA_RING * PA_SLICE + PA_XCH_FLOATS * 4 + PA_BQ * 4 + PA_NBARS * 8 + 16;
constexpr int PA_S_COL = 256;                        // score buffers at TMEM columns 256 and 384
constexpr int PA_MERGE_MAX_PARTS = 14;
static_assert(PA_SMEM_BYTES <= 232448, "pair attention: shared-memory budget exceeded");

struct AttnPairParams {
  int lq, lk, kv_blocks;
  float scale_log2;
  void* O;
  long long ldo, o_batch;
  float* lse;
  int heads, qtiles;
  int gs, ngq, groups;   // pairs per group, q-groups per (b, h), groups
  long long units;       // batch * heads * ngq * kv_blocks
  float* ws;             // per CTA: O half fp32 [W/32][32][128], m [128], l [128]
  unsigned int* flags;   // per CTA
  unsigned long long* trace;  // development: event timestamps of the first pair (NULL in production)
};

// development trace: (globaltimer-free) clock64 stamps of role `role` (0 TMA, 1 MMA, 2 softmax) of CTAs 0 / 1
#define PA_TR(role, code)                                                                                  \
  do {                                                                                                     \
    if (p.trace != nullptr && blockIdx.x < 2 && tr_n < 200) {                                              \
      p.trace[(blockIdx.x * 3 + (role)) * 256 + tr_n] = (static_cast<unsigned long long>(clock64()) << 8) | (code); \
      ++tr_n;                                                                                              \
    }                                                                                                      \
  } while (0)

constexpr long long pa_slot_floats() { return static_cast<long long>(PA_BQ) * 256 + 2 * PA_BQ; }

__device__ __forceinline__ void pa_group_range(const AttnPairParams& p, int g, long long& u0, long long& u1) {
  u0 = static_cast<long long>(g) * p.units / p.groups;
  u1 = static_cast<long long>(g + 1) * p.units / p.groups;
}

__device__ __forceinline__ float pa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// D[tmem] (+)= A[tmem] * B[smem desc]  (TS mode: the A operand -- P, bf16 pairs in 32-bit columns -- comes from TMEM)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.commit arriving on an mbarrier of ANOTHER CTA of the cluster (shared::cluster address from mapa)
__device__ __forceinline__ void umma_commit_remote(uint32_t cluster_bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(cluster_bar_addr)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {  // data written by the peer CTA
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 0x3FF) == 0 && clock64() - t0 > MAVLM_MBAR_TIMEOUT_CYCLES) {
      printf("mavlm: cluster mbarrier timeout block %d thread %d bar@%u parity %u\n", blockIdx.x, threadIdx.x,
             smem_u32(bar), parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void st_cluster_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ring slot bookkeeping: every slot keeps its own phase (a pair of slices must not wrap, so slot RING-1 is skipped
// now and then and the slots' use counts diverge)
struct PaRing {
  int stage = 0;
  uint32_t phases = 0;  // bit s = parity of the NEXT use of slot s
  __device__ __forceinline__ uint32_t phase(int s) const { return (phases >> s) & 1u; }
  __device__ __forceinline__ void used(int s) { phases ^= 1u << s; }
  __device__ __forceinline__ int take() {  // one slice
    const int s = stage;
    stage = (stage + 1 == PA_RING) ? 0 : stage + 1;
    return s;
  }
  __device__ __forceinline__ int take_pair() {  // two adjacent slices (s, s + 1)
    if (stage == PA_RING - 1) stage = 0;
    const int s = stage;
    stage = (stage + 2 >= PA_RING) ? (stage + 2 - PA_RING) : stage + 2;
    return s;
  }
};

// The issue order of one segment of N key blocks on the CTA with cluster rank r (global block counter g0 at its start):
// blocks are owned alternately (block with global counter g belongs to rank g & 1); QK^T of an own block is issued two
// blocks ahead of its P V.  `qk(n)` / `pv(n)` are called in that order by the TMA producer and the MMA issuer alike.
template <typename FQ, typename FP>
__device__ __forceinline__ void pa_schedule(int N, long long g0, uint32_t r, FQ&& qk, FP&& pv) {
  int nq = static_cast<int>((g0 ^ r) & 1);  // first own block of the segment
  for (int n = 0; n < N; ++n) {
    while (nq < N && nq <= n + 2) {
      qk(nq);
      nq += 2;
    }
    pv(n);
  }
}

template <typename T>
__global__ void __launch_bounds__(PA_THREADS, 1)
attn_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnPairParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sPin = sQ + PA_Q_BYTES;
  uint8_t* sKV = sPin + PA_PIN_BYTES;
  float* xchg = reinterpret_cast<float*>(sKV + PA_RING * PA_SLICE);  // [2][2][128]
  float* refmsg = xchg + PA_XCH_FLOATS;                             // [128]: the owner's reference after its block
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(refmsg + PA_BQ);
  uint64_t* kv_empty = kv_full + PA_RING;
  uint64_t* q_full = kv_empty + PA_RING;
  uint64_t* q_free = q_full + 1;
  uint64_t* s_full = q_free + 1;       // [2] QK^T of an own block complete
  uint64_t* s_free = s_full + 2;       // [2] P V of the own block that used this score / P buffer complete
  uint64_t* p_own_full = s_free + 2;   // [2] softmax wrote P into TMEM
  uint64_t* o_free = p_own_full + 2;   // epilogue of a segment has read O out of TMEM
  uint64_t* pin_full = o_free + 1;     // the PEER's softmax warps wrote their block's P tile into my sPin (8 remote arrivals)
  uint64_t* ref_full = pin_full + 1;   // the PEER's softmax warps wrote their block's reference into my refmsg (8 remote arrivals)
  uint64_t* ref_ready = ref_full + 1;  // my softmax warps have followed that reference (rescaled O if it changed)
  uint64_t* psend_free = ref_ready + 1;  // the peer's P V of the block I sent last is complete (remote commit)
  uint64_t* o_done = psend_free + 1;   // [4] P V of block g complete -> o_done[g & 3]
  uint64_t* lx_full = o_done + 4;      // the peer's row sums of this segment arrived (8 remote arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lx_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t peer = rank ^ 1u;
  const int J = p.kv_blocks;
  const int pairid = blockIdx.x >> 1;
  const int grp = pairid / p.gs, grp_r = pairid % p.gs;
  long long u_begin, u_end;
  pa_group_range(p, grp, u_begin, u_end);

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("mavlm: pair attention smem base not 1024-byte aligned\n");
      __trap();
    }
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    for (int s = 0; s < PA_RING; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(q_full, 1);
    mbar_init(q_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_free[b], 1);
      mbar_init(&p_own_full[b], PA_SM_WARPS);
    }
    mbar_init(o_free, PA_SM_WARPS);
    mbar_init(pin_full, PA_SM_WARPS);
    mbar_init(ref_full, PA_SM_WARPS);
    mbar_init(ref_ready, PA_SM_WARPS);
    mbar_init(psend_free, 1);
    for (int b = 0; b < 4; ++b) mbar_init(&o_done[b], 1);
    mbar_init(lx_full, PA_SM_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();  // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  const int o_width = rank == 0 ? 256 : 192;   // O columns this CTA owns
  const int o_col0 = rank == 0 ? 0 : 256;      // ... starting at this head column

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int tr_n = 0;
      PaRing ring;
      long long g0 = 0;  // global key-block counter of this pair at the start of the segment
      int seg = 0;
      for (long long u = u_begin; u < u_end; ++seg) {
        const int item = static_cast<int>(u / J), j0 = static_cast<int>(u - static_cast<long long>(item) * J);
        const int j1 = static_cast<int>(min(static_cast<long long>(J), j0 + (u_end - u)));
        const int qt = (item % p.ngq) * p.gs + grp_r;
        const int h = (item / p.ngq) % p.heads, b = item / (p.ngq * p.heads);
        bool q_loaded = false;
        auto load_q = [&]() {
          mbar_wait(q_free, (seg & 1) ^ 1);  // the previous segment's QK^T MMAs no longer read sQ
          mbar_expect_tx(q_full, PA_Q_BYTES);
          for (int s = 0; s < PA_NS; ++s) tma_load_3d(sQ + s * PA_SLICE, &tmQ, q_full, h * PA_DH + 64 * s, qt * PA_BQ, b);
          q_loaded = true;
        };
        auto load_slice = [&](const CUtensorMap* tm, int slot, int dh_slice, int jj) {
          mbar_wait(&kv_empty[slot], ring.phase(slot) ^ 1);
          mbar_expect_tx(&kv_full[slot], PA_SLICE);
          tma_load_3d(sKV + slot * PA_SLICE, tm, &kv_full[slot], h * PA_DH + 64 * dh_slice, jj * PA_BK, b);
          ring.used(slot);
        };
        pa_schedule(
            j1 - j0, g0, rank,
            [&](int n) {  // K block of an own block: 7 slices; the first K slices go out before Q (as in attn_tc.cu)
              PA_TR(0, 1);
              for (int s = 0; s < PA_NS; ++s) {
                load_slice(&tmK, ring.take(), s, j0 + n);
                if (!q_loaded && s == 1) load_q();
              }
            },
            [&](int n) {  // V of block n, this CTA's O columns: rank 0 slices 0-3 (two pairs), rank 1 slices 4-6
              PA_TR(0, 2);
              if (!q_loaded) load_q();
              if (rank == 0) {
                for (int pr = 0; pr < 2; ++pr) {
                  const int s = ring.take_pair();
                  load_slice(&tmV, s, 2 * pr, j0 + n);
                  load_slice(&tmV, s + 1, 2 * pr + 1, j0 + n);
                }
              } else {
                const int s = ring.take_pair();
                load_slice(&tmV, s, 4, j0 + n);
                load_slice(&tmV, s + 1, 5, j0 + n);
                load_slice(&tmV, ring.take(), 6, j0 + n);
              }
            });
        g0 += j1 - j0;
        u += j1 - j0;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(PA_BQ, PA_BK, 0, 0, Elem16<T>::kUmmaFormat);
      constexpr uint32_t idesc_pv128 = umma_idesc_bf16(PA_BQ, 128, 0, 1, Elem16<T>::kUmmaFormat);
      constexpr uint32_t idesc_pv64 = umma_idesc_bf16(PA_BQ, 64, 0, 1, Elem16<T>::kUmmaFormat);
      const uint32_t psend_free_peer = mapa_u32(smem_u32(psend_free), peer);
      int tr_n = 0;
      PaRing ring;
      long long g0 = 0;
      long long own_cnt = 0, peer_cnt = 0;  // own / peer blocks whose QK^T (own) / P V (peer) have been issued so far
      long long own_pv = 0;                 // own blocks whose P V has been issued
      int seg = 0;
      for (long long u = u_begin; u < u_end; ++seg) {
        const int item = static_cast<int>(u / J), j0 = static_cast<int>(u - static_cast<long long>(item) * J);
        const int j1 = static_cast<int>(min(static_cast<long long>(J), j0 + (u_end - u)));
        const int N = j1 - j0;
        mbar_wait(q_full, seg & 1);
        tc_fence_after();
        int last_own = -1;
        for (int n = N - 1; n >= 0; --n)
          if ((((g0 + n) ^ rank) & 1) == 0) { last_own = n; break; }
        if (last_own < 0) umma_commit(q_free);  // no QK^T in this segment: sQ is free at once
        bool first_pv = true;
        pa_schedule(
            N, g0, rank,
            [&](int n) {  // S[buf] = Q K^T for own block n
              const int buf = static_cast<int>(own_cnt & 1);
              if (own_cnt >= 2) {  // the P V that read P out of this buffer (two own blocks ago) is complete
                mbar_wait(&s_free[buf], static_cast<uint32_t>(((own_cnt >> 1) - 1) & 1));
                tc_fence_after();
              }
              const uint32_t s_tmem = tmem_base + PA_S_COL + buf * 128;
              PA_TR(1, 1);
              for (int s = 0; s < PA_NS; ++s) {
                const int slot = ring.take();
                mbar_wait(&kv_full[slot], ring.phase(slot));
                ring.used(slot);
                tc_fence_after();
                const uint64_t q_desc = umma_desc_kmajor(smem_u32(sQ + s * PA_SLICE));
                const uint64_t k_desc = umma_desc_kmajor(smem_u32(sKV + slot * PA_SLICE));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(s_tmem, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, (s | k) != 0);
                umma_commit(&kv_empty[slot]);
              }
              umma_commit(&s_full[buf]);
              PA_TR(1, 2);
              if (n == last_own) umma_commit(q_free);
              ++own_cnt;
            },
            [&](int n) {  // O (+)= P V for block n
              const long long g = g0 + n;
              const bool own = ((g ^ rank) & 1) == 0;
              if (first_pv) {  // this P V overwrites O: the previous segment's epilogue must be done with it
                mbar_wait(o_free, (seg & 1) ^ 1);
                tc_fence_after();
              }
              uint32_t a_tmem = 0;
              uint64_t a_desc = 0;
              PA_TR(1, own ? 3 : 4);
              if (own) {
                const int buf = static_cast<int>(own_pv & 1);
                mbar_wait(&p_own_full[buf], static_cast<uint32_t>((own_pv >> 1) & 1));
                a_tmem = tmem_base + PA_S_COL + buf * 128;
              } else {
                mbar_wait(ref_ready, static_cast<uint32_t>(peer_cnt & 1));       // O follows the block's reference
                mbar_wait_cluster(pin_full, static_cast<uint32_t>(peer_cnt & 1));  // the P tile has landed in sPin
                a_desc = umma_desc_kmajor(smem_u32(sPin));
              }
              tc_fence_after();
              PA_TR(1, 5);
              for (int part = 0; part < 2; ++part) {  // rank 0: two slice pairs; rank 1: one pair + one slice
                const bool is_pair = rank == 0 || part == 0;
                const int slot = is_pair ? ring.take_pair() : ring.take();
                mbar_wait(&kv_full[slot], ring.phase(slot));
                ring.used(slot);
                if (is_pair) {
                  mbar_wait(&kv_full[slot + 1], ring.phase(slot + 1));
                  ring.used(slot + 1);
                }
                tc_fence_after();
                // V slice [128 keys x 64 dh] as an MN-major B operand: 8-key atoms 1 KB apart, 64-wide dh groups one
                // slice (16 KB) apart; 16 keys per MMA = 2 KB
                const uint64_t v_desc = umma_desc_mnmajor(smem_u32(sKV + slot * PA_SLICE), PA_SLICE);
                const uint32_t d_tmem = tmem_base + part * 128;
                const uint32_t idesc = is_pair ? idesc_pv128 : idesc_pv64;
#pragma unroll
                for (int k = 0; k < PA_BK / 16; ++k) {
                  const uint32_t accum = (!first_pv || k != 0) ? 1u : 0u;
                  if (own) umma_bf16_ts(d_tmem, a_tmem + 8 * k, v_desc + 128 * k, idesc, accum);
                  else umma_bf16(d_tmem, a_desc + (k >> 2) * (PA_SLICE >> 4) + 2 * (k & 3), v_desc + 128 * k, idesc, accum);
                }
                umma_commit(&kv_empty[slot]);
                if (is_pair) umma_commit(&kv_empty[slot + 1]);
              }
              umma_commit(&o_done[g & 3]);
              PA_TR(1, 6);
              if (own) {
                umma_commit(&s_free[own_pv & 1]);
                ++own_pv;
              } else {
                umma_commit_remote(psend_free_peer);  // the peer may overwrite my sPin / refmsg with its next block
                ++peer_cnt;
              }
              first_pv = false;
            });
        g0 += N;
        u += N;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue warps
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;  // which 64 of the block's 128 key columns
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const int och = o_width / 64;      // 32-column O chunks per half: 4 (rank 0) / 3 (rank 1)
    const uint32_t sPin_peer = mapa_u32(smem_u32(sPin), peer);
    const uint32_t refmsg_peer = mapa_u32(smem_u32(refmsg), peer);
    const uint32_t xchg_peer = mapa_u32(smem_u32(xchg), peer);
    const uint32_t pin_full_peer = mapa_u32(smem_u32(pin_full), peer);
    const uint32_t ref_full_peer = mapa_u32(smem_u32(ref_full), peer);
    const uint32_t lx_full_peer = mapa_u32(smem_u32(lx_full), peer);
    long long g0 = 0, own_cnt = 0, peer_cnt = 0;
    int seg = 0;
    int tr_n = (warp == 2 && lane == 0) ? 0 : 1000;
    for (long long u = u_begin; u < u_end; ++seg) {
      const int item = static_cast<int>(u / J), j0 = static_cast<int>(u - static_cast<long long>(item) * J);
      const int j1 = static_cast<int>(min(static_cast<long long>(J), j0 + (u_end - u)));
      const int N = j1 - j0;
      const int qt = (item % p.ngq) * p.gs + grp_r, h = (item / p.ngq) % p.heads, b = item / (p.ngq * p.heads);
      float m_ref = -INFINITY, l = 0.f;

      // O[row, my columns] *= alpha once every P V issued so far (blocks < g) is complete
      auto rescale_o = [&](long long g, float alpha) {
        mbar_wait(&o_done[(g - 1) & 3], static_cast<uint32_t>(((g - 1) >> 2) & 1));
        tc_fence_after();
#pragma unroll 1
        for (int c = half * och; c < (half + 1) * och; ++c) {
          uint32_t o[32];
          tmem_ld32(tmem_base + lane_off + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tmem_base + lane_off + c * 32, o);
        }
        tmem_st_wait();
        tc_fence_before();
      };

      for (int n = 0; n < N; ++n) {
        const long long g = g0 + n;
        const bool own = ((g ^ rank) & 1) == 0;
        if (own) {
          const int buf = static_cast<int>(own_cnt & 1);
          const uint32_t s_tmem = tmem_base + lane_off + PA_S_COL + buf * 128;
          PA_TR(2, 1);
          mbar_wait(&s_full[buf], static_cast<uint32_t>((own_cnt >> 1) & 1));
          tc_fence_after();
          PA_TR(2, 2);
          float s[64];
          {
            uint32_t r0[32], r1[32];
            tmem_ld32(s_tmem + 64 * half, r0);
            tmem_ld32(s_tmem + 64 * half + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              s[i] = __uint_as_float(r0[i]);
              s[32 + i] = __uint_as_float(r1[i]);
            }
          }
          if (j0 + n == J - 1) {
            const int valid = p.lk - (J - 1) * PA_BK - 64 * half;  // keys beyond lk were zero-filled by TMA
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (i >= valid) s[i] = -INFINITY;
          }
          float mx = s[0];
#pragma unroll
          for (int i = 1; i < 64; ++i) mx = fmaxf(mx, s[i]);
          float* xb = xchg + (own_cnt & 1) * (2 * PA_BQ);
          xb[half * PA_BQ + row] = mx;
          // (also: every thread has read its score columns, so the P columns may now be written over them)
          named_bar_sync(1, 32 * PA_SM_WARPS);
          mx = fmaxf(mx, xb[(half ^ 1) * PA_BQ + row]) * p.scale_log2;
          // The block's reference: fold this block's maxima into the running one (lazily), and publish it to the peer
          // AT ONCE -- its next block's softmax waits for nothing else from me, so the serial chain over key blocks is
          // max -> DSMEM hop -> max -> ..., while the exponentials, the P stores and the P V MMAs run off that chain
          float alpha = 1.f;
          bool need = false;
          if (n == 0) {
            m_ref = mx;  // first block of the segment: O is overwritten, nothing to rescale
          } else {
            need = mx > m_ref + 8.f;
            if (need) {
              alpha = pa_ex2(m_ref - mx);
              m_ref = mx;
            }
          }
          if (half == 0) st_cluster_f32(refmsg_peer + row * 4, m_ref);
          __syncwarp();
          if (lane == 0) mbar_arrive_release_cluster(ref_full_peer);
          PA_TR(2, 3);
          if (__any_sync(0xffffffffu, need)) {
            rescale_o(g, alpha);
            l *= alpha;
          }
          float sum = 0.f;
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e0 = pa_ex2(fmaf(s[2 * i], p.scale_log2, -m_ref));
            const float e1 = pa_ex2(fmaf(s[2 * i + 1], p.scale_log2, -m_ref));
            sum += e0 + e1;
            pk[i] = Elem16<T>::pack2(e0, e1);
          }
          l += sum;
          // P -> own TMEM (A operand of my P V, bf16 pairs: 64 keys = 32 columns) ...
          tmem_st32(tmem_base + lane_off + PA_S_COL + buf * 128 + 32 * half, pk);
          // ... and -> the peer's shared memory (K-major 128B-swizzled A tile, one 64-key half per softmax half)
          PA_TR(2, 4);
          if (own_cnt >= 1) mbar_wait(psend_free, static_cast<uint32_t>((own_cnt - 1) & 1));
          PA_TR(2, 5);
          {
            const uint32_t base = sPin_peer + half * PA_SLICE + row * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              st_cluster_u4(base + ((c ^ (row & 7)) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          }
          tmem_st_wait();
          tc_fence_before();
          fence_proxy_async_all();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&p_own_full[buf]);
            mbar_arrive_release_cluster(pin_full_peer);
          }
          PA_TR(2, 6);
          ++own_cnt;
        } else {
          // the peer's block: follow its reference as soon as it arrives (my MMA warp waits for the P tile itself)
          PA_TR(2, 7);
          mbar_wait_cluster(ref_full, static_cast<uint32_t>(peer_cnt & 1));
          PA_TR(2, 8);
          const float ref_new = refmsg[row];
          if (n == 0) {
            m_ref = ref_new;
          } else {
            const bool need = ref_new > m_ref;
            if (__any_sync(0xffffffffu, need)) {
              const float alpha = need ? pa_ex2(m_ref - ref_new) : 1.f;
              rescale_o(g, alpha);
              if (need) {
                l *= alpha;
                m_ref = ref_new;
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(ref_ready);
          ++peer_cnt;
        }
      }
      // ---- segment epilogue: total row sum = my two halves + the peer's two halves (same reference everywhere)
      {
        named_bar_sync(1, 32 * PA_SM_WARPS);  // nobody still reads the exchange buffer as a max buffer
        xchg[(rank * 2 + half) * PA_BQ + row] = l;
        st_cluster_f32(xchg_peer + ((rank * 2 + half) * PA_BQ + row) * 4, l);
        __syncwarp();
        if (lane == 0) mbar_arrive_release_cluster(lx_full_peer);
        named_bar_sync(1, 32 * PA_SM_WARPS);
        mbar_wait_cluster(lx_full, static_cast<uint32_t>(seg & 1));
        l = (xchg[0 * PA_BQ + row] + xchg[1 * PA_BQ + row]) + (xchg[2 * PA_BQ + row] + xchg[3 * PA_BQ + row]);
        named_bar_sync(1, 32 * PA_SM_WARPS);  // everybody has read the sums before the buffer serves the next block's maxima
      }
      const long long g_last = g0 + N - 1;
      mbar_wait(&o_done[g_last & 3], static_cast<uint32_t>((g_last >> 2) & 1));
      tc_fence_after();
      const int q = qt * PA_BQ + row;
      float m_used = m_ref;
      if (j0 > 0) {
        // later part of an item that starts in an earlier group: unnormalised fp32 O half + (m, l) into my slot
        float* wsb = p.ws + static_cast<long long>(blockIdx.x) * pa_slot_floats();
#pragma unroll 1
        for (int c = half * och; c < (half + 1) * och; ++c) {
          uint32_t o[32];
          tmem_ld32(tmem_base + lane_off + c * 32, o);
          tmem_ld_wait();
          float* wc = wsb + static_cast<long long>(c) * 32 * PA_BQ + row;
#pragma unroll
          for (int i = 0; i < 32; ++i) wc[i * PA_BQ] = __uint_as_float(o[i]);
        }
        if (half == 0) {
          wsb[static_cast<long long>(PA_BQ) * 256 + row] = m_used;
          wsb[static_cast<long long>(PA_BQ) * 256 + PA_BQ + row] = l;
        }
        __threadfence();
        named_bar_sync(1, 32 * PA_SM_WARPS);
        if (warp == 2 && lane == 0) st_release_gpu(p.flags + blockIdx.x, 1u);
      } else {
        float w_own = 1.f;
        int np = 0;
        int g_part[PA_MERGE_MAX_PARTS];
        float w_part[PA_MERGE_MAX_PARTS];
        if (j1 < J) {
          const long long item_end = (static_cast<long long>(item) + 1) * J;
          float m_all = m_used;
          for (int g2 = grp + 1; g2 < p.groups && np < PA_MERGE_MAX_PARTS; ++g2) {
            long long v0, v1;
            pa_group_range(p, g2, v0, v1);
            if (v0 >= item_end) break;
            g_part[np++] = g2;
          }
          auto slot_of = [&](int g2) { return (static_cast<long long>(g2) * p.gs + grp_r) * 2 + rank; };  // same rank there
          if (warp == 2 && lane == 0)
            for (int i = 0; i < np; ++i) wait_flag_gpu(p.flags + slot_of(g_part[i]));
          named_bar_sync(1, 32 * PA_SM_WARPS);
          __threadfence();
          for (int i = 0; i < np; ++i) {
            const float* sl = p.ws + slot_of(g_part[i]) * pa_slot_floats();
            w_part[i] = ld_cg_f32(sl + static_cast<long long>(PA_BQ) * 256 + row);
            m_all = fmaxf(m_all, w_part[i]);
          }
          w_own = pa_ex2(m_used - m_all);
          l *= w_own;
          for (int i = 0; i < np; ++i) {
            const float* sl = p.ws + slot_of(g_part[i]) * pa_slot_floats();
            w_part[i] = pa_ex2(w_part[i] - m_all);
            l += w_part[i] * ld_cg_f32(sl + static_cast<long long>(PA_BQ) * 256 + PA_BQ + row);
          }
          m_used = m_all;
        }
        const float inv = 1.f / l;
        const float own_scale = w_own * inv;
        T* orow = static_cast<T*>(p.O) + b * p.o_batch + static_cast<long long>(q) * p.ldo + h * PA_DH + o_col0;
#pragma unroll 1
        for (int c = half * och; c < (half + 1) * och; ++c) {
          uint32_t o[32];
          tmem_ld32(tmem_base + lane_off + c * 32, o);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(o[i]) * own_scale;
          for (int k = 0; k < np; ++k) {
            const float* sl = p.ws + ((static_cast<long long>(g_part[k]) * p.gs + grp_r) * 2 + rank) * pa_slot_floats() +
                              static_cast<long long>(c) * 32 * PA_BQ + row;
            const float wk = w_part[k] * inv;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaf(wk, ld_cg_f32(sl + i * PA_BQ), v[i]);
          }
          if (q < p.lq) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              uint4 t;
              t.x = Elem16<T>::pack2(v[8 * gq], v[8 * gq + 1]);
              t.y = Elem16<T>::pack2(v[8 * gq + 2], v[8 * gq + 3]);
              t.z = Elem16<T>::pack2(v[8 * gq + 4], v[8 * gq + 5]);
              t.w = Elem16<T>::pack2(v[8 * gq + 6], v[8 * gq + 7]);
              reinterpret_cast<uint4*>(orow + c * 32)[gq] = t;
            }
          }
        }
        if (rank == 0 && half == 0 && p.lse != nullptr && q < p.lq)
          p.lse[(static_cast<long long>(b) * p.heads + h) * p.lq + q] = (m_used + log2f(l)) * 0.69314718055994530942f;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      g0 += N;
      u += N;
    }
  }
  tc_fence_before();
  cluster_sync_all();  // neither CTA exits (or frees TMEM) while the other may still write into its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------- host
struct PairGeom {
  int qtiles, ngq, gs, groups;
  long long units;
};
static int g_pair_force_groups = 0;
void attn_pair_force_groups(int n) { g_pair_force_groups = n; }
static unsigned long long* g_pair_trace = nullptr;
void attn_pair_set_trace(unsigned long long* buf) { g_pair_trace = buf; }

static PairGeom pair_geometry(int batch, int heads, int lq, int lk) {
  PairGeom g;
  g.qtiles = ceil_div(lq, PA_BQ);
  const int pairs = sm_count() / 2;
  // q tiles are split into ngq q-groups of gs <= 16 tiles; pick the split that leaves the fewest pairs idle
  int best_ngq = ceil_div(g.qtiles, 16), best_idle = 1 << 30;
  for (int ngq = ceil_div(g.qtiles, 16); ngq <= g.qtiles && ngq <= 8; ++ngq) {
    const int gs = ceil_div(g.qtiles, ngq);
    const int idle = pairs % gs + (gs * ngq - g.qtiles) * (pairs / gs) / ngq;  // unused pairs + pairs on padding q tiles
    if (pairs / gs >= 1 && idle < best_idle) {
      best_idle = idle;
      best_ngq = ngq;
    }
  }
  g.ngq = best_ngq;
  g.gs = ceil_div(g.qtiles, g.ngq);
  g.units = static_cast<long long>(batch) * heads * g.ngq * ceil_div(lk, PA_BK);
  // CO-RESIDENCY (see attn_tc.cu): merged parts spin on flags of other CTAs of the grid -> groups * gs pairs <= SMs / 2
  const long long resident = pairs / g.gs;
  long long groups = resident;
  if (g_pair_force_groups > 0 && g_pair_force_groups < resident) groups = g_pair_force_groups;
  const long long bh = static_cast<long long>(batch) * heads * g.ngq;
  if (groups > bh * (PA_MERGE_MAX_PARTS - 1)) groups = bh * (PA_MERGE_MAX_PARTS - 1);
  if (groups < 1) groups = 1;
  if (groups > g.units) groups = g.units;
  g.groups = static_cast<int>(groups);
  return g;
}

size_t xattn_pair_workspace_bytes(int batch, int heads, int lq, int lk) {
  const PairGeom g = pair_geometry(batch, heads, lq, lk);
  const size_t ctas = static_cast<size_t>(g.groups) * g.gs * 2;
  return ctas * static_cast<size_t>(pa_slot_floats()) * sizeof(float) + ctas * sizeof(unsigned int);
}

template <typename T>
static int launch_pair(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const AttnPairParams& p,
                       cudaStream_t st) {
  static bool configured_dev[64] = {};
  int dev_id = 0;
  MAVLM_CUDA_OK(cudaGetDevice(&dev_id));
  bool& configured = configured_dev[dev_id & 63];
  if (!configured) {
    MAVLM_CUDA_OK(cudaFuncSetAttribute(attn_pair_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, PA_SMEM_BYTES));
    configured = true;
  }
  const int ctas = p.groups * p.gs * 2;
  if (p.units % p.groups != 0 || (p.units / p.groups) % p.kv_blocks != 0)
    MAVLM_CUDA_OK(cudaMemsetAsync(p.flags, 0, static_cast<size_t>(ctas) * sizeof(unsigned int), st));
  LaunchCfg lc;
  make_launch(lc, dim3(ctas), dim3(PA_THREADS), PA_SMEM_BYTES, st, 2, 2);
  MAVLM_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, attn_pair_kernel<T>, tmQ, tmK, tmV, p));
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

int xattn_bf16_pair(const __nv_bfloat16* Q, long long ldq, long long qb, const __nv_bfloat16* K, long long ldk, long long kb,
                    const __nv_bfloat16* V, long long ldv, long long vb, __nv_bfloat16* O, long long ldo, long long ob,
                    float* lse, int batch, int heads, int lq, int lk, float scale, void* ws, size_t ws_bytes,
                    cudaStream_t st, int half) {
  if (batch == 0 || lq == 0) return MAVLM_OK;
  MAVLM_REQUIRE(lk > 0, MAVLM_E_INVALID, "xattn: empty key set");
  MAVLM_REQUIRE(scale > 0.f, MAVLM_E_INVALID, "xattn: scale must be positive");
  MAVLM_REQUIRE(ldo % 8 == 0 && ob % 8 == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, MAVLM_E_INVALID,
                "bf16 xattn: O must be 16-byte aligned with ldo %% 8 == 0");
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t cols = static_cast<uint64_t>(heads) * PA_DH;
  auto mk = [&](CUtensorMap* tm, const void* base, long long ld, long long bs, int rows) {
    const uint64_t dims[3] = {cols, static_cast<uint64_t>(rows), static_cast<uint64_t>(batch)};
    const uint64_t bstride = batch > 1 ? static_cast<uint64_t>(bs) * 2 : static_cast<uint64_t>(ld) * 2 * rows;
    const uint64_t str[2] = {static_cast<uint64_t>(ld) * 2, bstride};
    const uint32_t box[3] = {64, 128, 1};
    return make_tmap_bf16(tm, base, 3, dims, str, box);
  };
  int rc;
  if ((rc = mk(&tmQ, Q, ldq, qb, lq))) return rc;
  if ((rc = mk(&tmK, K, ldk, kb, lk))) return rc;
  if ((rc = mk(&tmV, V, ldv, vb, lk))) return rc;
  const size_t need = xattn_pair_workspace_bytes(batch, heads, lq, lk);
  MAVLM_REQUIRE(ws != nullptr && ws_bytes >= need && (reinterpret_cast<uintptr_t>(ws) & 15) == 0, MAVLM_E_WORKSPACE,
                "bf16 xattn: 16-byte aligned workspace of %zu bytes needed, %zu given", need, ws_bytes);
  AttnPairParams p{};
  p.lq = lq; p.lk = lk; p.kv_blocks = ceil_div(lk, PA_BK);
  p.scale_log2 = scale * 1.44269504088896340736f;
  p.O = O; p.ldo = ldo; p.o_batch = ob; p.lse = lse; p.heads = heads;
  const PairGeom geo = pair_geometry(batch, heads, lq, lk);
  p.qtiles = geo.qtiles; p.ngq = geo.ngq; p.gs = geo.gs; p.groups = geo.groups; p.units = geo.units;
  p.ws = static_cast<float*>(ws);
  p.flags = reinterpret_cast<unsigned int*>(p.ws + static_cast<long long>(p.groups) * p.gs * 2 * pa_slot_floats());
  p.trace = g_pair_trace;
  return half ? launch_pair<__half>(tmQ, tmK, tmV, p, st) : launch_pair<__nv_bfloat16>(tmQ, tmK, tmV, p, st);
}

}  // namespace mavlm
