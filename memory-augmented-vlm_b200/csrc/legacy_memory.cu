// the Flash-VStream-style memories and the scene segmentation that sit beside the recurrent
// memory in the reference tree (memory_module/compress_functions.py, memory_builder.py, segment.py).
//
// The reference runs these as Python loops over frames with a host decision (argmax -> slicing / torch.cat) and
// 10-20 small ATen kernels per streamed frame.  Here every streamed frame is ONE kernel launch and no host
// round trip: the CTAs first stream the frames involved (cosine partial sums, and the pairwise average a merge
// needs, written to a scratch row), then the last CTA to finish (atomic ticket) folds the partial sums in a fixed
// order, applies the reference's decision rule on a small state block in device memory and plans the rows the
// next launch has to read.  Frames never move: the state holds row handles (input row or scratch row), not data.
// Everything is HBM-bound: per streamed frame the algorithmic traffic is 2 rows (drop / merge), T0 + 1 rows
// (k_drop) or 2 T0 + 3 rows (k_merge) of P*D elements.
//
// Also here: adjacent-frame cosine + depth scores for the scene segmentation, frame means, the avg_pool2d spatial
// compression, the k-means distance / centroid kernels and the row softmax of the Turing-memory update (whose
// contractions run on the tcgen05 GEMM through mavlm_gemm_ex).
#include <algorithm>
#include <type_traits>

#include "vec.cuh"

namespace mavlm {

constexpr int LM_THREADS = 256;
constexpr int LM_CAP = 64;                 // most frames a streaming memory may keep
constexpr int LM_MAX_JOBS = 2048;          // >= LM_CAP * (LM_CAP - 1) / 2 + slack (the initial all-pairs pass)
constexpr int LM_MAX_SPLITS = 64;
constexpr float LM_NEG = -100.0f;

enum { LM_DROP = 0, LM_MERGE = 1, LM_KDROP = 2, LM_KMERGE = 3 };
enum { LM_PHASE_INIT = 0, LM_PHASE_FRAME = 1, LM_PHASE_FLUSH = 2 };

template <typename T>
__device__ __forceinline__ float round_through(float v);
template <>
__device__ __forceinline__ float round_through<float>(float v) { return v; }
template <>
__device__ __forceinline__ float round_through<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16(v)); }
template <>
__device__ __forceinline__ float round_through<__half>(float v) { return __half2float(__float2half(v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of up to 3 values; result valid in thread 0
template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float* red /* [N][8] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    v[i] = warp_sum(v[i]);
    if (lane == 0) red[i * 8 + warp] = v[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float s = 0.f;
      for (int w = 0; w < LM_THREADS / 32; ++w) s += red[i * 8 + w];
      v[i] = s;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// streaming compression (drop / merge / k_drop / k_merge)
// ---------------------------------------------------------------------------------------------------------------
// Small part of the state: the deciding CTA works on a shared-memory copy (one parallel round trip in, one out)
// instead of chasing dependent global loads from a single thread.
struct LmSmall {
  int mode, keep, n_in;
  int n_fix, n_new;                       // jobs [0, n_fix): refresh similarities of the last decision; then n_new
                                          // jobs pairing the kept rows with the incoming frame
  int avg_a, avg_b, avg_dst;              // planned average (row handles), avg_dst < 0: none
  int fix_pos;
  int free_pos;
  int n_free;
  unsigned ticket;
  int fix_target[2 * LM_CAP + 4];
  // drop / merge: handles and adjacent similarities in temporal order
  int order[LM_CAP + 2];
  float sim[LM_CAP + 2];
  // k_drop / k_merge: LM_CAP + 1 physical positions and a logical (temporal) order over them
  int lorder[LM_CAP + 2];
  int handle[LM_CAP + 2];
  int free_slots[LM_CAP + 4];
};
constexpr int LM_SMALL_WORDS = sizeof(LmSmall) / 4;
struct LmState {
  LmSmall sm;
  float S[(LM_CAP + 1) * (LM_CAP + 1)];   // all-pairs similarities by physical position (k modes)
  int jobs[LM_MAX_JOBS][2];
};

template <typename T>
__device__ __forceinline__ const T* lm_row(const T* x, const T* scratch, int n_in, long long L, int h) {
  return h < n_in ? x + static_cast<long long>(h) * L : scratch + static_cast<long long>(h - n_in) * L;
}

// (a + b) / 2 with the reference's two roundings in the storage dtype
template <typename T>
__device__ __forceinline__ float lm_avg(float a, float b) { return round_through<T>(round_through<T>(a + b) * 0.5f); }

__device__ __forceinline__ float lm_cos(float ab, float aa, float bb, float eps) {
  return ab / (fmaxf(sqrtf(aa), eps) * fmaxf(sqrtf(bb), eps));
}

__device__ void lm_plan_new(LmSmall* m, int (*jobs)[2], int next) {
  const int base = m->n_fix;
  if (next >= m->n_in) { m->n_new = 0; return; }
  if (m->mode == LM_DROP || m->mode == LM_MERGE) {
    jobs[base][0] = m->order[m->keep - 1];
    jobs[base][1] = next;
    m->n_new = 1;
  } else {
    for (int j = 0; j < m->keep; ++j) {
      jobs[base + j][0] = m->handle[m->lorder[j]];
      jobs[base + j][1] = next;
    }
    m->n_new = m->keep;
  }
}

// first maximum of S over the logical (row-major) order of the keep + 1 live positions
__device__ void lm_argmax_pairs(const LmSmall* m, const float* S, int n, int* out_left, int* out_right, float* sv, int* si) {
  float best = -3.0e38f;
  int bi = 0x7fffffff;
  const int stride = LM_CAP + 1;
  for (int f = threadIdx.x; f < n * n; f += LM_THREADS) {
    const int l = f / n, r = f - l * n;
    const float v = S[m->lorder[l] * stride + m->lorder[r]];
    if (v > best) { best = v; bi = f; }      // f ascending per thread: keeps the first maximum
  }
  sv[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = LM_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float v = sv[threadIdx.x + o];
      const int i = si[threadIdx.x + o];
      if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && i < si[threadIdx.x])) { sv[threadIdx.x] = v; si[threadIdx.x] = i; }
    }
    __syncthreads();
  }
  *out_left = si[0] / n;
  *out_right = si[0] - (si[0] / n) * n;
  __syncthreads();
}

// m, S: shared-memory copies of the state; jobs: the job list in global memory (write only here)
template <typename T>
__device__ void lm_decide(LmSmall* m, float* S, int (*jobs)[2], const float* dots /* [jobs][3] */, int phase, int step,
                          int coin, int* decisions, float* sv, int* si) {
  const int keep = m->keep, mode = m->mode;
  const bool adjacent = mode == LM_DROP || mode == LM_MERGE;
  const float eps = adjacent ? 1e-8f : 1e-12f;
  const int stride = LM_CAP + 1;
  auto cosj = [&](int j) { return round_through<T>(lm_cos(dots[3 * j], dots[3 * j + 1], dots[3 * j + 2], eps)); };

  if (phase == LM_PHASE_INIT) {
    if (threadIdx.x == 0) {
      if (adjacent) {
        for (int j = 0; j + 1 < keep; ++j) m->sim[j] = cosj(j);
      } else {
        int j = 0;
        for (int a = 0; a < keep; ++a)
          for (int b = a + 1; b < keep; ++b, ++j) S[a * stride + b] = S[b * stride + a] = cosj(j);
      }
      m->n_fix = 0;
      lm_plan_new(m, jobs, keep);
    }
    return;
  }
  // similarities left over from the previous decision
  if (threadIdx.x == 0) {
    for (int j = 0; j < m->n_fix; ++j) {
      const int t = m->fix_target[j];
      if (t < 0) continue;
      if (adjacent) m->sim[t] = cosj(j);
      else S[m->fix_pos * stride + t] = S[t * stride + m->fix_pos] = cosj(j);
    }
    m->avg_dst = -1;
  }
  __syncthreads();
  if (phase == LM_PHASE_FLUSH || m->n_new == 0) {
    __syncthreads();
    if (threadIdx.x == 0) { m->n_fix = 0; m->n_new = 0; }
    return;
  }
  const int nf = m->n_fix;   // new-frame jobs start here
  const int d = step - keep; // decision slot
  if (adjacent) {
    if (threadIdx.x != 0) return;
    // temporal list of keep + 1 rows with keep adjacent similarities
    m->order[keep] = step;
    m->sim[keep - 1] = cosj(nf);
    int idx = 0;
    for (int j = 1; j < keep; ++j)
      if (m->sim[j] > m->sim[idx]) idx = j;
    if (mode == LM_DROP) {
      if (coin) ++idx;
      decisions[2 * d] = idx;
      decisions[2 * d + 1] = idx;
      m->n_fix = 0;
      if (idx < keep) {
        for (int j = idx; j < keep; ++j) m->order[j] = m->order[j + 1];
        if (idx == 0) {
          for (int j = 0; j + 1 < keep; ++j) m->sim[j] = m->sim[j + 1];
        } else {
          for (int j = idx; j + 1 < keep; ++j) m->sim[j] = m->sim[j + 1];
          jobs[0][0] = m->order[idx - 1];
          jobs[0][1] = m->order[idx];
          m->fix_target[0] = idx - 1;
          m->n_fix = 1;
        }
      }
    } else {
      decisions[2 * d] = idx;
      decisions[2 * d + 1] = idx + 1;
      const int ha = m->order[idx], hb = m->order[idx + 1];
      const int dst = m->free_slots[--m->n_free];
      if (ha >= m->n_in) m->free_slots[m->n_free++] = ha;
      if (hb >= m->n_in) m->free_slots[m->n_free++] = hb;
      m->avg_a = ha; m->avg_b = hb; m->avg_dst = dst;
      m->order[idx + 1] = dst;
      for (int j = idx; j < keep; ++j) m->order[j] = m->order[j + 1];
      for (int j = idx; j + 1 < keep; ++j) m->sim[j] = m->sim[j + 1];
      int n = 0;
      jobs[n][0] = dst; jobs[n][1] = dst; m->fix_target[n++] = -1;
      if (idx > 0) { jobs[n][0] = dst; jobs[n][1] = m->order[idx - 1]; m->fix_target[n++] = idx - 1; }
      if (idx + 1 < keep) { jobs[n][0] = dst; jobs[n][1] = m->order[idx + 1]; m->fix_target[n++] = idx; }
      m->n_fix = n;
    }
    lm_plan_new(m, jobs, step + 1);
    return;
  }
  // all-pairs modes: the new frame takes the free position at the end of the temporal order
  const int pf = m->free_pos;
  if (threadIdx.x == 0) {
    m->handle[pf] = step;
    m->lorder[keep] = pf;
    S[pf * stride + pf] = LM_NEG;
  }
  for (int j = threadIdx.x; j < keep; j += LM_THREADS) {
    const int q = m->lorder[j];
    S[q * stride + pf] = S[pf * stride + q] = cosj(nf + j);
  }
  __syncthreads();
  int left, right;
  lm_argmax_pairs(m, S, keep + 1, &left, &right, sv, si);
  if (threadIdx.x != 0) return;
  decisions[2 * d] = left;
  decisions[2 * d + 1] = right;
  if (mode == LM_KDROP) {
    const int idx = coin ? left : right;
    m->free_pos = m->lorder[idx];
    for (int j = idx; j < keep; ++j) m->lorder[j] = m->lorder[j + 1];
    m->n_fix = 0;
  } else {
    const int pl = m->lorder[left], pr = m->lorder[right];
    const int ha = m->handle[pl], hb = m->handle[pr];
    const int dst = m->free_slots[--m->n_free];
    if (ha >= m->n_in) m->free_slots[m->n_free++] = ha;
    if (hb >= m->n_in) m->free_slots[m->n_free++] = hb;
    m->avg_a = ha; m->avg_b = hb; m->avg_dst = dst;
    m->handle[pr] = dst;
    m->free_pos = pl;
    for (int j = left; j < keep; ++j) m->lorder[j] = m->lorder[j + 1];
    int n = 0;
    jobs[n][0] = dst; jobs[n][1] = dst; m->fix_target[n++] = -1;
    for (int j = 0; j < keep; ++j) {
      const int q = m->lorder[j];
      if (q == pr) continue;
      jobs[n][0] = dst; jobs[n][1] = m->handle[q]; m->fix_target[n++] = q;
    }
    m->fix_pos = pr;
    m->n_fix = n;
  }
  lm_plan_new(m, jobs, step + 1);
}

// Videos of a batch are independent chains: grid.z = video, each with its own state / partial sums / scratch rows
// (one workspace block per video), coins and decision log.  A video shorter than the longest one simply has no
// new-frame jobs left (n_new = 0) and its CTAs fall through.
struct LmBatch {
  const void* x_single;        // batch == 1: the video itself
  const void* const* xs;       // batch > 1: DEVICE array of per-video base pointers
  char* ws;                    // workspace, `per_video` bytes per video
  size_t per_video, partial_off, scratch_off;
  const uint8_t* coins;        // [batch, coin_stride] or NULL
  int* decisions;              // [batch, coin_stride, 2]
  long long coin_stride;
};

// 4 CTAs per SM (<= 64 registers): the 512 CTAs of a drop launch must be co-resident (one wave)
template <typename T>
__global__ void __launch_bounds__(LM_THREADS, 4) lm_stream_kernel(LmBatch b, long long L, int phase, int step) {
  constexpr int V = Vec<T>::N;
  const int vid = blockIdx.z;
  const T* __restrict__ x = static_cast<const T*>(b.xs != nullptr ? b.xs[vid] : b.x_single);
  char* wsv = b.ws + static_cast<size_t>(vid) * b.per_video;
  LmState* s = reinterpret_cast<LmState*>(wsv);
  float* partial = reinterpret_cast<float*>(wsv + b.partial_off);
  T* scratch = reinterpret_cast<T*>(wsv + b.scratch_off);
  const uint8_t* coins = b.coins != nullptr ? b.coins + vid * b.coin_stride : nullptr;
  int* decisions = b.decisions + vid * b.coin_stride * 2;
  extern __shared__ __align__(16) unsigned char lm_smem[];
  // the deciding CTA's workspace: dots [jobs][3] | state copy | argmax scratch | (k modes) similarity matrix
  float* s_dots = reinterpret_cast<float*>(lm_smem);
  __shared__ float s_red[24];
  __shared__ bool s_last;
  pdl_trigger();
  pdl_wait();                                   // state, scratch rows and partial sums of the previous launch
  const int job = blockIdx.y, sp = blockIdx.x, splits = gridDim.x;
  // everything a CTA needs from the state in ONE round trip (independent loads)
  const int n_jobs = s->sm.n_fix + s->sm.n_new;
  const int keep_ = s->sm.keep;
  const int n_in = s->sm.n_in;
  const int ha = s->jobs[job][0], hb = s->jobs[job][1];
  const int avg_dst = s->sm.avg_dst, avg_a = s->sm.avg_a, avg_b = s->sm.avg_b;
  // issued early so that the deciding CTA does not pay a dependent round trip for it
  const int coin = (coins != nullptr && phase == LM_PHASE_FRAME && threadIdx.x == 0 && step < n_in) ? coins[step - keep_] : 0;
  if (job < n_jobs) {
    const bool a_avg = avg_dst >= 0 && ha == avg_dst, b_avg = avg_dst >= 0 && hb == avg_dst;
    const T* pa = lm_row(x, scratch, n_in, L, ha);
    const T* pb = lm_row(x, scratch, n_in, L, hb);
    const long long nvec = L / V;
    const long long per = (nvec + splits - 1) / splits;
    const long long v0 = sp * per, v1 = v0 + per < nvec ? v0 + per : nvec;
    float acc[3] = {0.f, 0.f, 0.f};
    auto fma3 = [&](const float (&fa)[V], const float (&fb)[V]) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        acc[0] = fmaf(fa[k], fb[k], acc[0]);
        acc[1] = fmaf(fa[k], fa[k], acc[1]);
        acc[2] = fmaf(fb[k], fb[k], acc[2]);
      }
    };
    if (a_avg || b_avg) {
      const T* qa = lm_row(x, scratch, n_in, L, avg_a);
      const T* qb = lm_row(x, scratch, n_in, L, avg_b);
      T* pd = job == 0 ? scratch + static_cast<long long>(avg_dst - n_in) * L : nullptr;   // job 0 always names the average
      for (long long v = v0 + threadIdx.x; v < v1; v += LM_THREADS) {
        float fa[V], fb[V], ma[V], mb[V];
        Vec<T>::load_plain(qa + v * V, ma);
        Vec<T>::load_plain(qb + v * V, mb);
        if (!a_avg) Vec<T>::load_plain(pa + v * V, fa);
        if (!b_avg) Vec<T>::load_plain(pb + v * V, fb);
#pragma unroll
        for (int k = 0; k < V; ++k) ma[k] = lm_avg<T>(ma[k], mb[k]);
        if (pd != nullptr) Vec<T>::store(pd + v * V, ma);
        if (a_avg) {
#pragma unroll
          for (int k = 0; k < V; ++k) fa[k] = ma[k];
        }
        if (b_avg) {
#pragma unroll
          for (int k = 0; k < V; ++k) fb[k] = ma[k];
        }
        fma3(fa, fb);
      }
    } else {
      // two vectors per row in flight per thread: the loop is a chain of DRAM round trips otherwise
      for (long long v = v0 + threadIdx.x; v < v1; v += 2 * LM_THREADS) {
        float fa[V], fb[V], ga[V], gb[V];
        const bool two = v + LM_THREADS < v1;
        Vec<T>::load_plain(pa + v * V, fa);
        Vec<T>::load_plain(pb + v * V, fb);
        if (two) {
          Vec<T>::load_plain(pa + (v + LM_THREADS) * V, ga);
          Vec<T>::load_plain(pb + (v + LM_THREADS) * V, gb);
        }
        fma3(fa, fb);
        if (two) fma3(ga, gb);
      }
    }
    block_sum<3>(acc, s_red);
    if (threadIdx.x == 0) {
      float* p = partial + (static_cast<long long>(job) * splits + sp) * 3;
      p[0] = acc[0]; p[1] = acc[1]; p[2] = acc[2];
    }
  }
  // ticket: the last CTA of the grid folds the partial sums (fixed order) and takes the decision
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&s->sm.ticket, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const bool kmode = s->sm.mode == LM_KDROP || s->sm.mode == LM_KMERGE;
  LmSmall* m = reinterpret_cast<LmSmall*>(s_dots + LM_MAX_JOBS * 3);
  float* s_val = reinterpret_cast<float*>(m) + LM_SMALL_WORDS;
  int* s_idx = reinterpret_cast<int*>(s_val + LM_THREADS);
  float* S = reinterpret_cast<float*>(s_idx + LM_THREADS);
  // state in: one parallel round trip
  {
    const int* src = reinterpret_cast<const int*>(&s->sm);
    int* dst = reinterpret_cast<int*>(m);
    for (int i = threadIdx.x; i < LM_SMALL_WORDS; i += LM_THREADS) dst[i] = __ldcg(src + i);
    if (kmode)       // positions 0..keep are the only live ones
      for (int i = threadIdx.x; i < (keep_ + 1) * (keep_ + 1); i += LM_THREADS) {
        const int e = (i / (keep_ + 1)) * (LM_CAP + 1) + i % (keep_ + 1);
        S[e] = __ldcg(s->S + e);
      }
  }
  // fold: one warp per job, lanes over the splits, fixed shuffle tree
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = warp; j < n_jobs; j += LM_THREADS / 32) {
      const float* p = partial + static_cast<long long>(j) * splits * 3;
      float ab = 0.f, aa = 0.f, bb = 0.f;
      for (int i = lane; i < splits; i += 32) {
        ab += __ldcg(p + 3 * i);
        aa += __ldcg(p + 3 * i + 1);
        bb += __ldcg(p + 3 * i + 2);
      }
      ab = warp_sum(ab); aa = warp_sum(aa); bb = warp_sum(bb);
      if (lane == 0) { s_dots[3 * j] = ab; s_dots[3 * j + 1] = aa; s_dots[3 * j + 2] = bb; }
    }
  }
  __syncthreads();
  lm_decide<T>(m, S, s->jobs, s_dots, phase, step, coin, decisions, s_val, s_idx);
  __syncthreads();
  if (threadIdx.x == 0) m->ticket = 0;
  __syncthreads();
  {
    int* dst = reinterpret_cast<int*>(&s->sm);
    const int* src = reinterpret_cast<const int*>(m);
    for (int i = threadIdx.x; i < LM_SMALL_WORDS; i += LM_THREADS) dst[i] = src[i];
    if (kmode)
      for (int i = threadIdx.x; i < (keep_ + 1) * (keep_ + 1); i += LM_THREADS) {
        const int e = (i / (keep_ + 1)) * (LM_CAP + 1) + i % (keep_ + 1);
        s->S[e] = S[e];
      }
  }
}
constexpr size_t LM_STREAM_SMEM = LM_MAX_JOBS * 3 * sizeof(float) + sizeof(LmSmall) + LM_THREADS * 8 +
                                  (LM_CAP + 1) * (LM_CAP + 1) * sizeof(float);

__global__ void lm_init_kernel(LmState* s, int mode, int keep, int n_in) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  LmSmall* m = &s->sm;
  m->mode = mode; m->keep = keep; m->n_in = n_in;
  m->avg_a = m->avg_b = m->avg_dst = -1;
  m->fix_pos = 0;
  m->ticket = 0;
  m->n_fix = 0;
  const int stride = LM_CAP + 1;
  int n = 0;
  if (mode == LM_DROP || mode == LM_MERGE) {
    for (int j = 0; j < keep; ++j) m->order[j] = j;
    for (int j = 0; j + 1 < keep; ++j, ++n) { s->jobs[n][0] = j; s->jobs[n][1] = j + 1; }
  } else {
    for (int j = 0; j < keep; ++j) { m->lorder[j] = j; m->handle[j] = j; }
    for (int p = 0; p <= keep; ++p) s->S[p * stride + p] = LM_NEG;
    m->free_pos = keep;
    for (int a = 0; a < keep; ++a)
      for (int b = a + 1; b < keep; ++b, ++n) { s->jobs[n][0] = a; s->jobs[n][1] = b; }
  }
  m->n_new = n;
  m->n_free = 0;
  for (int j = keep + 2; j >= 0; --j) m->free_slots[m->n_free++] = n_in + j;   // scratch rows, slot 0 handed out first
}

template <typename T>
__global__ void __launch_bounds__(LM_THREADS) lm_finish_kernel(LmBatch b, T* __restrict__ out_all,
                                                               float* __restrict__ out_sim_all, int n_sim, long long L) {
  constexpr int V = Vec<T>::N;
  const int vid = blockIdx.z;
  const T* __restrict__ x = static_cast<const T*>(b.xs != nullptr ? b.xs[vid] : b.x_single);
  const char* wsv = b.ws + static_cast<size_t>(vid) * b.per_video;
  const LmState* s = reinterpret_cast<const LmState*>(wsv);
  const T* scratch = reinterpret_cast<const T*>(wsv + b.scratch_off);
  const LmSmall* m = &s->sm;
  T* out = out_all + static_cast<long long>(vid) * m->keep * L;
  float* out_sim = out_sim_all != nullptr ? out_sim_all + static_cast<long long>(vid) * n_sim : nullptr;
  const int j = blockIdx.y, keep = m->keep;
  const bool adjacent = m->mode == LM_DROP || m->mode == LM_MERGE;
  const int h = adjacent ? m->order[j] : m->handle[m->lorder[j]];
  const T* src = lm_row(x, scratch, m->n_in, L, h);
  T* dst = out + static_cast<long long>(j) * L;
  const long long nvec = L / V;
  for (long long v = static_cast<long long>(blockIdx.x) * LM_THREADS + threadIdx.x; v < nvec;
       v += static_cast<long long>(gridDim.x) * LM_THREADS) {
    float f[V];
    Vec<T>::load_plain(src + v * V, f);
    Vec<T>::store(dst + v * V, f);
  }
  if (blockIdx.x == 0 && out_sim != nullptr) {
    if (adjacent) {
      if (threadIdx.x == 0 && j + 1 < keep) out_sim[j] = m->sim[j];
    } else {
      for (int r = threadIdx.x; r < keep; r += LM_THREADS)
        out_sim[j * keep + r] = s->S[m->lorder[j] * (LM_CAP + 1) + m->lorder[r]];
    }
  }
}

constexpr int LM_STREAM_MAX_SPLITS = 256;
static int lm_splits(long long nvec) {
  long long s = nvec / (LM_THREADS * 4);
  if (s < 1) s = 1;
  if (s > LM_MAX_SPLITS) s = LM_MAX_SPLITS;
  return static_cast<int>(s);
}
// k-means: frames x splits CTAs.  Thin CTAs (a few vectors per thread) measured faster than a few waves of fat ones
// (2.4 vs 3.2 ms for 512 frames x 3 iterations): the loops are latency-bound per thread, so parallelism across CTAs
// is what keeps HBM busy.
static int lm_kmeans_splits(long long nvec, long long frames) {
  (void)frames;
  return lm_splits(nvec);
}
// streaming kernels: a launch is a chain of round trips, so spread each row pair over as many CTAs as stay
// co-resident (4 per SM) with at most ~2 vectors per thread and row
static int lm_stream_splits(long long nvec, int jobs) {
  long long s = (4LL * sm_count()) / std::max(jobs, 1);
  const long long by_work = (nvec + LM_THREADS - 1) / LM_THREADS;
  if (s > by_work) s = by_work;
  if (s > LM_STREAM_MAX_SPLITS) s = LM_STREAM_MAX_SPLITS;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}

static size_t lm_align(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

struct LmLayout {
  size_t state, partial, scratch, total;
  int splits, max_jobs;
};
static LmLayout lm_layout(int64_t row_elems, int keep, int mode, int dtype, int batch = 1) {
  LmLayout l;
  const int vec = dtype == MAVLM_F32 ? 4 : 8;
  const size_t esz = dtype == MAVLM_F32 ? 4 : 2;
  const int init_jobs = (mode == LM_DROP || mode == LM_MERGE) ? keep - 1 : keep * (keep - 1) / 2;
  const int step_jobs = mode == LM_DROP ? 2 : mode == LM_MERGE ? 4 : mode == LM_KDROP ? keep : 2 * keep + 1;
  l.max_jobs = std::max(std::max(init_jobs, step_jobs), 1);
  l.splits = lm_stream_splits(row_elems / vec, step_jobs * std::max(batch, 1));
  l.state = 0;
  l.partial = lm_align(sizeof(LmState));
  l.scratch = l.partial + lm_align(static_cast<size_t>(l.max_jobs) * l.splits * 3 * sizeof(float));
  const bool merges = mode == LM_MERGE || mode == LM_KMERGE;
  l.total = l.scratch + (merges ? lm_align(static_cast<size_t>(keep + 3) * row_elems * esz) : 0);
  return l;
}

template <typename T>
static int lm_stream_launch(const void* x_single, const void* const* xs, const int64_t* n_frames, int batch, int64_t L,
                            int keep, int mode, const uint8_t* coins, int64_t coin_stride, void* out, float* out_sim,
                            int32_t* decisions, void* ws, const LmLayout& lay, cudaStream_t st) {
  LmBatch b;
  b.x_single = x_single; b.xs = xs; b.ws = static_cast<char*>(ws);
  b.per_video = lay.total; b.partial_off = lay.partial; b.scratch_off = lay.scratch;
  b.coins = coins; b.decisions = decisions; b.coin_stride = coin_stride;
  int64_t longest = 0;
  for (int v = 0; v < batch; ++v) {
    lm_init_kernel<<<1, 32, 0, st>>>(reinterpret_cast<LmState*>(b.ws + static_cast<size_t>(v) * lay.total), mode, keep,
                                     static_cast<int>(n_frames[v]));
    MAVLM_LAUNCH_OK();
    longest = std::max(longest, n_frames[v]);
  }
  static bool configured[64][3] = {};     // the attribute is per device and per instantiation
  const int ti = sizeof(T) == 4 ? 0 : (std::is_same<T, __nv_bfloat16>::value ? 1 : 2);
  int dev_id = 0;
  MAVLM_CUDA_OK(cudaGetDevice(&dev_id));
  if (!configured[dev_id & 63][ti]) {
    MAVLM_CUDA_OK(cudaFuncSetAttribute(lm_stream_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(LM_STREAM_SMEM)));
    configured[dev_id & 63][ti] = true;
  }
  const int init_jobs = (mode == LM_DROP || mode == LM_MERGE) ? keep - 1 : keep * (keep - 1) / 2;
  const int step_jobs = mode == LM_DROP ? 2 : mode == LM_MERGE ? 4 : mode == LM_KDROP ? keep : 2 * keep + 1;
  auto launch = [&](int jobs_y, int phase, int step) -> int {
    LaunchCfg lc;
    make_launch(lc, dim3(lay.splits, jobs_y, batch), dim3(LM_THREADS), LM_STREAM_SMEM, st, 1, 8);
    MAVLM_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, lm_stream_kernel<T>, b, static_cast<long long>(L), phase, step));
    MAVLM_LAUNCH_OK();
    return MAVLM_OK;
  };
  int rc = launch(std::max(init_jobs, 1), LM_PHASE_INIT, 0);
  if (rc != MAVLM_OK) return rc;
  for (int64_t i = keep; i < longest; ++i) {
    rc = launch(step_jobs, LM_PHASE_FRAME, static_cast<int>(i));
    if (rc != MAVLM_OK) return rc;
  }
  rc = launch(step_jobs, LM_PHASE_FLUSH, static_cast<int>(longest));
  if (rc != MAVLM_OK) return rc;
  const long long nvec = L / Vec<T>::N;
  const int gx = static_cast<int>(std::min<long long>((nvec + LM_THREADS - 1) / LM_THREADS,
                                                      std::max(1LL, 2LL * sm_count() / (keep * batch))));
  const int n_sim = (mode == LM_DROP || mode == LM_MERGE) ? std::max(keep - 1, 1) : keep * keep;
  lm_finish_kernel<T><<<dim3(std::max(gx, 1), keep, batch), LM_THREADS, 0, st>>>(b, static_cast<T*>(out), out_sim, n_sim, L);
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// scene segmentation: frame means, adjacent cosine, depth scores
// ---------------------------------------------------------------------------------------------------------------
// x [T, P, D] -> out [T, D]: one thread per (frame, 16-byte channel group), P coalesced loads, fp32 sum.
template <typename T>
__global__ void __launch_bounds__(128) frame_mean_kernel(const T* __restrict__ x, T* __restrict__ out, int P, int D) {
  constexpr int V = Vec<T>::N;
  const int g = blockIdx.x * 128 + threadIdx.x;
  if (g * V >= D) return;
  const T* src = x + static_cast<long long>(blockIdx.y) * P * D + g * V;
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  for (int p = 0; p < P; ++p) {
    float f[V];
    Vec<T>::load(src + static_cast<long long>(p) * D, f);
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] += f[k];
  }
  const float inv = 1.0f / static_cast<float>(P);
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] *= inv;
  Vec<T>::store(out + static_cast<long long>(blockIdx.y) * D + g * V, acc);
}

// sim[t] = cos(x[t], x[t+1]) over rows of L elements (ld elements apart); grid (splits, T - 1) partial sums,
// the last CTA of each pair folds them (per-pair ticket).
template <typename T>
__global__ void __launch_bounds__(LM_THREADS) adjacent_cos_kernel(const T* __restrict__ x, long long ld, long long L,
                                                                  float eps, float* partial, unsigned* tickets,
                                                                  float* __restrict__ sim) {
  constexpr int V = Vec<T>::N;
  __shared__ float red[24];
  __shared__ bool s_last;
  const int t = blockIdx.y, sp = blockIdx.x, splits = gridDim.x;
  const T* pa = x + static_cast<long long>(t) * ld;
  const T* pb = pa + ld;
  const long long nvec = L / V;
  const long long per = (nvec + splits - 1) / splits;
  const long long v0 = sp * per, v1 = v0 + per < nvec ? v0 + per : nvec;
  float acc[3] = {0.f, 0.f, 0.f};
  for (long long v = v0 + threadIdx.x; v < v1; v += LM_THREADS) {
    float fa[V], fb[V];
    Vec<T>::load(pa + v * V, fa);
    Vec<T>::load(pb + v * V, fb);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      acc[0] = fmaf(fa[k], fb[k], acc[0]);
      acc[1] = fmaf(fa[k], fa[k], acc[1]);
      acc[2] = fmaf(fb[k], fb[k], acc[2]);
    }
  }
  block_sum<3>(acc, red);
  if (threadIdx.x == 0) {
    float* p = partial + (static_cast<long long>(t) * splits + sp) * 3;
    p[0] = acc[0]; p[1] = acc[1]; p[2] = acc[2];
    __threadfence();
    s_last = atomicAdd(tickets + t, 1u) == static_cast<unsigned>(splits) - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  float ab = 0.f, aa = 0.f, bb = 0.f;
  const float* p = partial + static_cast<long long>(t) * splits * 3;
  for (int i = 0; i < splits; ++i) { ab += __ldcg(p + 3 * i); aa += __ldcg(p + 3 * i + 1); bb += __ldcg(p + 3 * i + 2); }
  sim[t] = round_through<T>(lm_cos(ab, aa, bb, eps));
  tickets[t] = 0;
}

// segment.py:3-25 / :210-223.  One thread per position; the scans stop at the first decrease.
__global__ void depth_score_kernel(const float* __restrict__ sim, float* __restrict__ depth, int n, int left_only) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = sim[i];
  float lpeak = s;
  for (int j = i - 1; j >= 0; --j) {
    const float v = sim[j];
    if (v >= lpeak) lpeak = v; else break;
  }
  if (left_only) { depth[i] = __fsub_rn(lpeak, s); return; }
  float rpeak = s;
  for (int j = i + 1; j < n; ++j) {
    const float v = sim[j];
    if (v >= rpeak) rpeak = v; else break;
  }
  depth[i] = __fsub_rn(__fadd_rn(lpeak, rpeak), __fmul_rn(2.0f, s));
}

// ---------------------------------------------------------------------------------------------------------------
// spatial compression: avg_pool2d (window = stride = k, floor mode) on channels-last tokens, or the mean token
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) avg_pool_kernel(const T* __restrict__ x, T* __restrict__ out, int side, int k,
                                                       int oside, int D) {
  constexpr int V = Vec<T>::N;
  const int g = blockIdx.x * 128 + threadIdx.x;
  if (g * V >= D) return;
  const int o = blockIdx.y, f = blockIdx.z;
  const int oy = o / oside, ox = o - oy * oside;
  const T* src = x + (static_cast<long long>(f) * side * side) * D + g * V;
  float acc[V];
#pragma unroll
  for (int c = 0; c < V; ++c) acc[c] = 0.f;
  for (int dy = 0; dy < k; ++dy)
    for (int dx = 0; dx < k; ++dx) {
      float v[V];
      Vec<T>::load(src + static_cast<long long>((oy * k + dy) * side + ox * k + dx) * D, v);
#pragma unroll
      for (int c = 0; c < V; ++c) acc[c] += v[c];
    }
  const float inv = 1.0f / static_cast<float>(k * k);
#pragma unroll
  for (int c = 0; c < V; ++c) acc[c] *= inv;
  Vec<T>::store(out + (static_cast<long long>(f) * oside * oside + o) * D + g * V, acc);
}

// ---------------------------------------------------------------------------------------------------------------
// k-means over whole frames (rows of L elements)
// ---------------------------------------------------------------------------------------------------------------
// partial[t][split][k] = sum over the split of (x[t] - c[k])^2.  One pass per centroid; the frame slice (26 KB per CTA)
// is re-read from L1/L2.  A variant that kept the slice in registers against 8 centroids at once was slower
// (2.9 vs 2.4 ms for 512 frames x 3 iterations: thin CTAs, the 8-wide block reduction dominated).
template <typename T>
__global__ void __launch_bounds__(LM_THREADS) kmeans_dist_kernel(const T* __restrict__ x, const T* __restrict__ cent,
                                                                 long long L, int K, float* __restrict__ partial) {
  constexpr int V = Vec<T>::N;
  __shared__ float red[8];
  const int t = blockIdx.y, sp = blockIdx.x, splits = gridDim.x;
  const long long nvec = L / V;
  const long long per = (nvec + splits - 1) / splits;
  const long long v0 = sp * per, v1 = v0 + per < nvec ? v0 + per : nvec;
  const T* px = x + static_cast<long long>(t) * L;
  for (int k = 0; k < K; ++k) {
    const T* pc = cent + static_cast<long long>(k) * L;
    float acc[1] = {0.f};
    for (long long v = v0 + threadIdx.x; v < v1; v += LM_THREADS) {
      float fx[V], fc[V];
      Vec<T>::load(px + v * V, fx);
      Vec<T>::load(pc + v * V, fc);
#pragma unroll
      for (int c = 0; c < V; ++c) { const float d = fx[c] - fc[c]; acc[0] = fmaf(d, d, acc[0]); }
    }
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) partial[(static_cast<long long>(t) * splits + sp) * K + k] = acc[0];
  }
}

// labels[t] = first argmin_k sqrt(sum of partials); members grouped by label (counting sort); wsum[k].
// Single CTA (T, K small).
__global__ void kmeans_assign_kernel(const float* __restrict__ partial, int T, int K, int splits,
                                     const float* __restrict__ weights, int* __restrict__ labels,
                                     int* __restrict__ members, int* __restrict__ offsets, float* __restrict__ wsum,
                                     float* __restrict__ dist_out) {
  __shared__ int cursor[1024];
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float best = 0.f;
    int bk = 0;
    for (int k = 0; k < K; ++k) {
      float d2 = 0.f;
      for (int i = 0; i < splits; ++i) d2 += partial[(static_cast<long long>(t) * splits + i) * K + k];
      const float d = sqrtf(d2);
      if (dist_out != nullptr) dist_out[t * K + k] = d;
      if (k == 0 || d < best) { best = d; bk = k; }
    }
    labels[t] = bk;
  }
  __syncthreads();
  if (threadIdx.x == 0) {                      // counting sort by label, O(T + K); weights summed in frame order
    for (int k = 0; k <= K; ++k) offsets[k] = 0;
    for (int k = 0; k < K; ++k) wsum[k] = 0.f;
    for (int t = 0; t < T; ++t) {
      ++offsets[labels[t] + 1];
      wsum[labels[t]] += weights != nullptr ? weights[t] : 1.0f;
    }
    for (int k = 0; k < K; ++k) offsets[k + 1] += offsets[k];
    for (int k = 0; k < K; ++k) cursor[k] = offsets[k];
    for (int t = 0; t < T; ++t) members[cursor[labels[t]]++] = t;
  }
}

// new_cent[k] = sum_{t in k} w_t x[t] / wsum[k]  (plain k-means: w = 1);  diff_partial[k][block] = sum (cent - new)^2.
// grid (column blocks, K); every frame is read once.
template <typename T>
__global__ void __launch_bounds__(LM_THREADS) kmeans_update_kernel(const T* __restrict__ x, const T* __restrict__ cent,
                                                                   T* __restrict__ new_cent, long long L,
                                                                   const int* __restrict__ members,
                                                                   const int* __restrict__ offsets,
                                                                   const float* __restrict__ weights,
                                                                   const float* __restrict__ wsum,
                                                                   float* __restrict__ diff_partial) {
  constexpr int V = Vec<T>::N;
  __shared__ float red[8];
  const int k = blockIdx.y;
  const int m0 = offsets[k], m1 = offsets[k + 1];
  const long long nvec = L / V;
  float dacc[1] = {0.f};
  if (m1 > m0) {
    const float inv = 1.0f / wsum[k];
    for (long long v = static_cast<long long>(blockIdx.x) * LM_THREADS + threadIdx.x; v < nvec;
         v += static_cast<long long>(gridDim.x) * LM_THREADS) {
      float acc[V];
#pragma unroll
      for (int c = 0; c < V; ++c) acc[c] = 0.f;
      int m = m0;
      for (; m + 4 <= m1; m += 4) {          // 4 frames in flight per thread
        float fx[4][V], w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int t = members[m + u];
          w[u] = weights != nullptr ? weights[t] : 1.0f;
          Vec<T>::load(x + static_cast<long long>(t) * L + v * V, fx[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int c = 0; c < V; ++c) acc[c] = fmaf(w[u], fx[u][c], acc[c]);
      }
      for (; m < m1; ++m) {
        const int t = members[m];
        const float w = weights != nullptr ? weights[t] : 1.0f;
        float fx[V];
        Vec<T>::load(x + static_cast<long long>(t) * L + v * V, fx);
#pragma unroll
        for (int c = 0; c < V; ++c) acc[c] = fmaf(w, fx[c], acc[c]);
      }
      float fc[V];
      Vec<T>::load(cent + static_cast<long long>(k) * L + v * V, fc);
#pragma unroll
      for (int c = 0; c < V; ++c) {
        acc[c] = round_through<T>(acc[c] * inv);
        const float d = fc[c] - acc[c];
        dacc[0] = fmaf(d, d, dacc[0]);
      }
      Vec<T>::store(new_cent + static_cast<long long>(k) * L + v * V, acc);
    }
  }
  block_sum<1>(dacc, red);
  if (threadIdx.x == 0) diff_partial[k * gridDim.x + blockIdx.x] = dacc[0];
}

// diff_partial[k][block] = sum (a_k - b_k)^2 for the rows listed (empty clusters re-seeded by the host)
template <typename T>
__global__ void __launch_bounds__(LM_THREADS) row_diff_kernel(const T* __restrict__ a, const T* __restrict__ b, long long L,
                                                              float* __restrict__ diff_partial) {
  constexpr int V = Vec<T>::N;
  __shared__ float red[8];
  const int k = blockIdx.y;
  const long long nvec = L / V;
  float dacc[1] = {0.f};
  for (long long v = static_cast<long long>(blockIdx.x) * LM_THREADS + threadIdx.x; v < nvec;
       v += static_cast<long long>(gridDim.x) * LM_THREADS) {
    float fa[V], fb[V];
    Vec<T>::load(a + static_cast<long long>(k) * L + v * V, fa);
    Vec<T>::load(b + static_cast<long long>(k) * L + v * V, fb);
#pragma unroll
    for (int c = 0; c < V; ++c) { const float d = fa[c] - fb[c]; dacc[0] = fmaf(d, d, dacc[0]); }
  }
  block_sum<1>(dacc, red);
  if (threadIdx.x == 0) diff_partial[k * gridDim.x + blockIdx.x] = dacc[0];
}

// norms[k] = sqrt(sum_blocks diff_partial[k][*])
__global__ void kmeans_norm_kernel(const float* __restrict__ diff_partial, int blocks, int K, float* __restrict__ norms) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float s = 0.f;
  for (int i = 0; i < blocks; ++i) s += diff_partial[k * blocks + i];
  norms[k] = sqrtf(s);
}

// ---------------------------------------------------------------------------------------------------------------
// Turing memory: w = ratio * softmax(scores * scale) per row (fp32 scores from the tensor-core GEMM), and
// mem_scaled = mem * (1 - sum_j w_ij), the term the second GEMM accumulates onto.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(LM_THREADS) ntm_softmax_kernel(const float* __restrict__ scores, long long lds, int n,
                                                                 float scale, float ratio, T* __restrict__ w, long long ldw,
                                                                 int n_pad, const T* __restrict__ mem, T* __restrict__ mem_scaled,
                                                                 int D) {
  __shared__ float red[8];
  __shared__ float s_b[2];
  const long long row = blockIdx.x;
  const float* sr = scores + row * lds;
  float m = -3.0e38f;
  for (int j = threadIdx.x; j < n; j += LM_THREADS) m = fmaxf(m, sr[j] * scale);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = red[0];
    for (int i = 1; i < LM_THREADS / 32; ++i) v = fmaxf(v, red[i]);
    s_b[0] = v;
  }
  __syncthreads();
  m = s_b[0];
  float acc[1] = {0.f};
  for (int j = threadIdx.x; j < n; j += LM_THREADS) acc[0] += __expf(sr[j] * scale - m);
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) s_b[1] = acc[0];
  __syncthreads();
  const float inv = 1.0f / s_b[1];
  // weights in the storage dtype (what the reference's softmax * ratio yields); their rounded sum is the decay
  float dsum[1] = {0.f};
  T* wr = w + row * ldw;
  for (int j = threadIdx.x; j < n_pad; j += LM_THREADS) {
    float v = 0.f;
    if (j < n) v = round_through<T>(round_through<T>(__expf(sr[j] * scale - m) * inv) * ratio);
    if constexpr (sizeof(T) == 4) wr[j] = v; else wr[j] = Elem16<T>::from_float(v);
    dsum[0] += v;
  }
  if (mem == nullptr) return;
  block_sum<1>(dsum, red);
  if (threadIdx.x == 0) s_b[0] = round_through<T>(dsum[0]);
  __syncthreads();
  const float keepf = round_through<T>(1.0f - s_b[0]);
  for (int c = threadIdx.x; c < D; c += LM_THREADS) {
    float v;
    if constexpr (sizeof(T) == 4) v = mem[row * D + c]; else v = Elem16<T>::to_float(mem[row * D + c]);
    v = round_through<T>(v * keepf);
    if constexpr (sizeof(T) == 4) mem_scaled[row * D + c] = v; else mem_scaled[row * D + c] = Elem16<T>::from_float(v);
  }
}

}  // namespace mavlm

using namespace mavlm;

extern "C" {

size_t mavlm_stream_compress_workspace_bytes(int64_t row_elems, int keep, int mode, int dtype) {
  if (row_elems <= 0 || keep < 1 || keep > LM_CAP || mode < 0 || mode > 3 || !dtype_ok(dtype)) return 0;
  return lm_layout(row_elems, keep, mode, dtype).total;
}

size_t mavlm_stream_compress_batched_workspace_bytes(int batch, int64_t row_elems, int keep, int mode, int dtype) {
  if (batch < 1 || row_elems <= 0 || keep < 1 || keep > LM_CAP || mode < 0 || mode > 3 || !dtype_ok(dtype)) return 0;
  return lm_layout(row_elems, keep, mode, dtype, batch).total * static_cast<size_t>(batch);
}

static int lm_check_common(const char* what, int64_t row_elems, int keep, int mode, int dtype, const void* out,
                           const void* decisions, const void* workspace, const uint8_t* coins) {
  MAVLM_REQUIRE(dtype_ok(dtype), MAVLM_E_INVALID, "%s: bad dtype %d", what, dtype);
  MAVLM_REQUIRE(mode >= 0 && mode <= 3, MAVLM_E_INVALID, "%s: mode %d (0 drop, 1 merge, 2 k_drop, 3 k_merge)", what, mode);
  MAVLM_REQUIRE(keep >= 1 && keep <= LM_CAP, MAVLM_E_INVALID, "%s: keep %d must be in [1, %d]", what, keep, LM_CAP);
  const int vec = dtype == MAVLM_F32 ? 4 : 8;
  MAVLM_REQUIRE(row_elems > 0 && row_elems % vec == 0, MAVLM_E_INVALID, "%s: P*D = %lld must be a positive multiple of %d",
                what, static_cast<long long>(row_elems), vec);
  MAVLM_REQUIRE(out != nullptr && decisions != nullptr, MAVLM_E_INVALID, "%s: NULL buffer", what);
  MAVLM_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                MAVLM_E_INVALID, "%s: out must be 16-byte, workspace 256-byte aligned", what);
  MAVLM_REQUIRE((mode != LM_DROP && mode != LM_KDROP) || coins != nullptr, MAVLM_E_INVALID,
                "%s: drop / k_drop need one coin per streamed frame", what);
  return MAVLM_OK;
}

int mavlm_stream_compress_fwd(const void* x, int64_t n_frames, int64_t row_elems, int keep, int mode,
                              const uint8_t* coins, void* out, float* out_sim, int32_t* decisions, void* workspace,
                              size_t workspace_bytes, int dtype, void* stream) {
  const int rc = lm_check_common("stream_compress", row_elems, keep, mode, dtype, out, decisions, workspace, coins);
  if (rc != MAVLM_OK) return rc;
  MAVLM_REQUIRE(n_frames > keep && n_frames < (1 << 30), MAVLM_E_INVALID,
                "stream_compress: needs more frames (%lld) than kept (%d); shorter videos pass through on the host",
                static_cast<long long>(n_frames), keep);
  MAVLM_REQUIRE(x != nullptr && (reinterpret_cast<uintptr_t>(x) & 15) == 0, MAVLM_E_INVALID,
                "stream_compress: x must be a 16-byte aligned device pointer");
  const LmLayout lay = lm_layout(row_elems, keep, mode, dtype);
  MAVLM_REQUIRE(workspace != nullptr && workspace_bytes >= lay.total, MAVLM_E_WORKSPACE,
                "stream_compress: workspace %zu < %zu bytes", workspace_bytes, lay.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MAVLM_DISPATCH_DTYPE(dtype, return lm_stream_launch<T>(x, nullptr, &n_frames, 1, row_elems, keep, mode, coins,
                                                         n_frames - keep, out, out_sim, decisions, workspace, lay, st));
}

int mavlm_stream_compress_batched_fwd(const void* const* x_ptrs, const int64_t* n_frames, int batch, int64_t row_elems,
                                      int keep, int mode, const uint8_t* coins, int64_t coin_stride, void* out,
                                      float* out_sim, int32_t* decisions, void* workspace, size_t workspace_bytes,
                                      int dtype, void* stream) {
  const int rc = lm_check_common("stream_compress_batched", row_elems, keep, mode, dtype, out, decisions, workspace, coins);
  if (rc != MAVLM_OK) return rc;
  MAVLM_REQUIRE(batch >= 1 && batch <= 65535 && x_ptrs != nullptr && n_frames != nullptr, MAVLM_E_INVALID,
                "stream_compress_batched: batch %d / NULL pointer table", batch);
  int64_t longest = 0;
  for (int v = 0; v < batch; ++v) {
    MAVLM_REQUIRE(n_frames[v] > keep && n_frames[v] < (1 << 30), MAVLM_E_INVALID,
                  "stream_compress_batched: video %d has %lld frames, needs more than keep = %d", v,
                  static_cast<long long>(n_frames[v]), keep);
    longest = std::max(longest, n_frames[v]);
  }
  MAVLM_REQUIRE(coin_stride >= longest - keep, MAVLM_E_INVALID, "stream_compress_batched: coin / decision stride %lld < %lld",
                static_cast<long long>(coin_stride), static_cast<long long>(longest - keep));
  const LmLayout lay = lm_layout(row_elems, keep, mode, dtype, batch);
  MAVLM_REQUIRE(workspace != nullptr && workspace_bytes >= lay.total * static_cast<size_t>(batch), MAVLM_E_WORKSPACE,
                "stream_compress_batched: workspace %zu < %zu bytes", workspace_bytes, lay.total * static_cast<size_t>(batch));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MAVLM_DISPATCH_DTYPE(dtype, return lm_stream_launch<T>(nullptr, x_ptrs, n_frames, batch, row_elems, keep, mode, coins,
                                                         coin_stride, out, out_sim, decisions, workspace, lay, st));
}

int mavlm_frame_mean_fwd(const void* x, void* out, int frames, int tokens, int dim, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype_ok(dtype), MAVLM_E_INVALID, "frame_mean: bad dtype %d", dtype);
  const int vec = dtype == MAVLM_F32 ? 4 : 8;
  MAVLM_REQUIRE(dim > 0 && dim % vec == 0 && tokens > 0, MAVLM_E_INVALID, "frame_mean: dim %d must be a multiple of %d", dim, vec);
  MAVLM_REQUIRE(frames >= 0 && frames <= 65535, MAVLM_E_INVALID, "frame_mean: at most 65535 frames per call (got %d)", frames);
  if (frames == 0) return MAVLM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(ceil_div(dim / vec, 128), frames);
  MAVLM_DISPATCH_DTYPE(dtype, (frame_mean_kernel<T><<<grid, 128, 0, st>>>(static_cast<const T*>(x), static_cast<T*>(out),
                                                                         tokens, dim)));
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

size_t mavlm_adjacent_cosine_workspace_bytes(int64_t rows, int64_t row_elems, int dtype) {
  if (rows < 2 || row_elems <= 0 || !dtype_ok(dtype)) return 0;
  const int splits = lm_splits(row_elems / (dtype == MAVLM_F32 ? 4 : 8));
  return lm_align(static_cast<size_t>(rows - 1) * splits * 3 * sizeof(float)) + lm_align((rows - 1) * sizeof(unsigned));
}

int mavlm_adjacent_cosine_fwd(const void* x, int64_t rows, int64_t row_elems, int64_t ld, float eps, float* sim,
                              void* workspace, size_t workspace_bytes, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype_ok(dtype), MAVLM_E_INVALID, "adjacent_cosine: bad dtype %d", dtype);
  const int vec = dtype == MAVLM_F32 ? 4 : 8;
  MAVLM_REQUIRE(row_elems > 0 && row_elems % vec == 0 && ld % vec == 0 && ld >= row_elems, MAVLM_E_INVALID,
                "adjacent_cosine: row length %lld / stride %lld must be multiples of %d", static_cast<long long>(row_elems),
                static_cast<long long>(ld), vec);
  if (rows < 2) return MAVLM_OK;
  MAVLM_REQUIRE(rows <= 65536, MAVLM_E_INVALID, "adjacent_cosine: at most 65536 rows");
  const size_t need = mavlm_adjacent_cosine_workspace_bytes(rows, row_elems, dtype);
  MAVLM_REQUIRE(workspace != nullptr && workspace_bytes >= need, MAVLM_E_WORKSPACE, "adjacent_cosine: workspace %zu < %zu bytes",
                workspace_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int splits = lm_splits(row_elems / vec);
  float* partial = static_cast<float*>(workspace);
  unsigned* tickets = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) +
                                                  lm_align(static_cast<size_t>(rows - 1) * splits * 3 * sizeof(float)));
  MAVLM_CUDA_OK(cudaMemsetAsync(tickets, 0, (rows - 1) * sizeof(unsigned), st));
  const dim3 grid(splits, static_cast<unsigned>(rows - 1));
  MAVLM_DISPATCH_DTYPE(dtype, (adjacent_cos_kernel<T><<<grid, LM_THREADS, 0, st>>>(static_cast<const T*>(x), ld, row_elems, eps,
                                                                                  partial, tickets, sim)));
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

int mavlm_depth_scores_fwd(const float* sim, float* depth, int n, int left_only, void* stream) {
  MAVLM_REQUIRE(n >= 0 && (n == 0 || (sim != nullptr && depth != nullptr)), MAVLM_E_INVALID, "depth_scores: NULL buffer");
  if (n == 0) return MAVLM_OK;
  depth_score_kernel<<<ceil_div(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(sim, depth, n, left_only);
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

int mavlm_avg_pool_fwd(const void* x, void* out, int frames, int side, int window, int dim, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype_ok(dtype), MAVLM_E_INVALID, "avg_pool: bad dtype %d", dtype);
  const int vec = dtype == MAVLM_F32 ? 4 : 8;
  MAVLM_REQUIRE(dim > 0 && dim % vec == 0, MAVLM_E_INVALID, "avg_pool: dim %d must be a multiple of %d", dim, vec);
  MAVLM_REQUIRE(side > 0 && window > 0 && window <= side, MAVLM_E_INVALID, "avg_pool: window %d on a %d x %d grid", window, side, side);
  if (frames == 0) return MAVLM_OK;
  const int oside = (side - window) / window + 1;
  MAVLM_REQUIRE(frames <= 65535, MAVLM_E_INVALID, "avg_pool: at most 65535 frames per call");
  const dim3 grid(ceil_div(dim / vec, 128), oside * oside, frames);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MAVLM_DISPATCH_DTYPE(dtype, (avg_pool_kernel<T><<<grid, 128, 0, st>>>(static_cast<const T*>(x), static_cast<T*>(out), side,
                                                                       window, oside, dim)));
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

size_t mavlm_kmeans_workspace_bytes(int64_t n_frames, int64_t row_elems, int clusters, int dtype) {
  if (n_frames <= 0 || row_elems <= 0 || clusters < 1 || !dtype_ok(dtype)) return 0;
  const int splits = lm_kmeans_splits(row_elems / (dtype == MAVLM_F32 ? 4 : 8), n_frames);
  return lm_align(static_cast<size_t>(n_frames) * splits * clusters * sizeof(float)) +   // distance partials
         lm_align(static_cast<size_t>(n_frames) * sizeof(int)) +                         // members
         lm_align(static_cast<size_t>(clusters + 1) * sizeof(int)) +                     // offsets
         lm_align(static_cast<size_t>(clusters) * 2 * sm_count() * sizeof(float));       // diff partials
}

// One Lloyd iteration: labels / weight sums against `cent`, new centroids for the non-empty clusters, and
// diff[k] = |cent_k - new_k| for them (0 for empty clusters, which the caller re-seeds; see mavlm_row_distance_fwd).
int mavlm_kmeans_iter_fwd(const void* x, const float* weights, const void* cent, void* new_cent, int32_t* labels,
                          float* wsum, float* diff, float* dist, int64_t n_frames, int64_t row_elems, int clusters,
                          void* workspace, size_t workspace_bytes, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype_ok(dtype), MAVLM_E_INVALID, "kmeans: bad dtype %d", dtype);
  const int vec = dtype == MAVLM_F32 ? 4 : 8;
  MAVLM_REQUIRE(row_elems > 0 && row_elems % vec == 0, MAVLM_E_INVALID, "kmeans: P*D must be a multiple of %d", vec);
  MAVLM_REQUIRE(n_frames > 0 && n_frames <= 65535 && clusters >= 1 && clusters <= 1024, MAVLM_E_INVALID,
                "kmeans: frames %lld / clusters %d out of range", static_cast<long long>(n_frames), clusters);
  const size_t need = mavlm_kmeans_workspace_bytes(n_frames, row_elems, clusters, dtype);
  MAVLM_REQUIRE(workspace != nullptr && workspace_bytes >= need, MAVLM_E_WORKSPACE, "kmeans: workspace %zu < %zu bytes",
                workspace_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int splits = lm_kmeans_splits(row_elems / vec, n_frames);
  char* base = static_cast<char*>(workspace);
  float* partial = reinterpret_cast<float*>(base);
  base += lm_align(static_cast<size_t>(n_frames) * splits * clusters * sizeof(float));
  int* members = reinterpret_cast<int*>(base);
  base += lm_align(static_cast<size_t>(n_frames) * sizeof(int));
  int* offsets = reinterpret_cast<int*>(base);
  base += lm_align(static_cast<size_t>(clusters + 1) * sizeof(int));
  float* dpart = reinterpret_cast<float*>(base);
  const int blocks = static_cast<int>(std::min<long long>((row_elems / vec + LM_THREADS - 1) / LM_THREADS, 2LL * sm_count()));
  MAVLM_DISPATCH_DTYPE(dtype, (kmeans_dist_kernel<T><<<dim3(splits, static_cast<unsigned>(n_frames)), LM_THREADS, 0, st>>>(
                                  static_cast<const T*>(x), static_cast<const T*>(cent), row_elems, clusters, partial)));
  MAVLM_LAUNCH_OK();
  kmeans_assign_kernel<<<1, 256, 0, st>>>(partial, static_cast<int>(n_frames), clusters, splits, weights, labels, members,
                                          offsets, wsum, dist);
  MAVLM_LAUNCH_OK();
  if (new_cent == nullptr) return MAVLM_OK;   // distances / labels only
  MAVLM_DISPATCH_DTYPE(dtype, (kmeans_update_kernel<T><<<dim3(blocks, clusters), LM_THREADS, 0, st>>>(
                                  static_cast<const T*>(x), static_cast<const T*>(cent), static_cast<T*>(new_cent), row_elems,
                                  members, offsets, weights, wsum, dpart)));
  MAVLM_LAUNCH_OK();
  kmeans_norm_kernel<<<ceil_div(clusters, 128), 128, 0, st>>>(dpart, blocks, clusters, diff);
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

// dist[k] = |a_k - b_k| over rows of row_elems (k < rows); workspace as for mavlm_kmeans_iter_fwd with clusters = rows.
int mavlm_row_distance_fwd(const void* a, const void* b, float* dist, int rows, int64_t row_elems, void* workspace,
                           size_t workspace_bytes, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype_ok(dtype), MAVLM_E_INVALID, "row_distance: bad dtype %d", dtype);
  const int vec = dtype == MAVLM_F32 ? 4 : 8;
  MAVLM_REQUIRE(row_elems > 0 && row_elems % vec == 0 && rows >= 0 && rows <= 65535, MAVLM_E_INVALID, "row_distance: bad shape");
  if (rows == 0) return MAVLM_OK;
  const int blocks = static_cast<int>(std::min<long long>((row_elems / vec + LM_THREADS - 1) / LM_THREADS, 2LL * sm_count()));
  MAVLM_REQUIRE(workspace != nullptr && workspace_bytes >= static_cast<size_t>(rows) * blocks * sizeof(float), MAVLM_E_WORKSPACE,
                "row_distance: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* dpart = static_cast<float*>(workspace);
  MAVLM_DISPATCH_DTYPE(dtype, (row_diff_kernel<T><<<dim3(blocks, rows), LM_THREADS, 0, st>>>(static_cast<const T*>(a),
                                                                                           static_cast<const T*>(b), row_elems, dpart)));
  MAVLM_LAUNCH_OK();
  kmeans_norm_kernel<<<ceil_div(rows, 128), 128, 0, st>>>(dpart, blocks, rows, dist);
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

int mavlm_ntm_softmax_fwd(const float* scores, int64_t ld_scores, int64_t rows, int n, float scale, float ratio, void* w,
                          int64_t ld_w, int n_pad, const void* mem, void* mem_scaled, int dim, int dtype, void* stream) {
  MAVLM_REQUIRE(dtype_ok(dtype), MAVLM_E_INVALID, "ntm_softmax: bad dtype %d", dtype);
  MAVLM_REQUIRE(n > 0 && n_pad >= n && ld_w >= n_pad && ld_scores >= n, MAVLM_E_INVALID, "ntm_softmax: bad row geometry");
  MAVLM_REQUIRE((mem == nullptr) == (mem_scaled == nullptr), MAVLM_E_INVALID, "ntm_softmax: mem and mem_scaled go together");
  if (rows == 0) return MAVLM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MAVLM_DISPATCH_DTYPE(dtype, (ntm_softmax_kernel<T><<<static_cast<unsigned>(rows), LM_THREADS, 0, st>>>(
                                  scores, ld_scores, n, scale, ratio, static_cast<T*>(w), ld_w, n_pad,
                                  static_cast<const T*>(mem), static_cast<T*>(mem_scaled), dim)));
  MAVLM_LAUNCH_OK();
  return MAVLM_OK;
}

}  // extern "C"
