// K2: alpha / beta forward-backward log-sum-exp recursion + loss read-out.
//
// Replaces the three tf.while_loop / TensorArray "unfold" loops (tf_seq2seq_losses/tools.py:191-277) driving
//   simplified: alpha_step simplified_ctc_loss.py:393-424, beta_step :327-343, loss :73-83
//   classic:    _alpha_step classic_ctc_loss.py:415-451, beta_step :349-364, loss :152-165
// One warp per (utterance, direction): alpha and beta run concurrently as independent CTAs.  The state vector
// lives in registers, NS consecutive states per lane; the l-1 / l+1 neighbour crosses lanes with one shuffle per
// step; the per-frame inputs (h[t], d[t,.]) stream through a kStages-deep cp.async shared-memory ring.
//
// Classic transition algebra (s=0 closed, s=1 open; rep[l] = label[l]==label[l-1]; r[l] = d[l-1] when label[l-1]
// is not the blank):
//   A'[l,0] = h + S[l],  S[l] = lse(A[l,0], A[l,1])
//   A'[l,1] = lse(r[l] + A[l,1], d[l-1] + (rep[l-1] ? A[l-1,0] : S[l-1]))
//   B'[l,0] = lse(h + B[l,0], d[l] + B[l+1,1])
//   B'[l,1] = lse(rep[l] ? h + B[l,0] : B'[l,0], r[l] + B[l,1])
// which is the reference's [next,prev] table form (classic_ctc_loss.py:464-563) with the -inf entries removed.
#include "common.cuh"
#include "recursion.cuh"

namespace ctcb200 {

constexpr int kGroup = 4;      // frames per unrolled group (the renormalisation cadence)
// cp.async ring depth: frames in flight per warp (DRAM latency / step time).  The wide state vectors beyond the fused
// kernel's range (NS > 16, 4 KB frames and a longer step) take half the depth, which keeps the ring under 48 KB.
template <int NS>
struct K2Stages { static constexpr int value = NS > 16 ? 8 : 16; };

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Ring slot = one frame: Upad label-token log-probs (private layout) followed by h.  Always commits a group (an
// empty one when there is nothing left to fetch) so that wait_group counts stay uniform.  Upad = 32 * NS is a
// compile-time constant, so the copy is a fixed, predicated sequence of 16-byte cp.async with no loop.
template <int NS>
__device__ __forceinline__ void issue_frame(const float* d_src, const float* h_src, float* ring, int slot, bool valid,
                                            int lane) {
  constexpr int kUpad = NS * kWarp, kChunks = kUpad / 4;
  if (valid) {
    float* dst = ring + slot * (kUpad + 4);
#pragma unroll
    for (int k = 0; k < (kChunks + kWarp - 1) / kWarp; ++k) {
      const int c = k * kWarp + lane;
      if (c < kChunks) cp_async16(dst + 4 * c, d_src + 4 * c);
    }
    if (lane == 0) cp_async4(dst + kUpad, h_src);
  }
  cp_async_commit();
}

template <int NS>
__device__ __forceinline__ void read_frame(const float* ring, int slot, int lane, float* d, float& h) {
  constexpr int kUpad = NS * kWarp;
  const float* src = ring + slot * (kUpad + 4);
#pragma unroll
  for (int j = 0; j < NS; ++j) d[j] = src[j * kWarp + lane];
  h = src[kUpad];
}

// ---- the kernel ------------------------------------------------------------------------------------------------
// grid (B, 2): blockIdx.y == 0 alpha (forward), 1 beta (backward).  full_states: also cover the padded frames
// t >= logit_length and beta[0] so that the reference-layout export can read every row.
template <int NS, bool CLASSIC>
__global__ void __launch_bounds__(kWarp) k2_recursion(Problem p, Scratch s, float* loss, bool full_states) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int L = utt_label_len(p, b), n_t = utt_frames(p, b);
  constexpr int S = CLASSIC ? 2 : 1;
  const size_t row_pitch = (size_t)S * p.Upad;
  const long long frame0 = (long long)b * p.T;
  LabelBits<NS> lb;
  if (CLASSIC) lb = make_label_bits<NS>(p, b, L, lane);
  extern __shared__ __align__(16) float ring[];
  constexpr int kStages = K2Stages<NS>::value;

  if (blockIdx.y == 0) {
    // ------------------------------------------------ alpha: t = 0 .. n_t-1 ----------------------------------
    float a0[NS], a1[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      a0[j] = (lane * NS + j == 0) ? 0.0f : kNegInf;     // first alpha slice: log one_hot(0)
      a1[j] = kNegInf;
    }
    float* out = s.alphaT + (size_t)b * (p.T + 1) * row_pitch;
    double* offs = s.ca + (size_t)b * (p.T + 1);
    double c = 0.0;
    float m_pend = kNegInf;
    store_row<NS>(out, a0, lane);
    if (CLASSIC) store_row<NS>(out + p.Upad, a1, lane);
    if (lane == 0) offs[0] = 0.0;
    const int n_run = full_states ? p.T : n_t;
    const float* dsrc = s.dT + (size_t)frame0 * p.Upad;
    const float* hsrc = s.h + frame0;
    for (int k = 0; k < kStages - 1; ++k) issue_frame<NS>(dsrc + (size_t)k * (NS * kWarp), hsrc + k, ring, k, k < n_t, lane);
    for (int t0 = 0; t0 < n_run; t0 += kGroup) {
#pragma unroll
      for (int k = 0; k < kGroup; ++k) {
        const int t = t0 + k;
        if (t < n_run) {
          float d[NS], h;
          if (t < n_t) {
            cp_async_wait<kStages - 2>();     // frame t has landed
            __syncwarp();
            read_frame<NS>(ring, t % kStages, lane, d, h);
            const int nf = t + kStages - 1;   // refill the slot consumed one frame ago
            issue_frame<NS>(dsrc + (size_t)nf * (NS * kWarp), hsrc + nf, ring, nf % kStages, nf < n_t, lane);
          } else {   // padded frame: blank with probability one (base_loss.py:378-393)
#pragma unroll
            for (int j = 0; j < NS; ++j) d[j] = kNegInf;
            h = 0.0f;
          }
          if (CLASSIC) alpha_step_classic<NS>(a0, a1, d, h, lane, lb);
          else alpha_step_simplified<NS>(a0, d, h, lane);
          if (k == 0) m_pend = state_max<NS, CLASSIC>(a0, a1);
          if (k == 2) apply_offset<NS, CLASSIC>(a0, a1, m_pend, c);
          float* o = out + (size_t)(t + 1) * row_pitch;
          store_row<NS>(o, a0, lane);
          if (CLASSIC) store_row<NS>(o + p.Upad, a1, lane);
          if (lane == 0) offs[t + 1] = c;
        }
      }
    }
    // loss = -alpha[T, label_length] (classic: logsumexp over the two states); frames >= n_t leave it unchanged
#pragma unroll
    for (int j = 0; j < NS; ++j)
      if (lane * NS + j == L) {
        const double ld = -((double)(CLASSIC ? lse2(a0[j], a1[j]) : a0[j]) + c);
        s.lossd[b] = ld;
        if (loss != nullptr) loss[b] = (float)ld;
      }
  } else {
    // ------------------------------------------------ beta: t = n_t-1 .. 0 -----------------------------------
    float b0[NS], b1[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      b0[j] = (lane * NS + j == L) ? 0.0f : kNegInf;      // last beta slice: log one_hot(label_length), both states
      b1[j] = b0[j];
    }
    float* out = s.betaT + (size_t)b * (p.T + 1) * row_pitch;
    double* offs = s.cb + (size_t)b * (p.T + 1);
    double c = 0.0;
    float m_pend = kNegInf;
    // frames t >= n_t leave beta unchanged: rows n_t .. T all equal the initial slice
    const int t_hi = full_states ? p.T : n_t;
    for (int t = n_t; t <= t_hi; ++t) {
      float* o = out + (size_t)t * row_pitch;
      store_row<NS>(o, b0, lane);
      if (CLASSIC) store_row<NS>(o + p.Upad, b1, lane);
      if (lane == 0) offs[t] = 0.0;
    }
    const int t_lo = full_states ? 0 : 1;                  // beta[0] is not needed by the gradient
    // i-th processed frame is n_t-1-i; it lives in ring slot i % kStages
    const float* dsrc = s.dT + (size_t)frame0 * p.Upad;
    const float* hsrc = s.h + frame0;
    for (int k = 0; k < kStages - 1; ++k) {
      const int f = n_t - 1 - k;
      issue_frame<NS>(dsrc + (ptrdiff_t)f * (NS * kWarp), hsrc + f, ring, k, f >= t_lo, lane);
    }
    for (int t0 = n_t - 1; t0 >= t_lo; t0 -= kGroup) {
#pragma unroll
      for (int k = 0; k < kGroup; ++k) {
        const int t = t0 - k;
        if (t >= t_lo) {
          float d[NS], h;
          const int i = n_t - 1 - t;
          cp_async_wait<kStages - 2>();
          __syncwarp();
          read_frame<NS>(ring, i % kStages, lane, d, h);
          const int ni = i + kStages - 1, nf = n_t - 1 - ni;
          issue_frame<NS>(dsrc + (ptrdiff_t)nf * (NS * kWarp), hsrc + nf, ring, ni % kStages, nf >= t_lo, lane);
          if (CLASSIC) beta_step_classic<NS>(b0, b1, d, h, lane, lb);
          else beta_step_simplified<NS>(b0, d, h, lane);
          if (k == 0) m_pend = state_max<NS, CLASSIC>(b0, b1);
          if (k == 2) apply_offset<NS, CLASSIC>(b0, b1, m_pend, c);
          float* o = out + (size_t)t * row_pitch;
          store_row<NS>(o, b0, lane);
          if (CLASSIC) store_row<NS>(o + p.Upad, b1, lane);
          if (lane == 0) offs[t] = c;
        }
      }
    }
  }
}

template <int NS>
static cudaError_t launch_ns(const Problem& p, const Scratch& s, float* loss, bool full, cudaStream_t st) {
  dim3 grid(p.B, 2);
  const size_t smem = (size_t)K2Stages<NS>::value * (p.Upad + 4) * sizeof(float);     // <= 33 KB
  if (p.variant == CTCB200_CLASSIC) k2_recursion<NS, true><<<grid, kWarp, smem, st>>>(p, s, loss, full);
  else k2_recursion<NS, false><<<grid, kWarp, smem, st>>>(p, s, loss, full);
  return cudaGetLastError();
}

cudaError_t launch_recursion(const Problem& p, const Scratch& s, float* loss, bool full_states, cudaStream_t st) {
  if (p.B == 0) return cudaSuccess;
  switch (p.NS) {
#define CTCB200_CASE(n) \
  case n:               \
    return launch_ns<n>(p, s, loss, full_states, st);
    CTCB200_CASE(1) CTCB200_CASE(2) CTCB200_CASE(3) CTCB200_CASE(4) CTCB200_CASE(5) CTCB200_CASE(6)
    CTCB200_CASE(7) CTCB200_CASE(8) CTCB200_CASE(9) CTCB200_CASE(10) CTCB200_CASE(11) CTCB200_CASE(12)
    CTCB200_CASE(13) CTCB200_CASE(14) CTCB200_CASE(15) CTCB200_CASE(16)
    CTCB200_CASE(20) CTCB200_CASE(24) CTCB200_CASE(28) CTCB200_CASE(32)      // U > 512: make_problem rounds NS up
#undef CTCB200_CASE
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace ctcb200
