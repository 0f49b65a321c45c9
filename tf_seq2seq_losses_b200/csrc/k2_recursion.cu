// K2: alpha / beta forward-backward log-sum-exp recursion + loss read-out.
//
// Replaces the three tf.while_loop / TensorArray "unfold" loops (tf_seq2seq_losses/tools.py:191-277) driving
//   simplified: alpha_step simplified_ctc_loss.py:393-424, beta_step :327-343, loss :73-83
//   classic:    _alpha_step classic_ctc_loss.py:415-451, beta_step :349-364, loss :152-165
// One warp per (utterance, direction): alpha and beta run concurrently as independent CTAs.  The state vector
// lives in registers, NS consecutive states per lane; the l-1 / l+1 neighbour crosses lanes with one shuffle per
// step; the per-frame inputs (h[t], d[t,.]) are prefetched kPrefetch frames ahead with coalesced loads.
//
// Classic transition algebra (s=0 closed, s=1 open; rep[l] = label[l]==label[l-1]; r[l] = d[l-1] when label[l-1]
// is not the blank):
//   A'[l,0] = h + S[l],  S[l] = lse(A[l,0], A[l,1])
//   A'[l,1] = lse(r[l] + A[l,1], d[l-1] + (rep[l-1] ? A[l-1,0] : S[l-1]))
//   B'[l,0] = lse(h + B[l,0], d[l] + B[l+1,1])
//   B'[l,1] = lse(rep[l] ? h + B[l,0] : B'[l,0], r[l] + B[l,1])
// which is the reference's [next,prev] table form (classic_ctc_loss.py:464-563) with the -inf entries removed.
#include "common.cuh"

namespace ctcb200 {

constexpr int kPrefetch = 4;

template <int NS>
struct FrameQueue {
  float d[kPrefetch][NS];
  float h[kPrefetch];
};

template <int NS>
__device__ __forceinline__ void load_frame(const Problem& p, const Scratch& s, long long row, int lane, float* d,
                                           float& h) {
  const float* src = s.dT + (size_t)row * p.Upad + lane;
#pragma unroll
  for (int j = 0; j < NS; ++j) d[j] = __ldg(src + j * kWarp);
  h = __ldg(s.h + row);
}

// per-lane static label facts for the classic variant
template <int NS>
struct LabelBits {
  unsigned rep;    // bit j: label[l] == label[l-1]        (l = lane*NS + j; label[-1] := blank)
  unsigned nb;     // bit j: label[l] != blank
  bool rep_left;   // rep / nb of state lane*NS - 1 (lives in lane-1)
  bool nb_left;
};

template <int NS>
__device__ __forceinline__ LabelBits<NS> make_label_bits(const Problem& p, int b, int L, int lane) {
  LabelBits<NS> lb;
  lb.rep = 0u;
  lb.nb = 0u;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int l = lane * NS + j;
    const int tok = utt_token(p, b, l, L);
    const int prev = utt_token(p, b, l - 1, L);
    if (tok == prev) lb.rep |= 1u << j;
    if (tok != p.blank) lb.nb |= 1u << j;
  }
  const unsigned rl = __shfl_up_sync(kFull, lb.rep, 1), nl = __shfl_up_sync(kFull, lb.nb, 1);
  lb.rep_left = lane > 0 && ((rl >> (NS - 1)) & 1u);
  lb.nb_left = lane > 0 && ((nl >> (NS - 1)) & 1u);
  return lb;
}

// ---- one frame of each recursion ------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void alpha_step_simplified(float* a, const float* d, float h, int lane) {
  float carry = __shfl_up_sync(kFull, d[NS - 1] + a[NS - 1], 1);
  if (lane == 0) carry = kNegInf;
#pragma unroll
  for (int j = NS - 1; j >= 1; --j) a[j] = lse2(h + a[j], d[j - 1] + a[j - 1]);
  a[0] = lse2(h + a[0], carry);
}

template <int NS>
__device__ __forceinline__ void beta_step_simplified(float* bt, const float* d, float h, int lane) {
  float carry = __shfl_down_sync(kFull, bt[0], 1);
  if (lane == 31) carry = kNegInf;
#pragma unroll
  for (int j = 0; j < NS - 1; ++j) bt[j] = lse2(h + bt[j], d[j] + bt[j + 1]);
  bt[NS - 1] = lse2(h + bt[NS - 1], d[NS - 1] + carry);
}

template <int NS>
__device__ __forceinline__ void alpha_step_classic(float* a0, float* a1, const float* d, float h, int lane,
                                                   const LabelBits<NS>& lb) {
  // d of the left neighbour's top state: pure data, off the dependency chain
  float d_left = __shfl_up_sync(kFull, d[NS - 1], 1);
  if (lane == 0) d_left = kNegInf;
  float S[NS];
#pragma unroll
  for (int j = NS - 1; j >= 0; --j) S[j] = lse2(a0[j], a1[j]);
  const float x_top = ((lb.rep >> (NS - 1)) & 1u) ? a0[NS - 1] : S[NS - 1];
  float x_left = __shfl_up_sync(kFull, x_top, 1);
  if (lane == 0) x_left = kNegInf;
#pragma unroll
  for (int j = NS - 1; j >= 1; --j) {
    const float x = ((lb.rep >> (j - 1)) & 1u) ? a0[j - 1] : S[j - 1];
    const float r = ((lb.nb >> (j - 1)) & 1u) ? d[j - 1] : kNegInf;
    a1[j] = lse2(r + a1[j], d[j - 1] + x);
  }
  {
    const float r = lb.nb_left ? d_left : kNegInf;
    a1[0] = lse2(r + a1[0], d_left + x_left);
  }
#pragma unroll
  for (int j = 0; j < NS; ++j) a0[j] = h + S[j];
}

template <int NS>
__device__ __forceinline__ void beta_step_classic(float* b0, float* b1, const float* d, float h, int lane,
                                                  const LabelBits<NS>& lb) {
  float d_left = __shfl_up_sync(kFull, d[NS - 1], 1);
  if (lane == 0) d_left = kNegInf;
  float carry = __shfl_down_sync(kFull, b1[0], 1);
  if (lane == 31) carry = kNegInf;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const float nxt = (j < NS - 1) ? b1[j + 1] : carry;      // old B[l+1,1]
    const float stay = h + b0[j];
    const float n0 = lse2(stay, d[j] + nxt);
    const float dl = (j > 0) ? d[j - 1] : d_left;
    const bool nbl = (j > 0) ? ((lb.nb >> (j - 1)) & 1u) : lb.nb_left;
    const float r = nbl ? dl : kNegInf;
    const float base = ((lb.rep >> j) & 1u) ? stay : n0;
    b1[j] = lse2(base, r + b1[j]);                            // uses old b1[j]; b1[j+1] already consumed above
    b0[j] = n0;
  }
}

// Offset renormalisation (see Scratch in common.cuh).  The warp maximum is taken right after frame k == 0 of every
// kPrefetch-frame group and subtracted two frames later, so its five dependent shuffles overlap the next frames'
// arithmetic instead of lengthening the serial chain.
template <int NS, bool CLASSIC>
__device__ __forceinline__ float state_max(const float* v0, const float* v1) {
  float m = kNegInf;
#pragma unroll
  for (int j = 0; j < NS; ++j) m = fmaxf(m, CLASSIC ? fmaxf(v0[j], v1[j]) : v0[j]);
  return warp_max(m);
}
template <int NS, bool CLASSIC>
__device__ __forceinline__ void apply_offset(float* v0, float* v1, float m, double& c) {
  if (m == kNegInf) return;              // nothing reachable: leave the -inf vector alone
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    v0[j] -= m;
    if (CLASSIC) v1[j] -= m;
  }
  c += (double)m;
}

template <int NS>
__device__ __forceinline__ void store_row(float* dst, const float* v, int lane) {
#pragma unroll
  for (int j = 0; j < NS; ++j) dst[j * kWarp + lane] = v[j];
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
  FrameQueue<NS> q;

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
#pragma unroll
    for (int k = 0; k < kPrefetch; ++k)
      if (k < n_t) load_frame<NS>(p, s, frame0 + k, lane, q.d[k], q.h[k]);
    for (int t0 = 0; t0 < n_run; t0 += kPrefetch) {
#pragma unroll
      for (int k = 0; k < kPrefetch; ++k) {
        const int t = t0 + k;
        if (t < n_run) {
          float d[NS], h;
          if (t < n_t) {
#pragma unroll
            for (int j = 0; j < NS; ++j) d[j] = q.d[k][j];
            h = q.h[k];
            if (t + kPrefetch < n_t) load_frame<NS>(p, s, frame0 + t + kPrefetch, lane, q.d[k], q.h[k]);
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
#pragma unroll
    for (int k = 0; k < kPrefetch; ++k)
      if (n_t - 1 - k >= t_lo) load_frame<NS>(p, s, frame0 + n_t - 1 - k, lane, q.d[k], q.h[k]);
    for (int t0 = n_t - 1; t0 >= t_lo; t0 -= kPrefetch) {
#pragma unroll
      for (int k = 0; k < kPrefetch; ++k) {
        const int t = t0 - k;
        if (t >= t_lo) {
          float d[NS], h;
#pragma unroll
          for (int j = 0; j < NS; ++j) d[j] = q.d[k][j];
          h = q.h[k];
          if (t - kPrefetch >= t_lo) load_frame<NS>(p, s, frame0 + t - kPrefetch, lane, q.d[k], q.h[k]);
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
  if (p.variant == CTCB200_CLASSIC) k2_recursion<NS, true><<<grid, kWarp, 0, st>>>(p, s, loss, full);
  else k2_recursion<NS, false><<<grid, kWarp, 0, st>>>(p, s, loss, full);
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
#undef CTCB200_CASE
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace ctcb200
