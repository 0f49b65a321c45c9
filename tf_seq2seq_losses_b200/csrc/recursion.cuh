// Register-resident alpha / beta recursion steps shared by the staged recursion kernel (K2) and the fused kernel.
//
// A recursion warp keeps NS consecutive states per lane (state l = lane*NS + j); the l-1 / l+1 neighbour crosses
// lanes with one shuffle per frame.  Reference: simplified alpha_step simplified_ctc_loss.py:393-424, beta_step
// :327-343; classic _alpha_step classic_ctc_loss.py:415-451, beta_step :349-364 with the tables of :464-563.
//
// Classic transition algebra (s=0 closed, s=1 open; rep[l] = label[l]==label[l-1]; r[l] = d[l-1] when label[l-1]
// is not the blank):
//   A'[l,0] = h + S[l],  S[l] = lse(A[l,0], A[l,1])
//   A'[l,1] = lse(r[l] + A[l,1], d[l-1] + (rep[l-1] ? A[l-1,0] : S[l-1]))
//   B'[l,0] = lse(h + B[l,0], d[l] + B[l+1,1])
//   B'[l,1] = lse(rep[l] ? h + B[l,0] : B'[l,0], r[l] + B[l,1])
// which is the reference's [next,prev] table form with the -inf entries removed.
#pragma once
#include "common.cuh"

namespace ctcb200 {

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

// The classic alpha step in two halves, so that a caller can hand S / x of the CURRENT frame to someone else before the
// state moves on (the fused kernel publishes x and the open plane: they are all the occupancies of the frame need):
//   prepare  S[l] = lse(A[l,0], A[l,1]),  x[l] = rep[l] ? A[l,0] : S[l]      (what a diagonal step out of state l carries)
//   finish   A'[l,1] = lse(r[l] + A[l,1], d[l-1] + x[l-1]),  A'[l,0] = h + S[l]
template <int NS>
__device__ __forceinline__ void alpha_classic_prepare(const float* a0, const float* a1, const LabelBits<NS>& lb, float* S,
                                                      float* x) {
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    S[j] = lse2(a0[j], a1[j]);
    x[j] = ((lb.rep >> j) & 1u) ? a0[j] : S[j];
  }
}

template <int NS>
__device__ __forceinline__ void alpha_classic_finish(float* a0, float* a1, const float* S, const float* x, const float* d,
                                                     float h, int lane, const LabelBits<NS>& lb) {
  // d of the left neighbour's top state: pure data, off the dependency chain
  float d_left = __shfl_up_sync(kFull, d[NS - 1], 1);
  if (lane == 0) d_left = kNegInf;
  float x_left = __shfl_up_sync(kFull, x[NS - 1], 1);
  if (lane == 0) x_left = kNegInf;
#pragma unroll
  for (int j = NS - 1; j >= 1; --j) {
    const float r = ((lb.nb >> (j - 1)) & 1u) ? d[j - 1] : kNegInf;
    a1[j] = lse2(r + a1[j], d[j - 1] + x[j - 1]);
  }
  {
    const float r = lb.nb_left ? d_left : kNegInf;
    a1[0] = lse2(r + a1[0], d_left + x_left);
  }
#pragma unroll
  for (int j = 0; j < NS; ++j) a0[j] = h + S[j];
}

template <int NS>
__device__ __forceinline__ void alpha_step_classic(float* a0, float* a1, const float* d, float h, int lane,
                                                   const LabelBits<NS>& lb) {
  float S[NS], x[NS];
  alpha_classic_prepare<NS>(a0, a1, lb, S, x);
  alpha_classic_finish<NS>(a0, a1, S, x, d, h, lane, lb);
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
// kGroup-frame group and subtracted two frames later, so its five dependent shuffles overlap the next frames'
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

}  // namespace ctcb200
