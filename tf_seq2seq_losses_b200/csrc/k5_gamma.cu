// K5: the gamma tensor -- log-probability of going from state sigma1 at time t1 to state sigma2 at time t2.
//
// Replaces ClassicCtcLossData.gamma / gamma_step / diagonal_gamma (tf_seq2seq_losses/classic_ctc_loss.py:167-308,
// layout [B,T+1,U,2,T+1,U,2]) and SimplifiedCtcLossData.gamma (simplified_ctc_loss.py:85-191,279-289, layout
// [B,T+1,U,T+1,U]).  The reference unfolds it with a tf.while_loop over slices of [B,T+1,U,2,U,2]; here one warp owns
// one source (b, t1, sigma1): it starts from the unit vector at sigma1 (diagonal_gamma), applies the ordinary alpha
// step frame by frame (recursion.cuh) and writes the row gamma[b,t1,sigma1,t2,:] for every t2 >= t1; rows t2 < t1 are
// -inf (the band mask of classic_ctc_loss.py:204-213).  The tensor is O(T^2 U^2): only small shapes are practical,
// and neither the gradient nor the Hessian kernels need it -- it exists to complete the data-class surface.
#include "common.cuh"
#include "recursion.cuh"

namespace ctcb200 {

constexpr int kK5Warps = 4;

template <int NS, bool CLASSIC>
__global__ void __launch_bounds__(kK5Warps * kWarp) k5_gamma(Problem p, Scratch s, float* __restrict__ gamma) {
  constexpr int S = CLASSIC ? 2 : 1, kUpad = NS * kWarp;
  const int lane = threadIdx.x & 31;
  const long long src = (long long)blockIdx.x * kK5Warps + (threadIdx.x >> 5);     // (b, t1, l1, s1)
  const long long per_b = (long long)(p.T + 1) * p.U * S;
  if (src >= (long long)p.B * per_b) return;
  const int b = (int)(src / per_b);
  int rem = (int)(src % per_b);
  const int t1 = rem / (p.U * S);
  rem %= p.U * S;
  const int l1 = rem / S, s1 = rem % S;
  const int L = utt_label_len(p, b), n_t = utt_frames(p, b);
  LabelBits<NS> lb;
  if (CLASSIC) lb = make_label_bits<NS>(p, b, L, lane);
  float v0[NS], v1[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const bool here = (lane * NS + j == l1);
    v0[j] = (here && s1 == 0) ? 0.0f : kNegInf;
    v1[j] = (here && s1 == 1) ? 0.0f : kNegInf;
  }
  const size_t row_len = (size_t)p.U * S;                                   // one (t2) row of the output
  float* out = gamma + (size_t)src * (p.T + 1) * row_len;
  for (int t2 = 0; t2 <= p.T; ++t2) {
    float* o = out + (size_t)t2 * row_len;
    if (t2 < t1) {
      for (int q = lane; q < (int)row_len; q += kWarp) o[q] = kNegInf;
      continue;
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int l = lane * NS + j;
      if (l < p.U) {
        o[(size_t)l * S] = v0[j];
        if (CLASSIC) o[(size_t)l * S + 1] = v1[j];
      }
    }
    if (t2 == p.T) break;
    float d[NS], h;
    if (t2 < n_t) {
      const float* dsrc = s.dT + ((size_t)b * p.T + t2) * kUpad;
#pragma unroll
      for (int j = 0; j < NS; ++j) d[j] = dsrc[j * kWarp + lane];
      h = s.h[(size_t)b * p.T + t2];
    } else {            // frames beyond logit_length: blank with probability one (base_loss.py:378-393)
#pragma unroll
      for (int j = 0; j < NS; ++j) d[j] = kNegInf;
      h = 0.0f;
    }
    if (CLASSIC) alpha_step_classic<NS>(v0, v1, d, h, lane, lb);
    else alpha_step_simplified<NS>(v0, d, h, lane);
  }
}

template <int NS>
static cudaError_t launch_k5_ns(const Problem& p, const Scratch& s, float* gamma, cudaStream_t st) {
  const long long sources = (long long)p.B * (p.T + 1) * p.U * p.S;
  const unsigned grid = (unsigned)((sources + kK5Warps - 1) / kK5Warps);
  if (p.variant == CTCB200_CLASSIC) k5_gamma<NS, true><<<grid, kK5Warps * kWarp, 0, st>>>(p, s, gamma);
  else k5_gamma<NS, false><<<grid, kK5Warps * kWarp, 0, st>>>(p, s, gamma);
  return cudaGetLastError();
}

cudaError_t launch_gamma(const Problem& p, const Scratch& s, float* gamma, cudaStream_t st) {
  if (p.B == 0) return cudaSuccess;
  switch (p.NS) {
    case 1: return launch_k5_ns<1>(p, s, gamma, st);
    case 2: return launch_k5_ns<2>(p, s, gamma, st);
    case 3: return launch_k5_ns<3>(p, s, gamma, st);
    case 4: return launch_k5_ns<4>(p, s, gamma, st);
    default: return cudaErrorInvalidValue;      // U > 128: a [T+1,U,2,T+1,U,2] tensor is out of reach anyway
  }
}

}  // namespace ctcb200
