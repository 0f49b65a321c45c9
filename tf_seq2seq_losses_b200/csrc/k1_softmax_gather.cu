// K1: fused log-softmax statistics + label-state gather.
//
// Replaces logit_to_logproba (tf_seq2seq_losses/tools.py:27-40), the logit-length masking of the log-probabilities
// (base_loss.py:378-393), the label cleaning (base_loss.py:395-418) and the per-utterance gathers
// _expected_token_logproba / _blank_logproba (base_loss.py:328-371).  One warp owns one logits row: it reads the row
// once from HBM with 128-bit loads (the second pass and the <= U gathers hit L1), and emits
//   rowlse[b,t] = logsumexp_k logits[b,t,k]
//   h[b,t]      = logits[b,t,blank] - rowlse
//   dT[b,t,pos(l)] = logits[b,t,label[b,l]] - rowlse   for l < label_length[b], else -inf
// Rows t >= logit_length[b] are never touched: the recursion treats them as "blank with probability one".
#include "common.cuh"

namespace ctcb200 {

constexpr int kK1Warps = 8;

__global__ void __launch_bounds__(kK1Warps * kWarp) k1_softmax_gather(Problem p, Scratch s) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kK1Warps + (threadIdx.x >> 5);
  if (row >= (long long)p.B * p.T) return;
  const int b = (int)(row / p.T), t = (int)(row % p.T);
  if (t >= utt_frames(p, b)) return;
  const int L = utt_label_len(p, b);
  const float* x = p.logits + row_offset(p, b, t);

  float lse = 0.0f;
  if (!p.input_logprobas) {
    const bool vec = ((p.V & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    float m = kNegInf;
    if (vec) {
      const float4* x4 = reinterpret_cast<const float4*>(x);
      for (int i = lane; i < (p.V >> 2); i += kWarp) {
        float4 v = __ldg(x4 + i);
        m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
      }
    } else {
      for (int i = lane; i < p.V; i += kWarp) m = fmaxf(m, __ldg(x + i));
    }
    m = warp_max(m);
    // tf.reduce_logsumexp replaces a non-finite max by 0
    const float m0 = (m == kNegInf || m == INFINITY) ? 0.0f : m;
    float sum = 0.0f;
    if (vec) {
      const float4* x4 = reinterpret_cast<const float4*>(x);
      for (int i = lane; i < (p.V >> 2); i += kWarp) {
        float4 v = __ldg(x4 + i);
        sum += (__expf(v.x - m0) + __expf(v.y - m0)) + (__expf(v.z - m0) + __expf(v.w - m0));
      }
    } else {
      for (int i = lane; i < p.V; i += kWarp) sum += __expf(__ldg(x + i) - m0);
    }
    sum = warp_sum(sum);
    lse = m0 + logf(sum);   // once per row: full-precision log
  }
  if (lane == 0) {
    s.rowlse[row] = lse;
    s.h[row] = __ldg(x + p.blank) - lse;
  }
  float* drow = s.dT + (size_t)row * p.Upad;
  for (int pos = lane; pos < p.Upad; pos += kWarp) {
    const int l = lane * p.NS + (pos >> 5);
    float v = kNegInf;
    if (l < L) {
      const int tok = utt_token(p, b, l, L);
      if (tok >= 0 && tok < p.V) v = __ldg(x + tok) - lse;
    }
    drow[pos] = v;
  }
}

cudaError_t launch_softmax_gather(const Problem& p, const Scratch& s, cudaStream_t st) {
  const long long rows = (long long)p.B * p.T;
  if (rows == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((rows + kK1Warps - 1) / kK1Warps);
  k1_softmax_gather<<<grid, kK1Warps * kWarp, 0, st>>>(p, s);
  return cudaGetLastError();
}

}  // namespace ctcb200
