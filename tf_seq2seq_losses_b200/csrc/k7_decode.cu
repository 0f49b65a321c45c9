// K7: greedy (best-path) CTC decoding -- the step that follows the loss at inference time.
//
// Not part of tf_seq2seq_losses itself: it is what tf.nn.ctc_greedy_decoder does to the same logits the reference's
// classic_ctc_loss (tf_seq2seq_losses/classic_ctc_loss.py:33-70) is trained on.  Per frame t < logit_length the arg-max
// token (lowest index on ties, like tf.argmax); repeated tokens are merged (merge_repeated), blanks dropped;
// neg_sum_logits = -sum_t max_k logits[b,t,k].
//   kd_argmax    one warp per logits row, 128-bit loads: HBM-bound, reads [B,T,V] once, writes 8 bytes per row
//   kd_collapse  one warp per utterance: ballot + prefix count over 32 frames at a time
#include "common.cuh"

namespace ctcb200 {

constexpr int kK7Warps = 8;

__global__ void __launch_bounds__(kK7Warps * kWarp) kd_argmax(Problem p, int* __restrict__ best, float* __restrict__ bestv) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kK7Warps + (threadIdx.x >> 5);
  if (row >= (long long)p.B * p.T) return;
  const int b = (int)(row / p.T), t = (int)(row % p.T);
  if (t >= utt_frames(p, b)) return;
  const float* x = p.logits + row_offset(p, b, t);
  float m = kNegInf;
  int arg = 0x7fffffff;
  auto take = [&](float v, int k) {
    if (v > m || (v == m && k < arg)) { m = v; arg = k; }
  };
  if (((p.V & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int i = lane; i < (p.V >> 2); i += kWarp) {
      const float4 v = ldg_stream4(x4 + i);
      take(v.x, 4 * i); take(v.y, 4 * i + 1); take(v.z, 4 * i + 2); take(v.w, 4 * i + 3);
    }
  } else {
    for (int k = lane; k < p.V; k += kWarp) take(__ldg(x + k), k);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(kFull, m, o);
    const int oa = __shfl_xor_sync(kFull, arg, o);
    if (om > m || (om == m && oa < arg)) { m = om; arg = oa; }
  }
  if (lane == 0) {
    best[row] = (arg == 0x7fffffff) ? 0 : arg;      // a row of NaNs: token 0, like an arg-max that never updates
    bestv[row] = m;
  }
}

__global__ void __launch_bounds__(kK7Warps * kWarp) kd_collapse(Problem p, const int* __restrict__ best,
                                                                 const float* __restrict__ bestv, int merge_repeated,
                                                                 int* __restrict__ decoded, int* __restrict__ decoded_length,
                                                                 float* __restrict__ neg_sum_logits) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * kK7Warps + (threadIdx.x >> 5);
  if (b >= p.B) return;
  const int n_t = utt_frames(p, b);
  const int* bb = best + (size_t)b * p.T;
  int* out = decoded + (size_t)b * p.T;
  int count = 0;
  float acc = 0.0f;
  for (int t0 = 0; t0 < n_t; t0 += kWarp) {
    const int t = t0 + lane;
    const bool in = t < n_t;
    const int tok = in ? bb[t] : p.blank;
    const int prev = (in && t > 0) ? bb[t - 1] : -1;
    const bool keep = in && tok != p.blank && !(merge_repeated && tok == prev);
    const unsigned mask = __ballot_sync(kFull, keep);
    if (keep) out[count + __popc(mask & ((1u << lane) - 1u))] = tok;
    count += __popc(mask);
    if (in) acc += bestv[(size_t)b * p.T + t];
  }
  for (int t = count + lane; t < p.T; t += kWarp) out[t] = -1;
  acc = warp_sum(acc);
  if (lane == 0) {
    decoded_length[b] = count;
    if (neg_sum_logits != nullptr) neg_sum_logits[b] = -acc;
  }
}

cudaError_t launch_greedy_decode(const Problem& p, int* best, float* bestv, int merge_repeated, int* decoded,
                                 int* decoded_length, float* neg_sum_logits, cudaStream_t st) {
  if (p.B == 0) return cudaSuccess;
  const long long rows = (long long)p.B * p.T;
  if (rows > 0) kd_argmax<<<(unsigned)((rows + kK7Warps - 1) / kK7Warps), kK7Warps * kWarp, 0, st>>>(p, best, bestv);
  kd_collapse<<<(unsigned)((p.B + kK7Warps - 1) / kK7Warps), kK7Warps * kWarp, 0, st>>>(p, best, bestv, merge_repeated, decoded,
                                                                                       decoded_length, neg_sum_logits);
  return cudaGetLastError();
}

}  // namespace ctcb200
