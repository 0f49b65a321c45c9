// K6: the log-softmax chain around the matrix-free Hessian-vector product, so that the second-order backward w.r.t. the
// LOGITS is one C-ABI call.
//
// The reference gets this by TF autodiff: tape.gradient of gradient_fn.backprop (tf_seq2seq_losses/base_loss.py:167-173)
// chained through logit_to_logproba (tools.py:27-40).  With p = softmax(logits), g = d loss / d logproba, H = d2 loss /
// d logproba2, s_t = sum_k g[t,k] and J = d logproba / d logits = I - 1 p^T per frame:
//   (d2 loss / d logits2) v = J^T H J v - s * (p.v - p (p^T v))
// `hvp_pre` forms w = J v = v - p^T v, K4<HVP> contracts H with w, `hvp_post` applies J^T and the softmax curvature term.
// One warp per row; rows at or beyond logit_length are zero.
#include "common.cuh"

namespace ctcb200 {

constexpr int kK6Warps = 8;

__global__ void __launch_bounds__(kK6Warps * kWarp) k6_hvp_pre(Problem p, Scratch s, const float* __restrict__ v,
                                                                float* __restrict__ w, float* __restrict__ pv) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kK6Warps + (threadIdx.x >> 5);
  if (row >= (long long)p.B * p.T) return;
  const int b = (int)(row / p.T), t = (int)(row % p.T);
  const float* x = p.logits + (size_t)row * p.V;
  const float* vr = v + (size_t)row * p.V;
  float* wr = w + (size_t)row * p.V;
  if (t >= utt_frames(p, b)) {
    for (int k = lane; k < p.V; k += kWarp) wr[k] = 0.0f;
    if (lane == 0) pv[row] = 0.0f;
    return;
  }
  const float lse = s.rowlse[row];
  float acc = 0.0f;
  for (int k = lane; k < p.V; k += kWarp) acc += __expf(__ldg(x + k) - lse) * __ldg(vr + k);
  acc = warp_sum(acc);
  for (int k = lane; k < p.V; k += kWarp) wr[k] = __ldg(vr + k) - acc;
  if (lane == 0) pv[row] = acc;
}

__global__ void __launch_bounds__(kK6Warps * kWarp) k6_hvp_post(Problem p, Scratch s, const float* __restrict__ v,
                                                                 const float* __restrict__ y, const float* __restrict__ g,
                                                                 const float* __restrict__ pv,
                                                                 const float* __restrict__ d_loss, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kK6Warps + (threadIdx.x >> 5);
  if (row >= (long long)p.B * p.T) return;
  const int b = (int)(row / p.T), t = (int)(row % p.T);
  float* o = out + (size_t)row * p.V;
  if (t >= utt_frames(p, b)) {
    for (int k = lane; k < p.V; k += kWarp) o[k] = 0.0f;
    return;
  }
  const float* x = p.logits + (size_t)row * p.V;
  const float* vr = v + (size_t)row * p.V;
  const float* yr = y + (size_t)row * p.V;
  const float* gr = g + (size_t)row * p.V;
  float sy = 0.0f, sg = 0.0f;
  for (int k = lane; k < p.V; k += kWarp) {
    sy += __ldg(yr + k);
    sg += __ldg(gr + k);
  }
  sy = warp_sum(sy);
  sg = warp_sum(sg);
  const float lse = s.rowlse[row], pvr = pv[row], dl = d_loss ? d_loss[b] : 1.0f;
  for (int k = lane; k < p.V; k += kWarp) {
    const float pk = __expf(__ldg(x + k) - lse);
    o[k] = dl * (__ldg(yr + k) - pk * sy - sg * (pk * __ldg(vr + k) - pk * pvr));
  }
}

cudaError_t launch_hvp_pre(const Problem& p, const Scratch& s, const float* v, float* w, float* pv, cudaStream_t st) {
  const long long rows = (long long)p.B * p.T;
  if (rows == 0) return cudaSuccess;
  k6_hvp_pre<<<(unsigned)((rows + kK6Warps - 1) / kK6Warps), kK6Warps * kWarp, 0, st>>>(p, s, v, w, pv);
  return cudaGetLastError();
}

cudaError_t launch_hvp_post(const Problem& p, const Scratch& s, const float* v, const float* y, const float* g,
                            const float* pv, const float* d_loss, float* out, cudaStream_t st) {
  const long long rows = (long long)p.B * p.T;
  if (rows == 0) return cudaSuccess;
  k6_hvp_post<<<(unsigned)((rows + kK6Warps - 1) / kK6Warps), kK6Warps * kWarp, 0, st>>>(p, s, v, y, g, pv, d_loss, out);
  return cudaGetLastError();
}

}  // namespace ctcb200
