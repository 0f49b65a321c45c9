// Re-lays the private (lane-transposed) alpha / beta scratch rows into the reference's tensor layout:
// ClassicCtcLossData.alpha/.beta [B,T+1,U,2] (tf_seq2seq_losses/classic_ctc_loss.py:310-462) and
// SimplifiedCtcLossData.alpha/.beta [B,T+1,U] (simplified_ctc_loss.py:291-438).
#include "common.cuh"

namespace ctcb200 {

__global__ void __launch_bounds__(256)
    k_export_states(Problem p, const float* __restrict__ src, const double* __restrict__ offs, float* __restrict__ dst) {
  const size_t per_row = (size_t)p.U * p.S;
  const size_t total = (size_t)p.B * (p.T + 1) * per_row;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / per_row;
    const int rem = (int)(i % per_row);
    const int l = rem / p.S, st = rem % p.S;
    // stored rows are relative to the running offset (see Scratch); -inf + offset stays -inf
    dst[i] = (float)((double)src[(row * p.S + st) * p.Upad + state_pos(l, p.NS)] + offs[row]);
  }
}

cudaError_t launch_export_states(const Problem& p, const Scratch& s, float* alpha, float* beta, cudaStream_t st) {
  const size_t total = (size_t)p.B * (p.T + 1) * p.U * p.S;
  if (total == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (alpha) k_export_states<<<grid, 256, 0, st>>>(p, s.alphaT, s.ca, alpha);
  if (beta) k_export_states<<<grid, 256, 0, st>>>(p, s.betaT, s.cb, beta);
  return cudaGetLastError();
}

}  // namespace ctcb200
