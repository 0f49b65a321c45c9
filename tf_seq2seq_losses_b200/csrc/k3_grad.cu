// K3: gradient writer.  Combines alpha[t] * beta[t+1] per (frame, token) and writes the dense gradient rows.
//
// Replaces _combine_transition_probabilities (tf_seq2seq_losses/classic_ctc_loss.py:565-669,
// simplified_ctc_loss.py:456-534) with its _select_from_act token scatter (base_loss.py:420-468,
// tools.py:95-119), logarithmic_logproba_gradient / gradient (base_loss.py:262-298), forward_fn.backprop
// (base_loss.py:150-153) and the TF autodiff of the log-softmax (tools.py:37-39):
//     occ[t,k]          = exp(loss + c[t,k])                              (= -gradient)
//     grad_logprobas    = -d_loss * occ
//     grad_logits[t,k]  = d_loss * (softmax[t,k] * sum_k' occ[t,k'] - occ[t,k])
// One CTA = k3_rows(p) consecutive frames of one utterance.  The CTA builds the utterance's token -> slot map once in
// shared memory (slot = a label position carrying that token; blank has its own slot), each warp then owns
// whole frames: it accumulates the <= U+1 per-state occupancies into its slot array with shared-memory atomics
// and streams the logits row once (128-bit loads / stores) to produce the dense output row.
#include "common.cuh"
#include "occupancy.cuh"

namespace ctcb200 {

constexpr int kK3Warps = 8;
// Frames per CTA.  The per-utterance tables are rebuilt by every CTA, so narrow vocabularies (cheap rows) take longer
// tiles -- as long as that still leaves four CTAs for each of the 148 SMs.
__host__ __device__ inline int k3_rows(const Problem& p) {
  if (p.V >= 512) return 16;
  for (int rows = 64; rows > 16; rows >>= 1)
    if ((long long)p.B * ((p.T + rows - 1) / rows) >= 4 * 148) return rows;
  return 16;
}
__device__ __forceinline__ void zero_row(float* dst, int V, int lane) {
  if (dst == nullptr) return;
  if (((V & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = lane; i < (V >> 2); i += kWarp) stg_stream4(d4 + i, z);
  } else {
    for (int i = lane; i < V; i += kWarp) dst[i] = 0.0f;
  }
}

template <int NS, bool CLASSIC>
__global__ void __launch_bounds__(kK3Warps * kWarp)
    k3_grad(Problem p, Scratch s, const float* __restrict__ d_loss, float* __restrict__ grad_logits,
            float* __restrict__ grad_logprobas) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Vpad = (p.V + 7) & ~7;
  unsigned short* map = reinterpret_cast<unsigned short*>(smem_raw);
  int* toks = reinterpret_cast<int*>(smem_raw + (size_t)Vpad * sizeof(unsigned short));
  float* acc_all = reinterpret_cast<float*>(toks + p.Upad);
  const int acc_pitch = p.Upad + kWarp;     // slots 0..Upad-1 = label positions, slot Upad = blank

  const int kK3Rows = k3_rows(p);
  const int tiles = (p.T + kK3Rows - 1) / kK3Rows;
  const int b = blockIdx.x / tiles, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t_begin = (blockIdx.x % tiles) * kK3Rows, t_end = min(p.T, t_begin + kK3Rows);
  const int L = utt_label_len(p, b), n_t = utt_frames(p, b);
  const double lossd = s.lossd[b];
  const float dl = d_loss ? d_loss[b] : 1.0f;
  const bool dead = (lossd == (double)INFINITY) || (t_begin >= n_t);
  if (!dead) build_utterance_tables(p, b, L, toks, map, Vpad);

  const int S = CLASSIC ? 2 : 1;
  const size_t row_pitch = (size_t)S * p.Upad;
  float* acc = acc_all + (size_t)warp * acc_pitch;

  for (int t = t_begin + warp; t < t_end; t += kK3Warps) {
    const size_t row = (size_t)b * p.T + t;
    float* gl = grad_logits ? grad_logits + row_offset(p, b, t) : nullptr;
    float* gp = grad_logprobas ? grad_logprobas + row_offset(p, b, t) : nullptr;
    if (dead || t >= n_t) {          // frames beyond logit_length and infeasible samples: exact zeros
      zero_row(gl, p.V, lane);
      zero_row(gp, p.V, lane);
      continue;
    }
    const float* A = s.alphaT + ((size_t)b * (p.T + 1) + t) * row_pitch;
    const float* Bn = s.betaT + ((size_t)b * (p.T + 1) + t + 1) * row_pitch;
    // loss + offset of alpha[t] + offset of beta[t+1]: a small number, formed in double
    const size_t crow = (size_t)b * (p.T + 1) + t;
    const float lossb = (float)(lossd + s.ca[crow] + s.cb[crow + 1]);
    const float occ_sum = row_occupancies_t<NS, CLASSIC>(p, L, lane, A, Bn, s.dT + row * p.Upad, s.h[row], lossb, toks,
                                                         map, acc);

    // dense row
    const float* x = p.logits + row_offset(p, b, t);
    const float lse_row = s.rowlse[row];
    const float scale = dl * occ_sum;
    const bool vec = ((p.V & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                     (gl == nullptr || (reinterpret_cast<uintptr_t>(gl) & 15) == 0) &&
                     (gp == nullptr || (reinterpret_cast<uintptr_t>(gp) & 15) == 0);
    if (vec) {
      const float4* x4 = reinterpret_cast<const float4*>(x);
      const ushort4* m4 = reinterpret_cast<const ushort4*>(map);
      const int n4 = p.V >> 2;
      // kUnroll independent 128-bit loads are issued before any of them is consumed (memory-level parallelism)
      constexpr int kUnroll = 4;
      for (int i0 = 0; i0 < n4; i0 += kUnroll * kWarp) {
        float4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int i = i0 + u * kWarp + lane;
          if (gl && i < n4) v[u] = ldg_stream4(x4 + i);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int i = i0 + u * kWarp + lane;
          if (i >= n4) continue;
          const ushort4 m = m4[i];
          float4 o;
          o.x = (m.x != kNoSlot) ? acc[m.x] : 0.0f;
          o.y = (m.y != kNoSlot) ? acc[m.y] : 0.0f;
          o.z = (m.z != kNoSlot) ? acc[m.z] : 0.0f;
          o.w = (m.w != kNoSlot) ? acc[m.w] : 0.0f;
          if (gl) {
            float4 g;
            g.x = scale * __expf(v[u].x - lse_row) - dl * o.x;
            g.y = scale * __expf(v[u].y - lse_row) - dl * o.y;
            g.z = scale * __expf(v[u].z - lse_row) - dl * o.z;
            g.w = scale * __expf(v[u].w - lse_row) - dl * o.w;
            stg_stream4(reinterpret_cast<float4*>(gl) + i, g);
          }
          if (gp) stg_stream4(reinterpret_cast<float4*>(gp) + i, make_float4(-dl * o.x, -dl * o.y, -dl * o.z, -dl * o.w));
        }
      }
    } else {
      for (int k = lane; k < p.V; k += kWarp) {
        const unsigned short m = map[k];
        const float o = (m != kNoSlot) ? acc[m] : 0.0f;
        if (gl) gl[k] = scale * __expf(__ldg(x + k) - lse_row) - dl * o;
        if (gp) gp[k] = -dl * o;
      }
    }
    __syncwarp();     // acc is reused by this warp's next frame
  }
}

// K3-log: the same frames in the log domain -> logarithmic_logproba_gradient (base_loss.py:270-298): log of the
// occupancy per (frame, token), -inf where impossible, for frames beyond logit_length and for infeasible samples.
template <int NS, bool CLASSIC>
__global__ void __launch_bounds__(kK3Warps * kWarp)
    k3_log_grad(Problem p, Scratch s, float* __restrict__ log_grad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Vpad = (p.V + 7) & ~7;
  unsigned short* map = reinterpret_cast<unsigned short*>(smem_raw);
  int* toks = reinterpret_cast<int*>(smem_raw + (size_t)Vpad * sizeof(unsigned short));
  int* mx_all = toks + p.Upad;
  float* sm_all = reinterpret_cast<float*>(mx_all + kK3Warps * p.Upad);

  const int kK3Rows = k3_rows(p);
  const int tiles = (p.T + kK3Rows - 1) / kK3Rows;
  const int b = blockIdx.x / tiles, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t_begin = (blockIdx.x % tiles) * kK3Rows, t_end = min(p.T, t_begin + kK3Rows);
  const int L = utt_label_len(p, b), n_t = utt_frames(p, b);
  const double lossd = s.lossd[b];
  const bool dead = (lossd == (double)INFINITY) || (t_begin >= n_t);
  if (!dead) build_utterance_tables(p, b, L, toks, map, Vpad);
  const size_t row_pitch = (size_t)(CLASSIC ? 2 : 1) * p.Upad;
  int* mx = mx_all + (size_t)warp * p.Upad;
  float* sm = sm_all + (size_t)warp * p.Upad;
  for (int t = t_begin + warp; t < t_end; t += kK3Warps) {
    const size_t row = (size_t)b * p.T + t;
    float* out = log_grad + row_offset(p, b, t);
    if (dead || t >= n_t) {
      for (int k = lane; k < p.V; k += kWarp) out[k] = kNegInf;
      continue;
    }
    const float* A = s.alphaT + ((size_t)b * (p.T + 1) + t) * row_pitch;
    const float* Bn = s.betaT + ((size_t)b * (p.T + 1) + t + 1) * row_pitch;
    const size_t crow = (size_t)b * (p.T + 1) + t;
    const double lossb = lossd + s.ca[crow] + s.cb[crow + 1];
    const float lg_blank = row_log_occupancies_t<NS, CLASSIC>(p, L, lane, A, Bn, s.dT + row * p.Upad, s.h[row], lossb, toks,
                                                              map, mx, sm);
    for (int k = lane; k < p.V; k += kWarp) {
      const unsigned short m = map[k];
      float v = kNegInf;
      if (k == p.blank) v = lg_blank;
      else if (m != kNoSlot && sm[m] > 0.0f) v = (float)(lossb + (double)float_unorder(mx[m])) + __logf(sm[m]);
      out[k] = v;
    }
    __syncwarp();
  }
}

size_t log_grad_smem_bytes(const Problem& p) {
  const int Vpad = (p.V + 7) & ~7;
  return (size_t)Vpad * sizeof(unsigned short) + (size_t)p.Upad * sizeof(int) + (size_t)kK3Warps * p.Upad * 8;
}

template <int NS>
static cudaError_t launch_k3_log_ns(const Problem& p, const Scratch& s, float* log_grad, cudaStream_t st) {
  const size_t smem = log_grad_smem_bytes(p);
  const int kK3Rows = k3_rows(p);
  const unsigned grid = (unsigned)(((p.T + kK3Rows - 1) / kK3Rows) * (long long)p.B);
  cudaError_t e;
  if (p.variant == CTCB200_CLASSIC) {
    if (smem > 48 * 1024 &&
        (e = cudaFuncSetAttribute(k3_log_grad<NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return e;
    k3_log_grad<NS, true><<<grid, kK3Warps * kWarp, smem, st>>>(p, s, log_grad);
  } else {
    if (smem > 48 * 1024 &&
        (e = cudaFuncSetAttribute(k3_log_grad<NS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return e;
    k3_log_grad<NS, false><<<grid, kK3Warps * kWarp, smem, st>>>(p, s, log_grad);
  }
  return cudaGetLastError();
}

cudaError_t launch_log_grad(const Problem& p, const Scratch& s, float* log_grad, cudaStream_t st) {
  if (p.B == 0 || p.T == 0 || log_grad == nullptr) return cudaSuccess;
  switch (p.NS) {
#define CTCB200_CASE(n) \
  case n:               \
    return launch_k3_log_ns<n>(p, s, log_grad, st);
    CTCB200_CASE(1) CTCB200_CASE(2) CTCB200_CASE(3) CTCB200_CASE(4) CTCB200_CASE(5) CTCB200_CASE(6)
    CTCB200_CASE(7) CTCB200_CASE(8) CTCB200_CASE(9) CTCB200_CASE(10) CTCB200_CASE(11) CTCB200_CASE(12)
    CTCB200_CASE(13) CTCB200_CASE(14) CTCB200_CASE(15) CTCB200_CASE(16)
    CTCB200_CASE(20) CTCB200_CASE(24) CTCB200_CASE(28) CTCB200_CASE(32)      // U > 512: make_problem rounds NS up
#undef CTCB200_CASE
    default:
      return cudaErrorInvalidValue;
  }
}

size_t grad_smem_bytes(const Problem& p) {
  const int Vpad = (p.V + 7) & ~7;
  return (size_t)Vpad * sizeof(unsigned short) + (size_t)p.Upad * sizeof(int) +
         (size_t)kK3Warps * (p.Upad + kWarp) * sizeof(float);
}

template <int NS>
static cudaError_t launch_k3_ns(const Problem& p, const Scratch& s, const float* d_loss, float* grad_logits,
                                float* grad_logprobas, cudaStream_t st) {
  const size_t smem = grad_smem_bytes(p);
  const int kK3Rows = k3_rows(p);
  const unsigned grid = (unsigned)(((p.T + kK3Rows - 1) / kK3Rows) * (long long)p.B);
  cudaError_t e;
  if (p.variant == CTCB200_CLASSIC) {
    if (smem > 48 * 1024 &&
        (e = cudaFuncSetAttribute(k3_grad<NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return e;
    k3_grad<NS, true><<<grid, kK3Warps * kWarp, smem, st>>>(p, s, d_loss, grad_logits, grad_logprobas);
  } else {
    if (smem > 48 * 1024 &&
        (e = cudaFuncSetAttribute(k3_grad<NS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return e;
    k3_grad<NS, false><<<grid, kK3Warps * kWarp, smem, st>>>(p, s, d_loss, grad_logits, grad_logprobas);
  }
  return cudaGetLastError();
}

cudaError_t launch_grad(const Problem& p, const Scratch& s, const float* d_loss, float* grad_logits,
                        float* grad_logprobas, cudaStream_t st) {
  if (p.B == 0 || p.T == 0 || (grad_logits == nullptr && grad_logprobas == nullptr)) return cudaSuccess;
  switch (p.NS) {
#define CTCB200_CASE(n) \
  case n:               \
    return launch_k3_ns<n>(p, s, d_loss, grad_logits, grad_logprobas, st);
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
