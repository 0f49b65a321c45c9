// K4: batched Hessian of the loss w.r.t. the log-probabilities, H[b,t,k,t',k'] (and its matrix-free contraction).
//
// Replaces BaseCtcLossData.hessian (tf_seq2seq_losses/base_loss.py:186-260) together with the gamma recursion it
// consumes (classic_ctc_loss.py:167-308, simplified_ctc_loss.py:85-191) and gradient_fn.backprop's contraction
// (base_loss.py:167-173).  gamma ([B,T+1,U,2,T+1,U,2]) is never materialised: for a frame t and a token k the
// alpha vector is pushed through "emit k at t" and then propagated frame by frame with the ordinary alpha step;
// at every later frame t' it is combined with beta[t'+1] exactly like the gradient is (row_occupancies), which
// yields exp(loss + J[t,k,t',.]).  The lower triangle (t' < t) is produced the same way backwards in time from
// beta, so that one CTA owns the whole contiguous output slab H[b,t,:,:,:] and writes it row by row.
//     H[t,k,t',k'] = -exp(loss + J_sym) + g[t,k] g[t',k']          (t' != t)
//     H[t,k,t ,k'] = [k == k'] g[t,k]   + g[t,k] g[t ,k']          (g = -occupancy, base_loss.py:200-237)
// and zero for infeasible samples and frames beyond logit_length (base_loss.py:240-258).
#include "common.cuh"
#include "occupancy.cuh"

namespace ctcb200 {

constexpr int kK4Warps = 8;

struct ChainBufs {
  float* cur;   // [S*Upad]
  float* nxt;   // [S*Upad]
  __device__ __forceinline__ void swap() { float* t = cur; cur = nxt; nxt = t; }
};

template <bool CLASSIC>
__device__ __forceinline__ void chain_fill_neginf(const Problem& p, float* v, int lane) {
  for (int i = lane; i < (CLASSIC ? 2 : 1) * p.Upad; i += kWarp) v[i] = kNegInf;
}

// state after consuming frame t with token k emitted (A = alpha[t], global, private layout)
template <bool CLASSIC>
__device__ __forceinline__ void emit_forward(const Problem& p, int L, int lane, int k, const float* A, const float* d,
                                             float h, const int* toks, float* out) {
  for (int pos = lane; pos < p.Upad; pos += kWarp) {
    const int l = lane * p.NS + (pos >> 5);
    if (l > L) continue;
    const int pp = pos_prev(pos, lane, p.NS);
    if (k == p.blank) {
      if (CLASSIC) { out[pos] = h + lse2(A[pos], A[p.Upad + pos]); out[p.Upad + pos] = kNegInf; }
      else out[pos] = h + A[pos];
      continue;
    }
    const bool hit = (l >= 1) && (pp >= 0) && (toks[l - 1] == k);
    if (!CLASSIC) {
      out[pos] = hit ? A[pp] + d[pp] : kNegInf;
    } else {
      float v = kNegInf;
      if (hit) {
        const bool rep = toks[l - 1] == tok_at(p, toks, l - 2);
        const float x = rep ? A[pp] : lse2(A[pp], A[p.Upad + pp]);
        v = lse2(A[p.Upad + pos] + d[pp], d[pp] + x);       // stay open on label[l-1] / move from l-1
      }
      out[pos] = kNegInf;
      out[p.Upad + pos] = v;
    }
  }
}

// state weights *before* frame t given that frame t emits k and is followed by Bn = beta[t+1]
template <bool CLASSIC>
__device__ __forceinline__ void emit_backward(const Problem& p, int L, int lane, int k, const float* Bn, const float* d,
                                              float h, const int* toks, float* out) {
  for (int pos = lane; pos < p.Upad; pos += kWarp) {
    const int l = lane * p.NS + (pos >> 5);
    if (l > L) continue;
    const int pn = pos_next(pos, lane, p.NS), pp = pos_prev(pos, lane, p.NS);
    if (k == p.blank) {
      const float v = h + Bn[pos];
      out[pos] = v;
      if (CLASSIC) out[p.Upad + pos] = v;
      continue;
    }
    const bool move = (l < L) && (toks[l] == k) && (pn >= 0);
    if (!CLASSIC) {
      out[pos] = move ? d[pos] + Bn[pn] : kNegInf;
    } else {
      const float mv = move ? d[pos] + Bn[p.Upad + pn] : kNegInf;
      const bool rep = toks[l] == tok_at(p, toks, l - 1);
      const bool stay = (l >= 1) && (pp >= 0) && (toks[l - 1] == k);
      const float sv = stay ? d[pp] + Bn[p.Upad + pos] : kNegInf;
      out[pos] = mv;
      out[p.Upad + pos] = lse2(sv, rep ? kNegInf : mv);
    }
  }
}

template <bool CLASSIC>
__device__ __forceinline__ void chain_alpha_step(const Problem& p, int L, int lane, const float* cur, const float* d,
                                                 float h, const int* toks, float* nxt) {
  for (int pos = lane; pos < p.Upad; pos += kWarp) {
    const int l = lane * p.NS + (pos >> 5);
    if (l > L) continue;
    const int pp = pos_prev(pos, lane, p.NS);
    if (!CLASSIC) {
      nxt[pos] = lse2(h + cur[pos], (pp >= 0) ? d[pp] + cur[pp] : kNegInf);
    } else {
      const float c0 = cur[pos], c1 = cur[p.Upad + pos];
      nxt[pos] = h + lse2(c0, c1);
      float v = kNegInf;
      if (l >= 1 && pp >= 0) {
        const int tp = toks[l - 1];
        const bool rep = tp == tok_at(p, toks, l - 2);
        const float x = rep ? cur[pp] : lse2(cur[pp], cur[p.Upad + pp]);
        const float r = (tp != p.blank) ? d[pp] : kNegInf;
        v = lse2(r + c1, d[pp] + x);
      }
      nxt[p.Upad + pos] = v;
    }
  }
}

template <bool CLASSIC>
__device__ __forceinline__ void chain_beta_step(const Problem& p, int L, int lane, const float* cur, const float* d,
                                                float h, const int* toks, float* nxt) {
  for (int pos = lane; pos < p.Upad; pos += kWarp) {
    const int l = lane * p.NS + (pos >> 5);
    if (l > L) continue;
    const int pn = pos_next(pos, lane, p.NS), pp = pos_prev(pos, lane, p.NS);
    if (!CLASSIC) {
      nxt[pos] = lse2(h + cur[pos], d[pos] + ((pn >= 0) ? cur[pn] : kNegInf));
    } else {
      const float stay = h + cur[pos];
      const float nx = (pn >= 0) ? cur[p.Upad + pn] : kNegInf;
      const float n0 = lse2(stay, d[pos] + nx);
      const bool rep = toks[l] == tok_at(p, toks, l - 1);
      const float r = (l >= 1 && pp >= 0 && toks[l - 1] != p.blank) ? d[pp] : kNegInf;
      nxt[pos] = n0;
      nxt[p.Upad + pos] = lse2(rep ? stay : n0, r + cur[p.Upad + pos]);
    }
  }
}

// one output row H[b,t,k,t2,:] (dense) or its dot product with d_gradient[b,t2,:] (hvp)
template <bool HVP>
__device__ __forceinline__ float emit_row(const Problem& p, int lane, float gk, const float* g2, const float* acc,
                                          const unsigned short* map, int diag_k, float* hrow, const float* dg2) {
  float partial = 0.0f;
  for (int k2 = lane; k2 < p.V; k2 += kWarp) {
    float v = gk * g2[k2];
    if (acc != nullptr) {
      const unsigned short m = map[k2];
      if (m != kNoSlot) v -= acc[m];
    }
    if (k2 == diag_k) v += gk;
    if (HVP) partial += v * dg2[k2];
    else hrow[k2] = v;
  }
  return partial;
}

template <bool CLASSIC, bool HVP>
__global__ void __launch_bounds__(kK4Warps * kWarp)
    k4_hessian(Problem p, Scratch s, const float* __restrict__ g, float* __restrict__ hessian,
               const float* __restrict__ d_gradient, float* __restrict__ hvp_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int S = CLASSIC ? 2 : 1;
  const int Vpad = (p.V + 7) & ~7;
  unsigned short* map = reinterpret_cast<unsigned short*>(smem_raw);
  int* toks = reinterpret_cast<int*>(smem_raw + (size_t)Vpad * sizeof(unsigned short));
  float* warp_base = reinterpret_cast<float*>(toks + p.Upad);
  const int per_warp = (p.Upad + kWarp) + 2 * S * p.Upad;

  const int b = blockIdx.x / p.T, t = blockIdx.x % p.T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = utt_label_len(p, b), n_t = utt_frames(p, b);
  const double lossd = s.lossd[b];
  const bool dead = (lossd == (double)INFINITY) || (t >= n_t);
  const size_t TV = (size_t)p.T * p.V;
  float* slab = HVP ? nullptr : hessian + ((size_t)b * p.T + t) * p.V * TV;

  if (dead) {
    if (HVP) {
      for (int k = tid; k < p.V; k += blockDim.x) hvp_out[((size_t)b * p.T + t) * p.V + k] = 0.0f;
    } else {
      for (size_t i = tid; i < (size_t)p.V * TV; i += blockDim.x) slab[i] = 0.0f;
    }
    return;
  }
  build_utterance_tables(p, b, L, toks, map, Vpad);

  float* acc = warp_base + (size_t)warp * per_warp;
  ChainBufs cb;
  cb.cur = acc + p.Upad + kWarp;
  cb.nxt = cb.cur + S * p.Upad;
  const size_t row_pitch = (size_t)S * p.Upad;
  const float* alpha_b = s.alphaT + (size_t)b * (p.T + 1) * row_pitch;
  const float* beta_b = s.betaT + (size_t)b * (p.T + 1) * row_pitch;
  const float* d_b = s.dT + (size_t)b * p.T * p.Upad;
  const float* h_b = s.h + (size_t)b * p.T;
  const float* g_b = g + (size_t)b * TV;
  const float* dg_b = HVP ? d_gradient + (size_t)b * TV : nullptr;
  const double* ca_b = s.ca + (size_t)b * (p.T + 1);
  const double* cb_b = s.cb + (size_t)b * (p.T + 1);

  for (int k = warp; k < p.V; k += kK4Warps) {
    float* hk = HVP ? nullptr : slab + (size_t)k * TV;
    const unsigned short slot = map[k];
    if (slot == kNoSlot) {       // token not in the label: g[t,k] == 0 and every path term vanishes
      if (HVP) { if (lane == 0) hvp_out[((size_t)b * p.T + t) * p.V + k] = 0.0f; }
      else for (size_t i = lane; i < TV; i += kWarp) hk[i] = 0.0f;
      continue;
    }
    const float gk = g_b[(size_t)t * p.V + k];
    float hv = 0.0f;
    // frames beyond logit_length
    if (!HVP) for (size_t i = (size_t)n_t * p.V + lane; i < TV; i += kWarp) hk[i] = 0.0f;
    // same frame
    hv += emit_row<HVP>(p, lane, gk, g_b + (size_t)t * p.V, nullptr, map, k, HVP ? nullptr : hk + (size_t)t * p.V,
                        HVP ? dg_b + (size_t)t * p.V : nullptr);
    // ---- later frames: push alpha[t] through "emit k", propagate forward ----
    chain_fill_neginf<CLASSIC>(p, cb.cur, lane);
    chain_fill_neginf<CLASSIC>(p, cb.nxt, lane);
    __syncwarp();
    emit_forward<CLASSIC>(p, L, lane, k, alpha_b + (size_t)t * row_pitch, d_b + (size_t)t * p.Upad, h_b[t], toks, cb.cur);
    __syncwarp();
    for (int t2 = t + 1; t2 < n_t; ++t2) {
      const float* d2 = d_b + (size_t)t2 * p.Upad;
      const float lossb = (float)(lossd + ca_b[t] + cb_b[t2 + 1]);   // chain starts from alpha[t]'s offset
      row_occupancies<CLASSIC>(p, L, lane, cb.cur, beta_b + (size_t)(t2 + 1) * row_pitch, d2, h_b[t2], lossb, toks, map, acc);
      hv += emit_row<HVP>(p, lane, gk, g_b + (size_t)t2 * p.V, acc, map, -1, HVP ? nullptr : hk + (size_t)t2 * p.V,
                          HVP ? dg_b + (size_t)t2 * p.V : nullptr);
      chain_alpha_step<CLASSIC>(p, L, lane, cb.cur, d2, h_b[t2], toks, cb.nxt);
      __syncwarp();
      cb.swap();
    }
    // ---- earlier frames: pull beta[t+1] back through "emit k", propagate backward ----
    chain_fill_neginf<CLASSIC>(p, cb.cur, lane);
    chain_fill_neginf<CLASSIC>(p, cb.nxt, lane);
    __syncwarp();
    emit_backward<CLASSIC>(p, L, lane, k, beta_b + (size_t)(t + 1) * row_pitch, d_b + (size_t)t * p.Upad, h_b[t], toks, cb.cur);
    __syncwarp();
    for (int t2 = t - 1; t2 >= 0; --t2) {
      const float* d2 = d_b + (size_t)t2 * p.Upad;
      const float lossb = (float)(lossd + ca_b[t2] + cb_b[t + 1]);   // chain starts from beta[t+1]'s offset
      row_occupancies<CLASSIC>(p, L, lane, alpha_b + (size_t)t2 * row_pitch, cb.cur, d2, h_b[t2], lossb, toks, map, acc);
      hv += emit_row<HVP>(p, lane, gk, g_b + (size_t)t2 * p.V, acc, map, -1, HVP ? nullptr : hk + (size_t)t2 * p.V,
                          HVP ? dg_b + (size_t)t2 * p.V : nullptr);
      chain_beta_step<CLASSIC>(p, L, lane, cb.cur, d2, h_b[t2], toks, cb.nxt);
      __syncwarp();
      cb.swap();
    }
    if (HVP) {
      hv = warp_sum(hv);
      if (lane == 0) hvp_out[((size_t)b * p.T + t) * p.V + k] = hv;
    }
  }
}

static size_t hessian_smem_bytes(const Problem& p) {
  const int Vpad = (p.V + 7) & ~7;
  const int per_warp = (p.Upad + kWarp) + 2 * p.S * p.Upad;
  return (size_t)Vpad * sizeof(unsigned short) + (size_t)p.Upad * sizeof(int) + (size_t)kK4Warps * per_warp * sizeof(float);
}

template <bool CLASSIC, bool HVP>
static cudaError_t launch_k4(const Problem& p, const Scratch& s, const float* g, float* hessian,
                             const float* d_gradient, float* hvp_out, cudaStream_t st) {
  const size_t smem = hessian_smem_bytes(p);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k4_hessian<CLASSIC, HVP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const unsigned grid = (unsigned)((long long)p.B * p.T);
  k4_hessian<CLASSIC, HVP><<<grid, kK4Warps * kWarp, smem, st>>>(p, s, g, hessian, d_gradient, hvp_out);
  return cudaGetLastError();
}

// g = d loss / d logproba [B,T,V] (K3's grad_logprobas output with d_loss == 1)
cudaError_t launch_hessian(const Problem& p, const Scratch& s, const float* g, float* hessian,
                           const float* d_gradient, float* hvp_out, cudaStream_t st) {
  if (p.B == 0 || p.T == 0) return cudaSuccess;
  const bool classic = p.variant == CTCB200_CLASSIC;
  if (hessian != nullptr) {
    cudaError_t e = classic ? launch_k4<true, false>(p, s, g, hessian, nullptr, nullptr, st)
                            : launch_k4<false, false>(p, s, g, hessian, nullptr, nullptr, st);
    if (e != cudaSuccess) return e;
  }
  if (hvp_out != nullptr) {
    return classic ? launch_k4<true, true>(p, s, g, nullptr, d_gradient, hvp_out, st)
                   : launch_k4<false, true>(p, s, g, nullptr, d_gradient, hvp_out, st);
  }
  return cudaSuccess;
}

}  // namespace ctcb200
