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
#include <cstdlib>

#include "common.cuh"
#include "occupancy.cuh"
#include "recursion.cuh"

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

// ----------------------------------------------------------------------------------------------------------------------
// Register-resident variant for U <= 128 (NS <= 4), which covers every shape for which the O(T^2 V^2) output is
// affordable.  Same algorithm; the chain state lives in registers (recursion.cuh steps, neighbour by shuffle), the
// occupancies of a frame are computed from registers and subtracted from a shared-memory row that was pre-loaded with
// g[t,k] * g[t',:], and the finished row is written (or contracted with d_gradient) at once.  ~5x fewer instructions
// per (t, k, t') than the generic kernel above.
template <int NS, bool CLASSIC>
struct ChainRegs {
  float v0[NS], v1[NS];
};

// occupancies of one frame given alpha-side state (a0,a1) and beta-side state (b0,b1); subtracts them from row[]
template <int NS, bool CLASSIC>
__device__ __forceinline__ void subtract_frame_occupancies(const float* a0, const float* a1, const float* b0,
                                                           const float* b1, const float* dd, float h, float K,
                                                           const int* tok, int tok_left, int blank, int V, int lane,
                                                           float* row, float total) {
  constexpr float kLog2e = 1.4426950408889634f;
  float occ[NS], occ_stay[NS];
  (void)h;
  if (!CLASSIC) {
    float bx = __shfl_down_sync(kFull, b0[0], 1);
    if (lane == 31) bx = kNegInf;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const float bn = (j < NS - 1) ? b0[j + 1] : bx;
      occ[j] = ex2_approx((K + (a0[j] + dd[j] + bn)) * kLog2e);
      if (!((tok[j] != blank) && (tok[j] >= 0) && (tok[j] < V))) occ[j] = 0.0f;
    }
  } else {
    float bx = __shfl_down_sync(kFull, b1[0], 1);
    if (lane == 31) bx = kNegInf;
    float d_left = __shfl_up_sync(kFull, dd[NS - 1], 1);
    if (lane == 0) d_left = kNegInf;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int tp = (j > 0) ? tok[j - 1] : tok_left;
      const float sj = lse2(a0[j], a1[j]);
      const float bn = (j < NS - 1) ? b1[j + 1] : bx;
      occ[j] = ex2_approx((K + (dd[j] + ((tok[j] == tp) ? a0[j] : sj) + bn)) * kLog2e);
      if (!((tok[j] != blank) && (tok[j] >= 0) && (tok[j] < V))) occ[j] = 0.0f;
      const float dp = (j > 0) ? dd[j - 1] : d_left;
      occ_stay[j] = ex2_approx((K + (a1[j] + dp + b1[j])) * kLog2e);
      if (!((tp != blank) && (tp >= 0) && (tp < V))) occ_stay[j] = 0.0f;
    }
    float s_next = __shfl_down_sync(kFull, occ_stay[0], 1);
    if (lane == 31) s_next = 0.0f;
#pragma unroll
    for (int j = 0; j < NS; ++j) occ[j] += (j < NS - 1) ? occ_stay[j + 1] : s_next;
  }
  // The conditioned chain carries the probability mass `total` (= the occupancy of the conditioning emission, -g[t,k]) and
  // every alignment emits exactly one symbol per frame, so the blank's joint occupancy is `total` minus the others: one
  // warp sum instead of the log-sum-exp over states the reference forms (classic_ctc_loss.py:608-614,
  // simplified_ctc_loss.py:498-501).  Measured 6-7 % faster on the dense Hessian and on the HVP.
  float osum = 0.0f;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    osum += occ[j];
    if (occ[j] > 0.0f) atomicAdd(&row[tok[j]], -occ[j]);
  }
  osum = warp_sum(osum);
  __syncwarp();
  if (lane == 0) row[blank] -= total - osum;      // no state scatters into the blank column: a plain update
}

template <int NS>
__device__ __forceinline__ void load_state(const float* src, int planes, int lane, float* v0, float* v1) {
  constexpr int kUpad = NS * kWarp;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    v0[j] = src[j * kWarp + lane];
    v1[j] = (planes == 2) ? src[kUpad + j * kWarp + lane] : kNegInf;
  }
}

// resident CTAs per SM the register-resident form is compiled for (A/B switch, tools/build_variant.sh)
#ifndef CTCB200_K4_MIN_BLOCKS
#define CTCB200_K4_MIN_BLOCKS 1
#endif
template <int NS, bool CLASSIC, bool HVP>
__global__ void __launch_bounds__(kK4Warps * kWarp, CTCB200_K4_MIN_BLOCKS)
    k4_hessian_regs(Problem p, Scratch s, const float* __restrict__ g, float* __restrict__ hessian,
                    const float* __restrict__ d_gradient, float* __restrict__ hvp_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int S = CLASSIC ? 2 : 1, kUpad = NS * kWarp;
  const int b = blockIdx.x / p.T, t = blockIdx.x % p.T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = utt_label_len(p, b), n_t = utt_frames(p, b);
  const double lossd = s.lossd[b];
  const bool dead = (lossd == (double)INFINITY) || (t >= n_t);
  const size_t TV = (size_t)p.T * p.V;
  float* slab = HVP ? nullptr : hessian + ((size_t)b * p.T + t) * p.V * TV;
  float* out_hv = HVP ? hvp_out + ((size_t)b * p.T + t) * p.V : nullptr;
  // 128-bit zero fill of `n` floats at `dst` by `nthr` cooperating threads (scalar when the run is not 16-byte aligned)
  auto zero_fill = [&](float* dst, size_t n, int me, int nthr) {
    if (((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((n & 3) == 0)) {
      float4* d4 = reinterpret_cast<float4*>(dst);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (size_t i = me; i < (n >> 2); i += nthr) d4[i] = z;
    } else {
      for (size_t i = me; i < n; i += nthr) dst[i] = 0.0f;
    }
  };
  if (dead) {
    if (HVP) for (int k = tid; k < p.V; k += blockDim.x) out_hv[k] = 0.0f;
    else zero_fill(slab, (size_t)p.V * TV, tid, blockDim.x);
    return;
  }
  float* row = reinterpret_cast<float*>(smem_raw) + (size_t)warp * p.V;
  // Work list of the CTA.  Only the blank and the tokens of this utterance's label have non-zero second derivatives
  // (g[t,k] == 0 and no alignment emits k otherwise); which warp got which token used to be k % 8, i.e. luck: with 13 of
  // 32 tokens in the label, some warps propagated four chain pairs while others had none (ncu: 21 % active warps of 37.5 %
  // resident).  Now the distinct label tokens are compacted into `list` (a bitmap over the vocabulary dedups them) and
  // the warps draw (token, direction) items -- the longer direction of every token first, then the shorter ones, then (dense
  // form) the zero rows of the tokens outside the label -- from a shared counter.
  const int nW = (p.V + 31) >> 5;
  unsigned* bitmap = reinterpret_cast<unsigned*>(reinterpret_cast<float*>(smem_raw) + (size_t)kK4Warps * p.V);
  int* list = reinterpret_cast<int*>(bitmap + nW);          // [kUpad + 1] distinct tokens: blank first
  float* hvacc = reinterpret_cast<float*>(list + kUpad + 1);  // [kUpad + 1] HVP: the two directions of a token meet here
  int* ctr = reinterpret_cast<int*>(hvacc + kUpad + 1);      // [0] list length, [1] next work item
  for (int i = tid; i < nW; i += blockDim.x) bitmap[i] = 0u;
  for (int i = tid; i < kUpad + 1; i += blockDim.x) hvacc[i] = 0.0f;
  __syncthreads();
  if (tid == 0) {
    bitmap[p.blank >> 5] = 1u << (p.blank & 31);
    list[0] = p.blank;
    ctr[0] = 1;
    ctr[1] = 0;
  }
  __syncthreads();
  for (int l = tid; l < L; l += blockDim.x) {
    const int tk = utt_token(p, b, l, L);
    if (tk >= 0 && tk < p.V) {
      const unsigned bit = 1u << (tk & 31);
      if (!(atomicOr(&bitmap[tk >> 5], bit) & bit)) list[atomicAdd(&ctr[0], 1)] = tk;
    }
  }
  __syncthreads();
  const int n_lab = ctr[0];
  if (HVP) for (int k = tid; k < p.V; k += blockDim.x) if (!((bitmap[k >> 5] >> (k & 31)) & 1u)) out_hv[k] = 0.0f;
  int tok[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) tok[j] = utt_token(p, b, lane * NS + j, L);
  const int tok_left = utt_token(p, b, lane * NS - 1, L);
  LabelBits<NS> lb;
  if (CLASSIC) lb = make_label_bits<NS>(p, b, L, lane);

  const size_t row_pitch = (size_t)S * kUpad;
  const float* alpha_b = s.alphaT + (size_t)b * (p.T + 1) * row_pitch;
  const float* beta_b = s.betaT + (size_t)b * (p.T + 1) * row_pitch;
  const float* d_b = s.dT + (size_t)b * p.T * kUpad;
  const float* h_b = s.h + (size_t)b * p.T;
  const float* g_b = g + (size_t)b * TV;
  const float* dg_b = HVP ? d_gradient + (size_t)b * TV : nullptr;
  const double* ca_b = s.ca + (size_t)b * (p.T + 1);
  const double* cb_b = s.cb + (size_t)b * (p.T + 1);

  // frame t's own inputs, reused by every token's two chains
  float a0t[NS], a1t[NS], b0t[NS], b1t[NS], dt[NS];
  load_state<NS>(alpha_b + (size_t)t * row_pitch, S, lane, a0t, a1t);
  load_state<NS>(beta_b + (size_t)(t + 1) * row_pitch, S, lane, b0t, b1t);
#pragma unroll
  for (int j = 0; j < NS; ++j) dt[j] = d_b[(size_t)t * kUpad + j * kWarp + lane];
  const float ht = h_b[t];

  // emits one finished row: dense store or contraction with d_gradient
  auto finish_row = [&](int t2, float& hv, float* hk) {
    __syncwarp();
    if (HVP) {
      const float* dg2 = dg_b + (size_t)t2 * p.V;
      for (int k2 = lane; k2 < p.V; k2 += kWarp) hv += row[k2] * dg2[k2];
    } else {
      float* dst = hk + (size_t)t2 * p.V;
      for (int k2 = lane; k2 < p.V; k2 += kWarp) dst[k2] = row[k2];
    }
    __syncwarp();
  };

  const bool fwd_first = (n_t - 1 - t) >= t;          // the direction with more frames goes first
  const int n_items = 2 * n_lab + (HVP ? 0 : p.V);
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(&ctr[1], 1);
    item = __shfl_sync(kFull, item, 0);
    if (item >= n_items) break;
    if (item >= 2 * n_lab) {       // a token outside the label: g[t,k] == 0 and every path term vanishes
      const int kz = item - 2 * n_lab;
      if (!((bitmap[kz >> 5] >> (kz & 31)) & 1u)) zero_fill(slab + (size_t)kz * TV, TV, lane, kWarp);
      continue;
    }
    const int li = (item < n_lab) ? item : item - n_lab;
    const bool fwd = (item < n_lab) == fwd_first;
    const int k = list[li];
    float* hk = HVP ? nullptr : slab + (size_t)k * TV;
    const float gk = g_b[(size_t)t * p.V + k];
    float hv = 0.0f;
    float v0[NS], v1[NS];
    if (fwd) {
    if (!HVP) for (size_t i = (size_t)n_t * p.V + lane; i < TV; i += kWarp) hk[i] = 0.0f;   // frames beyond logit_length
    // same frame: [k == k'] g[t,k] + g[t,k] g[t,k']
    {
      const float* g2 = g_b + (size_t)t * p.V;
      for (int k2 = lane; k2 < p.V; k2 += kWarp) row[k2] = gk * g2[k2] + ((k2 == k) ? gk : 0.0f);
      finish_row(t, hv, hk);
    }
    // ---- later frames: push alpha[t] through "emit k", propagate forward ----
    {
      float a_left0 = __shfl_up_sync(kFull, a0t[NS - 1], 1), a_left1 = __shfl_up_sync(kFull, a1t[NS - 1], 1);
      float d_left = __shfl_up_sync(kFull, dt[NS - 1], 1);
      if (lane == 0) { a_left0 = kNegInf; a_left1 = kNegInf; d_left = kNegInf; }
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int tp = (j > 0) ? tok[j - 1] : tok_left;             // label[l-1]
        const float pa0 = (j > 0) ? a0t[j - 1] : a_left0, pa1 = (j > 0) ? a1t[j - 1] : a_left1;
        const float pd = (j > 0) ? dt[j - 1] : d_left;
        if (k == p.blank) {
          v0[j] = ht + (CLASSIC ? lse2(a0t[j], a1t[j]) : a0t[j]);
          v1[j] = kNegInf;
        } else if (!CLASSIC) {
          v0[j] = (tp == k) ? pa0 + pd : kNegInf;
          v1[j] = kNegInf;
        } else {
          const bool rep_prev = (j > 0) ? ((lb.rep >> (j - 1)) & 1u) : lb.rep_left;
          const float xprev = rep_prev ? pa0 : lse2(pa0, pa1);
          v0[j] = kNegInf;
          v1[j] = (tp == k) ? lse2(a1t[j] + pd, pd + xprev) : kNegInf;
        }
      }
    }
    for (int t2 = t + 1; t2 < n_t; ++t2) {
      float bn0[NS], bn1[NS], d2[NS];
      load_state<NS>(beta_b + (size_t)(t2 + 1) * row_pitch, S, lane, bn0, bn1);
#pragma unroll
      for (int j = 0; j < NS; ++j) d2[j] = d_b[(size_t)t2 * kUpad + j * kWarp + lane];
      const float h2 = h_b[t2];
      const float* g2 = g_b + (size_t)t2 * p.V;
      for (int k2 = lane; k2 < p.V; k2 += kWarp) row[k2] = gk * g2[k2];
      __syncwarp();
      const float K = (float)(lossd + ca_b[t] + cb_b[t2 + 1]);
      subtract_frame_occupancies<NS, CLASSIC>(v0, v1, bn0, bn1, d2, h2, K, tok, tok_left, p.blank, p.V, lane, row, -gk);
      finish_row(t2, hv, hk);
      if (CLASSIC) alpha_step_classic<NS>(v0, v1, d2, h2, lane, lb);
      else alpha_step_simplified<NS>(v0, d2, h2, lane);
    }
    } else {
    // ---- earlier frames: pull beta[t+1] back through "emit k", propagate backward ----
    {
      float b_right1 = __shfl_down_sync(kFull, CLASSIC ? b1t[0] : b0t[0], 1);
      if (lane == 31) b_right1 = kNegInf;
      float d_left = __shfl_up_sync(kFull, dt[NS - 1], 1);
      if (lane == 0) d_left = kNegInf;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int l = lane * NS + j;
        const int tp = (j > 0) ? tok[j - 1] : tok_left;
        const float bnext = (j < NS - 1) ? (CLASSIC ? b1t[j + 1] : b0t[j + 1]) : b_right1;   // beta[t+1, l+1(,open)]
        if (k == p.blank) {
          v0[j] = ht + b0t[j];
          v1[j] = v0[j];
        } else {
          const float mv = (l < L && tok[j] == k) ? dt[j] + bnext : kNegInf;
          v0[j] = mv;
          if (CLASSIC) {
            const float pd = (j > 0) ? dt[j - 1] : d_left;
            const float sv = (l >= 1 && tp == k) ? pd + b1t[j] : kNegInf;
            const bool rep = (lb.rep >> j) & 1u;
            v1[j] = lse2(sv, rep ? kNegInf : mv);
          } else {
            v1[j] = kNegInf;
          }
        }
      }
    }
    for (int t2 = t - 1; t2 >= 0; --t2) {
      float al0[NS], al1[NS], d2[NS];
      load_state<NS>(alpha_b + (size_t)t2 * row_pitch, S, lane, al0, al1);
#pragma unroll
      for (int j = 0; j < NS; ++j) d2[j] = d_b[(size_t)t2 * kUpad + j * kWarp + lane];
      const float h2 = h_b[t2];
      const float* g2 = g_b + (size_t)t2 * p.V;
      for (int k2 = lane; k2 < p.V; k2 += kWarp) row[k2] = gk * g2[k2];
      __syncwarp();
      const float K = (float)(lossd + ca_b[t2] + cb_b[t + 1]);
      subtract_frame_occupancies<NS, CLASSIC>(al0, al1, v0, v1, d2, h2, K, tok, tok_left, p.blank, p.V, lane, row, -gk);
      finish_row(t2, hv, hk);
      if (CLASSIC) beta_step_classic<NS>(v0, v1, d2, h2, lane, lb);
      else beta_step_simplified<NS>(v0, d2, h2, lane);
    }
    }
    if (HVP) {
      hv = warp_sum(hv);
      if (lane == 0) atomicAdd(&hvacc[li], hv);      // two terms onto zero: the order does not matter
    }
  }
  if (HVP) {
    __syncthreads();
    for (int i = tid; i < n_lab; i += blockDim.x) out_hv[list[i]] = hvacc[i];
  }
}

// rows of the eight warps + vocabulary bitmap + token list + HVP accumulators + two counters
template <int NS>
static size_t k4_regs_smem_bytes(const Problem& p) {
  return (size_t)kK4Warps * p.V * sizeof(float) + (size_t)((p.V + 31) >> 5) * sizeof(unsigned) +
         (size_t)(NS * kWarp + 1) * (sizeof(int) + sizeof(float)) + 2 * sizeof(int);
}

template <int NS, bool CLASSIC, bool HVP>
static cudaError_t launch_k4_regs(const Problem& p, const Scratch& s, const float* g, float* hessian,
                                  const float* d_gradient, float* hvp_out, cudaStream_t st) {
  const size_t smem = k4_regs_smem_bytes<NS>(p);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k4_hessian_regs<NS, CLASSIC, HVP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const unsigned grid = (unsigned)((long long)p.B * p.T);
  k4_hessian_regs<NS, CLASSIC, HVP><<<grid, kK4Warps * kWarp, smem, st>>>(p, s, g, hessian, d_gradient, hvp_out);
  return cudaGetLastError();
}

template <bool CLASSIC, bool HVP>
static cudaError_t launch_k4_regs_ns(const Problem& p, const Scratch& s, const float* g, float* hessian,
                                     const float* d_gradient, float* hvp_out, cudaStream_t st) {
  switch (p.NS) {
    case 1: return launch_k4_regs<1, CLASSIC, HVP>(p, s, g, hessian, d_gradient, hvp_out, st);
    case 2: return launch_k4_regs<2, CLASSIC, HVP>(p, s, g, hessian, d_gradient, hvp_out, st);
    case 3: return launch_k4_regs<3, CLASSIC, HVP>(p, s, g, hessian, d_gradient, hvp_out, st);
    case 4: return launch_k4_regs<4, CLASSIC, HVP>(p, s, g, hessian, d_gradient, hvp_out, st);
    default: return cudaErrorInvalidValue;
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
  const bool regs = (p.NS <= 4) && ((size_t)kK4Warps * p.V * sizeof(float) <= 200 * 1024) && getenv("CTCB200_K4_GENERIC") == nullptr;
  if (hessian != nullptr) {
    cudaError_t e;
    if (regs) e = classic ? launch_k4_regs_ns<true, false>(p, s, g, hessian, nullptr, nullptr, st)
                          : launch_k4_regs_ns<false, false>(p, s, g, hessian, nullptr, nullptr, st);
    else e = classic ? launch_k4<true, false>(p, s, g, hessian, nullptr, nullptr, st)
                     : launch_k4<false, false>(p, s, g, hessian, nullptr, nullptr, st);
    if (e != cudaSuccess) return e;
  }
  if (hvp_out != nullptr) {
    if (regs) return classic ? launch_k4_regs_ns<true, true>(p, s, g, nullptr, d_gradient, hvp_out, st)
                             : launch_k4_regs_ns<false, true>(p, s, g, nullptr, d_gradient, hvp_out, st);
    return classic ? launch_k4<true, true>(p, s, g, nullptr, d_gradient, hvp_out, st)
                   : launch_k4<false, true>(p, s, g, nullptr, d_gradient, hvp_out, st);
  }
  return cudaSuccess;
}

}  // namespace ctcb200
