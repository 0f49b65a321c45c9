// Per-frame transition occupancies: the device form of _combine_transition_probabilities
// (tf_seq2seq_losses/classic_ctc_loss.py:565-669, simplified_ctc_loss.py:456-534) followed by exp(loss + .).
// Shared by the gradient writer (K3) and the Hessian kernel (K4).
#pragma once
#include "common.cuh"

namespace ctcb200 {

constexpr unsigned short kNoSlot = 0xFFFFu;

// online log-sum-exp accumulator (per lane), merged across the warp at the end
struct LseAcc {
  float m = kNegInf, s = 0.0f;
  __device__ __forceinline__ void add(float v) {
    if (v == kNegInf) return;
    if (v > m) {
      s = s * __expf(m - v) + 1.0f;     // m == -inf -> s == 0 and exp(-inf) == 0
      m = v;
    } else {
      s += __expf(v - m);
    }
  }
  __device__ __forceinline__ float warp_result() const {
    const float M = warp_max(m);
    const float part = (m == kNegInf) ? 0.0f : s * __expf(m - M);
    const float tot = warp_sum(part);
    return (M == kNegInf) ? kNegInf : M + __logf(tot);
  }
};

// neighbours of a private-layout position (state l+1 / l-1); -1 when outside the 32*NS states
__device__ __forceinline__ int pos_next(int pos, int lane, int ns) {
  return ((pos >> 5) + 1 < ns) ? pos + kWarp : ((lane < 31) ? lane + 1 : -1);
}
__device__ __forceinline__ int pos_prev(int pos, int lane, int ns) {
  return ((pos >> 5) > 0) ? pos - kWarp : ((lane > 0) ? (ns - 1) * kWarp + lane - 1 : -1);
}

// Shared-memory tables of one utterance, built once per CTA:
//   toks[l]  cleaned label (base_loss.py:395-418), l in [0, Upad): blank for l >= label_length
//   map[k]   token -> slot (one of the label positions carrying k); blank -> slot Upad; kNoSlot for tokens not in the label
__device__ __forceinline__ void build_utterance_tables(const Problem& p, int b, int L, int* toks,
                                                       unsigned short* map, int Vpad) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int k = tid; k < Vpad; k += nthr) map[k] = kNoSlot;
  for (int l = tid; l < p.Upad; l += nthr) toks[l] = utt_token(p, b, l, L);
  __syncthreads();
  // Any label position carrying the token serves as its slot: the racing 16-bit stores leave exactly one of them, and
  // every reader comes after the barrier.  (A compare-and-swap minimum was 4x slower for character vocabularies, where
  // a hundred labels compete for thirty tokens.)
  for (int l = tid; l < L; l += nthr) {
    const int tok = toks[l];
    if (tok >= 0 && tok < p.V) map[tok] = (unsigned short)l;
  }
  __syncthreads();
  if (tid == 0) map[p.blank] = (unsigned short)p.Upad;   // the blank column is overridden (blank_mask tf.where)
  __syncthreads();
}

__device__ __forceinline__ int tok_at(const Problem& p, const int* toks, int l) {
  return (l >= 0 && l < p.Upad) ? toks[l] : p.blank;
}

// One warp, one frame.  A: state log-weights before the frame, Bn: state log-weights after it, both in the private
// layout with pitch Upad between the closed / open planes; d: label-token log-probs of the frame; h: blank log-prob.
// Fills acc[slot] = sum over the transitions emitting that slot's token of exp(lossb + A + emission + Bn) for every
// slot (acc[Upad] = blank) and returns the total over all tokens.  acc must hold Upad + 1 floats.
template <bool CLASSIC>
__device__ __forceinline__ float row_occupancies(const Problem& p, int L, int lane, const float* A, const float* Bn,
                                                 const float* d, float h, float lossb, const int* toks,
                                                 const unsigned short* map, float* acc) {
  for (int i = lane; i <= p.Upad; i += kWarp) acc[i] = 0.0f;
  __syncwarp();
  LseAcc blank_acc;
  float occ_sum = 0.0f;
  for (int pos = lane; pos < p.Upad; pos += kWarp) {
    const int l = lane * p.NS + (pos >> 5);
    if (l > L) continue;
    const int pn = pos_next(pos, lane, p.NS), pp = pos_prev(pos, lane, p.NS);
    const int tok = toks[l];
    const unsigned short slot = (tok >= 0 && tok < p.V) ? map[tok] : kNoSlot;
    if (!CLASSIC) {
      const float a = A[pos];
      blank_acc.add(a + Bn[pos]);
      if (l < L) {
        const float bnext = (pn >= 0) ? Bn[pn] : kNegInf;
        const float o = __expf(lossb + (a + d[pos] + bnext));
        if (slot != kNoSlot && o > 0.0f) atomicAdd(&acc[slot], o);
        occ_sum += (slot != kNoSlot && slot != p.Upad) ? o : 0.0f;
      }
    } else {
      const float a0 = A[pos], a1 = A[p.Upad + pos];
      blank_acc.add(lse2(a0, a1) + Bn[pos]);
      const int tok_prev = tok_at(p, toks, l - 1);
      if (l < L) {     // diagonal step emitting label[l]: any state of l -> open state of l+1
        const float bnext = (pn >= 0) ? Bn[p.Upad + pn] : kNegInf;
        const float dv = d[pos];
        const float v1 = (tok == tok_prev) ? kNegInf : a1 + dv;
        const float o = __expf(lossb + (lse2(a0 + dv, v1) + bnext));
        if (slot != kNoSlot && o > 0.0f) atomicAdd(&acc[slot], o);
        occ_sum += (slot != kNoSlot && slot != p.Upad) ? o : 0.0f;
      }
      if (l >= 1 && pp >= 0) {   // horizontal step re-emitting label[l-1]: open l -> open l
        const unsigned short sp = (tok_prev >= 0 && tok_prev < p.V) ? map[tok_prev] : kNoSlot;
        const float o = __expf(lossb + (a1 + d[pp] + Bn[p.Upad + pos]));
        if (sp != kNoSlot && o > 0.0f) atomicAdd(&acc[sp], o);
        occ_sum += (sp != kNoSlot && sp != p.Upad) ? o : 0.0f;
      }
    }
  }
  const float occ_blank = __expf(lossb + (h + blank_acc.warp_result()));
  occ_sum = warp_sum(occ_sum) + occ_blank;
  __syncwarp();
  if (lane == 0) acc[p.Upad] = occ_blank;
  __syncwarp();
  return occ_sum;
}

// Same as row_occupancies with the states-per-lane count known at compile time: the loop over the lane's states unrolls
// and every global load of the frame (alpha, beta, d: 3 to 5 per state) is issued before the first is consumed.  The
// run-time-NS form above serialises them (load -> use -> next load), which left K3 waiting on L2 for most of its time
// (ncu: long-scoreboard stall 8.6 per issued instruction at B=32 T=500 V=29).
template <int NS, bool CLASSIC>
__device__ __forceinline__ float row_occupancies_t(const Problem& p, int L, int lane, const float* A, const float* Bn,
                                                   const float* d, float h, float lossb, const int* toks,
                                                   const unsigned short* map, float* acc) {
  constexpr int kUpad = NS * kWarp;
  for (int i = lane; i <= kUpad; i += kWarp) acc[i] = 0.0f;
  float a0[NS], a1[NS], b0[NS], b1[NS], dv[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int pos = j * kWarp + lane;
    a0[j] = A[pos];
    b0[j] = Bn[pos];
    dv[j] = d[pos];
    if (CLASSIC) {
      a1[j] = A[kUpad + pos];
      b1[j] = Bn[kUpad + pos];
    }
  }
  // neighbours across lanes: state l+1 of the lane's last state, state l-1 of its first
  float b_up = __shfl_down_sync(kFull, CLASSIC ? b1[0] : b0[0], 1);
  if (lane == 31) b_up = kNegInf;
  float d_left = __shfl_up_sync(kFull, dv[NS - 1], 1);
  if (lane == 0) d_left = kNegInf;
  __syncwarp();
  LseAcc blank_acc;
  float occ_sum = 0.0f;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int l = lane * NS + j;
    if (l > L) continue;
    const int tok = toks[l];
    const unsigned short slot = (tok >= 0 && tok < p.V) ? map[tok] : kNoSlot;
    if (!CLASSIC) {
      blank_acc.add(a0[j] + b0[j]);
      if (l < L) {
        const float bnext = (j < NS - 1) ? b0[j < NS - 1 ? j + 1 : j] : b_up;
        const float o = __expf(lossb + (a0[j] + dv[j] + bnext));
        if (slot != kNoSlot && o > 0.0f) atomicAdd(&acc[slot], o);
        occ_sum += (slot != kNoSlot && slot != p.Upad) ? o : 0.0f;
      }
    } else {
      blank_acc.add(lse2(a0[j], a1[j]) + b0[j]);
      const int tok_prev = tok_at(p, toks, l - 1);
      if (l < L) {     // diagonal step emitting label[l]: any state of l -> open state of l+1
        const float bnext = (j < NS - 1) ? b1[j < NS - 1 ? j + 1 : j] : b_up;
        const float v1 = (tok == tok_prev) ? kNegInf : a1[j] + dv[j];
        const float o = __expf(lossb + (lse2(a0[j] + dv[j], v1) + bnext));
        if (slot != kNoSlot && o > 0.0f) atomicAdd(&acc[slot], o);
        occ_sum += (slot != kNoSlot && slot != p.Upad) ? o : 0.0f;
      }
      if (l >= 1) {    // horizontal step re-emitting label[l-1]: open l -> open l
        const unsigned short sp = (tok_prev >= 0 && tok_prev < p.V) ? map[tok_prev] : kNoSlot;
        const float dprev = (j > 0) ? dv[j > 0 ? j - 1 : 0] : d_left;
        const float o = __expf(lossb + (a1[j] + dprev + b1[j]));
        if (sp != kNoSlot && o > 0.0f) atomicAdd(&acc[sp], o);
        occ_sum += (sp != kNoSlot && sp != p.Upad) ? o : 0.0f;
      }
    }
  }
  const float occ_blank = __expf(lossb + (h + blank_acc.warp_result()));
  occ_sum = warp_sum(occ_sum) + occ_blank;
  __syncwarp();
  if (lane == 0) acc[kUpad] = occ_blank;
  __syncwarp();
  return occ_sum;
}

// ---- the same combine in the LOG domain (logarithmic_logproba_gradient, base_loss.py:270-298) ------------------------
// The reference returns log occupancies, finite down to the smallest alignment weight and -inf exactly where a
// (frame, token) pair is impossible; exp(.) of them underflows below e^-87.  Per slot this is the reference's
// unsorted_segment_logsumexp (tools.py:95-119): segment maximum first, then log sum exp(x - max).
// Order-preserving map float <-> int so that shared-memory integer atomicMax implements a float maximum.
__device__ __forceinline__ int float_order(float v) {
  const int i = __float_as_int(v);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float float_unorder(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// One warp, one frame.  Fills mx[slot] (as ordered ints) and sm[slot] so that the log occupancy of the slot's token is
// lossb + mx + log(sm) (-inf when sm == 0), for slots 0..Upad-1; returns the blank's log occupancy.  lossb (the loss plus
// both renormalisation offsets) stays in double until it has met the state term: the loss may be hundreds of nats while
// the result is a small number.
template <int NS, bool CLASSIC>
__device__ __forceinline__ float row_log_occupancies_t(const Problem& p, int L, int lane, const float* A, const float* Bn,
                                                       const float* d, float h, double lossb, const int* toks,
                                                       const unsigned short* map, int* mx, float* sm) {
  constexpr int kUpad = NS * kWarp;
  for (int i = lane; i < kUpad; i += kWarp) {
    mx[i] = float_order(kNegInf);
    sm[i] = 0.0f;
  }
  float a0[NS], a1[NS], b0[NS], b1[NS], dv[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int pos = j * kWarp + lane;
    a0[j] = A[pos];
    b0[j] = Bn[pos];
    dv[j] = d[pos];
    if (CLASSIC) {
      a1[j] = A[kUpad + pos];
      b1[j] = Bn[kUpad + pos];
    }
  }
  float b_up = __shfl_down_sync(kFull, CLASSIC ? b1[0] : b0[0], 1);
  if (lane == 31) b_up = kNegInf;
  float d_left = __shfl_up_sync(kFull, dv[NS - 1], 1);
  if (lane == 0) d_left = kNegInf;
  __syncwarp();
  // log-weights of the transitions leaving / staying in this lane's states, and the slots they land in
  float mv[NS], st[NS];
  unsigned short s_mv[NS], s_st[NS];
  LseAcc blank_acc;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int l = lane * NS + j;
    mv[j] = kNegInf; st[j] = kNegInf;
    s_mv[j] = kNoSlot; s_st[j] = kNoSlot;
    if (l > L) continue;
    const int tok = toks[l];
    const unsigned short slot = (tok >= 0 && tok < p.V) ? map[tok] : kNoSlot;
    if (!CLASSIC) {
      blank_acc.add(a0[j] + b0[j]);
      if (l < L && slot != kNoSlot && slot != kUpad) {
        const float bnext = (j < NS - 1) ? b0[j < NS - 1 ? j + 1 : j] : b_up;
        mv[j] = a0[j] + dv[j] + bnext;
        s_mv[j] = slot;
      }
    } else {
      blank_acc.add(lse2(a0[j], a1[j]) + b0[j]);
      const int tok_prev = tok_at(p, toks, l - 1);
      if (l < L && slot != kNoSlot && slot != kUpad) {
        const float bnext = (j < NS - 1) ? b1[j < NS - 1 ? j + 1 : j] : b_up;
        const float v1 = (tok == tok_prev) ? kNegInf : a1[j] + dv[j];
        mv[j] = lse2(a0[j] + dv[j], v1) + bnext;
        s_mv[j] = slot;
      }
      if (l >= 1) {
        const unsigned short sp = (tok_prev >= 0 && tok_prev < p.V) ? map[tok_prev] : kNoSlot;
        if (sp != kNoSlot && sp != kUpad) {
          const float dprev = (j > 0) ? dv[j > 0 ? j - 1 : 0] : d_left;
          st[j] = a1[j] + dprev + b1[j];
          s_st[j] = sp;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    if (s_mv[j] != kNoSlot && mv[j] != kNegInf) atomicMax(&mx[s_mv[j]], float_order(mv[j]));
    if (CLASSIC && s_st[j] != kNoSlot && st[j] != kNegInf) atomicMax(&mx[s_st[j]], float_order(st[j]));
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    if (s_mv[j] != kNoSlot && mv[j] != kNegInf) atomicAdd(&sm[s_mv[j]], __expf(mv[j] - float_unorder(mx[s_mv[j]])));
    if (CLASSIC && s_st[j] != kNoSlot && st[j] != kNegInf) atomicAdd(&sm[s_st[j]], __expf(st[j] - float_unorder(mx[s_st[j]])));
  }
  const float lb = blank_acc.warp_result();
  __syncwarp();
  return (lb == kNegInf || h == kNegInf) ? kNegInf : (float)(lossb + ((double)h + (double)lb));
}

}  // namespace ctcb200
