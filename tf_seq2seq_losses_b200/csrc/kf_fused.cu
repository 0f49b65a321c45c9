// Host side of the fused kernel: configuration choice and dispatch to the per-variant instantiations.
#include "kf_fused.cuh"

namespace ctcb200 {

// Developer / test hook (ctcb200_debug_fused_plan in ctc_b200.h): a process-wide override of the plan below.
static int g_plan_override[5] = {0, 0, 0, 0, 0};     // W, SL, XA, R, mode; W == 0: no override
void fused_set_plan_override(int W, int SL, int XA, int R, int mode) {     // mode: bit 0 split, 1 no HALF scratch, 2 no idle warps, 3 no row helpers
  const int v[5] = {W, SL, XA, R, mode};
  for (int i = 4; i >= 0; --i) __atomic_store_n(&g_plan_override[i], v[i], __ATOMIC_RELEASE);   // W last
}

// One SM holds 228 KB of shared memory and every resident CTA reserves 1 KB of it; one CTA may opt in to 227 KB.
constexpr int kSmemHalfSm = 228 * 1024 / 2 - 1024;
constexpr int kNumSms = 148;

// (workers per side, row buffers per worker, extra phase-A buffer, ring depth, split) for this problem; W == 0 when the
// fused kernel cannot take it.
//   B > 148  one CTA per utterance, both sides inside (W <= 4): prefer plans that leave room for two CTAs per SM.
//   B <= 74  split: a cluster of two CTAs per utterance, one side each (W <= 8), every CTA with an SM to itself.  Measured at T=1000 V=1024 L=200 (simplified): B=32 275 vs 427 us, B=64 279 vs
//            431 us.
//   B <= 148 split with 6 workers per side, two CTAs per SM (see below).
// The HALF state scratch (fused_layout) is used whenever the plan allows it and its hand-over vectors fit the budget.
static bool fits(const Problem& p, int W, int SL, int XA, int R, int sides, int budget, int want_half, int* half) {
  for (int h = (want_half && fused_half_ok(p.S, W, R)) ? 1 : 0; h >= 0; --h)
    if (fused_layout(p.V, p.Upad, p.S, W, SL, XA, R, sides, h).total <= budget) {
      *half = h;
      return true;
    }
  return false;
}

static bool fused_pick(const Problem& p, int* W, int* SL, int* XA, int* R, int* split, int* half) {
  *W = 0; *SL = 0; *XA = 0; *R = 0; *split = 0; *half = 0;
  if (p.NS > kMaxNS) return false;
  const int ow = __atomic_load_n(&g_plan_override[0], __ATOMIC_ACQUIRE);
  const int mode = ow > 0 ? g_plan_override[4] : 0, want_half = (mode & 2) ? 0 : 1;
  if (ow > 0) {
    const int sl = g_plan_override[1], xa = g_plan_override[2], r = g_plan_override[3], sp = mode & 1;
    if (ow <= (sp ? (p.NS > 8 ? 6 : kMaxWorkersSplit) : kMaxWorkers) && sl >= 2 && sl <= 3 && xa >= 0 && xa <= 1 && r >= 1 && r <= 2 * ow &&
        fits(p, ow, sl, xa, r, sp ? 1 : 2, kSmemPerSm, want_half, half)) {
      *W = ow; *SL = sl; *XA = xa; *R = r; *split = sp;
      return true;
    }
  }
  if (2 * p.B <= kNumSms) {
    static const int scand[8][4] = {{8, 3, 1, 16}, {8, 2, 1, 16}, {8, 2, 0, 16}, {6, 3, 1, 12}, {6, 2, 1, 12}, {6, 2, 0, 12},
                                    {4, 3, 1, 8}, {4, 2, 0, 8}};
    for (int c = 0; c < 8; ++c) {
      // Registers are a per-scheduler resource (16 K each): nine warps put three on one scheduler, which holds at most
      // 112 registers per thread.  The wide-state kernels (NS > 8, up to 224 registers) run seven warps, two per scheduler.
      if (p.NS > 8 && scand[c][0] > 6) continue;
      if (fits(p, scand[c][0], scand[c][1], scand[c][2], scand[c][3], 1, kSmemPerSm, want_half, half)) {
        *W = scand[c][0]; *SL = scand[c][1]; *XA = scand[c][2]; *R = scand[c][3]; *split = 1;
        return true;
      }
    }
  }
  // 74 < B <= 148 (the per-GPU slice of the named batch on two GPUs): still split, two CTAs per SM -- 6 workers per side
  // (1 + 6 + 1 idle = 8 warps at 112 registers: two such CTAs fill the register file exactly).  Measured at B=128
  // (simplified, T=1000 V=1024 L=200): 371 us against 464 us for the one-CTA plan, 440 us for split W=4, 611 us for W=8.
  if (p.B <= kNumSms && p.NS <= 8) {
    static const int hcand[5][4] = {{6, 2, 1, 12}, {6, 2, 0, 12}, {4, 3, 1, 8}, {4, 2, 1, 8}, {4, 2, 0, 8}};
    for (int c = 0; c < 5; ++c)
      if (fits(p, hcand[c][0], hcand[c][1], hcand[c][2], hcand[c][3], 1, kSmemHalfSm, want_half, half)) {
        *W = hcand[c][0]; *SL = hcand[c][1]; *XA = hcand[c][2]; *R = hcand[c][3]; *split = 1;
        return true;
      }
  }
  // {workers per side, row buffers per worker, extra phase-A row buffer, ring depth}.  Measured on B200 (B=256 T=1000
  // V=1024): 4 workers beat 3 for both variants; the classic variant (two state planes) only fits 4 workers next to a
  // second CTA with a 6-frame ring.
  static const int cand[11][4] = {{4, 2, 1, 8}, {4, 2, 0, 8}, {4, 2, 0, 6}, {3, 3, 1, 6}, {3, 3, 0, 6}, {3, 2, 1, 6},
                                  {3, 2, 0, 6}, {2, 3, 0, 4}, {2, 2, 0, 4}, {1, 2, 0, 2}, {1, 2, 0, 1}};
  const int budgets[2] = {kSmemHalfSm, kSmemPerSm};
  for (int bi = 0; bi < 2; ++bi)
    for (int c = 0; c < 11; ++c) {
      if (fits(p, cand[c][0], cand[c][1], cand[c][2], cand[c][3], 2, budgets[bi], want_half, half)) {
        *W = cand[c][0]; *SL = cand[c][1]; *XA = cand[c][2]; *R = cand[c][3];
        return true;
      }
    }
  return false;
}

int fused_pick_workers(const Problem& p) {
  int W, SL, XA, R, split, half;
  fused_pick(p, &W, &SL, &XA, &R, &split, &half);
  return W;
}

cudaError_t launch_fused(const Problem& p, const Scratch& s, const float* d_loss, float* loss, float* grad, int W,
                         cudaStream_t st) {
  if (p.B == 0) return cudaSuccess;
  FusedArgs a;
  a.p = p;
  a.rowlse = s.rowlse;
  a.stateT = s.alphaT;      // the staged path's alpha scratch is large enough ([B,(T+1),S,Upad])
  a.coff = s.ca;
  a.d_loss = d_loss;
  a.loss = loss;
  a.grad = grad;
  a.dbg = nullptr;
#ifdef CTCB200_FUSED_TIMING
  a.dbg = reinterpret_cast<long long*>(s.betaT);   // the staged path's beta scratch is unused by the fused kernel
#endif
  a.tma = ((p.V & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.logits) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0) ? 1 : 0;
  if (p.logits_bf16 && (!a.tma || (p.V & 7) != 0)) return cudaErrorInvalidValue;    // checked by the caller (api.cu)
  fused_pick(p, &a.W, &a.SL, &a.XA, &a.R, &a.split, &a.half);
  // Row helpers: one-CTA plans of wide fp32 rows with at most two workers per side (see fused_layout); the extra barriers
  // must still fit next to the plan's shared memory.  Debug mode bit 3 switches them off.
  a.helpers = 0;
  if (a.W > 0 && !a.split && a.tma && !p.logits_bf16 && fused_helpers_ok(p.V, a.W, a.R, 2) &&
      !(g_plan_override[4] & 8)) {
    const int with = fused_layout(p.V, p.Upad, p.S, a.W, a.SL, a.XA, a.R, 2, a.half, 1).total;
    const int without = fused_layout(p.V, p.Upad, p.S, a.W, a.SL, a.XA, a.R, 2, a.half, 0).total;
    // the same residency as the plan without them: two CTAs per SM stay two CTAs per SM
    if (with <= kSmemPerSm && (without > kSmemHalfSm || with <= kSmemHalfSm)) a.helpers = 1;
  }
  a.rec_alone = (a.split && !(g_plan_override[0] > 0 && (g_plan_override[4] & 4))) ? 1 : 0;     // mode bit 2: off
  (void)W;
  const bool classic = p.variant == CTCB200_CLASSIC;
  if (p.logits_bf16) return classic ? launch_fused_variant<true, true, true>(a, st) : launch_fused_variant<false, true, true>(a, st);
  if (classic) return a.tma ? launch_fused_variant<true, true, false>(a, st) : launch_fused_variant<true, false, false>(a, st);
  return a.tma ? launch_fused_variant<false, true, false>(a, st) : launch_fused_variant<false, false, false>(a, st);
}

}  // namespace ctcb200
