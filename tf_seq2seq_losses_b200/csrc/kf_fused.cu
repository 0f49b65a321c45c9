// Host side of the fused kernel: configuration choice and dispatch to the per-variant instantiations.
#include "kf_fused.cuh"

namespace ctcb200 {

// (workers per side, row buffers per worker) for this problem; W == 0 when the fused kernel cannot take it.
// Preference: configurations that leave room for two CTAs per SM (more warps to hide latency), then one CTA per SM.
static bool fused_pick(const Problem& p, int* W, int* SL, int* XA) {
  *W = 0; *SL = 0; *XA = 0;
  if (p.NS > kMaxNS) return false;
  const int wmax = (p.S == 2) ? 3 : kMaxWorkers;
  // developer override for experiments: CTCB200_FUSED_W / CTCB200_FUSED_SL / CTCB200_FUSED_XA
  const char* ew = getenv("CTCB200_FUSED_W");
  const char* es = getenv("CTCB200_FUSED_SL");
  const char* ex = getenv("CTCB200_FUSED_XA");
  if (ew != nullptr && es != nullptr) {
    const int w = atoi(ew), sl = atoi(es), xa = ex ? atoi(ex) : 0;
    if (w >= 1 && w <= wmax && sl >= 2 && sl <= 3 && xa >= 0 && xa <= 1 &&
        fused_layout(p.V, p.Upad, p.S, w, sl, xa).total <= kSmemPerSm) {
      *W = w; *SL = sl; *XA = xa;
      return true;
    }
  }
  // {workers per side, row buffers per worker, extra phase-A row buffer}.  Measured on B200 (B=256 T=1000 V=1024):
  // 4 workers beat 3; the classic variant keeps 3 workers because its recursion warps need the 128-register budget of
  // a 256-thread CTA.
  static const int cand[9][3] = {{4, 2, 1}, {4, 2, 0}, {3, 3, 1}, {3, 3, 0}, {3, 2, 1}, {3, 2, 0}, {2, 3, 0}, {2, 2, 0}, {1, 2, 0}};
  // an SM has 228 KB of shared memory and every resident CTA reserves 1 KB of it; one CTA may opt in to 227 KB
  const int budgets[2] = {228 * 1024 / 2 - 1024, kSmemPerSm};
  for (int bi = 0; bi < 2; ++bi)
    for (int c = 0; c < 9; ++c) {
      if (cand[c][0] > wmax) continue;
      if (fused_layout(p.V, p.Upad, p.S, cand[c][0], cand[c][1], cand[c][2]).total <= budgets[bi]) {
        *W = cand[c][0]; *SL = cand[c][1]; *XA = cand[c][2];
        return true;
      }
    }
  return false;
}

int fused_pick_workers(const Problem& p) {
  int W, SL, XA;
  fused_pick(p, &W, &SL, &XA);
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
  fused_pick(p, &a.W, &a.SL, &a.XA);
  (void)W;
  a.tma = ((p.V & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.logits) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0) ? 1 : 0;
  const bool classic = p.variant == CTCB200_CLASSIC;
  if (classic) return a.tma ? launch_fused_variant<true, true>(a, st) : launch_fused_variant<true, false>(a, st);
  return a.tma ? launch_fused_variant<false, true>(a, st) : launch_fused_variant<false, false>(a, st);
}

}  // namespace ctcb200
