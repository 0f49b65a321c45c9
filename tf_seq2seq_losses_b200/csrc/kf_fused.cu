// Host side of the fused kernel: configuration choice and dispatch to the per-variant instantiations.
#include "kf_fused.cuh"

namespace ctcb200 {

// (workers per side, row buffers per worker) for this problem; W == 0 when the fused kernel cannot take it.
// Preference: configurations that leave room for two CTAs per SM (more warps to hide latency), then one CTA per SM.
static bool fused_pick(const Problem& p, int* W, int* SL) {
  *W = 0; *SL = 0;
  if (p.NS > kMaxNS) return false;
  // developer override for experiments: CTCB200_FUSED_W / CTCB200_FUSED_SL
  const char* ew = getenv("CTCB200_FUSED_W");
  const char* es = getenv("CTCB200_FUSED_SL");
  if (ew != nullptr && es != nullptr) {
    const int w = atoi(ew), sl = atoi(es);
    if (w >= 1 && w <= (p.S == 2 ? 3 : kMaxWorkers) && sl >= 2 && sl <= kMaxRowSlots &&
        fused_layout(p.V, p.Upad, p.S, w, sl).total <= kSmemPerSm) {
      *W = w; *SL = sl;
      return true;
    }
  }
  // measured on B200 (cfg B=256 T=1000 V=1024): 4 workers x 2 buffers beats 3 x 3; the classic variant keeps 3 workers
  // because its recursion warps need the 128-register budget of a 256-thread CTA.
  static const int cand[6][2] = {{4, 2}, {3, 3}, {3, 2}, {2, 3}, {2, 2}, {1, 2}};
  const int budgets[2] = {kSmemPerSm / 2 - 1024, kSmemPerSm};
  for (int bi = 0; bi < 2; ++bi)
    for (int c = (p.S == 2 ? 1 : 0); c < 6; ++c)
      if (fused_layout(p.V, p.Upad, p.S, cand[c][0], cand[c][1]).total <= budgets[bi]) {
        *W = cand[c][0]; *SL = cand[c][1];
        return true;
      }
  return false;
}

int fused_pick_workers(const Problem& p) {
  int W, SL;
  fused_pick(p, &W, &SL);
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
  fused_pick(p, &a.W, &a.SL);
  (void)W;
  a.tma = ((p.V & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.logits) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0) ? 1 : 0;
  const bool classic = p.variant == CTCB200_CLASSIC;
  if (classic) return a.tma ? launch_fused_variant<true, true>(a, st) : launch_fused_variant<true, false>(a, st);
  return a.tma ? launch_fused_variant<false, true>(a, st) : launch_fused_variant<false, false>(a, st);
}

}  // namespace ctcb200
