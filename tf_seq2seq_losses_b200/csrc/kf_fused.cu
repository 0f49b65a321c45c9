// Host side of the fused kernel: configuration choice and dispatch to the per-variant instantiations.
#include "kf_fused.cuh"

namespace ctcb200 {

// (workers per side, row buffers per worker, extra phase-A buffer, ring depth) for this problem; W == 0 when the fused
// kernel cannot take it.  Preference: configurations that leave room for two CTAs per SM (more warps to hide latency),
// then one CTA per SM.
static bool fused_pick(const Problem& p, int* W, int* SL, int* XA, int* R) {
  *W = 0; *SL = 0; *XA = 0; *R = 0;
  if (p.NS > kMaxNS) return false;
  // developer override for experiments: CTCB200_FUSED_W / CTCB200_FUSED_SL / CTCB200_FUSED_XA / CTCB200_FUSED_R
  const char* ew = getenv("CTCB200_FUSED_W");
  const char* es = getenv("CTCB200_FUSED_SL");
  const char* ex = getenv("CTCB200_FUSED_XA");
  const char* er = getenv("CTCB200_FUSED_R");
  if (ew != nullptr && es != nullptr) {
    const int w = atoi(ew), sl = atoi(es), xa = ex ? atoi(ex) : 0, r = er ? atoi(er) : 2 * w;
    if (w >= 1 && w <= kMaxWorkers && sl >= 2 && sl <= 3 && xa >= 0 && xa <= 1 && r >= w && r <= 2 * w &&
        fused_layout(p.V, p.Upad, p.S, w, sl, xa, r).total <= kSmemPerSm) {
      *W = w; *SL = sl; *XA = xa; *R = r;
      return true;
    }
  }
  // {workers per side, row buffers per worker, extra phase-A row buffer, ring depth}.  Measured on B200 (B=256 T=1000
  // V=1024): 4 workers beat 3 for both variants; the classic variant (two state planes) only fits 4 workers next to a
  // second CTA with a 6-frame ring.
  static const int cand[11][4] = {{4, 2, 1, 8}, {4, 2, 0, 8}, {4, 2, 0, 6}, {3, 3, 1, 6}, {3, 3, 0, 6}, {3, 2, 1, 6},
                                  {3, 2, 0, 6}, {2, 3, 0, 4}, {2, 2, 0, 4}, {1, 2, 0, 2}, {1, 2, 0, 1}};
  // an SM has 228 KB of shared memory and every resident CTA reserves 1 KB of it; one CTA may opt in to 227 KB
  const int budgets[2] = {228 * 1024 / 2 - 1024, kSmemPerSm};
  for (int bi = 0; bi < 2; ++bi)
    for (int c = 0; c < 11; ++c) {
      if (fused_layout(p.V, p.Upad, p.S, cand[c][0], cand[c][1], cand[c][2], cand[c][3]).total <= budgets[bi]) {
        *W = cand[c][0]; *SL = cand[c][1]; *XA = cand[c][2]; *R = cand[c][3];
        return true;
      }
    }
  return false;
}

int fused_pick_workers(const Problem& p) {
  int W, SL, XA, R;
  fused_pick(p, &W, &SL, &XA, &R);
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
  fused_pick(p, &a.W, &a.SL, &a.XA, &a.R);
  (void)W;
  a.tma = ((p.V & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.logits) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0) ? 1 : 0;
  const bool classic = p.variant == CTCB200_CLASSIC;
  if (classic) return a.tma ? launch_fused_variant<true, true>(a, st) : launch_fused_variant<true, false>(a, st);
  return a.tma ? launch_fused_variant<false, true>(a, st) : launch_fused_variant<false, false>(a, st);
}

}  // namespace ctcb200
