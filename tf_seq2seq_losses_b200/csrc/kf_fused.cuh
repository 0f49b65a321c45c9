// KF: the fused loss + d/dlogits kernel -- ONE launch for the whole hot path, one CTA per utterance.
//
// What it replaces in the reference: the log-softmax (tf_seq2seq_losses/tools.py:27-40), the label-state gathers
// (base_loss.py:328-418), the alpha/beta tf.while_loop recursions (classic_ctc_loss.py:310-462,
// simplified_ctc_loss.py:291-438), the loss read-out, _combine_transition_probabilities with its token scatter
// (classic_ctc_loss.py:565-669, simplified_ctc_loss.py:456-534, base_loss.py:420-468), the gradient
// (base_loss.py:262-298) and TF's autodiff of the log-softmax -- i.e. everything the staged kernels K1+K2+K3 do, with
// the [B,T,U] gathered-probability scratch and one of the two state tensors never leaving the SM.
//
// Schedule ("meet in the middle").  For an utterance with n frames, M = n/2:
//   phase A  alpha runs forward over frames 0..M-1 while beta runs backward over frames n-1..M.  Each side stores the
//            state it held *before* consuming a frame (alpha[t] for t < M, beta[t+1] for t >= M) to global scratch.
//   middle   logZ = logsumexp_l(alpha[M,l] + beta[M,l]) -- the normaliser every occupancy needs -- is known half-way.
//   phase B  alpha continues over frames M..n-1 and beta over M-1..0.  At frame t the running side's state and the
//            other side's stored state give the occupancies of that frame, and the gradient row is written at once.
// Every logits row is therefore read twice (once per phase) and every gradient row written once; the serial chain is
// n/2 + n/2 steps instead of 2n.
//
// Warp roles (per side s in {alpha, beta}; W row workers per side):
//   recursion warp   NS states per lane in registers, one shuffle per frame (recursion.cuh); consumes per-frame
//                    inputs (h, d[.]) from a shared-memory ring and publishes its pre-step state.
//   row workers      each owns whole logits rows: a 1-D TMA bulk copy (cp.async.bulk + mbarrier complete_tx) lands the
//                    row in shared memory (double buffered), the warp reduces it (phase A: row log-sum-exp), gathers
//                    the <= U label columns into the ring, and in phase B combines alpha*beta into per-token
//                    occupancies (occupancy.cuh) and streams out the dense gradient row with 128-bit stores.
// Warps hand frames to each other through rings of mbarriers in shared memory (full_d / full_s / empty, see SideView): a
// waiting warp is suspended by the hardware (mbarrier.try_wait), nobody spins on a counter, and there is no CTA-wide
// barrier inside a phase.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "occupancy.cuh"
#include "recursion.cuh"

namespace ctcb200 {

constexpr int kMaxRowSlots = 4;     // row buffers per worker: current + prefetch(es); phase A may own one more (see XA)
constexpr int kFusedGroup = 4;      // frames per unrolled group (renormalisation cadence, see recursion.cuh)
constexpr int kMaxWorkers = 4;       // workers per side when one CTA serves both sides of an utterance
constexpr int kMaxWorkersSplit = 8;  // ... when each side has a CTA (and an SM, or half of one) to itself: see SPLIT below

// A/B switches (developer builds, tools/build_variant.sh): the recursion warp's frame order, the wait watchdog
#ifndef CTCB200_REC_REORDER
#define CTCB200_REC_REORDER 0      // 1: await the next frame's inputs before stepping (measured slower: 622 vs 608 us, classic 1011 vs 926)
#endif
#ifndef CTCB200_WATCHDOG
#define CTCB200_WATCHDOG 0      // 1: every mbarrier wait counts its polls and traps instead of hanging (costs ~15 % at B=256)
#endif

// ---- PTX helpers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// thread-block cluster primitives (split mode: the two sides of an utterance are the two CTAs of a cluster)
__device__ __forceinline__ void cluster_sync_all() {     // every thread of every CTA of the cluster; release / acquire
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_map(unsigned smem_addr, unsigned cta_rank) {   // my address -> the peer's copy
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ float ld_dsmem_f32(unsigned addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double ld_dsmem_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {   // release at CTA scope
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait suspends the warp in hardware until the phase completes or the time hint (ns) expires; the loop lives inside
// the PTX block so a wake-up costs a few instructions, and the long hint keeps idle warps off the issue slots.
// Watchdog: a wait that polls kWatchdogSpins times can only be a protocol bug (the longest legitimate wait is a few
// microseconds of DRAM latency, a handful of polls; the limit is >= 0.1 s even if every poll returned at once): the kernel
// traps -- a CUDA error the caller sees -- instead of hanging the GPU.
#ifndef CTCB200_MBAR_HINT_NS
#define CTCB200_MBAR_HINT_NS 20000
#endif
constexpr unsigned kWatchdogSpins = 1u << 22;
#ifdef CTCB200_WATCHDOG_VERBOSE
// developer build: a C-level poll loop that says who is stuck where before it traps (`site` names the call site)
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity, int site = 0) {
  for (unsigned n = 0;; ++n) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(20000) : "memory");
    if (ok) break;
    if (n == (1u << 15) && (threadIdx.x & 31) == 0)      // report, keep waiting so that every stuck warp gets to report ...
      printf("libctc_b200 watchdog: block %d warp %d stuck at site %d, mbarrier +%u, parity %u\n", (int)blockIdx.x,
             (int)(threadIdx.x >> 5), site, smem_u32(bar), parity);
    if (n > (1u << 17)) __trap();                       // ... then give up
  }
  __syncwarp();
}
#else
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity, int site = 0) {
  (void)site;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
#if CTCB200_WATCHDOG
      ".reg .pred q;\n\t.reg .u32 n;\n\tmov.u32 n, 0;\n\t"
#endif
      "CTCB200_WAIT:\n\t"
#if CTCB200_MBAR_HINT_NS > 0
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
#endif
      "@p bra CTCB200_DONE;\n\t"
#if CTCB200_WATCHDOG
      "add.u32 n, n, 1;\n\tsetp.gt.u32 q, n, %3;\n\t@q trap;\n\t"
#endif
      "bra CTCB200_WAIT;\n\t"
      "CTCB200_DONE:\n\t}"
      :
      : "r"(smem_u32(bar)), "r"(parity), "r"(CTCB200_MBAR_HINT_NS), "r"(kWatchdogSpins)
      : "memory");
#if CTCB200_WATCHDOG
  // The trap is a second way out of the poll loop, and with it the compiler no longer re-converges the warp behind the
  // loop on its own (measured: lanes ran on diverged and the row / ring hand-offs, written for converged warps, raced).
  // Every call site is warp-uniform, so the warp is re-converged here explicitly.
  __syncwarp();
#endif
}
#endif
// Optional L2 eviction priorities (-DCTCB200_L2_HINTS=1; off by default).  The kernel streams 3.1 GB of logits /
// gradient rows through the 126 MB L2 exactly once per phase, while the 0.23 GB of stored recursion states are written
// in phase A and read back last-in-first-out in phase B: with the hints rows are marked evict_first, states evict_last,
// and a state row is discarded from L2 once consumed.  Measured on B200 (B=256 T=1000 V=1024): DRAM traffic drops
// from 3.55 to 3.42 GB (L2 hit rate 2.7 % -> 8.9 %) but the launch is not faster (600 vs 592 us), so they stay off.
#ifndef CTCB200_L2_HINTS
#define CTCB200_L2_HINTS 0
#endif

#ifndef CTCB200_L2_ROW_HINTS
#define CTCB200_L2_ROW_HINTS CTCB200_L2_HINTS
#endif
#ifndef CTCB200_L2_DISCARD
#define CTCB200_L2_DISCARD 1
#endif
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// 1-D TMA: global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
#if CTCB200_L2_ROW_HINTS
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(l2_policy_evict_first())
      : "memory");
#else
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
#endif
}
// 1-D TMA store: shared -> global bulk copy tracked by the per-thread bulk async-group
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, unsigned bytes) {
#if CTCB200_L2_ROW_HINTS
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes), "l"(l2_policy_evict_first())
               : "memory");
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
#endif
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// recursion-state scratch: stores and 16-byte async loads that ask L2 to keep the line, and the final discard
__device__ __forceinline__ void stg_keep(float* p, float v) {
#if CTCB200_L2_HINTS
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(l2_policy_evict_last()) : "memory");
#else
  *p = v;
#endif
}
__device__ __forceinline__ void l2_discard128(const void* p) {
#if CTCB200_L2_HINTS && CTCB200_L2_DISCARD
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
#endif
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fused_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void fused_cp_async16_keep(void* smem_dst, const void* gsrc) {
#if CTCB200_L2_HINTS
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
               "l"(l2_policy_evict_last())
               : "memory");
#else
  fused_cp_async16(smem_dst, gsrc);
#endif
}
__device__ __forceinline__ void fused_cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// the mbarrier receives one (pre-counted) arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fused_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fused_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2: two lanes of fp32 per issue slot) -----------------------------
// exp((v - m)) for the four elements of a float4, as 2^((v - m) * log2e): subtract first (exact for |v| ~ 1e10), then
// scale -- both as packed ops -- then four MUFU.EX2.
__device__ __forceinline__ float4 exp4_shifted(float4 v, float m) {
  const float2 nm = make_float2(-m, -m), k = make_float2(1.4426950408889634f, 1.4426950408889634f);
  const float2 lo = __fmul2_rn(__fadd2_rn(make_float2(v.x, v.y), nm), k);
  const float2 hi = __fmul2_rn(__fadd2_rn(make_float2(v.z, v.w), nm), k);
  return make_float4(ex2_approx(lo.x), ex2_approx(lo.y), ex2_approx(hi.x), ex2_approx(hi.y));
}
__device__ __forceinline__ float hsum4(float4 e) {
  const float2 s = __fadd2_rn(make_float2(e.x, e.y), make_float2(e.z, e.w));
  return s.x + s.y;
}

// ---- bf16 rows (CTCB200_LOGITS_BF16): a row lands in the UPPER half of its fp32-sized buffer and is widened on the fly ----
__device__ __forceinline__ float bf16_lo(unsigned w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(unsigned w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ void widen8(uint4 q, float4& a, float4& b) {
  a = make_float4(bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y));
  b = make_float4(bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w));
}
__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {      // round to nearest even
  unsigned r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---- optional wait-time instrumentation (compile with -DCTCB200_FUSED_TIMING; results go to FusedArgs::dbg) --------
// per warp: [0] phase A cycles, [1] phase B cycles, [2] TMA wait, [3] dcount wait, [4] ccount wait, [5] scount wait,
//           [6] done wait, [7] state cp.async wait; phase-B worker segments: [8] softmax pass, [9] occupancies,
//           [10] scatter, [11] blank term + row store
#ifdef CTCB200_FUSED_TIMING
#define TIMED(slot, stmt)                       \
  do {                                          \
    const long long t0__ = clock64();           \
    stmt;                                       \
    tm[slot] += clock64() - t0__;               \
  } while (0)
#else
#define TIMED(slot, stmt) \
  do {                    \
    stmt;                 \
  } while (0)
#endif

#ifdef CTCB200_WATCHDOG_VERBOSE
#define CTCB200_TRACE(fmt, ...)                                                                         \
  do {                                                                                                  \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) printf("trace warp %d: " fmt "\n", (int)(threadIdx.x >> 5), __VA_ARGS__); \
  } while (0)
#else
#define CTCB200_TRACE(fmt, ...) \
  do {                          \
  } while (0)
#endif

// ---- shared-memory layout (one definition for host sizing and device carving) ---------------------------------------
struct FusedLayout {
  int W, R, SL, XA;         // workers per side, ring depth (W..2W), row buffers per worker, extra phase-A row buffers (0/1)
  int half;                 // 1: phase A stores every second state row only, phase B derives the others (see HALF below)
  int helpers;              // 1: every row worker has a helper warp that takes the upper part of its rows (see helper_phase)
  int off_xch, off_xoff, off_flag, off_side0, off_hbar, hbar_side_bytes, total, xch_aliased;
  // offsets inside a side block
  int s_ctl, s_bar, s_row, s_aux, s_ringd, s_ringh, side_bytes;
  // offsets inside the aux block (phase B view)
  int x_rings, x_ringc, x_stbuf, x_fwd;
};

__host__ __device__ inline int fl_align(int x, int a) { return (x + a - 1) / a * a; }

// Phase A is latency-bound by the bytes it keeps in flight, phase B needs the state ring and the stored-state staging
// buffers instead: the `aux` block is the union of the two (XA extra row buffers per side in phase A; rings / ringc /
// stbuf in phase B), which is what lets 4 workers x 3 row buffers fit next to a second CTA on the SM.
// R = ring depth in frames (>= W; 2W unless shared memory is short).  The exchange vectors of the middle (S*Upad floats
// per side, used only between the phases) alias each side's own input ring when that is large enough.
// `sides` = 2 when one CTA holds both sides of the utterance, 1 in split mode (a CTA per side).
// HALF (simplified variant, even W and R): phase A writes the recursion state of every SECOND frame to global scratch.  In
// phase B the worker of a frame whose row was stored (odd frames, counted from the middle) takes one recursion step from
// it with the inputs of its own frame -- it holds them anyway -- and hands the result to the worker of the frame before
// through shared memory (`fwd`, one vector per even ring slot, barrier fwd_full).  Halves the state traffic (0.40 ->
// 0.20 GB at B=256 T=1000 U=201) and the global stores of the recursion warps.
// Compiled out by default: measured on B200 (simplified, T=1000 V=1024 L=200) it is SLOWER than storing every row -- B=256
// 609 vs 592 us, B=128 507 vs 474 us: the even frames now wait for the worker of the next frame, and at 95 % of the
// copy bandwidth the 5 % of traffic it saves does not buy that back.  -DCTCB200_HALF_SCRATCH=1 brings it back.
#ifndef CTCB200_HALF_SCRATCH
#define CTCB200_HALF_SCRATCH 0
#endif
__host__ __device__ inline bool fused_half_ok(int S, int W, int R) {
  return CTCB200_HALF_SCRATCH && S == 1 && (W & 1) == 0 && (R & 1) == 0;
}
// Row helpers (wide rows, at most two workers per side -- shared memory holds no more at 20 KB a row): a row pass of a
// single warp is ~3 us at V = 5000 and sets the pace of the whole kernel, so every worker gets a second warp that reduces
// (phase A) and exponentiates (phase B) the part of the row from float4 `fused_helper_split(n4)` on.  The pair shares the
// row's TMA barrier; the helper answers through `hdone` (one mbarrier per row buffer and phase) and, in phase A, two
// floats (`hres`: its maximum and its sum of exponentials).
__host__ __device__ inline int fused_helper_split(int n4) { return ((n4 >> 1) / (8 * kWarp)) * (8 * kWarp); }
// (R % W == 0: the frames that share a ring slot belong to one worker, which is what lets the helper wait on the slot's
// full_d barrier in phase B without ever meeting a later use of it.)
__host__ __device__ inline bool fused_helpers_ok(int V, int W, int R, int sides) {
  return sides == 2 && W <= 2 && R % W == 0 && (V & 3) == 0 && fused_helper_split(V >> 2) >= 8 * kWarp;
}
__host__ __device__ inline FusedLayout fused_layout(int V, int Upad, int S, int W, int SL, int XA, int R, int sides = 2,
                                                    int half = 0, int helpers = 0) {
  FusedLayout f;
  f.half = (half && fused_half_ok(S, W, R)) ? 1 : 0;
  f.helpers = (helpers && fused_helpers_ok(V, W, R, sides)) ? 1 : 0;
  f.W = W;
  f.R = R;
  f.SL = SL;
  f.XA = XA;
  const int Vp = (V + 3) & ~3;
  int o = 0;
  f.xch_aliased = (R >= S) ? 1 : 0;       // a side's exchange vector (S*Upad floats) fits its input ring (R*Upad floats)
  f.off_xch = o;  o += f.xch_aliased ? 0 : sides * S * Upad * 4;
  f.off_xoff = o; o += 2 * 8;
  f.off_flag = o; o += 8;                 // CTA-wide facts gathered in the prologue (see fused_body)
  o = fl_align(o, 128);
  int s = 0;
  // Two sets of every barrier, one per phase: all of them are initialised once at kernel start and none is ever
  // re-initialised (see fused_body).
  f.s_ctl = s;   s += 2 * 4 * f.R * 8;                        // ring barriers: full_d[R], full_s[R], empty[R], fwd_full[R]
  f.s_bar = s;   s += 2 * W * kMaxRowSlots * 8;
  s = fl_align(s, 128);
  f.s_row = s;   s += W * SL * Vp * 4;
  int x = 0;
  f.x_rings = x; x += f.R * S * Upad * 4;
  f.x_ringc = x; x += f.R * 8;
  x = fl_align(x, 16);
  f.x_stbuf = x; x += W * S * Upad * 4;
  f.x_fwd = x;   x += f.half ? (f.R / 2) * Upad * 4 : 0;
  const int xa_bytes = XA * W * Vp * 4;
  f.s_aux = s;   s += fl_align(x > xa_bytes ? x : xa_bytes, 16);
  f.s_ringd = s; s += f.R * Upad * 4;
  f.s_ringh = s; s += fl_align(f.R * 4, 16);
  f.side_bytes = fl_align(s, 128);
  f.off_side0 = o;
  f.total = o + sides * f.side_bytes;
  f.off_hbar = f.total;                                       // per side: hdone[2 phases][W][kMaxRowSlots], hres[W][kMaxRowSlots][2]
  f.hbar_side_bytes = f.helpers ? 3 * W * kMaxRowSlots * 8 : 0;
  f.total += sides * f.hbar_side_bytes;
  return f;
}

struct FusedArgs {
  Problem p;
  float* rowlse;        // [B*T]          row log-sum-exp (phase A -> phase B)
  float* stateT;        // [B*T*S*Upad]   row t: alpha[t] if t < M(b) else beta[t+1]; private layout
  double* coff;         // [B*T]          renormalisation offset of that row
  const float* d_loss;  // [B] or null
  float* loss;          // [B]
  float* grad;          // [B,T,V]
  int W, SL, XA, R;
  int half;             // 1: HALF state scratch (see fused_layout)
  int helpers;          // 1: one helper warp per row worker (wide rows, see fused_layout)
  int rec_alone;        // split mode: leave the warps that share the recursion warp's scheduler idle
  int split;            // 1: a cluster of two CTAs per utterance, one per side (small batches); 0: one CTA per utterance
  int tma;              // 1: rows move by 1-D TMA (V % 4 == 0, 16-byte aligned bases); 0: by 4-byte cp.async / plain stores
  long long* dbg;       // [B][warps][12] when built with CTCB200_FUSED_TIMING, else unused
};

// view of one side's shared memory
struct SideView {
  // Ring hand-off barriers (mbarriers, arrival count 1; a waiting warp is suspended by the hardware instead of
  // spinning on an issue slot).  The k-th use of ring slot q completes phase k of its barriers.
  unsigned long long* full_d;   // [R] worker -> recursion: the frame's inputs (h, d[.]) are in the ring slot
  unsigned long long* full_s;   // [R] recursion -> worker (phase B): the pre-step state of the frame is in the ring slot
  unsigned long long* empty;    // [R] slot released: by the recursion once it has read the inputs (phase A), by the
                                //     worker once the frame's gradient row is finished (phase B)
  unsigned long long* fwd_full; // [R] HALF: the derived other-side state of the (even) frame in this slot is in `fwd`
  unsigned long long* bar;   // [W][kMaxRowSlots]
  float* row;         // [W][SL][Vp]
  float* aux_rows;    // [W][Vp]  phase A only: one extra row buffer per worker (aliases rings / ringc / stbuf)
  float* ringd;       // [R][Upad]
  float* ringh;       // [R]
  float* rings;       // [R][S*Upad]
  double* ringc;      // [R]
  float* stbuf;       // [W][S*Upad]
  float* fwd;         // [R/2][Upad]  HALF only
  unsigned long long* hdone;   // [W][kMaxRowSlots]  helpers only: the helper is through with the row in this buffer
  float* hres;                 // [W][kMaxRowSlots][2]  helpers only, phase A: the helper's row maximum and sum of exponentials
};

__device__ __forceinline__ SideView side_view(unsigned char* smem, const FusedLayout& f, int side, int phase) {
  unsigned char* base = smem + f.off_side0 + side * f.side_bytes;
  SideView v;
  unsigned long long* ctl = reinterpret_cast<unsigned long long*>(base + f.s_ctl) + phase * (4 * f.R);
  v.full_d = ctl;
  v.full_s = ctl + f.R;
  v.empty = ctl + 2 * f.R;
  v.fwd_full = ctl + 3 * f.R;
  v.bar = reinterpret_cast<unsigned long long*>(base + f.s_bar) + phase * (f.W * kMaxRowSlots);
  v.row = reinterpret_cast<float*>(base + f.s_row);
  v.aux_rows = reinterpret_cast<float*>(base + f.s_aux);
  v.ringd = reinterpret_cast<float*>(base + f.s_ringd);
  v.ringh = reinterpret_cast<float*>(base + f.s_ringh);
  v.rings = reinterpret_cast<float*>(base + f.s_aux + f.x_rings);
  v.ringc = reinterpret_cast<double*>(base + f.s_aux + f.x_ringc);
  v.stbuf = reinterpret_cast<float*>(base + f.s_aux + f.x_stbuf);
  v.fwd = reinterpret_cast<float*>(base + f.s_aux + f.x_fwd);
  unsigned char* hb = smem + f.off_hbar + side * f.hbar_side_bytes;
  v.hdone = reinterpret_cast<unsigned long long*>(hb) + phase * (f.W * kMaxRowSlots);
  v.hres = reinterpret_cast<float*>(hb + 2 * f.W * kMaxRowSlots * 8);
  return v;
}

__device__ __forceinline__ void fused_zero_row(float* dst, int V, int lane, bool vec) {
  if (vec) {
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = lane; i < (V >> 2); i += kWarp) stg_stream4(d4 + i, z);
  } else {
    for (int i = lane; i < V; i += kWarp) dst[i] = 0.0f;
  }
}

// ---- recursion warp, one phase ---------------------------------------------------------------------------------------
// Frame i of the phase is frame t = t_first + i * t_step of the utterance.  SIDE 0 = alpha (forward), 1 = beta.
// One rolled loop serves both phases (`phase_b` is a warp-uniform runtime flag) and the renormalisation cadence is
// counted at run time: unrolling four frames per (side, phase) instantiation made the classic kernel 156 KB (U = 201)
// to 253 KB (U = 401) of SASS, more than the instruction cache holds for ten warps in different roles.
template <int NS, bool CLASSIC, int SIDE>
__device__ __forceinline__ void rec_phase(const FusedArgs& a, const FusedLayout& f, const SideView& sv, int b, int count,
                                       int t_first, int t_step, bool phase_b, float* v0, float* v1, double& c,
                                       const LabelBits<NS>& lb, int lane, long long* tm) {
  constexpr int S_ = CLASSIC ? 2 : 1, kUpad = NS * kWarp;
  const int R = f.R;
  float m_pend = kNegInf;
  int slot = 0;                // i % R
  unsigned use_par = 0;        // (i / R) & 1: parity of this use of the slot
  float* g_state = a.stateT + ((size_t)b * a.p.T + t_first) * (size_t)(S_ * kUpad);
  double* g_off = a.coff + (size_t)b * a.p.T + t_first;
  const ptrdiff_t g_step = (ptrdiff_t)t_step * (S_ * kUpad);
  // CTCB200_REC_REORDER=1 (off: measured slower) changes the order inside a frame: the loads of the frame's inputs are
  // issued first (their barrier was passed in the previous iteration), the publication of the pre-step state hides their
  // latency, then the NEXT frame's barrier is awaited before the step.
#if CTCB200_REC_REORDER
  if (count > 0) TIMED(3, mbar_wait(sv.full_d, 0u, 1));         // the first frame's inputs are in the ring
#endif
#pragma unroll 1
  for (int i = 0; i < count; ++i) {
#if !CTCB200_REC_REORDER
    TIMED(3, mbar_wait(sv.full_d + slot, use_par, 1));
#endif
    float d[NS];
    const float* dsrc = sv.ringd + slot * kUpad;
#pragma unroll
    for (int j = 0; j < NS; ++j) d[j] = dsrc[j * kWarp + lane];
    const float h = sv.ringh[slot];
    float S[NS], x[NS];
    if (CLASSIC && SIDE == 0) alpha_classic_prepare<NS>(v0, v1, lb, S, x);
    const float* out0 = !CLASSIC ? v0 : (SIDE == 0 ? x : v1);
    if (!phase_b) {
      __syncwarp();
      if (lane == 0) mbar_arrive(sv.empty + slot);             // the inputs are in registers: the slot may be refilled
      // -> global scratch for the other side's phase B, which walks these frames in reverse order (frame i here is its
      // frame count-1-i): with HALF only its odd frames are stored.  A loss-only call stops at the middle: nobody will
      // read the scratch.
      if (a.grad != nullptr && (!(CTCB200_HALF_SCRATCH && f.half) || ((count - 1 - i) & 1))) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          stg_keep(g_state + j * kWarp + lane, out0[j]);
          if (CLASSIC && SIDE == 0) stg_keep(g_state + kUpad + j * kWarp + lane, v1[j]);
        }
        if (lane == 0) *g_off = c;
      }
      g_state += g_step;
      g_off += t_step;
    } else {
      if (i >= R) TIMED(6, mbar_wait(sv.empty + slot, use_par ^ 1u, 2));   // previous frame of the slot is finished
      float* dst = sv.rings + slot * (S_ * kUpad);
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        dst[j * kWarp + lane] = out0[j];
        if (CLASSIC && SIDE == 0) dst[kUpad + j * kWarp + lane] = v1[j];
      }
      if (lane == 0) sv.ringc[slot] = c;
      __syncwarp();
      if (lane == 0) mbar_arrive(sv.full_s + slot);
      CTCB200_TRACE("rec side %d frame %d: state published (slot %d)", SIDE, i, slot);
    }
    const int nslot = (slot + 1 == R) ? 0 : slot + 1;
    const unsigned npar = (nslot == 0) ? (use_par ^ 1u) : use_par;
#if CTCB200_REC_REORDER
    if (i + 1 < count) TIMED(3, mbar_wait(sv.full_d + nslot, npar, 1));     // the next frame's inputs are in the ring
#endif
    if (SIDE == 0) {
      if (CLASSIC) alpha_classic_finish<NS>(v0, v1, S, x, d, h, lane, lb);
      else alpha_step_simplified<NS>(v0, d, h, lane);
    } else {
      if (CLASSIC) beta_step_classic<NS>(v0, v1, d, h, lane, lb);
      else beta_step_simplified<NS>(v0, d, h, lane);
    }
    // lagged offset renormalisation: the warp maximum taken after frame 4n is subtracted two frames later
    const int k = i & (kFusedGroup - 1);
    if (k == 0) m_pend = state_max<NS, CLASSIC>(v0, v1);
    else if (k == 2) apply_offset<NS, CLASSIC>(v0, v1, m_pend, c);
    slot = nslot;
    use_par = npar;
  }
}

__device__ __forceinline__ unsigned long long* bars_of(const SideView& sv, int w) { return sv.bar + w * kMaxRowSlots; }

// Static facts about this lane's label states l = lane*NS + j, in registers for the whole kernel (bit j of each mask):
//   ok    state l emits a token: l < label_length and the label is a valid column.  Its log-probability feeds the
//         recursion (d[t,l], base_loss.py:328-344) whatever the token is -- including a label equal to the blank.
//   nb    ... and that token is not the blank: its occupancy lands in the token's gradient column.  (A real label equal
//         to the blank is undefined input; like the reference's blank_mask override, classic_ctc_loss.py:647-654, and K3,
//         its emission occupancy is dropped from the blank column and from the row total.)
struct LaneLabels {
  unsigned ok_nb;       // ok | nb << 16
  unsigned flags;       // bit 0: ok of state lane*NS-1, bit 1: its nb, bit 3: the label holds a blank-valued entry (uniform
                        // over the CTA)
  __device__ __forceinline__ bool ok(int j) const { return (ok_nb >> j) & 1u; }
  __device__ __forceinline__ bool nb(int j) const { return (ok_nb >> (16 + j)) & 1u; }
  __device__ __forceinline__ bool nb_left() const { return flags & 2u; }
  __device__ __forceinline__ bool any_blank_label() const { return flags & 8u; }
};

// ---- row worker, one phase -------------------------------------------------------------------------------------------
// tok[j] = cleaned label (base_loss.py:395-418) of this lane's states l = lane*NS + j (always a safe column index);
// `ll` holds their static facts (LaneLabels).  `side` is a runtime argument (one code body for both sides keeps the
// instruction footprint inside the instruction cache).
template <int NS, bool CLASSIC, bool PHASE_B, bool TMA, bool BF16, bool HELPERS>
__device__ __forceinline__ void worker_phase(const FusedArgs& a, const FusedLayout& f, const SideView& sv, int side, int b,
                                          int w, int count, int t_first, int t_step, int L, double lossd_mid, float dl,
                                          const int (&tok)[NS], const LaneLabels ll, int lane, long long* tm) {
  constexpr int S = CLASSIC ? 2 : 1, kUpad = NS * kWarp;
  constexpr float kLog2e = 1.4426950408889634f;
  const Problem& p = a.p;
  const int W = f.W, R = f.R, V = p.V, Vp = (V + 3) & ~3, n4 = Vp >> 2;
  // with a helper warp (fused_layout) this warp streams float4 [0, n4m) of its rows and the helper the rest
  const bool helped = HELPERS && f.helpers;
  const int n4m = helped ? fused_helper_split(n4) : n4;
  const int SL = PHASE_B ? f.SL : f.SL + f.XA;      // phase A may own one more row buffer (see fused_layout)
  const int n_my = (count > w) ? (count - w + W - 1) / W : 0;
  static_assert(!BF16 || TMA, "bf16 rows move by TMA only");
  const unsigned row_bytes = (unsigned)V * 4u, in_bytes = BF16 ? (unsigned)V * 2u : row_bytes;
  constexpr bool tma = TMA;
  const int n8 = V >> 3;                               // bf16 rows: 16-byte groups of eight (V % 8 == 0)
  const float* logits_b = p.logits + row_offset(p, b, 0);                                       // fp32 rows
  const char* logits_hb = reinterpret_cast<const char*>(p.logits) + 2 * row_offset(p, b, 0);    // bf16 rows
  float* rowbuf = sv.row + (size_t)w * f.SL * Vp;
  float* rowx = sv.aux_rows + (size_t)w * Vp;
  auto slot_ptr = [&](int q) { return (q < f.SL) ? rowbuf + (size_t)q * Vp : rowx; };
  // Brings logits row `t` into row buffer q; completion is signalled on bars[q] either by the TMA transaction count
  // (one elected lane) or, when rows are not 16-byte aligned, by every lane's cp.async completion.
  auto load_row = [&](int q, int t) {
    float* dst = slot_ptr(q);
    const float* src = logits_b + (size_t)t * p.stride_t;
    if (tma) {
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(bars_of(sv, w) + q, in_bytes);
        if (BF16) bulk_load(dst + (Vp >> 1), logits_hb + 2 * (size_t)t * p.stride_t, in_bytes, bars_of(sv, w) + q);
        else bulk_load(dst, src, row_bytes, bars_of(sv, w) + q);
      }
    } else {
      for (int k = lane; k < V; k += kWarp) fused_cp_async4(dst + k, src + k);
      cp_async_mbar_arrive(bars_of(sv, w) + q);
    }
  };
  unsigned long long* bars = bars_of(sv, w);
  float* stb = sv.stbuf + w * (S * kUpad);
  const float* rowlse_b = a.rowlse + (size_t)b * p.T;
  const double* coff_b = a.coff + (size_t)b * p.T;

  // prologue: the first SL-1 rows are in flight before any is consumed
  if (!tma)   // pad lanes of the (4-float aligned) row buffers never receive data: make them neutral once
    for (int q = 0; q < SL; ++q)
      for (int k = V + lane; k < Vp; k += kWarp) slot_ptr(q)[k] = kNegInf;
  for (int q = 0; q < SL - 1 && q < n_my; ++q) load_row(q, t_first + (w + q * W) * t_step);
  int slot = w % R;        // ring slot of frame i = w + n*W
  unsigned use_par = 0;    // (i / R) & 1
  int rs = 0;              // row buffer of row n (= n % SL)
  unsigned par = 0;        // bit q: parity of the next completion to wait for on row buffer q
  // HALF (see fused_layout): only the rows of odd frames (counted from the middle) were stored.  W is even, so a worker
  // sees frames of one parity: odd workers read stored rows -- fetched a whole iteration ahead -- and derive the state of
  // the frame before; even workers receive that through `fwd` and never touch the global scratch.
  const bool half = CTCB200_HALF_SCRATCH && PHASE_B && !CLASSIC && f.half;
  const bool odd_w = half && (w & 1), even_w = half && !(w & 1);
  // the other side's stored state of frame `tt`: async copy into this worker's staging buffer
  auto fetch_state = [&](int tt) {
    const float* src = a.stateT + ((size_t)b * p.T + tt) * (size_t)(S * kUpad);
    const int n16 = ((CLASSIC && side == 0) ? 1 : S) * (kUpad / 4);      // classic beta rows hold the open plane only
#pragma unroll
    for (int k = 0; k < (S * kUpad / 4 + kWarp - 1) / kWarp; ++k) {
      const int cidx = k * kWarp + lane;
      if (cidx < n16) fused_cp_async16_keep(stb + 4 * cidx, src + 4 * cidx);
    }
    fused_cp_async_commit();
  };
  // scalars of the row produced in phase A, fetched one row ahead (plain loads: written by this CTA in phase A).  The
  // offset of a derived state is that of the stored row it was derived from (the next frame's); the last frame of an
  // odd-length phase has no such row: its other-side state is the initial vector, offset 0.
  float lse_next = 0.0f;
  double cst_next = 0.0;
  if (PHASE_B && n_my > 0) {
    const int t0 = t_first + w * t_step;
    if (!p.input_logprobas) lse_next = rowlse_b[t0];
    if (!even_w) cst_next = coff_b[t0];
    else if (w + 1 < count) cst_next = coff_b[t0 + t_step];
    if (odd_w) fetch_state(t0);
  }
  for (int n = 0; n < n_my; ++n) {
    const int i = w + n * W;
    const int t = t_first + i * t_step;
    float lse = lse_next;
    const double cst = cst_next;
    if (PHASE_B) {
      if (!half) fetch_state(t);     // consumed after the recursion catches up
      if (n + 1 < n_my) {
        const int tn = t + W * t_step;
        if (!p.input_logprobas) lse_next = rowlse_b[tn];
        if (!even_w) cst_next = coff_b[tn];
        else cst_next = (i + W + 1 < count) ? coff_b[tn + t_step] : 0.0;
      }
    }
    CTCB200_TRACE("worker %d phase %d frame %d of %d: start (slot %d par %u)", w, (int)PHASE_B, i, count, slot, use_par);
    // ---- prefetch (phase A): the buffer row n-1 used is free as soon as this iteration starts ----
    if (!PHASE_B && n + SL - 1 < n_my) load_row((rs == 0) ? SL - 1 : rs - 1, t + (SL - 1) * W * t_step);
    const unsigned rowpar = (par >> rs) & 1u;
    TIMED(2, mbar_wait(bars + rs, rowpar, 3));          // the row has landed
    par ^= 1u << rs;
    float* row = slot_ptr(rs);
    float4* row4 = reinterpret_cast<float4*>(row);
    const uint4* rowh = reinterpret_cast<const uint4*>(row + (Vp >> 1));                  // bf16 row: upper half of the buffer
    const unsigned short* rowh16 = reinterpret_cast<const unsigned short*>(row + (Vp >> 1));
    auto in_at = [&](int k) { return BF16 ? __uint_as_float((unsigned)rowh16[k] << 16) : row[k]; };   // raw input element k

    // ---- stage 1: row log-sum-exp (phase A; one pass, the row chunk lives in registers) and the label gather ----
#ifdef CTCB200_FUSED_TIMING
    const long long t_s0 = clock64();
#endif
    if (!PHASE_B && !p.input_logprobas) {
      float m_run = kNegInf, s_run = 0.0f;
      if (BF16 && n4 <= kWarp) {   // narrow bf16 rows (V <= 128): one group of eight per lane
        float4 v0 = make_float4(kNegInf, kNegInf, kNegInf, kNegInf), v1 = v0;
        if (lane < n8) widen8(rowh[lane], v0, v1);
        m_run = fmaxf(fmaxf(fmaxf(v0.x, v0.y), fmaxf(v0.z, v0.w)), fmaxf(fmaxf(v1.x, v1.y), fmaxf(v1.z, v1.w)));
        const float mn0 = (m_run == kNegInf || m_run == INFINITY) ? 0.0f : m_run;
        s_run = hsum4(exp4_shifted(v0, mn0)) + hsum4(exp4_shifted(v1, mn0));
      } else if (n4 <= kWarp) {   // narrow rows (V <= 128): one float4 per lane
        const float4 v = (lane < n4) ? row4[lane] : make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
        m_run = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        const float mn0 = (m_run == kNegInf || m_run == INFINITY) ? 0.0f : m_run;
        s_run = hsum4(exp4_shifted(v, mn0));
      } else {
        // one 8 x float4 chunk per lane; full chunks need no bounds checks (V = 1024 is exactly one of them)
        auto chunk = [&](auto checked, int base) {
          constexpr bool kChecked = decltype(checked)::value;
          float4 v[8];
          if (BF16) {      // the same 1024 elements as four groups of eight per lane
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int g = (base >> 1) + u * kWarp + lane;
              v[2 * u] = v[2 * u + 1] = make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
              if (!kChecked || g < n8) widen8(rowh[g], v[2 * u], v[2 * u + 1]);
            }
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int c4 = base + u * kWarp + lane;
              v[u] = (!kChecked || c4 < n4m) ? row4[c4] : make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
            }
          }
          float pm[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) pm[u] = fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w));
          const float cm = fmaxf(fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3])), fmaxf(fmaxf(pm[4], pm[5]), fmaxf(pm[6], pm[7])));
          const float mn = fmaxf(m_run, cm);
          const float mn0 = (mn == kNegInf || mn == INFINITY) ? 0.0f : mn;     // tf.reduce_logsumexp convention
          s_run *= ex2_approx((m_run - mn0) * kLog2e);                         // 0 * 0 when m_run == -inf
          // (v - max) first, then the scale: a fused v*log2e - max*log2e would lose the exact 0 for |logit| ~ 1e10
          float ps[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) ps[u] = hsum4(exp4_shifted(v[u], mn0));
          s_run += ((ps[0] + ps[1]) + (ps[2] + ps[3])) + ((ps[4] + ps[5]) + (ps[6] + ps[7]));
          m_run = mn;
        };
        int base = 0;
        for (; base + 8 * kWarp <= n4m; base += 8 * kWarp) chunk(std::false_type{}, base);
        if (base < n4m) chunk(std::true_type{}, base);
      }
      float M = warp_max(m_run), h_max = kNegInf, h_sum = 0.0f;
      if (helped) {      // the helper's part of the row: its maximum and its sum of exponentials relative to that
        TIMED(2, mbar_wait(sv.hdone + w * kMaxRowSlots + rs, rowpar, 9));
        h_max = sv.hres[(w * kMaxRowSlots + rs) * 2];
        h_sum = sv.hres[(w * kMaxRowSlots + rs) * 2 + 1];
        M = fmaxf(M, h_max);
      }
      const float M0 = (M == kNegInf || M == INFINITY) ? 0.0f : M;
      const float mr0 = (m_run == kNegInf || m_run == INFINITY) ? 0.0f : m_run;
      float sum = warp_sum(s_run * ex2_approx((mr0 - M0) * kLog2e));
      if (helped && h_sum > 0.0f) {
        const float hm0 = (h_max == kNegInf || h_max == INFINITY) ? 0.0f : h_max;
        sum += h_sum * ex2_approx((hm0 - M0) * kLog2e);
      }
      lse = fmaf(lg2_approx(sum), 0.6931471805599453f, M0);    // sum is in [1, V]: no denormal / range handling needed
      if (lane == 0 && a.grad != nullptr) a.rowlse[(size_t)b * p.T + t] = lse;
    }
#ifdef CTCB200_FUSED_TIMING
    tm[6] += clock64() - t_s0;   // workers: row statistics
    const long long t_g0 = clock64();
#endif
    if (!PHASE_B && i >= R) TIMED(4, mbar_wait(sv.empty + slot, use_par ^ 1u, 4));   // slot consumed by the recursion
    float dd[NS];
    {
      float* dst = sv.ringd + slot * kUpad;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        dd[j] = ll.ok(j) ? in_at(tok[j]) - lse : kNegInf;
        dst[j * kWarp + lane] = dd[j];
      }
    }
    const float h = in_at(p.blank) - lse;
    if (lane == 0) sv.ringh[slot] = h;
    __syncwarp();
    if (lane == 0) mbar_arrive(sv.full_d + slot);
    CTCB200_TRACE("worker %d phase %d frame %d: inputs published", w, (int)PHASE_B, i);
#ifdef CTCB200_FUSED_TIMING
    tm[3] += clock64() - t_g0;   // workers: ring wait + gather + publish
#endif

    // ---- HALF, odd frames: one recursion step from the stored state with this frame's inputs gives the other side's
    // state of the frame before (beta[t] from beta[t+1] on the alpha side, alpha[t+1] from alpha[t] on the beta side).
    if (odd_w) {
      TIMED(7, fused_cp_async_wait_all());
      __syncwarp();
      const int sp = slot ? slot - 1 : R - 1;                       // ring slot of frame i-1
      const unsigned pp = slot ? use_par : use_par ^ 1u;            // parity of that use of it
      if (i - 1 >= R) mbar_wait(sv.empty + sp, pp ^ 1u, 7);            // the previous user of its fwd vector (frame i-1-R) is done
      float st[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) st[j] = stb[j * kWarp + lane];
      if (side == 0) beta_step_simplified<NS>(st, dd, h, lane);
      else alpha_step_simplified<NS>(st, dd, h, lane);
      float* fw = sv.fwd + (sp >> 1) * kUpad;
#pragma unroll
      for (int j = 0; j < NS; ++j) fw[j * kWarp + lane] = st[j];
      __syncwarp();
      if (lane == 0) mbar_arrive(sv.fwd_full + sp);
      CTCB200_TRACE("worker %d frame %d: derived state forwarded to slot %d", w, i, sp);
    }

    // ---- stage 1b (phase B): the dense softmax part of the gradient row, in place, while the recursion catches up.
    // d loss/d logit = d_loss * (softmax * sum_k occ - occ); sum_k occ is 1 for every frame of a feasible sample (it is
    // the total probability of being somewhere), so the softmax term does not have to wait for the occupancies.
#ifdef CTCB200_FUSED_TIMING
    const long long t_p0 = clock64();
#endif
    if (PHASE_B) {
      // batches of four float4 per lane with all loads first: left to itself the compiler runs one float4 at a time
      // through the same four registers (load -> 2^x -> store, ~100 dependent cycles each)
      auto softmax_in_place = [&](auto scaled) {
        constexpr bool kScaled = decltype(scaled)::value;
        const float2 dl2 = make_float2(dl, dl);
        auto fin = [&](float4 v) {
          float4 e = exp4_shifted(v, lse);
          if (kScaled) {
            const float2 lo = __fmul2_rn(make_float2(e.x, e.y), dl2), hi = __fmul2_rn(make_float2(e.z, e.w), dl2);
            e = make_float4(lo.x, lo.y, hi.x, hi.y);
          }
          return e;
        };
        if (BF16) {
          // Widening in place: group g (eight bf16 in the upper half) becomes float4 2g, 2g+1 of the lower part.  Elements
          // are taken in increasing order, every batch loads before it stores, and the fp32 image of elements < e ends at
          // byte 4e <= 2V + 2e, where the unread bf16 data begins: a batch can only overwrite what it has already read.
          int g = lane;
          for (; (g - lane) + 2 * kWarp <= n8; g += 2 * kWarp) {      // full batches: warp-uniform trip count
            const uint4 q0 = rowh[g], q1 = rowh[g + kWarp];
            float4 v[4];
            widen8(q0, v[0], v[1]);
            widen8(q1, v[2], v[3]);
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = fin(v[u]);
            __syncwarp();
            row4[2 * g] = v[0];
            row4[2 * g + 1] = v[1];
            row4[2 * (g + kWarp)] = v[2];
            row4[2 * (g + kWarp) + 1] = v[3];
          }
          for (; g - lane < n8; g += kWarp) {      // warp-uniform trip count (the barriers inside need every lane)
            float4 v0, v1;
            if (g < n8) widen8(rowh[g], v0, v1);
            __syncwarp();
            if (g < n8) {
              row4[2 * g] = fin(v0);
              row4[2 * g + 1] = fin(v1);
            }
          }
          __syncwarp();
        } else {
          int c4 = lane;
          for (; c4 + 3 * kWarp < n4m; c4 += 4 * kWarp) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = row4[c4 + u * kWarp];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = fin(v[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) row4[c4 + u * kWarp] = v[u];
          }
          for (; c4 < n4m; c4 += kWarp) row4[c4] = fin(row4[c4]);
        }
      };
      if (dl == 1.0f) softmax_in_place(std::false_type{});     // no upstream gradient (or ones): skip the scaling
      else softmax_in_place(std::true_type{});
    }

#ifdef CTCB200_FUSED_TIMING
    if (PHASE_B) tm[8] += clock64() - t_p0;
#endif
    // ---- prefetch (phase B): row n+SL-1 goes into the buffer row n-1 used; its TMA store must have drained first,
    // which is why this sits after the softmax pass rather than at the top of the iteration ----
    if (PHASE_B && n + SL - 1 < n_my) {
      if (tma && lane == 0) bulk_store_wait_read();
      load_row((rs == 0) ? SL - 1 : rs - 1, t + (SL - 1) * W * t_step);
    }

    // ---- stage 2 (phase B): occupancies of the frame, scattered into the row; then the row leaves by TMA ----
    if (PHASE_B) {
      CTCB200_TRACE("worker %d frame %d: softmax + prefetch done, waiting for the state", w, i);
      TIMED(5, mbar_wait(sv.full_s + slot, use_par, 5));          // the running side's state for this frame is published
      if (helped) TIMED(5, mbar_wait(sv.hdone + w * kMaxRowSlots + rs, rowpar, 10));   // ... and the helper's part of the row is softmax
      CTCB200_TRACE("worker %d frame %d: state there", w, i);
      const float* other = stb;                                  // the other side's state for this frame
      if (!even_w) {
        TIMED(7, fused_cp_async_wait_all());
        __syncwarp();
        // the stored state row is dead now: drop it from L2 instead of letting it be written back (rows are 128-byte
        // multiples at 128-byte aligned offsets of the workspace)
        const char* dead_row = reinterpret_cast<const char*>(a.stateT + ((size_t)b * p.T + t) * (size_t)(S * kUpad));
        if (lane < S * kUpad * 4 / 128) l2_discard128(dead_row + lane * 128);
      } else {
        float* fw = sv.fwd + (slot >> 1) * kUpad;
        __syncwarp();                                              // (every lane is past its part of the softmax pass)
        if (i + 1 < count) {
          TIMED(7, mbar_wait(sv.fwd_full + slot, use_par, 6));      // derived by the worker of frame i+1
        } else {
          // last frame of an odd-length phase: the other side's INITIAL vector (beta[n] = one-hot(label_length) on the
          // alpha side, alpha[0] = one-hot(0) on the beta side)
          if (i >= R) mbar_wait(sv.empty + slot, use_par ^ 1u, 8);
#pragma unroll
          for (int j = 0; j < NS; ++j) fw[j * kWarp + lane] = (lane * NS + j == (side == 0 ? L : 0)) ? 0.0f : kNegInf;
          __syncwarp();
        }
        other = fw;
      }
#ifdef CTCB200_FUSED_TIMING
      const long long t_o0 = clock64();
#endif
      CTCB200_TRACE("worker %d frame %d: other state ready", w, i);
      const float* ring_state = sv.rings + slot * (S * kUpad);
      const float K = (float)(lossd_mid + sv.ringc[slot] + cst);   // loss + both renormalisation offsets
      const float* A = (side == 0) ? ring_state : other;       // alpha[t]
      const float* Bn = (side == 0) ? other : ring_state;      // beta[t+1]
      float occ[NS], occ_stay[NS];
      if (!CLASSIC) {
        float a0[NS], b0[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          a0[j] = A[j * kWarp + lane];
          b0[j] = Bn[j * kWarp + lane];
        }
        if (odd_w && n + 1 < n_my) {      // the staging buffer is in registers now: fetch this worker's next stored row
          __syncwarp();
          fetch_state(t + W * t_step);
        }
        float bx = __shfl_down_sync(kFull, b0[0], 1);
        if (lane == 31) bx = kNegInf;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          const float bn = (j < NS - 1) ? b0[j + 1] : bx;
          occ[j] = ex2_approx((K + (a0[j] + dd[j] + bn)) * kLog2e);   // emit label[l]: simplified_ctc_loss.py:503-510
          if (!ll.ok(j)) occ[j] = 0.0f;
        }
      } else {
        // alpha rows: plane 0 = x[l] (the diagonal carrier, see rec_phase), plane 1 = A[l,1]; beta rows: B[l,1]
        float xa[NS], a1[NS], b1[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          xa[j] = A[j * kWarp + lane];
          a1[j] = A[kUpad + j * kWarp + lane];
          b1[j] = Bn[j * kWarp + lane];
        }
        float bx = __shfl_down_sync(kFull, b1[0], 1);
        if (lane == 31) bx = kNegInf;
        float d_left = __shfl_up_sync(kFull, dd[NS - 1], 1);
        if (lane == 0) d_left = kNegInf;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          const float bn = (j < NS - 1) ? b1[j + 1] : bx;
          // diagonal step emitting label[l] (classic_ctc_loss.py:629-639); open -> open is barred on a repeat (x)
          occ[j] = ex2_approx((K + (dd[j] + xa[j] + bn)) * kLog2e);
          if (!ll.ok(j)) occ[j] = 0.0f;
          // horizontal step re-emitting label[l-1] from the open state (classic_ctc_loss.py:617-626)
          const float dp = (j > 0) ? dd[j - 1] : d_left;
          occ_stay[j] = ex2_approx((K + (a1[j] + dp + b1[j])) * kLog2e);
          if (!((j > 0) ? ll.nb(j > 0 ? j - 1 : 0) : ll.nb_left())) occ_stay[j] = 0.0f;   // a blank is never re-emitted
        }
      }
      // Scatter the occupancies into the row: row[token] -= d_loss * occ.
      if (CLASSIC) {   // the horizontal (stay) occupancy of state l+1 re-emits label[l]: same target as state l's move
        float s_next = __shfl_down_sync(kFull, occ_stay[0], 1);
        if (lane == 31) s_next = 0.0f;
#pragma unroll
        for (int j = 0; j < NS; ++j) occ[j] += (j < NS - 1) ? occ_stay[j + 1] : s_next;
      }
#ifdef CTCB200_FUSED_TIMING
      const long long t_o1 = clock64();
      tm[9] += t_o1 - t_o0;
#endif
      // The blank's occupancy is the complement of the others: every alignment emits exactly one symbol per frame, so
      // sum_k occ[t,k] = 1 (simplified_ctc_loss.py:498-501 / classic_ctc_loss.py:608-614 compute the same number as
      // exp(loss + h + logsumexp_l(alpha + beta)); the complement needs one warp sum instead of a second log-sum-exp).
      CTCB200_TRACE("worker %d frame %d: occupancies computed", w, i);
      float osum = 0.0f;
#pragma unroll
      for (int j = 0; j < NS; ++j) osum += occ[j];
#ifdef CTCB200_WATCHDOG_VERBOSE
      {
        const unsigned f0 = __shfl_sync(kFull, ll.flags, 0), any = __ballot_sync(kFull, ll.flags & 8u);
        if (blockIdx.x == 0 && lane == 0 && i < 2) printf("trace warp %d: flags lane0 %x ballot(bit3) %08x ok_nb %x\n", (int)(threadIdx.x >> 5), f0, any, ll.ok_nb);
      }
#endif
#ifndef CTCB200_DBG_NO_BLANKBLOCK
      if (ll.any_blank_label()) {
        // Undefined input (a real label equal to the blank): the reference overrides the blank column with the horizontal
        // term alone, so the emission occupancy of such states vanishes from the row and from its total -- the softmax
        // part, already in the row with weight 1, shrinks accordingly (same numbers as the staged kernel K3).
        float drop = 0.0f;
#pragma unroll
        for (int j = 0; j < NS; ++j)
          if (ll.ok(j) && !ll.nb(j)) drop += occ[j];
        drop = warp_sum(drop);
        if (drop != 0.0f) {
          const float keep = 1.0f - drop;
          for (int c4 = lane; c4 < n4; c4 += kWarp) {
            float4 v = row4[c4];
            row4[c4] = make_float4(v.x * keep, v.y * keep, v.z * keep, v.w * keep);
          }
          __syncwarp();
        }
      }
#endif
      CTCB200_TRACE("worker %d frame %d: before scatter", w, i);
      // Scatter: row[token] -= d_loss * occ.  Shared-memory float atomics were measured against shuffle-combined,
      // conflict-mask-guarded and statically ranked conflict-free read-modify-write passes (round 1) and against a static
      // leader / follower plan that needs no atomics at all (round 2: 671 vs 573 us at B=256, 347 vs 323 us at B=32); the
      // atomics won every time at V = 1024 and tied at V = 5000.
#pragma unroll
      for (int j = 0; j < NS; ++j)
#ifdef CTCB200_DBG_PLAIN_SCATTER
        if (ll.nb(j) && occ[j] > 0.0f) row[tok[j]] -= dl * occ[j];
#else
        if (ll.nb(j) && occ[j] > 0.0f) atomicAdd(&row[tok[j]], -dl * occ[j]);
#endif
      CTCB200_TRACE("worker %d frame %d: atomics issued", w, i);
      __syncwarp();
#ifdef CTCB200_FUSED_TIMING
      const long long t_o2 = clock64();
      tm[10] += t_o2 - t_o1;
#endif
      CTCB200_TRACE("worker %d frame %d: scattered", w, i);
      const float occ_blank = 1.0f - warp_sum(osum);
      if (lane == 0) row[p.blank] -= dl * occ_blank;
      __syncwarp();
      CTCB200_TRACE("worker %d frame %d: blank done", w, i);
      float* gdst = a.grad + row_offset(p, b, t);
      if (BF16 && p.grad_bf16) {
        // narrow the finished row in place (element k -> bytes 2k..2k+1: again only already-read bytes are overwritten)
        uint4* out8 = reinterpret_cast<uint4*>(row);
        for (int g = lane; g - lane < n8; g += kWarp) {
          float4 v0, v1;
          if (g < n8) { v0 = row4[2 * g]; v1 = row4[2 * g + 1]; }
          __syncwarp();
          if (g < n8) out8[g] = make_uint4(pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w));
          __syncwarp();
        }
        if (lane == 0) {
          fence_proxy_async();
          bulk_store(reinterpret_cast<char*>(a.grad) + 2 * row_offset(p, b, t), row, in_bytes);
        }
      } else if (tma) {
        if (lane == 0) {
          fence_proxy_async();                                  // generic-proxy writes -> visible to the TMA store
          bulk_store(gdst, row, row_bytes);
        }
      } else {
        for (int k = lane; k < V; k += kWarp) gdst[k] = row[k];
        __syncwarp();
      }
      if (lane == 0) mbar_arrive(sv.empty + slot);              // ring slot and stb are free again
      CTCB200_TRACE("worker %d frame %d: row done", w, i);
#ifdef CTCB200_FUSED_TIMING
      tm[11] += clock64() - t_o2;
#endif
    }
    slot += W;
    if (slot >= R) { slot -= R; use_par ^= 1u; }
    if (++rs == SL) rs = 0;
  }
  if (PHASE_B && tma && lane == 0) bulk_store_wait_read();      // shared memory must outlive the stores reading it
}

// ---- row helper, one phase (see fused_layout) ---------------------------------------------------------------------------
// The helper of worker w walks the same rows through the same row buffers, waits on the same landing barrier and handles
// float4 [fused_helper_split(n4), n4) of every row: their maximum and sum of exponentials in phase A (answered through
// `hres`), their softmax in place in phase B -- but only once the worker has gathered the frame's label columns from the raw
// row (it says so on the ring slot's full_d barrier, the one the recursion warp waits on).  `hdone[buffer]` tells the
// worker the helper is through; the buffer is not reloaded before the worker has seen that, so the helper can never meet
// a later phase of the landing barrier.  fp32 rows moved by TMA only (no pad lanes, no in-place widening).
template <bool PHASE_B>
__device__ __forceinline__ void helper_phase(const FusedArgs& a, const FusedLayout& f, const SideView& sv, int b, int w,
                                             int count, int t_first, int t_step, float dl, int lane) {
  constexpr float kLog2e = 1.4426950408889634f;
  const Problem& p = a.p;
  const int W = f.W, Vp = (p.V + 3) & ~3, n4 = Vp >> 2, n4m = fused_helper_split(n4);
  const int SL = PHASE_B ? f.SL : f.SL + f.XA;
  const int n_my = (count > w) ? (count - w + W - 1) / W : 0;
  if (!PHASE_B && p.input_logprobas) return;        // no row statistics are taken: the worker does not wait either
  float* rowbuf = sv.row + (size_t)w * f.SL * Vp;
  float* rowx = sv.aux_rows + (size_t)w * Vp;
  unsigned long long* bars = bars_of(sv, w);
  unsigned long long* hdone = sv.hdone + w * kMaxRowSlots;
  float* hres = sv.hres + w * kMaxRowSlots * 2;
  const float* rowlse_b = a.rowlse + (size_t)b * p.T;
  const bool need_lse = PHASE_B && !p.input_logprobas;
  float lse_next = (need_lse && n_my > 0) ? rowlse_b[t_first + w * t_step] : 0.0f;
  int rs = 0;
  unsigned par = 0;
  int slot = w % f.R;       // ring slot and use parity of frame i = w + n*W, as in worker_phase
  unsigned use_par = 0;
  for (int n = 0; n < n_my; ++n) {
    const int t = t_first + (w + n * W) * t_step;
    const float lse = lse_next;
    if (need_lse && n + 1 < n_my) lse_next = rowlse_b[t + W * t_step];
    mbar_wait(bars + rs, (par >> rs) & 1u, 11);       // the row has landed
    par ^= 1u << rs;
    if (PHASE_B) mbar_wait(sv.full_d + slot, use_par, 12);     // ... and the worker has taken the raw label columns from it
    float4* row4 = reinterpret_cast<float4*>((rs < f.SL) ? rowbuf + (size_t)rs * Vp : rowx);
    if (!PHASE_B) {
      float m_run = kNegInf, s_run = 0.0f;
      auto chunk = [&](auto checked, int base) {       // the worker's chunk (worker_phase), on the helper's part of the row
        constexpr bool kChecked = decltype(checked)::value;
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c4 = base + u * kWarp + lane;
          v[u] = (!kChecked || c4 < n4) ? row4[c4] : make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
        }
        float pm[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) pm[u] = fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w));
        const float cm = fmaxf(fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3])), fmaxf(fmaxf(pm[4], pm[5]), fmaxf(pm[6], pm[7])));
        const float mn = fmaxf(m_run, cm);
        const float mn0 = (mn == kNegInf || mn == INFINITY) ? 0.0f : mn;
        s_run *= ex2_approx((m_run - mn0) * kLog2e);
        float ps[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) ps[u] = hsum4(exp4_shifted(v[u], mn0));
        s_run += ((ps[0] + ps[1]) + (ps[2] + ps[3])) + ((ps[4] + ps[5]) + (ps[6] + ps[7]));
        m_run = mn;
      };
      int base = n4m;
      for (; base + 8 * kWarp <= n4; base += 8 * kWarp) chunk(std::false_type{}, base);
      if (base < n4) chunk(std::true_type{}, base);
      const float M = warp_max(m_run);
      const float M0 = (M == kNegInf || M == INFINITY) ? 0.0f : M;
      const float mr0 = (m_run == kNegInf || m_run == INFINITY) ? 0.0f : m_run;
      const float sum = warp_sum(s_run * ex2_approx((mr0 - M0) * kLog2e));
      if (lane == 0) {
        hres[2 * rs] = M;
        hres[2 * rs + 1] = sum;
      }
    } else {
      const float2 dl2 = make_float2(dl, dl);
      const bool scaled = dl != 1.0f;
      auto fin = [&](float4 v) {
        float4 e = exp4_shifted(v, lse);
        if (scaled) {
          const float2 lo = __fmul2_rn(make_float2(e.x, e.y), dl2), hi = __fmul2_rn(make_float2(e.z, e.w), dl2);
          e = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        return e;
      };
      int c4 = n4m + lane;
      for (; c4 + 3 * kWarp < n4; c4 += 4 * kWarp) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = row4[c4 + u * kWarp];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = fin(v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) row4[c4 + u * kWarp] = v[u];
      }
      for (; c4 < n4; c4 += kWarp) row4[c4] = fin(row4[c4]);
      fence_proxy_async();          // this warp's part of the row leaves by the worker's TMA store
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(hdone + rs);
    slot += W;
    if (slot >= f.R) { slot -= f.R; use_par ^= 1u; }
    if (++rs == SL) rs = 0;
  }
}

// ---- the kernel body ----------------------------------------------------------------------------------------------------
// SPLIT = false: one CTA per utterance holds both sides (2 x (1 recursion warp + W workers), W <= 4, two CTAs per SM): the
//                plan for full batches, where the SMs are shared by two utterances and HBM bandwidth is the limit.
// SPLIT = true : a cluster of two CTAs per utterance, one side each (1 recursion warp + W workers, W <= 8): the plan for
//                small batches (B <= 148, e.g. the per-GPU slice of a batch sharded over 4 or 8 GPUs), where the T-step chain
//                is the runtime.  Every side gets twice the row workers, its own shared memory and a less crowded
//                scheduler; the sides only talk at the middle (state vectors through distributed shared memory, two cluster
//                barriers) and through the global scratch the other side reads back in phase B.
// HELPERS: the kernel carries the row helpers' code (a compile-time switch: the extra code cost the narrow-row kernels 1 % at
// B = 256 and 3 % in the split plan of the classic variant when it was merely branched around -- profiles/r2_ab_helper_code_cost.txt).
template <int NS, bool CLASSIC, bool TMA, bool SPLIT, bool BF16, bool HELPERS>
__device__ __forceinline__ void fused_body(const FusedArgs& a) {
  static_assert(!HELPERS || (TMA && !SPLIT && !BF16), "row helpers: one-CTA plans of fp32 rows moved by TMA");
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int S = CLASSIC ? 2 : 1, kUpad = NS * kWarp;
  const Problem& p = a.p;
  const FusedLayout f = fused_layout(p.V, kUpad, S, a.W, a.SL, a.XA, a.R, SPLIT ? 1 : 2, a.half, HELPERS ? a.helpers : 0);
  const int b = SPLIT ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = f.W;
  // Warp -> (side, role); role 0 = recursion warp, 1..W = row workers.  Other placements were measured on B200 (the two
  // recursion warps on the highest warp ids, on one scheduler, or rotated by the CTA index so co-resident CTAs do not
  // stack roles): none beat this one (simplified +-1 %, classic 3-8 % slower).  In split mode the side is the CTA's rank
  // in its cluster.
  // Split mode with `rec_alone`: warps 4, 8, ... (which share the recursion warp's scheduler) stay idle, so the T-step
  // chain of warp 0 has an issue port to itself; role -1 = idle (takes part in the barriers only).
  // (For the small one-CTA plans of wide rows -- W = 1 or 2, two CTAs per SM -- shifting the roles of second-wave CTAs by
  // one warp, so that recursion warps and workers of co-resident CTAs do not meet on the same scheduler, was measured at
  // V = 5000: 5.41 ms against 5.06 ms without.  Not done.)
  // With row helpers a side is 1 + 2W warps: roles W+1 .. 2W are the helpers of workers 1 .. W.
  const int wps = f.helpers ? 2 * W + 1 : W + 1;      // warps per side (one-CTA plans)
  const int side = SPLIT ? (int)(blockIdx.x & 1u) : warp / wps;
  const int role = !SPLIT ? warp % wps : !a.rec_alone ? warp : (warp == 0) ? 0 : ((warp & 3) == 0) ? -1 : warp - (warp >> 2);
  const int my = SPLIT ? 0 : side;        // index of this side's block in THIS CTA's shared memory
  const int L = utt_label_len(p, b), n_t = utt_frames(p, b), M = n_t >> 1;
  const float dl = a.d_loss ? a.d_loss[b] : 1.0f;

  // exchange slot of a side (S*Upad floats): a dedicated buffer, or the head of that side's input ring (see fused_layout)
  auto xch_of = [&](int s2) {
    return reinterpret_cast<float*>(smem + (f.xch_aliased ? f.off_side0 + s2 * f.side_bytes + f.s_ringd : f.off_xch + s2 * (S * kUpad * 4)));
  };
  double* xoff = reinterpret_cast<double*>(smem + f.off_xoff);

  // Every mbarrier is initialised exactly once, here, long before its first use; each phase has its own set.  (Round 2
  // found that RE-initialising the barriers at the middle is fragile: depending on register allocation the compiler
  // emitted the side-0 mbarrier.init as a plain 64-bit shared store, and arrivals of phase B that followed the closing
  // __syncthreads were lost on exactly those barriers -- wrong recursion inputs on short utterances, a deadlock on long
  // ones.  With one set per phase nothing is ever re-initialised while the kernel runs.)
  if (tid == 0) {
    for (int s2 = 0; s2 < (SPLIT ? 1 : 2); ++s2) {
      const SideView v = side_view(smem, f, s2, 0);      // the two sets are contiguous
      for (int k = 0; k < 2 * 4 * f.R; ++k) mbar_init(v.full_d + k, 1u);
      for (int k = 0; k < 2 * W * kMaxRowSlots; ++k) mbar_init(v.bar + k, TMA ? 1u : (unsigned)kWarp);
      if (HELPERS && f.helpers)
        for (int k = 0; k < 2 * W * kMaxRowSlots; ++k) mbar_init(v.hdone + k, 1u);
    }
    fence_mbar_init();
  }
  __syncthreads();

  // this lane's labels and their static facts, in registers for the whole kernel (see LaneLabels)
  int tok[NS];
  LaneLabels ll;
  ll.ok_nb = 0u; ll.flags = 0u;
  auto emits = [&](int l, int& t) {       // does state l emit a token?  t <- a column index that is always safe
    const int raw = utt_token(p, b, l, L);
    const bool in_range = raw >= 0 && raw < p.V;
    t = in_range ? raw : p.blank;
    return l >= 0 && l < L && in_range;
  };
#pragma unroll
  for (int j = 0; j < NS; ++j)
    if (emits(lane * NS + j, tok[j])) ll.ok_nb |= (tok[j] != p.blank) ? (0x10001u << j) : (1u << j);
  {
    int tl;
    if (emits(lane * NS - 1, tl)) ll.flags |= (tl != p.blank) ? 3u : 1u;
  }
  {
    // "Some label of this utterance equals the blank" is gathered through a shared-memory word and plain barriers.  (NOT
    // __syncthreads_or: its result is only consumed deep inside the workers' phase-B loop, and nvcc 12.9 sank the
    // BAR.RED.OR itself down to that use -- a CTA-wide barrier executed per row by the worker warps alone, which
    // deadlocked long utterances and let the recursion warps of short ones run ahead of their inputs.)
    volatile int* cta_flag = reinterpret_cast<volatile int*>(smem + f.off_flag);
    if (tid == 0) *cta_flag = 0;
    __syncthreads();
    bool blank_label = false;
#pragma unroll
    for (int j = 0; j < NS; ++j) blank_label |= ll.ok(j) && !ll.nb(j);
    if (blank_label) *cta_flag = 1;
    __syncthreads();
    if (*cta_flag != 0) ll.flags |= 8u;
  }

  long long tm[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  (void)tm;
#ifdef CTCB200_FUSED_TIMING
  const long long t_start = clock64();
#endif
  // ------------------------------------------------ phase A, the middle, phase B -----------------------------------------
  // The role branch sits OUTSIDE the phase loop: the recursion warps never carry the workers' label registers and the
  // workers never carry the recursion state (with one loop body serving both roles every one of them stayed live through
  // the other role's code, and the recursion loops of the classic variant reloaded eight spilled registers per frame).
  bool dead = false;                 // no feasible alignment: loss = +inf, zero gradient
  double lossd_mid = 0.0;            // -log Z, known after phase A
#ifdef CTCB200_FUSED_TIMING
  long long t_mid = 0;
#endif
  auto phase_frames = [&](int ph, int& cnt, int& tf) {
    if (ph == 0) {
      cnt = (side == 0) ? M : n_t - M;
      tf = (side == 0) ? 0 : n_t - 1;
    } else {
      cnt = (side == 0) ? n_t - M : M;
      tf = (side == 0) ? M : M - 1;
    }
  };
  const int ts = (side == 0) ? 1 : -1;
  // The middle, executed by every thread: log Z from the two exchange vectors, then the barrier that lets phase B reuse
  // the shared memory.  Returns false when the call ends here (loss-only call, or an infeasible sample).
  auto middle = [&]() {
#ifdef CTCB200_FUSED_TIMING
    tm[0] = clock64() - t_start;
    t_mid = clock64();
#endif
    LseAcc zacc;
    double off_sum;
    if (!SPLIT) {
      __syncthreads();     // also makes phase A's global scratch visible to the whole CTA
      const float *xa = xch_of(0), *xb = xch_of(1);
      for (int q = lane; q < S * kUpad; q += kWarp) zacc.add(xa[q] + xb[q]);
      off_sum = xoff[0] + xoff[1];
    } else {
      // The other side lives in the peer CTA of the cluster: its exchange vector is read through distributed shared
      // memory.  The cluster barrier (release / acquire) also publishes phase A's global scratch -- row log-sum-exps,
      // stored states, offsets -- to the peer, which reads it back in phase B.
      cluster_sync_all();
      const float* own = xch_of(0);
      const unsigned peer = cluster_map(smem_u32(own), (unsigned)(side ^ 1));
      for (int q = lane; q < S * kUpad; q += kWarp) zacc.add(own[q] + ld_dsmem_f32(peer + 4u * q));
      const double o_own = xoff[0], o_peer = ld_dsmem_f64(cluster_map(smem_u32(xoff), (unsigned)(side ^ 1)));
      off_sum = (side == 0) ? o_own + o_peer : o_peer + o_own;     // the same double on both sides
    }
    const float lz = zacc.warp_result();
    dead = (lz == kNegInf);
    lossd_mid = -((double)lz + off_sum);
    if (!SPLIT) __syncthreads();
    else cluster_sync_all();          // the peer has read this CTA's exchange slot: phase B may overwrite it
    if (a.grad == nullptr) {
      // Loss-only call (the forward pass of a training step, or evaluation): -log Z is known at the middle, so the call
      // costs phase A alone -- half the chain, one read of the logits, nothing written but the loss.
      if (side == 0 && tid == 0) a.loss[b] = dead ? INFINITY : (float)lossd_mid;
      return false;
    }
    return !dead;
  };
#ifdef CTCB200_FUSED_TIMING
  auto dump_timing = [&]() {
    tm[1] = clock64() - t_mid;
    if (a.dbg != nullptr && lane == 0)
      for (int q = 0; q < 12; ++q) a.dbg[((size_t)b * (2 * (W + 1)) + side * (W + 1) + role) * 12 + q] = tm[q];
  };
#endif

  if (role == 0) {
    // ---- a recursion warp: state in registers ----
    float v0[NS], v1[NS];
    double c = 0.0;
    LabelBits<NS> lb;
    if (CLASSIC) lb = make_label_bits<NS>(p, b, L, lane);
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int l = lane * NS + j;
      if (side == 0) {            // alpha[0] = log one_hot(state 0, closed)
        v0[j] = (l == 0) ? 0.0f : kNegInf;
        v1[j] = kNegInf;
      } else {                    // beta[n_t] = log one_hot(label_length), both states
        v0[j] = (l == L) ? 0.0f : kNegInf;
        v1[j] = v0[j];
      }
    }
#pragma unroll 1
    for (int ph = 0; ph < 2; ++ph) {
      int cnt, tf;
      phase_frames(ph, cnt, tf);
      const SideView sv = side_view(smem, f, my, ph);      // this phase's barrier set
      if (side == 0) rec_phase<NS, CLASSIC, 0>(a, f, sv, b, cnt, tf, ts, ph == 1, v0, v1, c, lb, lane, tm);
      else rec_phase<NS, CLASSIC, 1>(a, f, sv, b, cnt, tf, ts, ph == 1, v0, v1, c, lb, lane, tm);
      if (ph == 1) break;
      // The warp leaves its state vector in ITS OWN side's exchange slot.  When shared memory is short the slot aliases
      // that side's input ring, which is idle by now: the warp has consumed every frame its workers produced.
      float* dst = xch_of(my);
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        dst[j * kWarp + lane] = v0[j];
        if (CLASSIC) dst[kUpad + j * kWarp + lane] = v1[j];
      }
      if (lane == 0) xoff[my] = c;
      if (!middle()) {
        if (a.grad == nullptr) return;
        break;
      }
    }
#ifdef CTCB200_FUSED_TIMING
    dump_timing();
#endif
    if (side == 0) {
      // loss = -alpha[T, label_length] (classic_ctc_loss.py:152-165 / simplified_ctc_loss.py:73-83); frames beyond n_t
      // leave it unchanged.
#pragma unroll
      for (int j = 0; j < NS; ++j)
        if (lane * NS + j == L) {
          const double ld = dead ? (double)INFINITY : -((double)(CLASSIC ? lse2(v0[j], v1[j]) : v0[j]) + c);
          a.loss[b] = (float)ld;
        }
    }
  } else if (role < 0) {
    (void)middle();      // an idle warp: the CTA-wide barriers of the middle, nothing else
  } else if (HELPERS && role > W) {
    // ---- a row helper (only when f.helpers) ----
    int cnt, tf;
    phase_frames(0, cnt, tf);
    helper_phase<false>(a, f, side_view(smem, f, my, 0), b, role - 1 - W, cnt, tf, ts, dl, lane);
    if (middle()) {
      phase_frames(1, cnt, tf);
      helper_phase<true>(a, f, side_view(smem, f, my, 1), b, role - 1 - W, cnt, tf, ts, dl, lane);
    }
  } else {
    // ---- a row worker ----
#pragma unroll 1
    for (int ph = 0; ph < 2; ++ph) {
      int cnt, tf;
      phase_frames(ph, cnt, tf);
      const SideView sv = side_view(smem, f, my, ph);
      if (ph == 0) {
        worker_phase<NS, CLASSIC, false, TMA, BF16, HELPERS>(a, f, sv, side, b, role - 1, cnt, tf, ts, L, 0.0, dl, tok, ll, lane, tm);
      } else {
        worker_phase<NS, CLASSIC, true, TMA, BF16, HELPERS>(a, f, sv, side, b, role - 1, cnt, tf, ts, L, lossd_mid, dl, tok, ll, lane, tm);
        break;
      }
      if (!middle()) {
        if (a.grad == nullptr) return;
        break;
      }
    }
#ifdef CTCB200_FUSED_TIMING
    dump_timing();
#endif
    // frames beyond logit_length, or every frame of an infeasible sample: exact zeros
    const int widx = side * W + (role - 1);
    for (int r = (dead ? 0 : n_t) + widx; r < p.T; r += 2 * W) {
      if (BF16 && p.grad_bf16)       // a bf16 row is V/2 floats wide (V % 8 == 0)
        fused_zero_row(reinterpret_cast<float*>(reinterpret_cast<char*>(a.grad) + 2 * row_offset(p, b, r)), p.V >> 1, lane, true);
      else
        fused_zero_row(a.grad + row_offset(p, b, r), p.V, lane, TMA);
    }
  }
}

// ---- the two kernels ----------------------------------------------------------------------------------------------------
template <int NS, bool CLASSIC, bool TMA, bool BF16, bool HELPERS>
__global__ void __launch_bounds__(2 * (kMaxWorkers + 1) * kWarp, (NS <= 8) ? 2 : 1) kf_fused(const __grid_constant__ FusedArgs a) {
  fused_body<NS, CLASSIC, TMA, false, BF16, HELPERS>(a);
}
template <int NS, bool CLASSIC, bool TMA, bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__((NS <= 8) ? 112 : 224)
    kf_fused_split(const __grid_constant__ FusedArgs a) {
  fused_body<NS, CLASSIC, TMA, true, BF16, false>(a);
}

// ---- host side: one launcher per (variant, row-mover) pair, defined in kf_fused_*.cu --------------------------------
constexpr int kSmemPerSm = 227 * 1024;

template <bool CLASSIC, bool TMA, bool BF16>
cudaError_t launch_fused_variant(const FusedArgs& a, cudaStream_t st);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is issued once per (kernel, device) -- or again when a later problem
// needs more -- not on every launch.
constexpr int kMaxDevices = 64;
template <typename Kernel>
static cudaError_t ensure_smem(Kernel kernel, int* cache, int bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (__atomic_load_n(&cache[dev], __ATOMIC_ACQUIRE) >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) {      // racing callers all set a value that is large enough for themselves; keep the maximum
    int seen = __atomic_load_n(&cache[dev], __ATOMIC_RELAXED);
    while (seen < bytes && !__atomic_compare_exchange_n(&cache[dev], &seen, bytes, false, __ATOMIC_RELEASE, __ATOMIC_RELAXED)) {}
  }
  return e;
}

template <int NS, bool CLASSIC, bool TMA, bool BF16>
static cudaError_t launch_fused_ns(const FusedArgs& a, cudaStream_t st) {
  const FusedLayout f = fused_layout(a.p.V, a.p.Upad, a.p.S, a.W, a.SL, a.XA, a.R, a.split ? 1 : 2, a.half, (a.split || BF16 || !TMA) ? 0 : a.helpers);
  // NOTE: a concurrent caller on the same device may raise the attribute between this check and the launch; it is
  // never lowered, so the launch below always finds at least f.total bytes allowed.
  static int cache[3][kMaxDevices];
  constexpr bool kCanHelp = TMA && !BF16;      // the only instantiations that carry the row helpers' code
  if (a.split) {
    cudaError_t e = ensure_smem(kf_fused_split<NS, CLASSIC, TMA, BF16>, cache[1], f.total);
    if (e != cudaSuccess) return e;
    int warps = a.W + 1;
    if (a.rec_alone)       // 1 recursion warp + W workers + the idle warps 4, 8, ... in between
      for (warps = 1; warps - 1 - (warps - 1) / 4 < a.W; ++warps) {}
    kf_fused_split<NS, CLASSIC, TMA, BF16><<<2 * a.p.B, warps * kWarp, f.total, st>>>(a);
  } else if (kCanHelp && f.helpers) {
    cudaError_t e = ensure_smem(kf_fused<NS, CLASSIC, TMA, BF16, kCanHelp>, cache[2], f.total);
    if (e != cudaSuccess) return e;
    kf_fused<NS, CLASSIC, TMA, BF16, kCanHelp><<<a.p.B, 2 * (2 * a.W + 1) * kWarp, f.total, st>>>(a);
  } else {
    cudaError_t e = ensure_smem(kf_fused<NS, CLASSIC, TMA, BF16, false>, cache[0], f.total);
    if (e != cudaSuccess) return e;
    kf_fused<NS, CLASSIC, TMA, BF16, false><<<a.p.B, 2 * (a.W + 1) * kWarp, f.total, st>>>(a);
  }
  const cudaError_t err = cudaGetLastError();
  if (err == cudaErrorLaunchOutOfResources) {      // say which resource: the plan and the compiled kernel disagree
    cudaFuncAttributes fa{};
    if (a.split) (void)cudaFuncGetAttributes(&fa, kf_fused_split<NS, CLASSIC, TMA, BF16>);
    else if (kCanHelp && f.helpers) (void)cudaFuncGetAttributes(&fa, kf_fused<NS, CLASSIC, TMA, BF16, kCanHelp>);
    else (void)cudaFuncGetAttributes(&fa, kf_fused<NS, CLASSIC, TMA, BF16, false>);
    fprintf(stderr, "libctc_b200: kf_fused%s<NS=%d,classic=%d,tma=%d> bf16=%d W=%d SL=%d XA=%d R=%d: %d threads, %d B dynamic smem; kernel: %d regs, "
            "max %d threads/block, %zu B static smem, %d B max dynamic smem, %zu B local\n", a.split ? "_split" : "", NS, (int)CLASSIC,
            (int)TMA, (int)BF16, a.W, a.SL, a.XA, a.R, (a.split ? 1 : 2) * (a.W + 1) * kWarp, f.total, fa.numRegs, fa.maxThreadsPerBlock,
            fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes, fa.localSizeBytes);
    (void)cudaGetLastError();
  }
  return err;
}

#define CTCB200_FUSED_CASE(n) \
  case n:                     \
    return launch_fused_ns<n, CLASSIC_, TMA_, BF16_>(a, st);
#define CTCB200_DEFINE_FUSED_VARIANT(CLASSIC__, TMA__, BF16__)                                                       \
  template <>                                                                                                       \
  cudaError_t launch_fused_variant<CLASSIC__, TMA__, BF16__>(const FusedArgs& a, cudaStream_t st) {                 \
    constexpr bool CLASSIC_ = CLASSIC__, TMA_ = TMA__, BF16_ = BF16__;                                               \
    switch (a.p.NS) {                                                                                               \
      CTCB200_FUSED_CASE(1) CTCB200_FUSED_CASE(2) CTCB200_FUSED_CASE(3) CTCB200_FUSED_CASE(4)                       \
      CTCB200_FUSED_CASE(5) CTCB200_FUSED_CASE(6) CTCB200_FUSED_CASE(7) CTCB200_FUSED_CASE(8)                       \
      CTCB200_FUSED_CASE(9) CTCB200_FUSED_CASE(10) CTCB200_FUSED_CASE(11) CTCB200_FUSED_CASE(12)                    \
      CTCB200_FUSED_CASE(13) CTCB200_FUSED_CASE(14) CTCB200_FUSED_CASE(15) CTCB200_FUSED_CASE(16)                   \
      default: return cudaErrorInvalidValue;                                                                        \
    }                                                                                                               \
  }

}  // namespace ctcb200
