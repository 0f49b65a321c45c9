// Shared device helpers and the internal launch interface of libctc_b200.so (sm_100a).
//
// Notation follows the reference (tf_seq2seq_losses): B batch, T frames, V tokens, U = max label length + 1
// "tokens emitted so far" states, h[t] = log-prob of blank at frame t, d[t,l] = log-prob of label token l at
// frame t (-inf for l >= label_length).  See DESIGN.md for the data layout.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <cmath>
#include <stdint.h>

#include "../../include/ctc_b200.h"

namespace ctcb200 {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxNS = 16;          // states per lane the fused kernel carries -> U <= 512
constexpr int kMaxNSStaged = 32;    // the staged kernels go on to U <= 1024 (NS rounded up to 20, 24, 28 or 32 above 16)
constexpr int kMaxV = 32768;        // token -> slot map lives in shared memory as int16
#define kNegInf (-INFINITY)

// ---- numerics (tf_seq2seq_losses/tools.py:57-71) -------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// log(e^a + e^b) = max + log1p(exp(-|a-b|)); (-inf,-inf) -> -inf and a == b -> a + log 2, exactly as the reference's
// three-way tf.where.  2 MUFU (EX2, LG2) + 6 ALU ops, no branches: for a == b == -inf the difference is NaN, which
// fminf(NaN, 0) turns into 0, and -inf + log 2 is -inf again.  Abs error ~2^-21, inside the fp32 tolerance of the path.
__device__ __forceinline__ float lse2(float a, float b) {
  const float t = fminf(-fabsf(a - b) * 1.4426950408889634f, 0.0f);
  return fmaf(lg2_approx(1.0f + ex2_approx(t)), 0.6931471805599453f, fmaxf(a, b));
}

// warp-wide float max in one REDUX: floats are mapped monotonically to signed integers (flip the magnitude bits of
// negatives), reduced with redux.sync.max.s32 and mapped back; -inf/+inf keep their order, NaN inputs are don't-care.
__device__ __forceinline__ float warp_max(float v) {
#ifdef CTCB200_SHUFFLE_MAX
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
#endif
  int i = __float_as_int(v);
  i ^= (i >> 31) & 0x7fffffff;
  i = __reduce_max_sync(kFull, i);
  i ^= (i >> 31) & 0x7fffffff;
  return __int_as_float(i);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// ---- private state layout -------------------------------------------------------------------------------------
// A recursion warp keeps NS = ceil(U/32) consecutive states per lane: state l lives in lane l / NS, register
// l % NS.  Scratch rows are stored "transposed" so that a warp-wide load of register j is one coalesced 128 B
// line: position(l) = (l % NS) * 32 + l / NS, row pitch Upad = 32 * NS.
__host__ __device__ __forceinline__ int state_pos(int l, int ns) { return (l % ns) * kWarp + l / ns; }

// streaming (read-once / write-once) global accesses that do not pollute L1
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream4(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// ---- problem description shared by all launchers -------------------------------------------------------------
struct Problem {
  int B, T, V, Lw, blank, variant;
  int U;      // state count bound (desc.U or Lw+1)
  int NS;     // states per lane
  int Upad;   // 32 * NS
  int S;      // 1 simplified, 2 classic (closed/open)
  bool input_logprobas;
  bool logits_bf16, grad_bf16;   // CTCB200_LOGITS_BF16 / CTCB200_GRAD_BF16: the fused kernel's bf16 row formats
  size_t stride_b, stride_t;   // floats between consecutive utterances / frames of logits and gradients (see row_offset)
  const float* logits;
  const int32_t* labels;
  const int32_t* label_length;
  const int32_t* logit_length;
};

struct Scratch {
  float* rowlse;   // [B*T]      log-sum-exp of each logits row (0 when the input is log-probabilities)
  float* h;        // [B*T]      blank log-prob
  float* dT;       // [B*T*Upad] label-token log-probs, private layout
  float* alphaT;   // [B*(T+1)*S*Upad]
  float* betaT;    // [B*(T+1)*S*Upad]
  float* loss;     // [B] (when the caller passes no loss buffer)
  // Offset renormalisation: the stored rows are alpha[t,.] - ca[t] and beta[t,.] - cb[t]; the offsets are running
  // sums of lagged warp maxima kept in double, so magnitudes stay O(10) instead of O(T log V) and the fp32
  // rounding error of the path no longer grows with T.  lossd = -(alpha[T,L]) in double.
  double* ca;      // [B*(T+1)]
  double* cb;      // [B*(T+1)]
  double* lossd;   // [B]
};

// offset of row (b, t) in logits / grad_logits / grad_logprobas: [B,T,V], or [T,B,V] with CTCB200_TIME_MAJOR
__host__ __device__ __forceinline__ size_t row_offset(const Problem& p, int b, int t) {
  return (size_t)b * p.stride_b + (size_t)t * p.stride_t;
}

// per-utterance clamped lengths
__device__ __forceinline__ int utt_label_len(const Problem& p, int b) {
  int L = p.label_length[b];
  return max(0, min(L, p.U - 1));
}
__device__ __forceinline__ int utt_frames(const Problem& p, int b) {
  int n = p.logit_length[b];
  return max(0, min(n, p.T));
}
// the reference's cleaned label (base_loss.py:395-418): labels beyond Lw or label_length are blank
__device__ __forceinline__ int utt_token(const Problem& p, int b, int l, int L) {
  return (l >= 0 && l < L && l < p.Lw) ? p.labels[(size_t)b * p.Lw + l] : p.blank;
}

cudaError_t launch_softmax_gather(const Problem& p, const Scratch& s, cudaStream_t st);
cudaError_t launch_recursion(const Problem& p, const Scratch& s, float* loss, bool full_states, cudaStream_t st);
cudaError_t launch_grad(const Problem& p, const Scratch& s, const float* d_loss, float* grad_logits,
                        float* grad_logprobas, cudaStream_t st);
cudaError_t launch_log_grad(const Problem& p, const Scratch& s, float* log_grad, cudaStream_t st);
int fused_pick_workers(const Problem& p);
void fused_set_plan_override(int W, int SL, int XA, int R, int split);
cudaError_t launch_fused(const Problem& p, const Scratch& s, const float* d_loss, float* loss, float* grad, int W,
                         cudaStream_t st);
cudaError_t launch_export_states(const Problem& p, const Scratch& s, float* alpha, float* beta, cudaStream_t st);
cudaError_t launch_gamma(const Problem& p, const Scratch& s, float* gamma, cudaStream_t st);
cudaError_t launch_greedy_decode(const Problem& p, int* best, float* bestv, int merge_repeated, int* decoded,
                                 int* decoded_length, float* neg_sum_logits, cudaStream_t st);
cudaError_t launch_hvp_pre(const Problem& p, const Scratch& s, const float* v, float* w, float* pv, cudaStream_t st);
cudaError_t launch_hvp_post(const Problem& p, const Scratch& s, const float* v, const float* y, const float* g,
                            const float* pv, const float* d_loss, float* out, cudaStream_t st);
cudaError_t launch_hessian(const Problem& p, const Scratch& s, const float* g, float* hessian,
                           const float* d_gradient, float* hvp_out, cudaStream_t st);

}  // namespace ctcb200
