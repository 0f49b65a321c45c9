// C ABI of libctc_b200.so (include/ctc_b200.h): descriptor validation, workspace carving, kernel sequencing.
#include <new>

#include "common.cuh"

namespace ctcb200 {

static size_t align256(size_t n) { return (n + 255) & ~size_t(255); }

static int make_problem(const ctcb200_desc* d, Problem* p) {
  if (d == nullptr) return CTCB200_ERR_NULL_POINTER;
  if (d->B < 0 || d->T < 0 || d->V < 1 || d->Lw < 0 || d->U < 0) return CTCB200_ERR_BAD_DESCRIPTOR;
  if (d->blank < 0 || d->blank >= d->V) return CTCB200_ERR_BAD_DESCRIPTOR;
  if (d->variant != CTCB200_CLASSIC && d->variant != CTCB200_SIMPLIFIED) return CTCB200_ERR_BAD_DESCRIPTOR;
  if (d->flags & ~(CTCB200_INPUT_LOGPROBAS | CTCB200_FORCE_STAGED | CTCB200_FORCE_FUSED | CTCB200_TIME_MAJOR |
                   CTCB200_LOGITS_BF16 | CTCB200_GRAD_BF16 | CTCB200_STAGE_MASK))
    return CTCB200_ERR_BAD_DESCRIPTOR;
  if ((d->flags & CTCB200_GRAD_BF16) && !(d->flags & CTCB200_LOGITS_BF16)) return CTCB200_ERR_BAD_DESCRIPTOR;
  if ((d->flags & CTCB200_LOGITS_BF16) && (d->flags & CTCB200_FORCE_STAGED)) return CTCB200_ERR_BAD_DESCRIPTOR;
  p->B = d->B; p->T = d->T; p->V = d->V; p->Lw = d->Lw; p->blank = d->blank; p->variant = d->variant;
  p->U = d->U > 0 ? d->U : d->Lw + 1;
  p->NS = (p->U + kWarp - 1) / kWarp;
  if (p->NS > kMaxNS) p->NS = (p->NS + 3) & ~3;      // staged kernels only: instantiated in steps of four states per lane
  if (p->NS > kMaxNSStaged || d->V > kMaxV) return CTCB200_ERR_UNSUPPORTED_SIZE;
  if ((long long)d->B * (long long)(d->T + 1) > (1LL << 40)) return CTCB200_ERR_UNSUPPORTED_SIZE;
  p->Upad = p->NS * kWarp;
  p->S = d->variant == CTCB200_CLASSIC ? 2 : 1;
  p->input_logprobas = (d->flags & CTCB200_INPUT_LOGPROBAS) != 0;
  p->logits_bf16 = (d->flags & CTCB200_LOGITS_BF16) != 0;
  p->grad_bf16 = (d->flags & CTCB200_GRAD_BF16) != 0;
  const bool tm = (d->flags & CTCB200_TIME_MAJOR) != 0;
  p->stride_b = tm ? (size_t)d->V : (size_t)d->T * d->V;
  p->stride_t = tm ? (size_t)d->B * d->V : (size_t)d->V;
  p->logits = nullptr; p->labels = nullptr; p->label_length = nullptr; p->logit_length = nullptr;
  return CTCB200_OK;
}

struct Carve {
  size_t off = 0;
  size_t take(size_t bytes) { size_t o = off; off += align256(bytes); return o; }
};

// `fused_only`: the call is served by the fused kernel alone, which needs the row log-sum-exps, ONE state tensor and one
// offset vector; the gathered rows, the second state tensor and its offsets (staged kernels) are left out.
static size_t carve(const Problem& p, int what, bool fused_only, char* base, Scratch* s, float** grad_tmp) {
  Carve c;
  if (what == CTCB200_WS_DECODE) {       // arg-max token and its logit per row; nothing else
    const size_t o = c.take((size_t)p.B * p.T * 4);
    c.take((size_t)p.B * p.T * 4);
    if (base != nullptr && grad_tmp != nullptr) *grad_tmp = reinterpret_cast<float*>(base + o);
    return c.off;
  }
#ifdef CTCB200_FUSED_TIMING
  const bool staged = true;            // the instrumented kernel writes its counters into the (otherwise unused) beta scratch
  (void)fused_only;
#else
  const bool staged = !fused_only;
#endif
  const size_t rows = (size_t)p.B * p.T, srows = (size_t)p.B * (p.T + 1) * p.S * p.Upad;
  const size_t o_lse = c.take(rows * 4), o_h = c.take(staged ? rows * 4 : 0), o_d = c.take(staged ? rows * p.Upad * 4 : 0);
  const size_t o_a = c.take(srows * 4), o_b = c.take(staged ? srows * 4 : 0), o_loss = c.take((size_t)p.B * 4);
  const size_t o_ca = c.take((size_t)p.B * (p.T + 1) * 8), o_cb = c.take(staged ? (size_t)p.B * (p.T + 1) * 8 : 0);
  const size_t o_ld = c.take((size_t)p.B * 8);
  size_t o_g = 0;
  if (what == CTCB200_WS_HESSIAN) o_g = c.take(rows * p.V * 4);
  if (what == CTCB200_WS_HVP_LOGITS) o_g = c.take(3 * align256(rows * p.V * 4) + align256(rows * 4));   // g, w, y, p.v
  if (base != nullptr) {
    s->rowlse = reinterpret_cast<float*>(base + o_lse);
    s->h = reinterpret_cast<float*>(base + o_h);
    s->dT = reinterpret_cast<float*>(base + o_d);
    s->alphaT = reinterpret_cast<float*>(base + o_a);
    s->betaT = reinterpret_cast<float*>(base + o_b);
    s->loss = reinterpret_cast<float*>(base + o_loss);
    s->ca = reinterpret_cast<double*>(base + o_ca);
    s->cb = reinterpret_cast<double*>(base + o_cb);
    s->lossd = reinterpret_cast<double*>(base + o_ld);
    if (grad_tmp) *grad_tmp = (what == CTCB200_WS_HESSIAN || what == CTCB200_WS_HVP_LOGITS) ? reinterpret_cast<float*>(base + o_g) : nullptr;
  }
  return c.off;
}

// The fused kernel takes the loss+gradient call whenever its shared-memory plan fits (rows move by TMA when V % 4 == 0
// and the bases are 16-byte aligned, by 4-byte cp.async otherwise).
static int fused_workers(const ctcb200_desc* desc, const Problem& p) {
  if (desc->flags & CTCB200_FORCE_STAGED) return 0;
  if (p.logits_bf16) return fused_pick_workers(p);      // bf16 rows exist in the fused kernel only
  // Narrow vocabularies (character models, V < 64) in SMALL batches are latency-bound on the T-step chain rather than on
  // row traffic; there the staged recursion kernel, which streams the compact gathered rows, is as fast or faster
  // (classic T=500 V=29 L=100, B=32: 165 us staged vs 171 us fused with the split plan).  From about 48 utterances on the
  // fused kernel wins because the staged kernels grow with B while the fused one still fits one wave (same shape, B=64:
  // 171 vs 201 us, B=128: 189 vs 277 us; B=256 T=255 V=32: 125 vs 228 us).
  if (p.V < 64 && p.B < 48 && !(desc->flags & CTCB200_FORCE_FUSED)) return 0;
  return fused_pick_workers(p);
}

// WS_LOSS_GRAD_LOGITS shrinks to the fused kernel's needs whenever that kernel will take the call
static bool fused_only_ws(const ctcb200_desc* desc, const Problem& p, int what) {
  return what == CTCB200_WS_LOSS_GRAD_LOGITS && p.T > 0 && fused_workers(desc, p) > 0;
}

static int check_common(const ctcb200_desc* desc, Problem* p, int what, const float* logits, const int32_t* labels,
                        const int32_t* label_length, const int32_t* logit_length, void* ws, size_t ws_bytes,
                        Scratch* s, float** grad_tmp) {
  int rc = make_problem(desc, p);
  if (rc != CTCB200_OK) return rc;
  if (p->B > 0) {
    if (label_length == nullptr || logit_length == nullptr) return CTCB200_ERR_NULL_POINTER;
    if (p->T > 0 && logits == nullptr) return CTCB200_ERR_NULL_POINTER;
    if (p->Lw > 0 && labels == nullptr) return CTCB200_ERR_NULL_POINTER;
  }
  const bool fo = fused_only_ws(desc, *p, what);
  const size_t need = carve(*p, what, fo, nullptr, nullptr, nullptr);
  if (need > 0 && p->B > 0) {
    if (ws == nullptr) return CTCB200_ERR_NULL_POINTER;
    if (reinterpret_cast<uintptr_t>(ws) & 255) return CTCB200_ERR_MISALIGNED;
    if (ws_bytes < need) return CTCB200_ERR_WORKSPACE_TOO_SMALL;
  }
  carve(*p, what, fo, static_cast<char*>(ws), s, grad_tmp);
  p->logits = logits; p->labels = labels; p->label_length = label_length; p->logit_length = logit_length;
  return CTCB200_OK;
}

// the CUDA status behind the last CTCB200_ERR_CUDA returned on this thread (ctcb200_last_cuda_error)
static thread_local cudaError_t t_last_cuda_error = cudaSuccess;

#define CTCB200_CUDA(call)                       \
  do {                                           \
    const cudaError_t e__ = (call);              \
    if (e__ != cudaSuccess) {                    \
      t_last_cuda_error = e__;                   \
      (void)cudaGetLastError();                  \
      return CTCB200_ERR_CUDA;                   \
    }                                            \
  } while (0)

}  // namespace ctcb200

using namespace ctcb200;

extern "C" {

int ctcb200_version(void) { return CTCB200_VERSION; }

const char* ctcb200_strerror(int code) {
  switch (code) {
    case CTCB200_OK: return "ok";
    case CTCB200_ERR_NULL_POINTER: return "null pointer";
    case CTCB200_ERR_BAD_DESCRIPTOR: return "bad descriptor";
    case CTCB200_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case CTCB200_ERR_UNSUPPORTED_SIZE: return "unsupported size (U > 1024 states, V > 32768 tokens, or bf16 rows the fused kernel cannot take)";
    case CTCB200_ERR_CUDA: return "CUDA launch failure";
    case CTCB200_ERR_MISALIGNED: return "workspace must be 256-byte aligned";
    default: return "unknown error";
  }
}

const char* ctcb200_last_cuda_error(void) { return cudaGetErrorString(t_last_cuda_error); }

const char* ctcb200_stage_names(const ctcb200_desc* desc) {
  Problem p;
  if (make_problem(desc, &p) == CTCB200_OK && fused_workers(desc, p) > 0) return "kf_fused";
  return "k1_softmax_gather,k2_recursion,k3_grad";
}

int ctcb200_launches_per_call(const ctcb200_desc* desc) {
  Problem p;
  if (make_problem(desc, &p) != CTCB200_OK || p.B == 0) return 0;
  if (fused_workers(desc, p) > 0) return 1;
  return (p.T > 0 ? 1 : 0) + 1 + (p.T > 0 ? 1 : 0);
}

void ctcb200_debug_fused_plan(int workers, int row_buffers, int extra_phase_a_buffer, int ring_depth, int split) {
  fused_set_plan_override(workers, row_buffers, extra_phase_a_buffer, ring_depth, split);
}

size_t ctcb200_workspace_bytes(const ctcb200_desc* desc, int what) {
  Problem p;
  if (make_problem(desc, &p) != CTCB200_OK) return 0;
  if (what < CTCB200_WS_LOSS_GRAD || what > CTCB200_WS_DECODE) return 0;
  return carve(p, what, fused_only_ws(desc, p, what), nullptr, nullptr, nullptr);
}

int ctcb200_loss_grad(const ctcb200_desc* desc, const float* logits, const int32_t* labels,
                      const int32_t* label_length, const int32_t* logit_length, const float* d_loss, float* loss,
                      float* grad_logits, float* grad_logprobas, void* workspace, size_t workspace_bytes,
                      void* stream) {
  Problem p; Scratch s;
  // loss + d/dlogits alone -- or the loss alone -- is the fused kernel's call (when its plan fits) and then needs the
  // smaller workspace only
  const bool logits_only = grad_logprobas == nullptr;
  int rc = check_common(desc, &p, logits_only ? CTCB200_WS_LOSS_GRAD_LOGITS : CTCB200_WS_LOSS_GRAD, logits, labels,
                        label_length, logit_length, workspace, workspace_bytes, &s, nullptr);
  if (rc != CTCB200_OK) return rc;
  if (p.B == 0) return CTCB200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* loss_out = loss ? loss : s.loss;
  const int W = fused_workers(desc, p);
  if (p.logits_bf16) {
    // bf16 rows are a format of the fused kernel alone: TMA-movable rows (V % 8 == 0, 16-byte aligned bases), no
    // log-probability gradient, and a shape the fused plan takes
    const uintptr_t bases = reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(grad_logits);
    if (grad_logprobas != nullptr || (p.V & 7) != 0 || (bases & 15) != 0 || W == 0 || p.T == 0) return CTCB200_ERR_UNSUPPORTED_SIZE;
  }
  if (W > 0 && logits_only && p.T > 0) {
    CTCB200_CUDA(launch_fused(p, s, d_loss, loss_out, grad_logits, W, st));
    return CTCB200_OK;
  }
  const unsigned stages = (desc->flags & CTCB200_STAGE_MASK) >> CTCB200_STAGE_SHIFT;
  if (stages == 0 || (stages & 1u)) CTCB200_CUDA(launch_softmax_gather(p, s, st));
  if (stages == 0 || (stages & 2u)) CTCB200_CUDA(launch_recursion(p, s, loss_out, false, st));
  if (stages == 0 || (stages & 4u)) CTCB200_CUDA(launch_grad(p, s, d_loss, grad_logits, grad_logprobas, st));
  return CTCB200_OK;
}

int ctcb200_log_gradient(const ctcb200_desc* desc, const float* logits, const int32_t* labels,
                         const int32_t* label_length, const int32_t* logit_length, float* loss, float* log_gradient,
                         void* workspace, size_t workspace_bytes, void* stream) {
  Problem p; Scratch s;
  int rc = check_common(desc, &p, CTCB200_WS_LOSS_GRAD, logits, labels, label_length, logit_length, workspace,
                        workspace_bytes, &s, nullptr);
  if (rc != CTCB200_OK) return rc;
  if (desc->flags & (CTCB200_TIME_MAJOR | CTCB200_LOGITS_BF16)) return CTCB200_ERR_BAD_DESCRIPTOR;   // loss_grad only
  if (p.B == 0) return CTCB200_OK;
  if (log_gradient == nullptr && p.T > 0) return CTCB200_ERR_NULL_POINTER;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CTCB200_CUDA(launch_softmax_gather(p, s, st));
  CTCB200_CUDA(launch_recursion(p, s, loss ? loss : s.loss, false, st));
  CTCB200_CUDA(launch_log_grad(p, s, log_gradient, st));
  return CTCB200_OK;
}

int ctcb200_states(const ctcb200_desc* desc, const float* logits, const int32_t* labels,
                   const int32_t* label_length, const int32_t* logit_length, float* alpha, float* beta,
                   float* loss, void* workspace, size_t workspace_bytes, void* stream) {
  Problem p; Scratch s;
  int rc = check_common(desc, &p, CTCB200_WS_STATES, logits, labels, label_length, logit_length, workspace,
                        workspace_bytes, &s, nullptr);
  if (rc != CTCB200_OK) return rc;
  if (desc->flags & (CTCB200_TIME_MAJOR | CTCB200_LOGITS_BF16)) return CTCB200_ERR_BAD_DESCRIPTOR;   // loss_grad only
  if (desc->U <= 0) return CTCB200_ERR_BAD_DESCRIPTOR;   // the output shape depends on the true U
  if (p.B == 0) return CTCB200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CTCB200_CUDA(launch_softmax_gather(p, s, st));
  CTCB200_CUDA(launch_recursion(p, s, loss ? loss : s.loss, true, st));
  CTCB200_CUDA(launch_export_states(p, s, alpha, beta, st));
  return CTCB200_OK;
}

int ctcb200_gamma(const ctcb200_desc* desc, const float* logits, const int32_t* labels, const int32_t* label_length,
                  const int32_t* logit_length, float* gamma, void* workspace, size_t workspace_bytes, void* stream) {
  Problem p; Scratch s;
  int rc = check_common(desc, &p, CTCB200_WS_STATES, logits, labels, label_length, logit_length, workspace,
                        workspace_bytes, &s, nullptr);
  if (rc != CTCB200_OK) return rc;
  if (desc->flags & (CTCB200_TIME_MAJOR | CTCB200_LOGITS_BF16)) return CTCB200_ERR_BAD_DESCRIPTOR;   // loss_grad only
  if (desc->U <= 0) return CTCB200_ERR_BAD_DESCRIPTOR;   // the output shape depends on the true U
  if (p.NS > 4) return CTCB200_ERR_UNSUPPORTED_SIZE;
  if (p.B == 0) return CTCB200_OK;
  if (gamma == nullptr) return CTCB200_ERR_NULL_POINTER;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CTCB200_CUDA(launch_softmax_gather(p, s, st));
  CTCB200_CUDA(launch_gamma(p, s, gamma, st));
  return CTCB200_OK;
}

int ctcb200_hessian(const ctcb200_desc* desc, const float* logits, const int32_t* labels,
                    const int32_t* label_length, const int32_t* logit_length, float* hessian, float* loss,
                    float* grad_logprobas, void* workspace, size_t workspace_bytes, void* stream) {
  Problem p; Scratch s; float* gtmp = nullptr;
  int rc = check_common(desc, &p, CTCB200_WS_HESSIAN, logits, labels, label_length, logit_length, workspace,
                        workspace_bytes, &s, &gtmp);
  if (rc != CTCB200_OK) return rc;
  if (desc->flags & (CTCB200_TIME_MAJOR | CTCB200_LOGITS_BF16)) return CTCB200_ERR_BAD_DESCRIPTOR;   // loss_grad only
  if (p.B == 0 || p.T == 0) return CTCB200_OK;
  if (hessian == nullptr) return CTCB200_ERR_NULL_POINTER;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* loss_out = loss ? loss : s.loss;
  CTCB200_CUDA(launch_softmax_gather(p, s, st));
  CTCB200_CUDA(launch_recursion(p, s, loss_out, false, st));
  float* g = grad_logprobas ? grad_logprobas : gtmp;
  CTCB200_CUDA(launch_grad(p, s, nullptr, nullptr, g, st));
  CTCB200_CUDA(launch_hessian(p, s, g, hessian, nullptr, nullptr, st));
  return CTCB200_OK;
}

int ctcb200_hvp(const ctcb200_desc* desc, const float* logits, const int32_t* labels,
                const int32_t* label_length, const int32_t* logit_length, const float* d_gradient, float* out,
                void* workspace, size_t workspace_bytes, void* stream) {
  Problem p; Scratch s; float* gtmp = nullptr;
  int rc = check_common(desc, &p, CTCB200_WS_HESSIAN, logits, labels, label_length, logit_length, workspace,
                        workspace_bytes, &s, &gtmp);
  if (rc != CTCB200_OK) return rc;
  if (desc->flags & (CTCB200_TIME_MAJOR | CTCB200_LOGITS_BF16)) return CTCB200_ERR_BAD_DESCRIPTOR;   // loss_grad only
  if (p.B == 0 || p.T == 0) return CTCB200_OK;
  if (d_gradient == nullptr || out == nullptr) return CTCB200_ERR_NULL_POINTER;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CTCB200_CUDA(launch_softmax_gather(p, s, st));
  CTCB200_CUDA(launch_recursion(p, s, s.loss, false, st));
  CTCB200_CUDA(launch_grad(p, s, nullptr, nullptr, gtmp, st));
  CTCB200_CUDA(launch_hessian(p, s, gtmp, nullptr, d_gradient, out, st));
  return CTCB200_OK;
}

int ctcb200_hvp_logits(const ctcb200_desc* desc, const float* logits, const int32_t* labels,
                       const int32_t* label_length, const int32_t* logit_length, const float* d_loss, const float* v,
                       float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (desc != nullptr && (desc->flags & CTCB200_INPUT_LOGPROBAS)) return CTCB200_ERR_BAD_DESCRIPTOR;   // wants logits
  Problem p; Scratch s; float* tmp = nullptr;
  int rc = check_common(desc, &p, CTCB200_WS_HVP_LOGITS, logits, labels, label_length, logit_length, workspace,
                        workspace_bytes, &s, &tmp);
  if (rc != CTCB200_OK) return rc;
  if (desc->flags & (CTCB200_TIME_MAJOR | CTCB200_LOGITS_BF16)) return CTCB200_ERR_BAD_DESCRIPTOR;   // loss_grad only
  if (p.B == 0 || p.T == 0) return CTCB200_OK;
  if (v == nullptr || out == nullptr) return CTCB200_ERR_NULL_POINTER;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = align256((size_t)p.B * p.T * p.V * 4) / 4;
  float *g = tmp, *w = tmp + n, *y = tmp + 2 * n, *pv = tmp + 3 * n;
  CTCB200_CUDA(launch_softmax_gather(p, s, st));
  CTCB200_CUDA(launch_recursion(p, s, s.loss, false, st));
  CTCB200_CUDA(launch_grad(p, s, nullptr, nullptr, g, st));
  CTCB200_CUDA(launch_hvp_pre(p, s, v, w, pv, st));
  CTCB200_CUDA(launch_hessian(p, s, g, nullptr, w, y, st));
  CTCB200_CUDA(launch_hvp_post(p, s, v, y, g, pv, d_loss, out, st));
  return CTCB200_OK;
}

int ctcb200_greedy_decode(const ctcb200_desc* desc, const float* logits, const int32_t* logit_length, int merge_repeated,
                          int32_t* decoded, int32_t* decoded_length, float* neg_sum_logits, void* workspace,
                          size_t workspace_bytes, void* stream) {
  Problem p; Scratch s; float* tmp = nullptr;
  // labels / label_length play no part in decoding: the logit lengths stand in for them in the common pointer checks
  int rc = check_common(desc, &p, CTCB200_WS_DECODE, logits, logit_length, logit_length, logit_length, workspace,
                        workspace_bytes, &s, &tmp);
  if (rc != CTCB200_OK) return rc;
  if (p.B == 0) return CTCB200_OK;
  if (decoded == nullptr || decoded_length == nullptr) return CTCB200_ERR_NULL_POINTER;
  const size_t n = align256((size_t)p.B * p.T * 4) / 4;
  CTCB200_CUDA(launch_greedy_decode(p, reinterpret_cast<int*>(tmp), tmp + n, merge_repeated, decoded, decoded_length,
                                    neg_sum_logits, static_cast<cudaStream_t>(stream)));
  return CTCB200_OK;
}

}  // extern "C"
