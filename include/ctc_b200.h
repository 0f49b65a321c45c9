/*
 * ctc_b200.h -- C ABI of libctc_b200.so: the CTC loss / gradient / Hessian hot path of
 * alexeytochin/tf_seq2seq_losses rebuilt as hand-written CUDA kernels for NVIDIA B200 (sm_100a).
 *
 * The reference has no FFI of its own (it is Python over TensorFlow ops); each entry point below replaces the
 * reference *Python* interface cited beside it (paths relative to the reference repository root) and is what a
 * TensorFlow custom op (tf_seq2seq_losses_b200/tf_adapter/ctc_b200_tf_op.cc), a ctypes binding
 * (tf_seq2seq_losses_b200/_lib.py) or any other host would bind.  See INTEGRATION.md.
 *
 * Contract (all device entry points):
 *   - every pointer is a DEVICE pointer owned by the caller, including the workspace; the library never
 *     allocates, frees or synchronises; all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - inputs are read-only, outputs are fully overwritten (rows t >= logit_length and infeasible samples are
 *     written as exact zeros, their loss as +inf);
 *   - return value: CTCB200_OK (0) or a negative error code, never an exception / abort;
 *   - re-entrant: no mutable global state except one-time per-device kernel attribute caching; safe to call
 *     concurrently on different streams / devices.
 * Shapes: logits float32 [B,T,V] batch-major contiguous; labels int32 [B,Lw]; label_length, logit_length int32 [B].
 * All arithmetic is fp32 (the reference asserts float32, tf_seq2seq_losses/base_loss.py:131).
 */
#ifndef CTC_B200_H_
#define CTC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTCB200_VERSION 100 /* 0.1.0 */

/* variant: which data class of the reference is being replaced */
#define CTCB200_CLASSIC 0    /* ClassicCtcLossData,    tf_seq2seq_losses/classic_ctc_loss.py:73-669 */
#define CTCB200_SIMPLIFIED 1 /* SimplifiedCtcLossData, tf_seq2seq_losses/simplified_ctc_loss.py:70-534 */

/* flags */
#define CTCB200_INPUT_LOGPROBAS 1u /* `logits` already holds log-probabilities (ctc_loss_from_logproba,
                                      base_loss.py:71-99, and the data-class constructors, base_loss.py:105-114):
                                      the log-softmax of tools.py:27-40 is skipped */

#define CTCB200_FORCE_STAGED 2u    /* ctcb200_loss_grad: use the three staged kernels (K1 softmax+gather, K2 recursion,
                                      K3 gradient) even where the fused single-launch kernel applies */

#define CTCB200_FORCE_FUSED 4u     /* ctcb200_loss_grad: use the fused kernel even for narrow vocabularies (V < 64), where the
                                      staged kernels are the default */

#define CTCB200_TIME_MAJOR 8u      /* ctcb200_loss_grad only: logits and both gradient outputs are time-major, [T,B,V] (the layout
                                      tf.nn.ctc_loss calls logits_time_major; the reference itself is batch-major only,
                                      classic_ctc_loss.py:33-39).  Every other array keeps its layout. */

#define CTCB200_LOGITS_BF16 16u    /* ctcb200_loss_grad / ctcb200_host_* only (an extension: the reference asserts float32,
                                      base_loss.py:131): `logits` points at bfloat16 [B,T,V] (or [T,B,V]).  All arithmetic stays
                                      fp32; the HBM and PCIe bytes of the input halve.  Served by the fused kernel alone: needs
                                      V % 8 == 0, 16-byte aligned bases, grad_logprobas == NULL, and a shape the fused plan takes
                                      (else CTCB200_ERR_UNSUPPORTED_SIZE). */
#define CTCB200_GRAD_BF16 32u      /* with CTCB200_LOGITS_BF16: grad_logits is written as bfloat16 too (round to nearest even) */

/* Profiling aid: flags bits 8..15 select which stages of ctcb200_loss_grad are enqueued (bit 8+i = i-th name of
 * ctcb200_stage_names()); 0 = all.  A partial call must follow a full call on the same workspace and inputs. */
#define CTCB200_STAGE_SHIFT 8
#define CTCB200_STAGE_MASK 0xFF00u

/* error codes */
#define CTCB200_OK 0
#define CTCB200_ERR_NULL_POINTER (-1)
#define CTCB200_ERR_BAD_DESCRIPTOR (-2)
#define CTCB200_ERR_WORKSPACE_TOO_SMALL (-3)
#define CTCB200_ERR_UNSUPPORTED_SIZE (-4) /* U > 1024 states (512 for bf16 rows) or V > 32768 tokens */
#define CTCB200_ERR_CUDA (-5)             /* launch failure; cudaGetLastError() was consumed */
#define CTCB200_ERR_MISALIGNED (-6)       /* workspace not 256-byte aligned */

/* what a workspace is sized for */
#define CTCB200_WS_LOSS_GRAD 0
#define CTCB200_WS_STATES 1
#define CTCB200_WS_HESSIAN 2
/* ctcb200_loss_grad called with grad_logits != NULL and grad_logprobas == NULL (the training call): when the fused kernel
 * serves the shape this is about a third of CTCB200_WS_LOSS_GRAD (no gathered rows, one state tensor instead of two);
 * otherwise the two are equal.  A workspace sized with CTCB200_WS_LOSS_GRAD is always accepted as well.
 * NOTE: this size is NOT monotonic in B -- the kernel choice depends on the batch size (narrow vocabularies, V < 64, take
 * the fused kernel from 48 utterances on and the larger staged scratch below that), so a caller that sizes one workspace
 * for its largest batch and then passes smaller batches must size with CTCB200_WS_LOSS_GRAD, or take the maximum over the
 * batch sizes it will use (ctcb200_host_create does the latter for its tail slice). */
#define CTCB200_WS_LOSS_GRAD_LOGITS 3
#define CTCB200_WS_HVP_LOGITS 4       /* ctcb200_hvp_logits */
#define CTCB200_WS_DECODE 5           /* ctcb200_greedy_decode */

typedef struct ctcb200_desc {
  int32_t B;       /* batch size                      (>= 0) */
  int32_t T;       /* logits.shape[1]                 (>= 0) */
  int32_t V;       /* number of tokens incl. blank    (>= 1) */
  int32_t Lw;      /* labels.shape[1]                 (>= 0) */
  int32_t blank;   /* blank index, 0 <= blank < V     (base_loss.py:122-125) */
  int32_t variant; /* CTCB200_CLASSIC | CTCB200_SIMPLIFIED */
  int32_t U;       /* max_b(label_length)+1 = the reference's max_label_length_plus_one (base_loss.py:478-486);
                      0 => use Lw+1 (label_length is then clamped to Lw) */
  uint32_t flags;  /* CTCB200_INPUT_LOGPROBAS | ... */
} ctcb200_desc;

int ctcb200_version(void);
const char* ctcb200_strerror(int code);
/* the CUDA runtime's message for the failure behind the last CTCB200_ERR_CUDA returned on the calling thread */
const char* ctcb200_last_cuda_error(void);

/* Comma-separated stage (kernel) names ctcb200_loss_grad runs for this descriptor (the fused kernel is a single
 * stage), and the number of kernels one full call enqueues. */
const char* ctcb200_stage_names(const ctcb200_desc* desc);
int ctcb200_launches_per_call(const ctcb200_desc* desc);

/* Developer / test hook: pins the fused kernel's plan (row workers per side, row buffers per worker, extra phase-A row
 * buffer 0/1, ring depth in frames, mode: bit 0 = split, one CTA per side in a two-CTA cluster; bit 1 = store every state
 * row instead of every second one; bit 2 = no idle warps in split plans; bit 3 = no row helpers, also honoured with
 * workers = 0) for every later call in this process, so a test can walk every plan on one shape; plans that do not fit
 * are ignored.  workers = 0 restores the built-in choice.  Not part of the reference-facing surface. */
void ctcb200_debug_fused_plan(int workers, int row_buffers, int extra_phase_a_buffer, int ring_depth, int mode);

/* Bytes of device workspace needed by the entry point named by `what` (CTCB200_WS_*); 0 on a bad descriptor. */
size_t ctcb200_workspace_bytes(const ctcb200_desc* desc, int what);

/*
 * Fused loss + gradient.  Replaces classic_ctc_loss / simplified_ctc_loss followed by tape.gradient
 * (tf_seq2seq_losses/classic_ctc_loss.py:33-70, simplified_ctc_loss.py:32-67, base_loss.py:38-99 and the
 * forward_fn/gradient_fn custom gradients base_loss.py:140-175, gradient base_loss.py:262-298).
 *   d_loss         [B] upstream gradient of the per-sample loss, or NULL for all-ones
 *   loss           [B] out: per-sample loss, +inf when the label cannot be emitted
 *   grad_logits    [B,T,V] out or NULL: d(sum_b d_loss[b]*loss[b]) / d logits  (= d_loss*(softmax*sum(occ) - occ))
 *   grad_logprobas [B,T,V] out or NULL: d_loss * ClassicCtcLossData.gradient   (= -d_loss*occupancy, base_loss.py:268)
 * With both gradient pointers NULL only the loss is computed (forward_fn, base_loss.py:140-155): where the fused kernel
 * serves the shape this costs half a call (the recursions meet in the middle, where the normaliser is known).
 */
int ctcb200_loss_grad(const ctcb200_desc* desc, const float* logits, const int32_t* labels,
                      const int32_t* label_length, const int32_t* logit_length, const float* d_loss, float* loss,
                      float* grad_logits, float* grad_logprobas, void* workspace, size_t workspace_bytes,
                      void* stream);

/*
 * log(-gradient) per (frame, token), computed in the log domain.  Replaces BaseCtcLossData.logarithmic_logproba_gradient
 * (base_loss.py:270-298 = loss + _combine_transition_probabilities, classic_ctc_loss.py:565-669 /
 * simplified_ctc_loss.py:456-534, with the segment log-sum-exp of tools.py:95-119): finite down to the smallest alignment
 * weight (where exp underflows, below about -87), -inf exactly where a (frame, token) pair is impossible, for frames
 * t >= logit_length and for infeasible samples.  log_gradient [B,T,V] out; loss [B] out or NULL; workspace sized with
 * CTCB200_WS_LOSS_GRAD.
 */
int ctcb200_log_gradient(const ctcb200_desc* desc, const float* logits, const int32_t* labels,
                         const int32_t* label_length, const int32_t* logit_length, float* loss, float* log_gradient,
                         void* workspace, size_t workspace_bytes, void* stream);

/*
 * alpha / beta state tensors.  Replaces ClassicCtcLossData.alpha/.beta (classic_ctc_loss.py:310-462, layout
 * [B,T+1,U,2], s=0 closed / s=1 open) and SimplifiedCtcLossData.alpha/.beta (simplified_ctc_loss.py:291-438,
 * layout [B,T+1,U]).  desc->U must be the true max(label_length)+1.  alpha, beta or loss may be NULL.
 */
int ctcb200_states(const ctcb200_desc* desc, const float* logits, const int32_t* labels,
                   const int32_t* label_length, const int32_t* logit_length, float* alpha, float* beta,
                   float* loss, void* workspace, size_t workspace_bytes, void* stream);

/*
 * gamma: log-probability of reaching state (t2, l2[, s2]) from state (t1, l1[, s1]); -inf for t2 < t1.  Replaces
 * ClassicCtcLossData.gamma (classic_ctc_loss.py:167-308, float32 [B,T+1,U,2,T+1,U,2]) and SimplifiedCtcLossData.gamma
 * (simplified_ctc_loss.py:85-191, float32 [B,T+1,U,T+1,U]).  desc->U must be the true max(label_length)+1 and at most
 * 128 (the tensor is O(T^2 U^2)); workspace sized with CTCB200_WS_STATES.
 */
int ctcb200_gamma(const ctcb200_desc* desc, const float* logits, const int32_t* labels, const int32_t* label_length,
                  const int32_t* logit_length, float* gamma, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Dense Hessian of the per-sample loss w.r.t. the log-probabilities, [B,T,V,T,V].  Replaces
 * BaseCtcLossData.hessian (base_loss.py:186-260) without materialising gamma (classic_ctc_loss.py:167-308,
 * simplified_ctc_loss.py:85-191).  Also returns loss [B] and gradient [B,T,V] (either may be NULL).
 */
int ctcb200_hessian(const ctcb200_desc* desc, const float* logits, const int32_t* labels,
                    const int32_t* label_length, const int32_t* logit_length, float* hessian, float* loss,
                    float* grad_logprobas, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Hessian-vector product: out[b,t,k] = sum_{t',k'} d_gradient[b,t',k'] * hessian[b,t,k,t',k'], the contraction
 * gradient_fn.backprop performs (base_loss.py:167-173), computed matrix-free.
 */
int ctcb200_hvp(const ctcb200_desc* desc, const float* logits, const int32_t* labels,
                const int32_t* label_length, const int32_t* logit_length, const float* d_gradient, float* out,
                void* workspace, size_t workspace_bytes, void* stream);

/*
 * Hessian-vector product w.r.t. the LOGITS: out[b] = d_loss[b] * (d2 loss[b] / d logits2) v[b], the quantity TF autodiff
 * produces when tape.gradient is taken of the first derivative of classic_ctc_loss / simplified_ctc_loss
 * (gradient_fn.backprop base_loss.py:167-173 chained through logit_to_logproba tools.py:27-40; README.md:58-71).
 * Matrix-free: neither the [B,T,V,T,V] Hessian nor gamma is formed.  `desc` must describe logits (no
 * CTCB200_INPUT_LOGPROBAS); d_loss may be NULL (ones); workspace sized with CTCB200_WS_HVP_LOGITS.
 */
int ctcb200_hvp_logits(const ctcb200_desc* desc, const float* logits, const int32_t* labels,
                       const int32_t* label_length, const int32_t* logit_length, const float* d_loss, const float* v,
                       float* out, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Greedy (best-path) decoding of the same logits, the inference-time step after the loss: per frame t < logit_length the
 * arg-max token (lowest index on ties), repeats merged when merge_repeated != 0, blanks dropped -- what
 * tf.nn.ctc_greedy_decoder returns for the logits classic_ctc_loss (classic_ctc_loss.py:33-70) is trained on, in dense
 * form: decoded [B,T] int32 padded with -1, decoded_length [B], neg_sum_logits [B] = -sum_t max_k logits (may be NULL).
 * desc->Lw, U and variant are ignored; CTCB200_TIME_MAJOR is honoured; workspace sized with CTCB200_WS_DECODE.
 */
int ctcb200_greedy_decode(const ctcb200_desc* desc, const float* logits, const int32_t* logit_length, int merge_repeated,
                          int32_t* decoded, int32_t* decoded_length, float* neg_sum_logits, void* workspace,
                          size_t workspace_bytes, void* stream);

/*
 * Host-buffer convenience path (what a framework without device tensors, or the end-to-end benchmark, calls):
 * owns its device buffers, copies the HOST inputs to the device in batch slices and pipelines three streams (host->device
 * copies, kernels, device->host copies) so both PCIe directions overlap the kernels, and copies loss (and optionally
 * grad_logits) back.  All pointers are HOST pointers (pinned memory gives full PCIe bandwidth).  The gradient stays resident on the device (ctcb200_host_grad_device_ptr) unless
 * host_grad_logits is non-NULL.  Blocks until the results are in the host buffers.
 */
typedef struct ctcb200_host_ctx ctcb200_host_ctx;
int ctcb200_host_create(const ctcb200_desc* desc, int device, int num_slices, ctcb200_host_ctx** out);
int ctcb200_host_loss_grad(ctcb200_host_ctx* ctx, const float* host_logits, const int32_t* host_labels,
                           const int32_t* host_label_length, const int32_t* host_logit_length,
                           float* host_loss, float* host_grad_logits);
float* ctcb200_host_grad_device_ptr(ctcb200_host_ctx* ctx);
void ctcb200_host_destroy(ctcb200_host_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* CTC_B200_H_ */
