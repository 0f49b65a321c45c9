/* Body of the C restatement, compiled twice: REAL = double (checker) and REAL = float (reference-like arithmetic,
 * CPU baseline).  TEST INFRASTRUCTURE ONLY -- see ctc_oracle.c. */

static REAL FN(lse2)(REAL x, REAL y) {
  /* tf_seq2seq_losses/tools.py:57-71 */
  REAL m = x > y ? x : y, n = x > y ? y : x;
  if (m == -INFINITY) return -INFINITY;
  return m + (REAL)log1p(exp((double)(n - m)));
}

/* One utterance.  logits [T,V] float; writes loss and (optionally) grad [T,V] = d loss / d logits. */
static void FN(utterance)(int variant, int T, int V, int Lw, int U, int blank, const float* logits,
                          const int* labels, int label_length, int logit_length, REAL* loss_out, REAL* grad) {
  const int S = variant == 0 ? 2 : 1;
  int n_t = logit_length < 0 ? 0 : (logit_length > T ? T : logit_length);
  int L = label_length < 0 ? 0 : (label_length > U - 1 ? U - 1 : label_length);
  int* lab = (int*)malloc(sizeof(int) * (size_t)(U + 1));
  /* _cleaned_label, base_loss.py:395-418 */
  for (int l = 0; l < U; ++l) lab[l] = (l < L && l < Lw) ? labels[l] : blank;
  REAL* lp = (REAL*)malloc(sizeof(REAL) * (size_t)(T > 0 ? T : 1) * V);
  REAL* d = (REAL*)malloc(sizeof(REAL) * (size_t)(T > 0 ? T : 1) * U);
  REAL* h = (REAL*)malloc(sizeof(REAL) * (size_t)(T > 0 ? T : 1));
  REAL* alpha = (REAL*)malloc(sizeof(REAL) * (size_t)(T + 1) * U * S);
  REAL* beta = (REAL*)malloc(sizeof(REAL) * (size_t)(T + 1) * U * S);
  /* logit_to_logproba tools.py:27-40 + _logproba base_loss.py:378-393 + gathers base_loss.py:328-371 */
  for (int t = 0; t < T; ++t) {
    REAL* row = lp + (size_t)t * V;
    if (t < n_t) {
      const float* x = logits + (size_t)t * V;
      REAL m = -INFINITY;
      for (int k = 0; k < V; ++k) if ((REAL)x[k] > m) m = (REAL)x[k];
      if (!isfinite((double)m)) m = 0;
      REAL s = 0;
      for (int k = 0; k < V; ++k) s += (REAL)exp((double)((REAL)x[k] - m));
      REAL lse = m + (REAL)log((double)s);
      for (int k = 0; k < V; ++k) row[k] = (REAL)x[k] - lse;
    } else {
      for (int k = 0; k < V; ++k) row[k] = -INFINITY;
      row[blank] = 0;
    }
    h[t] = row[blank];
    for (int l = 0; l < U; ++l) {
      int tok = lab[l];
      d[(size_t)t * U + l] = (l < L && tok >= 0 && tok < V) ? row[tok] : -INFINITY;
    }
  }
#define A(t, l, s) alpha[((size_t)(t) * U + (l)) * S + (s)]
#define Bt(t, l, s) beta[((size_t)(t) * U + (l)) * S + (s)]
#define D(t, l) d[(size_t)(t) * U + (l)]
  for (int l = 0; l < U; ++l)
    for (int s = 0; s < S; ++s) { A(0, l, s) = -INFINITY; Bt(T, l, s) = (l == L) ? 0 : -INFINITY; }
  A(0, 0, 0) = 0;
  if (variant == 1) {
    /* simplified_ctc_loss.py:393-424, :327-343 */
    for (int t = 0; t < T; ++t)
      for (int l = 0; l < U; ++l)
        A(t + 1, l, 0) = FN(lse2)(h[t] + A(t, l, 0), l > 0 ? D(t, l - 1) + A(t, l - 1, 0) : -INFINITY);
    for (int t = T - 1; t >= 0; --t)
      for (int l = 0; l < U; ++l)
        Bt(t, l, 0) = FN(lse2)(h[t] + Bt(t + 1, l, 0), l + 1 < U ? D(t, l) + Bt(t + 1, l + 1, 0) : -INFINITY);
    *loss_out = -A(T, L, 0);
  } else {
    /* classic_ctc_loss.py:415-451, :349-364 with the tables of :464-563 */
    for (int t = 0; t < T; ++t)
      for (int l = 0; l < U; ++l) {
        A(t + 1, l, 0) = h[t] + FN(lse2)(A(t, l, 0), A(t, l, 1));
        REAL v = -INFINITY;
        if (l > 0) {
          int prev = lab[l - 1], pprev = l > 1 ? lab[l - 2] : blank;
          REAL r = prev != blank ? lp[(size_t)t * V + prev] : -INFINITY;      /* re-emit label[l-1] */
          if (!(prev >= 0 && prev < V)) r = -INFINITY;
          REAL d1 = (prev != pprev) ? D(t, l - 1) : -INFINITY;                /* open -> open only if no repeat */
          v = FN(lse2)(r + A(t, l, 1), FN(lse2)(D(t, l - 1) + A(t, l - 1, 0), d1 + A(t, l - 1, 1)));
        }
        A(t + 1, l, 1) = v;
      }
    for (int t = T - 1; t >= 0; --t)
      for (int l = 0; l < U; ++l) {
        REAL nx = l + 1 < U ? Bt(t + 1, l + 1, 1) : -INFINITY;
        int prev = l > 0 ? lab[l - 1] : blank;
        REAL r = (l > 0 && prev != blank && prev >= 0 && prev < V) ? lp[(size_t)t * V + prev] : -INFINITY;
        REAL d1 = (lab[l] != prev) ? D(t, l) : -INFINITY;
        Bt(t, l, 0) = FN(lse2)(h[t] + Bt(t + 1, l, 0), D(t, l) + nx);
        Bt(t, l, 1) = FN(lse2)(h[t] + Bt(t + 1, l, 0), FN(lse2)(r + Bt(t + 1, l, 1), d1 + nx));
      }
    *loss_out = -FN(lse2)(A(T, L, 0), A(T, L, 1));
  }
  if (grad != NULL) {
    const REAL loss = *loss_out;
    REAL* c = (REAL*)malloc(sizeof(REAL) * (size_t)V);
    for (int t = 0; t < T; ++t) {
      REAL* g = grad + (size_t)t * V;
      if (loss == INFINITY || t >= n_t) {       /* base_loss.py:284-295 */
        for (int k = 0; k < V; ++k) g[k] = 0;
        continue;
      }
      /* _combine_transition_probabilities: classic_ctc_loss.py:565-669 / simplified_ctc_loss.py:456-534 */
      for (int k = 0; k < V; ++k) c[k] = -INFINITY;
      REAL cb = -INFINITY;
      for (int l = 0; l < U; ++l) {
        if (variant == 1) {
          cb = FN(lse2)(cb, A(t, l, 0) + Bt(t + 1, l, 0));
          int tok = lab[l];
          if (l + 1 < U && tok >= 0 && tok < V) c[tok] = FN(lse2)(c[tok], A(t, l, 0) + D(t, l) + Bt(t + 1, l + 1, 0));
        } else {
          cb = FN(lse2)(cb, FN(lse2)(A(t, l, 0), A(t, l, 1)) + Bt(t + 1, l, 0));
          int tok = lab[l], prev = l > 0 ? lab[l - 1] : blank;
          if (prev >= 0 && prev < V)
            c[prev] = FN(lse2)(c[prev], A(t, l, 1) + lp[(size_t)t * V + prev] + Bt(t + 1, l, 1));
          if (l + 1 < U && tok >= 0 && tok < V) {
            REAL d1 = (tok != prev) ? D(t, l) : -INFINITY;
            c[tok] = FN(lse2)(c[tok], FN(lse2)(A(t, l, 0) + D(t, l), A(t, l, 1) + d1) + Bt(t + 1, l + 1, 1));
          }
        }
      }
      c[blank] = h[t] + cb;                     /* the blank column overrides (blank_mask tf.where) */
      /* gradient = -exp(loss + c) (base_loss.py:262-298), then the log-softmax chain (tools.py:37-39) */
      REAL sum = 0;
      for (int k = 0; k < V; ++k) { g[k] = -(REAL)exp((double)(loss + c[k])); sum += g[k]; }
      const REAL* row = lp + (size_t)t * V;
      for (int k = 0; k < V; ++k) g[k] = g[k] - (REAL)exp((double)row[k]) * sum;
    }
    free(c);
  }
#undef A
#undef Bt
#undef D
  free(lab); free(lp); free(d); free(h); free(alpha); free(beta);
}

typedef struct {
  int variant, B, T, V, Lw, blank, U;
  const float* logits; const int* labels; const int* label_length; const int* logit_length;
  REAL* loss; REAL* grad;
  int* next;                                     /* shared work counter: utterances are independent */
} FN(job);

static void* FN(worker)(void* arg) {
  FN(job)* j = (FN(job)*)arg;
  for (;;) {
    int b = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
    if (b >= j->B) break;
    FN(utterance)(j->variant, j->T, j->V, j->Lw, j->U, j->blank, j->logits + (size_t)b * j->T * j->V,
                  j->labels + (size_t)b * j->Lw, j->label_length[b], j->logit_length[b], j->loss + b,
                  j->grad ? j->grad + (size_t)b * j->T * j->V : NULL);
  }
  return NULL;
}

int FN(ctc_oracle_loss_grad)(int variant, int B, int T, int V, int Lw, int blank, const float* logits,
                             const int* labels, const int* label_length, const int* logit_length, REAL* loss,
                             REAL* grad, int num_threads) {
  if (variant != 0 && variant != 1) return -1;
  int lmax = 0;                                  /* base_loss.py:482-486 */
  for (int b = 0; b < B; ++b) if (label_length[b] > lmax) lmax = label_length[b];
  int next = 0;
  FN(job) j = {variant, B, T, V, Lw, blank, lmax + 1, logits, labels, label_length, logit_length, loss, grad, &next};
  if (num_threads <= 0) num_threads = ctc_oracle_max_threads();
  if (num_threads > B) num_threads = B > 0 ? B : 1;
  if (num_threads > 256) num_threads = 256;
  pthread_t th[256];
  int started = 0;
  for (int i = 1; i < num_threads; ++i)
    if (pthread_create(&th[started], NULL, FN(worker), &j) == 0) ++started;
  FN(worker)(&j);
  for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
  return 0;
}
