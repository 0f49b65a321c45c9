/*
 * C restatement of the CTC loss + gradient path of alexeytochin/tf_seq2seq_losses (CPU, one pthread per core over the batch).
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ as a full-size checker and by bench.py's cpu_baseline / --impl reference
 * legs.  The product (tf_seq2seq_losses_b200/) never links or loads it.  Pinned against the numpy oracle
 * (oracle/ctc_oracle.py, itself pinned against the reference's known-answer tests) in tests/test_oracle_c.py, and against
 * outputs of the reference's own code (tests/golden/reference/reference_outputs.npz) in tests/test_reference_golden.py.
 *
 * Exposes ctc_oracle_loss_grad_f64 (double arithmetic, the checker) and ctc_oracle_loss_grad_f32 (float arithmetic,
 * like the reference, which asserts float32: tf_seq2seq_losses/base_loss.py:131).
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

int ctc_oracle_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

#define REAL double
#define FN(name) name##_f64
#include "ctc_oracle_impl.h"
#undef REAL
#undef FN

#define REAL float
#define FN(name) name##_f32
#include "ctc_oracle_impl.h"
#undef REAL
#undef FN

