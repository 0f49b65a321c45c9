/* A host written in plain C binds libctc_b200.so through include/ctc_b200.h alone (no CUDA or torch headers): sizes a
 * workspace and reads the diagnostics.  Compiled and run by tests/test_abi_cpu.py; no kernel is launched. */
#include <stdio.h>
#include <string.h>

#include "ctc_b200.h"

int main(void) {
  ctcb200_desc d = {256, 1000, 1024, 200, 0, CTCB200_SIMPLIFIED, 201, 0};
  const size_t small = ctcb200_workspace_bytes(&d, CTCB200_WS_LOSS_GRAD_LOGITS);
  const size_t full = ctcb200_workspace_bytes(&d, CTCB200_WS_LOSS_GRAD);
  if (small == 0 || full < small) return 1;
  if (strcmp(ctcb200_stage_names(&d), "kf_fused") != 0) return 2;
  if (strcmp(ctcb200_strerror(CTCB200_OK), "ok") != 0) return 3;
  if (ctcb200_version() != CTCB200_VERSION) return 4;
  /* a NULL workspace for a non-empty batch is an error code, never a crash */
  if (ctcb200_loss_grad(&d, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL) != CTCB200_ERR_NULL_POINTER) return 5;
  printf("%zu %zu\n", small, full);
  return 0;
}
