// Instantiations of the fused kernel: simplified variant, rows moved by 4-byte cp.async (unaligned rows).
#include "kf_fused.cuh"

namespace ctcb200 {
CTCB200_DEFINE_FUSED_VARIANT(false, false)
}  // namespace ctcb200
