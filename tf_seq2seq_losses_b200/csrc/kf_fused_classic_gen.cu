// Instantiations of the fused kernel: classic variant, rows moved by 4-byte cp.async (unaligned rows).
#include "kf_fused.cuh"

namespace ctcb200 {
CTCB200_DEFINE_FUSED_VARIANT(true, false)
}  // namespace ctcb200
