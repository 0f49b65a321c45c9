// Instantiations of the fused kernel: simplified variant, rows moved by 1-D TMA.
#include "kf_fused.cuh"

namespace ctcb200 {
CTCB200_DEFINE_FUSED_VARIANT(false, true)
}  // namespace ctcb200
