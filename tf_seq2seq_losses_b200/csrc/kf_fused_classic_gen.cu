// Instantiations of the fused kernel: classic variant, 4-byte cp.async row mover (one translation unit per combination so that they compile in parallel).
#include "kf_fused.cuh"

namespace ctcb200 {
CTCB200_DEFINE_FUSED_VARIANT(true, false, false)
}
