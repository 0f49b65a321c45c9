// Instantiations of the fused kernel: classic variant, TMA row mover (one translation unit per combination so that they compile in parallel).
#include "kf_fused.cuh"

namespace ctcb200 {
CTCB200_DEFINE_FUSED_VARIANT(true, true, false)
}
