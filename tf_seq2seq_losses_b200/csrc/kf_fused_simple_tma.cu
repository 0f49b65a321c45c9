// Instantiations of the fused kernel: simplified variant, TMA row mover (one translation unit per combination so that they compile in parallel).
#include "kf_fused.cuh"

namespace ctcb200 {
CTCB200_DEFINE_FUSED_VARIANT(false, true, false)
}
