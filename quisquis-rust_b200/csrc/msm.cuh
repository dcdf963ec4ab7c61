// Large multiscalar multiplication (Pippenger bucket method) kernels -- see DESIGN.md section "MSM".
#pragma once
#include "kernels.cuh"

namespace qq {
// (bucket kernels are added below by the Pippenger milestone; the small-n path in qq_api_msm.inc uses k_varbase)
}  // namespace qq
