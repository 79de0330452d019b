// Launchers of nfi_heads.cu (SDF point evaluator of the regulariser heads), its own translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "nfi_heads.h"

namespace nfi {
namespace heads {
int launch_forward(const nfi_sdf_points_params& p, cudaStream_t st, char* err, size_t err_len);
int launch_backward(const nfi_sdf_points_params& p, const nfi_sdf_points_grads& g, cudaStream_t st,
                    char* err, size_t err_len);
}  // namespace heads
}  // namespace nfi
