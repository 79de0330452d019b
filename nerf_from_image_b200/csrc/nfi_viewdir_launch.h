// Launchers of nfi_viewdir.cu: the fp32 SIMT render kernels instantiated with view-direction
// conditioning (--use_viewdir, CARLA; models/generator.py:189-253,662-663), a translation unit
// of its own compiled in parallel with the others by build.sh.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "nfi_render.h"

namespace nfi {
// `p.workspace` already points at the SIMT scratch slabs (after the weight-image header).
int launch_forward_viewdir(const nfi_render_params& p, int nout_pad, bool normals, cudaStream_t st,
                           char* err, size_t err_len);
int launch_backward_viewdir(const nfi_render_params& p, const nfi_render_grads& g, int nout_pad,
                            cudaStream_t st, char* err, size_t err_len);
}  // namespace nfi
