// Launchers of nfi_field.cu (the sampler seam and the pose kernels), a translation unit of its
// own compiled in parallel with the others by build.sh.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "nfi_render.h"

namespace nfi {
// `p` carries the field (planes, decoder, palette, beta/alpha, plane_res, n_attention, use_sdf,
// scene_range); `io` the points and the requested outputs.
int launch_sample_field(const nfi_render_params& p, const nfi_sample_params& io, int nout_pad,
                        cudaStream_t st, char* err, size_t err_len);
int launch_pose_to_matrix(const float* z0, const float* t2, const float* s, const float* q,
                          int flipped, int batch, float* c2w, float* focal, cudaStream_t st,
                          char* err, size_t err_len);
int launch_pose_to_matrix_backward(const float* z0, const float* t2, const float* s,
                                   const float* q, int flipped, int batch, const float* g_c2w,
                                   const float* g_focal, float* g_z0, float* g_t2, float* g_s,
                                   float* g_q, cudaStream_t st, char* err, size_t err_len);
}  // namespace nfi
