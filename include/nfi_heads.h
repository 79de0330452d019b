/*
 * nfi_heads.h -- C ABI of the SDF point evaluator behind the generator's regulariser heads
 * (SURVEY.md section 8f, N2).
 *
 * Reference: Generator.forward evaluates its decoder on 31^3 stratified points per image and
 * builds four losses from the first decoder output d(x) (/root/reference/models/generator.py:
 * 520-585): the eikonal loss needs grad_x d -- obtained there with torch.autograd.grad through
 * the decoder and a twice-differentiable bilinear fetch (lib/ops.py:58-120), i.e. a DOUBLE
 * backward when the loss is optimised -- the distance / total-variation / entropy losses need d.
 *
 * Here (d, grad_x d) is one fused forward evaluation (tri-plane fetch with spatial derivatives,
 * 32->64 softplus ->1, analytic gradient W2[0,:] diag(sigmoid(pre)) W1 dF/dx) and the backward of
 * BOTH outputs with respect to the planes and the decoder parameters is one more kernel with the
 * second-order terms written out (no autograd graph, no [B, N, 32] intermediates in HBM).  The
 * Python side (nerf_from_image_b200/heads.py) wraps the pair as a torch.autograd.Function and
 * forms the four losses with elementwise torch ops on [B, N] tensors.
 *
 * Conventions as in nfi_render.h: device pointers, fp32, stream as void*, 0 = success.
 */
#ifndef NFI_HEADS_H_
#define NFI_HEADS_H_

#include <stddef.h>
#include <stdint.h>

#include "nfi_render.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nfi_sdf_points_params {
  int32_t batch;     /* B */
  int32_t plane_res; /* R */
  float scene_range;
  int64_t n_points;    /* N points per image */
  const float *planes; /* [B,3,R,R,32] channel-last */
  const float *w1;     /* [64,32] EFFECTIVE decoder weights (EqualizedLinear gains applied) */
  const float *b1;     /* [64] */
  const float *w2;     /* [1+A,64]: only row 0 is read */
  const float *b2;     /* [1+A]:    only element 0 is read */
  const float *points; /* [B,N,3] world units */
  float *d;            /* out [B,N]   first decoder output (SDF, or pre-density) */
  float *grad;         /* out [B,N,3] d d / d point (world units), or NULL */
} nfi_sdf_points_params;

/* Upstream gradients in, parameter gradients out; every grad_* is ACCUMULATED into (+=). */
typedef struct nfi_sdf_points_grads {
  const float *g_d;    /* [B,N]   dL/d d      or NULL */
  const float *g_grad; /* [B,N,3] dL/d grad   or NULL */
  float *grad_planes;  /* [B,3,R,R,32] or NULL */
  float *grad_w1;      /* [64,32]      or NULL (all four decoder outputs together) */
  float *grad_b1;      /* [64] */
  float *grad_w2_row0; /* [64] */
  float *grad_b2_0;    /* [1]  */
} nfi_sdf_points_grads;

NFI_API int nfi_sdf_points_forward(const nfi_sdf_points_params *params, void *stream);
NFI_API int nfi_sdf_points_backward(const nfi_sdf_points_params *params,
                                    const nfi_sdf_points_grads *grads, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NFI_HEADS_H_ */
