/* Plain-C consumer of include/nfi_render.h, nfi_synth.h and nfi_heads.h: proves the header is valid C99 (no C++-isms, no
 * torch types), that the structs have the layout the ctypes mirror assumes, and that the
 * library links and reports errors through return codes without a GPU.
 * Built and run by tests/test_abi.py::test_header_is_plain_c_and_links. */
#include <stdio.h>
#include <stddef.h>
#include <string.h>

#include "nfi_render.h"
#include "nfi_heads.h"
#include "nfi_synth.h"

int main(void) {
  nfi_render_params p;
  nfi_render_grads g;
  nfi_sample_params s;
  nfi_synth_params y;
  memset(&y, 0, sizeof y);
  memset(&p, 0, sizeof p);
  memset(&g, 0, sizeof g);
  memset(&s, 0, sizeof s);
  printf("abi %d\n", nfi_abi_version());
  printf("build %s\n", nfi_build_info());
  printf("sizeof params %zu grads %zu sample %zu\n", sizeof p, sizeof g, sizeof s);
  printf("offsets planes %zu workspace %zu noise_seed %zu points %zu\n",
         offsetof(nfi_render_params, planes), offsetof(nfi_render_params, workspace),
         offsetof(nfi_render_params, noise_seed), offsetof(nfi_sample_params, points));
  printf("synth sizeof %zu layer %zu offsets ws %zu conv1 %zu planes %zu\n", sizeof y,
         sizeof(nfi_synth_layer), offsetof(nfi_synth_params, ws), offsetof(nfi_synth_params, conv1),
         offsetof(nfi_synth_params, planes));
  if (nfi_abi_version() != NFI_ABI_VERSION) return 10;
  {
    nfi_sdf_points_params hp;
    nfi_sdf_points_grads hg;
    memset(&hp, 0, sizeof hp);
    memset(&hg, 0, sizeof hg);
    if (nfi_sdf_points_forward(&hp, NULL) == 0 || nfi_sdf_points_backward(&hp, &hg, NULL) == 0) return 20;
  }
  if (nfi_synthesis_forward(&y, NULL) == 0 || nfi_synthesis_workspace_bytes(&y) != 0) return 19;
  /* every entry point rejects an empty request with a non-zero code and a message */
  if (nfi_render_forward(&p, NULL) == 0 || strlen(nfi_last_error()) == 0) return 11;
  if (nfi_render_backward(&p, &g, NULL) == 0) return 12;
  if (nfi_sample_field(&s, NULL) == 0) return 13;
  if (nfi_pose_to_matrix(NULL, NULL, NULL, NULL, 0, 4, NULL, NULL, NULL) == 0) return 14;
  if (nfi_pose_to_matrix_backward(NULL, NULL, NULL, NULL, 0, 4, NULL, NULL, NULL, NULL, NULL,
                                  NULL, NULL) == 0) return 15;
  if (nfi_render_workspace_bytes(NULL) != 0) return 16;
  if (nfi_fill_uniform(NULL, 0, 0, 0, 0, NULL) == 0) return 17;
  if (nfi_decoder_forward(NULL, 0, NULL, NULL, NULL, NULL, 10, NULL, NFI_MLP_FP32_SIMT, NULL,
                          NULL) == 0) return 18;
  printf("last error: %s\n", nfi_last_error());
  return 0;
}
