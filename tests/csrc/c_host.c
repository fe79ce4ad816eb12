/* c_host.c - TEST-ONLY: the C-ABI (include/g2o_b200.h) driven from plain C99, the way a non-C++ host would bind it.
 * Builds a small SE2 pose graph through the standalone host API, runs the structure phase on a host-only context
 * (device -1) and prints the block ordering; every compute call must answer B200_ERR_NO_DEVICE there. */
#include <stdio.h>
#include <stdlib.h>

#include "g2o_b200.h"

int main(void) {
  b200_graph* g = NULL;
  b200_ctx* ctx = NULL;
  int32_t dims[8], perm[8];
  int i, rc;
  double chi2 = 0.0;
  if (b200_graph_create(&g) != B200_OK) return 1;
  for (i = 0; i < 6; ++i) {
    double est[3];
    est[0] = i; est[1] = 0.1 * i; est[2] = 0.05 * i;
    if (b200_graph_add_vertex(g, B200_VERTEX_SE2, i, est, 3) != B200_OK) return 2;
  }
  for (i = 0; i < 6; ++i) {  /* ring: 0-1-2-3-4-5-0 */
    double pay[9] = {1.0, 0.1, 0.05, 10, 0, 0, 10, 0, 10};
    if (b200_graph_add_edge(g, B200_EDGE_SE2, i, (i + 1) % 6, pay, 9) != B200_OK) return 3;
  }
  if (b200_graph_setup_cli(g, 1) < 0) return 4;   /* a gauge vertex gets fixed: one of the six ids */
  if (b200_graph_initialize(g) != B200_OK) return 5;
  if (b200_create(-1, &ctx) != B200_OK) return 6;
  if (b200_graph_upload(g, ctx, 0, 1) != B200_OK) return 7;
  if (b200_build_structure(ctx) != B200_OK) { fprintf(stderr, "%s\n", b200_last_error(ctx)); return 8; }
  if (b200_get_dims(ctx, dims) != B200_OK) return 9;
  if (b200_get_block_ordering(ctx, perm) != dims[0]) return 10;  /* returns the number of blocks */
  rc = b200_compute_active_errors(ctx, &chi2);
  printf("poses=%d edges=%d lnz=%lld nodevice=%d perm=", (int)dims[0], (int)dims[4],
         (long long)b200_get_factor_nnz(ctx), rc == B200_ERR_NO_DEVICE);
  for (i = 0; i < dims[0]; ++i) printf("%d%s", (int)perm[i], i + 1 < dims[0] ? "," : "\n");
  printf("%s\n", b200_version());
  b200_destroy(ctx);
  b200_graph_destroy(g);
  return 0;
}
