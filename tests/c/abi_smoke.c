/* A C (not C++) caller of the C ABI in include/zsaac.h — what a non-Python host of the reference
 * would link.  Runs without a GPU: version, the dry planner, and the loud failure of zs_create
 * on a machine without an sm_100 device (exit code 0 = everything as documented). */
#include <stdio.h>
#include <string.h>

#include "zsaac.h"

int main(void) {
  int chunks = 0, tiles = 0, ctas = 0, window = 0, cg = 0;
  zs_ctx* ctx = NULL;
  int rc;
  if (zs_abi_version() != ZS_ABI_VERSION) return 1;
  if (strcmp(zs_kernel_name(), "zs_simtopk_kernel") != 0) return 2;
  /* BASELINE config 4 on a 148-SM device: 65,536 queries vs 10 M rows, top-32 */
  rc = zs_plan_dry(148, 10000000, 65536, 32, 0, &chunks, &tiles, &ctas, &window, &cg);
  if (rc != ZS_OK || chunks < 1 || tiles < 1 || ctas != 148 || cg != 2) return 3;
  printf("plan chunks=%d tiles_per_chunk=%d ctas=%d lockstep_window=%d cta_group=%d\n", chunks, tiles,
         ctas, window, cg);
  /* argument errors come back as codes with a message, never as a crash */
  if (zs_plan_dry(148, 10000000, 65536, 0, 0, &chunks, &tiles, &ctas, &window, &cg) != ZS_ERR_INVALID) return 4;
  if (zs_last_error() == NULL || zs_last_error()[0] == '\0') return 5;
  if (zs_create(NULL, 0) != ZS_ERR_INVALID) return 6;
  rc = zs_create(&ctx, 0);
  if (rc == ZS_OK) {               /* a B200 is present: the context works, destroy it again */
    printf("zs_create: ok\n");
    return zs_destroy(ctx) == ZS_OK ? 0 : 7;
  }
  if (rc != ZS_ERR_NO_DEVICE || ctx != NULL) return 8;
  printf("zs_create: %s\n", zs_last_error());
  return 0;
}
