/* A non-Python host of the C ABI (tests/test_cabi.py compiles this with gcc -std=c99 and runs it without a GPU):
 * include/mvin_b200.h is plain C, the library links like any shared object, configuration errors come back as
 * negative codes with a message in mvin_last_error() -- no compute call is made. */
#include <stdio.h>
#include <string.h>

#include "mvin_b200.h"

int main(void) {
  mvin_config_t cfg;
  mvin_handle_t h = 0;
  int rc;
  if (mvin_abi_version() != MVIN_ABI_VERSION) {
    printf("abi mismatch: library %d, header %d\n", mvin_abi_version(), MVIN_ABI_VERSION);
    return 1;
  }
  memset(&cfg, 0, sizeof(cfg));
  cfg.dim = 24; /* not in {8,16,32,64,128} */
  cfg.neighbor_sample_size = 8;
  cfg.h_hop = 2;
  cfg.n_mix_hop = 1;
  cfg.p_hop = 2;
  cfg.n_memory = 16;
  cfg.n_user = 4;
  cfg.n_entity = 9;
  cfg.n_relation = 3;
  cfg.max_batch = 8;
  cfg.l2_weight = 1e-4f;
  cfg.l2_agg_weight = 1e-6f;
  cfg.flags = 0x1f;
  rc = mvin_create(&cfg, &h);
  if (rc != MVIN_ERR_UNSUPPORTED || h != 0) {
    printf("expected MVIN_ERR_UNSUPPORTED, got %d\n", rc);
    return 2;
  }
  if (!mvin_last_error() || !strlen(mvin_last_error())) return 3;
  printf("abi %d; refused: %s\n", mvin_abi_version(), mvin_last_error());
  return 0;
}
