// pbd_batch.cu -- batch-of-bodies entry points (placeholder).
#include "pbd_body.h"
extern "C" {
pbd_batch* pbd_batch_create(const pbd_params*, uint32_t, const uint64_t*, const uint64_t*, const uint64_t*, const float*,
                            const uint32_t*, const uint32_t*, int, const pbd_options*, int* status) {
  if (status) *status = PBD_ERR_UNSUPPORTED;
  return nullptr;
}
int pbd_batch_step(pbd_batch*, float, pbd_step_stats*) { return PBD_ERR_UNSUPPORTED; }
int pbd_batch_step_async(pbd_batch*, float, uint32_t) { return PBD_ERR_UNSUPPORTED; }
int pbd_batch_sync(pbd_batch*, double*) { return PBD_ERR_UNSUPPORTED; }
int pbd_batch_read_positions(pbd_batch*, float*, double*) { return PBD_ERR_UNSUPPORTED; }
int pbd_batch_get_info(const pbd_batch*, pbd_info*) { return PBD_ERR_UNSUPPORTED; }
void pbd_batch_destroy(pbd_batch*) {}
}
