// placeholder until the tcgen05 encoder lands
#include "common.cuh"
extern "C" {
int pm_has_tcgen05(void) { return 0; }
size_t pm_pointnet_encode_forward_tc_ws_bytes(int, int, int) { return 0; }
int pm_pointnet_encode_forward_tc(const float*, int64_t, int, int, int, const pm_encoder_params*, int, float*, int64_t,
                                  int32_t*, void*, size_t, pm_stream_t) {
  PM_FAIL(PM_ERR_UNSUPPORTED, "tcgen05 encoder not built");
}
}
