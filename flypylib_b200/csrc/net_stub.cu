// temporary placeholders so the library links while the conv stack is being written
#include "common.cuh"
extern "C" {
int fpl_net_create(fpl_ctx *, int, fpl_net **) { fpl::set_error("net: not built yet"); return FPL_ESTATE; }
int fpl_net_destroy(fpl_net *) { return FPL_OK; }
int fpl_net_info(const fpl_net *, int32_t *, int32_t *, int32_t *, int32_t *) { return FPL_ESTATE; }
int fpl_net_num_weights(const fpl_net *, int32_t *) { return FPL_ESTATE; }
int fpl_net_weight_size(const fpl_net *, int32_t, int64_t *) { return FPL_ESTATE; }
int fpl_net_set_weights(fpl_net *, const float *const *, int32_t, int) { return FPL_ESTATE; }
int fpl_net_out_size(const fpl_net *, int32_t, int32_t *) { return FPL_ESTATE; }
int fpl_net_forward_tiles(fpl_net *, const float *, int32_t, int32_t, float *, void *) { return FPL_ESTATE; }
int fpl_net_infer_volume(fpl_net *, const void *, int, float, float, int64_t, int64_t, int64_t, int32_t,
                         int32_t, float *, void *) { return FPL_ESTATE; }
}
