// K3b: min-sum belief propagation on the DEM Tanner graph (placeholder until the kernel lands).
#include "gq_common.h"

int gq_bp_create(gq_dec *) { GQ_FAIL(GQ_E_UNSUPPORTED, "BP decoder: not built yet"); }
void gq_bp_destroy(gq_dec *) {}
int gq_bp_decode(gq_dec *, const uint32_t *, uint64_t, uint32_t *, int32_t *, uint8_t *, uint32_t *, int *) {
    GQ_FAIL(GQ_E_UNSUPPORTED, "BP decoder: not built yet");
}
