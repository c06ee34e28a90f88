// Test scaffolding: compiles the CUDA decoder's matching core (blossom_warp.cuh) for the host with a
// single "lane" so that its logic can be checked on a CPU-only box.  Not part of the product.
#define GQ_HOST_SIM 1
#include <cstdlib>
#include <cstring>
#include "../../graphqec_b200/csrc/blossom_warp.cuh"

extern "C" int hostsim_solve(int n, const uint16_t *W, int16_t *mate_out, long long *weight_out) {
    int cap = (n + 1) & ~1;
    if (cap < 2) cap = 2;
    size_t bytes = gqb::ws_bytes(cap);
    void *mem = calloc(1, bytes + 64);
    gqb::Solver S;
    gqb::ws_bind(S.w, mem, cap);
    S.n = n; S.cap = cap; S.nb = cap / 2;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) S.w.W[i * S.w.wstride + j] = W[i * n + j];
    int rc = S.solve();
    long long tot = 0;
    for (int i = 0; i < n; ++i) {
        mate_out[i] = S.w.mate[i];
        if (rc == 0 && S.w.mate[i] > i) tot += W[i * n + S.w.mate[i]];
    }
    *weight_out = tot;
    free(mem);
    return rc;
}

extern "C" void hostsim_counters(long long *out) {
    for (int i = 0; i < 16; ++i) out[i] = gq_counters[i];
}
