// Test scaffolding: compiles the first-tier matching decoder (match_tree.cuh: dual start, single-tree
// blossom phases with native boundary, prediction) for the host with one lane.  Not part of the product.
#include "warp_sim.h"
#include <cstdlib>
#include <cstring>
#include <vector>
#define GQ_MT_STATS 1
static long long gqm_stats[16];
static inline void gqm_stat_add(int i, long long x) { gqm_stats[i] += x; }
#include "../../graphqec_b200/csrc/match_tree.cuh"

#ifndef HS_CAP
#define HS_CAP 128
#endif

// dist/pobs: [D+2][D+2] tables as the CUDA decoder builds them (row/col D = boundary, D+1 = nothing).
// dets: nshots rows of DW uint32 words.  Outputs per shot: prediction mask, weight, return code
// (0 ok, 1 no perfect matching, 2/3 solver gave up, 4 more defects than the tier holds).
extern "C" int hostsim_mt_decode(int D, const uint16_t *dist, const uint8_t *pobs, int DW, const uint32_t *dets,
                                 long nshots, uint32_t *pred, int32_t *weight, int32_t *rcs, int16_t *mates_out,
                                 const uint32_t *nbr) {
    gqt::Tables T;
    T.dist = dist; T.pobs = pobs; T.D = (uint32_t)D; T.N = (uint32_t)D + 2; T.nbr = nbr;
    const bool mirror = nbr && DW <= gqt::SYN_WORDS;
    std::vector<uint32_t> synw(gqt::SYN_WORDS + 1); std::vector<uint16_t> spre(gqt::SYN_WORDS + 2);
    std::vector<uint16_t> def(HS_CAP);
    static gqr::StaticForest<HS_CAP> F;
    for (long s = 0; s < nshots; ++s) {
        int n = gqt::extract_defects(dets + (size_t)s * DW, (uint32_t)DW, (uint32_t)D, 0, def.data(), HS_CAP, mirror ? synw.data() : nullptr, spre.data());
        uint32_t pm = 0; int wt = 0; int rc;
        if (n > HS_CAP) rc = 4;
        else rc = gqt::decode_shot<HS_CAP>(T, F, def.data(), n, 0, true, &pm, &wt, mirror ? synw.data() : nullptr, spre.data());
        pred[s] = rc == 0 ? pm : 0; weight[s] = rc == 0 ? wt : -1; rcs[s] = rc;
        if (mates_out) for (int i = 0; i < HS_CAP; ++i) mates_out[(size_t)s * HS_CAP + i] = i < n ? F.mate[i] : -3;
    }
    return 0;
}

extern "C" void hostsim_mt_stats(long long *out, int reset) {
    for (int i = 0; i < 16; ++i) out[i] = gqm_stats[i];
    if (reset) memset(gqm_stats, 0, sizeof(gqm_stats));
}
