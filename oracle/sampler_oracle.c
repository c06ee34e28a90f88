/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see mwpm_oracle.c header for the rules).
 *
 * CPU restatement of what sinter's worker does with Stim for this path (SURVEY.md section 8a,
 * rows A7/A9): draw every DEM error mechanism independently with its probability, form
 * detectors = H.e and observables = L.e over GF(2), and count shots whose prediction differs from
 * the observable.  PARITY UNPINNED against Stim itself (stim>=1.14.0,
 * /root/reference/pyproject.toml:10, is not installed here); the DEM it samples is pinned by the
 * notebook circuits and SURVEY Appendix C (tests/test_dem.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t *s) { return ((splitmix64(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

/* Dense GF(2) product on explicit error vectors.
 * errs: uint8[nshots][M] (0/1).  dets: uint8[nshots][ceil(D/8)], obs: uint8[nshots][ceil(O/8)] (zeroed here). */
void orc_extract(int D, int O, int M, const uint32_t *det_indptr, const uint32_t *det_idx,
                 const uint32_t *obs_indptr, const uint32_t *obs_idx, const uint8_t *errs,
                 long nshots, uint8_t *dets, uint8_t *obs) {
    int db = (D + 7) / 8, ob = (O + 7) / 8;
    memset(dets, 0, (size_t)nshots * db);
    memset(obs, 0, (size_t)nshots * ob);
    for (long s = 0; s < nshots; ++s) {
        const uint8_t *e = errs + (size_t)s * M;
        uint8_t *dr = dets + (size_t)s * db, *orow = obs + (size_t)s * ob;
        for (int m = 0; m < M; ++m) {
            if (!e[m]) continue;
            for (uint32_t k = det_indptr[m]; k < det_indptr[m + 1]; ++k) dr[det_idx[k] >> 3] ^= 1u << (det_idx[k] & 7);
            for (uint32_t k = obs_indptr[m]; k < obs_indptr[m + 1]; ++k) orow[obs_idx[k] >> 3] ^= 1u << (obs_idx[k] & 7);
        }
    }
}

/* Sample nshots shots of the DEM (geometric skipping per mechanism over the shot axis). */
void orc_sample(int D, int O, int M, const double *p, const uint32_t *det_indptr,
                const uint32_t *det_idx, const uint32_t *obs_indptr, const uint32_t *obs_idx,
                uint64_t seed, long nshots, uint8_t *dets, uint8_t *obs) {
    int db = (D + 7) / 8, ob = (O + 7) / 8;
    memset(dets, 0, (size_t)nshots * db);
    memset(obs, 0, (size_t)nshots * ob);
    uint64_t st = seed * 0x2545F4914F6CDD1DULL + 0x1234567ULL;
    for (int m = 0; m < M; ++m) {
        double pm = p[m];
        if (pm <= 0) continue;
        double lq = pm >= 1 ? -INFINITY : log1p(-pm);
        long s = -1;
        for (;;) {
            double u = u01(&st);
            double gap = pm >= 1 ? 0 : floor(log(u) / lq);
            if (gap > (double)(nshots - s)) break;
            s += 1 + (long)gap;
            if (s >= nshots) break;
            uint8_t *dr = dets + (size_t)s * db, *orow = obs + (size_t)s * ob;
            for (uint32_t k = det_indptr[m]; k < det_indptr[m + 1]; ++k) dr[det_idx[k] >> 3] ^= 1u << (det_idx[k] & 7);
            for (uint32_t k = obs_indptr[m]; k < obs_indptr[m + 1]; ++k) orow[obs_idx[k] >> 3] ^= 1u << (obs_idx[k] & 7);
        }
    }
}

/* rows of pred (uint8[nshots][ob]) that differ from obs */
long orc_count_errors(const uint8_t *pred, const uint8_t *obs, long nshots, int ob) {
    long n = 0;
    for (long s = 0; s < nshots; ++s) n += memcmp(pred + (size_t)s * ob, obs + (size_t)s * ob, ob) != 0;
    return n;
}
