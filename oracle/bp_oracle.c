/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see mwpm_oracle.c header for the rules).
 *
 * Min-sum belief propagation on the DEM Tanner graph (checks = detectors, variables = error
 * mechanisms), the decoder BASELINE.json's north_star names for non-matchable codes.  The
 * reference's own decoder for that config is BP+OSD from stimbposd/ldpc (unpinned,
 * /root/reference/pyproject.toml:13; threshold_lab.py:19,30), which is not installed: PARITY
 * UNPINNED.  This file fixes the schedule the CUDA kernel must reproduce bit-for-bit in fp32:
 * flooding schedule, messages initialised to the prior LLRs, check update
 * m_{c->v} = alpha * (-1)^{s_c} * prod_{v'!=v} sign(m_{v'->c}) * min_{v'!=v} |m_{v'->c}|,
 * variable update m_{v->c} = prior_v + sum_{c'!=c} m_{c'->v}, hard decision on the posterior
 * after every iteration, stop at the first iteration whose decision reproduces the syndrome
 * (or after max_iter; the last decision is returned and flagged as not converged).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Tanner graph in CSR by check (detector): chk_indptr[D+1], chk_var[nnz] (sorted by variable),
 * and the transpose map var_indptr[M+1], var_edge[nnz] (edge ids in the check-major numbering).
 * syndrome: uint8[D] 0/1.  out_err: uint8[M].  Returns iterations used (negative if not converged). */
int orc_bp_minsum(int D, int M, const uint32_t *chk_indptr, const uint32_t *chk_var,
                  const uint32_t *var_indptr, const uint32_t *var_edge, const float *prior_llr,
                  const uint8_t *syndrome, int max_iter, float alpha, uint8_t *out_err) {
    uint32_t nnz = chk_indptr[D];
    float *v2c = malloc(sizeof(float) * nnz), *c2v = malloc(sizeof(float) * nnz);
    for (int c = 0; c < D; ++c)
        for (uint32_t e = chk_indptr[c]; e < chk_indptr[c + 1]; ++e) v2c[e] = prior_llr[chk_var[e]];
    int it, conv = 0;
    for (it = 1; it <= max_iter; ++it) {
        for (int c = 0; c < D; ++c) {
            float m1 = INFINITY, m2 = INFINITY;
            uint32_t arg = 0;
            int sgn = syndrome[c] ? 1 : 0;
            for (uint32_t e = chk_indptr[c]; e < chk_indptr[c + 1]; ++e) {
                float a = fabsf(v2c[e]);
                if (v2c[e] < 0) sgn ^= 1;
                if (a < m1) { m2 = m1; m1 = a; arg = e; }
                else if (a < m2) m2 = a;
            }
            for (uint32_t e = chk_indptr[c]; e < chk_indptr[c + 1]; ++e) {
                float mag = (e == arg) ? m2 : m1;
                int s = sgn ^ (v2c[e] < 0 ? 1 : 0);
                c2v[e] = alpha * (s ? -mag : mag);
            }
        }
        for (int v = 0; v < M; ++v) {
            float tot = prior_llr[v];
            for (uint32_t k = var_indptr[v]; k < var_indptr[v + 1]; ++k) tot += c2v[var_edge[k]];
            out_err[v] = tot < 0;
            for (uint32_t k = var_indptr[v]; k < var_indptr[v + 1]; ++k) v2c[var_edge[k]] = tot - c2v[var_edge[k]];
        }
        conv = 1;
        for (int c = 0; c < D && conv; ++c) {
            int par = 0;
            for (uint32_t e = chk_indptr[c]; e < chk_indptr[c + 1]; ++e) par ^= out_err[chk_var[e]];
            if (par != syndrome[c]) conv = 0;
        }
        if (conv) break;
    }
    free(v2c); free(c2v);
    return conv ? it : -max_iter;
}
