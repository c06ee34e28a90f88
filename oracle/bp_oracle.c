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
static int bp_minsum_impl(int D, int M, const uint32_t *chk_indptr, const uint32_t *chk_var,
                          const uint32_t *var_indptr, const uint32_t *var_edge, const float *prior_llr,
                          const uint8_t *syndrome, int max_iter, float alpha, uint8_t *out_err, float *out_post) {
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
            if (out_post) out_post[v] = tot;
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

int orc_bp_minsum(int D, int M, const uint32_t *chk_indptr, const uint32_t *chk_var,
                  const uint32_t *var_indptr, const uint32_t *var_edge, const float *prior_llr,
                  const uint8_t *syndrome, int max_iter, float alpha, uint8_t *out_err) {
    return bp_minsum_impl(D, M, chk_indptr, chk_var, var_indptr, var_edge, prior_llr, syndrome, max_iter, alpha, out_err, NULL);
}

/* ---- order-0 ordered statistics decoding (the "OSD" of the reference's BP+OSD decoder, stimbposd / ldpc;
 * published algorithm: Panteleev & Kalachev 2021, Roffe et al. 2020 -- osd_order = 0).
 * Variables sorted by posterior LLR ascending (most likely in error first; ties by index).  Walking that
 * order, a column of H joins the basis iff it is linearly independent of the columns already in it; the
 * answer is the unique combination of basis columns that equals the syndrome.  Because the representation
 * in a basis is unique, the walk may stop at the first moment the syndrome lies in the span of the basis
 * built so far: the later basis columns would get coefficient 0.  max_candidates bounds the walk
 * (0 = all M); max_basis bounds the basis size (0 = D).  max_candidates < 0 is the CUDA kernel's schedule:
 * the first |max_candidates| variables of the reliability order; then, only if they did not suffice, the
 * variables of the detector that is the lowest bit left of the syndrome, in index order, moving on to the
 * next such detector whenever that bit changes; and if a detector's variables are used up with its bit still
 * there, EVERY variable in index order (so a syndrome in the column space of H is always reached).
 * Returns 0 and the error vector on success, 1 if the syndrome was not reached within the bounds. */
typedef struct { float key; uint32_t idx; } osd_item;
static int osd_cmp(const void *a, const void *b) {
    const osd_item *x = a, *y = b;
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

int orc_osd0(int D, int M, const uint32_t *chk_indptr, const uint32_t *chk_var, const uint32_t *var_indptr,
             const uint32_t *var_edge, const float *posterior, const uint8_t *syndrome, int max_candidates,
             int max_basis, uint8_t *out_err, int *basis_size, int *candidates_used) {
    const int W = (D + 63) / 64;
    int tail = 0;
    if (max_candidates < 0) { tail = M; max_candidates = -max_candidates; }
    if (max_candidates <= 0 || max_candidates > M) max_candidates = M;
    if (max_basis <= 0 || max_basis > D) max_basis = D;
    const int CW = (max_basis + 63) / 64;
    /* edge id -> check */
    uint32_t nnz = chk_indptr[D];
    uint32_t *edge_chk = malloc(sizeof(uint32_t) * (nnz ? nnz : 1));
    for (int c = 0; c < D; ++c)
        for (uint32_t e = chk_indptr[c]; e < chk_indptr[c + 1]; ++e) edge_chk[e] = (uint32_t)c;
    osd_item *items = malloc(sizeof(osd_item) * (size_t)M);
    for (int v = 0; v < M; ++v) { items[v].key = posterior[v]; items[v].idx = (uint32_t)v; }
    qsort(items, (size_t)M, sizeof(osd_item), osd_cmp);
    uint64_t *basis = calloc((size_t)max_basis * W, 8), *coef = calloc((size_t)max_basis * CW, 8);
    uint32_t *col = malloc(sizeof(uint32_t) * (size_t)max_basis);
    int32_t *piv = malloc(sizeof(int32_t) * (size_t)D);
    for (int r = 0; r < D; ++r) piv[r] = -1;
    uint64_t *s = calloc((size_t)W, 8), *sc = calloc((size_t)CW, 8), *v = calloc((size_t)W, 8), *vc = calloc((size_t)CW, 8);
    for (int r = 0; r < D; ++r) if (syndrome[r]) s[r >> 6] |= 1ull << (r & 63);
    memset(out_err, 0, (size_t)M);
    int nb = 0, used = 0, done = 0;
#define LOWBIT(vec, r)                                   \
    {                                                    \
        r = -1;                                          \
        for (int w_ = 0; w_ < W; ++w_)                   \
            if (vec[w_]) { r = w_ * 64 + __builtin_ctzll(vec[w_]); break; } \
    }
    /* reduce the syndrome against the basis; done when it vanishes */
#define REDUCE_S()                                                          \
    for (;;) {                                                              \
        int r_; LOWBIT(s, r_);                                              \
        if (r_ < 0) { done = 1; break; }                                    \
        if (piv[r_] < 0) break;                                             \
        for (int w_ = 0; w_ < W; ++w_) s[w_] ^= basis[(size_t)piv[r_] * W + w_];    \
        for (int w_ = 0; w_ < CW; ++w_) sc[w_] ^= coef[(size_t)piv[r_] * CW + w_];  \
    }
    REDUCE_S();
    int stage = 0, i0 = 0, i2 = 0, tr = -1;
    uint32_t tk = 0, tend = 0;
    while (!done && nb < max_basis) {
        uint32_t c;
        if (stage == 0) {                           /* the reliability order */
            if (i0 >= max_candidates) { if (!tail) break; stage = 1; continue; }
            c = items[i0++].idx;
        } else if (stage == 1) {                    /* the variables of the lowest detector left of the syndrome */
            int rs; LOWBIT(s, rs);
            if (rs != tr) { tr = rs; tk = chk_indptr[tr]; tend = chk_indptr[tr + 1]; }
            if (tk >= tend) { stage = 2; continue; }
            c = chk_var[tk++];
        } else {                                    /* every variable */
            if (i2 >= M) break;
            c = (uint32_t)i2++;
        }
        ++used;
        memset(v, 0, (size_t)W * 8); memset(vc, 0, (size_t)CW * 8);
        for (uint32_t k = var_indptr[c]; k < var_indptr[c + 1]; ++k) { uint32_t r = edge_chk[var_edge[k]]; v[r >> 6] ^= 1ull << (r & 63); }
        for (;;) {
            int r; LOWBIT(v, r);
            if (r < 0) break;                       /* dependent on the basis: skipped */
            if (piv[r] >= 0) {
                for (int w = 0; w < W; ++w) v[w] ^= basis[(size_t)piv[r] * W + w];
                for (int w = 0; w < CW; ++w) vc[w] ^= coef[(size_t)piv[r] * CW + w];
            } else {
                piv[r] = nb;
                memcpy(basis + (size_t)nb * W, v, (size_t)W * 8);
                vc[nb >> 6] |= 1ull << (nb & 63);
                memcpy(coef + (size_t)nb * CW, vc, (size_t)CW * 8);
                col[nb] = c;
                ++nb;
                REDUCE_S();
                break;
            }
        }
    }
    if (done)
        for (int b = 0; b < nb; ++b) if ((sc[b >> 6] >> (b & 63)) & 1) out_err[col[b]] = 1;
    if (basis_size) *basis_size = nb;
    if (candidates_used) *candidates_used = used;
    free(edge_chk); free(items); free(basis); free(coef); free(col); free(piv); free(s); free(sc); free(v); free(vc);
    return done ? 0 : 1;
}

/* BP, then OSD-0 on the shots BP did not converge on.  Returns the BP iteration count (negative if BP did
 * not converge); *osd_status = -1 OSD not needed, 0 OSD solved, 1 OSD gave up (out_err = BP's last decision). */
int orc_bp_osd0(int D, int M, const uint32_t *chk_indptr, const uint32_t *chk_var, const uint32_t *var_indptr,
                const uint32_t *var_edge, const float *prior_llr, const uint8_t *syndrome, int max_iter, float alpha,
                int max_candidates, int max_basis, uint8_t *out_err, int *osd_status, int *basis_size, int *candidates_used) {
    float *post = malloc(sizeof(float) * (size_t)M);
    int it = bp_minsum_impl(D, M, chk_indptr, chk_var, var_indptr, var_edge, prior_llr, syndrome, max_iter, alpha, out_err, post);
    *osd_status = -1;
    if (basis_size) *basis_size = 0;
    if (candidates_used) *candidates_used = 0;
    if (it < 0) {
        uint8_t *e2 = malloc((size_t)M);
        int rc = orc_osd0(D, M, chk_indptr, chk_var, var_indptr, var_edge, post, syndrome, max_candidates, max_basis, e2, basis_size, candidates_used);
        *osd_status = rc;
        if (rc == 0) memcpy(out_err, e2, (size_t)M);
        free(e2);
    }
    free(post);
    return it;
}
