// K3b: batched min-sum belief propagation on the DEM Tanner graph, one shot per thread block.
//
// Checks = detectors (D), variables = error mechanisms (M), edges = nnz(H).  Stands in for stimbposd's BP+OSD
// (reference threshold_lab.py:19,30): min-sum BP as BASELINE.json's north star asks (stimbposd
// defaults to product-sum), then -- with params.bp_osd -- order-0 OSD (below; stimbposd defaults to OSD-CS(60)).
// Flooding schedule, fp32 messages:
//   check c -> variable v :  alpha * (-1)^{s_c} * prod_{v' != v} sign(m_{v'->c}) * min_{v' != v} |m_{v'->c}|
//   variable v posterior  :  prior_v + sum_c m_{c->v}      (m_{v->c} = posterior - m_{c->v}, formed on the fly)
// The block keeps one message per edge (check -> variable) and one posterior per variable; they
// live in shared memory when the model fits (Steane, small surface codes) and in a per-block
// HBM/L2 scratch otherwise (rotated d=11: 0.4 MB, BB-144: 1.8 MB per shot).
// Check update: one warp per check, lanes stride over the check's CSR row; min1 / min2 / argmin /
// sign parity are combined with warp shuffles, and the same sweep evaluates the parity of the
// current hard decision, so convergence (H.e == s) costs no extra pass.
// Summation order and tie rules equal oracle/bp_oracle.c, so results match it bit for bit.
//
// OSD-0 (params.bp_osd; the "OSD" of the reference's decoder="bposd") for the shots BP
// leaves unconverged, in two steps:
//  * bp_kernel's epilogue picks the osd_k variables with the smallest posterior LLR, in order (ties by
//    index): 4-pass radix select of the k-th key over the block's posterior array, index-ordered
//    compaction of everything below it, bitonic sort of the k (key, index) pairs in shared memory;
//  * osd_solve_kernel, one warp per shot, walks those candidates: a column of H joins the basis iff it is
//    independent of the columns already in it (incremental GF(2) elimination on D-bit vectors spread over
//    the lanes, pivot = lowest set bit, found with a ballot), and the walk stops the moment the syndrome
//    reduces to zero against the basis -- its representation in a basis is unique, so the later basis
//    columns of a full OSD-0 would get coefficient 0.  The basis may grow to D columns (rank(H) <= D), and a
//    shot whose syndrome the osd_k ordered candidates do not span goes on through every variable in index
//    order, so every syndrome in the column space of H gets a correction; the (impossible for a DEM's own
//    shots) rest is counted and reported as sinter's "discards".  Identical to oracle/bp_oracle.c::orc_osd0
//    with max_candidates = -osd_k.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "gq_common.h"

struct BpArgs {
    const uint32_t *dets;
    uint32_t *pred;
    int32_t *iters_out;
    uint8_t *flags_out;
    uint32_t *corr_out;
    uint64_t nshots;
    uint32_t D, M, DW, OW, MW, nnz;
    const uint32_t *chk_ptr, *chk_var, *var_ptr, *var_edge, *var_obs;
    const float *prior;
    float *scratch;          // per block: nnz + M floats (unused when the model fits in shared)
    size_t scratch_stride;   // floats
    int max_iter;
    float alpha;
    int use_shared;
    // OSD-0 hand-off (osd_k = 0: off)
    uint32_t osd_k;
    uint32_t *cand;          // [capacity][osd_k] candidate variables of every unconverged shot, most likely first
    uint32_t *fail_list;     // [capacity] the unconverged shots
    uint32_t *fail_count;
};


static constexpr int BP_BATCH = 8;        // shots per block of the global-scratch (shot-interleaved) BP kernel
static constexpr int OSD_NCAND = 2048;    // reliability-ordered candidates kept per shot (BB-144 at p = 3e-3: up to ~1400 are usually walked,
                                          // basis <= ~650); a shot that needs more goes on through all variables in index order
// the basis can grow to rank(H) <= D columns: coefficient sets are WPL words per lane like the vectors themselves
// (32 * 32 * WPL >= D ordinals)

__device__ __forceinline__ uint32_t osd_key(float x) {   // order-preserving map float -> uint32
    const uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// The k variables with the smallest posterior, ascending (ties by index), written to out[0..k); k <= OSD_NCAND.
// All 256 threads of the block call it; tot[] holds the block's M posteriors.
__device__ void osd_select(const float *tot, uint32_t stride, uint32_t M, uint32_t k, uint32_t *out) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_rank;
    __shared__ uint32_t cnt_less[256], cnt_tie[256];
    __shared__ unsigned long long pairs[OSD_NCAND];
    __shared__ uint32_t s_fill;
    const int tid = threadIdx.x;
    if (k > M) k = M;
    // ---- k-th smallest key by 8-bit digits, most significant first
    uint32_t prefix = 0, mask = 0, rank = k;        // rank: 1-based rank still to find inside the prefix group
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[tid] = 0;
        __syncthreads();
        for (uint32_t v = tid; v < M; v += 256) {
            const uint32_t key = osd_key(tot[(size_t)v * stride]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t cum = 0, d = 0;
            for (; d < 256; ++d) { if (cum + hist[d] >= rank) break; cum += hist[d]; }
            s_prefix = prefix | (d << shift);
            s_rank = rank - cum;
        }
        __syncthreads();
        prefix = s_prefix; rank = s_rank; mask |= 255u << shift;
        __syncthreads();
    }
    const uint32_t kth = prefix, need_ties = rank;   // keys < kth all belong; of the keys == kth the first need_ties by index
    // ---- index-ordered compaction: thread t owns the contiguous chunk [t * C, (t + 1) * C)
    const uint32_t C = (M + 255) / 256, lo = tid * C, hi = min(M, lo + C);
    uint32_t nl = 0, nt = 0;
    for (uint32_t v = lo; v < hi; ++v) { const uint32_t key = osd_key(tot[(size_t)v * stride]); nl += key < kth; nt += key == kth; }
    cnt_less[tid] = nl; cnt_tie[tid] = nt;
    if (tid == 0) s_fill = 0;
    for (uint32_t i = tid; i < OSD_NCAND; i += 256) pairs[i] = ~0ull;
    __syncthreads();
    uint32_t tie_before = 0;
    for (int t = 0; t < tid; ++t) tie_before += cnt_tie[t];
    for (uint32_t v = lo; v < hi; ++v) {
        const uint32_t key = osd_key(tot[(size_t)v * stride]);
        bool take = key < kth;
        if (key == kth) { take = tie_before < need_ties; ++tie_before; }
        if (take) pairs[atomicAdd(&s_fill, 1u)] = ((unsigned long long)key << 32) | v;
    }
    __syncthreads();
    // ---- bitonic sort of the (key, index) pairs (padding sorts last)
    for (uint32_t size = 2; size <= OSD_NCAND; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = tid; i < OSD_NCAND / 2; i += 256) {
                const uint32_t a = 2 * i - (i & (stride - 1)), b = a + stride;
                const bool up = (a & size) == 0;
                const unsigned long long x = pairs[a], y = pairs[b];
                if ((x > y) == up) { pairs[a] = y; pairs[b] = x; }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = tid; i < OSD_NCAND; i += 256) out[i] = i < k ? (uint32_t)pairs[i] : 0xFFFFFFFFu;
    __syncthreads();
}

__global__ void __launch_bounds__(256) bp_kernel(BpArgs a) {
    extern __shared__ float smem[];
    __shared__ int s_notconv;
    __shared__ uint32_t s_obs;
    float *c2v, *tot;
    if (a.use_shared) { c2v = smem; tot = smem + a.nnz; }
    else { c2v = a.scratch + (size_t)blockIdx.x * a.scratch_stride; tot = c2v + a.nnz; }
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;

    for (uint64_t shot = blockIdx.x; shot < a.nshots; shot += gridDim.x) {
        const uint32_t *syn = a.dets + shot * a.DW;
        for (uint32_t e = tid; e < a.nnz; e += blockDim.x) c2v[e] = 0.0f;
        for (uint32_t v = tid; v < a.M; v += blockDim.x) tot[v] = a.prior[v];
        __syncthreads();
        int used = 0, converged = 0;
        for (int it = 1; it <= a.max_iter + 1; ++it) {
            if (tid == 0) s_notconv = 0;
            __syncthreads();
            // ---- check pass (+ parity of the current decision)
            for (uint32_t c = wid; c < a.D; c += nw) {
                const uint32_t b = a.chk_ptr[c], e = a.chk_ptr[c + 1];
                const int sc = (syn[c >> 5] >> (c & 31)) & 1;
                float m1 = INFINITY, m2 = INFINITY;
                uint32_t arg = 0xFFFFFFFFu;
                int sg = 0, par = 0;
                // rows of up to 256 edges (8 per lane) keep their variable->check messages in registers between
                // the two sweeps and have all their gathers in flight at once (the kernel is latency-bound: the
                // scratch of the resident blocks is far bigger than L2); longer rows take the plain loops
                const bool cached = e - b <= 256u;
                float mv[8];
                if (cached) {
                    uint32_t var[8];
                    float tt[8], cc[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const uint32_t k = b + lane + 32u * j; var[j] = k < e ? a.chk_var[k] : 0u; }
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const uint32_t k = b + lane + 32u * j; tt[j] = k < e ? tot[var[j]] : 0.0f; cc[j] = k < e ? c2v[k] : 0.0f; }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t k = b + lane + 32u * j;
                        if (k < e) {
                            const float t = tt[j], m = t - cc[j], av = fabsf(m);
                            mv[j] = m;
                            sg ^= (m < 0.0f);
                            par ^= (t < 0.0f);
                            if (av < m1) { m2 = m1; m1 = av; arg = k; }
                            else if (av < m2) m2 = av;
                        }
                    }
                } else {
                    for (uint32_t k = b + lane; k < e; k += 32) {
                        float t = tot[a.chk_var[k]];
                        float m = t - c2v[k];
                        float av = fabsf(m);
                        sg ^= (m < 0.0f);
                        par ^= (t < 0.0f);
                        if (av < m1) { m2 = m1; m1 = av; arg = k; }
                        else if (av < m2) m2 = av;
                    }
                }
                sg = __reduce_xor_sync(0xffffffffu, (unsigned)sg);
                par = __reduce_xor_sync(0xffffffffu, (unsigned)par);
                float g1 = m1;
#pragma unroll
                for (int o = 16; o; o >>= 1) g1 = fminf(g1, __shfl_xor_sync(0xffffffffu, g1, o));
                uint32_t garg = __reduce_min_sync(0xffffffffu, m1 == g1 ? arg : 0xFFFFFFFFu);
                float g2 = (arg == garg) ? m2 : m1;
#pragma unroll
                for (int o = 16; o; o >>= 1) g2 = fminf(g2, __shfl_xor_sync(0xffffffffu, g2, o));
                if (par != sc && lane == 0) s_notconv = 1;
                if (cached) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t k = b + lane + 32u * j;
                        if (k < e) {
                            const float mag = (k == garg) ? g2 : g1;
                            const int sgn = sc ^ sg ^ (mv[j] < 0.0f);
                            c2v[k] = a.alpha * (sgn ? -mag : mag);
                        }
                    }
                } else {
                    for (uint32_t k = b + lane; k < e; k += 32) {
                        float m = tot[a.chk_var[k]] - c2v[k];
                        float mag = (k == garg) ? g2 : g1;
                        int s = sc ^ sg ^ (m < 0.0f);
                        c2v[k] = a.alpha * (s ? -mag : mag);
                    }
                }
            }
            __syncthreads();
            if (it > 1 && s_notconv == 0) { converged = 1; used = it - 1; break; }
            if (it == a.max_iter + 1) { used = a.max_iter; break; }
            // ---- variable pass
            // (two variables per thread at a time, edge ids and messages of four edges of each in flight; every
            // sum keeps the oracle's order)
            for (uint32_t v = tid; v < a.M; v += 2 * blockDim.x) {
                const uint32_t v2 = v + blockDim.x;
                const bool has2 = v2 < a.M;
                float t1 = a.prior[v], t2 = has2 ? a.prior[v2] : 0.0f;
                const uint32_t b1 = a.var_ptr[v], e1 = a.var_ptr[v + 1];
                const uint32_t b2 = has2 ? a.var_ptr[v2] : 0u, e2 = has2 ? a.var_ptr[v2 + 1] : 0u;
                const uint32_t n = max(e1 - b1, e2 - b2);
                for (uint32_t o = 0; o < n; o += 4) {
                    uint32_t ed1[4], ed2[4];
                    float c1[4], c2[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        ed1[j] = b1 + o + j < e1 ? a.var_edge[b1 + o + j] : 0u;
                        ed2[j] = b2 + o + j < e2 ? a.var_edge[b2 + o + j] : 0u;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        c1[j] = b1 + o + j < e1 ? c2v[ed1[j]] : 0.0f;
                        c2[j] = b2 + o + j < e2 ? c2v[ed2[j]] : 0.0f;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (b1 + o + j < e1) t1 += c1[j];
                        if (b2 + o + j < e2) t2 += c2[j];
                    }
                }
                tot[v] = t1;
                if (has2) tot[v2] = t2;
            }
            __syncthreads();
        }
        // ---- hard decision, prediction = L.e
        if (tid == 0) s_obs = 0;
        __syncthreads();
        uint32_t om = 0;
        for (uint32_t v0 = 0; v0 < a.M; v0 += blockDim.x) {
            uint32_t v = v0 + tid;
            int bit = v < a.M && tot[v] < 0.0f;
            if (bit) om ^= a.var_obs[v];
            if (a.corr_out) {
                uint32_t word = __ballot_sync(0xffffffffu, bit);
                if (lane == 0 && (v >> 5) < a.MW) a.corr_out[shot * a.MW + (v >> 5)] = word;
            }
        }
        om = __reduce_xor_sync(0xffffffffu, om);
        if (lane == 0 && om) atomicXor(&s_obs, om);
        __syncthreads();
        if (tid == 0) {
            a.pred[shot * a.OW] = s_obs;
            for (uint32_t w = 1; w < a.OW; ++w) a.pred[shot * a.OW + w] = 0;
            if (a.iters_out) a.iters_out[shot] = converged ? used : -a.max_iter;
            if (a.flags_out) a.flags_out[shot] = converged ? 0 : 1;
        }
        __syncthreads();
        if (a.osd_k && !converged) {
            // hand the shot to OSD-0: its osd_k most likely error mechanisms, in order
            __shared__ uint32_t s_slot;
            if (tid == 0) { s_slot = atomicAdd(a.fail_count, 1u); a.fail_list[s_slot] = (uint32_t)shot; }
            __syncthreads();
            osd_select(tot, 1u, a.M, a.osd_k, a.cand + (size_t)s_slot * OSD_NCAND);
        }
    }
}


// ---- big models (messages in global scratch): S shots per block with SHOT-INTERLEAVED messages, c2v[edge][S] and
// tot[variable][S].  The passes are gathers through the Tanner graph's index lists -- posteriors by variable in the
// check pass, messages by edge id in the variable pass -- and a 4-byte gather costs a whole 32-byte sector of HBM / L2
// traffic; with eight shots side by side every gathered sector is eight useful floats (measured on BB-144: the
// one-shot-per-block kernel moved 162 MB per shot at 3.0 TB/s).  The index lists are read once per S shots.
// Arithmetic per shot is unchanged (same order of every sum, same tie rules), so decisions and iteration counts stay
// bit-identical to oracle/bp_oracle.c.  A shot that converges is finalised at once (prediction, correction, iteration
// count); the block goes on until all its shots are done or out of iterations -- at BB-144 error rates most shots that
// converge do so within a few iterations of each other and four in five never do (they go to OSD).
template <int S>
__global__ void __launch_bounds__(256) bp_batched_kernel(BpArgs a) {
    extern __shared__ uint32_t s_syn[];          // [S][DW] the group's syndromes
    __shared__ uint32_t s_notconv, s_obs, s_slot;
    float *c2v = a.scratch + (size_t)blockIdx.x * a.scratch_stride;   // [nnz][S]
    float *tot = c2v + (size_t)a.nnz * S;                             // [M][S]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    const uint64_t ngroups = (a.nshots + S - 1) / S;

    for (uint64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const uint64_t shot0 = grp * S;
        const int ns = (int)min((uint64_t)S, a.nshots - shot0);
        for (uint32_t i = tid; i < (uint32_t)S * a.DW; i += blockDim.x) {
            const uint32_t sh = i / a.DW, w = i - sh * a.DW;
            s_syn[i] = (int)sh < ns ? a.dets[(shot0 + sh) * a.DW + w] : 0u;
        }
        {
            float4 *c4 = reinterpret_cast<float4 *>(c2v);
            const size_t n4 = (size_t)a.nnz * S / 4;
            for (size_t i = tid; i < n4; i += blockDim.x) c4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (uint32_t v = tid; v < a.M; v += blockDim.x) {
            const float pr = a.prior[v];
#pragma unroll
            for (int sh = 0; sh < S; ++sh) tot[(size_t)v * S + sh] = pr;
        }
        __syncthreads();
        uint32_t active = (1u << ns) - 1u;       // shots still iterating (block-uniform)
        uint32_t unconverged = 0;                // shots that ran out of iterations

        // writes the outcome of shot sh from the current posteriors (all threads)
        auto finalise = [&](int sh, bool converged, int used) {
            if (tid == 0) s_obs = 0;
            __syncthreads();
            const uint64_t shot = shot0 + sh;
            uint32_t om = 0;
            for (uint32_t v0 = 0; v0 < a.M; v0 += blockDim.x) {
                const uint32_t v = v0 + tid;
                const int bit = v < a.M && tot[(size_t)v * S + sh] < 0.0f;
                if (bit) om ^= a.var_obs[v];
                if (a.corr_out) {
                    const uint32_t word = __ballot_sync(0xffffffffu, bit);
                    if (lane == 0 && (v >> 5) < a.MW) a.corr_out[shot * a.MW + (v >> 5)] = word;
                }
            }
            om = __reduce_xor_sync(0xffffffffu, om);
            if (lane == 0 && om) atomicXor(&s_obs, om);
            __syncthreads();
            if (tid == 0) {
                a.pred[shot * a.OW] = s_obs;
                for (uint32_t w = 1; w < a.OW; ++w) a.pred[shot * a.OW + w] = 0;
                if (a.iters_out) a.iters_out[shot] = converged ? used : -a.max_iter;
                if (a.flags_out) a.flags_out[shot] = converged ? 0 : 1;
            }
            __syncthreads();
        };

        for (int it = 1; it <= a.max_iter + 1 && active; ++it) {
            if (tid == 0) s_notconv = 0;
            __syncthreads();
            // ---- check pass (+ parity of the current decisions)
            for (uint32_t c = wid; c < a.D; c += nw) {
                const uint32_t b = a.chk_ptr[c], e = a.chk_ptr[c + 1];
                uint32_t scbits = 0;
#pragma unroll
                for (int sh = 0; sh < S; ++sh) scbits |= ((s_syn[sh * a.DW + (c >> 5)] >> (c & 31)) & 1u) << sh;
                float m1[S], m2[S];
                uint32_t arg[S], sg = 0, par = 0;
#pragma unroll
                for (int sh = 0; sh < S; ++sh) { m1[sh] = INFINITY; m2[sh] = INFINITY; arg[sh] = 0xFFFFFFFFu; }
                for (uint32_t k = b + lane; k < e; k += 32) {
                    const uint32_t v = a.chk_var[k];
                    float t[S], cc[S];
#pragma unroll
                    for (int q = 0; q < S / 4; ++q) {
                        const float4 x = reinterpret_cast<const float4 *>(tot + (size_t)v * S)[q];
                        const float4 y = reinterpret_cast<const float4 *>(c2v + (size_t)k * S)[q];
                        t[4 * q] = x.x; t[4 * q + 1] = x.y; t[4 * q + 2] = x.z; t[4 * q + 3] = x.w;
                        cc[4 * q] = y.x; cc[4 * q + 1] = y.y; cc[4 * q + 2] = y.z; cc[4 * q + 3] = y.w;
                    }
#pragma unroll
                    for (int sh = 0; sh < S; ++sh) {
                        const float m = t[sh] - cc[sh], av = fabsf(m);
                        sg ^= (uint32_t)(m < 0.0f) << sh;
                        par ^= (uint32_t)(t[sh] < 0.0f) << sh;
                        if (av < m1[sh]) { m2[sh] = m1[sh]; m1[sh] = av; arg[sh] = k; }
                        else if (av < m2[sh]) m2[sh] = av;
                    }
                }
                sg = __reduce_xor_sync(0xffffffffu, sg);
                par = __reduce_xor_sync(0xffffffffu, par);
                float g1[S], g2[S];
                uint32_t garg[S];
#pragma unroll
                for (int sh = 0; sh < S; ++sh) {
                    // magnitudes are >= 0 (or +inf): their bit patterns order like the floats
                    g1[sh] = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(m1[sh])));
                    garg[sh] = __reduce_min_sync(0xffffffffu, m1[sh] == g1[sh] ? arg[sh] : 0xFFFFFFFFu);
                    const float x = (arg[sh] == garg[sh]) ? m2[sh] : m1[sh];
                    g2[sh] = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(x)));
                }
                if (lane == 0) {
                    const uint32_t bad = (par ^ scbits) & active;
                    if (bad) atomicOr(&s_notconv, bad);
                }
                for (uint32_t k = b + lane; k < e; k += 32) {
                    const uint32_t v = a.chk_var[k];
                    float o[S];
#pragma unroll
                    for (int q = 0; q < S / 4; ++q) {
                        const float4 x = reinterpret_cast<const float4 *>(tot + (size_t)v * S)[q];
                        const float4 y = reinterpret_cast<const float4 *>(c2v + (size_t)k * S)[q];
                        o[4 * q] = x.x - y.x; o[4 * q + 1] = x.y - y.y; o[4 * q + 2] = x.z - y.z; o[4 * q + 3] = x.w - y.w;
                    }
#pragma unroll
                    for (int sh = 0; sh < S; ++sh) {
                        const float mag = (k == garg[sh]) ? g2[sh] : g1[sh];
                        const uint32_t sgn = ((scbits >> sh) ^ (sg >> sh) ^ (uint32_t)(o[sh] < 0.0f)) & 1u;
                        o[sh] = a.alpha * (sgn ? -mag : mag);
                    }
#pragma unroll
                    for (int q = 0; q < S / 4; ++q)
                        reinterpret_cast<float4 *>(c2v + (size_t)k * S)[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                }
            }
            __syncthreads();
            const uint32_t notconv = s_notconv;
            if (it > 1) {
                uint32_t newly = active & ~notconv;   // their decision reproduces the syndrome: converged after it - 1 iterations
                active &= notconv;
                while (newly) {
                    const int sh = __ffs(newly) - 1;
                    newly &= newly - 1;
                    finalise(sh, true, it - 1);
                }
            }
            if (it == a.max_iter + 1) { unconverged = active; break; }
            if (!active) break;
            // ---- variable pass (every sum in the oracle's order)
            for (uint32_t v = tid; v < a.M; v += blockDim.x) {
                float t[S];
                const float pr = a.prior[v];
#pragma unroll
                for (int sh = 0; sh < S; ++sh) t[sh] = pr;
                const uint32_t b1 = a.var_ptr[v], e1 = a.var_ptr[v + 1];
                for (uint32_t k = b1; k < e1; ++k) {
                    const uint32_t ed = a.var_edge[k];
#pragma unroll
                    for (int q = 0; q < S / 4; ++q) {
                        const float4 y = reinterpret_cast<const float4 *>(c2v + (size_t)ed * S)[q];
                        t[4 * q] += y.x; t[4 * q + 1] += y.y; t[4 * q + 2] += y.z; t[4 * q + 3] += y.w;
                    }
                }
#pragma unroll
                for (int q = 0; q < S / 4; ++q)
                    reinterpret_cast<float4 *>(tot + (size_t)v * S)[q] = make_float4(t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3]);
            }
            __syncthreads();
        }
        // ---- shots that ran out of iterations: BP's last decision, and the hand-off to OSD-0
        while (unconverged) {
            const int sh = __ffs(unconverged) - 1;
            unconverged &= unconverged - 1;
            finalise(sh, false, a.max_iter);
            if (a.osd_k) {
                if (tid == 0) { s_slot = atomicAdd(a.fail_count, 1u); a.fail_list[s_slot] = (uint32_t)(shot0 + sh); }
                __syncthreads();
                osd_select(tot + sh, (uint32_t)S, a.M, a.osd_k, a.cand + (size_t)s_slot * OSD_NCAND);
            }
        }
        __syncthreads();
    }
}

// ---- OSD-0 walk, one warp per unconverged shot.  A D-bit vector is spread over the lanes: word w lives in
// slot w / 32 of lane w % 32, so the lowest set bit is one ballot per slot.  Basis vectors and their
// coefficient sets (one bit per basis ordinal, 32 ordinals per lane) live in a per-warp HBM/L2 scratch; every
// lane only ever reads back the words it wrote itself.
struct OsdArgs {
    const uint32_t *dets;
    uint32_t *pred;
    uint8_t *flags_out;
    uint32_t *corr_out;
    const uint32_t *cand, *fail_list;
    uint32_t nfail;
    uint32_t D, DW, OW, MW, M;
    const uint32_t *var_ptr, *var_chk, *var_obs;
    uint32_t *scratch;        // per warp: kmax x (32 WPL basis words + 32 WPL coefficient words) + kmax columns
    size_t scratch_stride;    // words
    uint32_t kmax;            // basis capacity = D (every independent column fits)
    uint32_t *giveups;        // shots whose syndrome is not in the column space of H (cannot happen for a DEM's own shots)
    uint32_t sentinel;        // prediction written for such a shot (0xffffffff: counted as a discard), 0 = leave BP's decision
};

template <int WPL>
__device__ __forceinline__ int osd_lowbit(const uint32_t (&v)[WPL], int lane) {
#pragma unroll
    for (int j = 0; j < WPL; ++j) {
        const unsigned m = __ballot_sync(0xffffffffu, v[j] != 0u);
        if (m) {
            const int l = __ffs(m) - 1;
            const uint32_t w = __shfl_sync(0xffffffffu, v[j], l);
            return ((j * 32 + l) << 5) + __ffs(w) - 1;
        }
    }
    return -1;
}

template <int WPL>
__global__ void __launch_bounds__(128) osd_solve_kernel(OsdArgs a) {
    extern __shared__ uint16_t piv_all[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    uint16_t *piv = piv_all + (size_t)wib * ((a.D + 1) & ~1u);
    uint32_t *basis = a.scratch + (size_t)warp * a.scratch_stride;
    uint32_t *coef = basis + (size_t)a.kmax * 32 * WPL;
    uint32_t *col = coef + (size_t)a.kmax * 32 * WPL;
    for (uint32_t item = warp; item < a.nfail; item += nwarps) {
        const uint32_t shot = a.fail_list[item];
        const uint32_t *cand = a.cand + (size_t)item * OSD_NCAND;
        for (uint32_t r = lane; r < a.D; r += 32) piv[r] = 0xFFFFu;
        uint32_t s[WPL], sc[WPL];
#pragma unroll
        for (int j = 0; j < WPL; ++j) {
            const uint32_t w = j * 32 + lane;
            uint32_t word = w < a.DW ? a.dets[(size_t)shot * a.DW + w] : 0u;
            if (w == a.DW - 1 && (a.D & 31)) word &= (1u << (a.D & 31)) - 1u;
            s[j] = word;
            sc[j] = 0u;
        }
        __syncwarp();
        int nb = 0;
        bool done = false;
        // reduce the syndrome against the basis; done when nothing is left of it
        auto reduce_s = [&]() {
            for (;;) {
                const int r = osd_lowbit<WPL>(s, lane);
                if (r < 0) { done = true; return; }
                const uint32_t p = piv[r];
                if (p == 0xFFFFu) return;
#pragma unroll
                for (int j = 0; j < WPL; ++j) { s[j] ^= basis[((size_t)p * WPL + j) * 32 + lane]; sc[j] ^= coef[((size_t)p * WPL + j) * 32 + lane]; }
            }
        };
        reduce_s();
        // the reliability-ordered candidates first, then (rare) every variable in index order: the walk ends when the
        // syndrome is reduced or the basis is complete
        for (uint32_t i = 0; i < OSD_NCAND + a.M && !done && nb < (int)a.kmax; ++i) {
            const uint32_t c = i < OSD_NCAND ? cand[i] : i - OSD_NCAND;
            if (c == 0xFFFFFFFFu) continue;
            uint32_t v[WPL], vc[WPL];
#pragma unroll
            for (int j = 0; j < WPL; ++j) { v[j] = 0u; vc[j] = 0u; }
            for (uint32_t k = a.var_ptr[c]; k < a.var_ptr[c + 1]; ++k) {
                const uint32_t d = a.var_chk[k], w = d >> 5;
#pragma unroll
                for (int j = 0; j < WPL; ++j)
                    if ((int)(w >> 5) == j && (int)(w & 31u) == lane) v[j] ^= 1u << (d & 31u);
            }
            for (;;) {
                const int r = osd_lowbit<WPL>(v, lane);
                if (r < 0) break;                          // dependent on the basis: skipped
                const uint32_t p = piv[r];
                if (p != 0xFFFFu) {
#pragma unroll
                    for (int j = 0; j < WPL; ++j) { v[j] ^= basis[((size_t)p * WPL + j) * 32 + lane]; vc[j] ^= coef[((size_t)p * WPL + j) * 32 + lane]; }
                } else {
#pragma unroll
                    for (int j = 0; j < WPL; ++j) {
                        if ((nb >> 10) == j && lane == ((nb >> 5) & 31)) vc[j] |= 1u << (nb & 31);
                        basis[((size_t)nb * WPL + j) * 32 + lane] = v[j];
                        coef[((size_t)nb * WPL + j) * 32 + lane] = vc[j];
                    }
                    if (lane == 0) { piv[r] = (uint16_t)nb; col[nb] = c; }
                    ++nb;
                    __syncwarp();
                    reduce_s();
                    break;
                }
            }
        }
        if (done) {
            // the error vector = the basis columns whose coefficient bit is set in sc
            uint32_t *corr_row = a.corr_out ? a.corr_out + (size_t)shot * a.MW : nullptr;
            if (corr_row) for (uint32_t w = lane; w < a.MW; w += 32) corr_row[w] = 0u;
            __syncwarp();
            uint32_t om = 0;
#pragma unroll
            for (int j = 0; j < WPL; ++j) {
                uint32_t bits = sc[j];
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const uint32_t c = col[((j * 32 + lane) << 5) + b];
                    om ^= a.var_obs[c];
                    if (corr_row) atomicXor(&corr_row[c >> 5], 1u << (c & 31u));
                }
            }
            om = __reduce_xor_sync(0xffffffffu, om);
            if (lane == 0) {
                a.pred[(size_t)shot * a.OW] = om;
                if (a.flags_out) a.flags_out[shot] = 0;
            }
        } else if (lane == 0) {
            atomicAdd(a.giveups, 1u);
            if (a.sentinel) a.pred[(size_t)shot * a.OW] = a.sentinel;
        }
        __syncwarp();
    }
}

static int osd_wpl(uint32_t DW) { return DW <= 32 ? 1 : (DW <= 64 ? 2 : (DW <= 128 ? 4 : (DW <= 256 ? 8 : 0))); }

int gq_bp_create(gq_dec *dec) {
    gq_dem *dem = dec->dem;
    const uint32_t D = dem->D, M = dem->M;
    if (M == 0 || D == 0) GQ_FAIL(GQ_E_INVALID, "BP decoder: empty DEM");
    const uint32_t nnz = (uint32_t)dem->det_idx.size();
    // check-major CSR (variables ascending inside a row), and for every variable the ids of its edges
    std::vector<uint32_t> chk_ptr(D + 1, 0), chk_var(nnz), var_ptr(M + 1, 0), var_edge(nnz), var_obs(M, 0);
    for (uint32_t d : dem->det_idx) chk_ptr[d + 1]++;
    for (uint32_t c = 0; c < D; ++c) chk_ptr[c + 1] += chk_ptr[c];
    {
        std::vector<uint32_t> fill(chk_ptr.begin(), chk_ptr.end() - 1);
        std::vector<uint32_t> vfill(M, 0);
        for (uint32_t m = 0; m < M; ++m) var_ptr[m + 1] = var_ptr[m] + (dem->det_indptr[m + 1] - dem->det_indptr[m]);
        // mechanisms are visited in ascending order, so each check row ends up sorted by variable;
        // a variable's edge list is sorted by edge id after the sort below
        for (uint32_t m = 0; m < M; ++m)
            for (uint32_t q = dem->det_indptr[m]; q < dem->det_indptr[m + 1]; ++q) {
                uint32_t c = dem->det_idx[q];
                uint32_t e = fill[c]++;
                chk_var[e] = m;
                var_edge[var_ptr[m] + vfill[m]++] = e;
            }
        for (uint32_t m = 0; m < M; ++m) std::sort(var_edge.begin() + var_ptr[m], var_edge.begin() + var_ptr[m + 1]);
    }
    // detectors of every variable (CSR by variable, ascending): the columns of H for OSD
    std::vector<uint32_t> var_chk(dem->det_idx.begin(), dem->det_idx.end());
    for (uint32_t m = 0; m < M; ++m) std::sort(var_chk.begin() + var_ptr[m], var_chk.begin() + var_ptr[m + 1]);
    std::vector<float> prior(M);
    for (uint32_t m = 0; m < M; ++m) {
        double p = std::min(std::max(dem->p[m], 1e-30), 1 - 1e-7);
        prior[m] = (float)log((1 - p) / p);
        for (uint32_t q = dem->obs_indptr[m]; q < dem->obs_indptr[m + 1]; ++q) var_obs[m] ^= 1u << dem->obs_idx[q];
    }
    dec->bp_nnz = nnz;
    auto up = [&](uint32_t **dst, const std::vector<uint32_t> &v) -> int {
        GQ_CUDA(cudaMalloc((void **)dst, std::max<size_t>(v.size(), 1) * 4));
        GQ_CUDA(cudaMemcpy(*dst, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
        return GQ_OK;
    };
    int rc;
    if ((rc = up(&dec->d_chk_ptr, chk_ptr)) || (rc = up(&dec->d_chk_var, chk_var)) ||
        (rc = up(&dec->d_var_ptr, var_ptr)) || (rc = up(&dec->d_var_edge, var_edge)) ||
        (rc = up(&dec->d_var_obs, var_obs)) || (rc = up(&dec->d_var_chk, var_chk)))
        return rc;
    if (dec->params.bp_osd) {
        if (osd_wpl(dem->DW) == 0) GQ_FAIL(GQ_E_UNSUPPORTED, "BP+OSD: more than 8192 detectors are not supported");
        const int osmem = 4 * (int)((D + 1) & ~1u) * 2;
        cudaFuncSetAttribute(osd_solve_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, osmem);
        cudaFuncSetAttribute(osd_solve_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, osmem);
        cudaFuncSetAttribute(osd_solve_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, osmem);
        cudaFuncSetAttribute(osd_solve_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, osmem);
    }
    GQ_CUDA(cudaMalloc((void **)&dec->d_prior, (size_t)M * 4));
    GQ_CUDA(cudaMemcpy(dec->d_prior, prior.data(), (size_t)M * 4, cudaMemcpyHostToDevice));
    dec->bp_msgs_per_block = ((size_t)nnz + M + 31) & ~(size_t)31;
    size_t bytes = dec->bp_msgs_per_block * 4;
    if (bytes <= 192 * 1024 && !dec->params.bp_scratch) {
        dec->bp_blocks = dem->num_sms * std::max<int>(1, std::min<int>(8, (int)(192 * 1024 / bytes)));
        if (cudaFuncSetAttribute(bp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) GQ_FAIL(GQ_E_CUDA, "BP decoder: cudaFuncSetAttribute failed");   // + 19 KB static (OSD selection)
    } else {
        // shot-interleaved kernel: BP_BATCH shots per block, c2v[nnz][BP_BATCH] then tot[M][BP_BATCH]
        dec->bp_msgs_per_block = (size_t)BP_BATCH * ((size_t)nnz + M) + 32;
        bytes = dec->bp_msgs_per_block * 4;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bp_batched_kernel<BP_BATCH>, 256, (size_t)BP_BATCH * dem->DW * 4) != cudaSuccess || per_sm < 1) per_sm = 2;
        dec->bp_blocks = dem->num_sms * per_sm;
        GQ_CUDA(cudaMalloc((void **)&dec->d_bp_msgs, bytes * dec->bp_blocks));
    }
    return GQ_OK;
}

void gq_bp_destroy(gq_dec *dec) {
    cudaFree(dec->d_chk_ptr); cudaFree(dec->d_chk_var); cudaFree(dec->d_var_ptr);
    cudaFree(dec->d_var_edge); cudaFree(dec->d_var_obs); cudaFree(dec->d_prior); cudaFree(dec->d_bp_msgs);
    cudaFree(dec->d_var_chk); cudaFree(dec->d_osd_cand); cudaFree(dec->d_osd_fail); cudaFree(dec->d_osd_scratch);
}

int gq_bp_decode(gq_dec *dec, const uint32_t *dets_sm, uint64_t nshots, uint32_t *pred_sm,
                 int32_t *weight_out, uint8_t *flags_out, uint32_t *corr_out, int *launches) {
    gq_dem *dem = dec->dem;
    BpArgs a;
    memset(&a, 0, sizeof(a));
    a.D = dem->D; a.M = dem->M; a.DW = dem->DW; a.OW = dem->OW; a.MW = (dem->M + 31) / 32; a.nnz = dec->bp_nnz;
    a.chk_ptr = dec->d_chk_ptr; a.chk_var = dec->d_chk_var; a.var_ptr = dec->d_var_ptr;
    a.var_edge = dec->d_var_edge; a.var_obs = dec->d_var_obs; a.prior = dec->d_prior;
    a.scratch = dec->d_bp_msgs; a.scratch_stride = dec->bp_msgs_per_block;
    a.max_iter = dec->params.bp_max_iter; a.alpha = dec->params.bp_alpha;
    a.use_shared = dec->d_bp_msgs == nullptr;
    size_t smem = a.use_shared ? dec->bp_msgs_per_block * 4 : 0;
    const bool osd = dec->params.bp_osd != 0;
    const uint64_t chunk = osd ? 16384 : nshots;       // OSD keeps 8 KB of candidates per unconverged shot
    const int wpl = osd_wpl(dem->DW);
    const int osd_blocks = dem->num_sms * 2, osd_warps = osd_blocks * 4;
    const uint32_t osd_kmax = dem->D;
    const size_t osd_stride = (size_t)osd_kmax * (64 * (size_t)wpl + 1);
    if (osd) {
        if (!dec->d_osd_cand) {
            GQ_CUDA(cudaMalloc((void **)&dec->d_osd_cand, chunk * OSD_NCAND * 4));
            GQ_CUDA(cudaMalloc((void **)&dec->d_osd_fail, (chunk + 2) * 4));   // [0] unconverged count, [1] OSD give-ups, [2..] shots
            GQ_CUDA(cudaMalloc((void **)&dec->d_osd_scratch, (size_t)osd_warps * osd_stride * 4));
        }
        a.osd_k = OSD_NCAND; a.cand = dec->d_osd_cand; a.fail_list = dec->d_osd_fail + 2; a.fail_count = dec->d_osd_fail;
    }
    for (uint64_t s0 = 0; s0 < nshots; s0 += chunk) {
        const uint64_t ns = std::min(chunk, nshots - s0);
        a.dets = dets_sm + s0 * dem->DW; a.pred = pred_sm + s0 * dem->OW;
        a.iters_out = weight_out ? weight_out + s0 : nullptr;
        a.flags_out = flags_out ? flags_out + s0 : nullptr;
        a.corr_out = corr_out ? corr_out + s0 * a.MW : nullptr;
        a.nshots = ns;
        if (osd) GQ_CUDA(cudaMemsetAsync(dec->d_osd_fail, 0, 8, dem->stream));
        if (a.use_shared) {
            int blocks = (int)std::min<uint64_t>(ns, (uint64_t)dec->bp_blocks);
            bp_kernel<<<blocks, 256, smem, dem->stream>>>(a);
        } else {
            int blocks = (int)std::min<uint64_t>((ns + BP_BATCH - 1) / BP_BATCH, (uint64_t)dec->bp_blocks);
            bp_batched_kernel<BP_BATCH><<<blocks, 256, (size_t)BP_BATCH * dem->DW * 4, dem->stream>>>(a);
        }
        GQ_CUDA(cudaGetLastError());
        ++*launches;
        if (!osd) continue;
        uint32_t nfail = 0;
        GQ_CUDA(cudaMemcpyAsync(&nfail, dec->d_osd_fail, 4, cudaMemcpyDeviceToHost, dem->stream));
        GQ_CUDA(cudaStreamSynchronize(dem->stream));
        dec->osd_shots += nfail;
        if (nfail == 0) continue;
        OsdArgs o;
        memset(&o, 0, sizeof(o));
        o.dets = a.dets; o.pred = a.pred; o.flags_out = a.flags_out; o.corr_out = a.corr_out;
        o.cand = dec->d_osd_cand; o.fail_list = dec->d_osd_fail + 2; o.nfail = nfail;
        o.kmax = osd_kmax; o.giveups = dec->d_osd_fail + 1;
        o.sentinel = dem->O < 32 ? 0xFFFFFFFFu : 0u;
        o.D = dem->D; o.DW = dem->DW; o.OW = dem->OW; o.MW = a.MW; o.M = dem->M;
        o.var_ptr = dec->d_var_ptr; o.var_chk = dec->d_var_chk; o.var_obs = dec->d_var_obs;
        o.scratch = dec->d_osd_scratch; o.scratch_stride = osd_stride;
        const int ob = (int)std::min<uint32_t>((nfail + 3) / 4, (uint32_t)osd_blocks);
        const size_t osmem = 4 * (size_t)((dem->D + 1) & ~1u) * 2;
        switch (wpl) {
        case 1: osd_solve_kernel<1><<<ob, 128, osmem, dem->stream>>>(o); break;
        case 2: osd_solve_kernel<2><<<ob, 128, osmem, dem->stream>>>(o); break;
        case 4: osd_solve_kernel<4><<<ob, 128, osmem, dem->stream>>>(o); break;
        default: osd_solve_kernel<8><<<ob, 128, osmem, dem->stream>>>(o); break;
        }
        GQ_CUDA(cudaGetLastError());
        ++*launches;
        uint32_t ngive = 0;
        GQ_CUDA(cudaMemcpyAsync(&ngive, dec->d_osd_fail + 1, 4, cudaMemcpyDeviceToHost, dem->stream));
        GQ_CUDA(cudaStreamSynchronize(dem->stream));
        dec->osd_giveups += ngive;
    }
    return GQ_OK;
}
