// K3b: batched min-sum belief propagation on the DEM Tanner graph, one shot (or eight) per thread block.
//
// Checks = detectors (D), variables = error mechanisms (M), edges = nnz(H).  Stands in for stimbposd's BP+OSD
// (reference threshold_lab.py:19,30): min-sum BP as BASELINE.json's north star asks (stimbposd
// defaults to product-sum), then -- with params.bp_osd -- order-0 OSD (below; stimbposd defaults to OSD-CS(60)).
// Flooding schedule, fp32 messages:
//   check c -> variable v :  alpha * (-1)^{s_c} * prod_{v' != v} sign(m_{v'->c}) * min_{v' != v} |m_{v'->c}|
//   variable v posterior  :  prior_v + sum_c m_{c->v}      (m_{v->c} = posterior - m_{c->v}, formed on the fly)
// Three kernels, one arithmetic (summation order and tie rules equal oracle/bp_oracle.c: results match it bit for bit):
//  * bp_kernel: the model's messages fit shared memory (Steane, small surface codes): one message per edge + one
//    posterior per variable in shared memory, check update by one warp per check (min1 / min2 / argmin / sign parity by
//    warp shuffles; the same sweep evaluates the parity of the hard decision, so convergence costs no extra pass);
//  * bp_cms_kernel ("compressed min-sum", the default for everything bigger: rotated d = 11, BB-144): NO message in
//    memory at all -- per check {min1, argmin, min2, sign} and one sign bit per edge, all in shared memory, passes turned
//    inside out (variables offer their messages to the checks' next records with shared-memory atomics); see there;
//  * bp_batched_kernel<8>: messages in a global scratch, eight shots interleaved per block (models whose records do not
//    fit shared memory or with more than 16 detectors per mechanism; params.bp_scratch = 1 forces it).
//
// OSD-0 (params.bp_osd; the "OSD" of the reference's decoder="bposd") for the shots BP
// leaves unconverged, in two steps:
//  * the BP kernel's epilogue picks the osd_k variables with the smallest posterior LLR, in order (ties by
//    index): 4-pass radix select of the k-th key over the block's posterior array, index-ordered
//    compaction of everything below it, bitonic sort of the k (key, index) pairs in shared memory;
//  * osd_rref_kernel, one warp per shot, walks those candidates: a column of H joins the basis iff it is
//    independent of the columns already in it (incremental GF(2) elimination on D-bit vectors spread over
//    the lanes, basis kept fully reduced so that a column is reduced by independent loads instead of a dependent chain),
//    and the walk stops the moment the syndrome
//    reduces to zero against the basis -- its representation in a basis is unique, so the later basis
//    columns of a full OSD-0 would get coefficient 0.  The basis may grow to D columns (rank(H) <= D), and a
//    shot whose syndrome the osd_k ordered candidates do not span goes on through the variables of the detectors left
//    of its syndrome (then, if need be, through every variable in index
//    order), so every syndrome in the column space of H gets a correction; the (impossible for a DEM's own
//    shots) rest is counted and reported as sinter's "discards".  Identical to oracle/bp_oracle.c::orc_osd0
//    with max_candidates = -osd_k.
#include <math.h>
#include <stdlib.h>
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
// All threads of the block (>= 256) call it; tot[] holds the block's M posteriors.
// ws: OSD_WS_WORDS 8-byte words of shared memory (the callers that have nothing to spare declare it statically).
static constexpr int OSD_WS_WORDS = OSD_NCAND + 3 * 128 + 2;
__device__ void osd_select(const float *tot, uint32_t stride, uint32_t M, uint32_t k, uint32_t *out, unsigned long long *ws) {
    unsigned long long *pairs = ws;
    uint32_t *hist = reinterpret_cast<uint32_t *>(ws + OSD_NCAND), *cnt_less = hist + 256, *cnt_tie = hist + 512;
    uint32_t &s_prefix = hist[768], &s_rank = hist[769], &s_fill = hist[770];
    const int tid = threadIdx.x, nt = blockDim.x;     // nt >= 256; the chunked compaction is done by the first 256 threads
    if (k > M) k = M;
    // ---- k-th smallest key by 8-bit digits, most significant first
    uint32_t prefix = 0, mask = 0, rank = k;        // rank: 1-based rank still to find inside the prefix group
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        for (uint32_t v = tid; v < M; v += nt) {
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
    // ---- index-ordered compaction: thread t < 256 owns the contiguous chunk [t * C, (t + 1) * C)
    const uint32_t C = (M + 255) / 256, lo = min(M, tid * C), hi = tid < 256 ? min(M, lo + C) : lo;
    uint32_t nl = 0, nt_ = 0;
    for (uint32_t v = lo; v < hi; ++v) { const uint32_t key = osd_key(tot[(size_t)v * stride]); nl += key < kth; nt_ += key == kth; }
    if (tid < 256) { cnt_less[tid] = nl; cnt_tie[tid] = nt_; }
    if (tid == 0) s_fill = 0;
    for (uint32_t i = tid; i < OSD_NCAND; i += nt) pairs[i] = ~0ull;
    __syncthreads();
    uint32_t tie_before = 0;
    for (int t = 0; t < tid && t < 256; ++t) tie_before += cnt_tie[t];
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
            for (uint32_t i = tid; i < OSD_NCAND / 2; i += nt) {
                const uint32_t a = 2 * i - (i & (stride - 1)), b = a + stride;
                const bool up = (a & size) == 0;
                const unsigned long long x = pairs[a], y = pairs[b];
                if ((x > y) == up) { pairs[a] = y; pairs[b] = x; }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = tid; i < OSD_NCAND; i += nt) out[i] = i < k ? (uint32_t)pairs[i] : 0xFFFFFFFFu;
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
            __shared__ unsigned long long osd_ws[OSD_WS_WORDS];
            osd_select(tot, 1u, a.M, a.osd_k, a.cand + (size_t)s_slot * OSD_NCAND, osd_ws);
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
                __shared__ unsigned long long osd_ws[OSD_WS_WORDS];
                osd_select(tot + sh, (uint32_t)S, a.M, a.osd_k, a.cand + (size_t)s_slot * OSD_NCAND, osd_ws);
            }
        }
        __syncthreads();
    }
}

// ---- compressed min-sum ("cms"): the same flooding schedule with NO per-edge message in memory.
// A min-sum check-to-variable message takes one of two magnitudes: c2v(c, v) = alpha * sign * (v == argmin_c ? min2_c : min1_c).
// So a check is fully described by {min1, argmin, min2, sign product} (12 bytes) plus ONE BIT per edge (the sign of the
// variable-to-check message it was computed from), and a posterior is a sum a variable re-forms from the <= 16 records of
// its checks.  One block = one shot, everything in shared memory (BB-144 x 12 rounds: 100 KB, against 1.8 MB of fp32
// messages per shot in the kernels above), and the passes turn inside out: instead of a check gathering the posteriors of
// its ~230 variables (a 4-byte gather per edge from a per-shot array that only HBM / L2 can hold), a VARIABLE reads the
// records of its checks, forms its posterior and its variable-to-check messages in registers, and offers each message to
// the check's next record with shared-memory atomics: min1/argmin = 64-bit atomicMin of (|m| bits << 32 | variable), min2
// = atomicMin of whatever loses or is displaced at min1, sign product and decision parity = atomicXor.  A message at or
// above the record's current min2 cannot change it and is filtered by a plain load first, so only a few of a check's
// ~230 offers reach an atomic.  The only global traffic is the Tanner graph itself (coalesced, L2-resident, shared by all
// blocks).  Same fp32 operations in the same order as oracle/bp_oracle.c (the sum over a variable's checks in ascending
// check order, argmin ties to the lowest variable -- immaterial, min2 == min1 then): bit-identical decisions and
// iteration counts.
// Variables are processed in tiles of 32 with equal degree (sorted by degree at set-up, a permutation: tie rules use the
// original index), so the per-degree bodies are straight-line code over registers.
struct CmsArgs {
    const uint32_t *dets;
    uint32_t *pred;
    int32_t *iters_out;
    uint8_t *flags_out;
    uint32_t *corr_out;
    uint64_t nshots;
    uint32_t D, M, DW, OW, MW;
    uint32_t ntiles, FW, SW;           // tiles of 32 variables; words of packed check flags; words of edge sign bits
    uint32_t grp_tile[18];             // tiles of degree g: [grp_tile[g], grp_tile[g + 1]) (tiles are sorted by degree)
    uint32_t grp_word[17];             // first sign word of degree g's first tile; tile t of the group starts g * (t - first) words later,
                                       // and its checks at 32 * that
    const uint32_t *tile_var;          // [ntiles * 32] original variable index (M: padding lane)
    const float *tile_prior;           // [ntiles * 32]
    const uint32_t *tile_obs;          // [ntiles * 32] observable mask
    const uint16_t *tile_chk;          // [32 * SW] check of edge j of lane l of a tile at 32 * (first word + j) + l
    int max_iter;
    float alpha;
    uint32_t *cursor;                  // next shot to claim
    float *post;                       // per block: M posteriors of an unconverged shot, for the OSD candidate selection
    uint32_t osd_k;
    uint32_t *cand, *fail_list, *fail_count;
};

static constexpr unsigned long long CMS_INFKEY = 0x7f800000ffffffffull;

// Records of check c: key1[buffer][c] = (min1 bits << 32) | argmin variable (bit 31 of the low word, once the record
// is closed: the total sign), key2[buffer][c] = min2 bits (8- and 4-byte strides: an interleaved 16-byte layout doubled the
// shared-memory bank conflicts of the random record reads, measured).  Check D is a dummy whose closed record always reads zero:
// the padding lanes of a tile (variable index M, prior 1e30) point at it, so the tile bodies carry no "valid lane" tests.
template <int DG, bool POST>
__device__ __forceinline__ void cms_group(const CmsArgs &a, int lane, int wid, int nw, const unsigned long long *key1o,
                                          const uint32_t *key2o, unsigned long long *key1n, uint32_t *key2n, uint32_t *flagsn,
                                          uint32_t *sgn, uint32_t *decw, float *post) {
    const uint32_t t0 = a.grp_tile[DG], t1 = a.grp_tile[DG + 1], w0 = a.grp_word[DG];
    const uint32_t lsh = 31u - (uint32_t)lane;
    for (uint32_t t = t0 + wid; t < t1; t += nw) {
        const uint32_t slot = t * 32u + lane, wbase = w0 + (t - t0) * DG;
        const uint32_t v = a.tile_var[slot];
        float tot = a.tile_prior[slot];
        uint32_t c[DG > 0 ? DG : 1];
        float cv[DG > 0 ? DG : 1];
        const uint16_t *cp = a.tile_chk + (size_t)wbase * 32u + lane;
        uint32_t *sp = sgn + wbase;
#pragma unroll
        for (int j = 0; j < DG; ++j) c[j] = cp[j * 32];
#pragma unroll
        for (int j = 0; j < DG; ++j) {
            const unsigned long long k1 = key1o[c[j]];
            const uint32_t lo = (uint32_t)k1;
            uint32_t mag = (uint32_t)(k1 >> 32);
            if ((lo & 0x7fffffffu) == v) mag = key2o[c[j]];
            const uint32_t sg = (lo ^ (sp[j] << lsh)) & 0x80000000u;
            cv[j] = __fmul_rn(a.alpha, __uint_as_float(mag | sg));
        }
#pragma unroll
        for (int j = 0; j < DG; ++j) tot = __fadd_rn(tot, cv[j]);
        const bool bit = tot < 0.0f;
        const uint32_t word = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) decw[t] = word;
        if (POST) post[v] = tot;
        if (word) {                  // rare: some variable of the tile is decided "error" -- its checks' decision parities flip
            if (bit) {
#pragma unroll
                for (int j = 0; j < DG; ++j) atomicXor(&flagsn[c[j] >> 4], 2u << (2u * (c[j] & 15u)));
            }
        }
#pragma unroll
        for (int j = 0; j < DG; ++j) {
            const float m = __fsub_rn(tot, cv[j]);
            const bool neg = m < 0.0f;
            const uint32_t nsg = __ballot_sync(0xffffffffu, neg);
            if (lane == 0) sp[j] = nsg;
            const uint32_t cc = c[j];
            const uint32_t magb = __float_as_uint(m) & 0x7fffffffu;
            const uint32_t cur2 = key2n[cc];
            if (neg || magb < cur2) {            // one divergent region per edge for everything that is rare
                if (neg) atomicXor(&flagsn[cc >> 4], 1u << (2u * (cc & 15u)));
                if (magb < cur2) {
                    const unsigned long long key = ((unsigned long long)magb << 32) | v;
                    const unsigned long long old = atomicMin(&key1n[cc], key);
                    const uint32_t loser = key < old ? (uint32_t)(old >> 32) : magb;
                    if (loser < cur2) atomicMin(&key2n[cc], loser);
                }
            }
        }
    }
}

// one phase: the degree groups in turn (no barrier between them); the bodies are straight-line code per degree, entered
// once per group -- a per-tile switch compiled into a chain of indirect branches that cost more than a small tile's work
template <bool POST>
__device__ __forceinline__ void cms_phase(const CmsArgs &a, int lane, int wid, int nw, const unsigned long long *key1o, const uint32_t *key2o,
                                          unsigned long long *key1n, uint32_t *key2n, uint32_t *flagsn, uint32_t *sgn, uint32_t *decw, float *post) {
#define CMS_GROUP(G) cms_group<G, POST>(a, lane, wid, nw, key1o, key2o, key1n, key2n, flagsn, sgn, decw, post);
    CMS_GROUP(0) CMS_GROUP(1) CMS_GROUP(2) CMS_GROUP(3) CMS_GROUP(4) CMS_GROUP(5) CMS_GROUP(6) CMS_GROUP(7) CMS_GROUP(8)
    CMS_GROUP(9) CMS_GROUP(10) CMS_GROUP(11) CMS_GROUP(12) CMS_GROUP(13) CMS_GROUP(14) CMS_GROUP(15) CMS_GROUP(16)
#undef CMS_GROUP
}

__global__ void __launch_bounds__(1024, 1) bp_cms_kernel(CmsArgs a) {
    extern __shared__ unsigned long long cms_smem[];
    __shared__ uint32_t s_nc[2], s_obs, s_shot, s_slot;
    const uint32_t D1 = a.D + 1;
    unsigned long long *key1 = cms_smem;                               // [2][D + 1]
    uint32_t *key2 = reinterpret_cast<uint32_t *>(key1 + 2 * (size_t)D1);    // [2][D + 1]
    uint32_t *flags = key2 + 2 * (size_t)D1;                           // [2][FW]  per check: bit 0 sign product, bit 1 decision parity
    uint32_t *sgn = flags + 2 * (size_t)a.FW;                          // [SW]
    uint32_t *decw = sgn + a.SW;                                       // [ntiles]
    uint32_t *syn = decw + a.ntiles;                                   // [DW]
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    float *post_blk = a.post ? a.post + (size_t)blockIdx.x * (a.M + 1) : nullptr;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_shot = atomicAdd(a.cursor, 1u);
        __syncthreads();
        const uint64_t shot = s_shot;
        if (shot >= a.nshots) break;
        for (uint32_t i = tid; i < D1; i += nt) { key1[i] = 0ull; key1[D1 + i] = CMS_INFKEY; key2[i] = 0u; key2[D1 + i] = 0x7f800000u; }
        for (uint32_t i = tid; i < 2 * a.FW; i += nt) flags[i] = 0u;
        for (uint32_t i = tid; i < a.SW; i += nt) sgn[i] = 0u;
        for (uint32_t i = tid; i < a.DW; i += nt) syn[i] = a.dets[shot * a.DW + i];
        if (tid == 0) { s_nc[0] = 0; s_nc[1] = 0; s_obs = 0; }
        __syncthreads();
        int converged = 0, used = a.max_iter;
        for (int it = 1; it <= a.max_iter + 1; ++it) {
            const int o = (it + 1) & 1, n = it & 1;          // old / new record buffers
            unsigned long long *key1n = key1 + (size_t)n * D1, *key1x = key1 + (size_t)o * D1;
            uint32_t *key2n = key2 + (size_t)n * D1, *key2x = key2 + (size_t)o * D1, *flagsn = flags + (size_t)n * a.FW;
            if (a.osd_k && it == a.max_iter + 1) cms_phase<true>(a, lane, wid, nw, key1x, key2x, key1n, key2n, flagsn, sgn, decw, post_blk);
            else cms_phase<false>(a, lane, wid, nw, key1x, key2x, key1n, key2n, flagsn, sgn, decw, nullptr);
            __syncthreads();
            // ---- close the new records: decision parity against the syndrome, total sign into the argmin word; the
            // old buffer becomes the next phase's new one
            {
                uint32_t *flagsx = flags + (size_t)o * a.FW;
                uint32_t bad = 0;
                for (uint32_t cc = tid; cc < D1; cc += nt) {
                    if (cc < a.D) {
                        const uint32_t f = (flagsn[cc >> 4] >> (2u * (cc & 15u))) & 3u;
                        const uint32_t sc = (syn[cc >> 5] >> (cc & 31u)) & 1u;
                        bad |= (f >> 1) ^ sc;
                        key1n[cc] |= (unsigned long long)((f ^ sc) & 1u) << 31;
                    } else { key1n[cc] = 0ull; key2n[cc] = 0u; }
                    key1x[cc] = CMS_INFKEY;
                    key2x[cc] = 0x7f800000u;
                }
                if (bad) s_nc[it & 1] = 1u;
                if (tid == 0) s_nc[(it + 1) & 1] = 0u;
                __syncthreads();
                for (uint32_t i = tid; i < a.FW; i += nt) flagsx[i] = 0u;
            }
            const uint32_t notconv = s_nc[it & 1];
            if (it > 1 && notconv == 0) { converged = 1; used = it - 1; break; }
            __syncthreads();
        }
        // ---- outcome: the decision bits of the last phase
        uint32_t om = 0;
        uint32_t *corr_row = a.corr_out ? a.corr_out + shot * a.MW : nullptr;
        if (corr_row) {
            for (uint32_t i = tid; i < a.MW; i += nt) corr_row[i] = 0u;
            __syncthreads();
        }
        for (uint32_t t = wid; t < a.ntiles; t += nw) {
            if ((decw[t] >> lane) & 1u) {
                om ^= a.tile_obs[t * 32u + lane];
                if (corr_row) { const uint32_t v = a.tile_var[t * 32u + lane]; atomicOr(&corr_row[v >> 5], 1u << (v & 31u)); }
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
        if (a.osd_k && !converged) {
            if (tid == 0) { s_slot = atomicAdd(a.fail_count, 1u); a.fail_list[s_slot] = (uint32_t)shot; }
            __syncthreads();
            osd_select(post_blk, 1u, a.M, a.osd_k, a.cand + (size_t)s_slot * OSD_NCAND, cms_smem);   // the records are free now
        }
    }
}

// ---- OSD-0 walk, one warp per unconverged shot.  A D-bit vector is spread over the lanes: word w lives in
// slot w / 32 of lane w % 32, so the lowest set bit is one ballot per slot.  Basis vectors and, per basis vector, the
// set of earlier basis vectors that were folded into it (one bit per basis ordinal, 32 ordinals per lane; written once,
// read once in the final back-substitution) live in a per-warp HBM/L2 scratch; every lane only ever reads back the words
// it wrote itself.
struct OsdArgs {
    const uint32_t *dets;
    uint32_t *pred;
    uint8_t *flags_out;
    uint32_t *corr_out;
    const uint32_t *cand, *fail_list;
    uint32_t nfail;
    uint32_t D, DW, OW, MW, M;
    const uint32_t *var_ptr, *var_chk, *var_obs;
    const uint32_t *chk_ptr, *chk_var;   // rows of H (variables of every detector, ascending): the targeted part of the walk
    uint32_t *cursor;         // next unconverged shot to claim
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
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint16_t *piv = piv_all + (size_t)wib * ((a.D + 1) & ~1u);
    uint32_t *basis = a.scratch + (size_t)warp * a.scratch_stride;
    uint32_t *coef = basis + (size_t)a.kmax * 32 * WPL;
    uint32_t *col = coef + (size_t)a.kmax * 32 * WPL;
    // shots differ by two orders of magnitude in the length of their walk: warps claim the next one with an atomic
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(a.cursor, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= a.nfail) break;
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
        // A reduction step XORs the basis vector whose pivot (lowest set bit) is the vector's lowest set bit r: both are zero
        // below r, so only the words from r / 32 on are loaded.  A vector never meets the same basis vector twice (later
        // steps only touch bits above r), so "which basis vectors went into it" is a plain bit set -- one bit set per
        // step in registers, no second row to load and XOR per step (the kernel is bound by the bytes it moves).
        auto set_bit = [&](uint32_t (&x)[WPL], uint32_t p) {
#pragma unroll
            for (int j = 0; j < WPL; ++j)
                if ((int)(p >> 10) == j && lane == (int)((p >> 5) & 31u)) x[j] |= 1u << (p & 31u);
        };
        // reduce the syndrome against the basis; done when nothing is left of it
        auto reduce_s = [&]() -> int {       // the lowest bit left of the syndrome (it has no pivot), -1 = nothing left
            for (;;) {
                const int r = osd_lowbit<WPL>(s, lane);
                if (r < 0) { done = true; return -1; }
                const uint32_t p = piv[r];
                if (p == 0xFFFFu) return r;
                const uint32_t wr = (uint32_t)r >> 5;
#pragma unroll
                for (int j = 0; j < WPL; ++j)
                    if ((uint32_t)(j * 32 + lane) >= wr) s[j] ^= basis[((size_t)p * WPL + j) * 32 + lane];
                set_bit(sc, p);
            }
        };
        int rs = reduce_s();
        // The walk: (0) the reliability-ordered candidates; then, for the few shots they do not suffice for, (1) the
        // variables of the detector that is the lowest bit left of the syndrome, in index order, moving on to the next
        // such detector whenever that bit changes (a blind walk through all variables spent ~0.7 s of one warp on such a
        // shot: tens of thousands of dependent columns, each a full reduction chain); (2) if a detector's row is used up
        // and its bit is still there, every variable in index order.  It ends when the syndrome is reduced or the basis is
        // complete.  Same schedule as oracle/bp_oracle.c::orc_osd0 with max_candidates < 0.
        int stage = 0, tr = -1;
        uint32_t i0 = 0, tk = 0, tend = 0, i2 = 0;
        while (!done && nb < (int)a.kmax) {
            uint32_t c;
            if (stage == 0) {
                if (i0 >= OSD_NCAND) { stage = 1; continue; }
                c = cand[i0++];
                if (c == 0xFFFFFFFFu) continue;
            } else if (stage == 1) {
                if (tr != rs) { tr = rs; tk = a.chk_ptr[tr]; tend = a.chk_ptr[tr + 1]; }
                if (tk >= tend) { stage = 2; continue; }
                c = a.chk_var[tk++];
            } else {
                if (i2 >= a.M) break;
                c = i2++;
            }
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
                    const uint32_t wr = (uint32_t)r >> 5;
#pragma unroll
                    for (int j = 0; j < WPL; ++j)
                        if ((uint32_t)(j * 32 + lane) >= wr) v[j] ^= basis[((size_t)p * WPL + j) * 32 + lane];
                    set_bit(vc, p);
                } else {
                    // joins the basis as ordinal nb: the reduced vector, and the set of earlier basis vectors folded into it
#pragma unroll
                    for (int j = 0; j < WPL; ++j) {
                        basis[((size_t)nb * WPL + j) * 32 + lane] = v[j];
                        coef[((size_t)nb * WPL + j) * 32 + lane] = vc[j];
                    }
                    if (lane == 0) { piv[r] = (uint16_t)nb; col[nb] = c; }
                    ++nb;
                    __syncwarp();
                    rs = reduce_s();
                    break;
                }
            }
        }
        if (done) {
            // sc = the basis VECTORS that sum to the syndrome; vector i = column i + the vectors in its set (all of lower
            // ordinal), so a descending sweep turns it into the basis COLUMNS that sum to the syndrome
            for (int i = nb - 1; i >= 0; --i) {
                uint32_t wv = 0;
#pragma unroll
                for (int j = 0; j < WPL; ++j) if ((i >> 10) == j) wv = sc[j];
                const uint32_t word = __shfl_sync(0xffffffffu, wv, (i >> 5) & 31);
                if ((word >> (i & 31)) & 1u) {
#pragma unroll
                    for (int j = 0; j < WPL; ++j) sc[j] ^= coef[((size_t)i * WPL + j) * 32 + lane];
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

// ---- OSD-0 with a fully reduced basis (the default; osd_solve_kernel above keeps the echelon form and is selected by
// GQ_OSD_CHAIN=1 for A/B measurements).  The echelon walk reduces a column by a CHAIN of dependent steps -- find the lowest
// bit, look up its pivot, load that basis vector from L2, XOR, repeat: ~700 cycles per step, ~100 steps per candidate, and
// one unlucky shot (thousands of candidates against a basis of ~900) holds its warp for a quarter of a second.  Here every
// basis vector is kept zero at all OTHER pivots, so a raw column is reduced in ONE step: the vectors of the pivots among
// its <= ~10 detectors, all loads independent and in flight together; what is left is zero (dependent) or a new basis
// vector, whose pivot bit is then cleared from the earlier vectors that carry it (about half of them; independent
// read-modify-writes, eight in flight).  Each vector travels with its composition in basis columns (one bit per ordinal,
// only the words in use are touched), so the syndrome's composition is the answer with no back-substitution.
// Same candidates in the same order, same independence decisions (those do not depend on the form of the basis), same
// unique solution as oracle/bp_oracle.c::orc_osd0 -- including the detector the targeted stage looks at: the lowest bit
// left of the syndrome is the same for every echelon basis of the same span.
template <int WPL>
__global__ void __launch_bounds__(128) osd_rref_kernel(OsdArgs a) {
    extern __shared__ uint16_t piv_all[];
    constexpr int NB = WPL <= 2 ? 8 : (WPL == 4 ? 4 : 2);     // rows in flight per batch
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint16_t *piv = piv_all + (size_t)wib * ((a.D + 1) & ~1u);
    uint32_t *basis = a.scratch + (size_t)warp * a.scratch_stride;
    uint32_t *coef = basis + (size_t)a.kmax * 32 * WPL;
    uint32_t *col = coef + (size_t)a.kmax * 32 * WPL;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(a.cursor, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= a.nfail) break;
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
        int rs = osd_lowbit<WPL>(s, lane);          // lowest bit left of the syndrome (never a pivot), -1 = solved
        int stage = 0, tr = -1;
        uint32_t i0 = 0, tk = 0, tend = 0, i2 = 0;
        while (rs >= 0 && nb < (int)a.kmax) {
            uint32_t c;
            if (stage == 0) {
                if (i0 >= OSD_NCAND) { stage = 1; continue; }
                c = cand[i0++];
                if (c == 0xFFFFFFFFu) continue;
            } else if (stage == 1) {
                if (tr != rs) { tr = rs; tk = a.chk_ptr[tr]; tend = a.chk_ptr[tr + 1]; }
                if (tk >= tend) { stage = 2; continue; }
                c = a.chk_var[tk++];
            } else {
                if (i2 >= a.M) break;
                c = i2++;
            }
            // ---- the column, minus the basis vectors of the pivots among its detectors
            uint32_t v[WPL], vc[WPL];
#pragma unroll
            for (int j = 0; j < WPL; ++j) { v[j] = 0u; vc[j] = 0u; }
            const uint32_t cw = (uint32_t)(nb + 31) >> 5;           // composition words in use
            const uint32_t k0 = a.var_ptr[c], k1 = a.var_ptr[c + 1];
            for (uint32_t kk = k0; kk < k1; kk += 32) {
                const uint32_t k = kk + lane;
                const uint32_t d = k < k1 ? a.var_chk[k] : 0xFFFFFFFFu;
                const uint32_t pm = k < k1 ? (uint32_t)piv[d] : 0xFFFFu;
                const int cnt = (int)min(32u, k1 - kk);
                for (int t = 0; t < cnt; ++t) {
                    const uint32_t dd = __shfl_sync(0xffffffffu, d, t), w = dd >> 5;
#pragma unroll
                    for (int j = 0; j < WPL; ++j)
                        if ((int)(w >> 5) == j && (int)(w & 31u) == lane) v[j] ^= 1u << (dd & 31u);
                }
                unsigned m = __ballot_sync(0xffffffffu, pm != 0xFFFFu);
                while (m) {
                    uint32_t pp[NB];
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        pp[q] = 0xFFFFu;
                        if (m) { const int l = __ffs(m) - 1; m &= m - 1; pp[q] = __shfl_sync(0xffffffffu, pm, l); }
                    }
                    uint32_t rb[NB][WPL], rc[NB][WPL];
#pragma unroll
                    for (int q = 0; q < NB; ++q)
#pragma unroll
                        for (int j = 0; j < WPL; ++j) {
                            const bool on = pp[q] != 0xFFFFu;
                            rb[q][j] = on ? basis[((size_t)pp[q] * WPL + j) * 32 + lane] : 0u;
                            rc[q][j] = (on && (uint32_t)(j * 32 + lane) < cw) ? coef[((size_t)pp[q] * WPL + j) * 32 + lane] : 0u;
                        }
#pragma unroll
                    for (int q = 0; q < NB; ++q)
#pragma unroll
                        for (int j = 0; j < WPL; ++j) { v[j] ^= rb[q][j]; vc[j] ^= rc[q][j]; }
                }
            }
            const int r = osd_lowbit<WPL>(v, lane);
            if (r < 0) continue;                                    // dependent on the basis: skipped
            // ---- joins the basis as ordinal nb, pivot r: clear bit r from the earlier vectors
#pragma unroll
            for (int j = 0; j < WPL; ++j)
                if ((nb >> 10) == j && lane == ((nb >> 5) & 31)) vc[j] |= 1u << (nb & 31);
            const uint32_t wr = (uint32_t)r >> 5, jr = wr >> 5, lr = wr & 31u, bitr = 1u << (r & 31);
            const uint32_t cw1 = ((uint32_t)nb >> 5) + 1u;          // composition words in use, own bit included
            for (int base = 0; base < nb; base += 32) {
                const int i = base + lane;
                const bool has = i < nb && (basis[((size_t)i * WPL + jr) * 32 + lr] & bitr);
                unsigned m = __ballot_sync(0xffffffffu, has);
                while (m) {
                    int idx[NB];
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        idx[q] = -1;
                        if (m) { idx[q] = base + __ffs(m) - 1; m &= m - 1; }
                    }
                    uint32_t rb[NB][WPL], rc[NB][WPL];
                    // v is zero below word wr, and the compositions beyond word cw1: those words are left alone
#pragma unroll
                    for (int q = 0; q < NB; ++q)
#pragma unroll
                        for (int j = 0; j < WPL; ++j) {
                            const uint32_t w = (uint32_t)(j * 32 + lane);
                            rb[q][j] = (idx[q] >= 0 && w >= wr) ? basis[((size_t)idx[q] * WPL + j) * 32 + lane] : 0u;
                            rc[q][j] = (idx[q] >= 0 && w < cw1) ? coef[((size_t)idx[q] * WPL + j) * 32 + lane] : 0u;
                        }
#pragma unroll
                    for (int q = 0; q < NB; ++q)
#pragma unroll
                        for (int j = 0; j < WPL; ++j) {
                            const uint32_t w = (uint32_t)(j * 32 + lane);
                            if (idx[q] >= 0 && w >= wr) basis[((size_t)idx[q] * WPL + j) * 32 + lane] = rb[q][j] ^ v[j];
                            if (idx[q] >= 0 && w < cw1) coef[((size_t)idx[q] * WPL + j) * 32 + lane] = rc[q][j] ^ vc[j];
                        }
                }
            }
#pragma unroll
            for (int j = 0; j < WPL; ++j) {
                basis[((size_t)nb * WPL + j) * 32 + lane] = v[j];
                coef[((size_t)nb * WPL + j) * 32 + lane] = vc[j];   // all words: the scratch still holds the previous shot's
            }
            if (lane == 0) { piv[r] = (uint16_t)nb; col[nb] = c; }
            ++nb;
            __syncwarp();
            // ---- the syndrome: only its bit r can have become reducible
            uint32_t wv = 0;
#pragma unroll
            for (int j = 0; j < WPL; ++j) if ((int)jr == j) wv = s[j];
            if (__shfl_sync(0xffffffffu, wv, (int)lr) & bitr) {
#pragma unroll
                for (int j = 0; j < WPL; ++j) { s[j] ^= v[j]; sc[j] ^= vc[j]; }
                rs = osd_lowbit<WPL>(s, lane);
            }
        }
        if (rs < 0) {
            // the error vector = the basis columns whose bit is set in the syndrome's composition
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
        cudaFuncSetAttribute(osd_rref_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, osmem);
        cudaFuncSetAttribute(osd_rref_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, osmem);
        cudaFuncSetAttribute(osd_rref_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, osmem);
        cudaFuncSetAttribute(osd_rref_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, osmem);
    }
    GQ_CUDA(cudaMalloc((void **)&dec->d_prior, (size_t)M * 4));
    GQ_CUDA(cudaMemcpy(dec->d_prior, prior.data(), (size_t)M * 4, cudaMemcpyHostToDevice));
    dec->bp_msgs_per_block = ((size_t)nnz + M + 31) & ~(size_t)31;
    size_t bytes = dec->bp_msgs_per_block * 4;
    // ---- which kernel: messages in shared memory when the whole model fits (bp_kernel); else the compressed min-sum
    // kernel when its records fit (bp_cms_kernel); else the shot-interleaved global-scratch kernel.  params.bp_scratch
    // forces one of the latter two (1 = scratch, 2 = compressed; tests and A/B measurements).
    const bool small = bytes <= 192 * 1024;
    if (dec->params.bp_scratch == 2 || (!small && dec->params.bp_scratch == 0)) {
        uint32_t maxdeg = 0;
        for (uint32_t m = 0; m < M; ++m) maxdeg = std::max(maxdeg, var_ptr[m + 1] - var_ptr[m]);
        // variables by degree (stable), every degree group padded to whole tiles of 32
        std::vector<uint32_t> order(M);
        for (uint32_t m = 0; m < M; ++m) order[m] = m;
        // inside a degree group: by check list (lexicographic).  The lanes of a tile then read records of the same or of
        // neighbouring checks -- same address = broadcast, consecutive records = distinct banks -- where a random order costs
        // a 3-4-way bank conflict per record read (GQ_CMS_ORDER = 1: by prior LLR, 2: by index; A/B measurements)
        int order_mode = 0;
        if (const char *env = getenv("GQ_CMS_ORDER")) order_mode = atoi(env);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
            const uint32_t dx = var_ptr[x + 1] - var_ptr[x], dy = var_ptr[y + 1] - var_ptr[y];
            if (dx != dy) return dx < dy;
            if (order_mode == 1) return prior[x] < prior[y];
            if (order_mode == 2) return false;
            return std::lexicographical_compare(var_chk.begin() + var_ptr[x], var_chk.begin() + var_ptr[x + 1],
                                                var_chk.begin() + var_ptr[y], var_chk.begin() + var_ptr[y + 1]);
        });
        std::vector<uint32_t> desc, tvar, tobs;
        uint32_t grp_tile[18], grp_word[17];
        for (int g = 0; g < 18; ++g) grp_tile[g] = 0xFFFFFFFFu;
        std::vector<float> tprior;
        std::vector<uint16_t> tchk;
        uint32_t words = 0;
        for (size_t i = 0; i < order.size();) {
            const uint32_t deg = var_ptr[order[i] + 1] - var_ptr[order[i]];
            size_t j = i;
            while (j < order.size() && j - i < 32 && var_ptr[order[j] + 1] - var_ptr[order[j]] == deg) ++j;
            if (deg <= 16 && grp_tile[deg] == 0xFFFFFFFFu) { grp_tile[deg] = (uint32_t)desc.size(); grp_word[deg] = words; }
            desc.push_back((words << 5) | deg);
            const size_t e0 = tchk.size();
            tchk.resize(e0 + (size_t)32 * deg, 0);
            for (size_t l = 0; l < 32; ++l) {
                if (i + l < j) {
                    const uint32_t m = order[i + l];
                    tvar.push_back(m); tprior.push_back(prior[m]); tobs.push_back(var_obs[m]);
                    for (uint32_t q = 0; q < deg; ++q) tchk[e0 + (size_t)32 * q + l] = (uint16_t)var_chk[var_ptr[m] + q];
                } else {      // padding lane: variable M, all edges at the dummy check D, never decided "error"
                    tvar.push_back(M); tprior.push_back(1e30f); tobs.push_back(0u);
                    for (uint32_t q = 0; q < deg; ++q) tchk[e0 + (size_t)32 * q + l] = (uint16_t)D;
                }
            }
            words += deg;
            i = j;
        }
        const uint32_t ntiles = (uint32_t)desc.size(), FW = (D + 1 + 15) / 16;
        grp_tile[17] = ntiles;
        for (int g = 16; g >= 0; --g) if (grp_tile[g] == 0xFFFFFFFFu) { grp_tile[g] = grp_tile[g + 1]; grp_word[g] = 0; }   // empty group
        const size_t smem = std::max((size_t)(D + 1) * 24 + ((size_t)2 * FW + words + ntiles + dem->DW) * 4, (size_t)OSD_WS_WORDS * 8);
        const bool eligible = maxdeg <= 16 && D < 65535 && M < (1u << 31) - 1 && words < (1u << 27) && smem <= 220 * 1024;
        if (eligible) {
            dec->cms = true;
            dec->cms_ntiles = ntiles; dec->cms_FW = FW; dec->cms_SW = words; dec->cms_smem = smem;
            // two blocks per SM when two sets of records fit (their barriers overlap), else one of 1024 threads
            const bool two = 2 * (smem + 1024) <= 227 * 1024;
            dec->cms_threads = two ? 512 : 1024;
            dec->cms_blocks = dem->num_sms * (two ? 2 : 1);
            if (const char *env = getenv("GQ_CMS_THREADS")) {      // A/B measurements: threads per block, blocks per SM = 1024 / threads
                const int t = atoi(env);
                if ((t == 256 || t == 512 || t == 1024) && (size_t)(1024 / t) * (smem + 1024) <= 227 * 1024) { dec->cms_threads = t; dec->cms_blocks = dem->num_sms * (1024 / t); }
            }
            if (cudaFuncSetAttribute(bp_cms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) GQ_FAIL(GQ_E_CUDA, "BP decoder: cudaFuncSetAttribute failed");
            memcpy(dec->cms_grp_tile, grp_tile, sizeof(grp_tile));
            memcpy(dec->cms_grp_word, grp_word, sizeof(grp_word));
            if ((rc = up(&dec->d_cms_var, tvar)) || (rc = up(&dec->d_cms_obs, tobs))) return rc;
            GQ_CUDA(cudaMalloc((void **)&dec->d_cms_prior, tprior.size() * 4));
            GQ_CUDA(cudaMemcpy(dec->d_cms_prior, tprior.data(), tprior.size() * 4, cudaMemcpyHostToDevice));
            GQ_CUDA(cudaMalloc((void **)&dec->d_cms_chk, std::max<size_t>(tchk.size(), 1) * 2));
            GQ_CUDA(cudaMemcpy(dec->d_cms_chk, tchk.data(), tchk.size() * 2, cudaMemcpyHostToDevice));
            GQ_CUDA(cudaMalloc((void **)&dec->d_cms_cursor, 4));
            if (dec->params.bp_osd) GQ_CUDA(cudaMalloc((void **)&dec->d_cms_post, (size_t)dec->cms_blocks * (M + 1) * 4));
            return GQ_OK;
        }
        if (dec->params.bp_scratch == 2) GQ_FAIL(GQ_E_UNSUPPORTED, "BP decoder: the compressed min-sum kernel needs <= 16 detectors per mechanism, <= 65535 detectors and records that fit shared memory");
    }
    if (small && !dec->params.bp_scratch) {
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
    cudaFree(dec->d_cms_var); cudaFree(dec->d_cms_obs); cudaFree(dec->d_cms_cursor);
    cudaFree(dec->d_cms_prior); cudaFree(dec->d_cms_post); cudaFree(dec->d_cms_chk);
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
    a.use_shared = dec->d_bp_msgs == nullptr && !dec->cms;
    size_t smem = a.use_shared ? dec->bp_msgs_per_block * 4 : 0;
    const bool osd = dec->params.bp_osd != 0;
    const uint64_t chunk = osd ? 32768 : nshots;       // OSD keeps 8 KB of candidates per unconverged shot; the OSD kernel ends with its
                                                        // longest walk, so bigger chunks amortise that tail (8192 shots 3.7e4/s, 16384 4.3e4/s)
    const int wpl = osd_wpl(dem->DW);
    int osd_mult = 1;
    if (const char *env = getenv("GQ_OSD_WARPS")) osd_mult = std::max(1, std::min(8, atoi(env)));
    const int osd_blocks = dem->num_sms * 2 * osd_mult, osd_warps = osd_blocks * 4;
    const uint32_t osd_kmax = dem->D;
    const size_t osd_stride = (size_t)osd_kmax * (64 * (size_t)wpl + 1);
    if (osd) {
        if (!dec->d_osd_cand) {
            GQ_CUDA(cudaMalloc((void **)&dec->d_osd_cand, chunk * OSD_NCAND * 4));
            GQ_CUDA(cudaMalloc((void **)&dec->d_osd_fail, (chunk + 4) * 4));   // [0] unconverged count, [1] OSD give-ups, [2] work cursor, [4..] shots
            GQ_CUDA(cudaMalloc((void **)&dec->d_osd_scratch, (size_t)osd_warps * osd_stride * 4));
        }
        a.osd_k = OSD_NCAND; a.cand = dec->d_osd_cand; a.fail_list = dec->d_osd_fail + 4; a.fail_count = dec->d_osd_fail;
    }
    for (uint64_t s0 = 0; s0 < nshots; s0 += chunk) {
        const uint64_t ns = std::min(chunk, nshots - s0);
        a.dets = dets_sm + s0 * dem->DW; a.pred = pred_sm + s0 * dem->OW;
        a.iters_out = weight_out ? weight_out + s0 : nullptr;
        a.flags_out = flags_out ? flags_out + s0 : nullptr;
        a.corr_out = corr_out ? corr_out + s0 * a.MW : nullptr;
        a.nshots = ns;
        if (osd) GQ_CUDA(cudaMemsetAsync(dec->d_osd_fail, 0, 16, dem->stream));
        if (dec->cms) {
            CmsArgs c;
            memset(&c, 0, sizeof(c));
            c.dets = a.dets; c.pred = a.pred; c.iters_out = a.iters_out; c.flags_out = a.flags_out; c.corr_out = a.corr_out;
            c.nshots = ns; c.D = a.D; c.M = a.M; c.DW = a.DW; c.OW = a.OW; c.MW = a.MW;
            c.ntiles = dec->cms_ntiles; c.FW = dec->cms_FW; c.SW = dec->cms_SW;
            memcpy(c.grp_tile, dec->cms_grp_tile, sizeof(c.grp_tile));
            memcpy(c.grp_word, dec->cms_grp_word, sizeof(c.grp_word));
            c.tile_var = dec->d_cms_var; c.tile_prior = dec->d_cms_prior;
            c.tile_obs = dec->d_cms_obs; c.tile_chk = dec->d_cms_chk;
            c.max_iter = a.max_iter; c.alpha = a.alpha; c.cursor = dec->d_cms_cursor; c.post = dec->d_cms_post;
            c.osd_k = a.osd_k; c.cand = a.cand; c.fail_list = a.fail_list; c.fail_count = a.fail_count;
            GQ_CUDA(cudaMemsetAsync(dec->d_cms_cursor, 0, 4, dem->stream));
            const int blocks = (int)std::min<uint64_t>(ns, (uint64_t)dec->cms_blocks);
            // the limit is per function, not per decoder: another decoder of this process may have set a smaller one
            GQ_CUDA(cudaFuncSetAttribute(bp_cms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec->cms_smem));
            bp_cms_kernel<<<blocks, dec->cms_threads, dec->cms_smem, dem->stream>>>(c);
        } else if (a.use_shared) {
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
        o.cand = dec->d_osd_cand; o.fail_list = dec->d_osd_fail + 4; o.nfail = nfail; o.cursor = dec->d_osd_fail + 2;
        o.chk_ptr = dec->d_chk_ptr; o.chk_var = dec->d_chk_var;
        o.kmax = osd_kmax; o.giveups = dec->d_osd_fail + 1;
        o.sentinel = dem->O < 32 ? 0xFFFFFFFFu : 0u;
        o.D = dem->D; o.DW = dem->DW; o.OW = dem->OW; o.MW = a.MW; o.M = dem->M;
        o.var_ptr = dec->d_var_ptr; o.var_chk = dec->d_var_chk; o.var_obs = dec->d_var_obs;
        o.scratch = dec->d_osd_scratch; o.scratch_stride = osd_stride;
        const int ob = (int)std::min<uint32_t>((nfail + 3) / 4, (uint32_t)osd_blocks);
        const size_t osmem = 4 * (size_t)((dem->D + 1) & ~1u) * 2;
        static const bool chain = getenv("GQ_OSD_CHAIN") != nullptr;      // A/B: the echelon-form walk
        {   // per-function limit (see above)
            const void *fn = chain ? (wpl == 1 ? (const void *)osd_solve_kernel<1> : wpl == 2 ? (const void *)osd_solve_kernel<2> : wpl == 4 ? (const void *)osd_solve_kernel<4> : (const void *)osd_solve_kernel<8>)
                                   : (wpl == 1 ? (const void *)osd_rref_kernel<1> : wpl == 2 ? (const void *)osd_rref_kernel<2> : wpl == 4 ? (const void *)osd_rref_kernel<4> : (const void *)osd_rref_kernel<8>);
            GQ_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)osmem));
        }
        if (chain) {
            switch (wpl) {
            case 1: osd_solve_kernel<1><<<ob, 128, osmem, dem->stream>>>(o); break;
            case 2: osd_solve_kernel<2><<<ob, 128, osmem, dem->stream>>>(o); break;
            case 4: osd_solve_kernel<4><<<ob, 128, osmem, dem->stream>>>(o); break;
            default: osd_solve_kernel<8><<<ob, 128, osmem, dem->stream>>>(o); break;
            }
        } else {
            switch (wpl) {
            case 1: osd_rref_kernel<1><<<ob, 128, osmem, dem->stream>>>(o); break;
            case 2: osd_rref_kernel<2><<<ob, 128, osmem, dem->stream>>>(o); break;
            case 4: osd_rref_kernel<4><<<ob, 128, osmem, dem->stream>>>(o); break;
            default: osd_rref_kernel<8><<<ob, 128, osmem, dem->stream>>>(o); break;
            }
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
