// K3a, first tier: one shot per warp, no materialised weight matrix.
//
// decode_shot<K>() takes the defect list of one shot (n <= 32*K defects, K = 1..8, 10, 12, 16, 24, 32; ascending
// detector index) and returns the exact minimum-weight matching's observable prediction:
//   1. nearest other defect of every defect: the NBR_LEN = 12 nearest detectors of its detector (a static sorted
//      list) are tested against the syndrome bits in shared memory; the few defects with no fired
//      detector among them take a full row of the L2-resident distance table instead;
//   2. dual start y(v) = min(nearest/2, boundary distance): a vertex whose boundary is at least as
//      close as half its nearest defect starts matched to the boundary; mutual nearest neighbours
//      are tight and start matched to each other (one lane-parallel step, no sequential greedy);
//   3. solve_tree<K>(): primal-dual blossom, one alternating tree per remaining free vertex, the
//      boundary as an always-free partner.  Vertex v lives in slot (v >> 5) of lane (v & 31): dual,
//      pending event time, label are registers; the blossom forest (lane 0 walks it) sits in shared
//      memory at compile-time offsets; weights are gathered straight from the distance table (row
//      of the scanning vertex, columns = this lane's defects);
//   4. XOR of the path observables of the matched pairs.
// Replaces PyMatching's decode_batch together with matching.cu.
//
// Dual bookkeeping: units are quadrupled (y_u + y_v <= 4 w, y_v <= 4 b_v; all start values even) so
// every event time is an integer.  Inside a phase values are stored against the phase clock D, so
// a dual step touches nothing:
//   vertex dual   y = yb + sg*D        (sg = +1 outer "S", -1 inner "T", 0 unlabelled)
//   blossom dual  z = zb + 2*sg*D
//   tm            clock value at which the best known edge from an outer vertex of another node
//                 turns tight (unlabelled target: D0 + slack, outer target: D0 + slack/2); for an
//                 inner vertex its frozen slack.
// The only stale entries are outer-outer ones whose source was swallowed by the target's blossom;
// they are lower bounds and are recomputed when they win the search.
//
// The same source compiles for the host with one lane (GQ_HOST_SIM, tests/hostsim/) so that the
// logic can be checked against the oracle on a CPU-only box; that build is test scaffolding.
#pragma once
#include <stdint.h>

#include "blossom_regs.cuh"

#ifdef GQ_HOST_SIM
#define GQW_LANES 1
#define GQW_LOG 0
typedef unsigned __int128 gqw_mask;   // one bit per slot; the single host lane owns every slot
#else
#define GQW_LANES 32
#define GQW_LOG 5
typedef unsigned long long gqw_mask;   // one bit per slot of a lane (K <= 64)
#endif
#define GQW_FULL 0xffffffffu
#define GQW_BIT(k) ((gqw_mask)1 << (k))

#ifdef GQ_MT_STATS
#define GQM_STAT(i, x) do { if (lane == 0) gqm_stat_add((i), (x)); } while (0)
#else
#define GQM_STAT(i, x) do { } while (0)
#endif

namespace gqt {

using gqr::i16;
using gqr::i32;
using gqr::u16;
using gqr::L_NONE;
using gqr::L_S;
using gqr::L_T;
using gqr::SL_INF;
using gqr::W_INF;

struct Tables {
    const u16 *dist;       // [N][N]: row/column D = the boundary, D + 1 = nothing (all W_INF); diagonal = W_INF
    const uint8_t *pobs;   // [N][N] observable mask of the shortest path
    uint32_t N, D;         // N = D + 2
    const uint32_t *nbr;   // [N][NBR_LEN]: the nearest other detectors of each detector, ascending distance,
                           // entry = distance << 16 | detector; padding = 0xffff << 16 | 32 * DW (a detector
                           // index whose syndrome word is the always-zero word behind the row)
};

#ifndef GQ_FP2_MAXK
#define GQ_FP2_MAXK 3   // classes up to this K try the second-event fast path: from K = 4 on its code costs more in
                      // instruction-cache misses (hit rate 83 % at K = 4) than it saves (measured: K >= 4 without it -1.7 % at the
                      // north star, -7 % at p = 0.008)
#endif
#ifndef GQ_NBR_LEN
#define GQ_NBR_LEN 12   // measured at the north star: 8 -> 37.5, 12 -> 36.8, 16 -> 37.8, 24 -> 39.1 ms per 2^22 shots
#endif
static constexpr int NBR_LEN = GQ_NBR_LEN;   // a multiple of 4 (16-byte loads)
static constexpr int SYN_WORDS = 256;   // syndrome rows up to this many words are mirrored in shared memory

// ---- defect list of one bit-packed syndrome row (ascending detector index); returns the count,
// stores the first cap entries
__device__ __forceinline__ int extract_defects(const uint32_t *row, uint32_t DW, uint32_t D, int lane,
                                               u16 *def, int cap, uint32_t *synw = nullptr, u16 *spre = nullptr) {
    int n0 = 0;
    for (uint32_t w0 = 0; w0 < DW; w0 += GQW_LANES) {
        const uint32_t wi = w0 + lane;
        uint32_t word = wi < DW ? row[wi] : 0u;
        if (wi == DW - 1 && (D & 31)) word &= (1u << (D & 31)) - 1;
        const int cnt = __popc(word);
        int pre = cnt;
#pragma unroll
        for (int o = 1; o < GQW_LANES; o <<= 1) {
            int t = __shfl_up_sync(GQW_FULL, pre, o);
            if (lane >= o) pre += t;
        }
        const int tot = __shfl_sync(GQW_FULL, pre, GQW_LANES - 1);
        int pos = n0 + pre - cnt;
        if (synw && wi <= DW) { synw[wi] = word; spre[wi] = (u16)pos; }   // word DW = the zero word padding points at
        while (word) {
            int b = __ffs(word) - 1;
            word &= word - 1;
            if (pos < cap) def[pos] = (u16)(wi * 32 + b);
            ++pos;
        }
        n0 += tot;
    }
    if (synw && (DW % GQW_LANES) == 0 && lane == 0) { synw[DW] = 0u; spre[DW] = (u16)n0; }
    __syncwarp();
    return n0;
}

// arr[slot] for a warp-uniform slot, written as a select tree on the bits of slot: a chain of
// "slot == k" selects gets folded back into a dynamically indexed (local-memory) array by the compiler
// generic select tree over [LO, LO + N), N a power of two; indices at or beyond K read as 0
template <int LO, int N, int K, class TV>
__device__ __forceinline__ int sel_rec(const TV (&arr)[K], const int slot) {
    if constexpr (N == 1) return LO < K ? (int)arr[LO < K ? LO : 0] : 0;
    else return (slot & (N / 2)) ? sel_rec<LO + N / 2, N / 2, K>(arr, slot) : sel_rec<LO, N / 2, K>(arr, slot);
}

template <int K, class TV>
__device__ __forceinline__ int sel_slot(const TV (&arr)[K], const int slot) {
    if constexpr (K == 1) return (int)arr[0];
    else if constexpr (K == 2) return (slot & 1) ? (int)arr[1] : (int)arr[0];
    else if constexpr (K == 3) return (slot & 2) ? (int)arr[2] : ((slot & 1) ? (int)arr[1] : (int)arr[0]);
    else if constexpr (K <= 8) {
        const int lo = (slot & 2) ? ((slot & 1) ? (int)arr[3] : (int)arr[2]) : ((slot & 1) ? (int)arr[1] : (int)arr[0]);
        if constexpr (K == 4) return lo;
        else {
            int hi = (int)arr[4];
            if constexpr (K >= 6) hi = (slot & 1) ? (int)arr[5] : hi;
            if constexpr (K >= 7) {
                int h2 = (int)arr[6];
                if constexpr (K >= 8) h2 = (slot & 1) ? (int)arr[7] : h2;
                hi = (slot & 2) ? h2 : hi;
            }
            return (slot & 4) ? hi : lo;
        }
    } else if constexpr (K <= 16) {
        // two halves of (up to) eight, each a select tree
        int lo, hi;
        {
            const int a0 = (slot & 2) ? ((slot & 1) ? (int)arr[3] : (int)arr[2]) : ((slot & 1) ? (int)arr[1] : (int)arr[0]);
            const int a1 = (slot & 2) ? ((slot & 1) ? (int)arr[7] : (int)arr[6]) : ((slot & 1) ? (int)arr[5] : (int)arr[4]);
            lo = (slot & 4) ? a1 : a0;
        }
        {
            auto at = [&](int i) -> int { return i < K ? (int)arr[i < K ? i : 0] : 0; };
            const int b0 = (slot & 2) ? ((slot & 1) ? at(11) : at(10)) : ((slot & 1) ? at(9) : at(8));
            const int b1 = (slot & 2) ? ((slot & 1) ? at(15) : at(14)) : ((slot & 1) ? at(13) : at(12));
            hi = (slot & 4) ? b1 : b0;
        }
        return (slot & 8) ? hi : lo;
    } else if constexpr (K <= 32) {
        return sel_rec<0, 32, K>(arr, slot);
    } else return (int)arr[slot];
}

// ---- single-tree blossom phases with native boundary.
// In: f.mate[] initial matching (-1 free, -2 boundary, else partner; every matched edge tight),
// yb[] feasible duals (quadrupled units, even), b4[] 4*boundary distance or SL_INF, freebits = slots
// of this lane that are free.  Out: f.mate[].  Returns 0 ok, 1 no perfect matching, 2 blossom ids
// exhausted, 3 step guard.
//
// Per-slot event keys (one register each, so the search is a handful of integer minima):
//   ek = pending edge event  : time << TS | type << VB | vertex   (type 0: an outer vertex reaches this
//        unlabelled vertex, 1: outer-outer edge turns tight); an inner vertex keeps its frozen slack
//        in the same layout with bit 31 set, which also keeps it out of the search
//   bk = boundary event of an outer vertex: time << TS | 3 << VB | vertex
// Label of a slot = the key bits it contributes: kb = vertex (unlabelled), vertex | 0x100 (outer),
// vertex | bit 31 (inner).  With X = 4w - y_u + (D - yb[k]) the candidate time of an edge from outer u
// is X for unlabelled and inner targets (time resp. frozen slack) and X / 2 for outer ones.
// Field widths: the vertex takes VB = 8 bits up to 256 defects per shot, 9 bits up to 512, 10 bits up to 1024 and 11 bits
// up to 2048, the type 2 bits above it, the time the rest below bit 31 (21, 20, 19 resp. 18 bits; times are bounded by
// 8 x the largest table distance, a u16: below 2^19 always, below 2^18 when that distance is below 2^15 -- the host only
// uses the 11-bit classes then).
static const unsigned EK_NONE = 0x7fffffffu;   // no pending edge event (time field all ones)
static const unsigned EK_FROZEN = 0x80000000u;
template <int K> struct KeyBits {
    static constexpr int VB = (K * GQW_LANES > 1024) ? 11 : ((K * GQW_LANES > 512) ? 10 : ((K * GQW_LANES > 256) ? 9 : 8));
    static constexpr int TS = VB + 2;
    static constexpr unsigned VM = (1u << VB) - 1u;
    static constexpr unsigned TY1 = 1u << VB, TY2 = 2u << VB, TY3 = 3u << VB;
    static constexpr unsigned LIM = (EK_NONE >> TS) << TS;     // keys at or above this hold no event
};

template <int TS> __device__ __forceinline__ bool ek_has(unsigned ek) { return (ek | 0x80000000u | ((1u << TS) - 1u)) != 0xffffffffu; }
template <int TS> __device__ __forceinline__ i32 ek_time(unsigned ek) { return (i32)((ek >> TS) & ((1u << (31 - TS)) - 1u)); }

template <int K, class FT>
__device__ __forceinline__ int solve_tree(FT &f, const u16 *__restrict__ dist, const uint32_t N,
                                          const u16 *def, const int n, const int lane,
                                          const uint32_t (&dj)[K], const i32 (&b4)[K], i32 (&yb)[K],
                                          gqw_mask freebits) {
    constexpr int cap = FT::cap;
    constexpr int VB = KeyBits<K>::VB, TS = KeyBits<K>::TS;
    constexpr unsigned VM = KeyBits<K>::VM, TY1 = KeyBits<K>::TY1, TY2 = KeyBits<K>::TY2, TY3 = KeyBits<K>::TY3, KB_OUTER = TY1, EK_LIM = KeyBits<K>::LIM;
    unsigned ek[K], bk[K], kb[K];
    int slfrom[K], vtop[K];
    i32 D = 0;
    int bhi = 0;

#define GQT_SEL(arr, u, out) { out = __shfl_sync(GQW_FULL, sel_slot<K>(arr, (u) >> GQW_LOG), (u) & (GQW_LANES - 1)); }
#define GQT_IS_OUTER(k_) ((kb[k_] & KB_OUTER) != 0)
#define GQT_IS_INNER(k_) ((int)kb[k_] < 0)
// slot k_ turns outer at clock D (yb[k_] already rebased): boundary event, key bits
#define GQT_MAKE_OUTER(k_)                                                                      \
    {                                                                                           \
        kb[k_] = (unsigned)(lane + GQW_LANES * (k_)) | KB_OUTER;                                \
        bk[k_] = b4[k_] < SL_INF ? (((unsigned)(b4[k_] - yb[k_]) << TS) | TY3 | (unsigned)(lane + GQW_LANES * (k_))) : 0xffffffffu; \
    }
// scan the row of outer vertex u_ (its dual is yb + D): offer its edges to every vertex of other nodes
#define GQT_SCAN(u_)                                                                            \
    {                                                                                           \
        const int su_ = (u_);                                                                   \
        const int tu_ = f.top[su_];                                                             \
        i32 yu_;                                                                                \
        GQT_SEL(yb, su_, yu_);                                                                  \
        const i32 e0_ = -yu_;            /* D - (yb_u + D) */                                   \
        const uint32_t ro_ = (uint32_t)def[su_] * N;                                            \
        _Pragma("unroll") for (int k_ = 0; k_ < K; ++k_) {                                      \
            const i32 wt_ = (i32)dist[ro_ + dj[k_]];                                            \
            const i32 x_ = 4 * wt_ + (e0_ - yb[k_]);                                            \
            const unsigned nk_ = ((unsigned)(x_ >> ((kb[k_] >> VB) & 1u)) << TS) + kb[k_];       \
            const bool up_ = wt_ != (i32)W_INF && vtop[k_] != tu_ && nk_ < ek[k_];              \
            ek[k_] = up_ ? nk_ : ek[k_];                                                        \
            slfrom[k_] = up_ ? su_ : slfrom[k_];                                                \
        }                                                                                       \
    }

#pragma unroll 1
    for (int x = lane; x < n; x += GQW_LANES) {
        f.base[x] = (i16)x; f.par[x] = -1; f.label[x] = L_NONE; f.mark[x] = 0; f.top[x] = (i16)x;
    }
#pragma unroll 1
    for (int i = lane; i < FT::nb; i += GQW_LANES) f.bused[i] = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        vtop[k] = lane + GQW_LANES * k;
        kb[k] = (unsigned)(lane + GQW_LANES * k);
        bk[k] = 0xffffffffu;
    }
    __syncwarp();

#pragma unroll 1
    for (int rk = 0; rk < K; ++rk) {
        unsigned rm = __ballot_sync(GQW_FULL, (freebits >> rk) & 1);
        while (rm) {
            const int rb = __ffs(rm) - 1;
            rm &= rm - 1;
            const int root = GQW_LANES * rk + rb;
            if (f.mate[root] != -1) continue;   // became the far end of an earlier augmenting path
            // ---- phase: one alternating tree rooted at root
            D = 0;
#ifdef GQ_MT_STATS
            int ph_events = 0;
#endif
            unsigned nm[K];          // warp-uniform masks of the vertices to scan next, per slot ...
            int sv;                  // ... or the single vertex to scan (the common case), -1: use the masks
            bool have_masks = false; // nm[] may hold vertices (set by the blossom events only)
            bool inner_blossoms = false;   // an inner (T) top-level blossom exists in this tree
            {
                // ---- the root's row, scanned once.  At the start of a phase every vertex is unlabelled, so the
                // offers are plain "reach" events (type 0) and the root's own boundary event.  When the earliest of
                // them ends the phase at once -- the boundary, a free vertex or a vertex that sits on the boundary
                // (a third of all phases at the north star) -- the match and the root's dual are written directly
                // and no tree is set up; otherwise the offers become the tree's first event keys.
                const bool mine = lane == (root & (GQW_LANES - 1));
                i32 rb4, ryb;
                GQT_SEL(b4, root, rb4);
                GQT_SEL(yb, root, ryb);
                const unsigned rbk = rb4 < SL_INF ? (((unsigned)(rb4 - ryb) << TS) | TY3 | (unsigned)root) : 0xffffffffu;
                const uint32_t ro = (uint32_t)def[root] * N;
                unsigned best0 = rbk;
                GQM_STAT(6, 1);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const i32 wt = (i32)dist[ro + dj[k]];      // the diagonal is infinite: the root offers nothing to itself
                    const i32 x = 4 * wt - ryb - yb[k];
                    ek[k] = wt != (i32)W_INF ? (((unsigned)x << TS) + kb[k]) : EK_NONE;
                    best0 = min(best0, ek[k]);
                }
                best0 = __reduce_min_sync(GQW_FULL, best0);
                if (best0 < EK_LIM) {
                    const int v0 = best0 & VM;
                    int done = 0;
                    if (((best0 >> VB) & 3) == 3) {
                        if (lane == 0) f.mate[root] = -2;
                        done = 1;
                    } else if (f.top[v0] == v0 && f.mate[v0] < 0) {
                        if (lane == 0) { f.mate[root] = (i16)v0; f.mate[v0] = (i16)root; }
                        done = 1;
                    }
                    if (done) {
                        GQM_STAT(0, 1); GQM_STAT(4, 1); GQM_STAT(11, 1);
#pragma unroll
                        for (int k = 0; k < K; ++k)
                            if (mine && k == rk) yb[k] += (i32)(best0 >> TS);
                        __syncwarp();
                        continue;
                    }
                    // ---- one step further (another 40 % of the phases): the root reaches a plain vertex v0 matched
                    // to a plain vertex w0 at clock t1; w0 turns outer and offers its row.  If the next event is an
                    // augmentation -- the root or w0 reaches the boundary, a free vertex or a vertex sitting on the
                    // boundary -- the two or three new mates and duals are written directly.  Candidates, with the
                    // source in the type bits: 0 root -> k (the offers above), 1 w0 -> k, 2 the edge root - w0 turning
                    // tight (a blossom: left to the tree code), 3 boundary of root / w0.
                    const int w0 = (K <= GQ_FP2_MAXK && f.top[v0] == v0) ? (int)f.mate[v0] : -1;
                    if (K <= GQ_FP2_MAXK && ((best0 >> VB) & 3) == 0 && w0 >= 0 && f.top[w0] == w0) {
                        const i32 t1 = (i32)(best0 >> TS);
                        i32 yw, wb4;
                        GQT_SEL(yb, w0, yw);
                        GQT_SEL(b4, w0, wb4);
                        const uint32_t rw = (uint32_t)def[w0] * N;
                        const i32 ew = t1 - yw;          // w0 reaches k at clock 4 d(w0, k) - y_k + ew
                        unsigned b2 = rbk;
                        if (wb4 < SL_INF) b2 = min(b2, ((unsigned)(wb4 + ew) << TS) | TY3 | (unsigned)w0);
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const int x = lane + GQW_LANES * k;
                            const i32 wt = (i32)dist[rw + dj[k]];
                            unsigned c = ek[k];
                            if (x == v0) c = 0xffffffffu;
                            if (x == w0) c = ek_has<TS>(ek[k]) ? (((unsigned)((ek_time<TS>(ek[k]) + t1) >> 1) << TS) | TY2 | (unsigned)x) : 0xffffffffu;
                            const unsigned c2 = (wt != (i32)W_INF && x != v0 && x != root)
                                                    ? ((((unsigned)(4 * wt - yb[k] + ew)) << TS) | TY1 | (unsigned)x) : 0xffffffffu;
                            b2 = min(b2, min(c, c2));
                        }
                        b2 = __reduce_min_sync(GQW_FULL, b2);
                        const int ty = (b2 >> VB) & 3, k2 = b2 & VM;
                        if (b2 < EK_LIM && ty != 2 && (ty == 3 || (f.top[k2] == k2 && f.mate[k2] < 0))) {
                            GQM_STAT(0, 2); GQM_STAT(2, 1); GQM_STAT(4, 1); GQM_STAT(6, 1); GQM_STAT(12, 1);
                            const bool from_root = ty == 0 || (ty == 3 && k2 == root);
                            if (lane == 0) {
                                const int src = from_root ? root : w0;
                                if (!from_root) { f.mate[root] = (i16)v0; f.mate[v0] = (i16)root; }
                                if (ty == 3) f.mate[src] = -2;
                                else { f.mate[src] = (i16)k2; f.mate[k2] = (i16)src; }
                            }
                            const i32 t2 = (i32)(b2 >> TS), dt = t2 - t1;
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                const int x = lane + GQW_LANES * k;
                                yb[k] += x == root ? t2 : (x == w0 ? dt : (x == v0 ? -dt : 0));
                            }
                            __syncwarp();
                            continue;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    slfrom[k] = ek[k] != EK_NONE ? root : -1;
                    if (mine && k == rk) { kb[k] = (unsigned)root | KB_OUTER; bk[k] = rbk; }
                }
                if (mine) f.label[root] = L_S;
            }
#pragma unroll
            for (int k = 0; k < K; ++k) nm[k] = 0u;
            sv = -1;
            __syncwarp();
            int fail = 0;
            int guard = 8 * (cap + FT::nb) + 64;   // bounds the blossom events of a phase (grow steps are bounded by n)
            for (;;) {
                // ---- scan the rows of the vertices that just turned outer (one body, runtime loops)
                for (;;) {
                    if (sv < 0) {
                        if (!have_masks) break;
                        // pop the lowest flagged vertex off the masks (blossom events only)
#pragma unroll
                        for (int k = 0; k < K; ++k)
                            if (sv < 0 && nm[k]) { sv = GQW_LANES * k + __ffs(nm[k]) - 1; nm[k] &= nm[k] - 1; }
                        if (sv < 0) { have_masks = false; break; }
                    }
                    GQM_STAT(6, 1);
                    GQT_SCAN(sv);
                    sv = -1;
                }
                // ---- earliest event
                unsigned best = 0xffffffffu;
#pragma unroll
                for (int k = 0; k < K; ++k) best = min(best, min(ek[k], bk[k]));
                if (inner_blossoms) {
#pragma unroll 1
                    for (int i = lane; i < bhi; i += GQW_LANES)
                        if (f.bused[i] && f.par[cap + i] < 0 && f.label[cap + i] == L_T)
                            best = min(best, ((unsigned)(f.yb[i] >> 1) << TS) | TY2 | (unsigned)i);
                }
                best = __reduce_min_sync(GQW_FULL, best);
                GQM_STAT(0, 1);
#ifdef GQ_MT_STATS
                ++ph_events;
#endif
                if (best >= EK_LIM) { fail = 1; break; }
                const i32 t = (i32)(best >> TS);
                const int type = (best >> VB) & 3;
                const int arg = best & VM;
                int src = -1;
                if (type < 2) GQT_SEL(slfrom, arg, src);
#ifdef GQ_HOST_SIM
                // stale entries (source and target in one node) are dropped when their blossom forms, so none
                // can win the search; the host build checks that
                if (type == 1 && f.top[src] == f.top[arg]) { GQM_STAT(1, 1); fail = 3; break; }
#endif
                D = t;
                if (type == 0) {
                    // ---- an outer vertex reaches an unlabelled node
                    const int v = arg, u = src;
                    const int tn = f.top[v];
                    const int m = f.mate[f.base[tn]];
                    if (m < 0) {
                        // the node is free (m == -1, a plain vertex) or sits on the boundary
                        // (m == -2): u takes it over; a boundary match is simply released
                        GQM_STAT(4, 1); GQM_STAT(10 + (ph_events > 4 ? 4 : ph_events), 1); if (m == -1) GQM_STAT(15, 1);
                        if (lane == 0) { if (tn >= cap) gqr::rebase(f, tn, v); gqr::augment(f, u, v); }
                        break;
                    }
                    GQM_STAT(2, 1);
                    const int sn = f.top[m];
                    if (lane == 0) {
                        f.label[tn] = L_T; f.pu[tn] = (i16)u; f.pv[tn] = (i16)v; f.label[sn] = L_S;
                        if (tn >= cap) f.yb[tn - cap] += 2 * D;   // z = zb - 2D from now on
                        if (sn >= cap) f.yb[sn - cap] -= 2 * D;   // z = zb + 2D
                    }
                    if (tn >= cap) inner_blossoms = true;
                    bool turned[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        turned[k] = false;
                        if (vtop[k] == tn) {
                            yb[k] += D;
                            ek[k] = ek_has<TS>(ek[k]) ? ((ek[k] - ((unsigned)D << TS)) | EK_FROZEN) : 0xffffffffu;
                            kb[k] = (unsigned)(lane + GQW_LANES * k) | EK_FROZEN;
                        } else if (vtop[k] == sn) {
                            yb[k] -= D; turned[k] = true;
                            if (ek_has<TS>(ek[k])) ek[k] = ((unsigned)((ek_time<TS>(ek[k]) + D) >> 1) << TS) | TY1 | (unsigned)(lane + GQW_LANES * k);
                            GQT_MAKE_OUTER(k);
                        }
                    }
                    if (sn < cap) sv = sn;
                    else {
#pragma unroll
                        for (int k = 0; k < K; ++k) nm[k] = __ballot_sync(GQW_FULL, turned[k]);
                        have_masks = true;
                    }
                    __syncwarp();
                } else if (type == 3) {
                    // ---- an outer vertex reaches the boundary: the tree augments onto it
                    GQM_STAT(4, 1); GQM_STAT(10 + (ph_events > 4 ? 4 : ph_events), 1);
                    if (lane == 0) gqr::augment_half(f, arg, -2);
                    break;
                } else if (type == 1) {
                    GQM_STAT(3, 1);
                    if (--guard < 0) { fail = 3; break; }
                    // ---- odd cycle: shrink
                    const int v = arg, u = src;
#pragma unroll 1
                    for (int x = lane; x < cap + bhi; x += GQW_LANES) f.mark[x] = 0;
                    __syncwarp();
                    int b = 0;
                    if (lane == 0) {
                        b = gqr::shrink(f, u, v);
                        if (b >= 0) {
                            // children that are blossoms freeze their dual at its current value
                            int c = f.first[b - cap];
                            const int c0 = c;
                            do {
                                if (c >= cap) f.yb[c - cap] += f.label[c] == L_S ? 2 * D : -2 * D;
                                c = f.nxt[c];
                            } while (c != c0);
                            f.yb[b - cap] = -2 * D;   // z = zb + 2D = 0 now
                        }
                    }
                    b = __shfl_sync(GQW_FULL, b, 0);
                    if (b < 0) { fail = 2; break; }
                    bhi = max(bhi, b - cap + 1);
                    __syncwarp();
                    bool turned[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int x = lane + GQW_LANES * k;
                        turned[k] = false;
                        if (x < n && f.par[vtop[k]] == b) {
                            if (GQT_IS_INNER(k)) {
                                turned[k] = true;
                                yb[k] -= 2 * D;
                                ek[k] = ek_has<TS>(ek[k]) ? (((unsigned)(D + (ek_time<TS>(ek[k]) >> 1)) << TS) | TY1 | (unsigned)x) : EK_NONE;
                                GQT_MAKE_OUTER(k);
                            }
                            vtop[k] = b;
                            f.top[x] = (i16)b;
                        }
                    }
                    __syncwarp();
                    // an entry whose source now sits in the same blossom is void: drop it and offer the vertex's
                    // row again (every edge to another node is then held by the far end at least), so that no
                    // stale entry can ever win the search
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        if (vtop[k] == b && slfrom[k] >= 0 && ek_has<TS>(ek[k]) && f.top[slfrom[k]] == b) {
                            ek[k] = EK_NONE; slfrom[k] = -1; turned[k] = true;
                        }
                        nm[k] = __ballot_sync(GQW_FULL, turned[k]);
                    }
                    have_masks = true;
                    __syncwarp();
                } else {
                    GQM_STAT(7, 1);
                    if (--guard < 0) { fail = 3; break; }
                    // ---- an inner blossom's dual reached zero: expand (its children may be inner blossoms)
                    const int b = cap + arg;
                    inner_blossoms = true;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int x = lane + GQW_LANES * k;
                        if (vtop[k] == b) {
                            int c = x;
#pragma unroll 1
                            while (f.par[c] != b) c = f.par[c];
                            vtop[k] = c;
                            f.top[x] = (i16)c;
                        }
                    }
                    __syncwarp();
                    if (lane == 0) {
                        const int fst = f.first[b - cap];
                        gqr::expand_relabel(f, b);
                        int c = fst;
                        do {
                            if (c >= cap && f.label[c] != L_NONE) f.yb[c - cap] += f.label[c] == L_S ? -2 * D : 2 * D;
                            c = f.nxt[c];
                        } while (c != fst);
                    }
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int x = lane + GQW_LANES * k;
                        bool turned = false;
                        if (x < n && GQT_IS_INNER(k)) {
                            const int nl = f.label[vtop[k]];
                            if (nl == L_S) {
                                turned = true; yb[k] -= 2 * D;
                                ek[k] = ek_has<TS>(ek[k]) ? (((unsigned)(D + (ek_time<TS>(ek[k]) >> 1)) << TS) | TY1 | (unsigned)x) : EK_NONE;
                                GQT_MAKE_OUTER(k);
                            } else if (nl == L_NONE) {
                                yb[k] -= D;
                                ek[k] = ek_has<TS>(ek[k]) ? (((unsigned)(D + ek_time<TS>(ek[k])) << TS) | (unsigned)x) : EK_NONE;
                                kb[k] = (unsigned)x;
                            }
                        }
                        nm[k] = __ballot_sync(GQW_FULL, turned);
                    }
                    have_masks = true;
                }
            }
            if (fail) return fail;
            // ---- phase end: the tree dissolves, duals leave the clock
            __syncwarp();
            if (bhi) {
#pragma unroll 1
                for (int i = lane; i < bhi; i += GQW_LANES) {
                    if (f.bused[i] && f.par[cap + i] < 0 && f.label[cap + i] != L_NONE)
                        f.yb[i] += f.label[cap + i] == L_S ? 2 * D : -2 * D;
                }
                __syncwarp();
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int x = lane + GQW_LANES * k;
                if (kb[k] != (unsigned)x) {
                    yb[k] += GQT_IS_OUTER(k) ? D : -D;
                    kb[k] = (unsigned)x; bk[k] = 0xffffffffu;
                    f.label[vtop[k]] = L_NONE;
                }
            }
            __syncwarp();
        }
    }
    return 0;
#undef GQT_SEL
#undef GQT_IS_OUTER
#undef GQT_IS_INNER
#undef GQT_MAKE_OUTER
#undef GQT_SCAN
}

// ---- dual start of one shot (part A): nearest other defect of every defect, y(v) = min(nearest / 2,
// boundary), boundary matches and mutual-nearest-neighbour matches.
// In: defects def[0..n) (shared), synw / spre (nullable): the shot's syndrome words and the number of
// defects before each word (shared); xch: shared scratch of n int16.  Out, per slot of this lane: yb[] =
// dual (quadrupled units, even), m0[] = initial mate (-1 free, -2 boundary, partner index; -3 for padding).
template <int K>
__device__ __forceinline__ void dual_start(const Tables &T, i16 *xch, const u16 *def, const int n, const int lane,
                                           const uint32_t *synw, const u16 *spre, i32 (&yb)[K], int (&m0)[K]) {
    const u16 *__restrict__ dist = T.dist;
    const uint32_t N = T.N;
    const uint32_t bro = T.D * N;
    constexpr int VB = KeyBits<K>::VB;
    constexpr unsigned VM = KeyBits<K>::VM;
    uint32_t dj[K];
    unsigned key[K];
    bool miss[K];
    // ---- nearest other defect of every defect: key = distance << VB | index.
    // Lane-parallel: test the NBR_LEN nearest detectors of each own defect against the syndrome bits
    // (descending, so the nearest hit is the one that sticks); any hit is the true nearest defect because
    // the list is a complete prefix of the detector's distance order.
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int v = lane + GQW_LANES * k;
        dj[k] = v < n ? def[v] : T.D + 1;     // padding slots read the "nothing" column: all infinite
        key[k] = 0xffffffffu;
        miss[k] = v < n;
        if (synw != nullptr) {
            const uint4 *L = reinterpret_cast<const uint4 *>(T.nbr + (size_t)dj[k] * NBR_LEN);
            uint32_t hit = 0xffffffffu;
            uint4 le[NBR_LEN / 4];   // all list loads in flight before the first test (the start kernel is small
                                     // enough for the unrolled body; it is latency-bound)
#pragma unroll
            for (int c = 0; c < NBR_LEN / 4; ++c) le[c] = L[c];
#pragma unroll
            for (int c = NBR_LEN / 4 - 1; c >= 0; --c) {
                const uint4 e = le[c];
                const uint32_t ent[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                for (int q = 3; q >= 0; --q) {
                    const uint32_t det = ent[q] & 0xffffu;
                    hit = ((synw[det >> 5] >> (det & 31)) & 1u) ? ent[q] : hit;
                }
            }
            if (hit != 0xffffffffu) {
                const uint32_t det = hit & 0xffffu;
                const uint32_t rank = (uint32_t)spre[det >> 5] + (uint32_t)__popc(synw[det >> 5] & ((1u << (det & 31)) - 1u));
                key[k] = ((hit >> 16) << VB) | rank;
                miss[k] = false;
            }
        }
    }
    // the rest (no fired detector among the listed ones; every defect when the syndrome is not mirrored):
    // one row of the distance table per defect, all lanes gather, the minimum goes to its owner
    {
        unsigned mm[K];
        unsigned any = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) { mm[k] = __ballot_sync(GQW_FULL, miss[k]); any |= mm[k]; }
        if (any) {
#pragma unroll 1
            for (int kk = 0; kk < K; ++kk) {
                unsigned m = (unsigned)sel_slot<K>(mm, kk);
#pragma unroll 1
                while (m) {
                    const int i = GQW_LANES * kk + __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t ro = (uint32_t)def[i] * N;
                    unsigned best = 0xffffffffu;
#pragma unroll
                    for (int k = 0; k < K; ++k) best = min(best, ((unsigned)dist[ro + dj[k]] << VB) + (unsigned)(lane + GQW_LANES * k));
                    best = __reduce_min_sync(GQW_FULL, best);
#pragma unroll
                    for (int k = 0; k < K; ++k) if (lane + GQW_LANES * k == i) key[k] = best;
                }
            }
        }
    }
    // ---- dual start, boundary matches, mutual nearest neighbours
    bool onb[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int v = lane + GQW_LANES * k;
        const uint32_t bb = dist[bro + dj[k]];
        const i32 b4 = bb == (uint32_t)W_INF ? SL_INF : 4 * (i32)bb;
        const uint32_t wmin = key[k] >> VB;
        const i32 y2 = wmin >= (uint32_t)W_INF ? SL_INF : 2 * (i32)wmin;
        onb[k] = b4 < SL_INF && b4 <= y2;
        yb[k] = onb[k] ? b4 : (y2 < SL_INF ? y2 : 0);
        if (v < n) xch[v] = (i16)((onb[k] || y2 == SL_INF) ? -1 : (int)(key[k] & VM));
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int v = lane + GQW_LANES * k;
        m0[k] = -3;
        if (v < n) {
            const int u = xch[v];
            int m = onb[k] ? -2 : -1;
            if (u >= 0 && xch[u] == v) m = u;
            m0[k] = m;
        }
    }
    __syncwarp();
}

// ---- solve and predict (part B).  In: defects def[0..n) (shared), f.mate[0..n) = the initial matching,
// yb[] = the lane's duals.  Out: prediction mask, weight, f.mate[] = the minimum-weight matching.
// Returns the solver's code (0 ok).
template <int K, class FT>
__device__ __forceinline__ int solve_predict(const Tables &T, FT &f, const u16 *def, const int n, const int lane,
                                             const bool want_weight, i32 (&yb)[K], uint32_t *pm_out, int *wt_out) {
    const u16 *__restrict__ dist = T.dist;
    const uint32_t N = T.N;
    const uint32_t bro = T.D * N;
    uint32_t dj[K];
    i32 b4[K];
    gqw_mask freebits = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int v = lane + GQW_LANES * k;
        dj[k] = v < n ? def[v] : T.D + 1;
        const uint32_t bb = dist[bro + dj[k]];
        b4[k] = bb == (uint32_t)W_INF ? SL_INF : 4 * (i32)bb;
        if (v < n && f.mate[v] == -1) freebits |= GQW_BIT(k);
    }
#ifdef GQ_MT_STATS
    {
        int nf = 0;
        for (int k = 0; k < K; ++k) nf += (int)((freebits >> k) & 1);
        GQM_STAT(8, __reduce_add_sync(GQW_FULL, nf));
        GQM_STAT(9, 1);
    }
#endif
    const int any_free = __any_sync(GQW_FULL, freebits != 0);
    int rc = 0;
    if (any_free) rc = solve_tree<K>(f, dist, N, def, n, lane, dj, b4, yb, freebits);
    __syncwarp();
    uint32_t pm = 0;
    int wt = 0, bad = 0;
    if (rc == 0) {
        // one gather per slot, no divergent branches: the boundary is row D of the path-observable table
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int v = lane + GQW_LANES * k;
            const int j = v < n ? (int)f.mate[v] : -3;
            const bool onb = j == -2, pair = j > v;     // a pair is counted by its lower end
            bad |= (j == -1) | (onb & (b4[k] == SL_INF));
            const uint32_t prow = onb ? T.D : (uint32_t)def[pair ? j : 0];
            const uint32_t at = prow * N + dj[k];
            if (onb | pair) pm ^= T.pobs[at];
            if (onb) wt += b4[k] >> 2;
            if (want_weight && pair) {
                const uint32_t d = dist[at];
                if (d >= (uint32_t)W_INF) bad = 1;
                wt += (int)d;
            }
        }
        pm = __reduce_xor_sync(GQW_FULL, pm);
        wt = __reduce_add_sync(GQW_FULL, wt);
        if (__any_sync(GQW_FULL, bad)) rc = 1;
    }
    *pm_out = pm;
    *wt_out = wt;
    return rc;
}

// ---- both parts in one call (host simulation and the fused use)
template <int K, class FT>
__device__ __forceinline__ int decode_shot(const Tables &T, FT &f, const u16 *def, const int n,
                                           const int lane, const bool want_weight, uint32_t *pm_out, int *wt_out,
                                           const uint32_t *synw = nullptr, const u16 *spre = nullptr) {
    i32 yb[K];
    int m0[K];
    dual_start<K>(T, f.stk, def, n, lane, synw, spre, yb, m0);
#pragma unroll
    for (int k = 0; k < K; ++k) if (m0[k] != -3) f.mate[lane + GQW_LANES * k] = (i16)m0[k];
    __syncwarp();
    return solve_predict<K>(T, f, def, n, lane, want_weight, yb, pm_out, wt_out);
}

}  // namespace gqt
