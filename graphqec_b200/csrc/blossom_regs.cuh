// Register-resident variant of the warp-cooperative blossom solver (fast tier, n <= 32*K).
//
// Same algorithm and invariants as blossom_warp.cuh (see there), different data placement:
//  * vertex v lives in slot (v >> 5) of lane (v & 31): its dual, best slack, slack source, top-level
//    node and that node's label are REGISTERS, so the O(n) sweeps (slack scan of a new outer vertex,
//    minimum-slack search, dual update) are a few ALU ops per slot plus one shuffle reduction;
//  * the blossom forest (mate, top, base, parent/sibling links, tree edges), which lane 0 walks
//    when it augments / shrinks / expands, lives in shared memory;
//  * the dense weight matrix W stays in an L2-resident scratch, read one coalesced row per scan.
// Two algorithmic savings over the memory-resident version: slacks of outer vertices are validated
// lazily (a stale entry -- source swallowed by the same blossom -- is a lower bound, so it is only
// recomputed if it wins the minimum search) instead of rebuilding all slacks after every shrink or
// expand; and label clearing / marking are lane-parallel.
//
// Everything is a free function over local arrays indexed by unrolled constants: keeping the
// per-lane state out of any object whose address is taken is what keeps it in registers.
#pragma once
#include <stdint.h>

#ifdef GQ_STATS
#define GQ_STAT(i) do { if (lane == 0) atomicAdd(&gq_stats[i], 1ull); } while (0)
#else
#define GQ_STAT(i) do { } while (0)
#endif

namespace gqr {

typedef int16_t i16;
typedef int32_t i32;
typedef uint16_t u16;

static const i32 SL_INF = 0x3fffffff;
static const u16 W_INF = 0xffff;
enum { L_NONE = 0, L_S = 1, L_T = 2 };

// shared-memory forest, capacity cap real vertices (even), nb = cap/2 blossoms, nn = cap + nb nodes
struct Forest {
    i16 *mate, *top;                                  // [cap]
    i16 *base, *par, *nxt, *prv, *eu, *ev, *pu, *pv;  // [nn]
    i16 *first;                                       // [nb]
    i16 *troot;                                       // [nn] root vertex of the tree a labelled node is in
    i16 *stk;                                         // [2*nn]
    i32 *yb;                                          // [nb] blossom duals
    int8_t *label, *mark;                             // [nn]
    int8_t *bused;                                    // [nb]
    int cap, nb;
};

__host__ __device__ inline size_t forest_bytes(int cap) {
    int nb = cap / 2, nn = cap + nb;
    size_t b = (size_t)cap * 2 * 2 + (size_t)nn * 2 * 9 + (size_t)nb * 2 + (size_t)nn * 2 * 2 + (size_t)nb * 4 +
               (size_t)nn * 2 + nb;
    return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ Forest forest_bind(char *p, int cap) {
    Forest f;
    int nb = cap / 2, nn = cap + nb;
    f.cap = cap; f.nb = nb;
    f.yb = (i32 *)p; p += (size_t)nb * 4;
    f.mate = (i16 *)p; p += cap * 2;
    f.top = (i16 *)p; p += cap * 2;
    f.base = (i16 *)p; p += nn * 2;
    f.par = (i16 *)p; p += nn * 2;
    f.nxt = (i16 *)p; p += nn * 2;
    f.prv = (i16 *)p; p += nn * 2;
    f.eu = (i16 *)p; p += nn * 2;
    f.ev = (i16 *)p; p += nn * 2;
    f.pu = (i16 *)p; p += nn * 2;
    f.pv = (i16 *)p; p += nn * 2;
    f.first = (i16 *)p; p += nb * 2;
    f.troot = (i16 *)p; p += nn * 2;
    f.stk = (i16 *)p; p += nn * 4;
    f.label = (int8_t *)p; p += nn;
    f.mark = (int8_t *)p; p += nn;
    f.bused = (int8_t *)p;
    return f;
}

// Same forest with every array at a compile-time offset from one base pointer (no pointer
// registers): the per-warp shared-memory block of the first tier (match_tree.cuh).
template <int CAP>
struct StaticForest {
    static constexpr int cap = CAP, nb = CAP / 2, nn = CAP + CAP / 2;
    i32 yb[nb];
    i16 mate[cap], top[cap];
    i16 base[nn], par[nn], nxt[nn], prv[nn], eu[nn], ev[nn], pu[nn], pv[nn];
    i16 first[nb];
    i16 stk[2 * nn];
    int8_t label[nn], mark[nn];
    int8_t bused[nb];
};

// ---------------------------------------------------------------- lane 0 forest surgery
// augment / augment_half run once per phase and stay inline (out of line they cost more in spills around
// the call than they save).  rebase, shrink and expand_relabel only run when a blossom is involved (about
// twice per shot at the north star) but are large -- rebase alone was inlined five times: out of line
// they take ~1 k instructions out of the event loop's footprint, which matters because the kernels sit at
// the edge of the instruction cache (sm__icc_request_hit_rate, profiles/).
#ifndef GQ_SURGERY_INLINE
#define GQ_SURGERY_INLINE __forceinline__
#endif
#ifndef GQ_RARE_INLINE
#define GQ_RARE_INLINE __noinline__
#endif
// (FT = Forest or StaticForest<CAP>)
template <class FT>
__device__ GQ_RARE_INLINE void rebase(FT &f, int x0, int a0) {
    const int cap = f.cap;
    int sp = 0;
    f.stk[sp++] = (i16)x0; f.stk[sp++] = (i16)a0;
#pragma unroll 1
    while (sp) {
        int a = f.stk[--sp];
        int B = f.stk[--sp];
        if (B < cap || f.base[B] == a) continue;
        int c = a;
#pragma unroll 1
        while (f.par[c] != B) c = f.par[c];
        f.stk[sp++] = (i16)c; f.stk[sp++] = (i16)a;
        const int fst = f.first[B - cap];
        int steps = 0;
#pragma unroll 1
        for (int t = c; t != fst; t = f.prv[t]) ++steps;
        int p = c;
        if ((steps & 1) == 0) {
#pragma unroll 1
            while (p != fst) {
                int q1 = f.prv[p], q2 = f.prv[q1];
                int a2 = f.eu[q2], a1 = f.ev[q2];
                f.mate[a2] = (i16)a1; f.mate[a1] = (i16)a2;
                f.stk[sp++] = (i16)q2; f.stk[sp++] = (i16)a2;
                f.stk[sp++] = (i16)q1; f.stk[sp++] = (i16)a1;
                p = q2;
            }
        } else {
#pragma unroll 1
            while (p != fst) {
                int q1 = f.nxt[p], q2 = f.nxt[q1];
                int a1 = f.eu[q1], a2 = f.ev[q1];
                f.mate[a1] = (i16)a2; f.mate[a2] = (i16)a1;
                f.stk[sp++] = (i16)q1; f.stk[sp++] = (i16)a1;
                f.stk[sp++] = (i16)q2; f.stk[sp++] = (i16)a2;
                p = q2;
            }
        }
        f.first[B - cap] = (i16)c;
        f.base[B] = (i16)a;
    }
}

template <class FT>
__device__ GQ_SURGERY_INLINE void augment(FT &f, int u, int v) {
    int a = u, nm = v;
#pragma unroll 1
    for (;;) {
        int x = f.top[a];
        int m = f.mate[f.base[x]];
        if (x >= f.cap) rebase(f, x, a);   // plain vertices need no rebasing (the common case)
        f.mate[a] = (i16)nm;
        if (nm >= 0) f.mate[nm] = (i16)a;
        if (m < 0) break;
        int t = f.top[m];
        int tpu = f.pu[t], tpv = f.pv[t];
        if (t >= f.cap) rebase(f, t, tpv);
        nm = tpv; a = tpu;
    }
}

// one side of an augmentation between two trees: flip the path from outer vertex u to its root and
// point u at its new mate nm (the caller's other half sets mate[nm])
template <class FT>
__device__ GQ_SURGERY_INLINE void augment_half(FT &f, int u, int nm0) {
    int a = u, nm = nm0, first = 1;
#pragma unroll 1
    for (;;) {
        int x = f.top[a];
        int m = f.mate[f.base[x]];
        if (x >= f.cap) rebase(f, x, a);   // plain vertices need no rebasing (the common case)
        f.mate[a] = (i16)nm;
        if (!first) f.mate[nm] = (i16)a;
        first = 0;
        if (m < 0) break;
        int t = f.top[m];
        int tpu = f.pu[t], tpv = f.pv[t];
        if (t >= f.cap) rebase(f, t, tpv);
        nm = tpv; a = tpu;
    }
}

template <class FT>
__device__ __forceinline__ int parent_S(FT &f, int x, int *tnode) {
    int m = f.mate[f.base[x]];
    if (m < 0) return -1;
    int t = f.top[m];
    *tnode = t;
    return f.top[f.pu[t]];
}

// mark[] must be zero on entry
template <class FT>
__device__ GQ_RARE_INLINE int shrink(FT &f, int u, int v) {
    const int cap = f.cap, nb = f.nb;
    int xu = f.top[u], xv = f.top[v];
#pragma unroll 1
    for (int x = xu; x >= 0;) { f.mark[x] = 1; int t; x = parent_S(f, x, &t); }
    int lca = xv;
#pragma unroll 1
    while (!f.mark[lca]) { int t; lca = parent_S(f, lca, &t); }
    int b = -1;
#pragma unroll 1
    for (int i = 0; i < nb; ++i) if (!f.bused[i]) { b = cap + i; f.bused[i] = 1; break; }
    if (b < 0) return -1;
    f.nxt[lca] = (i16)lca; f.prv[lca] = (i16)lca;
    {
        int have = 0, bu = 0, bv = 0, first_iter = 1;
        int x = xu;
#pragma unroll 1
        while (x != lca) {
            int t; int up = parent_S(f, x, &t);
            int old = f.nxt[lca];
            f.nxt[lca] = (i16)x; f.prv[x] = (i16)lca; f.nxt[x] = (i16)old; f.prv[old] = (i16)x;
            if (!first_iter) { f.eu[x] = (i16)bu; f.ev[x] = (i16)bv; }
            first_iter = 0;
            old = f.nxt[lca];
            f.nxt[lca] = (i16)t; f.prv[t] = (i16)lca; f.nxt[t] = (i16)old; f.prv[old] = (i16)t;
            f.eu[t] = f.mate[f.base[x]]; f.ev[t] = f.base[x];
            bu = f.pu[t]; bv = f.pv[t];
            have = 1;
            x = up;
        }
        if (have) { f.eu[lca] = (i16)bu; f.ev[lca] = (i16)bv; }
    }
    {
        int tail = f.prv[lca];
        f.eu[tail] = (i16)u; f.ev[tail] = (i16)v;
        int x = xv;
#pragma unroll 1
        while (x != lca) {
            int t; int up = parent_S(f, x, &t);
            f.nxt[tail] = (i16)x; f.prv[x] = (i16)tail;
            f.eu[x] = f.base[x]; f.ev[x] = f.mate[f.base[x]];
            f.nxt[x] = (i16)t; f.prv[t] = (i16)x; f.nxt[t] = (i16)lca; f.prv[lca] = (i16)t;
            f.eu[t] = f.pv[t]; f.ev[t] = f.pu[t];
            tail = t;
            x = up;
        }
    }
    f.first[b - cap] = (i16)lca;
    f.base[b] = f.base[lca];
    f.par[b] = -1;
    f.yb[b - cap] = 0;
    f.label[b] = L_S;
    int c = lca;
#pragma unroll 1
    do { f.par[c] = (i16)b; c = f.nxt[c]; } while (c != lca);
    return b;
}

// relabel the children of T blossom b (top[] of its vertices already points at them)
template <class FT>
__device__ GQ_RARE_INLINE void expand_relabel(FT &f, int b) {
    const int cap = f.cap;
    const int fst = f.first[b - cap];
    const int c = f.top[f.pv[b]];
    int steps = 0;
#pragma unroll 1
    for (int t = c; t != fst; t = f.prv[t]) ++steps;
    int x = fst;
#pragma unroll 1
    do { f.label[x] = L_NONE; int nx = f.nxt[x]; f.par[x] = -1; x = nx; } while (x != fst);
    f.label[c] = L_T; f.pu[c] = f.pu[b]; f.pv[c] = f.pv[b];
    int p = c;
    if ((steps & 1) == 0) {
#pragma unroll 1
        while (p != fst) {
            int s = f.prv[p], t = f.prv[s];
            f.label[s] = L_S;
            f.label[t] = L_T; f.pu[t] = f.ev[t]; f.pv[t] = f.eu[t];
            p = t;
        }
    } else {
#pragma unroll 1
        while (p != fst) {
            int s = f.nxt[p], t = f.nxt[s];
            f.label[s] = L_S;
            f.label[t] = L_T; f.pu[t] = f.eu[s]; f.pv[t] = f.ev[s];
            p = t;
        }
    }
    f.bused[b - cap] = 0;
    f.label[b] = L_NONE;
}

// ---------------------------------------------------------------- main
// yinit[k] = minimum weight incident to vertex lane + 32k (undoubled); n even; W row stride wstride.
// Returns 0 ok, 1 no perfect matching, 2 blossom ids exhausted, 3 step guard.  mate[] ends in f.mate.
template <int K>
__device__ __forceinline__ int solve(const Forest &f, const u16 *W, const int wstride, const int n,
                                     const int lane, const uint32_t (&yinit)[K]) {
    const int cap = f.cap, nb = f.nb, nn = cap + nb;
    const int kmax = (n + 31) >> 5;   // slots in use (uniform)
    int bhi = 0;                      // blossom ids in use are below cap + bhi
    i32 y[K], sl[K];
    int slfrom[K], vtop[K], vlab[K];

#define GQR_SEL(arr, u, out)                                           \
    {                                                                  \
        int sel_ = arr[0];                                             \
        _Pragma("unroll") for (int k_ = 1; k_ < K; ++k_) if (((u) >> 5) == k_) sel_ = arr[k_]; \
        out = __shfl_sync(0xffffffffu, sel_, (u) & 31);                \
    }
#define GQR_SCAN(u_)                                                   \
    {                                                                  \
        const int su_ = (u_);                                          \
        const int tu_ = f.top[su_];                                    \
        i32 yu_;                                                       \
        GQR_SEL(y, su_, yu_);                                          \
        const u16 *row_ = W + su_ * wstride;                           \
        _Pragma("unroll") for (int k_ = 0; k_ < K; ++k_) {             \
            const int v_ = lane + 32 * k_;                             \
            if (k_ < kmax && v_ < n && vtop[k_] != tu_) {              \
                u16 wt_ = row_[v_];                                    \
                if (wt_ != W_INF) {                                    \
                    i32 s_ = 2 * (i32)wt_ - yu_ - y[k_];               \
                    if (s_ < sl[k_]) { sl[k_] = s_; slfrom[k_] = su_; } \
                }                                                      \
            }                                                          \
        }                                                              \
    }
#define GQR_SCAN_WHERE(pred)                                           \
    {                                                                  \
        _Pragma("unroll") for (int kk_ = 0; kk_ < K; ++kk_) {          \
            unsigned m_ = kk_ < kmax ? __ballot_sync(0xffffffffu, pred[kk_]) : 0u; \
            while (m_) {                                               \
                int b_ = __ffs(m_) - 1;                                \
                m_ &= m_ - 1;                                          \
                GQ_STAT(6); GQR_SCAN(32 * kk_ + b_);                   \
            }                                                          \
        }                                                              \
    }

    for (int x = lane; x < nn; x += 32) { f.base[x] = (i16)(x < cap ? x : -1); f.par[x] = -1; f.label[x] = L_NONE; f.mark[x] = 0; }
    for (int i = lane; i < nb; i += 32) f.bused[i] = 0;
    unsigned freebits = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int v = lane + 32 * k;
        vtop[k] = v; vlab[k] = L_NONE; sl[k] = SL_INF; slfrom[k] = -1;
        y[k] = yinit[k] == (uint32_t)W_INF ? 0 : (i32)yinit[k];
        if (v < n) { f.mate[v] = -1; f.top[v] = (i16)v; freebits |= 1u << k; }
    }
    __syncwarp();
    // ---- greedy matching along tight edges
    for (int u = 0; u < n; ++u) {
        unsigned ufree = (__shfl_sync(0xffffffffu, freebits, u & 31) >> (u >> 5)) & 1u;
        if (!ufree) continue;
        i32 yu;
        GQR_SEL(y, u, yu);
        const u16 *row = W + u * wstride;
        int found = -1;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int v = lane + 32 * k;
            bool ok = false;
            if (k < kmax && found < 0 && v < n && v != u && ((freebits >> k) & 1u)) {
                u16 wt = row[v];
                ok = wt != W_INF && 2 * (i32)wt - yu - y[k] == 0;
            }
            unsigned m = __ballot_sync(0xffffffffu, ok);
            if (found < 0 && m) found = 32 * k + __ffs(m) - 1;
        }
        if (found >= 0) {
            if (lane == 0) { f.mate[u] = (i16)found; f.mate[found] = (i16)u; }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int v = lane + 32 * k;
                if (v == u || v == found) freebits &= ~(1u << k);
            }
        }
    }
    __syncwarp();
    GQ_STAT(7);
#ifdef GQ_STATS
    { int nf_ = __reduce_add_sync(0xffffffffu, __popc(freebits)); if (lane == 0) atomicAdd(&gq_stats[8], (unsigned long long)nf_); }
#endif
    // ---- phases: one alternating tree per free vertex
    for (int root = 0; root < n; ++root) {
        unsigned rfree = (__shfl_sync(0xffffffffu, freebits, root & 31) >> (root >> 5)) & 1u;
        if (!rfree) continue;
        for (int x = lane; x < nn; x += 32) f.label[x] = L_NONE;
#pragma unroll
        for (int k = 0; k < K; ++k) { vlab[k] = L_NONE; sl[k] = SL_INF; slfrom[k] = -1; }
        __syncwarp();
        if (lane == 0) f.label[root] = L_S;
#pragma unroll
        for (int k = 0; k < K; ++k) if (lane + 32 * k == root) vlab[k] = L_S;
        __syncwarp();
        GQR_SCAN(root);
        for (int guard = 0;; ++guard) {
            if (guard > 8 * nn + 64) return 3;
            // ---- minimum-slack event: key = delta << 10 | type << 8 | arg
            unsigned best = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int v = lane + 32 * k;
                if (k < kmax && v < n && sl[k] < SL_INF) {
                    if (vlab[k] == L_NONE) best = min(best, ((unsigned)sl[k] << 10) | (unsigned)v);
                    else if (vlab[k] == L_S) best = min(best, ((unsigned)(sl[k] >> 1) << 10) | 0x100u | (unsigned)v);
                }
            }
            for (int i = lane; i < bhi; i += 32)
                if (f.bused[i] && f.par[cap + i] < 0 && f.label[cap + i] == L_T)
                    best = min(best, ((unsigned)(f.yb[i] >> 1) << 10) | 0x200u | (unsigned)i);
            best = __reduce_min_sync(0xffffffffu, best);
            GQ_STAT(0);
            if (best == 0xffffffffu) return 1;
            const i32 delta = (i32)(best >> 10);
            const int type = (best >> 8) & 3;
            const int arg = best & 0xff;
            int src = -1;
            if (type != 2) GQR_SEL(slfrom, arg, src);
            if (type == 1 && f.top[src] == f.top[arg]) {
                GQ_STAT(1);
                // lazy validation: the source was swallowed by the same blossom -> recompute the exact
                // slack of outer vertex arg towards outer vertices of other nodes, then search again
                const int tv = f.top[arg];
                i32 yv;
                GQR_SEL(y, arg, yv);
                const u16 *row = W + arg * wstride;
                unsigned long long bb = ~0ull;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int u = lane + 32 * k;
                    if (k < kmax && u < n && vlab[k] == L_S && vtop[k] != tv) {
                        u16 wt = row[u];
                        if (wt != W_INF) {
                            i32 s = 2 * (i32)wt - yv - y[k];
                            unsigned long long key = ((unsigned long long)(unsigned)s << 16) | (unsigned)u;
                            if (key < bb) bb = key;
                        }
                    }
                }
                unsigned hi = (unsigned)(bb >> 32), lo = (unsigned)bb;
                unsigned mh = __reduce_min_sync(0xffffffffu, hi);
                unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
                bb = ((unsigned long long)mh << 32) | ml;
                const i32 ns = bb == ~0ull ? SL_INF : (i32)(bb >> 16);
                const int nf = bb == ~0ull ? -1 : (int)(bb & 0xffff);
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (lane + 32 * k == arg) { sl[k] = ns; slfrom[k] = nf; }
                continue;
            }
            // ---- dual update
            if (delta) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (k >= kmax) continue;
                    if (vlab[k] == L_S) { y[k] += delta; if (sl[k] < SL_INF) sl[k] -= 2 * delta; }
                    else if (vlab[k] == L_T) y[k] -= delta;
                    else if (sl[k] < SL_INF) sl[k] -= delta;
                }
                for (int i = lane; i < bhi; i += 32)
                    if (f.bused[i] && f.par[cap + i] < 0) {
                        if (f.label[cap + i] == L_S) f.yb[i] += 2 * delta;
                        else if (f.label[cap + i] == L_T) f.yb[i] -= 2 * delta;
                    }
                __syncwarp();
            }
            bool news[K];
            if (type == 0) {
                const int v = arg, u = src;
                const int t = f.top[v];
                const int m = f.mate[f.base[t]];
                if (m < 0) {
                    GQ_STAT(4);
                    if (lane == 0) augment(f, u, v);
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int x = lane + 32 * k;
                        if (x == root || x == v) freebits &= ~(1u << k);
                    }
                    break;
                }
                GQ_STAT(2);
                const int s = f.top[m];
                if (lane == 0) { f.label[t] = L_T; f.pu[t] = (i16)u; f.pv[t] = (i16)v; f.label[s] = L_S; }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    news[k] = false;
                    if (vtop[k] == t) vlab[k] = L_T;
                    else if (vtop[k] == s && lane + 32 * k < n) { vlab[k] = L_S; news[k] = true; }
                }
            } else if (type == 1) {
                const int v = arg, u = src;
                for (int x = lane; x < nn; x += 32) f.mark[x] = 0;
                __syncwarp();
                int b = 0;
                GQ_STAT(3);
                if (lane == 0) b = shrink(f, u, v);
                b = __shfl_sync(0xffffffffu, b, 0);
                if (b < 0) return 2;
                bhi = max(bhi, b - cap + 1);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    news[k] = false;
                    const int x = lane + 32 * k;
                    if (x < n && f.par[vtop[k]] == b) {
                        news[k] = vlab[k] == L_T;   // inner vertices turn outer: they must be offered
                        vtop[k] = b; vlab[k] = L_S;
                        f.top[x] = (i16)b;
                    }
                }
                __syncwarp();
            } else {
                GQ_STAT(5);
                const int b = cap + arg;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int x = lane + 32 * k;
                    if (x < n && vtop[k] == b) {
                        int c = x;
                        while (f.par[c] != b) c = f.par[c];
                        vtop[k] = c;
                        f.top[x] = (i16)c;
                    }
                }
                __syncwarp();
                if (lane == 0) expand_relabel(f, b);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int x = lane + 32 * k;
                    news[k] = false;
                    if (x < n) {
                        int nl = f.label[vtop[k]];
                        news[k] = nl == L_S && vlab[k] != L_S;
                        vlab[k] = nl;
                    }
                }
            }
            GQR_SCAN_WHERE(news);
        }
    }
    return 0;
#undef GQR_SEL
#undef GQR_SCAN
#undef GQR_SCAN_WHERE
}

// ---------------------------------------------------------------- native-boundary variant (default)
// The boundary is a special always-free partner: vertex v may be matched to it at cost b(v)
// (mate = -2).  W holds true distances d(i,j) only, there is no virtual vertex and n may be odd.
// An outer vertex whose boundary slack 2b - y reaches zero ends the phase (event type 3); reaching
// a vertex that is matched to the boundary also ends it (its boundary edge is an augmenting step).
// Compared with the min(d, b_i+b_j) form this halves the free vertices left after the greedy start
// (11 instead of 22 per shot at the north star) and keeps trees short near the boundary.
// yinit[k] = minimum weight incident to vertex lane + 32k (undoubled); n even; W row stride wstride.
// Returns 0 ok, 1 no perfect matching, 2 blossom ids exhausted, 3 step guard.  mate[] ends in f.mate.
template <int K>
__device__ __forceinline__ int solve_nb(const Forest &f, const u16 *W, const int wstride, const int n,
                                        const int lane, const uint32_t (&yinit)[K], const uint32_t (&bnd)[K]) {
    const int cap = f.cap, nb = f.nb, nn = cap + nb;
    const int kmax = (n + 31) >> 5;   // slots in use (uniform)
    int bhi = 0;                      // blossom ids in use are below cap + bhi
    i32 y[K], sl[K], b2[K];   // b2 = doubled boundary distance (SL_INF: none)
    int slfrom[K], vtop[K], vlab[K];

#define GQR_SEL(arr, u, out)                                           \
    {                                                                  \
        int sel_ = arr[0];                                             \
        _Pragma("unroll") for (int k_ = 1; k_ < K; ++k_) if (((u) >> 5) == k_) sel_ = arr[k_]; \
        out = __shfl_sync(0xffffffffu, sel_, (u) & 31);                \
    }
#define GQR_SCAN(u_)                                                   \
    {                                                                  \
        const int su_ = (u_);                                          \
        const int tu_ = f.top[su_];                                    \
        i32 yu_;                                                       \
        GQR_SEL(y, su_, yu_);                                          \
        const u16 *row_ = W + su_ * wstride;                           \
        _Pragma("unroll") for (int k_ = 0; k_ < K; ++k_) {             \
            const int v_ = lane + 32 * k_;                             \
            if (k_ < kmax && v_ < n && vtop[k_] != tu_) {              \
                u16 wt_ = row_[v_];                                    \
                if (wt_ != W_INF) {                                    \
                    i32 s_ = 2 * (i32)wt_ - yu_ - y[k_];               \
                    if (s_ < sl[k_]) { sl[k_] = s_; slfrom[k_] = su_; } \
                }                                                      \
            }                                                          \
        }                                                              \
    }
#define GQR_SCAN_WHERE(pred)                                           \
    {                                                                  \
        _Pragma("unroll") for (int kk_ = 0; kk_ < K; ++kk_) {          \
            unsigned m_ = kk_ < kmax ? __ballot_sync(0xffffffffu, pred[kk_]) : 0u; \
            while (m_) {                                               \
                int b_ = __ffs(m_) - 1;                                \
                m_ &= m_ - 1;                                          \
                GQ_STAT(6); GQR_SCAN(32 * kk_ + b_);                   \
            }                                                          \
        }                                                              \
    }

    for (int x = lane; x < nn; x += 32) { f.base[x] = (i16)(x < cap ? x : -1); f.par[x] = -1; f.label[x] = L_NONE; f.mark[x] = 0; }
    for (int i = lane; i < nb; i += 32) f.bused[i] = 0;
    unsigned freebits = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int v = lane + 32 * k;
        vtop[k] = v; vlab[k] = L_NONE; sl[k] = SL_INF; slfrom[k] = -1;
        b2[k] = bnd[k] == (uint32_t)W_INF ? SL_INF : 2 * (i32)bnd[k];
        y[k] = yinit[k] == (uint32_t)W_INF ? 0 : (i32)yinit[k];
        if (y[k] > b2[k] || yinit[k] == (uint32_t)W_INF) y[k] = b2[k] == SL_INF ? 0 : b2[k];
        if (v < n) {
            f.top[v] = (i16)v;
            if (y[k] == b2[k]) f.mate[v] = -2;            // tight boundary edge: start matched to it
            else { f.mate[v] = -1; freebits |= 1u << k; }
        }
    }
    __syncwarp();
    // ---- greedy matching along tight edges
    for (int u = 0; u < n; ++u) {
        unsigned ufree = (__shfl_sync(0xffffffffu, freebits, u & 31) >> (u >> 5)) & 1u;
        if (!ufree) continue;
        i32 yu;
        GQR_SEL(y, u, yu);
        const u16 *row = W + u * wstride;
        int found = -1;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int v = lane + 32 * k;
            bool ok = false;
            if (k < kmax && found < 0 && v < n && v != u && ((freebits >> k) & 1u)) {
                u16 wt = row[v];
                ok = wt != W_INF && 2 * (i32)wt - yu - y[k] == 0;
            }
            unsigned m = __ballot_sync(0xffffffffu, ok);
            if (found < 0 && m) found = 32 * k + __ffs(m) - 1;
        }
        if (found >= 0) {
            if (lane == 0) { f.mate[u] = (i16)found; f.mate[found] = (i16)u; }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int v = lane + 32 * k;
                if (v == u || v == found) freebits &= ~(1u << k);
            }
        }
    }
    __syncwarp();
    GQ_STAT(7);
#ifdef GQ_STATS
    { int nf_ = __reduce_add_sync(0xffffffffu, __popc(freebits)); if (lane == 0) atomicAdd(&gq_stats[8], (unsigned long long)nf_); }
#endif
    // ---- phases: one alternating tree per free vertex
    for (int root = 0; root < n; ++root) {
        unsigned rfree = (__shfl_sync(0xffffffffu, freebits, root & 31) >> (root >> 5)) & 1u;
        if (!rfree) continue;
        for (int x = lane; x < nn; x += 32) f.label[x] = L_NONE;
#pragma unroll
        for (int k = 0; k < K; ++k) { vlab[k] = L_NONE; sl[k] = SL_INF; slfrom[k] = -1; }
        __syncwarp();
        if (lane == 0) f.label[root] = L_S;
#pragma unroll
        for (int k = 0; k < K; ++k) if (lane + 32 * k == root) vlab[k] = L_S;
        __syncwarp();
        GQR_SCAN(root);
        for (int guard = 0;; ++guard) {
            if (guard > 8 * nn + 64) return 3;
            // ---- minimum-slack event: key = delta << 10 | type << 8 | arg
            unsigned best = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int v = lane + 32 * k;
                if (k < kmax && v < n && sl[k] < SL_INF) {
                    if (vlab[k] == L_NONE) best = min(best, ((unsigned)sl[k] << 10) | (unsigned)v);
                    else if (vlab[k] == L_S) best = min(best, ((unsigned)(sl[k] >> 1) << 10) | 0x100u | (unsigned)v);
                }
                if (k < kmax && v < n && vlab[k] == L_S && b2[k] < SL_INF)
                    best = min(best, ((unsigned)(b2[k] - y[k]) << 10) | 0x300u | (unsigned)v);
            }
            for (int i = lane; i < bhi; i += 32)
                if (f.bused[i] && f.par[cap + i] < 0 && f.label[cap + i] == L_T)
                    best = min(best, ((unsigned)(f.yb[i] >> 1) << 10) | 0x200u | (unsigned)i);
            best = __reduce_min_sync(0xffffffffu, best);
            GQ_STAT(0);
            if (best == 0xffffffffu) return 1;
            const i32 delta = (i32)(best >> 10);
            const int type = (best >> 8) & 3;
            const int arg = best & 0xff;
            int src = -1;
            if (type < 2) GQR_SEL(slfrom, arg, src);
            if (type == 1 && f.top[src] == f.top[arg]) {
                GQ_STAT(1);
                // lazy validation: the source was swallowed by the same blossom -> recompute the exact
                // slack of outer vertex arg towards outer vertices of other nodes, then search again
                const int tv = f.top[arg];
                i32 yv;
                GQR_SEL(y, arg, yv);
                const u16 *row = W + arg * wstride;
                unsigned long long bb = ~0ull;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int u = lane + 32 * k;
                    if (k < kmax && u < n && vlab[k] == L_S && vtop[k] != tv) {
                        u16 wt = row[u];
                        if (wt != W_INF) {
                            i32 s = 2 * (i32)wt - yv - y[k];
                            unsigned long long key = ((unsigned long long)(unsigned)s << 16) | (unsigned)u;
                            if (key < bb) bb = key;
                        }
                    }
                }
                unsigned hi = (unsigned)(bb >> 32), lo = (unsigned)bb;
                unsigned mh = __reduce_min_sync(0xffffffffu, hi);
                unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
                bb = ((unsigned long long)mh << 32) | ml;
                const i32 ns = bb == ~0ull ? SL_INF : (i32)(bb >> 16);
                const int nf = bb == ~0ull ? -1 : (int)(bb & 0xffff);
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (lane + 32 * k == arg) { sl[k] = ns; slfrom[k] = nf; }
                continue;
            }
            // ---- dual update
            if (delta) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (k >= kmax) continue;
                    if (vlab[k] == L_S) { y[k] += delta; if (sl[k] < SL_INF) sl[k] -= 2 * delta; }
                    else if (vlab[k] == L_T) y[k] -= delta;
                    else if (sl[k] < SL_INF) sl[k] -= delta;
                }
                for (int i = lane; i < bhi; i += 32)
                    if (f.bused[i] && f.par[cap + i] < 0) {
                        if (f.label[cap + i] == L_S) f.yb[i] += 2 * delta;
                        else if (f.label[cap + i] == L_T) f.yb[i] -= 2 * delta;
                    }
                __syncwarp();
            }
            bool news[K];
            if (type == 0) {
                const int v = arg, u = src;
                const int t = f.top[v];
                const int m = f.mate[f.base[t]];
                if (m < 0) {
                    GQ_STAT(4);
                    if (lane == 0) { rebase(f, t, v); augment(f, u, v); }
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int x = lane + 32 * k;
                        if (x == root || x == v) freebits &= ~(1u << k);
                    }
                    break;
                }
                GQ_STAT(2);
                const int s = f.top[m];
                if (lane == 0) { f.label[t] = L_T; f.pu[t] = (i16)u; f.pv[t] = (i16)v; f.label[s] = L_S; }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    news[k] = false;
                    if (vtop[k] == t) vlab[k] = L_T;
                    else if (vtop[k] == s && lane + 32 * k < n) { vlab[k] = L_S; news[k] = true; }
                }
            } else if (type == 3) {
                // ---- outer vertex arg reaches the boundary: match it there, flip the path to the root
                GQ_STAT(4);
                if (lane == 0) augment(f, arg, -2);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (lane + 32 * k == root) freebits &= ~(1u << k);
                break;
            } else if (type == 1) {
                const int v = arg, u = src;
                for (int x = lane; x < nn; x += 32) f.mark[x] = 0;
                __syncwarp();
                int b = 0;
                GQ_STAT(3);
                if (lane == 0) b = shrink(f, u, v);
                b = __shfl_sync(0xffffffffu, b, 0);
                if (b < 0) return 2;
                bhi = max(bhi, b - cap + 1);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    news[k] = false;
                    const int x = lane + 32 * k;
                    if (x < n && f.par[vtop[k]] == b) {
                        news[k] = vlab[k] == L_T;   // inner vertices turn outer: they must be offered
                        vtop[k] = b; vlab[k] = L_S;
                        f.top[x] = (i16)b;
                    }
                }
                __syncwarp();
            } else {
                GQ_STAT(5);
                const int b = cap + arg;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int x = lane + 32 * k;
                    if (x < n && vtop[k] == b) {
                        int c = x;
                        while (f.par[c] != b) c = f.par[c];
                        vtop[k] = c;
                        f.top[x] = (i16)c;
                    }
                }
                __syncwarp();
                if (lane == 0) expand_relabel(f, b);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int x = lane + 32 * k;
                    news[k] = false;
                    if (x < n) {
                        int nl = f.label[vtop[k]];
                        news[k] = nl == L_S && vlab[k] != L_S;
                        vlab[k] = nl;
                    }
                }
            }
            GQR_SCAN_WHERE(news);
        }
    }
    return 0;
#undef GQR_SEL
#undef GQR_SCAN
#undef GQR_SCAN_WHERE
}

// ---------------------------------------------------------------- multi-tree variant
// All free vertices grow alternating trees at once (one global dual clock D); an S-S edge between two
// trees augments and dissolves just those two trees, every other tree keeps its labels and slacks.
// Duals and slacks are stored against the clock, so a dual step costs nothing:
//   vertex dual    y = yb + sg*D            (sg = +1 outer, -1 inner, 0 unlabelled)
//   blossom dual   z = zb + 2*sg*D
//   tm = absolute time at which the best edge into the vertex turns tight (unlabelled: D0+slack,
//        outer: D0+slack/2) or, for an inner vertex, its frozen slack.
// Units are quadrupled (y_u + y_v <= 4w, initial duals even) so that every event time is an integer
// even across trees.  Entries are validated when they win the search (source still outer, other
// node, edge tight at that time); anything else is stale -- a lower bound -- and is recomputed.
template <int K>
__device__ __forceinline__ int solve_mt(const Forest &f, const u16 *W, const int wstride, const int n,
                                        const int lane, const uint32_t (&yinit)[K]) {
    const int cap = f.cap, nb = f.nb, nn = cap + nb;
    const int kmax = (n + 31) >> 5;
    i32 yb[K], tm[K];
    int slfrom[K], vtop[K], vlab[K];
    i32 D = 0;
    int bhi = 0;

#define GQM_YACT(k_, t_) (yb[k_] + (vlab[k_] == L_S ? (t_) : (vlab[k_] == L_T ? -(t_) : 0)))
#define GQM_SEL(arr, u, out)                                           \
    {                                                                  \
        int sel_ = arr[0];                                             \
        _Pragma("unroll") for (int k_ = 1; k_ < K; ++k_) if (((u) >> 5) == k_) sel_ = arr[k_]; \
        out = __shfl_sync(0xffffffffu, sel_, (u) & 31);                \
    }
#define GQM_YOF(u, t_, out)                                            \
    {                                                                  \
        int sel_ = GQM_YACT(0, t_);                                    \
        _Pragma("unroll") for (int k_ = 1; k_ < K; ++k_) if (((u) >> 5) == k_) sel_ = GQM_YACT(k_, t_); \
        out = __shfl_sync(0xffffffffu, sel_, (u) & 31);                \
    }
#define GQM_SCAN(u_)                                                   \
    {                                                                  \
        const int su_ = (u_);                                          \
        const int tu_ = f.top[su_];                                    \
        i32 yu_;                                                       \
        GQM_YOF(su_, D, yu_);                                          \
        const u16 *row_ = W + su_ * wstride;                           \
        _Pragma("unroll") for (int k_ = 0; k_ < K; ++k_) {             \
            if (k_ < kmax) {                                           \
                const int v_ = lane + 32 * k_;                         \
                if (v_ < n && vtop[k_] != tu_) {                       \
                    u16 wt_ = row_[v_];                                \
                    if (wt_ != W_INF) {                                \
                        i32 s_ = 4 * (i32)wt_ - yu_ - GQM_YACT(k_, D); \
                        i32 c_ = vlab[k_] == L_NONE ? D + s_ : (vlab[k_] == L_S ? D + (s_ >> 1) : s_); \
                        if (c_ < tm[k_]) { tm[k_] = c_; slfrom[k_] = su_; } \
                    }                                                  \
                }                                                      \
            }                                                          \
        }                                                              \
    }
#define GQM_SCAN_WHERE(pred)                                           \
    {                                                                  \
        _Pragma("unroll") for (int kk_ = 0; kk_ < K; ++kk_) {          \
            if (kk_ < kmax) {                                          \
                unsigned m_ = __ballot_sync(0xffffffffu, pred[kk_]);   \
                while (m_) {                                           \
                    int b_ = __ffs(m_) - 1;                            \
                    m_ &= m_ - 1;                                      \
                    GQ_STAT(6); GQM_SCAN(32 * kk_ + b_);               \
                }                                                      \
            }                                                          \
        }                                                              \
    }

    for (int x = lane; x < nn; x += 32) { f.base[x] = (i16)(x < cap ? x : -1); f.par[x] = -1; f.label[x] = L_NONE; f.mark[x] = 0; }
    for (int i = lane; i < nb; i += 32) f.bused[i] = 0;
    unsigned freebits = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int v = lane + 32 * k;
        vtop[k] = v; vlab[k] = L_NONE; tm[k] = SL_INF; slfrom[k] = -1;
        yb[k] = yinit[k] == (uint32_t)W_INF ? 0 : 2 * (i32)yinit[k];
        if (v < n) { f.mate[v] = -1; f.top[v] = (i16)v; freebits |= 1u << k; }
    }
    __syncwarp();
    // ---- greedy matching along tight edges
    for (int u = 0; u < n; ++u) {
        unsigned ufree = (__shfl_sync(0xffffffffu, freebits, u & 31) >> (u >> 5)) & 1u;
        if (!ufree) continue;
        i32 yu;
        GQM_SEL(yb, u, yu);
        const u16 *row = W + u * wstride;
        int found = -1;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (k < kmax) {
                const int v = lane + 32 * k;
                bool ok = false;
                if (found < 0 && v < n && v != u && ((freebits >> k) & 1u)) {
                    u16 wt = row[v];
                    ok = wt != W_INF && 4 * (i32)wt - yu - yb[k] == 0;
                }
                unsigned m = __ballot_sync(0xffffffffu, ok);
                if (found < 0 && m) found = 32 * k + __ffs(m) - 1;
            }
        }
        if (found >= 0) {
            if (lane == 0) { f.mate[u] = (i16)found; f.mate[found] = (i16)u; }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int v = lane + 32 * k;
                if (v == u || v == found) freebits &= ~(1u << k);
            }
        }
    }
    int nfree = __reduce_add_sync(0xffffffffu, __popc(freebits));
    GQ_STAT(7);
#ifdef GQ_STATS
    if (lane == 0) atomicAdd(&gq_stats[8], (unsigned long long)nfree);
#endif
    if (nfree == 0) return 0;
    // ---- every free vertex roots a tree
    bool news[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int v = lane + 32 * k;
        news[k] = v < n && ((freebits >> k) & 1u);
        if (news[k]) { vlab[k] = L_S; f.label[v] = L_S; f.troot[v] = (i16)v; }
    }
    __syncwarp();
    GQM_SCAN_WHERE(news);

    for (int guard = 0;; ++guard) {
        if (guard > 24 * nn + 256) return 3;
        // ---- earliest event: key = time << 10 | type << 8 | arg
        unsigned best = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (k < kmax) {
                const int v = lane + 32 * k;
                if (v < n && tm[k] < SL_INF && vlab[k] != L_T)
                    best = min(best, ((unsigned)tm[k] << 10) | (vlab[k] == L_S ? 0x100u : 0u) | (unsigned)v);
            }
        }
        for (int i = lane; i < bhi; i += 32)
            if (f.bused[i] && f.par[cap + i] < 0 && f.label[cap + i] == L_T)
                best = min(best, ((unsigned)(f.yb[i] >> 1) << 10) | 0x200u | (unsigned)i);
        best = __reduce_min_sync(0xffffffffu, best);
        GQ_STAT(0);
        if (best == 0xffffffffu) return 1;
        const i32 t = (i32)(best >> 10);
        const int type = (best >> 8) & 3;
        const int arg = best & 0xff;
        int src = -1;
        if (type != 2) {
            GQM_SEL(slfrom, arg, src);
            const int tu = f.top[src], tv = f.top[arg];
            bool valid = f.label[tu] == L_S && tu != tv;
            if (valid) {
                i32 yu, yv;
                GQM_YOF(src, t, yu);
                GQM_YOF(arg, t, yv);
                valid = 4 * (i32)W[src * wstride + arg] - yu - yv == 0;
            }
            if (!valid) {
                GQ_STAT(1);
                // stale entry: recompute the best edge from the current outer vertices into arg
                i32 yv;
                GQM_YOF(arg, D, yv);
                const u16 *row = W + arg * wstride;
                unsigned long long bb = ~0ull;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (k < kmax) {
                        const int u = lane + 32 * k;
                        if (u < n && vlab[k] == L_S && vtop[k] != tv) {
                            u16 wt = row[u];
                            if (wt != W_INF) {
                                i32 s = 4 * (i32)wt - yv - (yb[k] + D);
                                unsigned long long key = ((unsigned long long)(unsigned)s << 16) | (unsigned)u;
                                if (key < bb) bb = key;
                            }
                        }
                    }
                }
                unsigned hi = (unsigned)(bb >> 32), lo = (unsigned)bb;
                unsigned mh = __reduce_min_sync(0xffffffffu, hi);
                unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
                bb = ((unsigned long long)mh << 32) | ml;
                const i32 s = (i32)(bb >> 16);
                const i32 nt = bb == ~0ull ? SL_INF : (type == 0 ? D + s : D + (s >> 1));
                const int nf = bb == ~0ull ? -1 : (int)(bb & 0xffff);
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (lane + 32 * k == arg) { tm[k] = nt; slfrom[k] = nf; }
                continue;
            }
        }
        D = t;
#pragma unroll
        for (int k = 0; k < K; ++k) news[k] = false;
        if (type == 0) {
            GQ_STAT(2);
            // ---- grow: unlabelled node of arg becomes inner, its mate's node outer
            const int v = arg, u = src;
            const int tn = f.top[v];
            const int m = f.mate[f.base[tn]];
            if (m < 0) return 3;
            const int sn = f.top[m];
            if (lane == 0) {
                const i16 r = f.troot[f.top[u]];
                f.label[tn] = L_T; f.pu[tn] = (i16)u; f.pv[tn] = (i16)v; f.label[sn] = L_S;
                f.troot[tn] = r; f.troot[sn] = r;
                if (tn >= cap) f.yb[tn - cap] += 2 * D;   // z = zb - 2D from now on
                if (sn >= cap) f.yb[sn - cap] -= 2 * D;   // z = zb + 2D
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (lane + 32 * k < n) {
                    if (vtop[k] == tn) {
                        vlab[k] = L_T; yb[k] += D;
                        if (tm[k] < SL_INF) tm[k] -= D;
                    } else if (vtop[k] == sn) {
                        vlab[k] = L_S; yb[k] -= D; news[k] = true;
                        if (tm[k] < SL_INF) tm[k] = (tm[k] + D) >> 1;
                    }
                }
            }
        } else if (type == 1) {
            const int v = arg, u = src;
            const int tu = f.top[u], tv = f.top[v];
            const int r1 = f.troot[tu], r2 = f.troot[tv];
            if (r1 == r2) {
                GQ_STAT(3);
                // ---- shrink the odd cycle inside one tree
                for (int x = lane; x < nn; x += 32) f.mark[x] = 0;
                __syncwarp();
                int b = 0;
                if (lane == 0) {
                    b = shrink(f, u, v);
                    if (b >= 0) {
                        // children that are blossoms freeze their dual at its current value
                        int c = f.first[b - cap];
                        const int c0 = c;
                        do {
                            if (c >= cap) f.yb[c - cap] += f.label[c] == L_S ? 2 * D : -2 * D;
                            c = f.nxt[c];
                        } while (c != c0);
                        f.yb[b - cap] = -2 * D;   // z = zb + 2D = 0 now
                        f.troot[b] = (i16)r1;
                    }
                }
                b = __shfl_sync(0xffffffffu, b, 0);
                if (b < 0) return 2;
                bhi = max(bhi, b - cap + 1);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int x = lane + 32 * k;
                    if (x < n && f.par[vtop[k]] == b) {
                        if (vlab[k] == L_T) {
                            news[k] = true;
                            yb[k] -= 2 * D;
                            if (tm[k] < SL_INF) tm[k] = D + (tm[k] >> 1);
                        }
                        vtop[k] = b; vlab[k] = L_S;
                        f.top[x] = (i16)b;
                    }
                }
                __syncwarp();
            } else {
                GQ_STAT(4);
                // ---- augment between two trees, then dissolve exactly those two trees
                if (lane == 0) { augment_half(f, u, v); augment_half(f, v, u); }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int x = lane + 32 * k;
                    if (x < n && vlab[k] != L_NONE) {
                        const int r = f.troot[vtop[k]];
                        if (r == r1 || r == r2) {
                            if (vlab[k] == L_S) { yb[k] += D; if (tm[k] < SL_INF) tm[k] = 2 * tm[k] - D; }
                            else { yb[k] -= D; if (tm[k] < SL_INF) tm[k] += D; }
                            vlab[k] = L_NONE;
                        }
                    }
                    if (x == r1 || x == r2) freebits &= ~(1u << k);
                }
                __syncwarp();
                for (int x = lane; x < nn; x += 32) {
                    if (f.par[x] < 0 && f.label[x] != L_NONE && (f.troot[x] == r1 || f.troot[x] == r2)) {
                        if (x >= cap) f.yb[x - cap] += f.label[x] == L_S ? 2 * D : -2 * D;
                        f.label[x] = L_NONE;
                    }
                }
                __syncwarp();
                nfree -= 2;
                if (nfree <= 0) return 0;
            }
        } else {
            GQ_STAT(5);
            // ---- expand an inner blossom whose dual reached zero
            const int b = cap + arg;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int x = lane + 32 * k;
                if (x < n && vtop[k] == b) {
                    int c = x;
                    while (f.par[c] != b) c = f.par[c];
                    vtop[k] = c;
                    f.top[x] = (i16)c;
                }
            }
            __syncwarp();
            if (lane == 0) {
                const int fst = f.first[b - cap];
                const i16 r = f.troot[b];
                expand_relabel(f, b);
                int c = fst;
                do {
                    if (f.label[c] != L_NONE) {
                        f.troot[c] = r;
                        if (c >= cap) f.yb[c - cap] += f.label[c] == L_S ? -2 * D : 2 * D;
                    }
                    c = f.nxt[c];
                } while (c != fst);
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int x = lane + 32 * k;
                if (x < n && vlab[k] == L_T) {
                    const int nl = f.label[vtop[k]];
                    if (nl == L_S) {
                        news[k] = true; yb[k] -= 2 * D;
                        if (tm[k] < SL_INF) tm[k] = D + (tm[k] >> 1);
                    } else if (nl == L_NONE) {
                        yb[k] -= D;
                        if (tm[k] < SL_INF) tm[k] += D;
                    }
                    vlab[k] = nl;
                }
            }
        }
        GQM_SCAN_WHERE(news);
    }
#undef GQM_YACT
#undef GQM_SEL
#undef GQM_YOF
#undef GQM_SCAN
#undef GQM_SCAN_WHERE
}

}  // namespace gqr
