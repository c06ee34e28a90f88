// Warp-cooperative exact minimum-weight perfect matching on a small dense graph.
//
// One warp owns one matching problem: the n defects of one shot (n made even with a virtual
// boundary vertex) with integer weights W[i][j] = min(d(i,j), b(i)+b(j)) read from a workspace.
// The algorithm is the primal-dual blossom method (single alternating tree per phase, nested
// blossoms kept across phases); what is warp-shaped about it: every O(n) sweep -- slack scans of a
// newly outer vertex, the minimum-slack search, the dual update, relabelling -- is a lane-strided
// loop with shuffle reductions, while the pointer-chasing parts (augment, shrink, expand) run on
// lane 0.  It replaces PyMatching's sparse-blossom decode_batch.
//
// The same source compiles for the host with one "lane" (GQ_HOST_SIM) so that the logic can be
// checked on a CPU-only box against the oracle; that build is test scaffolding, never a fallback.
#pragma once
#include <stdint.h>

#ifdef GQ_HOST_SIM
#define GQ_DEV
#define GQ_HD inline
#define GQ_LANE 0
#define GQ_NLANES 1
#define GQ_SYNC() ((void)0)
static inline int gq_warp_min_i32(int v) { return v; }
static inline unsigned long long gq_warp_min_u64(unsigned long long v) { return v; }
static inline int gq_warp_any(int p) { return p; }
static inline int gq_warp_first_true(int p) { return p ? 0 : -1; }  // lane id of first true
static inline int gq_bcast(int v, int) { return v; }
#else
#define GQ_DEV __device__ __forceinline__
#define GQ_HD __host__ __device__ inline
#define GQ_LANE (threadIdx.x & 31)
#define GQ_NLANES 32
#define GQ_SYNC() __syncwarp()
__device__ __forceinline__ int gq_warp_min_i32(int v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ unsigned long long gq_warp_min_u64(unsigned long long v) {
    unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ int gq_warp_any(int p) { return __any_sync(0xffffffffu, p); }
__device__ __forceinline__ int gq_warp_first_true(int p) {
    unsigned b = __ballot_sync(0xffffffffu, p);
    return b ? (__ffs(b) - 1) : -1;
}
__device__ __forceinline__ int gq_bcast(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#endif

#ifdef GQ_HOST_SIM
#define GQ_COUNT(x) (++gq_counters[x])
static long long gq_counters[16];  // 0 shrink, 1 expand, 2 nested shrink, 3 augment through blossom, 4 odd-slack
#else
#define GQ_COUNT(x) ((void)0)
#endif

namespace gqb {

typedef int16_t i16;
typedef int32_t i32;
typedef uint16_t u16;

static const i32 SL_INF = 0x3fffffff;
static const u16 W_INF = 0xffff;

enum { L_NONE = 0, L_S = 1, L_T = 2 };

// Per-problem workspace.  cap = max real vertices (even); blossom ids are cap .. cap+cap/2-1, so
// node arrays hold cap + cap/2 entries.  All arrays may live in shared or global memory.
struct WS {
    u16 *W;       // [cap][wstride] weights (undoubled)
    int wstride;
    i32 *y;       // [nn]  vertex dual / blossom dual (doubled units)
    i32 *sl;      // [cap] min slack towards an outer vertex outside the own top node
    i16 *slfrom;  // [cap]
    i16 *mate;    // [cap] real mate or -1
    i16 *top;     // [cap] top-level node holding the vertex
    i16 *base;    // [nn]
    i16 *par;     // [nn] parent blossom or -1
    i16 *nxt;     // [nn] next / previous sibling around the parent's odd cycle
    i16 *prv;     // [nn]
    i16 *eu;      // [nn] edge (eu in x, ev in nxt[x]) closing x -> nxt[x]
    i16 *ev;      // [nn]
    i16 *pu;      // [nn] for a T node: tree edge (pu in parent S node, pv in this node)
    i16 *pv;      // [nn]
    i16 *first;   // [cap/2] child holding the base, per blossom
    i16 *stk;     // [2*nn] scratch stack
    int8_t *label;  // [nn]
    int8_t *mark;   // [nn]
    int8_t *bused;  // [cap/2] blossom id in use
};

GQ_HD size_t ws_bytes(int cap) {
    int nn = cap + cap / 2;
    int wstride = cap + 1;
    size_t b = 0;
    b += (size_t)cap * wstride * 2;
    b = (b + 3) & ~(size_t)3;
    b += (size_t)nn * 4 + (size_t)cap * 4;
    b += (size_t)cap * 2 * 3;            // slfrom mate top
    b += (size_t)nn * 2 * 8;             // base par nxt prv eu ev pu pv
    b += (size_t)(cap / 2) * 2;          // first
    b += (size_t)nn * 2 * 2;             // stk
    b += (size_t)nn * 2 + (size_t)(cap / 2);  // label mark bused
    return (b + 15) & ~(size_t)15;
}

GQ_HD void ws_bind(WS &w, void *mem, int cap) {
    int nn = cap + cap / 2;
    char *p = (char *)mem;
    w.wstride = cap + 1;
    w.W = (u16 *)p; p += (((size_t)cap * w.wstride * 2) + 3) & ~(size_t)3;
    w.y = (i32 *)p; p += (size_t)nn * 4;
    w.sl = (i32 *)p; p += (size_t)cap * 4;
    w.slfrom = (i16 *)p; p += (size_t)cap * 2;
    w.mate = (i16 *)p; p += (size_t)cap * 2;
    w.top = (i16 *)p; p += (size_t)cap * 2;
    w.base = (i16 *)p; p += (size_t)nn * 2;
    w.par = (i16 *)p; p += (size_t)nn * 2;
    w.nxt = (i16 *)p; p += (size_t)nn * 2;
    w.prv = (i16 *)p; p += (size_t)nn * 2;
    w.eu = (i16 *)p; p += (size_t)nn * 2;
    w.ev = (i16 *)p; p += (size_t)nn * 2;
    w.pu = (i16 *)p; p += (size_t)nn * 2;
    w.pv = (i16 *)p; p += (size_t)nn * 2;
    w.first = (i16 *)p; p += (size_t)(cap / 2) * 2;
    w.stk = (i16 *)p; p += (size_t)nn * 2 * 2;
    w.label = (int8_t *)p; p += nn;
    w.mark = (int8_t *)p; p += nn;
    w.bused = (int8_t *)p;
}

struct Solver {
    WS w;
    int n;    // real vertices (even)
    int cap;  // capacity the workspace was bound with; blossom ids start at cap
    int nb;   // cap / 2

    GQ_DEV i32 slack2(int u, int v) const {  // doubled slack of edge (u, v)
        return 2 * (i32)w.W[u * w.wstride + v] - w.y[u] - w.y[v];
    }

    // ---- lane-parallel: offer vertex u (just became outer) to every vertex outside its node
    GQ_DEV void scan(int u) {
        GQ_COUNT(5);
        const int tu = w.top[u];
        const i32 yu = w.y[u];
        const u16 *row = w.W + u * w.wstride;
        for (int v = GQ_LANE; v < n; v += GQ_NLANES) {
            if (w.top[v] == tu) continue;
            u16 wt = row[v];
            if (wt == W_INF) continue;
            i32 s = 2 * (i32)wt - yu - w.y[v];
            if (s < w.sl[v]) { w.sl[v] = s; w.slfrom[v] = (i16)u; }
        }
    }

    // scan every vertex of top-level node x (a vertex or a blossom)
    GQ_DEV void scan_node(int x) {
        if (x < cap) { scan(x); GQ_SYNC(); return; }
        for (int u = 0; u < n; ++u) {
            if (w.top[u] == x) { scan(u); GQ_SYNC(); }
        }
    }

    GQ_DEV void rescan_all() {
        GQ_COUNT(6);
        for (int v = GQ_LANE; v < n; v += GQ_NLANES) { w.sl[v] = SL_INF; w.slfrom[v] = -1; }
        GQ_SYNC();
        for (int u = 0; u < n; ++u) {
            if (w.label[w.top[u]] == L_S) { scan(u); GQ_SYNC(); }
        }
    }

    // ---- lane 0: make real vertex a the base of top-level-or-nested node x, flipping the
    // matching around the odd cycles on the way.
    GQ_DEV void rebase(int x0, int a0) {
        int sp = 0;
        w.stk[sp++] = (i16)x0; w.stk[sp++] = (i16)a0;
        while (sp) {
            int a = w.stk[--sp];
            int B = w.stk[--sp];
            if (B < cap) continue;
            if (w.base[B] == a) continue;
            GQ_COUNT(3);
            int c = a;
            while (w.par[c] != B) c = w.par[c];
            w.stk[sp++] = (i16)c; w.stk[sp++] = (i16)a;
            int f = w.first[B - cap];
            // parity of the position of c counted from f
            int steps = 0;
            for (int t = c; t != f; t = w.prv[t]) ++steps;
            if ((steps & 1) == 0) {
                int p = c;
                while (p != f) {
                    int q1 = w.prv[p];      // currently matched to p; will pair with prv[q1]
                    int q2 = w.prv[q1];
                    // new pair (q1, q2) through the edge stored at q2 (q2 -> q1)
                    int a2 = w.eu[q2], a1 = w.ev[q2];
                    w.mate[a2] = (i16)a1; w.mate[a1] = (i16)a2;
                    w.stk[sp++] = (i16)q2; w.stk[sp++] = (i16)a2;
                    w.stk[sp++] = (i16)q1; w.stk[sp++] = (i16)a1;
                    p = q2;
                }
            } else {
                int p = c;
                while (p != f) {
                    int q1 = w.nxt[p];
                    int q2 = w.nxt[q1];
                    // new pair (q1, q2) through the edge stored at q1 (q1 -> q2)
                    int a1 = w.eu[q1], a2 = w.ev[q1];
                    w.mate[a1] = (i16)a2; w.mate[a2] = (i16)a1;
                    w.stk[sp++] = (i16)q1; w.stk[sp++] = (i16)a1;
                    w.stk[sp++] = (i16)q2; w.stk[sp++] = (i16)a2;
                    p = q2;
                }
            }
            w.first[B - cap] = (i16)c;
            w.base[B] = (i16)a;
        }
    }

    // ---- lane 0: flip the alternating path from outer vertex u up to the root, after pairing
    // u with the free vertex v.
    GQ_DEV void augment(int u, int v) {
        int a = u, nm = v;
        for (;;) {
            int x = w.top[a];
            int m = w.mate[w.base[x]];
            rebase(x, a);
            w.mate[a] = (i16)nm; w.mate[nm] = (i16)a;
            if (m < 0) break;
            int t = w.top[m];
            int tpu = w.pu[t], tpv = w.pv[t];
            rebase(t, tpv);
            nm = tpv; a = tpu;
        }
    }

    // tree parent (an S node) of S node x, or -1 at the root; *tnode gets the T node in between
    GQ_DEV int parent_S(int x, int *tnode) const {
        int m = w.mate[w.base[x]];
        if (m < 0) return -1;
        int t = w.top[m];
        *tnode = t;
        return w.top[w.pu[t]];
    }

    // ---- lane 0: shrink the odd cycle closed by the tight edge (u, v) between two outer nodes
    GQ_DEV int shrink(int u, int v) {
        int nn = cap + nb;
        for (int i = 0; i < nn; ++i) w.mark[i] = 0;
        int xu = w.top[u], xv = w.top[v];
        for (int x = xu; x >= 0;) { w.mark[x] = 1; int t; x = parent_S(x, &t); }
        int lca = xv;
        while (!w.mark[lca]) { int t; lca = parent_S(lca, &t); }
        GQ_COUNT(0);
        int b = -1;
        for (int i = 0; i < nb; ++i) if (!w.bused[i]) { b = cap + i; w.bused[i] = 1; break; }
        if (b < 0) return -1;
        // cycle: lca -> ... -> xu -> xv -> ... -> lca ; build as a doubly linked ring
        // start: ring = {lca}
        w.nxt[lca] = (i16)lca; w.prv[lca] = (i16)lca;
        // u side: walk up from xu, inserting each node right after lca
        {
            int below = -1;  // node inserted previously (lower in the tree)
            int bu = 0, bv = 0;  // edge from the node being inserted to 'below'
            int x = xu;
            int first_iter = 1;
            while (x != lca) {
                // x is an S node; its T parent t; then S grandparent
                int t; int up = parent_S(x, &t);
                // insert x after lca
                int old = w.nxt[lca];
                w.nxt[lca] = (i16)x; w.prv[x] = (i16)lca; w.nxt[x] = (i16)old; w.prv[old] = (i16)x;
                if (first_iter) { first_iter = 0; }
                else { w.eu[x] = (i16)bu; w.ev[x] = (i16)bv; }
                // insert t after lca; edge t -> x is the matched edge (mate[base x] in t, base x)
                old = w.nxt[lca];
                w.nxt[lca] = (i16)t; w.prv[t] = (i16)lca; w.nxt[t] = (i16)old; w.prv[old] = (i16)t;
                w.eu[t] = w.mate[w.base[x]]; w.ev[t] = w.base[x];
                // edge from the next inserted node (up) down to t is t's tree edge (pu in up, pv in t)
                bu = w.pu[t]; bv = w.pv[t];
                below = t;
                x = up;
            }
            // edge lca -> nxt[lca]
            if (below >= 0) { w.eu[lca] = (i16)bu; w.ev[lca] = (i16)bv; }
            (void)below;
        }
        // the closing edge: stored at xu (xu -> next on the v side)
        // v side: walk up from xv appending before lca (at the ring's tail)
        {
            int tail = w.prv[lca];  // == xu if xu != lca, else lca
            // edge tail -> (first v-side node or lca) is (u, v)
            w.eu[tail] = (i16)u; w.ev[tail] = (i16)v;
            int x = xv;
            while (x != lca) {
                int t; int up = parent_S(x, &t);
                // append x
                w.nxt[tail] = (i16)x; w.prv[x] = (i16)tail; w.nxt[x] = (i16)lca; w.prv[lca] = (i16)x;
                // edge x -> t: (base x, mate[base x])
                w.eu[x] = w.base[x]; w.ev[x] = w.mate[w.base[x]];
                // append t
                w.nxt[x] = (i16)t; w.prv[t] = (i16)x; w.nxt[t] = (i16)lca; w.prv[lca] = (i16)t;
                // edge t -> up: (pv in t, pu in up)
                w.eu[t] = w.pv[t]; w.ev[t] = w.pu[t];
                tail = t;
                x = up;
            }
        }
        // register the blossom
        w.first[b - cap] = (i16)lca;
        w.base[b] = w.base[lca];
        w.par[b] = -1;
        w.y[b] = 0;
        w.label[b] = L_S;
        int c = lca;
        do { w.par[c] = (i16)b; if (c >= cap) GQ_COUNT(2); c = w.nxt[c]; } while (c != lca);
        return b;
    }

    // ---- expand a T blossom whose dual reached zero (mid-phase)
    GQ_DEV void expand(int b) {
        GQ_COUNT(1);
        // children become top-level: fix top[] of the member vertices first (lane-parallel)
        for (int v = GQ_LANE; v < n; v += GQ_NLANES) {
            if (w.top[v] == b) {
                int c = v;
                while (w.par[c] != b) c = w.par[c];
                w.top[v] = (i16)c;
            }
        }
        GQ_SYNC();
        if (GQ_LANE == 0) {
            int f = w.first[b - cap];
            int entry = w.pv[b];
            int c = w.top[entry];
            int steps = 0;
            for (int t = c; t != f; t = w.prv[t]) ++steps;
            // unlabel everything, detach
            int x = f;
            do { w.label[x] = L_NONE; int nx = w.nxt[x]; w.par[x] = -1; x = nx; } while (x != f);
            w.label[c] = L_T; w.pu[c] = w.pu[b]; w.pv[c] = w.pv[b];
            if ((steps & 1) == 0) {
                int p = c;
                while (p != f) {
                    int s = w.prv[p];   // S node matched to p
                    int t = w.prv[s];   // next T node, parent s, edge stored at t: (eu in t, ev in s)
                    w.label[s] = L_S;
                    w.label[t] = L_T; w.pu[t] = w.ev[t]; w.pv[t] = w.eu[t];
                    p = t;
                }
            } else {
                int p = c;
                while (p != f) {
                    int s = w.nxt[p];
                    int t = w.nxt[s];   // edge stored at s: (eu in s, ev in t)
                    w.label[s] = L_S;
                    w.label[t] = L_T; w.pu[t] = w.eu[s]; w.pv[t] = w.ev[s];
                    p = t;
                }
            }
            w.bused[b - cap] = 0;
            w.label[b] = L_NONE;
        }
        GQ_SYNC();
        // (simple and safe) rebuild every slack from the current outer set
        rescan_all();
    }

    // Solve.  Returns 0 on success, 1 if no perfect matching exists, 2 on internal overflow.
    GQ_DEV int solve() {
        const int nn = cap + nb;
        // ---- initial duals and greedy tight matching
        for (int v = GQ_LANE; v < n; v += GQ_NLANES) {
            const u16 *row = w.W + v * w.wstride;
            u16 m = W_INF;
            for (int j = 0; j < n; ++j) if (j != v && row[j] < m) m = row[j];
            w.y[v] = (m == W_INF) ? 0 : (i32)m;
            w.mate[v] = -1; w.top[v] = (i16)v;
        }
        for (int x = GQ_LANE; x < nn; x += GQ_NLANES) {
            w.base[x] = (i16)(x < cap ? x : -1); w.par[x] = -1; w.label[x] = L_NONE;
        }
        for (int i = GQ_LANE; i < nb; i += GQ_NLANES) w.bused[i] = 0;
        GQ_SYNC();
        for (int u = 0; u < n; ++u) {
            if (w.mate[u] >= 0) continue;  // uniform: mate[] only changes under GQ_SYNC below
            int found = -1;
            for (int v0 = 0; v0 < n && found < 0; v0 += GQ_NLANES) {
                int v = v0 + GQ_LANE;
                int ok = 0;
                if (v < n && v != u && w.mate[v] < 0 && w.W[u * w.wstride + v] != W_INF)
                    ok = slack2(u, v) == 0;
                int l = gq_warp_first_true(ok);
                if (l >= 0) found = v0 + l;
            }
            if (found >= 0 && GQ_LANE == 0) { w.mate[u] = (i16)found; w.mate[found] = (i16)u; }
            GQ_SYNC();
        }
        // ---- phases
        for (int root = 0; root < n; ++root) {
            if (w.mate[root] >= 0) continue;
            GQ_COUNT(7);
            for (int x = GQ_LANE; x < nn; x += GQ_NLANES) w.label[x] = L_NONE;
            for (int v = GQ_LANE; v < n; v += GQ_NLANES) { w.sl[v] = SL_INF; w.slfrom[v] = -1; }
            GQ_SYNC();
            if (GQ_LANE == 0) w.label[root] = L_S;
            GQ_SYNC();
            scan(root);
            GQ_SYNC();
            for (int guard = 0;; ++guard) {
                if (guard > 6 * nn + 64) return 3;
                // ---- minimum-slack event: key = (delta << 20 | type << 16 | arg)
                unsigned long long best = ~0ull;
                for (int v = GQ_LANE; v < n; v += GQ_NLANES) {
                    i32 s = w.sl[v];
                    if (s >= SL_INF) continue;
                    int lab = w.label[w.top[v]];
                    if (lab == L_NONE) {
                        unsigned long long k = ((unsigned long long)(unsigned)s << 20) | (0u << 16) | (unsigned)v;
                        if (k < best) best = k;
                    } else if (lab == L_S) {
                        if (s & 1) GQ_COUNT(4);
                        unsigned long long k = ((unsigned long long)(unsigned)(s >> 1) << 20) | (1u << 16) | (unsigned)v;
                        if (k < best) best = k;
                    }
                }
                for (int i = GQ_LANE; i < nb; i += GQ_NLANES) {
                    int b = cap + i;
                    if (w.bused[i] && w.par[b] < 0 && w.label[b] == L_T) {
                        unsigned long long k = ((unsigned long long)(unsigned)(w.y[b] >> 1) << 20) | (2u << 16) | (unsigned)i;
                        if (k < best) best = k;
                    }
                }
                best = gq_warp_min_u64(best);
                GQ_COUNT(8);
                if (best == ~0ull) return 1;
                const i32 delta = (i32)(best >> 20);
                const int type = (int)((best >> 16) & 0xf);
                const int arg = (int)(best & 0xffff);
                // ---- dual update
                if (delta) {
                    for (int v = GQ_LANE; v < n; v += GQ_NLANES) {
                        int lab = w.label[w.top[v]];
                        if (lab == L_S) { w.y[v] += delta; if (w.sl[v] < SL_INF) w.sl[v] -= 2 * delta; }
                        else if (lab == L_T) { w.y[v] -= delta; }
                        else if (w.sl[v] < SL_INF) w.sl[v] -= delta;
                    }
                    for (int i = GQ_LANE; i < nb; i += GQ_NLANES) {
                        int b = cap + i;
                        if (w.bused[i] && w.par[b] < 0) {
                            if (w.label[b] == L_S) w.y[b] += 2 * delta;
                            else if (w.label[b] == L_T) w.y[b] -= 2 * delta;
                        }
                    }
                    GQ_SYNC();
                }
                if (type == 0) {
                    const int v = arg, u = w.slfrom[v];
                    const int t = w.top[v];
                    const int m = w.mate[w.base[t]];
                    if (m < 0) {
                        if (GQ_LANE == 0) augment(u, v);
                        GQ_SYNC();
                        break;
                    }
                    const int s = w.top[m];
                    if (GQ_LANE == 0) {
                        w.label[t] = L_T; w.pu[t] = (i16)u; w.pv[t] = (i16)v;
                        w.label[s] = L_S;
                    }
                    GQ_SYNC();
                    scan_node(s);
                } else if (type == 1) {
                    const int v = arg, u = w.slfrom[v];
                    int b = 0;
                    if (GQ_LANE == 0) b = shrink(u, v);
                    b = gq_bcast(b, 0);
                    if (b < 0) return 2;
                    GQ_SYNC();
                    for (int x = GQ_LANE; x < n; x += GQ_NLANES) {
                        int t = w.top[x];
                        if (w.par[t] == b) w.top[x] = (i16)b;
                    }
                    GQ_SYNC();
                    rescan_all();
                } else {
                    expand(cap + arg);
                }
            }
        }
        return 0;
    }
};

}  // namespace gqb
