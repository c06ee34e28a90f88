/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
 * (graphqec_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED against the reference wheels: the arithmetic of this path lives in PyMatching 2
 * (pymatching>=2.2.1, /root/reference/pyproject.toml:11), which is selected by name at
 * /root/reference/graphqec/lab/threshold/threshold_lab.py:186,196 and is not installed in this
 * image, and the reference's only test of the path is stale.  What is
 * restated here is PyMatching's published behaviour: exact
 * minimum-weight perfect matching with a boundary on the detector graph, weight ln((1-p)/p),
 * prediction = XOR of the observable masks along the matched shortest paths.  It is pinned
 * instead against networkx.max_weight_matching and brute force (tests/test_oracle.py), and -- through the
 * whole pipeline -- against the error counts the reference's own sinter + PyMatching runs recorded
 * (tests/test_gpu_reference_stats.py).
 *
 * Method (deliberately different from the CUDA decoder): all-pairs shortest paths by Dijkstra
 * at graph creation; per shot, the defects become a complete graph with weights
 * w(i,j) = min(d(i,j), b(i)+b(j)) (b = distance to the boundary; one extra "boundary" node when
 * the defect count is odd) and the classic O(n^3) primal-dual blossom algorithm for maximum-weight
 * matching is run on (BIG - w).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef long long ll;

/* ------------------------------------------------------------------------------------------ */
/* O(n^3) maximum weight general matching (vertices 1..n, 0 = "none")                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int u, v; ll w; } Edge;

typedef struct {
    int cap;          /* max real vertices */
    int n, n_x;
    Edge *g;          /* (2cap+1)^2 */
    ll *lab;
    int *match, *slack, *st, *pa, *S, *vis;
    int *flower_from; /* (2cap+1) x (cap+1) */
    int *flower;      /* (2cap+1) x (2cap+1) children lists */
    int *flower_len;
    int *q; int qh, qt;
    int vis_t;
    int stride;
} Blossom;

#define G(B, a, b) ((B)->g[(size_t)(a) * (B)->stride + (b)])
#define FF(B, a, b) ((B)->flower_from[(size_t)(a) * ((B)->cap + 1) + (b)])
#define FL(B, a) ((B)->flower + (size_t)(a) * (B)->stride)

static Blossom *blossom_new(int cap) {
    Blossom *B = calloc(1, sizeof(Blossom));
    B->cap = cap;
    B->stride = 2 * cap + 1;
    size_t s = (size_t)B->stride;
    B->g = malloc(s * s * sizeof(Edge));
    B->lab = malloc(s * sizeof(ll));
    B->match = malloc(s * sizeof(int));
    B->slack = malloc(s * sizeof(int));
    B->st = malloc(s * sizeof(int));
    B->pa = malloc(s * sizeof(int));
    B->S = malloc(s * sizeof(int));
    B->vis = calloc(s, sizeof(int));
    B->flower_from = malloc(s * (size_t)(cap + 1) * sizeof(int));
    B->flower = malloc(s * s * sizeof(int));
    B->flower_len = calloc(s, sizeof(int));
    B->q = malloc(s * 4 * sizeof(int));
    B->vis_t = 0;
    return B;
}

static void blossom_free(Blossom *B) {
    if (!B) return;
    free(B->g); free(B->lab); free(B->match); free(B->slack); free(B->st); free(B->pa);
    free(B->S); free(B->vis); free(B->flower_from); free(B->flower); free(B->flower_len);
    free(B->q); free(B);
}

static int g_jump_start = 0;
/* 1 = greedy dual-feasible start + one root per phase (see blossom_solve).  Process-wide; set it before any decoding
 * thread runs.  Off for every parity check; bench.py turns it on for the timed CPU baseline. */
void orc_set_jump_start(int on) { g_jump_start = on ? 1 : 0; }

static inline ll e_delta(Blossom *B, Edge e) {
    return B->lab[e.u] + B->lab[e.v] - G(B, e.u, e.v).w * 2;
}
static inline void update_slack(Blossom *B, int u, int x) {
    if (!B->slack[x] || e_delta(B, G(B, u, x)) < e_delta(B, G(B, B->slack[x], x))) B->slack[x] = u;
}
static void set_slack(Blossom *B, int x) {
    B->slack[x] = 0;
    for (int u = 1; u <= B->n; ++u)
        if (G(B, u, x).w > 0 && B->st[u] != x && B->S[B->st[u]] == 0) update_slack(B, u, x);
}
static void q_push(Blossom *B, int x) {
    if (x <= B->n) { B->q[B->qt++] = x; return; }
    int *f = FL(B, x);
    for (int i = 0; i < B->flower_len[x]; ++i) q_push(B, f[i]);
}
static void set_st(Blossom *B, int x, int b) {
    B->st[x] = b;
    if (x > B->n) {
        int *f = FL(B, x);
        for (int i = 0; i < B->flower_len[x]; ++i) set_st(B, f[i], b);
    }
}
static void reverse_ints(int *a, int len) {
    for (int i = 0, j = len - 1; i < j; ++i, --j) { int t = a[i]; a[i] = a[j]; a[j] = t; }
}
static int get_pr(Blossom *B, int b, int xr) {
    int *f = FL(B, b);
    int len = B->flower_len[b], pr = 0;
    while (pr < len && f[pr] != xr) ++pr;
    if (pr % 2 == 1) {
        reverse_ints(f + 1, len - 1);
        return len - pr;
    }
    return pr;
}
static void set_match(Blossom *B, int u, int v) {
    B->match[u] = G(B, u, v).v;
    if (u > B->n) {
        Edge e = G(B, u, v);
        int xr = FF(B, u, e.u);
        int pr = get_pr(B, u, xr);
        int *f = FL(B, u);
        for (int i = 0; i < pr; ++i) set_match(B, f[i], f[i ^ 1]);
        set_match(B, xr, v);
        /* rotate left by pr */
        int len = B->flower_len[u];
        reverse_ints(f, pr);
        reverse_ints(f + pr, len - pr);
        reverse_ints(f, len);
    }
}
static void augment(Blossom *B, int u, int v) {
    for (;;) {
        int xnv = B->st[B->match[u]];
        set_match(B, u, v);
        if (!xnv) return;
        set_match(B, xnv, B->st[B->pa[xnv]]);
        u = B->st[B->pa[xnv]];
        v = xnv;
    }
}
static int get_lca(Blossom *B, int u, int v) {
    ++B->vis_t;
    while (u || v) {
        if (u) {
            if (B->vis[u] == B->vis_t) return u;
            B->vis[u] = B->vis_t;
            u = B->st[B->match[u]];
            if (u) u = B->st[B->pa[u]];
        }
        int t = u; u = v; v = t;
    }
    return 0;
}
static void add_blossom(Blossom *B, int u, int lca, int v) {
    int b = B->n + 1;
    while (b <= B->n_x && B->st[b]) ++b;
    if (b > B->n_x) ++B->n_x;
    B->lab[b] = 0; B->S[b] = 0;
    B->match[b] = B->match[lca];
    int *f = FL(B, b);
    int len = 0;
    f[len++] = lca;
    for (int x = u, y; x != lca; x = B->st[B->pa[y]]) {
        f[len++] = x;
        f[len++] = y = B->st[B->match[x]];
        q_push(B, y);
    }
    reverse_ints(f + 1, len - 1);
    for (int x = v, y; x != lca; x = B->st[B->pa[y]]) {
        f[len++] = x;
        f[len++] = y = B->st[B->match[x]];
        q_push(B, y);
    }
    B->flower_len[b] = len;
    set_st(B, b, b);
    for (int x = 1; x <= B->n_x; ++x) G(B, b, x).w = G(B, x, b).w = 0;
    for (int x = 1; x <= B->n; ++x) FF(B, b, x) = 0;
    for (int i = 0; i < len; ++i) {
        int xs = f[i];
        for (int x = 1; x <= B->n_x; ++x)
            if (G(B, b, x).w == 0 || e_delta(B, G(B, xs, x)) < e_delta(B, G(B, b, x))) {
                G(B, b, x) = G(B, xs, x);
                G(B, x, b) = G(B, x, xs);
            }
        for (int x = 1; x <= B->n; ++x)
            if (FF(B, xs, x)) FF(B, b, x) = xs;
    }
    set_slack(B, b);
}
static void expand_blossom(Blossom *B, int b) {
    int *f = FL(B, b);
    int len = B->flower_len[b];
    for (int i = 0; i < len; ++i) set_st(B, f[i], f[i]);
    int xr = FF(B, b, G(B, b, B->pa[b]).u);
    int pr = get_pr(B, b, xr);
    for (int i = 0; i < pr; i += 2) {
        int xs = f[i], xns = f[i + 1];
        B->pa[xs] = G(B, xns, xs).u;
        B->S[xs] = 1; B->S[xns] = 0;
        B->slack[xs] = 0;
        set_slack(B, xns);
        q_push(B, xns);
    }
    B->S[xr] = 1; B->pa[xr] = B->pa[b];
    for (int i = pr + 1; i < len; ++i) {
        int xs = f[i];
        B->S[xs] = -1;
        set_slack(B, xs);
    }
    B->st[b] = 0;
}
static int on_found_edge(Blossom *B, Edge e) {
    int u = B->st[e.u], v = B->st[e.v];
    if (B->S[v] == -1) {
        /* single-root phases (jump start) can meet a free vertex outside the tree: that is an augmenting path */
        if (!B->match[v]) { augment(B, u, v); augment(B, v, u); return 1; }
        B->pa[v] = e.u; B->S[v] = 1;
        int nu = B->st[B->match[v]];
        B->slack[v] = B->slack[nu] = 0;
        B->S[nu] = 0;
        q_push(B, nu);
    } else if (B->S[v] == 0) {
        int lca = get_lca(B, u, v);
        if (!lca) { augment(B, u, v); augment(B, v, u); return 1; }
        add_blossom(B, u, lca, v);
    }
    return 0;
}
static int matching_phase(Blossom *B) {
    for (int x = 1; x <= B->n_x; ++x) { B->S[x] = -1; B->slack[x] = 0; }
    B->qh = B->qt = 0;
    for (int x = 1; x <= B->n_x; ++x)
        if (B->st[x] == x && !B->match[x]) {
            B->pa[x] = 0; B->S[x] = 0; q_push(B, x);
            /* after a jump start the labels no longer share one parity; inside ONE tree they do (its edges are tight),
             * which keeps every dual step integral -- so one root per phase */
            if (g_jump_start) break;
        }
    if (B->qh == B->qt) return 0;
    for (;;) {
        while (B->qh < B->qt) {
            int u = B->q[B->qh++];
            if (B->S[B->st[u]] == 1) continue;
            for (int v = 1; v <= B->n; ++v)
                if (G(B, u, v).w > 0 && B->st[u] != B->st[v]) {
                    if (e_delta(B, G(B, u, v)) == 0) {
                        if (on_found_edge(B, G(B, u, v))) return 1;
                    } else
                        update_slack(B, u, B->st[v]);
                }
        }
        ll d = (ll)4e18;
        for (int b = B->n + 1; b <= B->n_x; ++b)
            if (B->st[b] == b && B->S[b] == 1 && B->lab[b] / 2 < d) d = B->lab[b] / 2;
        for (int x = 1; x <= B->n_x; ++x)
            if (B->st[x] == x && B->slack[x]) {
                ll ed = e_delta(B, G(B, B->slack[x], x));
                if (B->S[x] == -1) { if (ed < d) d = ed; }
                else if (B->S[x] == 0) { if (ed / 2 < d) d = ed / 2; }
            }
        for (int u = 1; u <= B->n; ++u) {
            if (B->S[B->st[u]] == 0) {
                if (B->lab[u] <= d) return 0;
                B->lab[u] -= d;
            } else if (B->S[B->st[u]] == 1)
                B->lab[u] += d;
        }
        for (int b = B->n + 1; b <= B->n_x; ++b)
            if (B->st[b] == b) {
                if (B->S[B->st[b]] == 0) B->lab[b] += d * 2;
                else if (B->S[B->st[b]] == 1) B->lab[b] -= d * 2;
            }
        B->qh = B->qt = 0;
        for (int x = 1; x <= B->n_x; ++x)
            if (B->st[x] == x && B->slack[x] && B->st[B->slack[x]] != x &&
                e_delta(B, G(B, B->slack[x], x)) == 0)
                if (on_found_edge(B, G(B, B->slack[x], x))) return 1;
        for (int b = B->n + 1; b <= B->n_x; ++b)
            if (B->st[b] == b && B->S[b] == 1 && B->lab[b] == 0) expand_blossom(B, b);
    }
}

/* weights w[i*n+j] (0-indexed, symmetric, > 0 means edge present). mate_out[i] = j or -1. */
static ll blossom_solve(Blossom *B, int n, const ll *w, int *mate_out) {
    B->n = B->n_x = n;
    ll wmax = 0;
    for (int u = 0; u <= 2 * n; ++u) { B->match[u] = 0; B->st[u] = (u <= n) ? u : 0; B->flower_len[u] = 0; }
    for (int u = 1; u <= n; ++u)
        for (int v = 1; v <= n; ++v) {
            Edge e = { u, v, (u == v) ? 0 : w[(size_t)(u - 1) * n + (v - 1)] };
            G(B, u, v) = e;
            FF(B, u, v) = (u == v) ? u : 0;
            if (e.w > wmax) wmax = e.w;
        }
    for (int u = 1; u <= n; ++u) B->lab[u] = wmax;
    if (g_jump_start) {
        /* Optional (orc_set_jump_start; off for every parity check, on for the timed CPU baseline): start from the tightest
         * feasible labels, lab[u] = the heaviest edge at u (lab[u] + lab[v] >= 2 w(u,v) holds for every edge), and match
         * greedily along edges that are tight under them -- mutual "nearest neighbours" in the minimum-weight reading.
         * The phases then only have to augment from the vertices left free (about one in six at the north star instead
         * of all of them), and the optimum they reach has the same weight: the start is dual feasible with a matching on
         * tight edges, which is all the primal-dual method asks of its starting point.  Labels stay far from zero (the
         * weights carry the offset that forces a perfect matching), so the template's "free vertex reached zero" exit
         * never fires. */
        for (int u = 1; u <= n; ++u) {
            ll best = 0;
            for (int v = 1; v <= n; ++v) if (G(B, u, v).w > best) best = G(B, u, v).w;
            B->lab[u] = best;
        }
        for (int u = 1; u <= n; ++u) {
            if (B->match[u]) continue;
            for (int v = u + 1; v <= n; ++v)
                if (!B->match[v] && G(B, u, v).w > 0 && B->lab[u] + B->lab[v] == 2 * G(B, u, v).w) {
                    B->match[u] = v; B->match[v] = u;
                    break;
                }
        }
    }
    while (matching_phase(B)) {}
    ll tot = 0;
    for (int u = 1; u <= n; ++u) {
        mate_out[u - 1] = B->match[u] ? B->match[u] - 1 : -1;
        if (B->match[u] && B->match[u] < u) tot += G(B, u, B->match[u]).w;
    }
    return tot;
}

/* exported for direct tests: maximum weight matching on a dense weight matrix */
ll orc_max_weight_matching(int n, const ll *w, int *mate_out) {
    Blossom *B = blossom_new(n > 0 ? n : 1);
    ll t = blossom_solve(B, n, w, mate_out);
    blossom_free(B);
    return t;
}

/* ------------------------------------------------------------------------------------------ */
/* detector graph with boundary, all-pairs shortest paths                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int nd;            /* detectors; node nd = boundary */
    int no;
    int32_t *dist;     /* (nd) x (nd+1): column nd = distance to boundary; INT32_MAX/4 = unreachable */
    uint8_t *par;      /* observable mask of that shortest path */
    int ne;
    int *eu, *ev; ll *ew; uint8_t *eo;
} OrcGraph;

#define UNREACH (INT32_MAX / 4)

typedef struct { ll d; int v; } HeapItem;

static void heap_push(HeapItem *h, int *n, HeapItem it) {
    int i = (*n)++;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (h[p].d <= it.d) break;
        h[i] = h[p]; i = p;
    }
    h[i] = it;
}
static HeapItem heap_pop(HeapItem *h, int *n) {
    HeapItem top = h[0];
    HeapItem last = h[--(*n)];
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= *n) break;
        if (c + 1 < *n && h[c + 1].d < h[c].d) ++c;
        if (h[c].d >= last.d) break;
        h[i] = h[c]; i = c;
    }
    h[i] = last;
    return top;
}

/* edges: v == -1 means boundary edge.  Parallel edges must already be merged by the caller.
 * Paths are NOT allowed to pass through the boundary node (it is a sink). */
OrcGraph *orc_graph_create(int nd, int no, int ne, const int *eu, const int *ev, const ll *ew,
                           const uint8_t *eo) {
    OrcGraph *g = calloc(1, sizeof(OrcGraph));
    g->nd = nd; g->no = no; g->ne = ne;
    g->eu = malloc(sizeof(int) * (ne + 1)); g->ev = malloc(sizeof(int) * (ne + 1));
    g->ew = malloc(sizeof(ll) * (ne + 1)); g->eo = malloc(ne + 1);
    memcpy(g->eu, eu, sizeof(int) * ne); memcpy(g->ev, ev, sizeof(int) * ne);
    memcpy(g->ew, ew, sizeof(ll) * ne); memcpy(g->eo, eo, ne);
    int N = nd + 1;
    int *deg = calloc(N + 1, sizeof(int));
    for (int e = 0; e < ne; ++e) {
        int a = eu[e], b = ev[e] < 0 ? nd : ev[e];
        deg[a]++; deg[b]++;
    }
    int *ptr = malloc(sizeof(int) * (N + 1));
    ptr[0] = 0;
    for (int i = 0; i < N; ++i) ptr[i + 1] = ptr[i] + deg[i];
    int *adj = malloc(sizeof(int) * (2 * ne + 1));
    int *aed = malloc(sizeof(int) * (2 * ne + 1));
    int *fill = calloc(N, sizeof(int));
    for (int e = 0; e < ne; ++e) {
        int a = eu[e], b = ev[e] < 0 ? nd : ev[e];
        adj[ptr[a] + fill[a]] = b; aed[ptr[a] + fill[a]++] = e;
        adj[ptr[b] + fill[b]] = a; aed[ptr[b] + fill[b]++] = e;
    }
    g->dist = malloc(sizeof(int32_t) * (size_t)nd * N);
    g->par = malloc((size_t)nd * N);
    ll *d = malloc(sizeof(ll) * N);
    uint8_t *po = malloc(N);
    char *done = malloc(N);
    HeapItem *heap = malloc(sizeof(HeapItem) * (2 * ne + N + 4));
    for (int s = 0; s < nd; ++s) {
        for (int i = 0; i < N; ++i) { d[i] = (ll)4e18; po[i] = 0; done[i] = 0; }
        int hn = 0;
        d[s] = 0;
        heap_push(heap, &hn, (HeapItem){0, s});
        while (hn) {
            HeapItem it = heap_pop(heap, &hn);
            int u = it.v;
            if (done[u]) continue;
            done[u] = 1;
            if (u == nd) continue; /* boundary is a sink */
            for (int k = ptr[u]; k < ptr[u + 1]; ++k) {
                int v = adj[k], e = aed[k];
                ll nd2 = it.d + g->ew[e];
                if (nd2 < d[v]) {
                    d[v] = nd2; po[v] = po[u] ^ g->eo[e];
                    heap_push(heap, &hn, (HeapItem){nd2, v});
                }
            }
        }
        for (int i = 0; i < N; ++i) {
            g->dist[(size_t)s * N + i] = d[i] >= UNREACH ? UNREACH : (int32_t)d[i];
            g->par[(size_t)s * N + i] = po[i];
        }
    }
    free(deg); free(ptr); free(adj); free(aed); free(fill); free(d); free(po); free(done); free(heap);
    return g;
}

void orc_graph_free(OrcGraph *g) {
    if (!g) return;
    free(g->dist); free(g->par); free(g->eu); free(g->ev); free(g->ew); free(g->eo); free(g);
}

const int32_t *orc_graph_dist(OrcGraph *g) { return g->dist; }
const uint8_t *orc_graph_par(OrcGraph *g) { return g->par; }

/* Decode one shot given its defect list.  Returns predicted observable mask; *weight_out gets the
 * total matching weight; pairs_out (optional, 2 ints per defect: defect node, partner node or -1
 * for boundary) lists the matching.  Returns 0xFF.. style failure through *ok = 0. */
static uint8_t decode_defects(OrcGraph *g, Blossom *B, ll *wbuf, int *mate, const int *def, int n,
                              ll *weight_out, int *pairs_out, int *ok) {
    int N = g->nd + 1;
    *ok = 1;
    if (weight_out) *weight_out = 0;
    if (n == 0) return 0;
    int m = n + (n & 1); /* extra boundary node when odd */
    ll maxw = 0;
    /* true weights */
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            ll w;
            if (i == j) w = 0;
            else if (i >= n || j >= n) {
                int r = i >= n ? j : i;
                w = g->dist[(size_t)def[r] * N + g->nd];
            } else {
                ll dd = g->dist[(size_t)def[i] * N + def[j]];
                ll bb = (ll)g->dist[(size_t)def[i] * N + g->nd] + g->dist[(size_t)def[j] * N + g->nd];
                w = dd < bb ? dd : bb;
            }
            wbuf[(size_t)i * m + j] = w;
            if (i != j && w < UNREACH && w > maxw) maxw = w;
        }
    ll big = (maxw + 1) * (m + 2);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            ll w = wbuf[(size_t)i * m + j];
            wbuf[(size_t)i * m + j] = (i == j || w >= UNREACH) ? 0 : big - w;
        }
    blossom_solve(B, m, wbuf, mate);
    uint8_t pred = 0;
    ll tot = 0;
    for (int i = 0; i < n; ++i) {
        int j = mate[i];
        if (j < 0) { *ok = 0; continue; }
        if (pairs_out) { pairs_out[2 * i] = def[i]; pairs_out[2 * i + 1] = -1; }
        if (j >= n) { /* the odd one out goes to the boundary */
            pred ^= g->par[(size_t)def[i] * N + g->nd];
            tot += g->dist[(size_t)def[i] * N + g->nd];
            continue;
        }
        ll dd = g->dist[(size_t)def[i] * N + def[j]];
        ll bi = g->dist[(size_t)def[i] * N + g->nd], bj = g->dist[(size_t)def[j] * N + g->nd];
        if (dd < bi + bj) {
            if (pairs_out) pairs_out[2 * i + 1] = def[j];
            if (i < j) { pred ^= g->par[(size_t)def[i] * N + def[j]]; tot += dd; }
        } else { /* both to the boundary */
            pred ^= g->par[(size_t)def[i] * N + g->nd];
            tot += bi;
        }
    }
    if (weight_out) *weight_out = tot;
    return pred;
}

typedef struct {
    OrcGraph *g;
    const uint8_t *dets; /* shot-major, row_bytes per shot, little-endian bits */
    int row_bytes;
    long s0, s1;
    uint8_t *pred;
    ll *weights;   /* optional */
    int *fails;
    int max_defects;
} Job;

static void *job_run(void *arg) {
    Job *J = arg;
    OrcGraph *g = J->g;
    int cap = J->max_defects + 2;
    Blossom *B = blossom_new(cap);
    ll *wbuf = malloc(sizeof(ll) * (size_t)cap * cap);
    int *mate = malloc(sizeof(int) * cap);
    int *def = malloc(sizeof(int) * cap);
    int nfail = 0;
    for (long s = J->s0; s < J->s1; ++s) {
        const uint8_t *row = J->dets + (size_t)s * J->row_bytes;
        int n = 0;
        for (int k = 0; k < g->nd; ++k)
            if (row[k >> 3] >> (k & 7) & 1) def[n++] = k;
        int ok;
        ll wt;
        J->pred[s] = decode_defects(g, B, wbuf, mate, def, n, &wt, NULL, &ok);
        if (J->weights) J->weights[s] = wt;
        if (!ok) ++nfail;
    }
    *J->fails = nfail;
    blossom_free(B); free(wbuf); free(mate); free(def);
    return NULL;
}

/* dets: uint8[nshots][row_bytes] (sinter b8 layout).  pred: uint8[nshots] observable masks.
 * weights (optional): matching weight per shot.  Returns number of shots without a perfect matching. */
long orc_decode_batch(OrcGraph *g, const uint8_t *dets, long nshots, int row_bytes, uint8_t *pred,
                      ll *weights, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    /* the blossom workspace is sized by the largest defect count in the batch */
    int maxdef = 0;
    for (long s = 0; s < nshots; ++s) {
        const uint8_t *row = dets + (size_t)s * row_bytes;
        int n = 0;
        for (int k = 0; k < row_bytes; ++k) n += __builtin_popcount(row[k]);
        if (n > maxdef) maxdef = n;
    }
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    Job *jobs = malloc(sizeof(Job) * nthreads);
    int *fails = calloc(nthreads, sizeof(int));
    long per = (nshots + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        long a = t * per, b = a + per;
        if (a > nshots) a = nshots;
        if (b > nshots) b = nshots;
        jobs[t] = (Job){g, dets, row_bytes, a, b, pred, weights, &fails[t], maxdef};
        pthread_create(&th[t], NULL, job_run, &jobs[t]);
    }
    long nf = 0;
    for (int t = 0; t < nthreads; ++t) { pthread_join(th[t], NULL); nf += fails[t]; }
    free(th); free(jobs); free(fails);
    return nf;
}

/* single shot with the matching spelled out: pairs_out = int[2*n] */
int orc_decode_one(OrcGraph *g, const int *defects, int n, int *pairs_out, ll *weight_out) {
    int cap = n + 2;
    Blossom *B = blossom_new(cap);
    ll *wbuf = malloc(sizeof(ll) * (size_t)cap * cap);
    int *mate = malloc(sizeof(int) * cap);
    int ok;
    uint8_t p = decode_defects(g, B, wbuf, mate, defects, n, weight_out, pairs_out, &ok);
    blossom_free(B); free(wbuf); free(mate);
    return ok ? (int)p : -1;
}
