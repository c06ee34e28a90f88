// K3a: batched exact matching decoder, one shot per warp.
//
// Set-up (host, once per DEM): graphlike pieces of the DEM are merged into the matching graph with
// PyMatching's conventions (parallel edges merge with p1(1-p2)+p2(1-p1), the first observable
// mask is kept, weight ln((1-p)/p) discretised to integers), then all-pairs shortest paths (one Dijkstra per
// node, one more from the boundary) give the tables resident in HBM/L2: distance (u16), observable mask of that
// path (u8), the path's next hop (u16, only if corrections are wanted) and, per detector, its nearest detectors.
//
// First tier (every shot of the surface-code sweeps): the producer of the syndromes (sampler, b8 re-stride kernel,
// or classify_kernel) sorts the shots of a slice by defect count into per-K lists (K = ceil(defects / 32) slots per
// lane); per class, tree_start_kernel<K> extracts the defects and computes a tight dual start and
// tree_solve_kernel<K> runs single-tree primal-dual blossom phases straight on the distance table (match_tree.cuh:
// no per-shot weight matrix), both claiming shots through an atomic cursor.  Shots with more than 1024 defects, or
// that the solver gives up on, climb the older dense-matrix ladder (match_fast_kernel<K>: register-resident solver
// on a gathered weight matrix, blossom_regs.cuh; match_kernel: memory-resident, blossom_warp.cuh).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <queue>
#include <thread>

#ifdef GQ_STATS
__device__ unsigned long long gq_stats[16];
#endif
#include "blossom_regs.cuh"
#include "blossom_warp.cuh"
#include "gq_common.h"
#include "match_tree.cuh"

#ifndef GQ_PER_SM
#define GQ_PER_SM 4
#endif

struct MatchArgs {
    const uint32_t *dets;
    uint32_t *pred;
    int32_t *weight_out;
    uint8_t *flags_out;
    uint32_t *corr_out;
    const uint32_t *list;     // optional shot list (overflow pass)
    uint64_t nitems;
    uint32_t D, DW, OW, EW;
    uint32_t N;               // row stride of the dist / pobs / next tables = D + 2 (D = boundary, D + 1 = nothing)
    uint32_t cap;             // even; real defects handled: cap - 1
    const uint16_t *dist;
    const uint8_t *pobs;
    const uint16_t *next;
    const uint32_t *nbr;      // [N][gqt::NBR_LEN] nearest detectors of each detector (match_tree.cuh)
    const uint32_t *adj_ptr, *adj_node, *adj_edge;
    char *ws;
    size_t ws_stride;
    uint32_t *overflow;       // [0] count, [1] max defects, [2..] shots
    uint32_t overflow_cap;
    uint32_t *cursor;         // first tier: next unclaimed item of the list (dynamic work distribution)
};

__device__ __forceinline__ void walk_path(const MatchArgs &a, uint32_t cur, uint32_t target, uint32_t *corr_row) {
    const uint32_t N = a.N;
    for (uint32_t guard = 0; guard <= a.D && cur != target; ++guard) {
        uint32_t nx = a.next[(size_t)target * N + cur];
        if (nx == 0xFFFFu) nx = a.D;  // hop onto the boundary
        uint32_t e = 0xFFFFFFFFu;
        for (uint32_t q = a.adj_ptr[cur]; q < a.adj_ptr[cur + 1]; ++q)
            if (a.adj_node[q] == nx) { e = a.adj_edge[q]; break; }
        if (e == 0xFFFFFFFFu) return;
        atomicXor(&corr_row[e >> 5], 1u << (e & 31));
        cur = nx;
    }
}

__global__ void __launch_bounds__(256) match_kernel(MatchArgs a) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t N = a.N;
    const int cap = (int)a.cap;
    char *mem = a.ws + warp * a.ws_stride;
    gqb::Solver S;
    gqb::ws_bind(S.w, mem, cap);
    S.cap = cap; S.nb = cap / 2;
    uint16_t *def = (uint16_t *)(mem + gqb::ws_bytes(cap));
    uint16_t *bd = def + cap;
    const uint16_t *brow = a.dist + (size_t)a.D * N;   // distances to the boundary
    const uint8_t *bobs = a.pobs + (size_t)a.D * N;

    for (uint64_t item = warp; item < a.nitems; item += nwarps) {
        const uint64_t shot = a.list ? a.list[item] : item;
        const uint32_t *row = a.dets + shot * a.DW;
        // ---- defect list (ascending detector index)
        int n0 = 0;
        for (uint32_t w0 = 0; w0 < a.DW; w0 += 32) {
            uint32_t wi = w0 + lane;
            uint32_t word = wi < a.DW ? row[wi] : 0u;
            if (wi == a.DW - 1 && (a.D & 31)) word &= (1u << (a.D & 31)) - 1;
            int cnt = __popc(word);
            int pre = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, pre, o);
                if (lane >= o) pre += t;
            }
            int tot = __shfl_sync(0xffffffffu, pre, 31);
            int pos = n0 + pre - cnt;
            while (word) {
                int b = __ffs(word) - 1;
                word &= word - 1;
                if (pos < cap) def[pos] = (uint16_t)(wi * 32 + b);
                ++pos;
            }
            n0 += tot;
        }
        __syncwarp();
        if (n0 > cap - 1) {
            if (lane == 0 && a.overflow) {
                uint32_t idx = atomicAdd(&a.overflow[0], 1u);
                atomicMax(&a.overflow[1], (uint32_t)n0);
                if (idx < a.overflow_cap) a.overflow[2 + idx] = (uint32_t)shot;
            }
            if (lane == 0 && !a.overflow && a.flags_out) a.flags_out[shot] = 2;
            continue;
        }
        uint32_t pm = 0;
        int wt = 0;
        int rc = 0;
        if (n0 > 0) {
            const int n = n0 + (n0 & 1);
            for (int i = lane; i < n0; i += 32) bd[i] = brow[def[i]];
            __syncwarp();
            for (int i = 0; i < n0; ++i) {
                const uint32_t di = def[i];
                const uint32_t bi = bd[i];
                const uint16_t *drow = a.dist + (size_t)di * N;
                uint16_t *wrow = S.w.W + i * S.w.wstride;
                for (int j = lane; j < n0; j += 32) {
                    uint32_t d = drow[def[j]];
                    uint32_t bj = bd[j];
                    uint32_t s = (bi == 0xFFFFu || bj == 0xFFFFu) ? 0xFFFFu : min(bi + bj, 0xFFFEu);
                    wrow[j] = (uint16_t)min(d, s);
                }
                if (n != n0 && lane == 0) {
                    wrow[n0] = (uint16_t)bi;
                    S.w.W[n0 * S.w.wstride + i] = (uint16_t)bi;
                }
            }
            if (n != n0 && lane == 0) S.w.W[n0 * S.w.wstride + n0] = 0xFFFFu;
            __syncwarp();
            S.n = n;
            rc = S.solve();
            __syncwarp();
            if (rc == 0) {
                uint32_t *corr_row = a.corr_out ? a.corr_out + shot * a.EW : nullptr;
                for (int i = lane; i < n0; i += 32) {
                    int j = S.w.mate[i];
                    const uint32_t di = def[i];
                    if (j == n0) {  // paired with the virtual boundary vertex
                        pm ^= bobs[di];
                        wt += bd[i];
                        if (corr_row) walk_path(a, di, a.D, corr_row);
                    } else if (j > i) {
                        const uint32_t dj = def[j];
                        uint32_t d = a.dist[(size_t)dj * N + di];
                        uint32_t s = (bd[i] == 0xFFFFu || bd[j] == 0xFFFFu) ? 0xFFFFu : (uint32_t)bd[i] + bd[j];
                        if (d <= s) {
                            pm ^= a.pobs[(size_t)dj * N + di];
                            wt += (int)d;
                            if (corr_row) walk_path(a, di, dj, corr_row);
                        } else {
                            pm ^= bobs[di] ^ bobs[dj];
                            wt += (int)s;
                            if (corr_row) { walk_path(a, di, a.D, corr_row); walk_path(a, dj, a.D, corr_row); }
                        }
                    }
                }
                pm = __reduce_xor_sync(0xffffffffu, pm);
                wt = __reduce_add_sync(0xffffffffu, wt);
            }
        }
        if (lane == 0) {
            a.pred[shot * a.OW] = rc == 0 ? pm : 0u;
            for (uint32_t w = 1; w < a.OW; ++w) a.pred[shot * a.OW + w] = 0u;
            if (a.weight_out) a.weight_out[shot] = rc == 0 ? wt : -1;
            if (a.flags_out) a.flags_out[shot] = (uint8_t)(rc == 0 ? 0 : (rc == 1 ? 1 : 2));
        }
        __syncwarp();
    }
}


// ---- fast tier: register-resident solver (blossom_regs.cuh), forest in shared memory, W in an
// L2-resident scratch; handles shots with fewer than 32*K defects, the rest go to the overflow list
template <int K>
__global__ void __launch_bounds__(256, (K <= 4 ? GQ_PER_SM : 2)) match_fast_kernel(MatchArgs a) {
    extern __shared__ __align__(16) char smem[];
    constexpr int CAP = 32 * K;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t N = a.N;
    const size_t per_warp = gqr::forest_bytes(CAP) + (size_t)CAP * 4;
    char *mem = smem + (size_t)wib * per_warp;
    const gqr::Forest F = gqr::forest_bind(mem, CAP);
    uint16_t *def = (uint16_t *)(mem + gqr::forest_bytes(CAP));
    uint16_t *bd = def + CAP;
    uint16_t *Wm = (uint16_t *)(a.ws + warp * a.ws_stride);
    const uint16_t *brow = a.dist + (size_t)a.D * N;
    const uint8_t *bobs = a.pobs + (size_t)a.D * N;

    for (uint64_t item = warp; item < a.nitems; item += nwarps) {
        const uint64_t shot = a.list ? a.list[item] : item;
        const uint32_t *row = a.dets + shot * a.DW;
        int n0 = 0;
        for (uint32_t w0 = 0; w0 < a.DW; w0 += 32) {
            uint32_t wi = w0 + lane;
            uint32_t word = wi < a.DW ? row[wi] : 0u;
            if (wi == a.DW - 1 && (a.D & 31)) word &= (1u << (a.D & 31)) - 1;
            int cnt = __popc(word);
            int pre = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, pre, o);
                if (lane >= o) pre += t;
            }
            int tot = __shfl_sync(0xffffffffu, pre, 31);
            int pos = n0 + pre - cnt;
            while (word) {
                int b = __ffs(word) - 1;
                word &= word - 1;
                if (pos < CAP) def[pos] = (uint16_t)(wi * 32 + b);
                ++pos;
            }
            n0 += tot;
        }
        __syncwarp();
        if (n0 > CAP - 1) {
            if (lane == 0 && a.overflow) {
                uint32_t idx = atomicAdd(&a.overflow[0], 1u);
                atomicMax(&a.overflow[1], (uint32_t)n0);
                if (idx < a.overflow_cap) a.overflow[2 + idx] = (uint32_t)shot;
            }
            continue;
        }
        uint32_t pm = 0;
        int wt = 0, rc = 0;
        if (n0 > 0) {
            // native-boundary form: W[i][j] = d(i, j) only, boundary distances travel separately
            const int n = n0;
            const int stride = n + 1;
            uint32_t dj[K], bj[K], ymin[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int j = lane + 32 * k;
                dj[k] = j < n0 ? def[j] : 0u;
                bj[k] = j < n0 ? brow[dj[k]] : 0xFFFFu;
                if (j < n0) bd[j] = (uint16_t)bj[k];
                ymin[k] = 0xFFFFu;
            }
            __syncwarp();
            for (int i = 0; i < n0; ++i) {
                const uint32_t di = def[i];
                const uint16_t *drow = a.dist + (size_t)di * N;
                uint16_t *wrow = Wm + i * stride;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int j = lane + 32 * k;
                    if (j < n0) {
                        uint32_t w = j == i ? 0xFFFFu : (uint32_t)drow[dj[k]];
                        wrow[j] = (uint16_t)w;
                        ymin[k] = min(ymin[k], w);
                    }
                }
            }
            __syncwarp();
            rc = gqr::solve_nb<K>(F, Wm, stride, n, lane, ymin, bj);
            __syncwarp();
            if (rc == 0) {
                uint32_t *corr_row = a.corr_out ? a.corr_out + shot * a.EW : nullptr;
                int bad = 0;
                for (int i = lane; i < n0; i += 32) {
                    int j = F.mate[i];
                    const uint32_t di = def[i];
                    if (j == -2) {                       // matched to the boundary
                        if (bd[i] == 0xFFFFu) bad = 1;
                        pm ^= bobs[di];
                        wt += bd[i];
                        if (corr_row) walk_path(a, di, a.D, corr_row);
                    } else if (j < 0) {
                        bad = 1;
                    } else if (j > i) {
                        const uint32_t dj2 = def[j];
                        uint32_t d = a.dist[(size_t)dj2 * N + di];
                        if (d >= 0xFFFFu) bad = 1;       // matched through a missing edge: no perfect matching
                        pm ^= a.pobs[(size_t)dj2 * N + di];
                        wt += (int)d;
                        if (corr_row) walk_path(a, di, dj2, corr_row);
                    }
                }
                pm = __reduce_xor_sync(0xffffffffu, pm);
                wt = __reduce_add_sync(0xffffffffu, wt);
                if (__any_sync(0xffffffffu, bad)) rc = 1;
            }
        }
        if (lane == 0) {
            a.pred[shot * a.OW] = rc == 0 ? pm : 0u;
            for (uint32_t w = 1; w < a.OW; ++w) a.pred[shot * a.OW + w] = 0u;
            if (a.weight_out) a.weight_out[shot] = rc == 0 ? wt : -1;
            if (a.flags_out) a.flags_out[shot] = (uint8_t)(rc == 0 ? 0 : (rc == 1 ? 1 : 2));
        }
        __syncwarp();
    }
}

// ---- first tier: single-tree blossom phases on the defect set, weights gathered straight from the
// distance table (match_tree.cuh).  The solver keeps K = ceil(n / 32) vertices per lane in registers,
// and the kernels are instruction-fetch bound, so every K gets its own kernel (a fused switch over K
// ran 2.4x slower: four solver bodies thrash the instruction cache) and classify_kernel sorts the
// shots of a slice into per-K lists first.  Shots with more than MT_CAP defects, and any shot the
// solver gives up on, go to the overflow list and climb the ladder of the older tiers.
#ifndef GQ_MT_PER_SM
#define GQ_MT_PER_SM 4
#endif
#ifndef GQ_SLICE_LOG2
#define GQ_SLICE_LOG2 22   // shots per decode slice (lists, hand-off records and launch overheads are per slice)
#endif
#ifndef GQ_MT_PER_SM_K4
#define GQ_MT_PER_SM_K4 4
#endif
#ifndef GQ_START_PER_SM
#define GQ_START_PER_SM 5   // start-kernel blocks per SM for K <= 4 (51 registers; measured: 8 -> 36.75, 6 -> 36.45, 5 -> 36.35 ms)
#endif
#ifndef GQ_MT_OVERSUB
#define GQ_MT_OVERSUB 1   // solve-kernel blocks per resident slot (> 1: the block scheduler balances the tail)
#endif
#ifndef GQ_MT_PER_SM_K56
#define GQ_MT_PER_SM_K56 2
#endif
#ifndef GQ_SIDE_STREAMS
#define GQ_SIDE_STREAMS 0   // 1: the smaller classes' kernels run on side streams (overlapping tails, but co-resident
                            // kernels share the instruction cache: measured slower once start/solve were split)
#endif
// classes 0..7: K = 1..8 (8-bit vertex field, up to 256 defects); 8..10: K = 10, 12, 16 (9-bit field, up to 512);
// 11, 12: K = 24, 32 (10-bit field, up to 1024; part of the per-lane state lives in local memory there);
// 13, 14: K = 48, 64 (11-bit field, up to 2048: d = 23 up to p = 1e-2, the reference example's largest point; only when
// the event times fit the 18-bit field that leaves, gq_dec::big_classes)
static constexpr int MT_CLASSES = GQ_MT_CLASSES;
static constexpr int MT_CAP = GQ_MT_CAP;
__host__ __device__ constexpr int mt_class_of(int n) { return gq_mt_class_of(n); }   // n >= 1 defects -> class, MT_CLASSES = beyond the first tier
template <int K> struct MtClassRange {     // defect counts a class-K list may hold: (LO, 32 K]
    static constexpr int LO = K <= 8 ? 32 * (K - 1) : (K == 10 ? 256 : (K == 12 ? 320 : (K == 16 ? 384 : (K == 24 ? 512 : (K == 32 ? 768 : (K == 48 ? 1024 : 1536))))));
    static constexpr int WPB = K <= 16 ? 8 : (K <= 32 ? 4 : 2);   // warps per block: the forests are 30 .. 41 KB per warp at K = 24, 32 and 61 .. 82 KB at K = 48, 64
};

#ifndef GQ_DYNAMIC
#define GQ_DYNAMIC 1   // 1: warps claim list items through an atomic cursor (shot times vary by an order of magnitude, so a
                       // static grid-stride split leaves the SMs idle while the unluckiest warp finishes: sm__warps_active
                       // was 88 % of its bound); 0: static grid-stride
#endif
// next item of a first-tier list for this warp; the first claim is the warp's own index (no atomic)
__device__ __forceinline__ uint32_t mt_next_item(uint32_t *cursor, uint32_t nwarps, uint32_t item, int lane) {
#if GQ_DYNAMIC
    uint32_t nx = 0;
    if (lane == 0) nx = nwarps + atomicAdd(cursor, 1u);
    return __shfl_sync(0xffffffffu, nx, 0);
#else
    return item + nwarps;
#endif
}

template <int K>
struct MtWarpMem {
    gqr::StaticForest<32 * K> F;
    uint16_t def[32 * K];
};

// counts[0..MT_CLASSES-1]: shots of class K = 1..MT_CLASSES, lists[c][...]: their indices; defect-free shots are answered
// here; overflow (> MT_CAP defects) is appended to the ladder's list
struct ClassifyArgs {
    const uint32_t *dets;
    uint32_t *pred;
    int32_t *weight_out;
    uint8_t *flags_out;
    uint32_t *counts;          // [MT_CLASSES]
    uint32_t *lists;           // [MT_CLASSES][list_stride]
    uint32_t list_stride;
    uint32_t *overflow;
    uint32_t overflow_cap;
    uint32_t nshots, D, DW, OW;
};

__global__ void __launch_bounds__(256) classify_kernel(ClassifyArgs a) {
    const int lane = threadIdx.x & 31;
    for (uint32_t base = blockIdx.x * blockDim.x; base < a.nshots; base += gridDim.x * blockDim.x) {
        const uint32_t shot = base + threadIdx.x;
        int n = -1;
        if (shot < a.nshots) {
            const uint32_t *row = a.dets + (size_t)shot * a.DW;
            n = 0;
            for (uint32_t w = 0; w + 1 < a.DW; ++w) n += __popc(row[w]);
            uint32_t last = row[a.DW - 1];
            if (a.D & 31) last &= (1u << (a.D & 31)) - 1;
            n += __popc(last);
        }
        const int cls = n < 0 ? -1 : (n == 0 ? -2 : mt_class_of(n));
        if (cls == -2) {
            for (uint32_t w = 0; w < a.OW; ++w) a.pred[(size_t)shot * a.OW + w] = 0u;
            if (a.weight_out) a.weight_out[shot] = 0;
            if (a.flags_out) a.flags_out[shot] = 0;
        }
#pragma unroll
        for (int c = 0; c <= MT_CLASSES; ++c) {
            const unsigned m = __ballot_sync(0xffffffffu, cls == c);
            if (m == 0) continue;
            uint32_t pos = 0;
            if (lane == __ffs(m) - 1) pos = atomicAdd(c < MT_CLASSES ? &a.counts[c] : &a.overflow[0], (uint32_t)__popc(m));
            pos = __shfl_sync(0xffffffffu, pos, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1));
            if (cls == c) {
                if (c < MT_CLASSES) a.lists[(size_t)c * a.list_stride + pos] = shot;
                else if (pos < a.overflow_cap) a.overflow[2 + pos] = shot;
            }
        }
        if (cls == MT_CLASSES) atomicMax(&a.overflow[1], (uint32_t)n);
    }
}

// The first tier runs as two kernels per class so that each one's hot code fits the 32 KB L1.5
// instruction cache (the fused kernel ran at an 80 % i-cache hit rate, 3 of 12 stall cycles per
// instruction in no_instruction): tree_start_kernel<K> extracts the defects and computes the dual start,
// tree_solve_kernel<K> runs the blossom phases and the prediction.  Hand-off through a per-shot record in
// global memory (6 bytes per defect, written and read once, coalesced).
template <int K>
struct StartRecord {
    uint16_t def[32 * K];     // detector index of every defect
    uint16_t yh[32 * K];      // dual / 2 (duals are even, quadrupled units)
    int16_t mate[32 * K];     // initial mate: -1 free, -2 boundary, else partner
    uint32_t n, pad[3];
};

// Per-warp shared memory of the start kernel, laid out at run time so that it is as small as the syndrome allows
// (what is not shared memory is L1 for the nearest-detector lists and the distance rows):
//   def u16[32 K] | xch i16[32 K] | synw u32[SW + 1] | spre u16[SW + 2]      (SW = DW when the syndrome is mirrored, else 0)
template <int K>
static size_t start_warp_bytes(uint32_t DW) {
    const uint32_t sw = DW <= (uint32_t)gqt::SYN_WORDS ? DW : 0u;
    return ((size_t)32 * K * 4 + ((size_t)sw + 1) * 4 + ((size_t)sw + 2) * 2 + 15) & ~(size_t)15;
}

template <int K>
__global__ void __launch_bounds__(32 * MtClassRange<K>::WPB, (K <= 4 ? GQ_START_PER_SM : (K <= 6 ? 4 : (K <= 8 ? 3 : 1)))) tree_start_kernel(MatchArgs a) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const bool mirror = a.nbr != nullptr && a.DW <= (uint32_t)gqt::SYN_WORDS;
    const uint32_t sw = mirror ? a.DW : 0u;
    char *wbase = smem + (size_t)wib * ((32u * K * 4 + (sw + 1) * 4 + (sw + 2) * 2 + 15u) & ~15u);
    uint16_t *m_def = reinterpret_cast<uint16_t *>(wbase);
    int16_t *m_xch = reinterpret_cast<int16_t *>(wbase + 32 * K * 2);
    uint32_t *m_synw = reinterpret_cast<uint32_t *>(wbase + 32 * K * 4);
    uint16_t *m_spre = reinterpret_cast<uint16_t *>(wbase + 32 * K * 4 + (sw + 1) * 4);
    StartRecord<K> *recs = reinterpret_cast<StartRecord<K> *>(a.ws);
    gqt::Tables T;
    T.dist = a.dist; T.pobs = a.pobs; T.D = a.D; T.N = a.N; T.nbr = a.nbr;
    for (uint32_t item = warp; item < a.nitems; item = mt_next_item(a.cursor, nwarps, item, lane)) {
        const uint32_t shot = a.list[item];
        const int n0 = gqt::extract_defects(a.dets + (size_t)shot * a.DW, a.DW, a.D, lane, m_def, 32 * K,
                                            mirror ? m_synw : nullptr, m_spre);
        StartRecord<K> &R = recs[item];
        if (n0 > 32 * K || n0 <= MtClassRange<K>::LO) {   // the list lied about the defect count: the solver will hand it on
            if (lane == 0) R.n = 0xffffffffu;
            __syncwarp();
            continue;
        }
        gqt::i32 yb[K];
        int m0[K];
        gqt::dual_start<K>(T, m_xch, m_def, n0, lane, mirror ? m_synw : nullptr, m_spre, yb, m0);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int v = lane + 32 * k;
            if (v < n0) { R.def[v] = m_def[v]; R.yh[v] = (uint16_t)(yb[k] >> 1); R.mate[v] = (int16_t)m0[k]; }
        }
        if (lane == 0) R.n = (uint32_t)n0;
        __syncwarp();
    }
}

// EXTRAS = weights / flags / corrections requested (tests, sinter's syndrome-reproduction checks): the
// throughput instantiation leaves that code out, which is worth it because the kernel is at the edge of
// the instruction cache
template <int K, bool EXTRAS>
__global__ void __launch_bounds__(32 * MtClassRange<K>::WPB, (K <= 3 ? GQ_MT_PER_SM : (K == 4 ? GQ_MT_PER_SM_K4 : (K <= 6 ? GQ_MT_PER_SM_K56 : 1)))) tree_solve_kernel(MatchArgs a) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    MtWarpMem<K> &M = reinterpret_cast<MtWarpMem<K> *>(smem)[wib];
    const StartRecord<K> *recs = reinterpret_cast<const StartRecord<K> *>(a.ws);
    gqt::Tables T;
    T.dist = a.dist; T.pobs = a.pobs; T.D = a.D; T.N = a.N; T.nbr = a.nbr;
    const bool want_weight = EXTRAS && a.weight_out != nullptr;

    for (uint32_t item = warp; item < a.nitems; item = mt_next_item(a.cursor, nwarps, item, lane)) {
        const uint32_t shot = a.list[item];
        const StartRecord<K> &R = recs[item];
        const uint32_t n0u = R.n;
        int rc = 4, wt = 0, n0 = (int)n0u;
        uint32_t pm = 0;
        if (n0u != 0xffffffffu) {
            gqt::i32 yb[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int v = lane + 32 * k;
                yb[k] = 0;
                if (v < n0) { M.def[v] = R.def[v]; M.F.mate[v] = R.mate[v]; yb[k] = 2 * (gqt::i32)R.yh[v]; }
            }
            __syncwarp();
            rc = gqt::solve_predict<K>(T, M.F, M.def, n0, lane, want_weight, yb, &pm, &wt);
        } else n0 = 32 * K;
        if (rc >= 2) {
            // the solver gave up (or the list lied about the defect count): hand the shot to the ladder
            if (lane == 0) {
                uint32_t idx = atomicAdd(&a.overflow[0], 1u);
                atomicMax(&a.overflow[1], (uint32_t)n0);
                if (idx < a.overflow_cap) a.overflow[2 + idx] = shot;
            }
            __syncwarp();
            continue;
        }
        if (EXTRAS && rc == 0 && a.corr_out) {
            uint32_t *corr_row = a.corr_out + (size_t)shot * a.EW;
#pragma unroll 1
            for (int i = lane; i < n0; i += 32) {
                const int j = M.F.mate[i];
                if (j == -2) walk_path(a, M.def[i], a.D, corr_row);
                else if (j > i) walk_path(a, M.def[i], M.def[j], corr_row);
            }
        }
        if (lane == 0) {
            a.pred[(size_t)shot * a.OW] = rc == 0 ? pm : 0u;
#pragma unroll 1
            for (uint32_t w = 1; w < a.OW; ++w) a.pred[(size_t)shot * a.OW + w] = 0u;
            if (EXTRAS && a.weight_out) a.weight_out[shot] = rc == 0 ? wt : -1;
            if (EXTRAS && a.flags_out) a.flags_out[shot] = (uint8_t)(rc == 0 ? 0 : 1);
        }
        __syncwarp();
    }
}

// classes 13, 14 of a decoder whose event times do not fit the 18-bit field: their shots join the overflow list
__global__ void append_list_to_overflow_kernel(const uint32_t *list, uint32_t n, uint32_t *overflow, uint32_t overflow_cap, uint32_t maxdef) {
    __shared__ uint32_t base;
    if (threadIdx.x == 0) { base = atomicAdd(&overflow[0], n); atomicMax(&overflow[1], maxdef); }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        if (base + i < overflow_cap) overflow[2 + base + i] = list[i];
}

template <int K>
static int launch_tree_class(MatchArgs b, uint32_t *cursors, gq_dem *dem, cudaStream_t st, int *launches, cudaEvent_t ev0, cudaEvent_t ev1) {
    constexpr int WPB = MtClassRange<K>::WPB;
    b.cursor = cursors;   // [0]: start kernel, [16]: solve kernel (zeroed with the class counts)
    const int blocks_a = (int)std::min<uint64_t>((b.nitems + WPB - 1) / WPB, (uint64_t)dem->num_sms * 8);   // latency-bound: all 64 warps/SM
    tree_start_kernel<K><<<blocks_a, 32 * WPB, WPB * start_warp_bytes<K>(b.DW), st>>>(b);
    if (cudaGetLastError() != cudaSuccess) return GQ_E_CUDA;
    const int per_sm = K <= 3 ? GQ_MT_PER_SM : (K == 4 ? GQ_MT_PER_SM_K4 : (K <= 6 ? GQ_MT_PER_SM_K56 : 1));
    const int blocks_b = (int)std::min<uint64_t>((b.nitems + WPB - 1) / WPB, (uint64_t)dem->num_sms * per_sm * GQ_MT_OVERSUB);
    if (ev0) cudaEventRecord(ev0, st);   // bench.py's roofline times the solve kernel of the biggest class
    b.cursor = cursors + 16;
    if (b.weight_out || b.flags_out || b.corr_out) tree_solve_kernel<K, true><<<blocks_b, 32 * WPB, WPB * sizeof(MtWarpMem<K>), st>>>(b);
    else tree_solve_kernel<K, false><<<blocks_b, 32 * WPB, WPB * sizeof(MtWarpMem<K>), st>>>(b);
    if (cudaGetLastError() != cudaSuccess) return GQ_E_CUDA;
    if (ev1) cudaEventRecord(ev1, st);
    *launches += 2;
    return GQ_OK;
}

template <int K>
static int configure_tree_class() {
    if (cudaFuncSetAttribute(tree_start_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MtClassRange<K>::WPB * start_warp_bytes<K>(gqt::SYN_WORDS))) != cudaSuccess) return GQ_E_CUDA;
    if (cudaFuncSetAttribute(tree_solve_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MtClassRange<K>::WPB * sizeof(MtWarpMem<K>))) != cudaSuccess) return GQ_E_CUDA;
    if (cudaFuncSetAttribute(tree_solve_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MtClassRange<K>::WPB * sizeof(MtWarpMem<K>))) != cudaSuccess) return GQ_E_CUDA;
    return GQ_OK;
}

static size_t start_record_bytes(int c) {   // class c = K - 1
    switch (c) {
    case 0: return sizeof(StartRecord<1>);
    case 1: return sizeof(StartRecord<2>);
    case 2: return sizeof(StartRecord<3>);
    case 3: return sizeof(StartRecord<4>);
    case 4: return sizeof(StartRecord<5>);
    case 5: return sizeof(StartRecord<6>);
    case 6: return sizeof(StartRecord<7>);
    case 7: return sizeof(StartRecord<8>);
    case 8: return sizeof(StartRecord<10>);
    case 9: return sizeof(StartRecord<12>);
    case 10: return sizeof(StartRecord<16>);
    case 11: return sizeof(StartRecord<24>);
    case 12: return sizeof(StartRecord<32>);
    case 13: return sizeof(StartRecord<48>);
    default: return sizeof(StartRecord<64>);
    }
}

// ------------------------------------------------------------------------------------------------
// host: graph + tables
// ------------------------------------------------------------------------------------------------
static size_t warp_ws_bytes(int cap) { return (gqb::ws_bytes(cap) + (size_t)cap * 4 + 15) & ~(size_t)15; }
static int even_up(int x) { return (x + 1) & ~1; }

int gq_matching_create(gq_dec *dec) {
    gq_dem *dem = dec->dem;
    const uint32_t D = dem->D, N = D + 1;
    if (dem->O > 8) GQ_FAIL(GQ_E_UNSUPPORTED, "matching decoder: more than 8 observables are not supported");
    if (D >= 0xFFFEu) GQ_FAIL(GQ_E_UNSUPPORTED, "matching decoder: too many detectors");
    if (D == 0) GQ_FAIL(GQ_E_INVALID, "matching decoder: the DEM has no detectors");
    // ---- graphlike pieces -> merged edges
    std::map<std::pair<uint32_t, uint32_t>, std::pair<double, uint64_t>> table;
    auto add_piece = [&](uint32_t a, uint32_t b, uint64_t om, double p) {
        if (a == GQ_NO_DET) return;
        if (b != GQ_NO_DET && b < a) std::swap(a, b);
        if (b != GQ_NO_DET && a == b) return;
        auto key = std::make_pair(a, b);
        auto it = table.find(key);
        if (it == table.end()) table[key] = std::make_pair(p, om);
        else it->second.first = it->second.first * (1 - p) + p * (1 - it->second.first);
    };
    for (uint32_t m = 0; m < dem->M; ++m) {
        if (dem->has_components) {
            for (uint32_t c = dem->comp_indptr[m]; c < dem->comp_indptr[m + 1]; ++c)
                add_piece(dem->comp_a[c], dem->comp_b[c], dem->comp_obs[c], dem->p[m]);
        } else {
            uint32_t b = dem->det_indptr[m], e = dem->det_indptr[m + 1];
            if (e - b < 1 || e - b > 2) continue;  // PyMatching ignores non-graphlike mechanisms
            uint64_t om = 0;
            for (uint32_t q = dem->obs_indptr[m]; q < dem->obs_indptr[m + 1]; ++q) om ^= 1ull << dem->obs_idx[q];
            add_piece(dem->det_idx[b], e - b == 2 ? dem->det_idx[b + 1] : GQ_NO_DET, om, dem->p[m]);
        }
    }
    int levels = dec->params.weight_levels > 0 ? dec->params.weight_levels : 1000;
    double wmax = 0;
    std::vector<double> wf;
    for (auto &kv : table) {
        double p = std::min(std::max(kv.second.first, 1e-300), 0.5);
        double w = log((1 - p) / p);
        wf.push_back(w);
        wmax = std::max(wmax, w);
        dec->eu.push_back(kv.first.first);
        dec->ev.push_back(kv.first.second);
        dec->eobs.push_back(kv.second.second);
        dec->eprob.push_back(kv.second.first);
        if (kv.first.second == GQ_NO_DET) dec->num_boundary++;
    }
    dec->num_edges = (uint32_t)dec->eu.size();
    auto set_weights = [&](int lv) {   // the heaviest edge maps to lv
        dec->ew.clear();
        for (double w : wf) {
            double v = wmax > 0 ? rint(w * (lv / wmax)) : 1.0;
            dec->ew.push_back((uint32_t)std::max(1.0, v));
        }
    };
    set_weights(levels);
    // expected defect count from the DEM (capacities below; and whether the 1025 .. 2048-defect classes matter)
    std::vector<double> keep(D, 1.0);
    for (uint32_t m = 0; m < dem->M; ++m)
        for (uint32_t q = dem->det_indptr[m]; q < dem->det_indptr[m + 1]; ++q) keep[dem->det_idx[q]] *= 1 - 2 * dem->p[m];
    double mu = 0;
    for (uint32_t d = 0; d < D; ++d) mu += 0.5 * (1 - keep[d]);
    // ---- adjacency (node D = boundary)
    std::vector<uint32_t> adj_ptr(N + 1, 0);
    for (uint32_t e = 0; e < dec->num_edges; ++e) {
        uint32_t a = dec->eu[e], b = dec->ev[e] == GQ_NO_DET ? D : dec->ev[e];
        adj_ptr[a + 1]++; adj_ptr[b + 1]++;
    }
    for (uint32_t i = 0; i < N; ++i) adj_ptr[i + 1] += adj_ptr[i];
    std::vector<uint32_t> adj_node(adj_ptr[N]), adj_edge(adj_ptr[N]);
    {
        std::vector<uint32_t> fill(adj_ptr.begin(), adj_ptr.end() - 1);
        for (uint32_t e = 0; e < dec->num_edges; ++e) {
            uint32_t a = dec->eu[e], b = dec->ev[e] == GQ_NO_DET ? D : dec->ev[e];
            adj_node[fill[a]] = b; adj_edge[fill[a]++] = e;
            adj_node[fill[b]] = a; adj_edge[fill[b]++] = e;
        }
    }
    // ---- all-pairs shortest paths: row s = Dijkstra tree rooted at s (row D: rooted at the boundary)
    const bool want_next = dec->params.with_corrections != 0;
    // tables have one more row / column than nodes: index D + 1 is "nothing" (all distances infinite),
    // the column that padding slots of the first tier read
    const uint32_t NT = D + 2;
    std::vector<uint16_t> dist((size_t)NT * NT, 0xFFFFu), nexttab;
    std::vector<uint8_t> pobs((size_t)NT * NT, 0);
    if (want_next) nexttab.assign((size_t)NT * NT, 0xFFFFu);
    std::vector<uint32_t> row_max(N, 0);
    std::atomic<bool> too_far(false);
    auto worker = [&](uint32_t s_begin, uint32_t s_end) {
        std::vector<uint64_t> d(N);
        std::vector<uint8_t> po(N);
        std::vector<uint32_t> pr(N);
        std::vector<char> done(N);
        typedef std::pair<uint64_t, uint32_t> QI;
        for (uint32_t s = s_begin; s < s_end; ++s) {
            std::fill(d.begin(), d.end(), ~0ull);
            std::fill(po.begin(), po.end(), 0);
            std::fill(pr.begin(), pr.end(), 0xFFFFFFFFu);
            std::fill(done.begin(), done.end(), 0);
            std::priority_queue<QI, std::vector<QI>, std::greater<QI>> pq;
            d[s] = 0;
            pq.push(QI(0, s));
            while (!pq.empty()) {
                QI it = pq.top(); pq.pop();
                uint32_t u = it.second;
                if (done[u]) continue;
                done[u] = 1;
                if (u == D && s != D) continue;  // the boundary is a sink, never a through-node
                for (uint32_t q = adj_ptr[u]; q < adj_ptr[u + 1]; ++q) {
                    uint32_t v = adj_node[q], e = adj_edge[q];
                    if (v == D && s == D) continue;
                    uint64_t nd = it.first + dec->ew[e];
                    if (nd < d[v]) {
                        d[v] = nd; po[v] = po[u] ^ (uint8_t)dec->eobs[e]; pr[v] = u;
                        pq.push(QI(nd, v));
                    }
                }
            }
            uint32_t mx = 0;
            for (uint32_t t = 0; t < N; ++t) {
                if (d[t] == ~0ull) continue;
                if (d[t] >= 0xFFFEull) { too_far = true; continue; }
                dist[(size_t)s * NT + t] = (uint16_t)d[t];
                pobs[(size_t)s * NT + t] = po[t];
                if (want_next && t != s) nexttab[(size_t)s * NT + t] = pr[t] == D ? 0xFFFFu : (uint16_t)pr[t];
                mx = std::max<uint32_t>(mx, (uint32_t)d[t]);
            }
            row_max[s] = mx;
        }
    };
    auto run_apsp = [&]() {
        std::fill(dist.begin(), dist.end(), (uint16_t)0xFFFFu);
        std::fill(pobs.begin(), pobs.end(), (uint8_t)0);
        if (want_next) std::fill(nexttab.begin(), nexttab.end(), (uint16_t)0xFFFFu);
        std::fill(row_max.begin(), row_max.end(), 0u);
        unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
        nt = std::min<unsigned>(nt, N);
        std::vector<std::thread> th;
        uint32_t per = (N + nt - 1) / nt;
        for (unsigned t = 0; t < nt; ++t) {
            uint32_t b = std::min<uint32_t>(N, t * per), e = std::min<uint32_t>(N, b + per);
            if (b < e) th.emplace_back(worker, b, e);
        }
        for (auto &t : th) t.join();
        dec->max_dist = *std::max_element(row_max.begin(), row_max.end());
    };
    run_apsp();
    // The 1025 .. 2048-defect classes of the first tier keep event times in 18 bits, which holds when the longest shortest
    // path is below 2^15.  A model that will produce such shots (mean + 6 sigma defects > 1024: d >= 20 codes) and whose
    // paths are a little longer -- unrotated d = 23 at low p: 33 k at 1000 levels -- gets a proportionally coarser weight
    // grid (>= 900 levels: the grid moves 4 predictions per million at 1000 levels) instead of
    // sending those shots to the memory-resident tier, which is two orders of magnitude slower.
    if (!too_far && dec->max_dist >= 32768u && dec->max_dist < 36000u && mu + 6 * sqrt(2 * mu + 1) > 1024 && dec->weight_levels_auto) {
        levels = (int)((double)levels * 32700.0 / dec->max_dist);
        set_weights(levels);
        run_apsp();
    }
    if (too_far) GQ_FAIL(GQ_E_UNSUPPORTED, "matching decoder: path lengths exceed 16 bits; lower weight_levels");
    for (uint32_t s = 0; s < N; ++s) dist[(size_t)s * NT + s] = 0xFFFFu;   // no self pairs (match_tree.cuh relies on it)
    dec->big_classes = dec->max_dist < 32768u;   // event times (<= 8 x a table distance) fit the 18 bits an 11-bit vertex field leaves
    // ---- the NBR_LEN nearest other detectors of every detector (ascending distance, ties by index) for the
    // first tier's dual start; rows D and D + 1 and the tails are padding (see match_tree.cuh)
    const uint32_t pad_entry = (0xFFFFu << 16) | (32u * dem->DW);
    std::vector<uint32_t> nbr((size_t)NT * gqt::NBR_LEN, pad_entry);
    if (32u * dem->DW <= 0xFFFFu) {
        auto lw = [&](uint32_t s_begin, uint32_t s_end) {
            std::vector<uint32_t> tmp;
            for (uint32_t s = s_begin; s < s_end; ++s) {
                tmp.clear();
                for (uint32_t t = 0; t < D; ++t)
                    if (t != s && dist[(size_t)s * NT + t] != 0xFFFFu) tmp.push_back(((uint32_t)dist[(size_t)s * NT + t] << 16) | t);
                size_t keep = std::min<size_t>(gqt::NBR_LEN, tmp.size());
                std::partial_sort(tmp.begin(), tmp.begin() + keep, tmp.end());
                std::copy(tmp.begin(), tmp.begin() + keep, nbr.begin() + (size_t)s * gqt::NBR_LEN);
            }
        };
        unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
        nt = std::min<unsigned>(nt, D);
        std::vector<std::thread> th;
        uint32_t per = (D + nt - 1) / nt;
        for (unsigned t = 0; t < nt; ++t) {
            uint32_t b = std::min<uint32_t>(D, t * per), e = std::min<uint32_t>(D, b + per);
            if (b < e) th.emplace_back(lw, b, e);
        }
        for (auto &t : th) t.join();
    }
    // ---- capacities
    int fast = dec->params.fast_capacity > 0 ? dec->params.fast_capacity : (int)(mu + 6 * sqrt(2 * mu + 1) + 8);
    fast = std::min<int>(even_up(fast + 1), even_up((int)D + 1));
    fast = std::max(fast, 4);
    dec->fast_cap = (uint32_t)fast;
    dec->max_cap = (uint32_t)even_up((int)D + 1);
    // ---- upload
    GQ_CUDA(cudaMalloc((void **)&dec->d_dist, dist.size() * 2));
    GQ_CUDA(cudaMemcpy(dec->d_dist, dist.data(), dist.size() * 2, cudaMemcpyHostToDevice));
    GQ_CUDA(cudaMalloc((void **)&dec->d_nbr, nbr.size() * 4));
    GQ_CUDA(cudaMemcpy(dec->d_nbr, nbr.data(), nbr.size() * 4, cudaMemcpyHostToDevice));
    GQ_CUDA(cudaMalloc((void **)&dec->d_pobs, pobs.size()));
    GQ_CUDA(cudaMemcpy(dec->d_pobs, pobs.data(), pobs.size(), cudaMemcpyHostToDevice));
    if (want_next) {
        GQ_CUDA(cudaMalloc((void **)&dec->d_next, nexttab.size() * 2));
        GQ_CUDA(cudaMemcpy(dec->d_next, nexttab.data(), nexttab.size() * 2, cudaMemcpyHostToDevice));
        GQ_CUDA(cudaMalloc((void **)&dec->d_adj_ptr, adj_ptr.size() * 4));
        GQ_CUDA(cudaMemcpy(dec->d_adj_ptr, adj_ptr.data(), adj_ptr.size() * 4, cudaMemcpyHostToDevice));
        GQ_CUDA(cudaMalloc((void **)&dec->d_adj_node, std::max<size_t>(adj_node.size(), 1) * 4));
        GQ_CUDA(cudaMemcpy(dec->d_adj_node, adj_node.data(), adj_node.size() * 4, cudaMemcpyHostToDevice));
        GQ_CUDA(cudaMalloc((void **)&dec->d_adj_edge, std::max<size_t>(adj_edge.size(), 1) * 4));
        GQ_CUDA(cudaMemcpy(dec->d_adj_edge, adj_edge.data(), adj_edge.size() * 4, cudaMemcpyHostToDevice));
    }
    // ---- the older ladder (shots beyond the first tier's 1024 defects, shots it gives up on, legacy_tiers = 1):
    // dense-matrix register-resident solver with K = 1, 2, 4, 8 slots per lane (< 32K defects); the first pass
    // uses the smallest tier that holds mu + 4 sigma defects, stragglers climb the ladder through overflow
    // lists and anything beyond 255 defects falls to the memory-resident solver
    {
        double want = mu + 2.5 * sqrt(2 * mu + 1);
        if (dec->params.fast_capacity > 0) want = dec->params.fast_capacity;
        int k0 = 1;
        while (k0 < 8 && 32 * k0 - 1 < want) k0 *= 2;
        dec->fast_cap = 32u * k0;
    }
    GQ_CUDA(cudaFuncSetAttribute(match_fast_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    GQ_CUDA(cudaFuncSetAttribute(match_fast_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    GQ_CUDA(cudaFuncSetAttribute(match_fast_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    GQ_CUDA(cudaFuncSetAttribute(match_fast_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    if (configure_tree_class<1>() || configure_tree_class<2>() || configure_tree_class<3>() || configure_tree_class<4>() ||
        configure_tree_class<5>() || configure_tree_class<6>() || configure_tree_class<7>() || configure_tree_class<8>() ||
        configure_tree_class<10>() || configure_tree_class<12>() || configure_tree_class<16>() ||
        configure_tree_class<24>() || configure_tree_class<32>() || configure_tree_class<48>() || configure_tree_class<64>())
        GQ_FAIL(GQ_E_CUDA, "matching decoder: cudaFuncSetAttribute failed for the first-tier kernels");
    GQ_CUDA(cudaMalloc((void **)&dec->d_lists, (size_t)(MT_CLASSES * ((size_t)1 << GQ_SLICE_LOG2) + 64) * 4));
    for (int i = 0; i < MT_CLASSES - 1; ++i) {
        GQ_CUDA(cudaStreamCreateWithFlags(&dec->side_streams[i], cudaStreamNonBlocking));
        GQ_CUDA(cudaEventCreateWithFlags(&dec->side_done[i], cudaEventDisableTiming));
    }
    GQ_CUDA(cudaEventCreate(&dec->kev0));
    GQ_CUDA(cudaEventCreate(&dec->kev1));
    dec->overflow_cap = 1u << GQ_SLICE_LOG2;   // a whole slice may overflow (dense syndromes)
    GQ_CUDA(cudaMalloc((void **)&dec->d_overflow, 2 * (2 + (size_t)dec->overflow_cap) * 4));
    return GQ_OK;
}

#ifdef GQ_STATS
extern "C" void gq_debug_stats(unsigned long long *out, int reset) {
    cudaMemcpyFromSymbol(out, gq_stats, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(gq_stats, z, sizeof(z)); }
}
#endif

void gq_matching_destroy(gq_dec *dec) {
    cudaFree(dec->d_dist); cudaFree(dec->d_pobs); cudaFree(dec->d_next); cudaFree(dec->d_nbr);
    cudaFree(dec->d_adj_ptr); cudaFree(dec->d_adj_node); cudaFree(dec->d_adj_edge);
    cudaFree(dec->d_ws_fast); cudaFree(dec->d_ws_big); cudaFree(dec->d_overflow); cudaFree(dec->d_lists); cudaFree(dec->d_recs);
    for (int i = 0; i < MT_CLASSES - 1; ++i) {
        if (dec->side_streams[i]) cudaStreamDestroy(dec->side_streams[i]);
        if (dec->side_done[i]) cudaEventDestroy(dec->side_done[i]);
    }
    if (dec->kev0) cudaEventDestroy(dec->kev0);
    if (dec->kev1) cudaEventDestroy(dec->kev1);
}

int gq_matching_classify_targets(gq_dec *dec, uint32_t *pred_sm, GqClassifyOut *out) {
    gq_dem *dem = dec->dem;
    GQ_CUDA(cudaMemsetAsync(dec->d_overflow, 0, 8, dem->stream));
    GQ_CUDA(cudaMemsetAsync(dec->d_lists, 0, 64 * 4, dem->stream));
    out->counts = dec->d_lists;
    out->lists = dec->d_lists + 64;
    out->list_stride = 1u << GQ_SLICE_LOG2;
    out->overflow = dec->d_overflow;
    out->overflow_cap = dec->overflow_cap;
    out->pred = pred_sm;
    out->OW = dem->OW;
    return GQ_OK;
}

int gq_matching_decode(gq_dec *dec, const uint32_t *dets_sm, uint64_t nshots, uint32_t *pred_sm,
                       int32_t *weight_out, uint8_t *flags_out, uint32_t *corr_out, int *launches, bool preclassified) {
    gq_dem *dem = dec->dem;
    if (preclassified && (nshots > (1ull << GQ_SLICE_LOG2) || dec->params.legacy_tiers))
        GQ_FAIL(GQ_E_INTERNAL, "matching decoder: a pre-classified batch is one first-tier slice");
    if (corr_out && !dec->d_next) GQ_FAIL(GQ_E_INVALID, "gq_decode: corrections need with_corrections=1 at creation");
    MatchArgs a;
    memset(&a, 0, sizeof(a));
    a.dets = dets_sm; a.pred = pred_sm; a.weight_out = weight_out; a.flags_out = flags_out;
    a.corr_out = corr_out;
    a.D = dem->D; a.DW = dem->DW; a.OW = dem->OW; a.EW = (dec->num_edges + 31) / 32;
    a.N = dem->D + 2;
    a.dist = dec->d_dist; a.pobs = dec->d_pobs; a.next = dec->d_next; a.nbr = dec->d_nbr;
    a.adj_ptr = dec->d_adj_ptr; a.adj_node = dec->d_adj_node; a.adj_edge = dec->d_adj_edge;
    // the overflow lists are bounded: process in slices so they can never overrun
    const uint64_t slice = 1ull << GQ_SLICE_LOG2;
    uint32_t *ovf_buf[2] = {dec->d_overflow, dec->d_overflow + 2 + dec->overflow_cap};
    for (uint64_t s0 = 0; s0 < nshots; s0 += slice) {
        uint64_t ns = std::min(slice, nshots - s0);
        MatchArgs b = a;
        b.dets = dets_sm + s0 * dem->DW;
        b.pred = pred_sm + s0 * dem->OW;
        if (weight_out) b.weight_out = weight_out + s0;
        if (flags_out) b.flags_out = flags_out + s0;
        if (corr_out) b.corr_out = corr_out + s0 * a.EW;
        b.list = nullptr; b.nitems = ns;
        int K = (int)dec->fast_cap / 32;
        int cur = 0;
        if (!dec->params.legacy_tiers) {
            // ---- first tier: per-K kernels over the lists classify_kernel builds (match_tree.cuh)
            uint32_t *d_counts = dec->d_lists;
            uint32_t *d_lists = dec->d_lists + 64;   // [0..16) class counts, [16..48) start / solve cursors
            if (!preclassified) {   // else the producer of the syndromes (sampler / re-stride kernel) has filled the lists
                GQ_CUDA(cudaMemsetAsync(ovf_buf[cur], 0, 8, dem->stream));
                GQ_CUDA(cudaMemsetAsync(d_counts, 0, 64 * 4, dem->stream));
                ClassifyArgs ca;
                ca.dets = b.dets; ca.pred = b.pred; ca.weight_out = b.weight_out; ca.flags_out = b.flags_out;
                ca.counts = d_counts; ca.lists = d_lists; ca.list_stride = (uint32_t)slice;
                ca.overflow = ovf_buf[cur]; ca.overflow_cap = dec->overflow_cap;
                ca.nshots = (uint32_t)ns; ca.D = a.D; ca.DW = a.DW; ca.OW = a.OW;
                classify_kernel<<<(int)std::min<uint64_t>((ns + 255) / 256, (uint64_t)dem->num_sms * 8), 256, 0, dem->stream>>>(ca);
                GQ_CUDA(cudaGetLastError());
                ++*launches;
            }
            uint32_t counts[MT_CLASSES];
            GQ_CUDA(cudaMemcpyAsync(counts, d_counts, sizeof(counts), cudaMemcpyDeviceToHost, dem->stream));
            GQ_CUDA(cudaStreamSynchronize(dem->stream));
            b.cap = MT_CAP;
            b.overflow = ovf_buf[cur]; b.overflow_cap = dec->overflow_cap;
            int order[MT_CLASSES];   // biggest class first
            for (int c = 0; c < MT_CLASSES; ++c) order[c] = c;
            std::sort(order, order + MT_CLASSES, [&](int x, int y) { return counts[x] > counts[y]; });
            int dom = 0;   // the class most shots of this slice fall in: its kernel is the one bench.py's roofline names
            for (int c = 1; c < MT_CLASSES; ++c) if (counts[c] > counts[dom]) dom = c;
            // the biggest class runs on the decoder's stream, the others on side streams: they start as soon
            // as the big kernel's persistent blocks begin to retire, so the tails of the launches overlap
            // (the host synchronised after classify_kernel, so nothing earlier is still in flight)
            // hand-off records of the slice, one region per class
            size_t rec_off[MT_CLASSES], rec_total = 0;
            for (int c = 0; c < MT_CLASSES; ++c) { rec_off[c] = rec_total; rec_total += ((size_t)counts[c] * start_record_bytes(c) + 255) & ~(size_t)255; }
            if (rec_total > dec->recs_bytes) {
                cudaFree(dec->d_recs);
                dec->d_recs = nullptr; dec->recs_bytes = 0;
                GQ_CUDA(cudaMalloc(&dec->d_recs, rec_total + rec_total / 8));
                dec->recs_bytes = rec_total + rec_total / 8;
            }
            int nside = 0;
            bool timed = false;   // the dominant class's kernels ran (its events were recorded)
            for (int oi = 0; oi < MT_CLASSES; ++oi) {
                const int c = order[oi];
                if (counts[c] == 0) continue;
                if (c >= 13 && !dec->big_classes) {
                    append_list_to_overflow_kernel<<<1, 1024, 0, dem->stream>>>(d_lists + (size_t)c * slice, counts[c], ovf_buf[cur], dec->overflow_cap, c == 13 ? 1536u : 2048u);
                    GQ_CUDA(cudaGetLastError());
                    ++*launches;
                    continue;
                }
                cudaStream_t st = (oi == 0 || !dec->side_streams[0] || !GQ_SIDE_STREAMS) ? dem->stream : dec->side_streams[nside++];
                b.list = d_lists + (size_t)c * slice; b.nitems = counts[c];
                b.ws = (char *)dec->d_recs + rec_off[c]; b.ws_stride = start_record_bytes(c);
                int lrc;
                switch (c) {
                case 0: lrc = launch_tree_class<1>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                case 1: lrc = launch_tree_class<2>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                case 2: lrc = launch_tree_class<3>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                case 3: lrc = launch_tree_class<4>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                case 4: lrc = launch_tree_class<5>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                case 5: lrc = launch_tree_class<6>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                case 6: lrc = launch_tree_class<7>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                case 7: lrc = launch_tree_class<8>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                case 8: lrc = launch_tree_class<10>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                case 9: lrc = launch_tree_class<12>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                case 10: lrc = launch_tree_class<16>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                case 11: lrc = launch_tree_class<24>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                case 12: lrc = launch_tree_class<32>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                case 13: lrc = launch_tree_class<48>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                default: lrc = launch_tree_class<64>(b, d_counts + 16 + c, dem, st, launches, c == dom ? dec->kev0 : nullptr, c == dom ? dec->kev1 : nullptr); break;
                }
                if (lrc) GQ_FAIL(lrc, "matching decoder: first-tier kernel launch failed");
                if (c == dom) timed = true;
                if (st != dem->stream) {
                    GQ_CUDA(cudaEventRecord(dec->side_done[nside - 1], st));
                    GQ_CUDA(cudaStreamWaitEvent(dem->stream, dec->side_done[nside - 1], 0));
                }
            }
            uint32_t ovf[2];
            GQ_CUDA(cudaMemcpyAsync(ovf, ovf_buf[cur], 8, cudaMemcpyDeviceToHost, dem->stream));
            GQ_CUDA(cudaStreamSynchronize(dem->stream));
            if (counts[dom] && timed) {
                float ms = 0;
                cudaEventElapsedTime(&ms, dec->kev0, dec->kev1);
                dec->kprof[0] += ms; dec->kprof[1] += 1; dec->kprof[2] += counts[dom]; dec->kprof[3] = dom < 8 ? dom + 1 : (dom == 8 ? 10 : (dom == 9 ? 12 : (dom == 10 ? 16 : (dom == 11 ? 24 : (dom == 12 ? 32 : (dom == 13 ? 48 : 64))))));
            }
            if (ovf[0] == 0) continue;
            if (ovf[0] > dec->overflow_cap) GQ_FAIL(GQ_E_INTERNAL, "matching decoder: overflow list overrun");
            b.list = ovf_buf[cur] + 2; b.nitems = ovf[0];
            cur ^= 1;
            K = 1;
            while (K < 8 && 32 * K - 1 < (int)ovf[1]) K *= 2;
            if (32 * K - 1 < (int)ovf[1]) K = 16;   // beyond the register tiers: straight to the memory tier
        }
        for (;;) {
            if (K > 8) {
                uint32_t ovf[2];
                GQ_CUDA(cudaMemcpyAsync(ovf, b.list - 2, 8, cudaMemcpyDeviceToHost, dem->stream));
                GQ_CUDA(cudaStreamSynchronize(dem->stream));
                int bigcap = std::min<int>(even_up((int)ovf[1] + 1), (int)dec->max_cap);
                size_t wsb = warp_ws_bytes(bigcap);
                int nw = (int)std::min<uint64_t>(ovf[0], (uint64_t)dem->num_sms * 8);
                size_t budget = (size_t)8 << 30;
                nw = (int)std::max<size_t>(1, std::min<size_t>(nw, budget / wsb));
                nw = std::max(8, (nw / 8) * 8);
                if (wsb * nw > dec->ws_big_bytes) {
                    cudaFree(dec->d_ws_big);
                    dec->d_ws_big = nullptr; dec->ws_big_bytes = 0;
                    GQ_CUDA(cudaMalloc(&dec->d_ws_big, wsb * nw));
                    dec->ws_big_bytes = wsb * nw;
                }
                MatchArgs c = b;
                c.cap = (uint32_t)bigcap;
                c.ws = (char *)dec->d_ws_big; c.ws_stride = wsb;
                c.overflow = nullptr; c.overflow_cap = 0;
                match_kernel<<<nw / 8, 256, 0, dem->stream>>>(c);
                GQ_CUDA(cudaGetLastError());
                ++*launches;
                break;
            }
            // ---- one pass of tier K over the current item list
            const int cap = 32 * K;
            const int per_sm = K <= 4 ? GQ_PER_SM : 2;
            const int nwarps = dem->num_sms * per_sm * 8;
            const size_t wbytes = ((size_t)cap * (cap + 1) * 2 + 255) & ~(size_t)255;
            if (wbytes * nwarps > dec->ws_fast_bytes) {
                cudaFree(dec->d_ws_fast);
                dec->d_ws_fast = nullptr; dec->ws_fast_bytes = 0;
                GQ_CUDA(cudaMalloc(&dec->d_ws_fast, wbytes * nwarps));
                dec->ws_fast_bytes = wbytes * nwarps;
            }
            GQ_CUDA(cudaMemsetAsync(ovf_buf[cur], 0, 8, dem->stream));
            b.cap = (uint32_t)cap;
            b.ws = (char *)dec->d_ws_fast; b.ws_stride = wbytes;
            b.overflow = ovf_buf[cur]; b.overflow_cap = dec->overflow_cap;
            int blocks = (int)std::min<uint64_t>((b.nitems + 7) / 8, (uint64_t)nwarps / 8);
            size_t smem = 8 * (gqr::forest_bytes(cap) + (size_t)cap * 4);
            if (K == 1) match_fast_kernel<1><<<blocks, 256, smem, dem->stream>>>(b);
            else if (K == 2) match_fast_kernel<2><<<blocks, 256, smem, dem->stream>>>(b);
            else if (K == 4) match_fast_kernel<4><<<blocks, 256, smem, dem->stream>>>(b);
            else match_fast_kernel<8><<<blocks, 256, smem, dem->stream>>>(b);
            GQ_CUDA(cudaGetLastError());
            ++*launches;
            uint32_t ovf[2];
            GQ_CUDA(cudaMemcpyAsync(ovf, ovf_buf[cur], 8, cudaMemcpyDeviceToHost, dem->stream));
            GQ_CUDA(cudaStreamSynchronize(dem->stream));
            if (ovf[0] == 0) break;
            if (ovf[0] > dec->overflow_cap) GQ_FAIL(GQ_E_INTERNAL, "matching decoder: overflow list overrun");
            b.list = ovf_buf[cur] + 2; b.nitems = ovf[0];
            cur ^= 1;
            int nextK = K;
            while (nextK < 8 && 32 * nextK - 1 < (int)ovf[1]) nextK *= 2;
            if (32 * nextK - 1 >= (int)ovf[1] && nextK != K) { K = nextK; continue; }
            // ---- beyond the largest register tier: memory-resident solver, workspace sized by the worst shot
            int bigcap = std::min<int>(even_up((int)ovf[1] + 1), (int)dec->max_cap);
            size_t wsb = warp_ws_bytes(bigcap);
            int nw = (int)std::min<uint64_t>(ovf[0], (uint64_t)dem->num_sms * 8);
            size_t budget = (size_t)8 << 30;
            nw = (int)std::max<size_t>(1, std::min<size_t>(nw, budget / wsb));
            nw = std::max(8, (nw / 8) * 8);
            if (wsb * nw > dec->ws_big_bytes) {
                cudaFree(dec->d_ws_big);
                dec->d_ws_big = nullptr; dec->ws_big_bytes = 0;
                GQ_CUDA(cudaMalloc(&dec->d_ws_big, wsb * nw));
                dec->ws_big_bytes = wsb * nw;
            }
            MatchArgs c = b;
            c.cap = (uint32_t)bigcap;
            c.ws = (char *)dec->d_ws_big; c.ws_stride = wsb;
            c.overflow = nullptr; c.overflow_cap = 0;
            match_kernel<<<nw / 8, 256, 0, dem->stream>>>(c);
            GQ_CUDA(cudaGetLastError());
            ++*launches;
            break;
        }
    }
    return GQ_OK;
}
