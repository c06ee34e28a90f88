// DEM handle + kernels K1 (Philox sampler), K2 (GF(2) extraction), bit transposes, K4 (count).
//
// K1/K2 fused (sample_kernel): a block owns a tile of T shots whose syndrome lives in shared memory
// as [T][DW+OW] words.  The DEM's mechanisms are grouped by probability; each group's flattened
// (mechanism, shot-in-tile) index space is cut into segments, and a thread walks its segment with
// geometric gaps floor(log(1-v)/log(1-p)), so the work is proportional to the errors that fire
// (sum p ~ 37 per shot at the north-star) instead of to mechanisms x shots (24 483 per shot).
// Every hit XORs the mechanism's detector/observable bits of that shot into the tile; the tile is
// then written to HBM with coalesced stores.  Replaces Stim's frame sampler.
//
// extract_dm_kernel is the bit-sliced form of K2 (32 shots per word): one warp-wide coalesced load
// per (mechanism row, 32-word column group), XOR-gathered along the detector's CSR row, used as the
// bit-exact parity hook on explicit error vectors.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <numeric>

#include "gq_common.h"

static thread_local std::string g_last_error;
void gq_set_error(const std::string &msg) { g_last_error = msg; }

extern "C" const char *gq_last_error(void) { return g_last_error.c_str(); }
extern "C" int gq_version(void) { return 100; }
extern "C" int gq_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

struct SampleArgs {
    const GqClass *classes;
    const uint32_t *seg_first;
    const uint32_t *mech_ptr, *mech_det, *mech_obs, *mech_orig;
    const uint4 *seg_rec;    // [nsegments] {1 / log(1 - p) as float bits, first mechanism of the class, first and
                             // one-past-last flat index of the segment}: one load instead of a binary search
    const uint4 *mech_rec;   // nullable: one 16-byte record per mechanism {det0 | det1 << 16, det2 | det3 << 16, observable
                             // mask, detector count} when every mechanism has <= 4 detectors below 65536 (one load per hit)
    uint32_t nclasses, nsegments;
    uint32_t D, DW, OW, T, Tlog2;
    uint64_t seed, tile0, nshots;  // tile0 = global index of the first tile of this call
    uint32_t ntiles;
    uint32_t *dets, *obs, *errs;
    uint32_t err_words;  // ceil(nshots / 32)
    GqClassifyOut cls;   // counts != null: sort the shots into the matching decoder's per-class lists while flushing
};

__global__ void __launch_bounds__(256) sample_kernel(SampleArgs a) {
    extern __shared__ uint32_t tile[];
    const uint32_t stride = a.DW + a.OW;
    const uint32_t tile_words = a.T * stride;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    for (uint32_t tl = blockIdx.x; tl < a.ntiles; tl += gridDim.x) {
        for (uint32_t i = threadIdx.x; i < tile_words; i += blockDim.x) tile[i] = 0;
        __syncthreads();
        const uint64_t gtile = a.tile0 + tl;
        const uint64_t shot_base = (uint64_t)tl * a.T;  // relative to this call's first shot
        for (uint32_t seg = threadIdx.x; seg < a.nsegments; seg += blockDim.x) {
            const uint4 sr = a.seg_rec[seg];
            struct { float inv_log1m; uint32_t mech_start; } c = {__uint_as_float(sr.x), sr.y};
            uint32_t pos = sr.z;
            const uint32_t end = sr.w;
            uint32_t draw = 0;
            bool done = false;
            while (!done) {
                uint4 r = philox4x32_10(make_uint4((uint32_t)gtile, (uint32_t)(gtile >> 32), seg, draw++), key);
                uint32_t rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (done) break;
                    // v uniform in (0, 1): gap = floor(log(1 - v) / log(1 - p)).  (float)rv rounds to 24 bits, so the top 128
                    // values of rv would give v == 1 and an infinite gap: clamped to the largest float below 1 (gap 16.6 / p,
                    // beyond every segment).  For v >= 1/2 the rounding moves 1 - v by up to 3e-8: a far hit (gap ~ 9 / p) can
                    // land a few positions off, symmetrically -- no bias on rates (tests: per-class z-scores, detector marginals;
                    // LER against the oracle's double-precision sampler at 1.3e8 shots, profiles/r02_ler_samplers.log).
                    float v = fminf(((float)rv[k] + 0.5f) * 2.3283064365386963e-10f, 0.99999994f);
                    float g = floorf(log1pf(-v) * c.inv_log1m);
                    float room = (float)(end - pos);
                    if (!(g < room)) { done = true; break; }
                    pos += (uint32_t)g;
                    if (pos >= end) { done = true; break; }  // float rounding of very long segments
                    const uint32_t m = c.mech_start + (pos >> a.Tlog2);
                    const uint32_t s = pos & (a.T - 1);
                    uint32_t *row = tile + s * stride;
                    if (a.mech_rec) {
                        const uint4 rec = a.mech_rec[m];
                        const uint32_t d0 = rec.x & 0xffffu, d1 = rec.x >> 16, d2 = rec.y & 0xffffu, d3 = rec.y >> 16;
                        if (rec.w > 0) atomicXor(&row[d0 >> 5], 1u << (d0 & 31));
                        if (rec.w > 1) atomicXor(&row[d1 >> 5], 1u << (d1 & 31));
                        if (rec.w > 2) atomicXor(&row[d2 >> 5], 1u << (d2 & 31));
                        if (rec.w > 3) atomicXor(&row[d3 >> 5], 1u << (d3 & 31));
                        if (rec.z) atomicXor(&row[a.DW], rec.z);
                    } else {
                        const uint32_t b = a.mech_ptr[m], e = a.mech_ptr[m + 1];
                        for (uint32_t q = b; q < e; ++q) {
                            uint32_t d = a.mech_det[q];
                            atomicXor(&row[d >> 5], 1u << (d & 31));
                        }
                        const uint32_t om = a.mech_obs[m];
                        if (om) atomicXor(&row[a.DW], om);
                    }
                    if (a.errs) {
                        uint64_t gs = shot_base + s;
                        if (gs < a.nshots)
                            atomicOr(&a.errs[(size_t)a.mech_orig[m] * a.err_words + (gs >> 5)], 1u << (gs & 31));
                    }
                    ++pos;
                    if (pos >= end) done = true;
                }
            }
        }
        __syncthreads();
        // flush: rows are contiguous in HBM, so consecutive threads write consecutive words
        const uint64_t rows_left = a.nshots - shot_base;
        const uint32_t rows = rows_left < a.T ? (uint32_t)rows_left : a.T;
        uint32_t *gd = a.dets + shot_base * a.DW;
        if (a.cls.counts) {
            // fused classification (what classify_kernel would do in a pass of its own): the defect count of every
            // row is a popcount of words that are in shared memory anyway; the tile's shots are appended to the
            // per-class lists with one global atomic per (tile, class)
            uint32_t *cl_cnt = tile + tile_words;            // [16] shots of the tile per class (15 = overflow)
            uint32_t *cl_base = cl_cnt + 16;                 // [16] their first position in the global lists
            uint16_t *cl_pos = reinterpret_cast<uint16_t *>(cl_base + 16);   // [T] class << 12 | position within (tile, class)
            if (threadIdx.x < 16) cl_cnt[threadIdx.x] = 0;
            __syncthreads();
            for (uint32_t s0 = threadIdx.x / 32; s0 < rows; s0 += blockDim.x / 32) {
                const uint32_t *row = tile + s0 * stride;
                uint32_t n = 0;
                for (uint32_t w = threadIdx.x & 31; w < a.DW; w += 32) { const uint32_t v = row[w]; gd[(size_t)s0 * a.DW + w] = v; n += __popc(v); }
                n = __reduce_add_sync(0xffffffffu, n);
                if ((threadIdx.x & 31) == 0) {
                    if (n == 0) {
                        cl_pos[s0] = 0xffffu;
                        for (uint32_t w = 0; w < a.cls.OW; ++w) a.cls.pred[(shot_base + s0) * a.cls.OW + w] = 0u;
                    } else {
                        const uint32_t c = (uint32_t)gq_mt_class_of((int)n);   // GQ_MT_CLASSES (13) = beyond the first tier
                        cl_pos[s0] = (uint16_t)((c << 12) | atomicAdd(&cl_cnt[c], 1u));
                        if (c == (uint32_t)GQ_MT_CLASSES) atomicMax(&a.cls.overflow[1], n);
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x <= (uint32_t)GQ_MT_CLASSES && cl_cnt[threadIdx.x])
                cl_base[threadIdx.x] = atomicAdd(threadIdx.x < (uint32_t)GQ_MT_CLASSES ? &a.cls.counts[threadIdx.x] : &a.cls.overflow[0], cl_cnt[threadIdx.x]);
            __syncthreads();
            for (uint32_t s0 = threadIdx.x; s0 < rows; s0 += blockDim.x) {
                const uint32_t cp = cl_pos[s0];
                if (cp == 0xffffu) continue;
                const uint32_t c = cp >> 12, at = cl_base[c] + (cp & 0xfffu);
                if (c < (uint32_t)GQ_MT_CLASSES) a.cls.lists[(size_t)c * a.cls.list_stride + at] = (uint32_t)(shot_base + s0);
                else if (at < a.cls.overflow_cap) a.cls.overflow[2 + at] = (uint32_t)(shot_base + s0);
            }
        } else {
            for (uint32_t s0 = threadIdx.x / 32; s0 < rows; s0 += blockDim.x / 32) {
                const uint32_t *row = tile + s0 * stride;
                for (uint32_t w = threadIdx.x & 31; w < a.DW; w += 32) gd[(size_t)s0 * a.DW + w] = row[w];
            }
        }
        if (a.obs) {
            uint32_t *go = a.obs + shot_base * a.OW;
            for (uint32_t i = threadIdx.x; i < rows * a.OW; i += blockDim.x) {
                uint32_t s = i / a.OW, w = i - s * a.OW;
                go[i] = tile[s * stride + a.DW + w];
            }
        }
        __syncthreads();
    }
}

// one thread per (row, word-column); a warp covers 32 consecutive columns of one row
__global__ void extract_dm_kernel(const uint32_t *row_ptr, const uint32_t *row_mech, uint32_t nrows,
                                  uint32_t D, const uint32_t *errs, uint32_t SW, uint32_t *dets,
                                  uint32_t *obs) {
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = blockIdx.y;
    if (r >= nrows || col >= SW) return;
    uint32_t acc = 0;
    const uint32_t b = row_ptr[r], e = row_ptr[r + 1];
    for (uint32_t q = b; q < e; ++q) acc ^= errs[(size_t)row_mech[q] * SW + col];
    if (r < D) dets[(size_t)r * SW + col] = acc;
    else obs[(size_t)(r - D) * SW + col] = acc;
}

// 32x32 bit transposes, one warp per tile
__global__ void transpose_dm_to_sm_kernel(const uint32_t *dm, uint32_t nbits, uint64_t nshots,
                                          uint32_t SW, uint32_t NW, uint32_t *sm) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t ntile = (uint64_t)SW * NW;
    if (warp >= ntile) return;
    const uint32_t wd = (uint32_t)(warp % NW);
    const uint32_t sw = (uint32_t)(warp / NW);
    const uint32_t bit = wd * 32 + lane;
    uint32_t word = bit < nbits ? dm[(size_t)bit * SW + sw] : 0;
    uint32_t outw = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        uint32_t bal = __ballot_sync(0xffffffffu, (word >> i) & 1);
        if (lane == i) outw = bal;
    }
    const uint64_t shot = (uint64_t)sw * 32 + lane;
    if (shot < nshots) sm[shot * NW + wd] = outw;
}

__global__ void transpose_sm_to_dm_kernel(const uint32_t *sm, uint32_t nbits, uint64_t nshots,
                                          uint32_t SW, uint32_t NW, uint32_t *dm) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t ntile = (uint64_t)SW * NW;
    if (warp >= ntile) return;
    const uint32_t wd = (uint32_t)(warp % NW);
    const uint32_t sw = (uint32_t)(warp / NW);
    const uint64_t shot = (uint64_t)sw * 32 + lane;
    uint32_t word = shot < nshots ? sm[shot * NW + wd] : 0;
    uint32_t outw = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        uint32_t bal = __ballot_sync(0xffffffffu, (word >> i) & 1);
        if (lane == i) outw = bal;
    }
    const uint32_t bit = wd * 32 + lane;
    if (bit < nbits) dm[(size_t)bit * SW + sw] = outw;
}

// discard_word != 0: a shot whose first prediction word equals it was given up by the decoder (BP+OSD sentinel) -- it is
// counted in count[1] (sinter's "discards") and not as an error
__global__ void count_errors_kernel(const uint32_t *pred, const uint32_t *obs, uint64_t nshots,
                                    uint32_t OW, uint32_t discard_word, unsigned long long *count) {
    uint32_t local = 0, disc = 0;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nshots;
         s += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t diff = 0;
        for (uint32_t w = 0; w < OW; ++w) diff |= pred[s * OW + w] ^ obs[s * OW + w];
        const bool discarded = discard_word != 0u && pred[s * OW] == discard_word;
        local += diff != 0 && !discarded;
        disc += discarded;
    }
    local = __reduce_add_sync(0xffffffffu, local);
    disc = __reduce_add_sync(0xffffffffu, disc);
    __shared__ uint32_t part[32], dpart[32];
    if ((threadIdx.x & 31) == 0) { part[threadIdx.x >> 5] = local; dpart[threadIdx.x >> 5] = disc; }
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0;
        uint32_t d = threadIdx.x < (blockDim.x >> 5) ? dpart[threadIdx.x] : 0;
        v = __reduce_add_sync(0xffffffffu, v);
        d = __reduce_add_sync(0xffffffffu, d);
        if (threadIdx.x == 0 && v) atomicAdd(count, (unsigned long long)v);
        if (threadIdx.x == 0 && d) atomicAdd(count + 1, (unsigned long long)d);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <typename T>
static int upload(T **dst, const std::vector<T> &src) {
    size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    GQ_CUDA(cudaMalloc((void **)dst, bytes));
    if (!src.empty()) GQ_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return GQ_OK;
}

extern "C" int gq_dem_create(uint32_t D, uint32_t O, uint32_t M, const double *p,
                             const uint32_t *det_indptr, const uint32_t *det_idx,
                             const uint32_t *obs_indptr, const uint32_t *obs_idx,
                             const uint32_t *comp_indptr, const uint32_t *comp_a,
                             const uint32_t *comp_b, const uint64_t *comp_obs, int device,
                             gq_dem **out) {
    if (!out) GQ_FAIL(GQ_E_INVALID, "gq_dem_create: out is NULL");
    *out = nullptr;
    if (M && (!p || !det_indptr || !obs_indptr)) GQ_FAIL(GQ_E_INVALID, "gq_dem_create: NULL array");
    if (O > 32) GQ_FAIL(GQ_E_UNSUPPORTED, "gq_dem_create: more than 32 observables are not supported");
    if (M >= (1u << 24)) GQ_FAIL(GQ_E_UNSUPPORTED, "gq_dem_create: too many error mechanisms");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        GQ_FAIL(GQ_E_CUDA, "gq_dem_create: no CUDA device available (the hot path has no CPU fallback)");
    if (device < 0 || device >= ndev) GQ_FAIL(GQ_E_INVALID, "gq_dem_create: bad device index");
    GQ_CUDA(cudaSetDevice(device));
    gq_dem *dem = new gq_dem();
    dem->device = device;
    dem->D = D; dem->O = O; dem->M = M;
    dem->DW = (D + 31) / 32; dem->OW = std::max(1u, (O + 31) / 32);
    if (dem->DW == 0) dem->DW = 1;
    cudaDeviceProp prop;
    GQ_CUDA(cudaGetDeviceProperties(&prop, device));
    dem->num_sms = prop.multiProcessorCount;
    GQ_CUDA(cudaStreamCreateWithFlags(&dem->stream, cudaStreamNonBlocking));
    dem->p.assign(p, p + M);
    dem->det_indptr.assign(det_indptr, det_indptr + M + 1);
    dem->obs_indptr.assign(obs_indptr, obs_indptr + M + 1);
    if (M == 0) { dem->det_indptr.assign(1, 0); dem->obs_indptr.assign(1, 0); }
    dem->det_idx.assign(det_idx, det_idx + dem->det_indptr[M]);
    dem->obs_idx.assign(obs_idx, obs_idx + dem->obs_indptr[M]);
    for (uint32_t d : dem->det_idx) if (d >= D) { gq_dem_destroy(dem); GQ_FAIL(GQ_E_INVALID, "gq_dem_create: detector index out of range"); }
    for (uint32_t o : dem->obs_idx) if (o >= O) { gq_dem_destroy(dem); GQ_FAIL(GQ_E_INVALID, "gq_dem_create: observable index out of range"); }
    for (uint32_t i = 0; i < M; ++i)
        if (!(p[i] >= 0.0 && p[i] <= 1.0)) { gq_dem_destroy(dem); GQ_FAIL(GQ_E_INVALID, "gq_dem_create: probability out of [0,1]"); }
    if (comp_indptr) {
        dem->has_components = true;
        dem->comp_indptr.assign(comp_indptr, comp_indptr + M + 1);
        uint32_t nc = dem->comp_indptr[M];
        dem->comp_a.assign(comp_a, comp_a + nc);
        dem->comp_b.assign(comp_b, comp_b + nc);
        dem->comp_obs.assign(comp_obs, comp_obs + nc);
    }
    // ---- tile size: largest power of two <= 256 whose tile fits comfortably in shared memory
    uint32_t stride = dem->DW + dem->OW;
    uint32_t T = 256;
    while (T > 32 && (size_t)T * stride * 4 > 160 * 1024) T >>= 1;
    if ((size_t)T * stride * 4 > 220 * 1024) { gq_dem_destroy(dem); GQ_FAIL(GQ_E_UNSUPPORTED, "gq_dem_create: too many detectors for one sampling tile"); }
    dem->tile_shots = T;
    dem->tile_log2 = 0;
    while ((1u << dem->tile_log2) < T) ++dem->tile_log2;
    // ---- probability classes (p rounded to 12 significant digits), mechanisms sorted by class
    std::vector<uint32_t> order;
    std::vector<double> key(M);
    for (uint32_t i = 0; i < M; ++i) {
        double v = p[i];
        if (v > 0) {
            double mag = pow(10.0, 11 - floor(log10(v)));
            v = round(v * mag) / mag;
        }
        key[i] = v;
        if (v > 0) order.push_back(i);
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    std::vector<GqClass> classes;
    std::vector<uint32_t> seg_first;
    std::vector<uint32_t> mech_ptr(1, 0), mech_det, mech_obs, mech_orig;
#ifndef GQ_SEG_LAMBDA
#define GQ_SEG_LAMBDA 16.0   // expected hits per RNG segment (a segment costs one wasted draw: measured 4 -> 1.96, 8 -> 1.72, 16 -> 1.65 ms per 2^22 north-star shots)
#endif
    const double lambda = GQ_SEG_LAMBDA;
    uint32_t nseg = 0;
    dem->sum_p = 0;
    for (size_t i = 0; i < order.size();) {
        size_t j = i;
        while (j < order.size() && key[order[j]] == key[order[i]]) ++j;
        double pc = key[order[i]];
        GqClass c;
        c.inv_log1m = pc >= 1.0 ? -0.0f : (float)(1.0 / log1p(-pc));
        c.mech_start = (uint32_t)i;
        c.mech_count = (uint32_t)(j - i);
        c.flat_size = c.mech_count * T;
        double len = ceil(lambda / pc);
        c.seg_len = (uint32_t)std::min<double>(std::max<double>(len, 1.0), (double)c.flat_size);
        c.seg_first = nseg;
        seg_first.push_back(nseg);
        nseg += (c.flat_size + c.seg_len - 1) / c.seg_len;
        classes.push_back(c);
        dem->sum_p += pc * c.mech_count;
        i = j;
    }
    seg_first.push_back(nseg);
    for (uint32_t m : order) {
        for (uint32_t q = dem->det_indptr[m]; q < dem->det_indptr[m + 1]; ++q) mech_det.push_back(dem->det_idx[q]);
        mech_ptr.push_back((uint32_t)mech_det.size());
        uint32_t om = 0;
        for (uint32_t q = dem->obs_indptr[m]; q < dem->obs_indptr[m + 1]; ++q) om ^= 1u << dem->obs_idx[q];
        mech_obs.push_back(om);
        mech_orig.push_back(m);
    }
    dem->nclasses = (uint32_t)classes.size();
    dem->nsegments = nseg;
    // one record per RNG segment
    std::vector<uint32_t> seg_rec((size_t)4 * nseg);
    for (const GqClass &c : classes) {
        const uint32_t ns = (c.flat_size + c.seg_len - 1) / c.seg_len;
        for (uint32_t k = 0; k < ns; ++k) {
            uint32_t bits;
            memcpy(&bits, &c.inv_log1m, 4);
            const uint32_t pos = k * c.seg_len;
            uint32_t *r = &seg_rec[(size_t)4 * (c.seg_first + k)];
            r[0] = bits; r[1] = c.mech_start; r[2] = pos; r[3] = std::min(pos + c.seg_len, c.flat_size);
        }
    }
    // one-load records for the sampler when every mechanism is small enough
    std::vector<uint32_t> mech_rec;
    {
        bool ok = D < 65536u;
        for (size_t i = 0; ok && i + 1 < mech_ptr.size(); ++i) ok = mech_ptr[i + 1] - mech_ptr[i] <= 4;
        if (ok) {
            mech_rec.assign(4 * (mech_ptr.size() - 1), 0u);
            for (size_t i = 0; i + 1 < mech_ptr.size(); ++i) {
                const uint32_t b = mech_ptr[i], n = mech_ptr[i + 1] - b;
                uint32_t d[4] = {0, 0, 0, 0};
                for (uint32_t q = 0; q < n; ++q) d[q] = mech_det[b + q];
                mech_rec[4 * i + 0] = d[0] | (d[1] << 16);
                mech_rec[4 * i + 1] = d[2] | (d[3] << 16);
                mech_rec[4 * i + 2] = mech_obs[i];
                mech_rec[4 * i + 3] = n;
            }
        }
    }
    // ---- transposed CSR for the bit-sliced extractor
    std::vector<uint32_t> row_ptr(D + O + 1, 0), row_mech(dem->det_idx.size() + dem->obs_idx.size());
    for (uint32_t d : dem->det_idx) row_ptr[d + 1]++;
    for (uint32_t o : dem->obs_idx) row_ptr[D + o + 1]++;
    for (uint32_t r = 0; r < D + O; ++r) row_ptr[r + 1] += row_ptr[r];
    {
        std::vector<uint32_t> fill(row_ptr.begin(), row_ptr.end() - 1);
        for (uint32_t m = 0; m < M; ++m) {
            for (uint32_t q = dem->det_indptr[m]; q < dem->det_indptr[m + 1]; ++q) row_mech[fill[dem->det_idx[q]]++] = m;
            for (uint32_t q = dem->obs_indptr[m]; q < dem->obs_indptr[m + 1]; ++q) row_mech[fill[D + dem->obs_idx[q]]++] = m;
        }
    }
    int rc;
    if ((rc = upload(&dem->d_classes, classes)) || (rc = upload(&dem->d_seg_first, seg_first)) ||
        (rc = upload(&dem->d_mech_ptr, mech_ptr)) || (rc = upload(&dem->d_mech_det, mech_det)) ||
        (rc = upload(&dem->d_mech_obs, mech_obs)) || (rc = upload(&dem->d_mech_orig, mech_orig)) ||
        (!mech_rec.empty() && (rc = upload(&dem->d_mech_rec, mech_rec))) || (rc = upload(&dem->d_seg_rec, seg_rec)) ||
        (rc = upload(&dem->d_row_ptr, row_ptr)) || (rc = upload(&dem->d_row_mech, row_mech))) {
        gq_dem_destroy(dem);
        return rc;
    }
    cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    *out = dem;
    return GQ_OK;
}

extern "C" void gq_dem_destroy(gq_dem *dem) {
    if (!dem) return;
    cudaSetDevice(dem->device);
    cudaFree(dem->d_classes); cudaFree(dem->d_seg_first); cudaFree(dem->d_mech_ptr);
    cudaFree(dem->d_mech_det); cudaFree(dem->d_mech_obs); cudaFree(dem->d_mech_orig); cudaFree(dem->d_mech_rec); cudaFree(dem->d_seg_rec);
    cudaFree(dem->d_row_ptr); cudaFree(dem->d_row_mech);
    if (dem->stream) cudaStreamDestroy(dem->stream);
    delete dem;
}

extern "C" int gq_dem_get_info(const gq_dem *dem, gq_dem_info *info) {
    if (!dem || !info) GQ_FAIL(GQ_E_INVALID, "gq_dem_get_info: NULL argument");
    info->num_detectors = dem->D; info->num_observables = dem->O; info->num_errors = dem->M;
    info->det_words = dem->DW; info->obs_words = dem->OW;
    info->tile_shots = dem->tile_shots;
    info->num_classes = dem->nclasses; info->num_segments = dem->nsegments;
    info->sum_p = dem->sum_p;
    return GQ_OK;
}

extern "C" int gq_sample(gq_dem *dem, uint64_t seed, uint64_t shot0, uint64_t nshots,
                         uint32_t *dets_sm, uint32_t *obs_sm, uint32_t *errs_mm) {
    return gq_sample_classified(dem, seed, shot0, nshots, dets_sm, obs_sm, errs_mm, nullptr);
}

int gq_sample_classified(gq_dem *dem, uint64_t seed, uint64_t shot0, uint64_t nshots, uint32_t *dets_sm,
                         uint32_t *obs_sm, uint32_t *errs_mm, const GqClassifyOut *cls) {
    if (!dem || !dets_sm) GQ_FAIL(GQ_E_INVALID, "gq_sample: NULL argument");
    if (shot0 % dem->tile_shots) GQ_FAIL(GQ_E_INVALID, "gq_sample: shot0 must be a multiple of tile_shots");
    if (nshots == 0) return GQ_OK;
    GQ_CUDA(cudaSetDevice(dem->device));
    SampleArgs a;
    a.classes = dem->d_classes; a.seg_first = dem->d_seg_first;
    a.mech_ptr = dem->d_mech_ptr; a.mech_det = dem->d_mech_det; a.mech_obs = dem->d_mech_obs;
    a.mech_orig = dem->d_mech_orig;
    a.mech_rec = reinterpret_cast<const uint4 *>(dem->d_mech_rec);
    a.seg_rec = reinterpret_cast<const uint4 *>(dem->d_seg_rec);
    a.nclasses = dem->nclasses; a.nsegments = dem->nsegments;
    a.D = dem->D; a.DW = dem->DW; a.OW = dem->OW; a.T = dem->tile_shots; a.Tlog2 = dem->tile_log2;
    a.seed = seed; a.tile0 = shot0 / dem->tile_shots; a.nshots = nshots;
    uint64_t ntiles = (nshots + dem->tile_shots - 1) / dem->tile_shots;
    if (ntiles > 0xFFFFFFFFull) GQ_FAIL(GQ_E_INVALID, "gq_sample: too many shots in one call");
    a.ntiles = (uint32_t)ntiles;
    a.dets = dets_sm; a.obs = obs_sm; a.errs = errs_mm;
    a.err_words = (uint32_t)((nshots + 31) / 32);
    if (cls) a.cls = *cls;
    if (a.cls.counts && (dem->tile_shots > 4096 || nshots > 0xFFFFFFFFull)) GQ_FAIL(GQ_E_INTERNAL, "gq_sample: classification limits");
    size_t smem = (size_t)dem->tile_shots * (dem->DW + dem->OW) * 4 + 128 + (size_t)dem->tile_shots * 2;
    int per_sm = std::max<int>(1, (int)(200 * 1024 / std::max<size_t>(smem, 1)));
    per_sm = std::min(per_sm, 8);
    uint32_t grid = (uint32_t)std::min<uint64_t>(ntiles, (uint64_t)dem->num_sms * per_sm);
    if (dem->nclasses == 0) {
        GQ_CUDA(cudaMemsetAsync(dets_sm, 0, nshots * dem->DW * 4, dem->stream));
        if (obs_sm) GQ_CUDA(cudaMemsetAsync(obs_sm, 0, nshots * dem->OW * 4, dem->stream));
        if (a.cls.counts) GQ_CUDA(cudaMemsetAsync(a.cls.pred, 0, nshots * a.cls.OW * 4, dem->stream));   // every shot is defect-free
        return GQ_OK;
    }
    sample_kernel<<<grid, 256, smem, dem->stream>>>(a);
    GQ_CUDA(cudaGetLastError());
    return GQ_OK;
}

extern "C" int gq_extract(gq_dem *dem, const uint32_t *errs_mm, uint64_t nshots, uint32_t *dets_dm,
                          uint32_t *obs_dm) {
    if (!dem || !errs_mm || !dets_dm || (dem->O && !obs_dm)) GQ_FAIL(GQ_E_INVALID, "gq_extract: NULL argument");
    if (nshots == 0) return GQ_OK;
    GQ_CUDA(cudaSetDevice(dem->device));
    uint32_t SW = (uint32_t)((nshots + 31) / 32);
    uint32_t nrows = dem->D + dem->O;
    dim3 block(128), grid((SW + 127) / 128, nrows);
    if (nrows > 65535) GQ_FAIL(GQ_E_UNSUPPORTED, "gq_extract: more than 65535 detectors+observables");
    extract_dm_kernel<<<grid, block, 0, dem->stream>>>(dem->d_row_ptr, dem->d_row_mech, nrows, dem->D,
                                                       errs_mm, SW, dets_dm, obs_dm);
    GQ_CUDA(cudaGetLastError());
    return GQ_OK;
}

extern "C" int gq_transpose_dm_to_sm(gq_dem *dem, const uint32_t *dm, uint32_t nbits, uint64_t nshots, uint32_t *sm) {
    if (!dem || !dm || !sm) GQ_FAIL(GQ_E_INVALID, "gq_transpose_dm_to_sm: NULL argument");
    if (!nbits || !nshots) return GQ_OK;
    GQ_CUDA(cudaSetDevice(dem->device));
    uint32_t SW = (uint32_t)((nshots + 31) / 32), NW = (nbits + 31) / 32;
    uint64_t warps = (uint64_t)SW * NW;
    uint64_t blocks = (warps + 7) / 8;
    transpose_dm_to_sm_kernel<<<(uint32_t)blocks, 256, 0, dem->stream>>>(dm, nbits, nshots, SW, NW, sm);
    GQ_CUDA(cudaGetLastError());
    return GQ_OK;
}

extern "C" int gq_transpose_sm_to_dm(gq_dem *dem, const uint32_t *sm, uint32_t nbits, uint64_t nshots, uint32_t *dm) {
    if (!dem || !dm || !sm) GQ_FAIL(GQ_E_INVALID, "gq_transpose_sm_to_dm: NULL argument");
    if (!nbits || !nshots) return GQ_OK;
    GQ_CUDA(cudaSetDevice(dem->device));
    uint32_t SW = (uint32_t)((nshots + 31) / 32), NW = (nbits + 31) / 32;
    uint64_t warps = (uint64_t)SW * NW;
    uint64_t blocks = (warps + 7) / 8;
    transpose_sm_to_dm_kernel<<<(uint32_t)blocks, 256, 0, dem->stream>>>(sm, nbits, nshots, SW, NW, dm);
    GQ_CUDA(cudaGetLastError());
    return GQ_OK;
}

extern "C" int gq_count_errors(gq_dem *dem, const uint32_t *pred_sm, const uint32_t *obs_sm,
                               uint64_t nshots, uint64_t *count) {
    return gq_count_errors_discards(dem, pred_sm, obs_sm, nshots, 0u, count);
}

// count[0] += errors, count[1] += discards (count[1] is only touched when discard_word != 0)
int gq_count_errors_discards(gq_dem *dem, const uint32_t *pred_sm, const uint32_t *obs_sm, uint64_t nshots,
                             uint32_t discard_word, uint64_t *count) {
    if (!dem || !pred_sm || !obs_sm || !count) GQ_FAIL(GQ_E_INVALID, "gq_count_errors: NULL argument");
    if (nshots == 0) return GQ_OK;
    GQ_CUDA(cudaSetDevice(dem->device));
    uint32_t grid = (uint32_t)std::min<uint64_t>((nshots + 255) / 256, (uint64_t)dem->num_sms * 8);
    count_errors_kernel<<<grid, 256, 0, dem->stream>>>(pred_sm, obs_sm, nshots, dem->OW, discard_word,
                                                       (unsigned long long *)count);
    GQ_CUDA(cudaGetLastError());
    return GQ_OK;
}

extern "C" int gq_sync(gq_dem *dem) {
    if (!dem) GQ_FAIL(GQ_E_INVALID, "gq_sync: NULL argument");
    GQ_CUDA(cudaSetDevice(dem->device));
    GQ_CUDA(cudaStreamSynchronize(dem->stream));
    return GQ_OK;
}
