// Decoder handle, host-buffer entry point (sinter's decode_shots_bit_packed) and the fused
// sample -> decode -> count pipeline (one sinter worker batch).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "gq_common.h"

// sinter's b8 rows (ceil(D/8) bytes, unpadded) -> 4-byte aligned "sm" rows.  One warp per row, lanes = words; the same
// pass counts the row's defects and appends the shot to the matching decoder's per-class list (cls.counts != null),
// with one global atomic per (block iteration, class) -- what classify_kernel would do in a pass of its own.
__global__ void __launch_bounds__(256) b8_to_sm_kernel(const uint8_t *src, uint32_t row_bytes, uint32_t DW, uint32_t D, uint64_t nshots,
                                                       uint32_t *dst, GqClassifyOut cls) {
    __shared__ uint32_t cl_cnt[16], cl_base[16];
    __shared__ uint16_t cl_pos[8];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t tail_mask = (D & 31) ? (1u << (D & 31)) - 1u : 0xffffffffu;
    for (uint64_t s0 = (uint64_t)blockIdx.x * 8; s0 < nshots; s0 += (uint64_t)gridDim.x * 8) {
        const uint64_t s = s0 + wib;
        if (cls.counts) {
            if (threadIdx.x < 16) cl_cnt[threadIdx.x] = 0;
            __syncthreads();
        }
        uint32_t n = 0;
        if (s < nshots) {
            const uint8_t *r = src + s * row_bytes;
            for (uint32_t w = lane; w < DW; w += 32) {
                const uint32_t left = row_bytes - w * 4;
                uint32_t v = r[4 * w];
                if (left > 1) v |= (uint32_t)r[4 * w + 1] << 8;
                if (left > 2) v |= (uint32_t)r[4 * w + 2] << 16;
                if (left > 3) v |= (uint32_t)r[4 * w + 3] << 24;
                if (w == DW - 1) v &= tail_mask;
                dst[s * DW + w] = v;
                n += __popc(v);
            }
        }
        if (!cls.counts) continue;
        n = __reduce_add_sync(0xffffffffu, n);
        if (lane == 0) {
            cl_pos[wib] = 0xffffu;
            if (s < nshots) {
                if (n == 0) {
                    for (uint32_t w = 0; w < cls.OW; ++w) cls.pred[s * cls.OW + w] = 0u;
                } else {
                    const uint32_t c = (uint32_t)gq_mt_class_of((int)n);
                    cl_pos[wib] = (uint16_t)((c << 12) | atomicAdd(&cl_cnt[c], 1u));
                    if (c == (uint32_t)GQ_MT_CLASSES) atomicMax(&cls.overflow[1], n);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x <= (uint32_t)GQ_MT_CLASSES && cl_cnt[threadIdx.x])
            cl_base[threadIdx.x] = atomicAdd(threadIdx.x < (uint32_t)GQ_MT_CLASSES ? &cls.counts[threadIdx.x] : &cls.overflow[0], cl_cnt[threadIdx.x]);
        __syncthreads();
        if (lane == 0 && cl_pos[wib] != 0xffffu) {
            const uint32_t c = cl_pos[wib] >> 12, at = cl_base[c] + (cl_pos[wib] & 0xfffu);
            if (c < (uint32_t)GQ_MT_CLASSES) cls.lists[(size_t)c * cls.list_stride + at] = (uint32_t)s;
            else if (at < cls.overflow_cap) cls.overflow[2 + at] = (uint32_t)s;
        }
        __syncthreads();
    }
}

__global__ void sm_to_b8_kernel(const uint32_t *src, uint32_t OW, uint32_t row_bytes, uint64_t nshots, uint8_t *dst) {
    const uint64_t total = nshots * row_bytes;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = i / row_bytes;
        const uint32_t b = (uint32_t)(i - s * row_bytes);
        dst[i] = (uint8_t)(src[s * OW + (b >> 2)] >> (8 * (b & 3)));
    }
}

#ifndef GQ_COLLECT_CHUNK_LOG2
#define GQ_COLLECT_CHUNK_LOG2 22   // = matching.cu's decode slice: per-slice launch overheads and tails amortise over 4M shots
#endif

// timing events of one call, destroyed on every exit path
struct EventSet {
    cudaEvent_t e[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ~EventSet() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
    cudaError_t make(int n) {
        for (int i = 0; i < n; ++i) { cudaError_t rc = cudaEventCreate(&e[i]); if (rc != cudaSuccess) return rc; }
        return cudaSuccess;
    }
};

static void default_params(gq_decoder_params *p) {
    memset(p, 0, sizeof(*p));
    p->weight_levels = 1000;
    p->bp_max_iter = 30;
    p->bp_alpha = 0.9f;
}

extern "C" int gq_decoder_create(gq_dem *dem, int kind, const gq_decoder_params *params, gq_dec **out) {
    if (!dem || !out) GQ_FAIL(GQ_E_INVALID, "gq_decoder_create: NULL argument");
    *out = nullptr;
    GQ_CUDA(cudaSetDevice(dem->device));
    gq_dec *dec = new gq_dec();
    dec->dem = dem;
    dec->kind = kind;
    default_params(&dec->params);
    if (params) {
        if (params->weight_levels > 0) { dec->params.weight_levels = params->weight_levels; dec->weight_levels_auto = false; }
        dec->params.fast_capacity = params->fast_capacity;
        dec->params.with_corrections = params->with_corrections;
        dec->params.legacy_tiers = params->legacy_tiers;
        if (params->bp_max_iter > 0) dec->params.bp_max_iter = params->bp_max_iter;
        if (params->bp_alpha > 0) dec->params.bp_alpha = params->bp_alpha;
        dec->params.bp_osd = params->bp_osd ? 1 : 0;
        dec->params.bp_scratch = params->bp_scratch;
    }
    int rc;
    if (kind == GQ_DECODER_MATCHING) rc = gq_matching_create(dec);
    else if (kind == GQ_DECODER_BP_MINSUM) rc = gq_bp_create(dec);
    else { delete dec; GQ_FAIL(GQ_E_INVALID, "gq_decoder_create: unknown decoder kind"); }
    if (rc) { gq_decoder_destroy(dec); return rc; }
    if (cudaMalloc((void **)&dec->d_count, 16) != cudaSuccess) { gq_decoder_destroy(dec); GQ_FAIL(GQ_E_NOMEM, "gq_decoder_create: out of device memory"); }
    *out = dec;
    return GQ_OK;
}

extern "C" void gq_decoder_destroy(gq_dec *dec) {
    if (!dec) return;
    cudaSetDevice(dec->dem->device);
    if (dec->kind == GQ_DECODER_MATCHING) gq_matching_destroy(dec);
    else if (dec->kind == GQ_DECODER_BP_MINSUM) gq_bp_destroy(dec);
    cudaFree(dec->d_dets); cudaFree(dec->d_obs); cudaFree(dec->d_pred); cudaFree(dec->d_count);
    cudaFree(dec->d_b8);
    if (dec->copy_stream) {
        cudaStreamDestroy(dec->copy_stream);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(dec->ev_copied[i]); cudaEventDestroy(dec->ev_consumed[i]); }
    }
    delete dec;
}

extern "C" int gq_decoder_get_graph_info(const gq_dec *dec, gq_graph_info *info) {
    if (!dec || !info) GQ_FAIL(GQ_E_INVALID, "gq_decoder_get_graph_info: NULL argument");
    memset(info, 0, sizeof(*info));
    info->num_nodes = dec->dem->D;
    info->num_edges = dec->num_edges;
    info->num_boundary_edges = dec->num_boundary;
    info->fast_capacity = dec->fast_cap;
    info->max_capacity = dec->max_cap;
    info->max_distance = dec->max_dist;
    return GQ_OK;
}

extern "C" int gq_decoder_get_edges(const gq_dec *dec, uint32_t *u, uint32_t *v, uint32_t *weight,
                                    uint64_t *obs_mask, double *prob) {
    if (!dec) GQ_FAIL(GQ_E_INVALID, "gq_decoder_get_edges: NULL argument");
    size_t n = dec->num_edges;
    if (u) memcpy(u, dec->eu.data(), n * 4);
    if (v) memcpy(v, dec->ev.data(), n * 4);
    if (weight) memcpy(weight, dec->ew.data(), n * 4);
    if (obs_mask) memcpy(obs_mask, dec->eobs.data(), n * 8);
    if (prob) memcpy(prob, dec->eprob.data(), n * 8);
    return GQ_OK;
}

static int decode_dispatch(gq_dec *dec, const uint32_t *dets_sm, uint64_t nshots, uint32_t *pred_sm,
                           int32_t *weight_out, uint8_t *flags_out, uint32_t *corr_out, int *launches,
                           bool preclassified = false) {
    if (dec->kind == GQ_DECODER_MATCHING)
        return gq_matching_decode(dec, dets_sm, nshots, pred_sm, weight_out, flags_out, corr_out, launches, preclassified);
    return gq_bp_decode(dec, dets_sm, nshots, pred_sm, weight_out, flags_out, corr_out, launches);
}

extern "C" int gq_decode(gq_dec *dec, const uint32_t *dets_sm, uint64_t nshots, uint32_t *pred_sm,
                         int32_t *weight_out, uint8_t *flags_out, uint32_t *corr_out) {
    if (!dec || !dets_sm || !pred_sm) GQ_FAIL(GQ_E_INVALID, "gq_decode: NULL argument");
    if (nshots == 0) return GQ_OK;
    GQ_CUDA(cudaSetDevice(dec->dem->device));
    int launches = 0;
    return decode_dispatch(dec, dets_sm, nshots, pred_sm, weight_out, flags_out, corr_out, &launches);
}

static int ensure_scratch(gq_dec *dec, uint64_t nshots);

// detector-major in, detector-major out (32 shots per word): what the bit-sliced extractor (gq_extract) produces
// and SURVEY 8b lists as gq_decode_dm.  Two bit-matrix transposes around the decoder, in slices of 2^20 shots (a
// multiple of 32, so a slice starts on a word boundary of the detector-major rows) so that the scratch stays bounded.
__global__ void dm_slice_gather_kernel(const uint32_t *dm, uint32_t nrows, uint64_t SW, uint64_t w0, uint32_t sw, uint32_t *out) {
    const uint64_t total = (uint64_t)nrows * sw;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / sw, c = i - r * sw;
        out[i] = dm[r * SW + w0 + c];
    }
}
__global__ void dm_slice_scatter_kernel(const uint32_t *in, uint32_t nrows, uint64_t SW, uint64_t w0, uint32_t sw, uint32_t *dm) {
    const uint64_t total = (uint64_t)nrows * sw;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / sw, c = i - r * sw;
        dm[r * SW + w0 + c] = in[i];
    }
}

extern "C" int gq_decode_dm(gq_dec *dec, const uint32_t *dets_dm, uint64_t nshots, uint32_t *pred_dm, uint32_t *corr_sm) {
    if (!dec || !dets_dm || !pred_dm) GQ_FAIL(GQ_E_INVALID, "gq_decode_dm: NULL argument");
    if (nshots == 0) return GQ_OK;
    gq_dem *dem = dec->dem;
    GQ_CUDA(cudaSetDevice(dem->device));
    const uint64_t slice = 1ull << 20, SW = (nshots + 31) / 32;
    const uint32_t orows = std::max<uint32_t>(dem->O, 1);
    const size_t corr_words = dec->kind == GQ_DECODER_MATCHING ? (dec->num_edges + 31) / 32 : (dem->M + 31) / 32;
    int rc = ensure_scratch(dec, std::min(slice, nshots));
    if (rc) return rc;
    // staging for one slice of the detector-major rows (contiguous [rows][slice / 32])
    const size_t stage_words = (size_t)(dem->D + orows) * (std::min(slice, nshots) + 31) / 32 + 64;
    uint32_t *stage = nullptr;
    const bool sliced = nshots > slice;
    if (sliced) GQ_CUDA(cudaMalloc((void **)&stage, stage_words * 4));
    int launches = 0;
    for (uint64_t s0 = 0; s0 < nshots && rc == 0; s0 += slice) {
        const uint64_t ns = std::min(slice, nshots - s0);
        const uint32_t sw = (uint32_t)((ns + 31) / 32);
        const uint32_t *src = dets_dm;
        uint32_t *dst = pred_dm;
        if (sliced) {
            dm_slice_gather_kernel<<<dem->num_sms * 4, 256, 0, dem->stream>>>(dets_dm, dem->D, SW, s0 / 32, sw, stage);
            src = stage; dst = stage + (size_t)dem->D * sw;
        }
        if ((rc = gq_transpose_dm_to_sm(dem, src, dem->D, ns, dec->d_dets))) break;
        if ((rc = decode_dispatch(dec, dec->d_dets, ns, dec->d_pred, nullptr, nullptr, corr_sm ? corr_sm + s0 * corr_words : nullptr, &launches))) break;
        if ((rc = gq_transpose_sm_to_dm(dem, dec->d_pred, orows, ns, dst))) break;
        if (sliced) dm_slice_scatter_kernel<<<dem->num_sms, 256, 0, dem->stream>>>(dst, orows, SW, s0 / 32, sw, pred_dm);
    }
    if (stage) { cudaStreamSynchronize(dem->stream); cudaFree(stage); }
    return rc;
}

static int ensure_scratch(gq_dec *dec, uint64_t nshots) {
    if (dec->scratch_shots >= nshots) return GQ_OK;
    cudaFree(dec->d_dets); cudaFree(dec->d_obs); cudaFree(dec->d_pred);
    dec->d_dets = dec->d_obs = dec->d_pred = nullptr;
    dec->scratch_shots = 0;
    gq_dem *dem = dec->dem;
    GQ_CUDA(cudaMalloc((void **)&dec->d_dets, nshots * dem->DW * 4));
    GQ_CUDA(cudaMalloc((void **)&dec->d_obs, nshots * dem->OW * 4));
    GQ_CUDA(cudaMalloc((void **)&dec->d_pred, nshots * dem->OW * 4));
    dec->scratch_shots = nshots;
    return GQ_OK;
}

extern "C" int gq_decode_b8(gq_dec *dec, const uint8_t *dets_b8, uint64_t nshots, uint8_t *pred_b8) {
    if (!dec || !dets_b8 || !pred_b8) GQ_FAIL(GQ_E_INVALID, "gq_decode_b8: NULL argument");
    if (nshots == 0) return GQ_OK;
    gq_dem *dem = dec->dem;
    GQ_CUDA(cudaSetDevice(dem->device));
    const size_t db = (dem->D + 7) / 8, ob = std::max<size_t>(1, (dem->O + 7) / 8);
    // Software pipeline over chunks: the H2D copy of chunk c+1 runs on a second stream while chunk c is
    // decoded (the decoder synchronises its own stream with the host; the copy stream keeps going).
    // Two raw-input buffers; a buffer is free again once its rows were re-strided into d_dets.
    // Chunk schedule: 2^F, 2^(F+1), ... doubling up to 2^X shots, then 2^X each (F = 18, X = 20 by default;
    // GQ_B8_FIRST_LOG2 / GQ_B8_MAX_LOG2 in the environment override).  The first copy is the only one that cannot
    // overlap anything, so it is short; every later copy hides behind the decode of the chunk before it as long as
    // PCIe delivers a chunk's 165 B/shot before the decoder is through the chunk before it: while the size doubles
    // that needs twice the decoder's 22 GB/s, from 2^X on only 22 GB/s.  Every chunk costs ~0.5 ms of launches,
    // kernel tails and host synchronisations, so chunks should not be small either.  Measured on one B200
    // (scripts/e2e_probe.py, 2^22 shots, ms per call): (16,19) 38.3, (17,20) 35.9, (18,20) 35.5, (18,21) 34.8,
    // (18,22) 34.7, one chunk 44.2; the doubling up to 2^21 of round 1 exposed 4.5 ms of copies when eight ranks
    // shared the host's memory bandwidth, hence X = 20.
    // A remainder shorter than half a chunk rides with the chunk before it.
    std::vector<uint64_t> cstart;
    uint64_t chunk = 0;
    {
        static const int first_log2 = []() { const char *e = getenv("GQ_B8_FIRST_LOG2"); int v = e ? atoi(e) : 18; return std::min(std::max(v, 10), 24); }();
        static const int max_log2 = []() { const char *e = getenv("GQ_B8_MAX_LOG2"); int v = e ? atoi(e) : 20; return std::min(std::max(v, 10), GQ_COLLECT_CHUNK_LOG2); }();
        uint64_t at = 0, sz = 1ull << std::min(first_log2, max_log2);
        const uint64_t cap = 1ull << max_log2;
        while (at < nshots) {
            uint64_t ns = std::min(sz, nshots - at);
            if (nshots - at - ns < sz / 2) ns = nshots - at;
            cstart.push_back(at);
            chunk = std::max(chunk, ns);
            at += ns;
            if (sz < cap) sz <<= 1;
        }
        cstart.push_back(nshots);
    }
    const uint64_t nchunks = cstart.size() - 1;
    if (!dec->copy_stream) {
        GQ_CUDA(cudaStreamCreateWithFlags(&dec->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            GQ_CUDA(cudaEventCreateWithFlags(&dec->ev_copied[i], cudaEventDisableTiming));
            GQ_CUDA(cudaEventCreateWithFlags(&dec->ev_consumed[i], cudaEventDisableTiming));
        }
    }
    EventSet evs;
    GQ_CUDA(evs.make(6));
    cudaEvent_t e0 = evs.e[0], e1 = evs.e[1], e2 = evs.e[2], e3 = evs.e[3], ea = evs.e[4], ez = evs.e[5];
    double t_wait = 0, t_dec = 0, t_d2h = 0;
    int launches = 0;
    memset(dec->kprof, 0, sizeof(dec->kprof));
    int rc = ensure_scratch(dec, std::min(chunk, nshots));
    if (rc) return rc;
    const size_t need = chunk * (2 * db + ob);
    if (dec->b8_bytes < need) {
        cudaFree(dec->d_b8);
        dec->d_b8 = nullptr; dec->b8_bytes = 0;
        GQ_CUDA(cudaMalloc((void **)&dec->d_b8, need));
        dec->b8_bytes = need;
    }
    uint8_t *d_in[2] = {dec->d_b8, dec->d_b8 + chunk * db};
    uint8_t *d_out = dec->d_b8 + 2 * chunk * db;
    auto issue_copy = [&](uint64_t c) -> int {
        const uint64_t s0 = cstart[c], ns = cstart[c + 1] - s0;
        const int buf = (int)(c & 1);
        if (c >= 2) GQ_CUDA(cudaStreamWaitEvent(dec->copy_stream, dec->ev_consumed[buf], 0));
        GQ_CUDA(cudaMemcpyAsync(d_in[buf], dets_b8 + s0 * db, ns * db, cudaMemcpyHostToDevice, dec->copy_stream));
        GQ_CUDA(cudaEventRecord(dec->ev_copied[buf], dec->copy_stream));
        return GQ_OK;
    };
    GQ_CUDA(cudaEventRecord(ea, dem->stream));
    rc = issue_copy(0);
    if (rc) return rc;
    for (uint64_t c = 0; c < nchunks; ++c) {
        const uint64_t s0 = cstart[c], ns = cstart[c + 1] - s0;
        const int buf = (int)(c & 1);
        if (c + 1 < nchunks) { rc = issue_copy(c + 1); if (rc) return rc; }
        GQ_CUDA(cudaEventRecord(e0, dem->stream));
        GQ_CUDA(cudaStreamWaitEvent(dem->stream, dec->ev_copied[buf], 0));
        // the b8 rows arrive contiguous; re-stride them to 4-byte aligned rows on the device, classifying on the way
        GqClassifyOut cls;
        const bool fused_cls = dec->kind == GQ_DECODER_MATCHING && !dec->params.legacy_tiers && ns <= (1ull << GQ_COLLECT_CHUNK_LOG2);
        if (fused_cls && (rc = gq_matching_classify_targets(dec, dec->d_pred, &cls))) return rc;
        b8_to_sm_kernel<<<dem->num_sms * 8, 256, 0, dem->stream>>>(d_in[buf], (uint32_t)db, dem->DW, dem->D, ns, dec->d_dets, cls);
        GQ_CUDA(cudaGetLastError());
        ++launches;
        GQ_CUDA(cudaEventRecord(dec->ev_consumed[buf], dem->stream));
        GQ_CUDA(cudaEventRecord(e1, dem->stream));
        rc = decode_dispatch(dec, dec->d_dets, ns, dec->d_pred, nullptr, nullptr, nullptr, &launches, fused_cls);
        if (rc) return rc;
        GQ_CUDA(cudaEventRecord(e2, dem->stream));
        sm_to_b8_kernel<<<dem->num_sms * 2, 256, 0, dem->stream>>>(dec->d_pred, dem->OW, (uint32_t)ob, ns, d_out);
        GQ_CUDA(cudaGetLastError());
        ++launches;
        GQ_CUDA(cudaMemcpyAsync(pred_b8 + s0 * ob, d_out, ns * ob, cudaMemcpyDeviceToHost, dem->stream));
        GQ_CUDA(cudaEventRecord(e3, dem->stream));
        GQ_CUDA(cudaStreamSynchronize(dem->stream));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1); t_wait += ms;   // exposed (not overlapped) copy time + re-stride
        cudaEventElapsedTime(&ms, e1, e2); t_dec += ms;
        cudaEventElapsedTime(&ms, e2, e3); t_d2h += ms;
    }
    GQ_CUDA(cudaEventRecord(ez, dem->stream));
    GQ_CUDA(cudaStreamSynchronize(dem->stream));
    GQ_CUDA(cudaStreamSynchronize(dec->copy_stream));
    float total = 0;
    cudaEventElapsedTime(&total, ea, ez);
    double *t = dec->timing;
    t[0] = 0; t[1] = t_dec; t[2] = 0; t[3] = total; t[4] = t_wait; t[5] = t_d2h; t[6] = launches; t[7] = 0;
    return GQ_OK;
}

extern "C" int gq_collect(gq_dem *dem, gq_dec *dec, uint64_t seed, uint64_t shot0, uint64_t nshots,
                          uint64_t max_errors, uint64_t out[3]) {
    if (!dem || !dec || !out) GQ_FAIL(GQ_E_INVALID, "gq_collect: NULL argument");
    if (dec->dem != dem) GQ_FAIL(GQ_E_INVALID, "gq_collect: decoder was built for another DEM");
    if (shot0 % dem->tile_shots) GQ_FAIL(GQ_E_INVALID, "gq_collect: shot0 must be a multiple of tile_shots");
    out[0] = out[1] = out[2] = 0;
    if (nshots == 0) return GQ_OK;
    GQ_CUDA(cudaSetDevice(dem->device));
    const uint64_t chunk = 1ull << GQ_COLLECT_CHUNK_LOG2;  // multiple of every tile size
    int rc = ensure_scratch(dec, std::min(chunk, nshots));
    if (rc) return rc;
    EventSet evs;
    GQ_CUDA(evs.make(4));
    cudaEvent_t *ev = evs.e;
    GQ_CUDA(cudaMemsetAsync(dec->d_count, 0, 16, dem->stream));
    // BP + OSD marks a shot it could not solve with an all-ones prediction word: sinter's "discards"
    const uint32_t discard_word = (dec->kind == GQ_DECODER_BP_MINSUM && dec->params.bp_osd && dem->O < 32) ? 0xFFFFFFFFu : 0u;
    double t_s = 0, t_d = 0, t_c = 0;
    int launches = 0;
    memset(dec->kprof, 0, sizeof(dec->kprof));
    uint64_t done = 0, errors = 0, counts[2] = {0, 0};
    while (done < nshots) {
        uint64_t ns = std::min(chunk, nshots - done);
        GQ_CUDA(cudaEventRecord(ev[0], dem->stream));
        // matching: the sampler sorts the shots it writes into the decoder's per-class lists (fused classification)
        GqClassifyOut cls;
        const bool fused_cls = dec->kind == GQ_DECODER_MATCHING && !dec->params.legacy_tiers;
        if (fused_cls && (rc = gq_matching_classify_targets(dec, dec->d_pred, &cls))) return rc;
        rc = gq_sample_classified(dem, seed, shot0 + done, ns, dec->d_dets, dec->d_obs, nullptr, fused_cls ? &cls : nullptr);
        if (rc) return rc;
        ++launches;
        GQ_CUDA(cudaEventRecord(ev[1], dem->stream));
        rc = decode_dispatch(dec, dec->d_dets, ns, dec->d_pred, nullptr, nullptr, nullptr, &launches, fused_cls);
        if (rc) return rc;
        GQ_CUDA(cudaEventRecord(ev[2], dem->stream));
        rc = gq_count_errors_discards(dem, dec->d_pred, dec->d_obs, ns, discard_word, dec->d_count);
        if (rc) return rc;
        ++launches;
        GQ_CUDA(cudaEventRecord(ev[3], dem->stream));
        done += ns;
        if (max_errors || done >= nshots) {
            GQ_CUDA(cudaMemcpyAsync(counts, dec->d_count, 16, cudaMemcpyDeviceToHost, dem->stream));
        }
        GQ_CUDA(cudaStreamSynchronize(dem->stream));
        errors = counts[0];
        float ms;
        cudaEventElapsedTime(&ms, ev[0], ev[1]); t_s += ms;
        cudaEventElapsedTime(&ms, ev[1], ev[2]); t_d += ms;
        cudaEventElapsedTime(&ms, ev[2], ev[3]); t_c += ms;
        if (max_errors && errors >= max_errors) break;
    }
    out[0] = done; out[1] = errors; out[2] = counts[1];
    double *t = dec->timing;
    t[0] = t_s; t[1] = t_d; t[2] = t_c; t[3] = t_s + t_d + t_c; t[4] = 0; t[5] = 8; t[6] = launches; t[7] = 0;
    return GQ_OK;
}

extern "C" int gq_last_kernel_profile(const gq_dec *dec, double out[4]) {
    if (!dec || !out) GQ_FAIL(GQ_E_INVALID, "gq_last_kernel_profile: NULL argument");
    memcpy(out, dec->kprof, sizeof(double) * 4);
    return GQ_OK;
}

extern "C" int gq_last_timing(const gq_dec *dec, double out_ms[8]) {
    if (!dec || !out_ms) GQ_FAIL(GQ_E_INVALID, "gq_last_timing: NULL argument");
    memcpy(out_ms, dec->timing, sizeof(double) * 8);
    return GQ_OK;
}
