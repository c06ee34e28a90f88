// Internal declarations shared by the translation units of libgqec_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/gqec_b200.h"

void gq_set_error(const std::string &msg);

#define GQ_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            gq_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));              \
            return GQ_E_CUDA;                                                              \
        }                                                                                  \
    } while (0)

#define GQ_FAIL(code, msg)      \
    do {                        \
        gq_set_error(msg);      \
        return (code);          \
    } while (0)

// one probability class of the sampler (all mechanisms with the same p)
struct GqClass {
    float inv_log1m;      // 1 / log(1 - p)   (negative; -0 for p == 1)
    uint32_t mech_start;  // first mechanism (class-sorted order)
    uint32_t mech_count;
    uint32_t seg_len;     // flat (mechanism, shot) indices per RNG segment
    uint32_t seg_first;   // index of the class's first segment
    uint32_t flat_size;   // mech_count * tile_shots
};

struct gq_dem {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t D = 0, O = 0, M = 0;
    uint32_t DW = 0, OW = 0;
    uint32_t tile_shots = 0, tile_log2 = 0;
    uint32_t nclasses = 0, nsegments = 0;
    double sum_p = 0;
    int num_sms = 148;
    // host copies (original order)
    std::vector<double> p;
    std::vector<uint32_t> det_indptr, det_idx, obs_indptr, obs_idx;
    std::vector<uint32_t> comp_indptr, comp_a, comp_b;
    std::vector<uint64_t> comp_obs;
    bool has_components = false;
    // device: class-sorted mechanisms for the sampler
    GqClass *d_classes = nullptr;
    uint32_t *d_seg_first = nullptr;   // [nclasses + 1]
    uint32_t *d_mech_ptr = nullptr;    // [M + 1]
    uint32_t *d_mech_det = nullptr;    // [nnz]
    uint32_t *d_mech_obs = nullptr;    // [M] observable mask
    uint32_t *d_mech_orig = nullptr;   // [M] original mechanism index
    uint32_t *d_seg_rec = nullptr;     // [nsegments][4] one-load segment records of the sampler
    uint32_t *d_mech_rec = nullptr;    // [M][4] one-load sampler records (null when a mechanism has > 4 detectors)
    // device: transposed CSR (row = detector, then observables) for the bit-sliced extractor
    uint32_t *d_row_ptr = nullptr;     // [D + O + 1]
    uint32_t *d_row_mech = nullptr;    // [nnz_H + nnz_L]
};

struct gq_dec {
    gq_dem *dem = nullptr;
    int kind = 0;
    gq_decoder_params params{};
    double timing[8] = {0};
    double kprof[4] = {0};             // dominant first-tier kernel of the last call: {ms, launches, shots, K}
    cudaEvent_t kev0 = nullptr, kev1 = nullptr;
    cudaStream_t side_streams[16] = {};   // first tier: the smaller classes' kernels
    cudaEvent_t side_done[16] = {};
    // ---- matching
    uint32_t num_edges = 0, num_boundary = 0;
    std::vector<uint32_t> eu, ev, ew;
    std::vector<uint64_t> eobs;
    std::vector<double> eprob;
    uint32_t fast_cap = 0, max_cap = 0, max_dist = 0;
    bool weight_levels_auto = true;    // the caller left weight_levels at its default (the decoder may coarsen the grid)
    bool big_classes = false;          // first-tier classes K = 48, 64 (1025 .. 2048 defects) usable: max_dist < 2^15
    uint16_t *d_dist = nullptr;        // [D][D+1]
    uint8_t *d_pobs = nullptr;         // [D][D+1] observable mask of the shortest path
    uint32_t *d_nbr = nullptr;         // [D+2][NBR_LEN] nearest detectors of each detector (distance << 16 | detector)
    uint16_t *d_next = nullptr;        // [D][D+1] next node on the shortest path (0xFFFF = boundary hop)
    uint32_t *d_adj_ptr = nullptr;     // [D+1]   adjacency for corrections
    uint32_t *d_adj_node = nullptr;    // neighbour (D = boundary)
    uint32_t *d_adj_edge = nullptr;    // edge id
    void *d_ws_fast = nullptr;         // per-warp workspaces
    void *d_ws_big = nullptr;
    size_t ws_fast_bytes = 0, ws_big_bytes = 0;
    int nwarps_fast = 0, nwarps_big = 0;
    uint32_t *d_overflow = nullptr;    // [1 + capacity] overflow shot list
    void *d_recs = nullptr;            // first tier: hand-off records between the start and the solve kernels
    size_t recs_bytes = 0;
    uint32_t *d_lists = nullptr;       // first tier: [64] class counts + work cursors, then one shot list per class
    uint32_t overflow_cap = 0;
    // ---- BP
    uint32_t *d_chk_ptr = nullptr, *d_chk_var = nullptr, *d_var_ptr = nullptr, *d_var_edge = nullptr;
    float *d_prior = nullptr;
    uint32_t *d_var_obs = nullptr;
    uint32_t *d_var_chk = nullptr;     // detectors of every variable (columns of H), for OSD
    uint32_t *d_osd_cand = nullptr, *d_osd_fail = nullptr, *d_osd_scratch = nullptr;
    uint64_t osd_shots = 0;            // shots handed to OSD so far
    uint64_t osd_giveups = 0;          // of those, shots OSD could not solve (syndrome outside the column space of H)
    float *d_bp_msgs = nullptr;
    size_t bp_msgs_per_block = 0;
    int bp_blocks = 0;
    uint32_t bp_nnz = 0;
    // compressed min-sum kernel (bp.cu: bp_cms_kernel): the Tanner graph as degree-sorted tiles of 32 variables
    bool cms = false;
    uint32_t cms_grp_tile[18] = {0}, cms_grp_word[17] = {0};
    uint32_t *d_cms_var = nullptr, *d_cms_obs = nullptr, *d_cms_cursor = nullptr;
    float *d_cms_prior = nullptr, *d_cms_post = nullptr;
    uint16_t *d_cms_chk = nullptr;
    uint32_t cms_ntiles = 0, cms_FW = 0, cms_SW = 0;
    size_t cms_smem = 0;
    int cms_blocks = 0, cms_threads = 0;
    // ---- scratch for gq_collect / gq_decode_b8
    uint32_t *d_dets = nullptr, *d_obs = nullptr, *d_pred = nullptr;
    uint64_t scratch_shots = 0;
    uint64_t *d_count = nullptr;
    uint8_t *d_b8 = nullptr;
    cudaStream_t copy_stream = nullptr;            // gq_decode_b8: H2D of the next chunk overlaps the decode
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    uint64_t b8_bytes = 0;
};

// first-tier classes of the matching decoder (matching.cu): K = ceil(defects / 32) slots per lane; classes 0..7: K = 1..8
// (up to 256 defects), 8..10: K = 10, 12, 16 (up to 512), 11, 12: K = 24, 32 (up to 1024), 13, 14: K = 48, 64 (up to 2048;
// most of the per-lane state lives in local memory there); GQ_MT_CLASSES = beyond the tier.
// Shared with the sampler and the b8 re-stride kernel, which classify the shots they write (no separate pass).
static constexpr int GQ_MT_CLASSES = 15;
static constexpr int GQ_MT_CAP = 2048;
__host__ __device__ constexpr int gq_mt_class_of(int n) {   // n >= 1 defects
    return n <= 256 ? (n - 1) >> 5
                    : (n <= 320 ? 8 : (n <= 384 ? 9 : (n <= 512 ? 10 : (n <= 768 ? 11 : (n <= 1024 ? 12 : (n <= 1536 ? 13 : (n <= 2048 ? 14 : GQ_MT_CLASSES)))))));
}
struct GqClassifyOut {
    uint32_t *counts = nullptr;     // [GQ_MT_CLASSES] shots per class (null: no classification)
    uint32_t *lists = nullptr;      // [GQ_MT_CLASSES][list_stride] shot indices, relative to the call's first shot
    uint32_t list_stride = 0;
    uint32_t *overflow = nullptr;   // [0] count, [1] max defects, [2..] shots beyond the first tier
    uint32_t overflow_cap = 0;
    uint32_t *pred = nullptr;       // predictions: rows of defect-free shots are zeroed here
    uint32_t OW = 0;
};
// dem.cu: gq_sample with the classification fused into the tile flush
int gq_sample_classified(gq_dem *dem, uint64_t seed, uint64_t shot0, uint64_t nshots, uint32_t *dets_sm,
                         uint32_t *obs_sm, uint32_t *errs_mm, const GqClassifyOut *cls);

int gq_count_errors_discards(gq_dem *dem, const uint32_t *pred_sm, const uint32_t *obs_sm, uint64_t nshots,
                             uint32_t discard_word, uint64_t *count);

// matching.cu
int gq_matching_create(gq_dec *dec);
// zeroes the class counts / work cursors / overflow header of the next slice and says where a producer kernel
// (sampler, re-stride) should append the shots; the following gq_matching_decode(..., preclassified = true) consumes them
int gq_matching_classify_targets(gq_dec *dec, uint32_t *pred_sm, GqClassifyOut *out);
void gq_matching_destroy(gq_dec *dec);
int gq_matching_decode(gq_dec *dec, const uint32_t *dets_sm, uint64_t nshots, uint32_t *pred_sm,
                       int32_t *weight_out, uint8_t *flags_out, uint32_t *corr_out, int *launches,
                       bool preclassified = false);
// bp.cu
int gq_bp_create(gq_dec *dec);
void gq_bp_destroy(gq_dec *dec);
int gq_bp_decode(gq_dec *dec, const uint32_t *dets_sm, uint64_t nshots, uint32_t *pred_sm,
                 int32_t *weight_out, uint8_t *flags_out, uint32_t *corr_out, int *launches);
