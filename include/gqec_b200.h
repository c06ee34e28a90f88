/*
 * gqec_b200 -- C ABI of the B200-native threshold-estimation hot path (libgqec_b200.so).
 *
 * The reference (adelshb/graphqec) has no FFI of its own: ThresholdLAB.collect_stats hands a task
 * generator to sinter.collect (graphqec/lab/threshold/threshold_lab.py:182-197), and all arithmetic
 * then happens inside third-party wheels.  Each entry point below therefore names the upstream call
 * it stands in for; INTEGRATION.md shows the ctypes
 * binding a graphqec maintainer would add at threshold_lab.py:182.
 *
 * Conventions: every function returns 0 on success or a negative code (GQ_E_*); the message of the
 * last failure on the calling thread is available from gq_last_error().  The caller owns every
 * buffer.  Handles are opaque, created by *_create and released by *_destroy; each handle owns one
 * CUDA stream; a handle is not thread-safe, distinct handles are independent.  "dev" pointers are
 * CUDA device pointers on the handle's device; "host" pointers are ordinary (ideally pinned) memory.
 *
 * Bit layouts.  "sm" (shot-major): uint32 words, row s = shot s, bit k of the row = bit (k & 31) of
 * word (k >> 5); a row of D detectors has DW = ceil(D/32) words.  Reinterpreted as bytes this is
 * exactly sinter's bit-packed layout ("b8": uint8[shots][ceil(D/8)], little-endian bits) with each
 * row zero-padded to a multiple of 4 bytes.  "dm" (detector-major): row d = detector d, bit s of
 * the row = shot s, SW = ceil(nshots/32) words per row -- the bit-sliced form (32 shots per word).
 */
#ifndef GQEC_B200_H
#define GQEC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GQ_OK 0
#define GQ_E_INVALID (-1)   /* bad argument */
#define GQ_E_CUDA (-2)      /* CUDA runtime failure (no device, launch error, ...) */
#define GQ_E_NOMEM (-3)
#define GQ_E_UNSUPPORTED (-4)
#define GQ_E_INTERNAL (-5)

#define GQ_NO_DET 0xFFFFFFFFu

#define GQ_DECODER_MATCHING 1   /* exact minimum-weight perfect matching (replaces PyMatching) */
#define GQ_DECODER_BP_MINSUM 2  /* min-sum belief propagation on the DEM Tanner graph */

typedef struct gq_dem gq_dem;
typedef struct gq_dec gq_dec;

typedef struct gq_decoder_params {
    int32_t weight_levels;     /* matching: integer weight of the heaviest edge (default 1000) */
    int32_t fast_capacity;     /* matching: defects handled by the fast tier; 0 = choose from the DEM */
    int32_t with_corrections;  /* matching: keep the next-hop table so corrections can be emitted */
    int32_t bp_max_iter;       /* BP: iteration cap (default 30) */
    float bp_alpha;            /* BP: min-sum normalisation factor (default 0.9) */
    int32_t legacy_tiers;      /* matching: 1 = skip the first tier (A/B testing only) */
    int32_t bp_osd;            /* BP: 1 = order-0 ordered-statistics decoding (OSD-0) of the shots BP leaves
                                  unconverged -- the "OSD" of the reference's decoder="bposd" (stimbposd) */
    int32_t bp_scratch;        /* BP kernel choice (tests, A/B): 0 = automatic (messages in shared memory when the model fits,
                                  else the compressed min-sum kernel, else the global-scratch kernel), 1 = force the
                                  global-scratch (shot-interleaved) kernel, 2 = force the compressed min-sum kernel */
    int32_t reserved[5];
} gq_decoder_params;

typedef struct gq_dem_info {
    uint32_t num_detectors, num_observables, num_errors;
    uint32_t det_words, obs_words;      /* DW, OW of the "sm" layout */
    uint32_t tile_shots;                /* shots per sampling tile: shot0 must be a multiple of it */
    uint32_t num_classes;               /* distinct probabilities */
    uint32_t num_segments;              /* RNG work items per tile */
    double sum_p;                       /* expected fired mechanisms per shot */
} gq_dem_info;

typedef struct gq_graph_info {
    uint32_t num_nodes, num_edges, num_boundary_edges;
    uint32_t fast_capacity, max_capacity;
    uint32_t max_distance;              /* largest finite entry of the distance table */
    uint32_t reserved[6];
} gq_graph_info;

const char *gq_last_error(void);
int gq_version(void);
int gq_device_count(void);

/* ---- DEM: replaces the object sinter's worker obtains from
 *      circuit.detector_error_model(decompose_errors=True, approximate_disjoint_errors=True)
 *      (SURVEY.md 8a row A6).  Mechanism i has probability p[i], flips detectors
 *      det_idx[det_indptr[i] .. det_indptr[i+1]) and observables obs_idx[...].  comp_* (nullable)
 *      lists its graphlike pieces (Stim's "^" separators): pieces comp_indptr[i]..comp_indptr[i+1],
 *      each (comp_a, comp_b or GQ_NO_DET, observable mask); pieces with comp_a == GQ_NO_DET are
 *      ignored by the matching-graph builder.  All arrays are host memory. */
int gq_dem_create(uint32_t num_detectors, uint32_t num_observables, uint32_t num_errors,
                  const double *p, const uint32_t *det_indptr, const uint32_t *det_idx,
                  const uint32_t *obs_indptr, const uint32_t *obs_idx,
                  const uint32_t *comp_indptr, const uint32_t *comp_a, const uint32_t *comp_b,
                  const uint64_t *comp_obs, int device, gq_dem **out);
void gq_dem_destroy(gq_dem *dem);
int gq_dem_get_info(const gq_dem *dem, gq_dem_info *info);

/* ---- circuit -> DEM, native host code (no device needed): replaces UPSTREAM
 *      stim.Circuit.detector_error_model(decompose_errors=..., approximate_disjoint_errors=True), which
 *      sinter calls for every task because the reference's Tasks carry no DEM
 *      (reference graphqec/lab/threshold/threshold_lab.py:151).
 *      stim_text: the circuit in Stim's text syntax (str(circuit); REPEAT blocks are unrolled); gate set
 *      R RX M MX MR MRX H X Y Z I S S_DAG CX CZ SWAP DEPOLARIZE1/2 X/Y/Z_ERROR DETECTOR
 *      OBSERVABLE_INCLUDE TICK QUBIT_COORDS SHIFT_COORDS.  Mechanisms come out sorted by (detector
 *      tuple, observable mask), equal symptoms merged.  decompose_errors != 0 also yields the graphlike
 *      pieces of every mechanism (GQ_E_UNSUPPORTED when one cannot be split -- sinter then falls back to
 *      the undecomposed model).  Export into caller-owned arrays sized from gq_dem_build_sizes:
 *      sizes = {D, O, M, nnz(H), nnz(L), number of pieces}; the comp_* arrays may be NULL when the model
 *      was built without decomposition.  The arrays are exactly gq_dem_create's inputs. */
typedef struct gq_dem_build gq_dem_build;
int gq_circuit_to_dem(const char *stim_text, int decompose_errors, gq_dem_build **out);
int gq_dem_build_sizes(const gq_dem_build *b, uint32_t sizes[6]);
int gq_dem_build_export(const gq_dem_build *b, double *p, uint32_t *det_indptr, uint32_t *det_idx,
                        uint32_t *obs_indptr, uint32_t *obs_idx, uint32_t *comp_indptr,
                        uint32_t *comp_a, uint32_t *comp_b, uint64_t *comp_obs);
void gq_dem_build_destroy(gq_dem_build *b);

/* ---- sampling + extraction: replaces
 *      circuit.compile_detector_sampler().sample(shots, bit_packed=True, separate_observables=True)
 *      (SURVEY.md 8a row A7).  Counter-based Philox4x32-10 keyed by (seed, global tile index,
 *      segment): the shots [shot0, shot0+nshots) come out the same whatever the batching or the
 *      number of GPUs.  shot0 must be a multiple of gq_dem_info.tile_shots.
 *      dets_sm: dev uint32[nshots][DW]; obs_sm: dev uint32[nshots][OW];
 *      errs_mm (nullable, debug/parity): dev uint32[num_errors][ceil(nshots/32)], must be zeroed by
 *      the caller; receives the sampled error vectors in mechanism-major bit-sliced form. */
int gq_sample(gq_dem *dem, uint64_t seed, uint64_t shot0, uint64_t nshots, uint32_t *dets_sm,
              uint32_t *obs_sm, uint32_t *errs_mm);

/* ---- sparse GF(2) extraction on explicit error vectors (the bit-exact parity hook):
 *      dets = H.e, obs = L.e, bit-sliced 32 shots per word.
 *      errs_mm: dev uint32[num_errors][SW]; dets_dm: dev uint32[D][SW]; obs_dm: dev uint32[O][SW]. */
int gq_extract(gq_dem *dem, const uint32_t *errs_mm, uint64_t nshots, uint32_t *dets_dm,
               uint32_t *obs_dm);

/* ---- bit-matrix transpose between the two layouts (dev pointers).
 *      dm: uint32[nbits][ceil(nshots/32)]  <->  sm: uint32[nshots][ceil(nbits/32)] */
int gq_transpose_dm_to_sm(gq_dem *dem, const uint32_t *dm, uint32_t nbits, uint64_t nshots, uint32_t *sm);
int gq_transpose_sm_to_dm(gq_dem *dem, const uint32_t *sm, uint32_t nbits, uint64_t nshots, uint32_t *dm);

/* ---- decoder: replaces sinter's compile_decoder_for_dem -> pymatching.Matching
 *      .from_detector_error_model(dem) (row A8) or stimbposd's BP+OSD (row A10): min-sum BP, and with params.bp_osd
 *      order-0 OSD for the shots BP leaves unconverged (stimbposd's default is product-sum BP + OSD-CS(60): the
 *      BASELINE north star asks for min-sum; higher OSD orders are not built). */
int gq_decoder_create(gq_dem *dem, int kind, const gq_decoder_params *params, gq_dec **out);
void gq_decoder_destroy(gq_dec *dec);
int gq_decoder_get_graph_info(const gq_dec *dec, gq_graph_info *info);
/* merged matching-graph edges, host arrays of num_edges entries each (any may be NULL):
 * node u, node v (GQ_NO_DET = boundary), integer weight, observable mask, merged probability. */
int gq_decoder_get_edges(const gq_dec *dec, uint32_t *u, uint32_t *v, uint32_t *weight,
                         uint64_t *obs_mask, double *prob);

/* ---- decode, device buffers ("sm" layout).  pred_sm: dev uint32[nshots][OW].
 *      weight_out (nullable): dev int32[nshots], total weight of the matching found (matching
 *      decoder) or BP iterations used (negative = not converged).
 *      flags_out (nullable): dev uint8[nshots], 0 = ok, 1 = no valid correction exists / BP did not
 *      converge, 2 = internal capacity error.
 *      corr_out (nullable; needs with_corrections): dev uint32[nshots][ceil(num_edges/32)] zeroed by
 *      the caller; receives the correction as a set of matching-graph edges (matching decoder) or
 *      uint32[nshots][ceil(num_errors/32)] error mechanisms (BP). */
int gq_decode(gq_dec *dec, const uint32_t *dets_sm, uint64_t nshots, uint32_t *pred_sm,
              int32_t *weight_out, uint8_t *flags_out, uint32_t *corr_out);

/* ---- decode, device buffers in the detector-major bit-sliced layout of gq_extract (32 shots per word):
 *      dets_dm: dev uint32[D][ceil(nshots/32)] -> pred_dm: dev uint32[max(O,1)][ceil(nshots/32)].
 *      corr_sm (nullable; SURVEY 8b "corr_optional"; needs with_corrections): the corrections, shot-major like
 *      gq_decode's corr_out -- dev uint32[nshots][ceil(num_edges/32)] (matching: edge set) or
 *      [nshots][ceil(num_errors/32)] (BP: mechanism set).  Works through the batch in slices of 2^20 shots. */
int gq_decode_dm(gq_dec *dec, const uint32_t *dets_dm, uint64_t nshots, uint32_t *pred_dm, uint32_t *corr_sm);

/* ---- decode, host buffers in sinter's layout: the body of
 *      CompiledDecoder.decode_shots_bit_packed(bit_packed_detection_event_data=...).
 *      dets_b8: host uint8[nshots][ceil(D/8)]; pred_b8: host uint8[nshots][ceil(O/8)]. */
int gq_decode_b8(gq_dec *dec, const uint8_t *dets_b8, uint64_t nshots, uint8_t *pred_b8);

/* ---- failure count (sinter's classify step, row A9): rows where pred != obs.
 *      count: dev uint64, incremented (not reset). */
int gq_count_errors(gq_dem *dem, const uint32_t *pred_sm, const uint32_t *obs_sm, uint64_t nshots,
                    uint64_t *count);

/* ---- the whole hot path for shots [shot0, shot0 + nshots): sample -> extract -> decode -> count,
 *      everything resident on the device; replaces one sinter worker batch (rows A7-A9).
 *      out (host): {shots, errors, discards}.  max_errors > 0 stops early at a chunk boundary once
 *      that many errors were seen (out[0] then holds the shots actually decoded). */
int gq_collect(gq_dem *dem, gq_dec *dec, uint64_t seed, uint64_t shot0, uint64_t nshots,
               uint64_t max_errors, uint64_t out[3]);

/* device-time of the last gq_collect / gq_decode_b8 call on this decoder, per stage, milliseconds:
 * {sample, decode, count, total, h2d, d2h, kernel launches} */
int gq_last_timing(const gq_dec *dec, double out_ms[8]);

/* the dominant kernel of the last gq_collect / gq_decode_b8 call (matching decoder: the first-tier
 * kernel most shots went through), timed with CUDA events on the decoder's stream:
 * {milliseconds summed over its launches, launches, shots it decoded, K = ceil(defects / 32) of its class} */
int gq_last_kernel_profile(const gq_dec *dec, double out[4]);

int gq_sync(gq_dem *dem);

#ifdef __cplusplus
}
#endif
#endif
