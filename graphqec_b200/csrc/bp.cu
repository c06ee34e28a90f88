// K3b: batched min-sum belief propagation on the DEM Tanner graph, one shot per thread block.
//
// Checks = detectors (D), variables = error mechanisms (M), edges = nnz(H).  Stands in for the BP
// stage of stimbposd's BP+OSD (reference threshold_lab.py:19,30) -- without OSD,
// as BASELINE.json's north star asks.  Flooding schedule, fp32 messages:
//   check c -> variable v :  alpha * (-1)^{s_c} * prod_{v' != v} sign(m_{v'->c}) * min_{v' != v} |m_{v'->c}|
//   variable v posterior  :  prior_v + sum_c m_{c->v}      (m_{v->c} = posterior - m_{c->v}, formed on the fly)
// The block keeps one message per edge (check -> variable) and one posterior per variable; they
// live in shared memory when the model fits (Steane, small surface codes) and in a per-block
// HBM/L2 scratch otherwise (rotated d=11: 0.4 MB, BB-144: 1.8 MB per shot).
// Check update: one warp per check, lanes stride over the check's CSR row; min1 / min2 / argmin /
// sign parity are combined with warp shuffles, and the same sweep evaluates the parity of the
// current hard decision, so convergence (H.e == s) costs no extra pass.
// Summation order and tie rules equal oracle/bp_oracle.c, so results match it bit for bit.
#include <math.h>
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
};

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
                for (uint32_t k = b + lane; k < e; k += 32) {
                    float t = tot[a.chk_var[k]];
                    float m = t - c2v[k];
                    float av = fabsf(m);
                    sg ^= (m < 0.0f);
                    par ^= (t < 0.0f);
                    if (av < m1) { m2 = m1; m1 = av; arg = k; }
                    else if (av < m2) m2 = av;
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
                for (uint32_t k = b + lane; k < e; k += 32) {
                    float m = tot[a.chk_var[k]] - c2v[k];
                    float mag = (k == garg) ? g2 : g1;
                    int s = sc ^ sg ^ (m < 0.0f);
                    c2v[k] = a.alpha * (s ? -mag : mag);
                }
            }
            __syncthreads();
            if (it > 1 && s_notconv == 0) { converged = 1; used = it - 1; break; }
            if (it == a.max_iter + 1) { used = a.max_iter; break; }
            // ---- variable pass
            for (uint32_t v = tid; v < a.M; v += blockDim.x) {
                float t = a.prior[v];
                for (uint32_t k = a.var_ptr[v]; k < a.var_ptr[v + 1]; ++k) t += c2v[a.var_edge[k]];
                tot[v] = t;
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
    }
}

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
        (rc = up(&dec->d_var_obs, var_obs)))
        return rc;
    GQ_CUDA(cudaMalloc((void **)&dec->d_prior, (size_t)M * 4));
    GQ_CUDA(cudaMemcpy(dec->d_prior, prior.data(), (size_t)M * 4, cudaMemcpyHostToDevice));
    dec->bp_msgs_per_block = ((size_t)nnz + M + 31) & ~(size_t)31;
    size_t bytes = dec->bp_msgs_per_block * 4;
    if (bytes <= 200 * 1024) {
        dec->bp_blocks = dem->num_sms * std::max<int>(1, std::min<int>(8, (int)(200 * 1024 / bytes)));
        cudaFuncSetAttribute(bp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    } else {
        dec->bp_blocks = dem->num_sms * 4;
        GQ_CUDA(cudaMalloc((void **)&dec->d_bp_msgs, bytes * dec->bp_blocks));
    }
    return GQ_OK;
}

void gq_bp_destroy(gq_dec *dec) {
    cudaFree(dec->d_chk_ptr); cudaFree(dec->d_chk_var); cudaFree(dec->d_var_ptr);
    cudaFree(dec->d_var_edge); cudaFree(dec->d_var_obs); cudaFree(dec->d_prior); cudaFree(dec->d_bp_msgs);
}

int gq_bp_decode(gq_dec *dec, const uint32_t *dets_sm, uint64_t nshots, uint32_t *pred_sm,
                 int32_t *weight_out, uint8_t *flags_out, uint32_t *corr_out, int *launches) {
    gq_dem *dem = dec->dem;
    BpArgs a;
    memset(&a, 0, sizeof(a));
    a.dets = dets_sm; a.pred = pred_sm; a.iters_out = weight_out; a.flags_out = flags_out; a.corr_out = corr_out;
    a.nshots = nshots;
    a.D = dem->D; a.M = dem->M; a.DW = dem->DW; a.OW = dem->OW; a.MW = (dem->M + 31) / 32; a.nnz = dec->bp_nnz;
    a.chk_ptr = dec->d_chk_ptr; a.chk_var = dec->d_chk_var; a.var_ptr = dec->d_var_ptr;
    a.var_edge = dec->d_var_edge; a.var_obs = dec->d_var_obs; a.prior = dec->d_prior;
    a.scratch = dec->d_bp_msgs; a.scratch_stride = dec->bp_msgs_per_block;
    a.max_iter = dec->params.bp_max_iter; a.alpha = dec->params.bp_alpha;
    a.use_shared = dec->d_bp_msgs == nullptr;
    size_t smem = a.use_shared ? dec->bp_msgs_per_block * 4 : 0;
    int blocks = (int)std::min<uint64_t>(nshots, (uint64_t)dec->bp_blocks);
    bp_kernel<<<blocks, 256, smem, dem->stream>>>(a);
    GQ_CUDA(cudaGetLastError());
    ++*launches;
    return GQ_OK;
}
