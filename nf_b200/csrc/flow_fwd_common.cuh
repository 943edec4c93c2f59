// Pieces shared by the forward kernels (flow_fwd.cu: shape-generic; flow_tiled.cu: register-tiled).
#pragma once
#include "common.cuh"

struct FwdArgs {
    const void* in; int in_dtype; int in_cols;   // external input rows (first cell, !from_state)
    const float* state_in;                       // fp32 [B][d+1] (from_state)
    float* state_out;                            // fp32 [B][d+1] or null
    void* out; int out_dtype;                    // external output (to_out)
    int from_state, to_out;
    float* saved;                                // [n_cells+1][B][d+1] or null
    int32_t* bins;
    const float* params; float* wpack; float* bn_running; float* bn_saved;
    double* partials; unsigned* counter;
    long long B; int c_begin, c_end, stats_layer;
    const float* zin;                            // tiled train path: pre-BN activations of the layer below,
    float* zout;                                 // tile-blocked [tile][64][M] (written by a statistics pass)
    int no_stats;                                // layer pass that only stores its activations (eval-mode split cell)
    float* z1out;                          // wide kernel, pass from the state: also store z_1 (backward recompute) or null
    float* scratch_state;                  // cooperative small-batch kernel: fp32 [B][d+1] state between cells when `saved` is null
    int zin_layer;                         // fp16-split kernel: which layer's pre-activations `zin` holds (0: the default,
                                           // stats_layer - 1 for a layer pass, depth for the final pass)
    int inverse;                           // nis_flow_inverse: cells backwards, inverse splines, J divided (shape-generic kernel only)
};


__device__ __forceinline__ float load_io(const void* p, int dtype, long long idx) {
    return dtype == NIS_F64 ? (float)reinterpret_cast<const double*>(p)[idx] : reinterpret_cast<const float*>(p)[idx];
}
__device__ __forceinline__ void store_io(void* p, int dtype, long long idx, float v) {
    if (dtype == NIS_F64) reinterpret_cast<double*>(p)[idx] = (double)v; else reinterpret_cast<float*>(p)[idx] = v;
}


// Train-mode BN statistics: every CTA of a statistics pass hands in its float64 per-feature sums
// (sacc[0..maxW) = sum z, sacc[maxW..2maxW) = sum z^2, in shared memory); the last CTA to arrive adds
// the per-CTA partials in index order (deterministic) and folds them into the layer's scale/shift
// (wpack), the saved batch statistics (backward) and the running statistics (torch BatchNorm1d
// semantics: biased variance for normalisation, unbiased for the running estimate, momentum 0.1).
// `scratch` (optional, 2 * NT doubles of shared memory): the last CTA then adds the per-CTA partials with all its threads
// (NT / Wp slices per feature, combined in slice order - still a fixed order) instead of one thread per feature walking
// all gridDim.x partials: 148 x 2 dependent-latency loads per feature were ~25 us at the tail of every statistics pass.
__device__ __forceinline__ void bn_stats_finalize(const DevFlow& F, const FwdArgs& A, const double* sacc, int NT,
                                                  double* scratch = nullptr) {
    const int tid = threadIdx.x;
    const int maxW = F.maxW;
    __syncthreads();
    const int c = A.c_begin, l = A.stats_layer;
    const DevCell& q = F.cells[c];
    const int W = F.W(c, l), Wp = F.Wp(c, l);
    double* mine = A.partials + (size_t)blockIdx.x * 2 * maxW;
    for (int i = tid; i < 2 * maxW; i += NT) mine[i] = sacc[i];
    __threadfence();
    __syncthreads();
    __shared__ bool s_last;
    if (tid == 0) s_last = atomicAdd(A.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* p = A.params + q.param_off + F.p_bn_gamma(c, l);
    float* aff = A.wpack + q.pk_off + q.aff_off[l];
    const int slices = scratch ? NT / Wp : 0;
    if (slices > 1) {
        const int sl = tid / Wp, j = tid - sl * Wp;
        if (sl < slices) {
            double s = 0.0, s2 = 0.0;
            for (unsigned b = sl; b < gridDim.x; b += slices) {
                s += __ldcg(A.partials + (size_t)b * 2 * maxW + j);
                s2 += __ldcg(A.partials + (size_t)b * 2 * maxW + maxW + j);
            }
            scratch[sl * Wp + j] = s;
            scratch[NT + sl * Wp + j] = s2;
        }
        __syncthreads();
    }
    for (int j = tid; j < Wp; j += NT) {
        float sc = 0.f, sh = 0.f;
        if (j < W) {
            double s = 0.0, s2 = 0.0;
            if (slices > 1) {
                for (int sl = 0; sl < slices; ++sl) { s += scratch[sl * Wp + j]; s2 += scratch[NT + sl * Wp + j]; }
            } else {
                for (unsigned b = 0; b < gridDim.x; ++b) {
                    s += __ldcg(A.partials + (size_t)b * 2 * maxW + j);
                    s2 += __ldcg(A.partials + (size_t)b * 2 * maxW + maxW + j);
                }
            }
            const double n = (double)A.B;
            const double mean = s / n;
            double var = s2 / n - mean * mean;
            var = var > 0.0 ? var : 0.0;
            const double invstd = 1.0 / sqrt(var + (double)F.eps);
            sc = (float)((double)p[j] * invstd);
            sh = (float)((double)p[W + j] - mean * (double)p[j] * invstd);
            if (A.bn_saved) {
                A.bn_saved[q.sv_off + l * 2 * maxW + j] = (float)mean;
                A.bn_saved[q.sv_off + l * 2 * maxW + maxW + j] = (float)invstd;
            }
            if (A.bn_running) {
                float* rs = A.bn_running + q.bn_off + F.r_mean(c, l);
                const double m = (double)F.momentum;
                const double unb = A.B > 1 ? var * n / (n - 1.0) : var;
                rs[j] = (float)((1.0 - m) * (double)rs[j] + m * mean);
                rs[W + j] = (float)((1.0 - m) * (double)rs[W + j] + m * unb);
            }
        }
        aff[j] = sc;
        aff[Wp + j] = sh;
    }
    if (tid == 0) *A.counter = 0u;
}
