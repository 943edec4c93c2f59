// Register-resident splines shared by the tensor-core forward kernels (flow_tc.cu, flow_tc_h.cu).
#pragma once
#include "spline.cuh"

// PWQuad with 32 bins on a register-resident logit vector z[0..64] (33 vertex heights, 32 widths): the
// reference's map (coupling_cells.py:167-225, see spline.cuh::pwquad_fwd) with static indexing only — the bin
// is found by counting edges, the per-bin quantities by predicated accumulation.
__device__ __forceinline__ void pwquad32_regs(const float* z, float x, float& y, float& f, int& kbin) {
    float mv = z[0], mw = z[33];
#pragma unroll
    for (int j = 1; j <= 32; ++j) mv = fmaxf(mv, z[j]);
#pragma unroll
    for (int j = 1; j < 32; ++j) mw = fmaxf(mw, z[33 + j]);
    float w[32], v[33];
    double Sw = 0.0;
#pragma unroll
    for (int j = 0; j < 32; ++j) { w[j] = expf(z[33 + j] - mw); Sw += (double)w[j]; }
#pragma unroll
    for (int j = 0; j <= 32; ++j) v[j] = expf(z[j] - mv);
    double Araw = 0.0;
#pragma unroll
    for (int j = 0; j < 32; ++j) Araw += 0.5 * ((double)v[j] + (double)v[j + 1]) * (double)w[j];
    const float xb = x > NIS_QUAD_CLAMP ? NIS_QUAD_CLAMP : x;
    const double target = (double)xb * Sw;
    int k = 0;
    double cum = 0.0;
#pragma unroll
    for (int j = 0; j < 31; ++j) { cum += (double)w[j]; k += cum <= target ? 1 : 0; }
    double cw = 0.0, ca = 0.0;
    float wk = 0.f, vk = 0.f, vk1 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const bool below = j < k;
        cw += below ? (double)w[j] : 0.0;
        ca += below ? 0.5 * ((double)v[j] + (double)v[j + 1]) * (double)w[j] : 0.0;
        wk = j == k ? w[j] : wk;
        vk = j == k ? v[j] : vk;
        vk1 = j == k ? v[j + 1] : vk1;
    }
    const float invA = (float)(Sw / Araw);
    const float alpha = (float)((target - cw) / (double)wk);
    const float Vk = vk * invA, Vk1 = vk1 * invA;
    const float Wk = (float)((double)wk / Sw);
    y = alpha * alpha * 0.5f * (Vk1 - Vk) * Wk + alpha * Vk * Wk + (float)(ca / Araw);
    f = Vk + alpha * (Vk1 - Vk);
    kbin = k;
}

