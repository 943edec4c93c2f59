// Per-(point, transformed-dim) spline maps of the coupling cells, forward and backward.
// Conditioner outputs (logits) live in a strided column: z[j * zs], j = 0..K-1.
//   PWLin : nisrep/normalizing_flows/layers/coupling_cells.py:114-141
//   PWQuad: nisrep/normalizing_flows/layers/coupling_cells.py:167-225
// All exponentials are max-shifted (mathematically identical to the reference's plain exp followed
// by normalisation; required in fp32).
// The functions are plain scalar code: with a host compiler (no __CUDACC__) they build as ordinary
// inline functions, which is how tests/test_host_math.py checks them against the oracle without a GPU.
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define NIS_DEV __device__ __forceinline__
#define NIS_DEV_MEMBER __device__ __forceinline__
#else
#include <math.h>
#define NIS_DEV static inline
#define NIS_DEV_MEMBER inline
#endif

#define NIS_QUAD_CLAMP 0.999999f   // float(1 - 1e-6), coupling_cells.py:167

// ---- PWLin --------------------------------------------------------------------------------------
// Overwrites z(j) with e_j = exp(z_j - max).  Returns y; f = bin height (Jacobian factor); k = bin.
// `z` is any accessor: z(j) -> float& (strided column, swizzled shared-memory row, ...).
template <typename Z>
NIS_DEV float pwlin_fwd_z(Z z, int nb, float x, float& f, int& k, float& S_out, float& alpha_out) {
    float m = z(0);
    for (int j = 1; j < nb; ++j) m = fmaxf(m, z(j));
    float S = 0.f;
    for (int j = 0; j < nb; ++j) { float e = expf(z(j) - m); z(j) = e; S += e; }
    float a = x * (float)nb;
    float fl = floorf(a);
    k = (int)fl;
    k = k < 0 ? 0 : (k > nb - 1 ? nb - 1 : k);       // reference: unclamped gather (x==1 raises)
    float alpha = a - (float)k;                      // in [0,1): fraction of the bin
    float C = 0.f;
    for (int j = 0; j < k; ++j) C += z(j);
    float ek = z(k);
    float inv = 1.f / S;
    f = ek * inv * (float)nb;
    S_out = S;
    alpha_out = alpha;
    return (ek * alpha + C) * inv;
}

struct StridedCol {
    float* p; int s;
    NIS_DEV_MEMBER float& operator()(int j) const { return p[j * s]; }
};

NIS_DEV float pwlin_fwd(float* z, int zs, int nb, float x, float& f, int& k, float& S_out, float& alpha_out) {
    StridedCol c{z, zs};
    return pwlin_fwd_z(c, nb, x, f, k, S_out, alpha_out);
}

// Inverse of the PWLin map (the reference lists the inverse as to do, README.md:68-69): given y, the bin is the
// largest k with C_k <= y and x = (k + (y S - sum_{j<k} e_j) / e_k) / nb; f = bin height at x (the inverse's Jacobian
// factor is 1/f).  Same float32 running sum as pwlin_fwd_z, so a round trip lands in the same bin.
NIS_DEV float pwlin_inv(float* z, int zs, int nb, float y, float& f, int& k) {
    float m = z[0];
    for (int j = 1; j < nb; ++j) m = fmaxf(m, z[j * zs]);
    float S = 0.f;
    for (int j = 0; j < nb; ++j) { float e = expf(z[j * zs] - m); z[j * zs] = e; S += e; }
    const float target = y * S;
    float C = 0.f;
    k = 0;
    for (int j = 0; j < nb - 1; ++j) {
        const float nc = C + z[j * zs];
        if (nc <= target) { C = nc; k = j + 1; } else break;
    }
    const float ek = z[k * zs];
    f = ek / S * (float)nb;
    float x = ((float)k + (target - C) / ek) / (float)nb;
    return x < 0.f ? 0.f : (x > 1.f ? 1.f : x);
}

// z holds e_j (after pwlin_fwd).  Overwrites z[j] with dL/dz_j.  Returns dL/dx.
//   gy = dL/dy, gJJ = dL/dJ_out * J_out (= dL/df * f)
NIS_DEV float pwlin_bwd(float* z, int zs, int nb, int k, float S, float alpha, float y,
                                           float f, float gy, float gJJ) {
    float inv = 1.f / S;
    for (int j = 0; j < nb; ++j) {
        float p = z[j * zs] * inv;
        float ind = j < k ? 1.f : (j == k ? alpha : 0.f);
        float g = gy * p * (ind - y) + gJJ * ((j == k ? 1.f : 0.f) - p);
        z[j * zs] = g;
    }
    return gy * f;
}

// ---- PWQuad -------------------------------------------------------------------------------------
struct QuadCtx {
    int k;
    float xb, Sw, A /*area with normalised widths, raw heights*/, alpha, Wk, Vk, Vk1, y, f;
    bool clamped;
};

// z[0..nb] raw heights, z[nb+1..2nb] raw widths.  Overwrites with v_j = exp(. - max) and
// w_j = exp(. - max).  Fills ctx.
// FAST: exponentials as ex2.approx(x log2 e) (__expf; arguments are <= 0 after the max shift) instead of the accurate expf --
// the streamed-weights kernels (flow_wide.cu final pass, flow_bwd_wide.cu head), like the resident-weights tensor-core kernels.
template <bool FAST = false>
NIS_DEV float nis_spline_exp(float x) {
#ifdef __CUDA_ARCH__
    if (FAST) return __expf(x);
#endif
    return expf(x);
}

template <bool FAST = false>
NIS_DEV void pwquad_fwd(float* z, int zs, int nb, float x, QuadCtx& c) {
    float* zv = z;
    float* zw = z + (nb + 1) * zs;
    // (FAST also unrolls the read-only loops by 8: with one thread per point and one warp per scheduler the streamed-weights
    //  kernels otherwise wait one shared-memory latency per bin)
    float mv = zv[0], mw = zw[0];
#pragma unroll(FAST ? 8 : 1)
    for (int j = 1; j <= nb; ++j) mv = fmaxf(mv, zv[j * zs]);
#pragma unroll(FAST ? 8 : 1)
    for (int j = 1; j < nb; ++j) mw = fmaxf(mw, zw[j * zs]);
    // The cumulative sums that locate x inside its bin are kept in float64: alpha = (x - E_k)/W_k
    // amplifies their rounding by 1/W_k, and a float64 add per bin is free next to the conditioner.
    double Sw = 0.0;
#pragma unroll(FAST ? 8 : 1)
    for (int j = 0; j < nb; ++j) { float e = nis_spline_exp<FAST>(zw[j * zs] - mw); zw[j * zs] = e; Sw += (double)e; }
    float vprev = nis_spline_exp<FAST>(zv[0] - mv);
    zv[0] = vprev;
    double Araw = 0.0;      // sum (v_j + v_{j+1})/2 * w_j  (unnormalised widths)
#pragma unroll(FAST ? 8 : 1)
    for (int j = 0; j < nb; ++j) {
        float vn = nis_spline_exp<FAST>(zv[(j + 1) * zs] - mv);
        zv[(j + 1) * zs] = vn;
        Araw += 0.5 * ((double)vprev + (double)vn) * (double)zw[j * zs];
        vprev = vn;
    }
    c.clamped = x > NIS_QUAD_CLAMP;
    float xb = c.clamped ? NIS_QUAD_CLAMP : x;
    const double target = (double)xb * Sw;
    // bin = number of right edges E_1..E_nb <= xb  (coupling_cells.py:199-202)
    int k = 0;
    double cw = 0.0, ca = 0.0;    // sum_{j<k} w_j, sum_{j<k} trapezoid_j (raw)
    for (int j = 0; j < nb - 1; ++j) {
        const double nw = cw + (double)zw[j * zs];
        if (nw <= target) {
            ca += 0.5 * ((double)zv[j * zs] + (double)zv[(j + 1) * zs]) * (double)zw[j * zs];
            cw = nw;
            k = j + 1;
        } else break;
    }
    const float wk = zw[k * zs];
    const float invA = (float)(Sw / Araw);       // 1/A, A = Araw/Sw
    const float alpha = (float)((target - cw) / (double)wk);
    const float Vk = zv[k * zs] * invA, Vk1 = zv[(k + 1) * zs] * invA;
    const float Wk = (float)((double)wk / Sw);
    c.k = k; c.xb = xb; c.Sw = (float)Sw; c.A = (float)(Araw / Sw); c.alpha = alpha; c.Wk = Wk; c.Vk = Vk; c.Vk1 = Vk1;
    c.y = alpha * alpha * 0.5f * (Vk1 - Vk) * Wk + alpha * Vk * Wk + (float)(ca / Araw);
    c.f = Vk + alpha * (Vk1 - Vk);
}

// Inverse of the PWQuad map: the bin is the largest k with S_k <= y (S = cdf at the edges); inside it
// y - S_k = alpha V_k W_k + alpha^2 (V_{k+1} - V_k) W_k / 2 is solved for alpha in its cancellation-free form.
// Returns x; f = density at x (the inverse's Jacobian factor is 1/f).  Float64 sums as in pwquad_fwd.
NIS_DEV float pwquad_inv(float* z, int zs, int nb, float y, float& f, int& kout) {
    float* zv = z;
    float* zw = z + (nb + 1) * zs;
    float mv = zv[0], mw = zw[0];
    for (int j = 1; j <= nb; ++j) mv = fmaxf(mv, zv[j * zs]);
    for (int j = 1; j < nb; ++j) mw = fmaxf(mw, zw[j * zs]);
    double Sw = 0.0;
    for (int j = 0; j < nb; ++j) { float e = expf(zw[j * zs] - mw); zw[j * zs] = e; Sw += (double)e; }
    float vprev = expf(zv[0] - mv);
    zv[0] = vprev;
    double Araw = 0.0;
    for (int j = 0; j < nb; ++j) {
        float vn = expf(zv[(j + 1) * zs] - mv);
        zv[(j + 1) * zs] = vn;
        Araw += 0.5 * ((double)vprev + (double)vn) * (double)zw[j * zs];
        vprev = vn;
    }
    const double target = (double)y * Araw;
    int k = 0;
    double cw = 0.0, ca = 0.0;
    for (int j = 0; j < nb - 1; ++j) {
        const double na = ca + 0.5 * ((double)zv[j * zs] + (double)zv[(j + 1) * zs]) * (double)zw[j * zs];
        if (na <= target) { ca = na; cw += (double)zw[j * zs]; k = j + 1; } else break;
    }
    const double wk = (double)zw[k * zs], vk = (double)zv[k * zs], vk1 = (double)zv[(k + 1) * zs];
    double c = target - ca;
    c = c > 0.0 ? c : 0.0;
    const double a = 0.5 * (vk1 - vk) * wk, b = vk * wk;
    double disc = b * b + 4.0 * a * c;
    disc = disc > 0.0 ? disc : 0.0;
    double alpha = 2.0 * c / (b + sqrt(disc));
    alpha = alpha < 0.0 ? 0.0 : (alpha > 1.0 ? 1.0 : alpha);
    f = (float)((vk + alpha * (vk1 - vk)) * (Sw / Araw));
    kout = k;
    const double x = (cw + alpha * wk) / Sw;
    return (float)(x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x));
}

// z holds v_j / w_j (after pwquad_fwd).  Overwrites with dL/dz.  Returns dL/dx.
//   gy = dL/dy, gf = dL/df
template <bool FAST = false>
NIS_DEV float pwquad_bwd(float* z, int zs, int nb, const QuadCtx& c, float gy, float gf) {
    float* zv = z;
    float* zw = z + (nb + 1) * zs;
    const int k = c.k;
    const float invA = 1.f / c.A, invSw = 1.f / c.Sw;
    const float D = c.Vk1 - c.Vk;
    const float galpha = gy * c.f * c.Wk + gf * D;
    // pass 1: R = sum_j GV_j V_j ; Q = sum_j GWdirect_j W_j ; Abar = sum_j (V_j+V_{j+1})/2 W_j = 1
    float R = 0.f, Q = 0.f;
#pragma unroll(FAST ? 8 : 1)
    for (int j = 0; j <= nb; ++j) {
        float Vj = zv[j * zs] * invA;
        float gv = 0.f;
        if (j < k) gv += 0.5f * (zw[j * zs] * invSw);
        if (j >= 1 && j <= k) gv += 0.5f * (zw[(j - 1) * zs] * invSw);
        gv *= gy;
        if (j == k) gv += gy * (c.alpha - 0.5f * c.alpha * c.alpha) * c.Wk + gf * (1.f - c.alpha);
        if (j == k + 1) gv += gy * 0.5f * c.alpha * c.alpha * c.Wk + gf * c.alpha;
        R += gv * Vj;
    }
#pragma unroll(FAST ? 8 : 1)
    for (int j = 0; j <= k; ++j) {
        float Wj = zw[j * zs] * invSw;
        float trap = 0.5f * (zv[j * zs] + zv[(j + 1) * zs]) * invA;
        float gw = j < k ? (-galpha / c.Wk + gy * trap) : (-galpha * c.alpha / c.Wk + gy * (0.5f * c.alpha * c.alpha * D + c.alpha * c.Vk));
        Q += gw * Wj;
    }
    // via A: GW_j += -R * trap_j ; sum_j (-R trap_j) W_j = -R
    Q -= R;
    // pass 2, descending j: at step j the widths w_j, w_{j-1} and heights v_j (memory), v_{j+1}
    // (register) are still the forward values; bin j and vertex j are overwritten with gradients.
    float v_hi_old = zv[nb * zs];          // v_{nb}
    {
        // j = nb height gradient
        float Vj = v_hi_old * invA;
        float gv = 0.f;                   // vertex nb is never inside the summed trapezoids (k <= nb-1)
        if (nb == k + 1) gv += gy * 0.5f * c.alpha * c.alpha * c.Wk + gf * c.alpha;
        float dz = Vj * (gv - R * 0.5f * (zw[(nb - 1) * zs] * invSw));
        zv[nb * zs] = dz;
    }
    for (int j = nb - 1; j >= 0; --j) {
        float v_old = zv[j * zs];
        float w_old = zw[j * zs];
        float Wj = w_old * invSw;
        float Wjm1 = j >= 1 ? zw[(j - 1) * zs] * invSw : 0.f;
        // width gradient of bin j (needs v_j(old), v_{j+1}(old))
        float trap = 0.5f * (v_old + v_hi_old) * invA;
        float gw = -R * trap;
        if (j < k) gw += -galpha / c.Wk + gy * trap;
        else if (j == k) gw += -galpha * c.alpha / c.Wk + gy * (0.5f * c.alpha * c.alpha * D + c.alpha * c.Vk);
        zw[j * zs] = Wj * (gw - Q);
        // height gradient of vertex j (needs W_j, W_{j-1} (old; bin j-1 not yet overwritten))
        float Vj = v_old * invA;
        float gv = 0.f;
        if (j < k) gv += 0.5f * Wj;
        if (j >= 1 && j <= k) gv += 0.5f * Wjm1;
        gv *= gy;
        if (j == k) gv += gy * (c.alpha - 0.5f * c.alpha * c.alpha) * c.Wk + gf * (1.f - c.alpha);
        if (j == k + 1) gv += gy * 0.5f * c.alpha * c.alpha * c.Wk + gf * c.alpha;
        zv[j * zs] = Vj * (gv - R * 0.5f * (Wj + Wjm1));
        v_hi_old = v_old;
    }
    return c.clamped ? 0.f : (gy * c.f + gf * D / c.Wk);
}
