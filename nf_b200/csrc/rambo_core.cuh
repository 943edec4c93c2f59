// RAMBO-on-diet, one event: intermediate masses, massive reweighting, sequential two-body decays with
// boosts, cuts.  float64 throughout (the reference is float64 and the cut masks must match it).
//
// Restates the pdf-inactive path of nisrep/PhaseSpace/flat_phase_space_generator.py:139-308 with its
// helpers (:81-113 weights / rho, :313-359 root of the mass polynomial, :363-441 intermediates and
// beams) and nisrep/PhaseSpace/utils.py:5-81 (set_square / boost), :151-187 (pseudo-rapidity, deltaR).
// Plain scalar code: compiles for the device (kernel in rambo.cu) and, without __CUDACC__, for the host
// (tests/test_host_math.py checks it against the oracle without a GPU).
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define NIS_DEV __device__ __forceinline__
#else
#include <math.h>
#define NIS_DEV static inline
#endif
#include <stdint.h>
#include "../../include/nis_b200.h"

#define NIS_TWO_PI 6.283185307179586
#define NIS_PI 3.141592653589793
#define NIS_SQRT_EPS 1.4901161193847656e-08   /* np.finfo(float).eps**0.5, utils.py:151 */
#define NIS_HUGE 1.7976931348623157e308       /* np.finfo(float).max */

struct RamboConst {
    int n;
    double m[NIS_MAX_FINAL];      // final masses
    double msum[NIS_MAX_FINAL];   // msum[j] = sum_{i>=j} m_i
    double K0;                    // E_cm - sum m
    double wconst;                // flat volume * (K0/M0)^(2n-4) / (2 E_cm^2)
    double beam[2][4];
    double pT_min, dR_min, rap_max;
};

// Host: fold the descriptor into per-launch constants.
static inline int rambo_fill_const(const NisRamboDesc* d, RamboConst* C) {
    const int n = d->n_final;
    if (n < 2 || n > NIS_MAX_FINAL) return NIS_EINVAL;
    C->n = n;
    double tot = 0.0;
    for (int j = n - 1; j >= 0; --j) { tot += d->final_masses[j]; C->m[j] = d->final_masses[j]; C->msum[j] = tot; }
    const double E = d->E_cm;
    if (!(E > tot)) return NIS_EINVAL;
    C->K0 = E - tot;
    // get_flatWeights, flat_phase_space_generator.py:81-97
    double fact1 = 1.0, fact2 = 1.0;
    for (int i = 2; i <= n - 1; ++i) fact1 *= i;
    for (int i = 2; i <= n - 2; ++i) fact2 *= i;
    const double flat = pow(2 * NIS_PI, 4 - 3 * n) * pow(NIS_PI / 2.0, n - 1) * (pow(E * E, n - 2) / (fact1 * fact2));
    C->wconst = flat * pow(C->K0 / E, 2 * n - 4) / (2.0 * E * E);   // :403, :307-308 (M_0 = E_cm)
    const double m1 = d->initial_masses[0], m2 = d->initial_masses[1];
    if (m1 == 0.0 || m2 == 0.0) {                                    // :415-419
        const double b[2][4] = {{E / 2.0, 0.0, 0.0, E / 2.0}, {E / 2.0, 0.0, 0.0, -E / 2.0}};
        for (int i = 0; i < 8; ++i) C->beam[i / 4][i % 4] = b[i / 4][i % 4];
    } else {                                                         // :425-433
        const double M1 = m1 * m1, M2 = m2 * m2;
        const double E1 = (E * E + M1 - M2) / E, E2 = (E * E - M1 + M2) / E;
        const double Z = sqrt(E * E * E * E - 2 * E * E * M1 - 2 * E * E * M2 + M1 * M1 - 2 * M1 * M2 + M2 * M2) / E;
        const double b[2][4] = {{E1 / 2.0, 0.0, 0.0, Z / 2.0}, {E2 / 2.0, 0.0, 0.0, -Z / 2.0}};
        for (int i = 0; i < 8; ++i) C->beam[i / 4][i % 4] = b[i / 4][i % 4];
    }
    C->pT_min = d->pT_mincut; C->dR_min = d->delR_mincut; C->rap_max = d->rap_maxcut;
    return NIS_OK;
}


// Root in [0,1] of  r = (e+1) u^e - e u^(e+1)  (flat_phase_space_generator.py:101-103,313-359; the
// reference bisects 120-180 levels).  Safeguarded Newton: the bracket [lo,hi] always contains the
// root, a Newton step leaving it is replaced by bisection.
template <int E>
NIS_DEV double rambo_root(double r) {
    if (E == 1) return r / (1.0 + sqrt(1.0 - r));             // u = 1 - sqrt(1-r), stable form
    if (r <= 0.0) return 0.0;
    if (r >= 1.0) return 1.0;
    const double e = (double)E;
    double lo = 0.0, hi = 1.0;
    // start: small-r asymptote u ~ (r/(e+1))^(1/e) (a lower bound of the root) or, past the inflection
    // point (e-1)/e, the large-r asymptote 1 - sqrt(2(1-r)/(e(e+1))) (an upper bound)
    const double ustar = (e - 1.0) / e;
    double us = 1.0;
    for (int i = 0; i < E; ++i) us *= ustar;
    const double rstar = us * ((e + 1.0) - e * ustar);
    double x;
    if (r < rstar) { x = pow(r / (e + 1.0), 1.0 / e); lo = x; hi = ustar; }
    else { x = 1.0 - sqrt(2.0 * (1.0 - r) / (e * (e + 1.0))); hi = x; lo = ustar; if (x < ustar) { x = ustar; hi = 1.0; } }
    for (int it = 0; it < 64; ++it) {
        double xe1 = 1.0;                                     // x^(e-1)
        for (int i = 0; i < E - 1; ++i) xe1 *= x;
        const double g = xe1 * x * ((e + 1.0) - e * x) - r;
        const double dg = e * (e + 1.0) * xe1 * (1.0 - x);
        if (g > 0.0) hi = x; else lo = x;
        double xn = x - g / dg;
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        const double dx = fabs(xn - x);
        x = xn;
        if (dx <= 4.5e-16 * x) break;
    }
    return x;
}

NIS_DEV double rambo_root_dyn(int e, double r) {
    switch (e) {
        case 1: return rambo_root<1>(r);
        case 2: return rambo_root<2>(r);
        case 3: return rambo_root<3>(r);
        case 4: return rambo_root<4>(r);
        case 5: return rambo_root<5>(r);
        default: return rambo_root<6>(r);
    }
}

NIS_DEV double rambo_rho(double M, double N, double m) {     // :107-113
    const double M2 = M * M;
    return sqrt((M2 - (N + m) * (N + m)) * (M2 - (N - m) * (N - m))) / (8.0 * M2);
}

NIS_DEV double rambo_pseudorap(double px, double py, double pz) {   // utils.py:151-157
    const double pt = sqrt(px * px + py * py);
    if (pt < NIS_SQRT_EPS && fabs(pz) < NIS_SQRT_EPS) return NIS_HUGE;
    const double th = atan2(pt, pz);
    return -log(tan(th / 2.0));
}

// One event.  r: 3n-4 uniforms at stride rs.  mom: (n+2)*4 doubles at stride ms, or null.
template <int N>
NIS_DEV void rambo_event(const RamboConst& C, const double* r, int rs, double* mom, int ms, double& weight,
                         uint8_t& pass) {
    double K[N > 1 ? N - 1 : 1], M[N];
    K[0] = C.K0;
#pragma unroll
    for (int j = 0; j < N - 2; ++j) {                          // :363-370
        const double u = rambo_root_dyn(N - 2 - j, r[j * rs]);
        K[j + 1] = sqrt(u * (K[j] * K[j]));
    }
#pragma unroll
    for (int j = 0; j < N - 1; ++j) M[j] = K[j] + C.msum[j];   // :391-392
    M[N - 1] = C.m[N - 1];
    double w = C.wconst * 8.0 * rambo_rho(M[N - 2], C.m[N - 1], C.m[N - 2]);   // :394-397
#pragma unroll
    for (int j = 0; j < N - 2; ++j)                            // :400-401
        w *= rambo_rho(M[j], M[j + 1], C.m[j]) / rambo_rho(K[j], K[j + 1], 0.0) * (M[j + 1] / K[j + 1]);

    double Q0 = M[0], Q1 = 0.0, Q2 = 0.0, Q3 = 0.0;
    double fx[N], fy[N], fz[N];
    double E_last = 0.0;
#pragma unroll
    for (int j = 0; j < N - 1; ++j) {
        const double q = 4.0 * M[j] * rambo_rho(M[j], M[j + 1], C.m[j]);   // :228
        const double ct = 2.0 * r[(N - 2 + 2 * j) * rs] - 1.0;             // :233-243
        const double st = sqrt(1.0 - ct * ct);
        const double phi = NIS_TWO_PI * r[(N - 1 + 2 * j) * rs];
        const double cp = cos(phi);
        const double sp0 = sqrt(1.0 - cp * cp);
        const double sp = phi > NIS_PI ? -sp0 : sp0;
        double p1 = q * st * cp, p2 = q * st * sp, p3 = q * ct;
        const double m2 = C.m[j] * C.m[j];
        const double p0 = sqrt(p1 * p1 + p2 * p2 + p3 * p3 + m2);          // set_square_t utils.py:5-19
        const double b1 = Q1 / Q0, b2_ = Q2 / Q0, b3 = Q3 / Q0;            // boostVector_t :31-36
        const double bb = b1 * b1 + b2_ * b2_ + b3 * b3;                   // boost_t :58-81
        const double gamma = 1.0 / sqrt(1.0 - bb);
        const double bp = p1 * b1 + p2 * b2_ + p3 * b3;
        const double gamma2 = bb > 0.0 ? (gamma - 1.0) / bb : 0.0;
        const double fac = gamma2 * bp + gamma * p0;
        p1 += fac * b1; p2 += fac * b2_; p3 += fac * b3;
        const double e = sqrt(p1 * p1 + p2 * p2 + p3 * p3 + m2);           // :265
        fx[j] = p1; fy[j] = p2; fz[j] = p3;
        if (mom) { double* o = mom + (2 + j) * 4 * ms; o[0] = e; o[ms] = p1; o[2 * ms] = p2; o[3 * ms] = p3; }
        Q1 -= p1; Q2 -= p2; Q3 -= p3;                                       // :271-275
        Q0 = sqrt(Q1 * Q1 + Q2 * Q2 + Q3 * Q3 + M[j + 1] * M[j + 1]);
        E_last = Q0;
    }
    fx[N - 1] = Q1; fy[N - 1] = Q2; fz[N - 1] = Q3;                         // :278
    if (mom) {
        double* o = mom + (N + 1) * 4 * ms;
        o[0] = E_last; o[ms] = Q1; o[2 * ms] = Q2; o[3 * ms] = Q3;
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int i = 0; i < 4; ++i) mom[(b * 4 + i) * ms] = C.beam[b][i];
    }
    // ---- cuts (:285-301); x1 = x2 = 1 so the lab frame is the CM frame -------------------------
    bool ok = true;
    double ptmin = NIS_HUGE;
#pragma unroll
    for (int j = 0; j < N; ++j) ptmin = fmin(ptmin, sqrt(fx[j] * fx[j] + fy[j] * fy[j]));
    if (ptmin < C.pT_min) ok = false;
    const bool need_eta = C.rap_max > 0.0 || C.dR_min > 0.0;
    if (need_eta) {
        double eta[N], pt[N];
        double etamax = -NIS_HUGE;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            eta[j] = rambo_pseudorap(fx[j], fy[j], fz[j]);
            pt[j] = sqrt(fx[j] * fx[j] + fy[j] * fy[j]);
            etamax = fmax(etamax, eta[j]);
        }
        if (C.rap_max > 0.0 && C.rap_max < fabs(etamax)) ok = false;          // |max eta|, not max |eta|
        if (C.dR_min > 0.0) {
#pragma unroll
            for (int i = 1; i < N; ++i)
#pragma unroll
                for (int j = 0; j < i; ++j) {
                    const double deta = eta[i] - eta[j];
                    double dphi;
                    if (pt[i] == 0.0 || pt[j] == 0.0) dphi = NIS_HUGE;      // utils.py:170-180
                    else {
                        double t = (fx[i] * fx[j] + fy[i] * fy[j]) / (pt[i] * pt[j]);
                        if (fabs(t) > 1.0) t = t / fabs(t);
                        dphi = acos(t);
                    }
                    const double dR = sqrt(deta * deta + dphi * dphi);
                    if (fabs(dR) < C.dR_min) ok = false;
                }
        }
    }
    pass = ok ? 1 : 0;
    weight = ok ? w : 0.0;
}
