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

#ifdef __CUDACC__
#define nis_fdiv(a, b) __fdividef((a), (b))       /* 2-ulp float division: only steers a Newton iterate */
#define nis_fpow(a, b) __powf((a), (b))           /* ex2.approx(b * lg2.approx(a)) */
#define nis_fsqrt(a) __fsqrt_rn(a)
#else
#define nis_fdiv(a, b) ((a) / (b))
#define nis_fpow(a, b) powf((a), (b))
#define nis_fsqrt(a) sqrtf(a)
#endif
#ifndef __CUDACC__
// host build (tests only): glibc has no sincospi
static inline void sincospi(double x, double* s, double* c) { *s = sin(3.141592653589793 * x); *c = cos(3.141592653589793 * x); }
#endif

// Float64 division / square root without the IEEE corner-case sequences: the hardware's 20-bit
// rcp / rsqrt seed + ONE Newton step + one residual correction, which is itself a refinement step (the error after it is
// the square of the error before it: 2^-20 -> 2^-40 -> 2^-80).  The IEEE sequences cost ~35 / ~30 instructions (a third of
// them integer exponent handling); these cost 6 / 7.  tools/{div,sqrt}_probe (built from /tmp sources):
// correctly rounded on 4.2e6 random arguments each.
#ifdef __CUDACC__
NIS_DEV double nis_div(double a, double b) {
    // one Newton step on the reciprocal (2^-20 -> 2^-40) is enough here: the residual correction below is itself a
    // refinement step, q' = q + (a - b q) y = (a / b)(1 - delta^2)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    const double e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    const double q = a * y;
    return fma(fma(-b, q, a), y, q);
}
NIS_DEV double nis_sqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
    y = y * fma(-h * y, y, 1.5);                    // seed ~2^-21 -> ~2^-41; the residual correction below squares the
    const double s = x * y;                         // error once more (checked: 4.2e6 random arguments, all correctly rounded)
    const double rs_ = fma(fma(-s, s, x), 0.5 * y, s);
    return x > 0.0 ? rs_ : 0.0;
}
#else
static inline double nis_div(double a, double b) { return a / b; }
static inline double nis_sqrt(double x) { return x > 0.0 ? sqrt(x) : 0.0; }
#endif

// sin and cos of 2 pi r for r in [0, 1] (the azimuth of a decay, :236-243): quadrant k = round(4 r), a = 2 pi (r - k/4)
// in [-pi/4, pi/4] (the subtraction is exact), Taylor polynomials to a^15 / a^16 (remainders 5e-17 / 2e-18), quadrant
// fix-up.  ~40 instructions; the library sincospi, which also handles arbitrary arguments, was 8.7 % of the kernel.
NIS_DEV void rambo_sincos2pi(double r, double* s, double* c) {
    const double k = rint(4.0 * r);
    const double a = (r - 0.25 * k) * 6.283185307179586;
    const double a2 = a * a;
    double sp = -7.6471637318198164759e-13;                    // -1/15!
    sp = fma(sp, a2, 1.6059043836821614599e-10);               //  1/13!
    sp = fma(sp, a2, -2.5052108385441718775e-08);              // -1/11!
    sp = fma(sp, a2, 2.7557319223985890653e-06);               //  1/9!
    sp = fma(sp, a2, -1.9841269841269841270e-04);              // -1/7!
    sp = fma(sp, a2, 8.3333333333333333333e-03);               //  1/5!
    sp = fma(sp, a2, -1.6666666666666666667e-01);              // -1/3!
    sp = fma(sp * a2, a, a);
    double cp = 4.7794773323873852974e-14;                     //  1/16!
    cp = fma(cp, a2, -1.1470745597729724714e-11);              // -1/14!
    cp = fma(cp, a2, 2.0876756987868098979e-09);               //  1/12!
    cp = fma(cp, a2, -2.7557319223985890653e-07);              // -1/10!
    cp = fma(cp, a2, 2.4801587301587301587e-05);               //  1/8!
    cp = fma(cp, a2, -1.3888888888888888889e-03);              // -1/6!
    cp = fma(cp, a2, 4.1666666666666666667e-02);               //  1/4!
    cp = fma(cp, a2, -0.5);
    cp = fma(cp, a2, 1.0);
    const int q = (int)k & 3;
    const double s0 = (q & 1) ? cp : sp, c0 = (q & 1) ? sp : cp;
    *s = (q & 2) ? -s0 : s0;
    *c = ((q + 1) & 2) ? -c0 : c0;
}

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
    double e2_rap, e2_rap_inv;    // exp(+-2 rap_max)
    double e2_dR;                 // exp(2 dR_min)
    double cos_dR;                // cos(min(dR_min, pi))
    int dR_ge_pi;                 // dR_min >= pi: the d-phi quick reject never fires
    double u_one[NIS_MAX_FINAL];  // u_one[e]: what the reference's lattice bisection returns for r == 1
    float rstar[NIS_MAX_FINAL];   // rstar[e]: rambo_rstar(e)
    // parton-density mode (flat_phase_space_generator.py:157-187): per-event partonic energy
    int pdf_active, tau_mode;
    double E_coll, tau_min, x_cut;
    double flat_norm;             // (2 pi)^(4-3n) (pi/2)^(n-1) / ((n-1)! (n-2)!)
    double m_in2[2];              // squared beam masses (0, 0 when either beam is massless, :415)
    const double* pdf_grid[2];
    int pdf_nodes;
    double pdf_lnx_lo, pdf_inv_h;
};

// The reference's root finder (flat_phase_space_generator.py:333-348) never returns u = 0 or u = 1:
//   * r == 0 makes its batch-wide error infinite at the first check, so the search stops after 60 levels on
//     the lowest lattice point u = 2^-60 (NIS_U_FLOOR; the same happens for any r whose root is below it);
//   * r == 1: the map value rounds to 1.0 once 1 - u < ~4e-9 and `v <= map(u)` then sends the search left,
//     so it settles on 1 - 2^-27/e.  Every operation of that search is an exact or correctly rounded
//     float64 operation, so it is replayed here on the host (independent of the level count >= 60) and
//     handed to the kernel as a per-exponent constant.
// With these two values the weight stays finite wherever the reference's is (massless: the rho ratio is 1).
#define NIS_U_FLOOR 8.673617379884035e-19   /* 2^-60 */
// r at the inflection point u* = (e-1)/e of the map: which asymptote the Newton start is taken from
#ifdef __CUDACC__
__host__ __device__
#endif
static inline float rambo_rstar(int e) {
    const float ef = (float)e;
    const float ustar = (ef - 1.f) / ef;
    float us = 1.f;
    for (int i = 0; i < e; ++i) us *= ustar;
    return us * ((ef + 1.f) - ef * ustar);
}

static inline double rambo_lattice_at_one(int e) {
    double left = 0.0, right = 1.0, u = 1.0, scale = 0.5;
    for (int level = 0; level < 60; ++level, scale *= 0.5) {
        u = (left + right) * scale;
        double p = 1.0;
        for (int i = 0; i < e; ++i) p *= u;
        const double check = p * ((double)(e + 1) - (double)e * u);
        left *= 2.0; right *= 2.0;
        const double adder = (1.0 <= check) ? -0.5 : 0.5;
        left += adder + 0.5; right += adder - 0.5;
    }
    return u;
}

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
    C->flat_norm = pow(2 * NIS_PI, 4 - 3 * n) * pow(NIS_PI / 2.0, n - 1) / (fact1 * fact2);
    const double flat = pow(2 * NIS_PI, 4 - 3 * n) * pow(NIS_PI / 2.0, n - 1) * (pow(E * E, n - 2) / (fact1 * fact2));
    C->wconst = flat * pow(C->K0 / E, 2 * n - 4) / (2.0 * E * E);   // :403, :307-308 (M_0 = E_cm)
    C->pdf_active = d->pdf_active != 0; C->tau_mode = d->tau_mode != 0;
    C->E_coll = E; C->tau_min = d->tau_min; C->x_cut = d->x_cut;
    C->pdf_grid[0] = d->pdf_grid[0]; C->pdf_grid[1] = d->pdf_grid[1];
    C->pdf_nodes = d->pdf_nodes; C->pdf_lnx_lo = d->pdf_lnx_lo;
    C->pdf_inv_h = d->pdf_nodes > 1 ? (double)(d->pdf_nodes - 1) / (0.0 - d->pdf_lnx_lo) : 0.0;
    if (C->pdf_active) {
        if ((d->pdf_grid[0] || d->pdf_grid[1]) && (d->pdf_nodes < 4 || !(d->pdf_lnx_lo < 0.0))) return NIS_EINVAL;
        if (C->tau_mode && !(d->tau_min > 0.0 && d->tau_min < 1.0)) return NIS_EINVAL;
    }
    const double m1 = d->initial_masses[0], m2 = d->initial_masses[1];
    C->m_in2[0] = (m1 == 0.0 || m2 == 0.0) ? 0.0 : m1 * m1;
    C->m_in2[1] = (m1 == 0.0 || m2 == 0.0) ? 0.0 : m2 * m2;
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
    C->e2_rap = exp(2.0 * d->rap_maxcut); C->e2_rap_inv = exp(-2.0 * d->rap_maxcut);
    C->e2_dR = exp(2.0 * d->delR_mincut);
    C->dR_ge_pi = d->delR_mincut >= NIS_PI;
    C->cos_dR = cos(d->delR_mincut < NIS_PI ? d->delR_mincut : NIS_PI);
    C->u_one[0] = 1.0;
    for (int e = 1; e < NIS_MAX_FINAL; ++e) C->u_one[e] = rambo_lattice_at_one(e);
    C->rstar[0] = 0.f;
    for (int e = 1; e < NIS_MAX_FINAL; ++e) C->rstar[e] = rambo_rstar(e);
    return NIS_OK;
}


// Root in [0,1] of  r = (e+1) u^e - e u^(e+1)  (flat_phase_space_generator.py:101-103,313-359; the
// reference bisects 120-180 levels to float64 resolution).  Here: float32 start from the small-r /
// large-r asymptote, four float32 Newton steps (FP32 pipe, no float64 division), then two
// straight-line float64 Newton steps (see below).
NIS_DEV double rambo_pow_em1(double X, int e) {           // X^(e-1), e >= 2
    if (e == 2) return X;
    if (e == 3) return X * X;
    double p = X * X * X;
    for (int i = 4; i < e; ++i) p *= X;
    return p;
}
// bracketed Newton until the step is below float64 resolution (the rare lanes rambo_root hands over)
#ifdef __CUDACC__
__device__ __noinline__
#else
static
#endif
double rambo_root_slow(int e, double r, double X) {
    const double ed = (double)e;
    double LO = 0.0, HI = 1.0;
    for (int it = 0; it < 80; ++it) {
        const double xe1 = rambo_pow_em1(X, e);
        const double g = xe1 * X * ((ed + 1.0) - ed * X) - r;
        const double dg = ed * (ed + 1.0) * xe1 * (1.0 - X);
        if (g > 0.0) HI = X; else LO = X;
        double xn = X - nis_div(g, dg);
        if (!(xn > LO && xn < HI)) xn = 0.5 * (LO + HI);
        const double dx = fabs(xn - X);
        X = xn;
        if (dx <= 4.5e-16 * X) break;
    }
    return X;
}


NIS_DEV double rambo_root(int e, double r, double u_one, float rstar) {
    if (r >= 1.0) return u_one;                                // see NIS_U_FLOOR above
    if (e == 1) return fmax(nis_div(r, 1.0 + nis_sqrt(1.0 - r)), NIS_U_FLOOR);   // u = 1 - sqrt(1-r), stable form
    if (r <= 0.0) return NIS_U_FLOOR;
    const float ef = (float)e;
    const float rf = (float)r;
    // start on the asymptote of the side the root is on, then four plain float32 Newton steps (a step that
    // leaves (0,1) is dropped).  No bracket: a bracket narrowed with float32 function values ends up one ulp
    // wide, and the bisection fallback then throws a converged iterate far away.  The start only steers the
    // iteration: fast lg2 / ex2 / rsqrt on the device.
    float x;
    if (rf < rstar) x = nis_fpow(nis_fdiv(fmaxf(rf, 1e-37f), ef + 1.f), nis_fdiv(1.f, ef));
    else x = 1.f - nis_fsqrt(nis_fdiv(2.f * fmaxf(1.f - rf, 0.f), ef * (ef + 1.f)));
    x = fminf(fmaxf(x, 0.f), 1.f);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        float xe1 = x;
        for (int i = 2; i < e; ++i) xe1 *= x;
        const float g = xe1 * x * ((ef + 1.f) - ef * x) - rf;
        const float dg = ef * (ef + 1.f) * xe1 * (1.f - x);
        const float xn = x - nis_fdiv(g, dg);
        if (xn > 0.f && xn < 1.f) x = xn;
    }
    // float64 polish.  The float32 iterate is within ~1e-7 of the root, so two Newton steps reach float64
    // resolution (quadratic convergence); they are straight-line code.  Only where the map is flat (r within
    // ~1e-12 of 0 or 1: a handful of events per 10^7) the second step is still large, and those lanes finish in
    // the bracketed loop of rambo_root_slow.
    const double ed = (double)e;
    double X = (double)x, dx = 0.0;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double xe1 = rambo_pow_em1(X, e);
        const double g = xe1 * X * ((ed + 1.0) - ed * X) - r;
        const double dg = ed * (ed + 1.0) * xe1 * (1.0 - X);
        const double xn = X - nis_div(g, dg);
        if (xn > 0.0 && xn < 1.0) { dx = fabs(xn - X); X = xn; }
        else dx = 1.0;                                   // flat region (dg = 0): leave it to the slow path
    }
    if (!(dx <= 1e-11 * X)) X = rambo_root_slow(e, r, X);
    return fmax(X, NIS_U_FLOOR);
}

NIS_DEV double rambo_root_dyn(int e, double r) { return rambo_root(e, r, rambo_lattice_at_one(e), rambo_rstar(e)); }

// x f(x) from the caller's grid (nodes uniform in ln x over [lnx_lo, 0]) by 4-point Lagrange interpolation in ln x,
// divided by x: the density get_pdfQ2 returns (flat_phase_space_generator.py:120-137).  NULL grid: 1.
NIS_DEV double rambo_pdf_density(const double* grid, int nodes, double lnx_lo, double inv_h, double x) {
    if (!grid) return 1.0;
    double t = log(x);
    t = t < lnx_lo ? lnx_lo : (t > 0.0 ? 0.0 : t);
    const double s = (t - lnx_lo) * inv_h;
    int i = (int)s;
    i = i < 1 ? 1 : (i > nodes - 3 ? nodes - 3 : i);
    const double f = s - (double)i;                             // in [0,1) inside the table, up to +-1 at its ends
    const double ym = grid[i - 1], y0 = grid[i], y1 = grid[i + 1], y2 = grid[i + 2];
    const double v = -f * (f - 1.0) * (f - 2.0) * (1.0 / 6.0) * ym + (f + 1.0) * (f - 1.0) * (f - 2.0) * 0.5 * y0
                     - (f + 1.0) * f * (f - 2.0) * 0.5 * y1 + (f + 1.0) * f * (f - 1.0) * (1.0 / 6.0) * y2;
    return v / x;
}

// One event, runtime multiplicity n = C.n.  r: 3n-4 uniforms at stride rs.  mo: scratch AND output row of
// (n+2)*4 doubles at stride ms (beams, final-state momenta; the two beam slots double as scratch for
// exp(2 eta_j) until the end).
//
// Same map as the reference, arranged for the FP64 pipe and a small instruction footprint (every loop is
// rolled: the first version, fully unrolled per multiplicity, was 72 KB of SASS for n=4 and spent 36 % of
// its issue slots waiting for instructions):
//   * q_j = sqrt(lambda_j)/(2 M_j) and rho_j = sqrt(lambda_j)/(8 M_j^2) share one square root; the
//     massive reweighting (:394-403) is accumulated as numerator / denominator products, one division;
//   * the decay energy in the parent rest frame is (M_j^2 + m_j^2 - M_{j+1}^2)/(2 M_j) (= sqrt(q^2+m^2), :262);
//   * the boost by Q (utils.py:58-81, gamma = Q0/M_j because Q is on shell) is
//     p' = p + [(p.Q)/(Q0+M_j) + p0]/M_j * Q,  E' = (Q0 p0 + p.Q)/M_j  — no square root;
//   * sin/cos of phi = 2 pi r from one sincospi (the reference takes cos and restores |sin| by a root);
//   * cuts are decided on squared quantities: pT^2, exp(2 eta) = (|p|+pz)/(|p|-pz); for deltaR two exact
//     quick rejects (|d eta| >= cut, d phi >= cut), then rigorous series bounds on atanh / asin decide
//     all but a ~1e-4-wide shell around deltaR = cut, where the reference's log/acos formula runs.
// All differences to the reference are at float64 rounding level (tests: momenta / weights rtol 1e-9,
// cut masks bit-exact on the golden vectors).
// KIN = false (weight-only call without cuts): only the intermediate masses and the reweighting run.
// PDF = true: two more uniforms sample the Bjorken x; the event is generated at E = sqrt(x1 x2) E_coll, the cuts see the
// lab frame (:157-187, 213-219, 283).
// BEAMS = false (the CUDA kernel without PDFs, where the beams are launch constants it stores itself): `mo` holds
// [n scratch slots | 4n final-state components] instead of [8 beam components | 4n final-state components].
template <bool KIN, bool PDF, bool BEAMS = true>
NIS_DEV void rambo_event(const RamboConst& C, const double* r, int rs, double* mo, int ms, double& weight,
                         uint8_t& pass) {
    const int n = C.n;
    double* const finw = mo + (BEAMS ? 8 : n) * ms;             // final-state particle j at finw + 4*j*ms
    double K0 = C.K0, wconst = C.wconst, E = C.E_coll, x1 = 1.0, x2 = 1.0;
    if (PDF) {
        const double ra = r[(3 * n - 4) * rs], rb = r[(3 * n - 3) * rs];
        double jac = 1.0, rt;
        if (C.tau_mode) {                                       // :161-176, uniform_distr utils.py:124-132
            const double tau = C.tau_min + (1.0 - C.tau_min) * ra;
            const double ymin = 0.5 * log(tau);
            const double ycm = ymin + (-ymin - ymin) * rb;
            rt = sqrt(tau);
            const double ey = exp(ycm);
            x1 = rt * ey; x2 = rt / ey;
            jac = (1.0 - C.tau_min) * (-ymin - ymin);
        } else {                                                // :177-182
            x1 = rb; x2 = ra;
            rt = sqrt(x1 * x2);
        }
        E = rt * C.E_coll;
        const double xc = (x1 < C.x_cut || x2 < C.x_cut) ? 0.0 : 1.0;      // :185-186
        jac *= rambo_pdf_density(C.pdf_grid[0], C.pdf_nodes, C.pdf_lnx_lo, C.pdf_inv_h, x1)
               * rambo_pdf_density(C.pdf_grid[1], C.pdf_nodes, C.pdf_lnx_lo, C.pdf_inv_h, x2) * xc;
        K0 = E - C.msum[0];
        double e2p = 1.0, kp = 1.0;                             // (E^2)^(n-2), (K0/E)^(2n-4)  (:95-97, :403)
        const double E2 = E * E, k2 = (K0 / E) * (K0 / E);
        for (int i = 0; i < n - 2; ++i) { e2p *= E2; kp *= k2; }
        wconst = C.flat_norm * e2p * kp * jac / (2.0 * (x1 * x2 * (C.E_coll * C.E_coll)));   // :306-308
    }
    double Kj = K0, Mj = Kj + C.msum[0];
    double num = 1.0, den = 1.0;
    double Q0 = Mj, Q1 = 0.0, Q2 = 0.0, Q3 = 0.0;
#pragma unroll 1
    for (int j = 0; j < n - 1; ++j) {
        double Kn = 0.0, Mn = C.m[n - 1];
        const double mj = C.m[j];
        if (j < n - 2) {                                        // :363-370, :391-392
            const double u = rambo_root(n - 2 - j, r[j * rs], C.u_one[n - 2 - j], C.rstar[n - 2 - j]);
            Kn = nis_sqrt(u) * Kj;
            Mn = Kn + C.msum[j + 1];
        }
        const double M2 = Mj * Mj, sm_ = Mn + mj, dm = Mn - mj;
        const double sl = nis_sqrt((M2 - sm_ * sm_) * (M2 - dm * dm));   // sqrt(lambda(M_j^2, M_{j+1}^2, m_j^2)), :107-113
        if (j < n - 2) {                                        // :400-401
            const double K2 = Kj * Kj;
            num *= sl * K2 * Mn;
            den *= M2 * (K2 - Kn * Kn) * Kn;
        } else {                                                // 8 rho(M_{n-2}, m_{n-1}, m_{n-2}), :394-397
            num *= sl;
            den *= M2;
        }
        if (!KIN) { Kj = Kn; Mj = Mn; continue; }
        const double i2M = nis_div(0.5, Mj);
        const double q = sl * i2M;                              // :228
        const double ct = 2.0 * r[(n - 2 + 2 * j) * rs] - 1.0;  // :233-243
        const double st = nis_sqrt(1.0 - ct * ct);
        double sp, cp;
        rambo_sincos2pi(r[(n - 1 + 2 * j) * rs], &sp, &cp);
        double p1 = q * st * cp, p2 = q * st * sp, p3 = q * ct;
        const double p0 = (M2 + mj * mj - Mn * Mn) * i2M;
        const double pQ = p1 * Q1 + p2 * Q2 + p3 * Q3;
        const double iM = 2.0 * i2M;
        const double coef = (nis_div(pQ, Q0 + Mj) + p0) * iM;
        p1 += coef * Q1; p2 += coef * Q2; p3 += coef * Q3;
        const double e = (Q0 * p0 + pQ) * iM;
        double* o = finw + j * 4 * ms;
        o[0] = e; o[ms] = p1; o[2 * ms] = p2; o[3 * ms] = p3;
        Q1 -= p1; Q2 -= p2; Q3 -= p3; Q0 -= e;                  // :271-275
        Kj = Kn; Mj = Mn;
    }
    const double w = wconst * nis_div(num, den);
    if (!KIN) { pass = 1; weight = w; return; }
    {
        double* o = finw + (n - 1) * 4 * ms;                    // :278
        o[0] = Q0; o[ms] = Q1; o[2 * ms] = Q2; o[3 * ms] = Q3;
    }
    // ---- cuts (:285-301); x1 = x2 = 1 so the lab frame is the CM frame -------------------------
    bool ok = true;
    // lab frame (boost_to_lab_frame utils.py:134-146): the beams' x1 p1 + x2 p2 moves along z with beta; only p_z
    // enters the cuts (pT and d-phi are invariant): p_z' = gamma (p_z + beta E).  The reference skips the boost for the
    // WHOLE batch when any event has beta = 0 (x1 == x2); here every event is boosted by its own beta (identity then).
    double lb_g = 1.0, lb_gb = 0.0;
    if (PDF) {
        double e1, e2_, z;
        if (C.m_in2[0] == 0.0 && C.m_in2[1] == 0.0) { e1 = 0.5 * E; e2_ = e1; z = e1; }
        else {
            const double E2 = E * E, M1 = C.m_in2[0], M2 = C.m_in2[1];
            e1 = 0.5 * (E2 + M1 - M2) / E; e2_ = 0.5 * (E2 - M1 + M2) / E;
            z = 0.5 * sqrt(E2 * E2 - 2 * E2 * M1 - 2 * E2 * M2 + M1 * M1 - 2 * M1 * M2 + M2 * M2) / E;
        }
        const double r0 = x1 * e1 + x2 * e2_, rz = (x1 - x2) * z;
        if (rz != 0.0) {
            const double beta = rz / r0;
            lb_g = 1.0 / sqrt(1.0 - beta * beta);
            lb_gb = lb_g * beta;
        }
    }
    const double* fin = finw;
    double* e2 = mo;                                            // exp(2 eta_j), j < n <= 8 (beam / scratch slots)
    const bool need_eta = C.rap_max > 0.0 || C.dR_min > 0.0;
    if (C.pT_min > 0.0 || need_eta) {
        double pt2min = NIS_HUGE, e2max = 0.0;
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const double px = fin[(4 * j + 1) * ms], py = fin[(4 * j + 2) * ms];
            const double pz = PDF ? lb_g * fin[(4 * j + 3) * ms] + lb_gb * fin[4 * j * ms] : fin[(4 * j + 3) * ms];
            const double pt2 = px * px + py * py;
            pt2min = fmin(pt2min, pt2);
            if (need_eta) {
                double v;
                if (pt2 < NIS_SQRT_EPS * NIS_SQRT_EPS && pz * pz < NIS_SQRT_EPS * NIS_SQRT_EPS) {
                    v = NIS_HUGE;                               // utils.py:157 `huge`
                } else {
                    const double P = nis_sqrt(pt2 + pz * pz);
                    const double a = P + fabs(pz);              // exp(|eta|) = a / pT
                    const double t = nis_div(a * a, pt2);
                    v = pz >= 0.0 ? t : nis_div(1.0, t);
                }
                e2[j * ms] = v;
                e2max = fmax(e2max, v);
            }
        }
        if (C.pT_min > 0.0 && pt2min < C.pT_min * C.pT_min) ok = false;              // :285-288
        // rap_max < |max_j eta_j|  (|max eta|, not max |eta|, :298-301)
        if (C.rap_max > 0.0 && (e2max > C.e2_rap || e2max < C.e2_rap_inv)) ok = false;
    }
    if (C.dR_min > 0.0) {                                       // :290-296
        const double cut2 = C.dR_min * C.dR_min;
        // Pass 1: the two exact quick rejects for every pair; the survivors (a few per cent of the pairs) are only
        // remembered.  Pass 2 runs the bound arithmetic for them.  Deciding each pair on the spot made a warp execute
        // the ~100-instruction bound code for almost every pair (some lane nearly always survives) with one or two
        // lanes active; deferred, a warp runs it max-over-lanes(survivors) times: ~1.4 instead of ~3.7 for n = 4.
        unsigned cand = 0;
        int pidx = 0;
#pragma unroll 1
        for (int i = 1; i < n; ++i) {
            const double xi = fin[(4 * i + 1) * ms], yi = fin[(4 * i + 2) * ms], ei = e2[i * ms];
            const double pti2 = xi * xi + yi * yi;
#pragma unroll 1
            for (int j = 0; j < i; ++j, ++pidx) {
                const double ej = e2[j * ms];
                // quick reject, exact: |d eta| >= cut  =>  dR >= cut
                if (!(ei < C.e2_dR * ej && ej < C.e2_dR * ei)) continue;
                const double xj = fin[(4 * j + 1) * ms], yj = fin[(4 * j + 2) * ms];
                const double dot = xi * xj + yi * yj;
                const double A = pti2 * (xj * xj + yj * yj);
                // quick reject, exact: d phi >= cut
                if (A > 0.0 && !C.dR_ge_pi && C.cos_dR >= 0.0 && ei < NIS_HUGE && ej < NIS_HUGE &&
                    !(dot > 0.0 && dot * dot > C.cos_dR * C.cos_dR * A)) continue;
                cand |= 1u << pidx;
            }
        }
#pragma unroll 1
        while (cand != 0u && ok) {
#ifdef __CUDACC__
            const int p = __ffs((int)cand) - 1;
#else
            const int p = __builtin_ctz(cand);
#endif
            cand &= cand - 1u;
            int i = 1, first = 0;                               // pair p = (i, j): pairs of row i start at i (i - 1) / 2
            while (first + i <= p) { first += i; ++i; }
            const int j = p - first;
            const double xi = fin[(4 * i + 1) * ms], yi = fin[(4 * i + 2) * ms], ei = e2[i * ms];
            const double xj = fin[(4 * j + 1) * ms], yj = fin[(4 * j + 2) * ms], ej = e2[j * ms];
            const double dot = xi * xj + yi * yj;
            const double A = (xi * xi + yi * yi) * (xj * xj + yj * yj);
            bool exact = true;
            if (A > 0.0 && !C.dR_ge_pi && C.cos_dR >= 0.0 && ei < NIS_HUGE && ej < NIS_HUGE) {
                // series bounds: d eta = atanh(z), z = (rho-1)/(rho+1) = |e_i - e_j| / (e_i + e_j), rho = exp(2 d eta);
                // d phi = asin(s), s = |cross| / (pT_i pT_j)   (0 <= d phi < cut <= pi/2 here)
                const double z = nis_div(fabs(ei - ej), ei + ej), z2 = z * z;
                const double Le = z * (1.0 + z2 * (1.0 / 3.0 + z2 * (1.0 / 5.0 + z2 * (1.0 / 7.0))));
                const double Ue = Le + nis_div(z2 * z2 * z2 * z2 * z, 9.0 * (1.0 - z2));
                const double cr = xi * yj - yi * xj;
                const double s2 = nis_div(cr * cr, A), sn = nis_sqrt(s2);
                const double Lp = sn * (1.0 + s2 * (1.0 / 6.0 + s2 * (3.0 / 40.0 + s2 * (15.0 / 336.0))));
                const double Up = Lp + nis_div(s2 * s2 * s2 * s2 * sn * (35.0 / 1152.0), 1.0 - s2);
                const double m = 1e-12 * cut2;                  // rounding guard of the bound arithmetic
                if (Le * Le + Lp * Lp >= cut2 + m) continue;
                if (Ue * Ue + Up * Up < cut2 - m) { ok = false; exact = false; }
            }
            if (exact) {
                // the reference's formula (utils.py:151-187)
                const double etai = ei >= NIS_HUGE ? NIS_HUGE : 0.5 * log(ei);
                const double etaj = ej >= NIS_HUGE ? NIS_HUGE : 0.5 * log(ej);
                const double deta = etai - etaj;
                double dphi;
                if (A == 0.0) dphi = NIS_HUGE;
                else {
                    double t = dot / sqrt(A);
                    if (fabs(t) > 1.0) t = t / fabs(t);
                    dphi = acos(t);
                }
                const double dR = sqrt(deta * deta + dphi * dphi);
                if (fabs(dR) < C.dR_min) ok = false;
            }
        }
    }
    if (PDF) {                                                  // setInitialStateMomenta_batch with a tensor E_cm, :420-441
        double e1, e2_, z;
        if (C.m_in2[0] == 0.0 && C.m_in2[1] == 0.0) { e1 = 0.5 * E; e2_ = e1; z = e1; }
        else {
            const double E2 = E * E, M1 = C.m_in2[0], M2 = C.m_in2[1];
            e1 = 0.5 * (E2 + M1 - M2) / E; e2_ = 0.5 * (E2 - M1 + M2) / E;
            z = 0.5 * sqrt(E2 * E2 - 2 * E2 * M1 - 2 * E2 * M2 + M1 * M1 - 2 * M1 * M2 + M2 * M2) / E;
        }
        mo[0] = e1; mo[ms] = 0.0; mo[2 * ms] = 0.0; mo[3 * ms] = z;
        mo[4 * ms] = e2_; mo[5 * ms] = 0.0; mo[6 * ms] = 0.0; mo[7 * ms] = -z;
    } else if (BEAMS) {
#pragma unroll 1
        for (int i = 0; i < 8; ++i) mo[i * ms] = C.beam[i >> 2][i & 3];
    }
    pass = ok ? 1 : 0;
    weight = ok ? w : 0.0;
}

// ---------------------------------------------------------------------------------------------------
// Inverse map (the reference lists it as to do, README.md:68-69): final-state momenta in the CM frame ->
// the 3n-4 uniforms generateKinematics_batch would have produced them from, and the weight of that point.
// fin: n x 4 components (E, px, py, pz) at stride ms, r: 3n-4 values at stride rs.  pdf-inactive map only.
//   Q_j = sum_{i >= j} p_i,  M_j = sqrt(Q_j^2) (M_0 = E_cm, M_{n-1} = m_{n-1}),  K_j = M_j - sum_{i >= j} m_i,
//   u_j = (K_{j+1} / K_j)^2,  r_j = (e+1) u^e - e u^{e+1}, e = n-2-j  (the polynomial the forward inverts, :101-105);
//   p_j boosted into the rest frame of Q_j (the forward's boost with Q -> (Q0, -Q)) gives cos(theta) = 2 r - 1 and
//   phi = 2 pi r (:233-243).
// The weight is the forward's own arithmetic on the recovered uniforms (rambo_event<false, false>).
NIS_DEV void rambo_invert_event(const RamboConst& C, const double* fin, int ms, double* r, int rs, double& weight) {
    const int n = C.n;
    double Mj = C.K0 + C.msum[0], Kj = C.K0;                    // M_0 = E_cm: the momenta are CM-frame momenta, Q_0 = (E_cm, 0)
    double Q0 = Mj, Q1 = 0.0, Q2 = 0.0, Q3 = 0.0;
    for (int j = 0; j < n - 1; ++j) {
        const double e = fin[(4 * j) * ms], p1 = fin[(4 * j + 1) * ms], p2 = fin[(4 * j + 2) * ms], p3 = fin[(4 * j + 3) * ms];
        // the angles of p_j in the rest frame of Q_j: the forward's boost with Q -> (Q0, -Q)
        const double pQ = -(p1 * Q1 + p2 * Q2 + p3 * Q3);
        const double coef = nis_div(nis_div(pQ, Q0 + Mj) + e, Mj);
        const double x = p1 - coef * Q1, y = p2 - coef * Q2, z = p3 - coef * Q3;
        const double pm = nis_sqrt(x * x + y * y + z * z);
        double ct = pm > 0.0 ? nis_div(z, pm) : 1.0;
        ct = ct > 1.0 ? 1.0 : (ct < -1.0 ? -1.0 : ct);
        double phi = atan2(y, x) * 0.15915494309189535;         // / (2 pi)
        if (phi < 0.0) phi += 1.0;
        r[(n - 2 + 2 * j) * rs] = 0.5 * (ct + 1.0);
        r[(n - 1 + 2 * j) * rs] = phi;
        if (j < n - 2) {                                        // the mass of what is left, and the uniform behind it
            Q0 -= e; Q1 -= p1; Q2 -= p2; Q3 -= p3;
            const double Mn = nis_sqrt(Q0 * Q0 - (Q1 * Q1 + Q2 * Q2 + Q3 * Q3));
            const double Kn = Mn - C.msum[j + 1];
            double u = nis_div(Kn, Kj);
            u = u * u;
            u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
            const int ex = n - 2 - j;
            const double ue1 = ex == 1 ? 1.0 : rambo_pow_em1(u, ex);          // u^(e-1)
            r[j * rs] = ue1 * u * ((double)(ex + 1) - (double)ex * u);
            Kj = Kn; Mj = Mn;
        }
    }
    uint8_t pass;
    rambo_event<false, false>(C, r, rs, nullptr, 1, weight, pass);
}
