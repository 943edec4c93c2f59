// Host build of the product's scalar device code (nf_b200/csrc/spline.cuh, rambo_core.cuh) so that the
// CPU test-suite can check the very same source the kernels run against the oracle.  Test harness only.
#include <math.h>
#include <stdint.h>
#include "../../nf_b200/csrc/spline.cuh"
#include "../../nf_b200/csrc/rambo_core.cuh"

extern "C" {

// z: [n][K] logits (K = nb), overwritten with dL/dz.  gy, gJJ per point.
void host_pwlin(int n, int nb, float* z, const float* x, const float* gy, const float* gJJ, float* y, float* f,
                int* k, float* dx) {
    for (int i = 0; i < n; ++i) {
        float S, al;
        float* zi = z + (size_t)i * nb;
        y[i] = pwlin_fwd(zi, 1, nb, x[i], f[i], k[i], S, al);
        dx[i] = pwlin_bwd(zi, 1, nb, k[i], S, al, y[i], f[i], gy[i], gJJ[i]);
    }
}

// z: [n][2nb+1] logits, overwritten with dL/dz.  gy, gf per point.
void host_pwquad(int n, int nb, float* z, const float* x, const float* gy, const float* gf, float* y, float* f,
                 int* k, float* dx) {
    for (int i = 0; i < n; ++i) {
        QuadCtx c;
        float* zi = z + (size_t)i * (2 * nb + 1);
        pwquad_fwd(zi, 1, nb, x[i], c);
        y[i] = c.y; f[i] = c.f; k[i] = c.k;
        dx[i] = pwquad_bwd(zi, 1, nb, c, gy[i], gf[i]);
    }
}

// inverse spline maps: z [n][K] logits (left untouched: a copy is transformed), y -> x, f = density at x, k = bin
void host_spline_inv(int kind, int n, int nb, const float* z, const float* y, float* x, float* f, int* k) {
    const int K = kind == 0 ? nb : 2 * nb + 1;
    float buf[2 * 512 + 1];
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < K; ++j) buf[j] = z[(size_t)i * K + j];
        x[i] = kind == 0 ? pwlin_inv(buf, 1, nb, y[i], f[i], k[i]) : pwquad_inv(buf, 1, nb, y[i], f[i], k[i]);
    }
}

int host_rambo(const NisRamboDesc* d, long long B, const double* r, double* mom, double* w, uint8_t* pass) {
    RamboConst C;
    int rc = rambo_fill_const(d, &C);
    if (rc) return rc;
    const int n = d->n_final, nd = 3 * n - 4 + (d->pdf_active ? 2 : 0), nm = (n + 2) * 4;
    double scratch[(NIS_MAX_FINAL + 2) * 4];
    for (long long i = 0; i < B; ++i) { if (C.pdf_active) rambo_event<true, true>(C, r + i * nd, 1, mom ? mom + i * nm : scratch, 1, w[i], pass[i]); else rambo_event<true, false>(C, r + i * nd, 1, mom ? mom + i * nm : scratch, 1, w[i], pass[i]); }
    return 0;
}

// inverse map: mom [B][n+2][4] -> r [B][3n-4], w [B] (no cuts)
int host_rambo_invert(const NisRamboDesc* d, long long B, const double* mom, double* r, double* w) {
    RamboConst C;
    int rc = rambo_fill_const(d, &C);
    if (rc) return rc;
    if (d->pdf_active) return NIS_EUNSUPPORTED;
    const int n = d->n_final, nd = 3 * n - 4, nm = (n + 2) * 4;
    for (long long i = 0; i < B; ++i) rambo_invert_event(C, mom + i * nm + 8, 1, r + i * nd, 1, w[i]);
    return 0;
}

double host_rambo_root(int e, double r) { return rambo_root_dyn(e, r); }
}
