#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "rambo_core.cuh"
__global__ void k(const double* a, const double* b, double* y, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) y[i] = nis_div(a[i], b[i]); }
int main() {
    const int n = 1 << 22; double *ha = new double[n], *hb = new double[n], *hy = new double[n], *da, *db, *dy;
    srand(7); for (int i = 0; i < n; ++i) { ha[i] = ldexp(1.0 + (double)rand() / RAND_MAX + (double)rand() / RAND_MAX * 1e-9, rand() % 400 - 200) * ((rand() & 1) ? 1 : -1);
                                           hb[i] = ldexp(1.0 + (double)rand() / RAND_MAX + (double)rand() / RAND_MAX * 1e-9, rand() % 400 - 200); }
    cudaMalloc(&da, n * 8); cudaMalloc(&db, n * 8); cudaMalloc(&dy, n * 8);
    cudaMemcpy(da, ha, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, n * 8, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(da, db, dy, n); cudaMemcpy(hy, dy, n * 8, cudaMemcpyDeviceToHost);
    int bad = 0, off1 = 0; double worst = 0;
    for (int i = 0; i < n; ++i) { double r = ha[i] / hb[i]; if (hy[i] != r) { double u = fabs(hy[i] - r) / fabs(nextafter(r, 2 * r) - r); if (u > worst) worst = u; if (u <= 1.0) ++off1; else ++bad; } }
    printf("nis_div vs IEEE division on %d pairs: %d differ by one ulp, %d by more (worst %.2f ulp)\n", n, off1, bad, worst);
    return bad != 0;
}
