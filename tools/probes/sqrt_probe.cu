#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "rambo_core.cuh"
__global__ void k(const double* x, double* y, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) y[i] = nis_sqrt(x[i]); }
int main() {
    const int n = 1 << 22; double *hx = new double[n], *hy = new double[n], *dx, *dy;
    srand(5); for (int i = 0; i < n; ++i) { double m = 1.0 + (double)rand() / RAND_MAX; int e = rand() % 600 - 300; hx[i] = ldexp(m + (double)rand() / RAND_MAX * 1e-9, e); }
    cudaMalloc(&dx, n * 8); cudaMalloc(&dy, n * 8); cudaMemcpy(dx, hx, n * 8, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(dx, dy, n); cudaMemcpy(hy, dy, n * 8, cudaMemcpyDeviceToHost);
    int bad = 0, off1 = 0; double worst = 0;
    for (int i = 0; i < n; ++i) { double r = sqrt(hx[i]); if (hy[i] != r) { double u = fabs(hy[i] - r) / (nextafter(r, 2 * r) - r); if (u > worst) worst = u; if (u <= 1.0) ++off1; else ++bad; } }
    printf("nis_sqrt vs IEEE sqrt on %d values: %d differ by one ulp, %d by more (worst %.2f ulp)\n", n, off1, bad, worst);
    return bad != 0;
}
