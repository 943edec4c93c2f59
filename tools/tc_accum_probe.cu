// Probe: how does tcgen05.mma.kind::tf32 round when it adds a K=8 step into the fp32 accumulator in TMEM?
// Every instruction adds exactly x = 8*a*b (a, b exact in TF32; x exact in fp32) to D.  The host simulates the
// chain with round-to-nearest and with truncation and prints which one the hardware matches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I nf_b200/csrc tools/tc_accum_probe.cu -o oracle/_ref/tc_accum_probe
#include <cstdio>
#include <cmath>
#include <cfenv>
#include <cuda_runtime.h>
#include "tc_common.cuh"

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__global__ void probe(float a, float b, int steps, float* out) {
    extern __shared__ char smraw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tb_s;
    char* sm = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~(uintptr_t)1023);
    float* A = reinterpret_cast<float*>(sm);              // [128][32]
    float* B = reinterpret_cast<float*>(sm + 16384);      // [16][32]
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) A[i] = a;
    for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) B[i] = b;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb_s)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tb_s;
    if (threadIdx.x == 0) {
        const uint32_t idesc = tc_idesc(128, 16);
        for (int s = 0; s < steps; ++s) mma_ss(tb, tc_desc(smem_u32(A)), tc_desc(smem_u32(B)), idesc, s > 0);
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    float v[16];
    tc_ld16(tb + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
    tc_ld_wait();
    if (threadIdx.x == 0) out[0] = v[0];
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(32));
}

static float chain(float x, int steps, int mode) {
    fesetround(mode);
    volatile float acc = 0.f;
    for (int s = 0; s < steps; ++s) acc = acc + x;
    fesetround(FE_TONEAREST);
    return acc;
}

int main() {
    float* out;
    cudaMalloc(&out, 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    const float cases[3][2] = {{1.f + 1.f / 1024, 1.f + 1.f / 1024}, {1.f + 3.f / 1024, 1.f + 5.f / 1024}, {-(1.f + 1.f / 1024), 1.f + 7.f / 1024}};
    for (int c = 0; c < 3; ++c) {
        const float a = cases[c][0], b = cases[c][1];
        const float x = 8.f * (float)((double)a * (double)b);
        for (int steps : {24, 96, 384}) {
            probe<<<1, 128, 32768>>>(a, b, steps, out);
            float h = 0.f;
            cudaError_t e = cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
            const double exact = (double)steps * 8.0 * (double)a * (double)b;
            printf("a=%.10g b=%.10g steps=%3d  tcgen05=%.9g  rn-chain=%.9g  rz-chain=%.9g  exact=%.12g  (%s)  rel err %.3g\n", a, b, steps, h,
                   chain(x, steps, FE_TONEAREST), chain(x, steps, FE_TOWARDZERO), exact, cudaGetErrorString(e), (h - exact) / exact);
        }
    }
    return 0;
}
