// RAMBO-on-diet 2 -> n: momenta + Jacobian weight + pT / deltaR / rapidity cuts in one pass.
// One thread per event, float64.  The uniforms of a 128-event tile are staged through shared memory
// with coalesced 16-byte loads and the momenta tile is written back the same way; rows are padded to
// an odd number of doubles so the per-thread row accesses are bank-conflict free.
// Reference: nisrep/PhaseSpace/flat_phase_space_generator.py:139-308 (see rambo_core.cuh).
#include <math.h>
#include "common.cuh"
#include "rambo_core.cuh"

#define RAMBO_NT 128

template <typename RT>
__global__ void __launch_bounds__(RAMBO_NT) rambo_kernel(const __grid_constant__ RamboConst C, const RT* __restrict__ r,
                                                         double* __restrict__ momenta, double* __restrict__ weight,
                                                         uint8_t* __restrict__ cutmask, long long B) {
    const int ND = 3 * C.n - 4;             // uniforms per event
    const int NDP = ND | 1;                 // odd row stride (doubles): conflict-free per-thread rows
    const int NM = (C.n + 2) * 4;           // momentum components per event
    const int NMP = NM | 1;
    extern __shared__ __align__(16) double smd[];
    double* rs = smd;                       // [NT][NDP]
    double* mo = smd + RAMBO_NT * NDP;      // [NT][NMP]  (scratch even when momenta are not requested)
    const int tid = threadIdx.x;
    // i / ND and i / NM by multiplication (exact for i < 2^14, divisor <= 128: i * (M d - 2^24) < 2^24)
    const unsigned long long magD = (1u << 24) / ND + 1, magM = (1u << 24) / NM + 1;
    const long long ntiles = (B + RAMBO_NT - 1) / RAMBO_NT;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * RAMBO_NT;
        const int cnt = (int)((B - base) < RAMBO_NT ? (B - base) : RAMBO_NT);
        const RT* src = r + base * ND;
        for (int i = tid; i < cnt * ND; i += RAMBO_NT) {
            const int ev = (int)(((unsigned long long)i * magD) >> 24), c = i - ev * ND;
            rs[ev * NDP + c] = (double)src[i];
        }
        __syncthreads();
        if (tid < cnt) {
            double w;
            uint8_t pass;
            rambo_event(C, rs + tid * NDP, 1, mo + tid * NMP, 1, w, pass);
            weight[base + tid] = w;
            if (cutmask) cutmask[base + tid] = pass;
        }
        __syncthreads();
        if (momenta) {
            double* dst = momenta + base * NM;
            for (int i = tid; i < cnt * NM; i += RAMBO_NT) {
                const int ev = (int)(((unsigned long long)i * magM) >> 24), c = i - ev * NM;
                dst[i] = mo[ev * NMP + c];
            }
        }
    }
}

template <typename RT>
static int rambo_launch(const RamboConst& C, const void* r, double* momenta, double* weight, uint8_t* cutmask,
                        long long B, cudaStream_t s) {
    const int NDP = (3 * C.n - 4) | 1, NMP = ((C.n + 2) * 4) | 1;
    const size_t smem = sizeof(double) * RAMBO_NT * (NDP + NMP);
    cudaFuncSetAttribute(rambo_kernel<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long ntiles = (B + RAMBO_NT - 1) / RAMBO_NT;
    int grid = (int)(ntiles < 148 * 16 ? ntiles : 148 * 16);
    rambo_kernel<RT><<<grid, RAMBO_NT, smem, s>>>(C, (const RT*)r, momenta, weight, cutmask, B);
    NIS_CUDA_CHECK_LAUNCH();
    return NIS_OK;
}

extern "C" int nis_rambo_generate(const NisRamboDesc* desc, const void* r, int32_t r_dtype, double* momenta,
                                  double* weight, uint8_t* cutmask, int64_t B, void* stream) {
    if (!desc || B < 0 || (B > 0 && (!r || !weight))) return NIS_EINVAL;
    if (r_dtype != NIS_F32 && r_dtype != NIS_F64) return NIS_EINVAL;
    RamboConst C;
    int rc = rambo_fill_const(desc, &C);
    if (rc) return rc;
    if (B == 0) return NIS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    return r_dtype == NIS_F64 ? rambo_launch<double>(C, r, momenta, weight, cutmask, B, s)
                              : rambo_launch<float>(C, r, momenta, weight, cutmask, B, s);
}
