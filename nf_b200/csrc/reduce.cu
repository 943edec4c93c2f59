// Moments of the Monte Carlo weights and the Philox uniform generator.
//   nis_reduce_moments <- torch.var / torch.mean of f*J (manager.py:255,399-400)
//   nis_uniform_fill   <- torch.nn.init.uniform_(w)      (manager.py:222,395)
#include "common.cuh"

#define RED_NT 256
#define RED_GRID 592   // 148 SMs x 4

// Two-stage deterministic reduction: every CTA writes the statistics of its grid-stride slice in float64; the
// last CTA to finish combines the partials in index order.  NS = 2: (sum, sum of squares); NS = 5: also max, min
// (NaN-propagating like torch.max / torch.min) and the number of non-finite entries.
__device__ __forceinline__ double nan_max(double a, double b) { return (a != a || b != b) ? a + b : fmax(a, b); }
__device__ __forceinline__ double nan_min(double a, double b) { return (a != a || b != b) ? a + b : fmin(a, b); }

template <typename T, int NS>
__global__ void __launch_bounds__(RED_NT) moments_kernel(const T* __restrict__ v, long long n, double* __restrict__ out,
                                                         int accumulate, double* __restrict__ partials,
                                                         unsigned* __restrict__ counter) {
    double s = 0.0, q = 0.0, mx = -INFINITY, mn = INFINITY, bad = 0.0;
    for (long long i = (long long)blockIdx.x * RED_NT + threadIdx.x; i < n; i += (long long)gridDim.x * RED_NT) {
        const double x = (double)v[i];
        s += x; q += x * x;
        if (NS > 2) {
            mx = nan_max(mx, x); mn = nan_min(mn, x);
            bad += (x - x == 0.0) ? 0.0 : 1.0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
        if (NS > 2) {
            mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            bad += __shfl_xor_sync(0xffffffffu, bad, o);
        }
    }
    __shared__ double wsm[5][RED_NT / 32];
    __shared__ bool last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { wsm[0][warp] = s; wsm[1][warp] = q; wsm[2][warp] = mx; wsm[3][warp] = mn; wsm[4][warp] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        s = 0.0; q = 0.0; mx = -INFINITY; mn = INFINITY; bad = 0.0;
        for (int i = 0; i < RED_NT / 32; ++i) {
            s += wsm[0][i]; q += wsm[1][i];
            if (NS > 2) { mx = nan_max(mx, wsm[2][i]); mn = nan_min(mn, wsm[3][i]); bad += wsm[4][i]; }
        }
        double* p = partials + NS * blockIdx.x;
        p[0] = s; p[1] = q;
        if (NS > 2) { p[2] = mx; p[3] = mn; p[4] = bad; }
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    s = 0.0; q = 0.0; mx = -INFINITY; mn = INFINITY; bad = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) {
        const double* p = partials + NS * b;
        s += __ldcg(p); q += __ldcg(p + 1);
        if (NS > 2) { mx = nan_max(mx, __ldcg(p + 2)); mn = nan_min(mn, __ldcg(p + 3)); bad += __ldcg(p + 4); }
    }
    if (accumulate) { out[0] += s; out[1] += q; out[2] += (double)n; }
    else { out[0] = s; out[1] = q; out[2] = (double)n; }
    if (NS > 2) {
        if (accumulate) { out[3] = nan_max(out[3], mx); out[4] = nan_min(out[4], mn); out[5] += bad; }
        else { out[3] = mx; out[4] = mn; out[5] = bad; }
    }
    *counter = 0u;
}

extern "C" size_t nis_reduce_workspace_bytes(void) { return sizeof(double) * 5 * RED_GRID + 256; }

template <int NS>
static int reduce_launch(const void* v, int32_t dtype, int64_t n, double* out, int32_t accumulate, void* workspace,
                         size_t workspace_bytes, void* stream) {
    if ((!v && n > 0) || !out || !workspace || n < 0) return NIS_EINVAL;
    if (workspace_bytes < nis_reduce_workspace_bytes()) return NIS_EWORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    double* partials = (double*)workspace;
    unsigned* counter = (unsigned*)((char*)workspace + sizeof(double) * 5 * RED_GRID);
    cudaMemsetAsync(counter, 0, 4, s);
    long long blocks = (n + RED_NT - 1) / RED_NT;
    int grid = (int)(blocks < 1 ? 1 : (blocks < RED_GRID ? blocks : RED_GRID));
    if (dtype == NIS_F64) moments_kernel<double, NS><<<grid, RED_NT, 0, s>>>((const double*)v, n, out, accumulate, partials, counter);
    else if (dtype == NIS_F32) moments_kernel<float, NS><<<grid, RED_NT, 0, s>>>((const float*)v, n, out, accumulate, partials, counter);
    else return NIS_EINVAL;
    NIS_CUDA_CHECK_LAUNCH();
    return NIS_OK;
}

extern "C" int nis_reduce_moments(const void* v, int32_t dtype, int64_t n, double* out, int32_t accumulate,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    return reduce_launch<2>(v, dtype, n, out, accumulate, workspace, workspace_bytes, stream);
}

extern "C" int nis_reduce_stats(const void* v, int32_t dtype, int64_t n, double* out, int32_t accumulate,
                                void* workspace, size_t workspace_bytes, void* stream) {
    return reduce_launch<5>(v, dtype, n, out, accumulate, workspace, workspace_bytes, stream);
}

// ---- Philox4x32-10 ---------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}

// element i <- word (i & 3) of Philox(counter = (offset + i) >> 2, key = seed); [0,1) with 24 (f32) or
// 32+21 (f64 from two words of the same block: elements are then 2 per block) bits
template <typename T>
__global__ void uniform_kernel(T* __restrict__ out, long long n, unsigned long long seed, unsigned long long offset) {
    const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long e = offset + (unsigned long long)i;
        if (sizeof(T) == 4) {
            const unsigned long long blk = e >> 2;
            const uint4 rnd = philox4x32_10(make_uint4((unsigned)blk, (unsigned)(blk >> 32), 0u, 0u), key);
            const unsigned w = (e & 3) == 0 ? rnd.x : ((e & 3) == 1 ? rnd.y : ((e & 3) == 2 ? rnd.z : rnd.w));
            out[i] = (T)((float)(w >> 8) * (1.0f / 16777216.0f));
        } else {
            const unsigned long long blk = e >> 1;
            const uint4 rnd = philox4x32_10(make_uint4((unsigned)blk, (unsigned)(blk >> 32), 1u, 0u), key);
            const unsigned a = (e & 1) ? rnd.z : rnd.x, b = (e & 1) ? rnd.w : rnd.y;
            const unsigned long long m = ((unsigned long long)a << 21) | (b >> 11);     // 53 bits
            out[i] = (T)((double)m * (1.0 / 9007199254740992.0));
        }
    }
}

extern "C" int nis_uniform_fill(void* out, int32_t dtype, int64_t n, uint64_t seed, uint64_t offset, void* stream) {
    if (!out || n < 0) return NIS_EINVAL;
    if (n == 0) return NIS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    long long blocks = (n + 255) / 256;
    int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
    if (dtype == NIS_F32) uniform_kernel<float><<<grid, 256, 0, s>>>((float*)out, n, seed, offset);
    else if (dtype == NIS_F64) uniform_kernel<double><<<grid, 256, 0, s>>>((double*)out, n, seed, offset);
    else return NIS_EINVAL;
    NIS_CUDA_CHECK_LAUNCH();
    return NIS_OK;
}

// ---- FP32 FMA pipe probe (roofline denominator for the compute-bound flow kernels) -------------------
// Every thread runs 16 independent FMA chains for `iters` rounds: 32 * iters flop per thread.
__global__ void __launch_bounds__(256) fma_probe_kernel(float* __restrict__ out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456f) out[0] = s;      // never true in practice; keeps the chains alive
}

extern "C" int64_t nis_probe_fp32_fma(float* out, int32_t iters, void* stream) {
    if (!out || iters <= 0) return NIS_EINVAL;
    const int grid = 148 * 8;
    fma_probe_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, iters, 0.999f, 0.001f);
    NIS_CUDA_CHECK_LAUNCH();
    return (int64_t)grid * 256 * 32 * (int64_t)iters;
}

extern "C" const char* nis_strerror(int code) {
    switch (code) {
        case NIS_OK: return "ok";
        case NIS_EINVAL: return "invalid descriptor or argument";
        case NIS_EWORKSPACE: return "workspace too small (see nis_flow_workspace_bytes)";
        case NIS_ECUDA: return "CUDA launch failed";
        case NIS_EUNSUPPORTED: return "configuration not supported by this build";
    }
    return "unknown error";
}

extern "C" size_t nis_sizeof_flow_desc(void) { return sizeof(NisFlowDesc); }
extern "C" size_t nis_sizeof_rambo_desc(void) { return sizeof(NisRamboDesc); }

extern "C" const char* nis_version(void) { return "nis_b200 0.1 (sm_100a)"; }
