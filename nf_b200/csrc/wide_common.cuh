// Shared by the streamed-weights tcgen05 kernels for wide conditioners (flow_wide.cu forward,
// flow_bwd_wide.cu backward): operand-pack geometry, shared-memory layout, float64 warp sums.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

#define WD_THREADS 192        // 4 point warps + MMA issuer + TMA producer
#define WD_COL_X 192          // cross-term accumulator D_x [192,384); D_hi is [0,192)
#define WD_COL_A 384          // A chunks: [384,448) and [448,512): 32 hi + 32 lo columns each
#define WD_NMAX 192
#define WD_MAX_SLOTS 4

__host__ __device__ static inline int wd_kp16(const DevFlow& F) { return (F.K + 15) & ~15; }
// floats of one cell's operand pack: hidden layers 1..depth-1 as W/32 panels of [W][32] (hi, lo), then per
// transformed dimension the output layer as W/32 panels of [Kp16][32] (hi, lo)
__host__ __device__ static inline size_t wd_cell_floats(const DevFlow& F) {
    const int W = F.widths[0];
    int T = 0;
    for (int c = 0; c < F.n_cells; ++c) T = F.cells[c].T > T ? F.cells[c].T : T;
    return (size_t)(F.depth - 1) * W * W * 2 + (size_t)T * wd_kp16(F) * W * 2;
}

// hidden layers: outputs per round (N of the MMAs) and rounds per tile
__host__ __device__ static inline int wd_hid_n(int W) { return W <= WD_NMAX ? W : W / 2; }
__host__ __device__ static inline int wd_hid_rounds(int W) { return W <= WD_NMAX ? 1 : 2; }

// float64 variant of tc_warp_feature_sums: sums of a[i] and a[i]^2 over the warp's 32 points, features
// (2 lane, 2 lane + 1) on each lane.  The pre-BN activations of a wide layer are sums of 256 products; their
// batch variance is formed as E[z^2] - E[z]^2, and float32 partial sums cost a factor two in log J here.
__device__ __forceinline__ void wd_warp_feature_sums64(const float* v, int lane, double* s, double* s2) {
    double a[32], b[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const bool up = lane & 16;
        const double x0 = (double)v[i], x1 = (double)v[i + 32];
        const double keep = up ? x1 : x0, send = up ? x0 : x1;
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        b[i] = keep * keep + __shfl_xor_sync(0xffffffffu, send * send, 16);
    }
#pragma unroll
    for (int w = 16; w >= 2; w >>= 1) {
        const int m = w >> 1;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < w) {
                const bool up = lane & m;
                const double ka = up ? a[i + w] : a[i], sa = up ? a[i] : a[i + w];
                const double kb = up ? b[i + w] : b[i], sb = up ? b[i] : b[i + w];
                a[i] = ka + __shfl_xor_sync(0xffffffffu, sa, m);
                b[i] = kb + __shfl_xor_sync(0xffffffffu, sb, m);
            }
        }
    }
    s[0] = a[0]; s[1] = a[1]; s2[0] = b[0]; s2[1] = b[1];
}

struct WdSmem { int ring, slots, slot_bytes, w0, aff, bias, st, stg, red, total; };
__host__ __device__ static inline WdSmem wd_layout(const DevFlow& F, int P, bool final_pass, bool from_state) {
    WdSmem s;
    const int W = F.widths[0], Kp = wd_kp16(F);
    int T = 0;
    for (int c = 0; c < F.n_cells; ++c) T = F.cells[c].T > T ? F.cells[c].T : T;
    const int npanel = final_pass ? Kp : wd_hid_n(W);
    s.slot_bytes = npanel * 256;
    int other = 0;
    const int w0b = from_state ? pad8(P) * W * 4 : 0;
    const int affb = (2 * 16 + 2 * W) * 4;
    const int biasb = final_pass ? T * Kp * 4 : 0;
    const int stb = (F.d + 1) * TCM * 4;
    const int stgb = final_pass ? Kp * TCM * 4 : 64 * TCM * 4;    // logits of one dimension / statistics staging tile [64][128]
    const int redb = 2 * W * 8;
    other = w0b + affb + biasb + stb + stgb + redb + 256;
    int slots = (226 * 1024 - other) / s.slot_bytes;
    if (slots > WD_MAX_SLOTS) slots = WD_MAX_SLOTS;
    s.slots = slots;
    int o = 0;
    s.ring = o; o += slots * s.slot_bytes;
    s.w0 = o; o += w0b;
    s.aff = o; o += affb;
    s.bias = o; o += biasb;
    s.st = o; o += stb;
    s.stg = o; o += stgb;
    o = (o + 7) & ~7;
    s.red = o; o += redb;
    s.total = o;
    return s;
}

