// Fused coupling-cell forward on the 5th-generation tensor cores (tcgen05 + TMEM), width-64 PWLin cells.
//
// The FP32-pipe kernel (flow_tiled.cu) is bound by shared-memory -> register fill (ncu: 83 % of the
// shared pipe, FMA pipe 46 %).  tcgen05.mma reads its operands straight from shared memory and
// accumulates in tensor memory, so that bound disappears.  To keep the float32 contract (1e-5 on points
// and log-Jacobians) every conditioner product runs as a 3xTF32 split: a = a_hi + a_lo with a_hi the
// value rounded to TF32's 10-bit mantissa and a_lo = a - a_hi (exact in float32, then rounded), likewise w;
// D += a_hi*w_hi + a_hi*w_lo + a_lo*w_hi drops only a_lo*w_lo (~2^-22 relative), the same order as the
// rounding of a float32 FMA chain of length 64.  (hi/lo are formed with round-to-nearest, see tf32_rn.)
//
// Warp-specialised CTA of 9 warps: two groups of 4 warps each own a tile of 128 points (thread m of a
// group owns TMEM lane m) and one warp issues the MMAs for both, so one group's epilogue / spline work on
// the FP32 pipe overlaps the other group's tensor-core work.  Layer 0 (K = P <= 16) runs on the FP32
// pipe; every 64x64 hidden layer and the 64 -> T*K output layer are tcgen05.mma.kind::tf32 with M=128,
// K=8 per instruction, the A operand (activations) read from TENSOR MEMORY, where the epilogue
// (tcgen05.ld of the accumulator row, BN scale/shift, ReLU, hi/lo split) writes it back with tcgen05.st,
// and the B operand (weights, split once per forward) from shared memory in the K-major 128-byte-swizzled
// canonical layout.  The spline (softmax, CDF, bin, Jacobian) runs on the thread's 32 logits in
// registers.  Groups and the MMA warp hand tiles to each other through mbarriers (128 arrivals when a
// group's A operand is in TMEM; tcgen05.commit when its accumulator is complete).
//
// The same kernel serves the train-mode layer passes of flow_tiled.cu's scheme (statistics pass L: read
// the stored pre-BN activations of layer L-1, one MMA layer, per-feature sums by recursive-halving warp
// shuffles, store layer L), whose tile-blocked [tile][64][128] buffers it shares.
#include <stdlib.h>
#include "common.cuh"
#include "spline.cuh"
#include "flow_fwd_common.cuh"

#include "tc_common.cuh"
#include "tc_spline.cuh"

__global__ void flow_tc_pack_kernel(DevFlow F, const float* __restrict__ params, float* __restrict__ tcpack) {
    const int c = blockIdx.y;
    const DevCell& q = F.cells[c];
    const float* p = params + q.param_off;
    char* dst = reinterpret_cast<char*>(tcpack + (size_t)c * tc_cell_floats(F));
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int l = 1; l < F.depth; ++l) {
        const float* w = p + F.p_lin(c, l);                     // [64][64] torch layout (out, in)
        char* hi = dst + (size_t)(l - 1) * 2 * TCH * TCH * 4;
        char* lo = hi + (size_t)TCH * TCH * 4;
        for (int i = tid; i < TCH * TCH; i += nth) {
            const int n = i / TCH, k = i - n * TCH;
            const float v = w[(size_t)n * TCH + k];
            const float h = tf32_rn(v);
            const int o = tc_off(TCH, n, k);
            *reinterpret_cast<float*>(hi + o) = h;
            *reinterpret_cast<float*>(lo + o) = tf32_rn(v - h);
        }
    }
    const float* wo = p + F.p_lin(c, F.depth);                  // [T*K][64]
    const int Nb = tc_out_n(F), tper = tc_out_tper(F), slot = tc_out_slot(F);
    for (int b = 0; b < tc_out_blocks(F, q.T); ++b) {
        char* hi = dst + (size_t)(F.depth - 1) * 2 * TCH * TCH * 4 + (size_t)b * 2 * Nb * TCH * 4;
        char* lo = hi + (size_t)Nb * TCH * 4;
        for (int i = tid; i < Nb * TCH; i += nth) {
            const int n = i / TCH, k = i - n * TCH;
            const int t = b * tper + n / slot, jj = n % slot;
            const float v = (t < q.T && jj < F.K) ? wo[((size_t)t * F.K + jj) * TCH + k] : 0.f;
            const float h = tf32_rn(v);
            const int o = tc_off(Nb, n, k);
            *reinterpret_cast<float*>(hi + o) = h;
            *reinterpret_cast<float*>(lo + o) = tf32_rn(v - h);
        }
    }
}


template <int KIND>
__global__ void __launch_bounds__(TC_THREADS, 1) flow_cell_tc_kernel(const __grid_constant__ DevFlow F, const FwdArgs A,
                                                                      const float* __restrict__ tcpack) {
    extern __shared__ char smraw[];
    __shared__ uint64_t a_ready[2], d_ready[2], z_full[2][2], s_full[2];
    __shared__ uint32_t tmem_base_s;
    char* sm = smem_align1024(smraw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = A.c_begin;
    const DevCell& q = F.cells[c];
    const int d = F.d, depth = F.depth;
    // which slice of the cell this launch computes (see the table in flow_tiled.cu / DESIGN.md)
    const bool stats = A.stats_layer >= 1;
    const bool from_z = A.zin != nullptr;
    const int lz = from_z ? (stats ? A.stats_layer - 1 : depth) : 1;      // v starts as z_{lz}
    const int l_end = stats ? A.stats_layer - 1 : depth;                   // MMA layers lz .. l_end
    // stored activations are staged through shared memory (bulk copies) unless the out-layer weights of a
    // PWQuad final pass leave no room: then the rows are read straight from global memory
    const bool zst = (from_z || A.zout != nullptr) && (stats || F.K == 32);
    const TcSmem L = tc_layout(F, q.P, lz, l_end, zst);
    float* w0s = reinterpret_cast<float*>(sm + L.w0);
    float* affs = reinterpret_cast<float*>(sm + L.aff);
    float* biass = reinterpret_cast<float*>(sm + L.bias);
    const float* pk = A.wpack + q.pk_off;

    // ---- one-time setup: weights, barriers, tensor memory ---------------------------------------------
    for (int l = lz; l <= l_end; ++l) {
        if (L.wl[l] < 0) continue;
        const int fl = l == depth ? tc_out_blocks(F, q.T) * 2 * tc_out_n(F) * TCH : 2 * TCH * TCH;
        const float4* src = reinterpret_cast<const float4*>(tcpack + (size_t)c * tc_cell_floats(F) + (size_t)(l - 1) * 2 * TCH * TCH);
        float4* dst = reinterpret_cast<float4*>(sm + L.wl[l]);
        for (int i = tid; i < fl / 4; i += TC_THREADS) dst[i] = src[i];
    }
    if (!from_z) {
        const float* s0 = pk + q.wt_off[0];                       // layer 0, [P][64] k-major
        for (int i = tid; i < q.P * TCH; i += TC_THREADS) w0s[i] = s0[i];
    }
    for (int l = 0; l <= depth; ++l) {
        const int W = l == 0 ? q.P : TCH, Wp = pad8(W);
        const float* s = pk + q.aff_off[l];
        for (int i = tid; i < W; i += TC_THREADS) { affs[l * 2 * TCH + i] = s[i]; affs[l * 2 * TCH + TCH + i] = s[Wp + i]; }
    }
    {
        const int Nb = tc_out_n(F), tper = tc_out_tper(F), slot = tc_out_slot(F);
        for (int i = tid; i < tc_out_blocks(F, q.T) * Nb; i += TC_THREADS) {
            const int b = i / Nb, n = i - b * Nb, t = b * tper + n / slot, jj = n % slot;
            biass[i] = (t < q.T && jj < F.K) ? pk[q.bo_off + t * F.Kpad + jj] : 0.f;
        }
    }
    if (tid == 0) {
        mbar_init(&a_ready[0], TCM); mbar_init(&a_ready[1], TCM);
        mbar_init(&d_ready[0], 1); mbar_init(&d_ready[1], 1);
        mbar_init(&z_full[0][0], 1); mbar_init(&z_full[0][1], 1); mbar_init(&z_full[1][0], 1); mbar_init(&z_full[1][1], 1);
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const long long ntiles = (A.B + TCM - 1) / TCM;
    const long long rowlen = d + 1;
    double dsum[2] = {0.0, 0.0}, dsq[2] = {0.0, 0.0};

    if (warp == 8) {
        // ===================== MMA issuer ======================================================
        if (lane == 0 && lz <= l_end) {
            uint32_t pa[2] = {0, 0};
            const int Nb = tc_out_n(F), nblk = tc_out_blocks(F, q.T);
            const uint32_t idesc64 = tc_idesc(TCM, TCH), idescO = tc_idesc(TCM, Nb);
            for (long long it = 0;; ++it) {
                const long long t0 = ((long long)blockIdx.x + it * gridDim.x) * 2;
                if (t0 >= ntiles) break;
                for (int l = lz; l <= l_end; ++l) {
                    const bool outl = l == depth;
                    const int nb_ = outl ? nblk : 1;
                    for (int b = 0; b < nb_; ++b) {
                        const int rows = outl ? Nb : TCH;
                        const uint32_t whi = smem_u32(sm + L.wl[l]) + (outl ? (uint32_t)b * 2 * Nb * TCH * 4 : 0u);
                        for (int g = 0; g < 2; ++g) {
                            if (t0 + g >= ntiles) continue;
                            mbar_wait(&a_ready[g], pa[g]);
                            pa[g] ^= 1;
                            tc_fence_after();
                            const uint32_t tb = tmem_base + g * TC_COLS_PER_GROUP;
                            tc_issue_layer(tb, tb + TC_COL_AHI, tb + TC_COL_ALO, whi, whi + rows * TCH * 4, rows,
                                           outl ? idescO : idesc64);
                            tc_commit(&d_ready[g]);
                        }
                    }
                }
            }
        }
    } else {
        // ===================== point groups ====================================================
        const int g = warp >> 2, gt = tid & (TCM - 1);
        float* st = reinterpret_cast<float*>(sm + L.st) + g * (d + 1) * TCM + gt;   // this thread's state row
        const uint32_t tg = tmem_base + g * TC_COLS_PER_GROUP + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t pd = 0;
        // stored activations arrive by bulk copy, one tile ahead: the copy of tile it+1 is issued when tile
        // it starts (its buffer was last read during tile it-1, which every thread of the group left
        // before the MMA of that tile could complete)
        float* zs = reinterpret_cast<float*>(sm + L.zb) + g * 2 * TCH * TCM;
        constexpr uint32_t ZBYTES = TCH * TCM * 4;
        if (from_z && zst && gt == 0) {
            const long long t0 = (long long)blockIdx.x * 2 + g;
            if (t0 < ntiles) bulk_load(zs, A.zin + (size_t)t0 * TCH * TCM, ZBYTES, &z_full[g][0]);
        }
        // the state rows of a full tile are one contiguous block: they arrive by bulk copy too, issued one tile
        // ahead (the landing zone is free as soon as every thread has moved its row into its state column)
        float* sst = reinterpret_cast<float*>(sm + L.sst) + g * (d + 1) * TCM;
        const uint32_t SBYTES = (uint32_t)((d + 1) * TCM * 4);
        const bool st_bulk = A.from_state && (reinterpret_cast<uintptr_t>(A.state_in) & 15) == 0;
        uint32_t sph = 0;
        if (st_bulk && gt == 0) {
            const long long t0 = (long long)blockIdx.x * 2 + g;
            if ((t0 + 1) * TCM <= A.B) bulk_load(sst, A.state_in + t0 * TCM * rowlen, SBYTES, &s_full[g]);
        }
        for (long long it = 0;; ++it) {
            const long long tile = ((long long)blockIdx.x + it * gridDim.x) * 2 + g;
            if (tile >= ntiles) break;
            if (from_z && zst && gt == 0) {
                const long long tn = ((long long)blockIdx.x + (it + 1) * gridDim.x) * 2 + g;
                if (A.zout) bulk_store_wait_read();       // tile it-1's output has left that buffer
                if (tn < ntiles) bulk_load(zs + ((it + 1) & 1) * TCH * TCM, A.zin + (size_t)tn * TCH * TCM, ZBYTES, &z_full[g][(it + 1) & 1]);
            }
            const long long pt = tile * TCM + gt;
            const bool valid = pt < A.B;
            // ---- this thread's point --------------------------------------------------------------
            if (st_bulk && (tile + 1) * TCM <= A.B) {
                mbar_wait(&s_full[g], sph);
                sph ^= 1;
                for (int i = 0; i <= d; ++i) st[i * TCM] = sst[gt * rowlen + i];
                proxy_fence();                       // our reads of the landing zone precede the next bulk write into it
                group_sync(g);
                if (gt == 0) {
                    const long long tn = ((long long)blockIdx.x + (it + 1) * gridDim.x) * 2 + g;
                    if ((tn + 1) * TCM <= A.B) bulk_load(sst, A.state_in + tn * TCM * rowlen, SBYTES, &s_full[g]);
                }
            } else if (valid) {
                if (A.from_state) {
                    for (int i = 0; i <= d; ++i) st[i * TCM] = A.state_in[pt * rowlen + i];
                } else {
                    for (int i = 0; i < d; ++i) st[i * TCM] = load_io(A.in, A.in_dtype, pt * A.in_cols + i);
                    st[d * TCM] = A.in_cols > d ? load_io(A.in, A.in_dtype, pt * A.in_cols + d) : 1.f;
                }
                if (!stats && A.saved && !A.from_state) {
                    float* sv = A.saved + ((long long)c * A.B + pt) * rowlen;
                    for (int i = 0; i <= d; ++i) sv[i] = st[i * TCM];
                }
            } else {
                for (int i = 0; i < d; ++i) st[i * TCM] = 0.5f;
                st[d * TCM] = 1.f;
            }
            // ---- v = z_{lz}: stored activations, or layer 0 on the FP32 pipe ---------------------------
            float v[TCH];
            if (from_z && zst) {
                mbar_wait(&z_full[g][it & 1], (uint32_t)((it >> 1) & 1));
                const float* zr = zs + (it & 1) * TCH * TCM + gt;
#pragma unroll
                for (int j = 0; j < TCH; ++j) v[j] = zr[j * TCM];
            } else if (from_z) {
                const float* zr = A.zin + (size_t)tile * TCH * TCM + gt;
#pragma unroll
                for (int j = 0; j < TCH; ++j) v[j] = zr[(size_t)j * TCM];
            } else {
#pragma unroll
                for (int j = 0; j < TCH; ++j) v[j] = 0.f;
                for (int k = 0; k < q.P; ++k) {
                    const float a = fmaf(st[q.feed[k] * TCM], affs[k], affs[TCH + k]);
                    const float4* wr = reinterpret_cast<const float4*>(w0s + k * TCH);
#pragma unroll
                    for (int j4 = 0; j4 < TCH / 4; ++j4) {
                        const float4 w = wr[j4];
                        v[4 * j4] = fmaf(a, w.x, v[4 * j4]); v[4 * j4 + 1] = fmaf(a, w.y, v[4 * j4 + 1]);
                        v[4 * j4 + 2] = fmaf(a, w.z, v[4 * j4 + 2]); v[4 * j4 + 3] = fmaf(a, w.w, v[4 * j4 + 3]);
                    }
                }
            }
            // ---- hidden MMA layers ------------------------------------------------------------------------
            const int l_hid = l_end < depth ? l_end : depth - 1;
            for (int l = lz; l <= l_hid; ++l) {
                tc_store_act(v, affs + l * 2 * TCH, affs + l * 2 * TCH + TCH, tg + TC_COL_AHI, tg + TC_COL_ALO);
                tc_fence_before();
                mbar_arrive(&a_ready[g]);
                mbar_wait(&d_ready[g], pd);
                pd ^= 1;
                tc_fence_after();
                tc_ld32(tg, v);
                tc_ld32(tg + 32, v + 32);
                tc_ld_wait();
            }
            if (stats) {
                // ---- statistics pass: z_L goes to the staging tile [64][128] (in place over the input tile),
                //      from there to HBM by ONE bulk store, and its per-feature sums are row sums of the tile
                float* zo = zs + (it & 1) * TCH * TCM;
                if (!from_z && gt == 0) bulk_store_wait_read();          // (no prefetch in this pass: guard reuse here)
                if (!from_z) group_sync(g);
#pragma unroll
                for (int j = 0; j < TCH; ++j) zo[j * TCM + gt] = valid ? v[j] : 0.f;
                proxy_fence();
                group_sync(g);
                if (gt == 0 && A.zout) bulk_store(A.zout + (size_t)tile * TCH * TCM, zo, ZBYTES);
                if (!A.no_stats) {
                    // thread gt sums half a row (64 points) of feature gt & 63, 16 bytes at a time; the start is
                    // rotated by the lane so that a warp's 32 rows do not hit the same banks
                    const float4* row = reinterpret_cast<const float4*>(zo + (gt & 63) * TCM + (gt >> 6) * 64);
                    float s_ = 0.f, q_ = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 x = row[(i + lane) & 15];
                        s_ += (x.x + x.y) + (x.z + x.w);
                        q_ = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, q_))));
                    }
                    dsum[0] += (double)s_; dsq[0] += (double)q_;
                }
                group_sync(g);
                continue;
            }
            // ---- output layer, one MMA block at a time, and the splines on this thread's logits ----------------
            tc_store_act(v, affs + depth * 2 * TCH, affs + depth * 2 * TCH + TCH, tg + TC_COL_AHI, tg + TC_COL_ALO);
            float jfac = 1.f;
            const int nblk = tc_out_blocks(F, q.T), tper = tc_out_tper(F), Nb = tc_out_n(F);
            tc_fence_before();
            mbar_arrive(&a_ready[g]);
            for (int b = 0; b < nblk; ++b) {
                mbar_wait(&d_ready[g], pd);
                pd ^= 1;
                tc_fence_after();
                for (int tt = 0; tt < tper; ++tt) {
                    const int t = b * tper + tt;
                    if (t >= q.T) break;
                    const float xv = st[q.trafo[t] * TCM];
                    float y, f;
                    int kb;
                    if (KIND == NIS_KIND_PWLIN) {
                        // PWLin, 32 bins (coupling_cells.py:114-141), registers only
                        float z[32];
                        tc_ld32(tg + tt * 32, z);
                        tc_ld_wait();
                        const float a = xv * 32.f;
                        kb = (int)floorf(a);
                        kb = kb < 0 ? 0 : (kb > 31 ? 31 : kb);
                        const float alpha = a - (float)kb;
                        float m = -3.0e38f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) { z[j] += biass[b * Nb + tt * 32 + j]; m = fmaxf(m, z[j]); }
                        float S = 0.f, C = 0.f, ek = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float e = __expf(z[j] - m);
                            S += e;
                            C += j < kb ? e : 0.f;
                            ek = j == kb ? e : ek;
                        }
                        const float inv = 1.f / S;
                        y = (ek * alpha + C) * inv;
                        f = ek * inv * 32.f;
                    } else {
                        float z[80];
                        tc_ld32(tg, z);
                        tc_ld32(tg + 32, z + 32);
                        tc_ld16(tg + 64, z + 64);
                        tc_ld_wait();
                        if (b + 1 < nblk) {                 // accumulator consumed: the next block's MMA may overwrite it
                            tc_fence_before();
                            mbar_arrive(&a_ready[g]);
                        }
#pragma unroll
                        for (int j = 0; j < 65; ++j) z[j] += biass[b * Nb + j];
                        pwquad32_regs(z, xv, y, f, kb);
                    }
                    st[q.trafo[t] * TCM] = y;
                    jfac *= f;
                    if (A.bins && valid) A.bins[((long long)c * A.B + pt) * d + t] = kb;
                }
            }
            st[d * TCM] *= jfac;
            // ---- store ----------------------------------------------------------------------------------
            if (valid) {
                if (A.state_out) {
                    float* so = A.state_out + pt * rowlen;
                    for (int i = 0; i <= d; ++i) so[i] = st[i * TCM];
                }
                if (A.to_out) {
                    for (int i = 0; i < d; ++i) store_io(A.out, A.out_dtype, pt * rowlen + i, st[F.out_perm[i] * TCM]);
                    store_io(A.out, A.out_dtype, pt * rowlen + d, st[d * TCM]);
                }
            }
        }
    }
    bulk_store_wait_read();
    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    if (!stats || A.no_stats) return;
    // ---- fold: thread gt of a group holds the sums of feature gt & 63 over half of each of its tiles -----------
    double* red = reinterpret_cast<double*>(sm + L.red);          // [2][256]: sum / sum of squares per group thread
    double* sacc = red + 8 * 2 * TCH;                              // [2 * maxW]
    if (warp < 8) { red[tid] = dsum[0]; red[256 + tid] = dsq[0]; }
    for (int i = tid; i < 2 * F.maxW; i += TC_THREADS) sacc[i] = 0.0;
    __syncthreads();
    if (tid < TCH) {
        double s = 0.0, s2 = 0.0;
        for (int k = 0; k < 4; ++k) { s += red[tid + 64 * k]; s2 += red[256 + tid + 64 * k]; }
        sacc[tid] = s; sacc[F.maxW + tid] = s2;
    }
    bn_stats_finalize(F, A, sacc, TC_THREADS);
}

// ---------------------------------------------------------------------------------------------------
size_t nis_tc_pack_floats(const DevFlow& F) {
    if (F.depth < 1) return 0;
    return (size_t)F.n_cells * tc_cell_floats(F);
}

// smallest batch the tensor-core kernels take (below it the shape-generic kernel runs); NIS_TC_MIN_B overrides (test knob)
int64_t nis_tc_min_batch(int64_t dflt) {
    const char* e = getenv("NIS_TC_MIN_B");
    return e ? atoll(e) : dflt;
}

bool nis_tc_supported(const DevFlow& F, int64_t B, int bn_mode) {
    (void)bn_mode;
    const char* off = getenv("NIS_TC");                   // NIS_TC=0 forces the FP32-pipe kernels (test knob)
    if (off && off[0] == '0') return false;
    if (F.depth < 1 || B < nis_tc_min_batch(256) || F.maxW != TCH || F.nb != 32) return false;
    for (int l = 0; l < F.depth; ++l) if (F.widths[l] != TCH) return false;
    if (F.kind == NIS_KIND_PWLIN ? F.K != 32 : F.K != 65) return false;
    for (int c = 0; c < F.n_cells; ++c) {
        const int P = F.cells[c].P, T = F.cells[c].T;
        if (P > 16) return false;
        if (F.K == 32 && T * 32 > TC_NOUT) return false;
        const size_t lim = 226 * 1024;
        if ((size_t)tc_layout(F, P, 1, F.depth, false).total + 1024 > lim && F.K == 32) return false;      // fused eval cell
        if ((size_t)tc_layout(F, P, 1, F.depth - 1, true).total + 1024 > lim) return false;               // hidden pass(es)
        if ((size_t)tc_layout(F, P, F.depth, F.depth, F.K == 32).total + 1024 > lim) return false;        // final pass
    }
    return true;
}

// PWQuad cells keep 160 KB of output-layer weights resident, so their eval-mode cell is split in two
// launches (hidden layers -> stored activations -> output layer + spline) like the train-mode passes.
bool nis_tc_split_eval(const DevFlow& F) { return F.K != 32; }

int nis_tc_pack(const DevFlow& F, const float* params, float* tcpack, cudaStream_t s) {
    flow_tc_pack_kernel<<<dim3(16, F.n_cells), 256, 0, s>>>(F, params, tcpack);
    NIS_CUDA_CHECK_LAUNCH();
    return NIS_OK;
}

int nis_launch_tc(const DevFlow& F, const FwdArgs& A, const float* tcpack, cudaStream_t s) {
    static int sms = 0;
    if (sms <= 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const bool stats = A.stats_layer >= 1;
    const int lz = A.zin ? (stats ? A.stats_layer - 1 : F.depth) : 1;
    const int l_end = stats ? A.stats_layer - 1 : F.depth;
    const bool zst = (A.zin != nullptr || A.zout != nullptr) && (stats || F.K == 32);
    const size_t smem = (size_t)tc_layout(F, F.cells[A.c_begin].P, lz, l_end, zst).total + 1024;
    auto kern = F.kind == NIS_KIND_PWLIN ? flow_cell_tc_kernel<NIS_KIND_PWLIN> : flow_cell_tc_kernel<NIS_KIND_PWQUAD>;
    if (F.kind == NIS_KIND_PWLIN) NIS_ENSURE_SMEM((flow_cell_tc_kernel<NIS_KIND_PWLIN>), (int)smem);
    else NIS_ENSURE_SMEM((flow_cell_tc_kernel<NIS_KIND_PWQUAD>), (int)smem);
    long long npairs = ((A.B + TCM - 1) / TCM + 1) / 2;
    int grid = (int)(npairs < sms ? npairs : sms);
    kern<<<grid, TC_THREADS, smem, s>>>(F, A, tcpack);
    NIS_CUDA_CHECK_LAUNCH();
    return NIS_OK;
}
