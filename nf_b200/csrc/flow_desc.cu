// Host-side descriptor handling for the flow kernels: validation, arena layouts, workspace carve-up.
#include <string.h>
#include "common.cuh"

static int cond_out(const NisFlowDesc* d) {
    return d->kind == NIS_KIND_PWLIN ? d->n_bins : (d->kind == NIS_KIND_AFFINE ? 2 : 2 * d->n_bins + 1);
}

extern "C" int64_t nis_flow_cell_param_count(const NisFlowDesc* d, int32_t c) {
    if (!d || c < 0 || c >= d->n_cells) return NIS_EINVAL;
    int P = d->cells[c].n_pass, T = d->n_flow - P, in = P;
    int64_t n = 2 * P;
    for (int l = 0; l < d->depth; ++l) { n += (int64_t)d->widths[l] * in + 2 * d->widths[l]; in = d->widths[l]; }
    n += (int64_t)T * cond_out(d) * in + (int64_t)T * cond_out(d);
    return n;
}

extern "C" int64_t nis_flow_cell_bn_count(const NisFlowDesc* d, int32_t c) {
    if (!d || c < 0 || c >= d->n_cells) return NIS_EINVAL;
    int64_t n = 2 * d->cells[c].n_pass;
    for (int l = 0; l < d->depth; ++l) n += 2 * d->widths[l];
    return n;
}

int nis_build_dev_flow(const NisFlowDesc* d, DevFlow* F) {
    if (!d || !F) return NIS_EINVAL;
    if (d->n_flow < 2 || d->n_flow > NIS_MAX_DIM) return NIS_EINVAL;
    if (d->n_cells < 1 || d->n_cells > NIS_MAX_CELLS) return NIS_EINVAL;
    if (d->kind != NIS_KIND_PWLIN && d->kind != NIS_KIND_PWQUAD && d->kind != NIS_KIND_AFFINE) return NIS_EINVAL;
    if (d->n_bins < 1 || d->n_bins > 512) return NIS_EINVAL;
    if (d->depth < 0 || d->depth > NIS_MAX_HIDDEN) return NIS_EINVAL;
    memset(F, 0, sizeof(*F));
    F->d = d->n_flow; F->n_cells = d->n_cells; F->kind = d->kind; F->nb = d->n_bins; F->depth = d->depth;
    F->K = cond_out(d); F->Kpad = pad8(F->K);
    F->eps = d->bn_eps; F->momentum = d->bn_momentum;
    int maxW = 8;
    for (int l = 0; l < d->depth; ++l) {
        if (d->widths[l] < 1 || d->widths[l] > NIS_MAX_WIDTH) return NIS_EINVAL;
        F->widths[l] = d->widths[l];
        if (pad8(d->widths[l]) > maxW) maxW = pad8(d->widths[l]);
    }
    bool seen[NIS_MAX_DIM];
    memset(seen, 0, sizeof(seen));
    for (int i = 0; i < d->n_flow; ++i) {
        int p = d->out_perm[i];
        if (p < 0 || p >= d->n_flow || seen[p]) return NIS_EINVAL;
        seen[p] = true;
        F->out_perm[i] = (uint8_t)p;
    }
    for (int c = 0; c < d->n_cells; ++c)
        if (pad8(d->cells[c].n_pass) > maxW) maxW = pad8(d->cells[c].n_pass);
    F->maxW = maxW;
    int pk = 0;
    for (int c = 0; c < d->n_cells; ++c) {
        const NisCellDesc& s = d->cells[c];
        DevCell& q = F->cells[c];
        if (s.n_pass < 1 || s.n_pass >= d->n_flow) return NIS_EINVAL;
        q.P = s.n_pass; q.T = d->n_flow - s.n_pass;
        memset(seen, 0, sizeof(seen));
        for (int i = 0; i < d->n_flow; ++i) {
            int col = i < q.P ? s.feed_idx[i] : s.trafo_idx[i - q.P];
            if (col < 0 || col >= d->n_flow || seen[col]) return NIS_EINVAL;
            seen[col] = true;
            if (i < q.P) q.feed[i] = (uint8_t)col; else q.trafo[i - q.P] = (uint8_t)col;
        }
        if (s.param_off < 0 || s.bn_off < 0) return NIS_EINVAL;
        q.param_off = s.param_off; q.bn_off = s.bn_off;
        q.pk_off = pk;
        int o = 0;
        // each layer's scale / shift block starts on its own 128-byte line: the cooperative small-batch kernels let one CTA
        // rewrite a block while the others still hold its neighbours in L1 (no stale line may be shared)
        for (int l = 0; l <= d->depth; ++l) { q.aff_off[l] = o; o += 2 * F->Wp(c, l); o = (o + 31) & ~31; }
        int in = q.P;
        for (int l = 0; l < d->depth; ++l) { q.wt_off[l] = o; o += in * pad8(d->widths[l]); in = d->widths[l]; }
        q.wo_off = o; o += q.T * in * F->Kpad;
        q.bo_off = o; o += q.T * F->Kpad;
        pk += (o + 31) & ~31;
        q.sv_off = c * (d->depth + 1) * 2 * maxW;
    }
    F->pack_total = pk;
    F->saved_total = d->n_cells * (d->depth + 1) * 2 * maxW;
    return NIS_OK;
}

extern "C" int64_t nis_flow_bn_saved_count(const NisFlowDesc* d) {
    DevFlow F;
    int rc = nis_build_dev_flow(d, &F);
    return rc ? rc : F.saved_total;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Backward scratch (floats), see flow_bwd.cu: gradient state [B][d+1], per-CTA parameter-gradient
// partials [NIS_BWD_GRID][max cell params], BN backward sums [n_cells][depth+1][2][maxW].

size_t nis_tc_pack_floats(const DevFlow& F);
size_t nis_wide_pack_floats(const DevFlow& F);

size_t nis_flow_carve(const DevFlow& F, int64_t B, void* base, FlowWorkspace* ws) {
    size_t off = 0;
    char* b = (char*)base;
    ws->wpack = (float*)(b + off); off = align256(off + sizeof(float) * (size_t)F.pack_total);
    {   // tensor-core operand pack: resident-weights layout (flow_tc.cu) or K-panels (flow_wide.cu)
        size_t fl = nis_tc_pack_floats(F), wf = nis_wide_pack_floats(F);
        ws->tcpack = (float*)(b + off); off = align256(off + sizeof(float) * (fl > wf ? fl : wf));
    }
    ws->state = (float*)(b + off); off = align256(off + sizeof(float) * (size_t)B * (F.d + 1));
    ws->partials = (double*)(b + off); off = align256(off + sizeof(double) * (size_t)NIS_MAX_GRID * 2 * F.maxW);
    ws->counter = (unsigned*)(b + off); off = align256(off + 256);
    ws->bwd = (float*)(b + off);
    ws->total = off;
    return off;
}
