// api_sketch_launch.h -- launch geometry of sketch_kernel<WINDOWED, Consumer>, shared by the sketch translation units.
#pragma once
#include "api_internal.h"
#include "sketch_kernels.cuh"

namespace {
uint64_t pick_span(const d2g_ctx *c, uint64_t total_len, uint32_t m) {
    // enough CTAs for ~8 waves, but spans long enough that the per-CTA register flush (m atomics) is noise
    const uint64_t min_span = std::max<uint64_t>(16ULL * d2g::SK_TILE, 16ULL * m);
    uint64_t span = total_len / ((uint64_t)c->sm_count * 32) + 1;
    span = std::max(span, min_span);
    span = (span + d2g::SK_TILE - 1) / d2g::SK_TILE * d2g::SK_TILE;
    return span;
}

// A launch may cover only part of a batch (chunked host uploads): start positions [pos_base, pos_end) of the
// sequence buffer, whose records are rec_off_d[0..n_rec] (absolute offsets) and whose entities start at ent_base.
d2g::SketchArgs make_sketch_args(const d2g_ctx *c, const d2g_sketch_params *p, const d2g::PackedSeq &seq_d, const uint64_t *rec_off_d,
                                 const uint32_t *rec_ent_d, uint64_t n_rec, uint64_t total_len, uint32_t m, const SketchRange &rg) {
    d2g::SketchArgs a;
    a.seq = seq_d; a.rec_off = rec_off_d; a.rec_entity = rec_ent_d;
    a.n_rec = n_rec; a.k = p->k; a.w = p->w; a.canon = p->canon; a.xormask = p->xormask;
    a.pos_base = rg.pos_base; a.pos_end = rg.pos_end; a.ent_base = rg.ent_base; a.ent_state = nullptr; a.want_state = 0;
    a.m = m; a.tile_stride = 1; a.score_slots = d2g::sketch_score_slots(p->k, p->w);
    a.span = pick_span(c, rg.pos_end - rg.pos_base, m);
    return a;
}

template <class Consumer>
int launch_sketch(d2g_ctx *c, const d2g::SketchArgs &a, const typename Consumer::Params &cp, bool windowed, int tcls = D2G_T_SKETCH_MAIN) {
    KernelTimer kt(c, tcls);
    const size_t smem = d2g::sketch_smem_bytes<Consumer>(a.m, a.score_slots);
    if (smem > 200 * 1024) return fail(D2G_EUNSUPPORTED, "sketch with %u registers needs %zu bytes of shared memory per CTA (max 200 KiB)", a.m, smem);
    const uint64_t grid = (a.pos_end - a.pos_base + a.span - 1) / a.span;
    if (windowed) {
        CU(cudaFuncSetAttribute(d2g::sketch_kernel<true, Consumer>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        d2g::sketch_kernel<true, Consumer><<<(unsigned)grid, d2g::SK_THREADS, smem, c->stream>>>(a, cp);
    } else {
        CU(cudaFuncSetAttribute(d2g::sketch_kernel<false, Consumer>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        d2g::sketch_kernel<false, Consumer><<<(unsigned)grid, d2g::SK_THREADS, smem, c->stream>>>(a, cp);
    }
    c->launches++;
    CU(cudaGetLastError());
    return D2G_OK;
}
} // namespace
