// api_sketch_launch.h -- launch geometry of sketch_kernel<WINDOWED, Consumer>, shared by the sketch translation units.
#pragma once
#include "api_internal.h"
#include "sketch_kernels.cuh"
#include "sketch_fast.cuh"
#include "stream_kernels.cuh"

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
    a.keymask = 0xFFFFFFFFu;
    a.filter = c->filter_n ? c->filter.p ? reinterpret_cast<const uint64_t *>(c->filter.p) : nullptr : nullptr; a.filter_n = c->filter_n;
    if (const char *ev = getenv("D2G_FAST_KEYMASK")) a.keymask = (uint32_t)strtoul(ev, nullptr, 0);   // test knob
    a.m = m; a.tile_stride = 1; a.score_slots = d2g::sketch_score_slots(p->k, p->w);
    a.span = pick_span(c, rg.pos_end - rg.pos_base, m);
    return a;
}

// element streams (stream_kernels.cuh): the same spans and Consumer protocol over item regions instead of sequence positions
template <class Consumer>
int launch_stream(d2g_ctx *c, const d2g::SketchArgs &a, const typename Consumer::Params &cp, int tcls) {
    KernelTimer kt(c, tcls);
    const size_t smem = d2g::stream_smem_bytes<Consumer>(a.m, a.w > a.k);
    if (smem > 200 * 1024) return fail(D2G_EUNSUPPORTED, "sketch with %u registers needs %zu bytes of shared memory per CTA (max 200 KiB)", a.m, smem);
    const uint64_t grid = (a.pos_end - a.pos_base + a.span - 1) / a.span;
    if (a.filter_n) {
        CU(cudaFuncSetAttribute(d2g::stream_kernel<Consumer, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        d2g::stream_kernel<Consumer, true><<<(unsigned)grid, d2g::SK_THREADS, smem, c->stream>>>(a, cp);
    } else {
        CU(cudaFuncSetAttribute(d2g::stream_kernel<Consumer>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        d2g::stream_kernel<Consumer><<<(unsigned)grid, d2g::SK_THREADS, smem, c->stream>>>(a, cp);
    }
    c->launches++;
    CU(cudaGetLastError());
    return D2G_OK;
}

template <class Consumer>
int launch_sketch(d2g_ctx *c, const d2g::SketchArgs &a, const typename Consumer::Params &cp, bool windowed, int tcls = D2G_T_SKETCH_MAIN) {
    if (a.seq.items) return launch_stream<Consumer>(c, a, cp, tcls);
    KernelTimer kt(c, tcls);
    const size_t smem = d2g::sketch_smem_bytes<Consumer>(a.m, a.score_slots);
    if (smem > 200 * 1024) return fail(D2G_EUNSUPPORTED, "sketch with %u registers needs %zu bytes of shared memory per CTA (max 200 KiB)", a.m, smem);
    const uint64_t grid = (a.pos_end - a.pos_base + a.span - 1) / a.span;
    if (a.filter_n) {            // --filterset: the same kernels with the membership test in front of the consumer
        if (windowed) {
            CU(cudaFuncSetAttribute(d2g::sketch_kernel<true, Consumer, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            d2g::sketch_kernel<true, Consumer, true><<<(unsigned)grid, d2g::SK_THREADS, smem, c->stream>>>(a, cp);
        } else {
            CU(cudaFuncSetAttribute(d2g::sketch_kernel<false, Consumer, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            d2g::sketch_kernel<false, Consumer, true><<<(unsigned)grid, d2g::SK_THREADS, smem, c->stream>>>(a, cp);
        }
    } else if (!windowed && a.k == 31 && !getenv("D2G_FAST_NO_K31")) {   // the unwindowed BASELINE shape (configs[0]): k = 31 folded into the kernel
        CU(cudaFuncSetAttribute(d2g::sketch_kernel<false, Consumer, false, 31>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        d2g::sketch_kernel<false, Consumer, false, 31><<<(unsigned)grid, d2g::SK_THREADS, smem, c->stream>>>(a, cp);
    } else if (windowed) {
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

// Windowed set sketches take the 32-bit-key kernel (sketch_fast.cuh) followed by the exact kernel over the tiles the fast pass put
// on its redo list (normally none); everything else takes the exact kernel.  D2G_NO_FAST=1 forces the exact kernel (test knob).
inline bool sketch_fast_eligible(const d2g::SketchArgs &a) {
    const int wsz = a.w - a.k + 1;
    return a.w > a.k && a.canon && wsz >= 2 && wsz <= d2g::SF_MAX_WSZ && a.tile_stride == 1 && !a.filter_n && !getenv("D2G_NO_FAST");
}
constexpr uint64_t kRedoCap = 1ULL << 16;

template <class Consumer>
int launch_sketch_windowed_set(d2g_ctx *c, const d2g::SketchArgs &a, const typename Consumer::Params &cp, int tcls = D2G_T_SKETCH_MAIN) {
    if (a.seq.items || !sketch_fast_eligible(a)) return launch_sketch<Consumer>(c, a, cp, true, tcls);
    if (int rc = c->redo.reserve((kRedoCap + 2) * 8)) return rc;
    unsigned long long *redo_count = c->redo.as<unsigned long long>();
    uint64_t *redo_list = c->redo.as<uint64_t>() + 2;
    CU(cudaMemsetAsync(redo_count, 0, 16, c->stream));
    const size_t cbytes = Consumer::smem_bytes(a.m, true);
    const size_t smem = d2g::sketch_fast_smem_bytes(cbytes);
    if (smem > 200 * 1024) return launch_sketch<Consumer>(c, a, cp, true, tcls);
    const uint64_t grid = (a.pos_end - a.pos_base + a.span - 1) / a.span;
    const d2g::FastAux fx{redo_count, redo_list, kRedoCap};
    {
        KernelTimer kt(c, tcls);
        if (a.w - a.k + 1 == 21 && a.k == 31 && !getenv("D2G_FAST_NO_K31")) {
            CU(cudaFuncSetAttribute(d2g::sketch_fast_kernel<21, Consumer, 31>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            d2g::sketch_fast_kernel<21, Consumer, 31><<<(unsigned)grid, d2g::SK_THREADS, smem, c->stream>>>(a, cp, fx);
        } else if (a.w - a.k + 1 == 21) {
            CU(cudaFuncSetAttribute(d2g::sketch_fast_kernel<21, Consumer>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            d2g::sketch_fast_kernel<21, Consumer><<<(unsigned)grid, d2g::SK_THREADS, smem, c->stream>>>(a, cp, fx);
        } else {
            CU(cudaFuncSetAttribute(d2g::sketch_fast_kernel<0, Consumer>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            d2g::sketch_fast_kernel<0, Consumer><<<(unsigned)grid, d2g::SK_THREADS, smem, c->stream>>>(a, cp, fx);
        }
        c->launches++;
        CU(cudaGetLastError());
    }
    // exact pass over the listed tiles; the grid is fixed (the count stays on the device) and returns at once when the list is empty
    const size_t smem_x = d2g::sketch_smem_bytes<Consumer>(a.m, a.score_slots);
    CU(cudaFuncSetAttribute(d2g::sketch_redo_kernel<Consumer>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x));
    d2g::sketch_redo_kernel<Consumer><<<(unsigned)(c->sm_count * 2), d2g::SK_THREADS, smem_x, c->stream>>>(a, cp, redo_list, redo_count, kRedoCap);
    c->launches++;
    CU(cudaGetLastError());
    if (getenv("D2G_DEBUG")) {
        unsigned long long h = 0;
        cudaMemcpyAsync(&h, redo_count, 8, cudaMemcpyDeviceToHost, c->stream); cudaStreamSynchronize(c->stream);
        fprintf(stderr, "[d2g] fast windowed kernel: %llu tile(s) recomputed by the exact kernel (of %llu)\n", h,
                (unsigned long long)((a.pos_end - a.pos_base + d2g::SF_TILE - 1) / d2g::SF_TILE));
    }
    return D2G_OK;
}
} // namespace
