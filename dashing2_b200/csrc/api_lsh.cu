// api_lsh.cu -- LSH-assisted top-k neighbour graphs of the C ABI (include/d2gpu.h).
#include "api_internal.h"
#include "lsh_kernels.cuh"
#include <cub/device/device_segmented_radix_sort.cuh>
#include <cub/device/device_radix_sort.cuh>

// -------------------------------------------------------------------------------------------------
// LSH top-k
// -------------------------------------------------------------------------------------------------
extern "C" int d2g_lsh_topk(d2g_ctx *c, const d2g_cmp_params *p, const double *regs, const double *cards, int32_t topk,
                            uint64_t *indptr_out, uint32_t **idx_out, float **val_out) {
    return d2g_lsh_topk_rows(c, p, regs, cards, topk, 0, p ? p->n : 0, indptr_out, idx_out, val_out);
}

extern "C" int d2g_lsh_topk_rows(d2g_ctx *c, const d2g_cmp_params *p, const double *regs, const double *cards, int32_t topk,
                                 uint64_t x0, uint64_t x1, uint64_t *indptr_out, uint32_t **idx_out, float **val_out) {
    if (topk <= 0) return fail(D2G_EINVAL, "topk must be > 0 (similarity-threshold graphs: d2g_lsh_graph)");
    if (p && p->cmp_kind >= D2G_CMP_SS_COMPRESSED) return fail(D2G_EINVAL, "compressed registers: the index is built over the f64 signatures, pass both to d2g_lsh_graph");
    return d2g_lsh_graph(c, p, nullptr, regs, cards, topk, 0., x0, x1, indptr_out, idx_out, val_out);
}

extern "C" int d2g_lsh_graph(d2g_ctx *c, const d2g_cmp_params *p, const double *index_regs, const double *regs, const double *cards, int32_t topk,
                             double min_similarity, uint64_t x0, uint64_t x1, uint64_t *indptr_out, uint32_t **idx_out, float **val_out) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_cmp_params(p)) return rc;
    if (x0 > x1 || x1 > p->n) return fail(D2G_EINVAL, "bad row range");
    const bool threshold = topk <= 0;
    if (!index_regs) index_regs = regs;
    if (p->cmp_kind >= D2G_CMP_SS_COMPRESSED && index_regs == regs)
        return fail(D2G_EINVAL, "compressed registers: pass the f64 signatures as index_regs (the reference builds the index before it compresses, src/cmp_core.cpp:741-799)");
    if (!indptr_out || !idx_out || !val_out) return fail(D2G_EINVAL, "null output");
    const uint64_t n = p->n; const uint32_t S = p->sketchsize;
    if (n < 2 || x0 == x1) { for (uint64_t i = 0; i <= x1 - x0; ++i) indptr_out[i] = 0; *idx_out = (uint32_t *)malloc(4); *val_out = (float *)malloc(4); return D2G_OK; }
    if (n >= 0x7FFFFFFFULL) return fail(D2G_EINVAL, "too many sketches for 32-bit LSH ids");
    if (S < 2) return fail(D2G_EINVAL, "sketchsize must be >= 2 for the default two LSH table types");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int nlsh = p->nlsh ? p->nlsh : 2;
    if (nlsh < 1 || nlsh > d2g::LSH_MAX_TYPES)
        return fail(D2G_EUNSUPPORTED, "--nLSH %d: 1 .. %d are implemented (keys of more than 16 registers leave XXH3's 128-byte branch, src/ssi.h:345-352)", p->nlsh, d2g::LSH_MAX_TYPES);
    d2g::LshGeom geom{};
    geom.ntypes = (uint32_t)nlsh;
    for (int ty = 0; ty < nlsh; ++ty) {                                // cmp_core.cpp:757-770
        const uint32_t nper = d2g::lsh_nper((uint32_t)ty);
        geom.cnt[ty] = nper <= 2 ? S / nper : (uint32_t)((uint64_t)S * 8 / nper);
        geom.start[ty] = ty ? geom.start[ty - 1] + geom.cnt[ty - 1] : 0;
        if (nper > S) return fail(D2G_EINVAL, "--nLSH %d needs at least %u registers", nlsh, nper);
    }
    for (int ty = nlsh - 1, acc = 0; ty >= 0; --ty) { geom.scan0[ty] = (uint32_t)acc; acc += (int)geom.cnt[ty]; }
    const uint32_t ntab = geom.ntab = geom.start[nlsh - 1] + geom.cnt[nlsh - 1];
    uint64_t ntoquery = threshold ? n - 1 : (uint64_t)((float)topk * 3.5f);   // index_build.cpp:56-60: no cap on the candidates of a threshold graph
    ntoquery = std::min<uint64_t>(ntoquery, n - 1);
    CU(cudaFuncSetAttribute(d2g::lsh_trim_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, d2g::LSH_TRIM_BIG_CAP * 8));
    if (ntoquery == 0 || (!threshold && ntoquery > 4096)) return fail(D2G_EUNSUPPORTED, "topk %d out of the supported range", topk);
    if (threshold && ntoquery * 8 > 200 * 1024)
        return fail(D2G_EUNSUPPORTED, "similarity-threshold graphs keep one uncapped candidate list per query in shared memory: at most %d sketches per call (got %llu)",
                    200 * 1024 / 8 + 1, (unsigned long long)n);
    const uint32_t maxcand = (uint32_t)ntoquery;
    if (int rc = c->cregs.reserve(n * S * 8)) return rc;
    if (int rc = c->ccards.reserve(n * 8)) return rc;
    CU(cudaMemcpyAsync(c->cregs.p, index_regs, n * S * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->ccards.p, cards, n * 8, cudaMemcpyHostToDevice, st));
    const double *regs_d = c->cregs.as<double>(), *cards_d = c->ccards.as<double>();
    const double *cmp_regs_d = regs_d;                                  // what refinement compares: the compressed registers under --fastcmp
    if (regs != index_regs) {
        if (int rc = c->lregs.reserve(n * S * 8)) return rc;
        CU(cudaMemcpyAsync(c->lregs.p, regs, n * S * 8, cudaMemcpyHostToDevice, st));
        cmp_regs_d = c->lregs.as<double>();
    }
    auto al = [](uint64_t b) { return (b + 255) / 256 * 256; };
    const uint64_t nk = (uint64_t)ntab * n, na = 2 * n * maxcand;
    if (na >= 0x7FFFFFF0ULL) return fail(D2G_EUNSUPPORTED, "n * topk too large for one call (%llu arrival slots)", (unsigned long long)na);
    uint64_t off = 0;
    const uint64_t o_kA = off; off += al(nk * 4); const uint64_t o_kB = off; off += al(nk * 4);
    const uint64_t o_iA = off; off += al(nk * 4); const uint64_t o_iB = off; off += al(nk * 4);
    const uint64_t o_offs = off; off += al(((uint64_t)ntab + 1) * 8);
    const uint64_t o_cand = off; off += al(n * maxcand * 4); const uint64_t o_cnt = off; off += al(n * maxcand * 4);
    const uint64_t o_nc = off; off += al(n * 4);
    const uint64_t o_alA = off; off += al(na * 4); const uint64_t o_alB = off; off += al(na * 4);
    const uint64_t o_apA = off; off += al(na * 8); const uint64_t o_apB = off; off += al(na * 8);
    const uint64_t o_seg = off; off += al((n + 1) * 4);
    const uint64_t o_lst = off; off += al(na * 8); const uint64_t o_dset = off; off += al(na * 4);
    const uint64_t o_lsz = off; off += al((n + 1) * 4); const uint64_t o_lsz64 = off; off += al((n + 1) * 8);
    const uint64_t o_indptr = off; off += al((n + 1) * 8);
    if (int rc = c->lbuf.reserve(off)) return rc;
    unsigned char *B = c->lbuf.as<unsigned char>();
    uint32_t *kA = (uint32_t *)(B + o_kA), *kB = (uint32_t *)(B + o_kB), *iA = (uint32_t *)(B + o_iA), *iB = (uint32_t *)(B + o_iB);
    int64_t *offs = (int64_t *)(B + o_offs);
    uint32_t *cand = (uint32_t *)(B + o_cand), *cnt = (uint32_t *)(B + o_cnt), *ncand = (uint32_t *)(B + o_nc);
    uint32_t *alA = (uint32_t *)(B + o_alA), *alB = (uint32_t *)(B + o_alB);
    uint64_t *apA = (uint64_t *)(B + o_apA), *apB = (uint64_t *)(B + o_apB);
    uint32_t *seg = (uint32_t *)(B + o_seg), *dset = (uint32_t *)(B + o_dset), *lsz = (uint32_t *)(B + o_lsz);
    d2g::Nb *lst = (d2g::Nb *)(B + o_lst);
    uint64_t *lsz64 = (uint64_t *)(B + o_lsz64), *indptr_d = (uint64_t *)(B + o_indptr);
    // 1. keys + 2. per-table sort (chunks of tables so one segmented sort stays below 2^30 items)
    const uint32_t tchunk = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(ntab, (1ULL << 30) / n));
    size_t tb = 0;
    for (uint32_t t0 = 0; t0 < ntab; t0 += tchunk) {
        const uint32_t nt = std::min(tchunk, ntab - t0);
        std::vector<int64_t> ho(nt + 1);
        for (uint32_t t = 0; t <= nt; ++t) ho[t] = (int64_t)((uint64_t)t * n);
        CU(cudaMemcpyAsync(offs, ho.data(), (nt + 1) * 8, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
        const uint64_t items = (uint64_t)nt * n;
        d2g::lsh_keys_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(regs_d, n, S, t0, nt, kA + (uint64_t)t0 * n, iA + (uint64_t)t0 * n, geom);
        size_t need = 0;
        cub::DeviceSegmentedRadixSort::SortPairs(nullptr, need, kA, kB, iA, iB, (int)items, (int)nt, offs, offs + 1, 0, 32, st);
        if (need > tb) { tb = need; if (int rc = c->wtmp.reserve(tb + 256)) return rc; }
        size_t tbytes = tb;
        KernelTimer kt(c, D2G_T_SORT);
        CU(cub::DeviceSegmentedRadixSort::SortPairs(c->wtmp.p, tbytes, kA + (uint64_t)t0 * n, kB + (uint64_t)t0 * n, iA + (uint64_t)t0 * n, iB + (uint64_t)t0 * n,
                                                   (int)items, (int)nt, offs, offs + 1, 0, 32, st));
        c->launches += 6;
    }
    // 3. ordered candidate scan, one warp per query
    {
        int wpb = 4;                                                  // queries (warps) per CTA; fewer when the candidate lists are long
        while (wpb > 1 && (size_t)wpb * 2 * maxcand * 4 > 96 * 1024) wpb >>= 1;
        if ((size_t)wpb * 2 * maxcand * 4 > 200 * 1024) return fail(D2G_EUNSUPPORTED, "candidate lists of %u entries do not fit shared memory", maxcand);
        const size_t smem = (size_t)wpb * 2 * maxcand * 4;
        CU(cudaFuncSetAttribute(d2g::lsh_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KernelTimer kt(c, D2G_T_CMP);
        d2g::lsh_query_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, smem, st>>>(regs_d, n, S, kB, iB, maxcand, cand, cnt, ncand, geom);
        c->launches++;
        CU(cudaGetLastError());
    }
    // 4. arrivals, stable sort by destination list, segment starts, replay
    d2g::lsh_arrivals_kernel<<<(unsigned)((n * maxcand + 255) / 256), 256, 0, st>>>(cand, cnt, ncand, n, maxcand, (uint32_t)x0, (uint32_t)x1, alA, apA);
    {
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, alA, alB, apA, apB, (int)na, 0, 32, st);
        if (need > tb) { tb = need; if (int rc = c->wtmp.reserve(tb + 256)) return rc; }
        size_t tbytes = tb;
        CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, alA, alB, apA, apB, (int)na, 0, 32, st));
    }
    d2g::lsh_segments_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(alB, na, n, seg);
    if (!threshold) {
        const bool warp_ok = maxcand <= (uint32_t)d2g::LSH_REPLAY_DCAP;
        if (warp_ok) d2g::lsh_replay_warp_kernel<<<(unsigned)((n + d2g::LSH_REPLAY_WARPS - 1) / d2g::LSH_REPLAY_WARPS), d2g::LSH_REPLAY_WARPS * 32, 0, st>>>(apB, seg, n, maxcand, lst, lsz);
        else d2g::lsh_replay_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(apB, seg, n, maxcand, 0u, lst, dset, lsz);
    }
    c->launches += 8;
    uint32_t h_total = 0;
    CU(cudaMemcpyAsync(&h_total, seg + n, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (threshold && h_total) {
        // first arrival of every id per list, ordered by (-hits, id): two segmented sorts over the arrivals (see lsh_kernels.cuh).
        // Buffers: apA / alA are free after the sort by list; lst / dset are not used yet.
        uint64_t *k1 = apA, *k1s = reinterpret_cast<uint64_t *>(lst); uint32_t *cd = alA, *cds = dset;
        const unsigned gt = (h_total + 255) / 256;
        d2g::lsh_thr_key1_kernel<<<gt, 256, 0, st>>>(alB, apB, seg, h_total, k1, cd);
        size_t need = 0, need2 = 0;
        cub::DeviceSegmentedRadixSort::SortPairs(nullptr, need, k1, k1s, cd, cds, (int)h_total, (int)n, seg, seg + 1, 0, 64, st);
        cub::DeviceSegmentedRadixSort::SortKeys(nullptr, need2, k1, k1s, (int)h_total, (int)n, seg, seg + 1, 0, 64, st);
        need = std::max(need, need2);
        if (need > tb) { tb = need; if (int rc = c->wtmp.reserve(tb + 256)) return rc; }
        size_t tbytes = tb;
        CU(cub::DeviceSegmentedRadixSort::SortPairs(c->wtmp.p, tbytes, k1, k1s, cd, cds, (int)h_total, (int)n, seg, seg + 1, 0, 64, st));
        d2g::lsh_thr_key2_kernel<<<gt, 256, 0, st>>>(alB, k1s, cds, seg, h_total, k1);
        tbytes = tb;
        CU(cub::DeviceSegmentedRadixSort::SortKeys(c->wtmp.p, tbytes, k1, k1s, (int)h_total, (int)n, seg, seg + 1, 0, 64, st));
        d2g::lsh_thr_lists_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(seg, n, k1s, lsz);
        c->launches += 9;
        CU(cudaGetLastError());
    } else if (threshold) CU(cudaMemsetAsync(lsz, 0, n * 4, st));
    if (getenv("D2G_DEBUG")) {
        std::vector<uint32_t> hs(n);
        cudaMemcpy(hs.data(), lsz, n * 4, cudaMemcpyDeviceToHost);
        uint64_t sum = 0, big256 = 0, big1024 = 0; uint32_t mx = 0;
        for (uint32_t v : hs) { sum += v; mx = std::max(mx, v); big256 += v > 256; big1024 += v > 1024; }
        fprintf(stderr, "[d2g] topk: %llu lists, %u arrival slots, list entries before refinement: total %llu, mean %.1f, max %u, >256: %llu, >1024: %llu\n",
                (unsigned long long)n, h_total, (unsigned long long)sum, (double)sum / n, mx, (unsigned long long)big256, (unsigned long long)big1024);
    }
    // 5. refine + trim
    d2g::CmpConsts k;
    if (int rc = make_consts(c, p, &k)) return rc;
    const int is_dist = !(p->measure == D2G_UNION_SIZE || p->measure == D2G_INTERSECTION || p->measure == D2G_SIMILARITY || p->measure == D2G_CONTAINMENT);
    const float mult = is_dist ? 1.f : -1.f;
    c->stats[D2G_STAT_REFINED] = 0;
    if (h_total) {
        {   // entries the replay kept = what refinement compares
            std::vector<uint32_t> hs(n);
            CU(cudaMemcpyAsync(hs.data(), lsz, n * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            uint64_t sum = 0; for (uint32_t v : hs) sum += v;
            c->stats[D2G_STAT_REFINED] = sum;
        }
        KernelTimer kt(c, D2G_T_LSH_REFINE);
        const uint64_t threads = (uint64_t)h_total * 32;
        const bool gtlt = p->cmp_kind == D2G_CMP_GTLT || p->cmp_kind == D2G_CMP_SS_COMPRESSED;
        if (gtlt) d2g::lsh_refine_kernel<0><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(cmp_regs_d, cards_d, n, seg, lsz, lst, k, mult);
        else d2g::lsh_refine_kernel<1><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(cmp_regs_d, cards_d, n, seg, lsz, lst, k, mult);
        c->launches++;
    }
    const uint32_t keep = threshold ? 0xFFFFFFFFu : (uint32_t)topk;
    if (threshold) {       // refine.cpp:45-68: threshold + 20-consecutive-failures walk in (-hits, id) order; the trim kernels then only sort
        d2g::lsh_thr_filter_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(seg, n, min_similarity, is_dist, lst, lsz);
        c->launches++;
    }
    // short lists first: a list the second kernel has trimmed (> LSH_TRIM_CAP entries before) must not be seen as short afterwards
    d2g::lsh_trim_kernel<<<(unsigned)((n + d2g::LSH_TRIM_WARPS - 1) / d2g::LSH_TRIM_WARPS), d2g::LSH_TRIM_WARPS * 32, 0, st>>>(seg, n, keep, is_dist, lst, lsz, threshold ? 1 : 0);
    d2g::lsh_trim_big_kernel<<<(unsigned)n, d2g::LSH_TRIM_BIG_THREADS, d2g::LSH_TRIM_BIG_CAP * 8, st>>>(seg, n, keep, is_dist, lst, lsz, threshold ? 1 : 0);
    c->launches += 2;
    // 6. CSR: indptr = exclusive scan of list sizes
    {
        // widen to u64 on the host side of the scan: sizes are small, sum may exceed 2^32
        std::vector<uint32_t> hs(n);
        CU(cudaMemcpyAsync(hs.data(), lsz, n * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        std::vector<uint64_t> full(n + 1, 0);                 // lists outside [x0, x1) are empty
        for (uint64_t i = 0; i < n; ++i) full[i + 1] = full[i] + ((i >= x0 && i < x1) ? hs[i] : 0);
        for (uint64_t i = x0; i <= x1; ++i) indptr_out[i - x0] = full[i];
        CU(cudaMemcpyAsync(indptr_d, full.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
    }
    (void)lsz64;
    const uint64_t nnz = indptr_out[x1 - x0];
    uint32_t *hidx = (uint32_t *)malloc((nnz + 1) * 4); float *hval = (float *)malloc((nnz + 1) * 4);
    if (!hidx || !hval) { free(hidx); free(hval); return fail(D2G_ENOMEM, "malloc failed for %llu neighbours", (unsigned long long)nnz); }
    if (nnz) {
        // reuse the arrival buffers for the CSR arrays
        uint32_t *idx_d = alA; float *val_d = (float *)alB;
        d2g::lsh_csr_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(seg, lsz, indptr_d, n, lst, idx_d, val_d);
        c->launches++;
        CU(cudaMemcpyAsync(hidx, idx_d, nnz * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hval, val_d, nnz * 4, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    *idx_out = hidx; *val_out = hval;
    return D2G_OK;
}
