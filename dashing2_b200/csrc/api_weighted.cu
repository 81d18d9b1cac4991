// api_weighted.cu -- counting sketches (BagMinHash / ProbMinHash), the one-permutation sketch with --count-threshold and the exact
// distinct-k-mer count: everything that goes emit -> radix sort by (entity, value) -> runs.
#include "api_sketch_launch.h"
#include "fss_kernels.cuh"
#include "weighted_kernels.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace {
// bits of the entity key a stable LSD pass has to sort: entities are 0 .. n_ent-1 and empty slots carry 0xFFFFFFFF, whose low b bits (all ones)
// exceed every entity as soon as 2^b > n_ent -- one 8-bit pass for up to 255 entities instead of four
inline int ent_sort_bits(uint32_t n_ent) { int b = 1; while (b < 32 && (1ull << b) <= (uint64_t)n_ent) ++b; return b; }
// BagMinHash / ProbMinHash (see weighted_kernels.cuh): emit -> sort -> run-length encode -> sketch -> verify loop.
// sig_d [n_ent][S], card_d [n_ent].  Synchronises (needs the number of distinct elements and the redo count).
__global__ void weighted_finalize_kernel(const uint64_t *keys, const unsigned long long *wsum, uint32_t n_ent, uint32_t m, double *sig, double *card) {
    const uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (sig && e < (uint64_t)n_ent * m) sig[e] = d2g::dunkey(keys[e]);
    if (card && e < n_ent) card[e] = (double)wsum[e];
}
} // namespace

int launch_weighted(d2g_ctx *c, const d2g_sketch_params *p, const d2g::PackedSeq &seq_d, const uint64_t *rec_off_d, const uint32_t *rec_ent_d,
                    uint64_t n_rec, uint32_t n_ent, uint64_t total_len, double *sig_d, double *card_d, uint64_t *ids_d) {
    const uint32_t m = p->sketchsize;
    const uint64_t n = total_len, nreg = (uint64_t)n_ent * m;
    if (n >= 0xFFFFFFF0ULL) return fail(D2G_EINVAL, "counting sketches: at most 2^32 bases per batch (got %llu)", (unsigned long long)n);
    const bool windowed = p->w > p->k;
    const uint64_t ovf_cap = 1ULL << 20;
    auto al = [](uint64_t b) { return (b + 255) / 256 * 256; };
    // wbuf layout
    uint64_t off = 0;
    const uint64_t o_hvA = off; off += al(n * 8 + 8);
    const uint64_t o_hvB = off; off += al(n * 8 + 8);
    const uint64_t o_entA = off; off += al(n * 4 + 4);
    const uint64_t o_entB = off; off += al(n * 4 + 4);
    const uint64_t o_flag = off; off += al(n * 4 + 4);
    const uint64_t o_excl = off; off += al(n * 4 + 4);
    const uint64_t o_pos = off; off += al(n * 4 + 4);
    const uint64_t cssize = p->countsketch_size;
    const uint64_t o_wts = off; off += cssize ? al(n * 4 + 4) : 0;
    const uint64_t o_keys = off; off += al(nreg * 8);
    const uint64_t o_wsum = off; off += al((uint64_t)n_ent * 8);
    const uint64_t o_T = off; off += al((uint64_t)n_ent * 8);
    const uint64_t o_state = off; off += al((uint64_t)n_ent * 4);
    const uint64_t o_misc = off; off += 256;            // n_valid, ovf_count, n_redo, error
    const uint64_t o_ovf = off; off += al(ovf_cap * 8);
    if (int rc = c->wbuf.reserve(off)) return rc;
    unsigned char *B = c->wbuf.as<unsigned char>();
    uint64_t *hvA = (uint64_t *)(B + o_hvA), *hvB = (uint64_t *)(B + o_hvB);
    uint32_t *entA = (uint32_t *)(B + o_entA), *entB = (uint32_t *)(B + o_entB);
    uint32_t *flag = (uint32_t *)(B + o_flag), *excl = (uint32_t *)(B + o_excl), *pos = (uint32_t *)(B + o_pos);
    uint32_t *wts = cssize ? (uint32_t *)(B + o_wts) : nullptr;
    const int id_shift = cssize ? 1 : 0;
    uint64_t *keys = (uint64_t *)(B + o_keys);
    unsigned long long *wsum = (unsigned long long *)(B + o_wsum);
    double *T = (double *)(B + o_T);
    uint32_t *state = (uint32_t *)(B + o_state);
    unsigned long long *n_valid = (unsigned long long *)(B + o_misc), *ovf_count = n_valid + 1;
    unsigned int *n_redo = (unsigned int *)(n_valid + 2), *error = n_redo + 1;
    uint64_t *ovf = (uint64_t *)(B + o_ovf);
    cudaStream_t st = c->stream;
    CU(cudaMemsetAsync(hvA, 0xFF, n * 8 + 8, st));
    CU(cudaMemsetAsync(entA, 0xFF, n * 4 + 4, st));
    CU(cudaMemsetAsync(B + o_wsum, 0, al((uint64_t)n_ent * 8), st));
    CU(cudaMemsetAsync(B + o_misc, 0, 256, st));
    fill_u64_kernel<<<(unsigned)std::min<uint64_t>((nreg + 255) / 256, 4096), 256, 0, st>>>(keys, nreg, d2g::FSS_KEY_EMPTY);
    c->launches++;
    uint64_t nu = 0;
    if (n && n_rec) {
        d2g::SketchArgs a = make_sketch_args(c, p, seq_d, rec_off_d, rec_ent_d, n_rec, total_len, 0, SketchRange{0, total_len, 0});
        if (a.span >= 0xFFFFFFFFULL) return fail(D2G_EINVAL, "span too large");
        d2g::EmitConsumer::Params ep{hvA, entA, a.span};
        if (int rc = launch_sketch<d2g::EmitConsumer>(c, a, ep, windowed, D2G_T_SKETCH_MAIN)) return rc;
        if (cssize) { d2g::cs_key_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hvA, entA, n, cssize); c->launches++; }
        // sort by (entity, value): LSD radix -- value first, then a stable pass over the entity
        const int ebits = ent_sort_bits(n_ent);   // the entity pass sorts only the bits entities use (the 0xFFFFFFFF sentinel still sorts last)
        size_t t1 = 0, t2 = 0, t3 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, t1, hvA, hvB, entA, entB, n, 0, 64, st);
        cub::DeviceRadixSort::SortPairs(nullptr, t2, entB, entA, hvB, hvA, n, 0, ebits, st);
        cub::DeviceScan::ExclusiveSum(nullptr, t3, flag, excl, n, st);
        const size_t tb = std::max(t1, std::max(t2, t3));
        if (int rc = c->wtmp.reserve(tb + 256)) return rc;
        {
        KernelTimer kt(c, D2G_T_SORT);
        size_t tbytes = tb;
        CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, hvA, hvB, entA, entB, n, 0, 64, st));
        tbytes = tb;
        CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, entB, entA, hvB, hvA, n, 0, ebits, st));
        c->launches += 2 * 9;
        const unsigned gb = (unsigned)((n + 255) / 256);
        d2g::rle_flag_kernel<<<gb, 256, 0, st>>>(hvA, entA, n, flag, id_shift);
        tbytes = tb;
        CU(cub::DeviceScan::ExclusiveSum(c->wtmp.p, tbytes, flag, excl, n, st));
        d2g::rle_scatter_kernel<<<gb, 256, 0, st>>>(flag, excl, entA, n, pos, n_valid);
        c->launches += 3;
        }
        uint32_t h_last[2] = {0, 0};
        CU(cudaMemcpyAsync(&h_last[0], excl + (n - 1), 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&h_last[1], flag + (n - 1), 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        nu = (uint64_t)h_last[0] + h_last[1];
    }
    // exact counts pass with count > threshold (counter.h:123); count-sketch buckets with |count| >= threshold (:135), and never with weight 0
    const double threshold = cssize ? (p->count_threshold >= 1 ? (double)p->count_threshold - 0.5 : 0.) : (double)p->count_threshold;
    if (nu) {
        const unsigned gu = (unsigned)((nu + 127) / 128);
        if (cssize) { d2g::cs_run_weight_kernel<<<(unsigned)((nu + 255) / 256), 256, 0, st>>>(hvA, pos, nu, n_valid, wts); c->launches++; }
        d2g::weight_sum_kernel<<<(unsigned)((nu + 255) / 256), 256, 0, st>>>(entA, pos, nu, n_valid, threshold, wsum, wts);
        d2g::weighted_guess_kernel<<<(n_ent + 255) / 256, 256, 0, st>>>(wsum, n_ent, m, T, state);
        c->launches += 2;
        unsigned long long *bmh_next = n_valid + 4;                  // in the misc block
        const unsigned bmh_grid = (unsigned)std::min<uint64_t>((nu + 127) / 128, (uint64_t)c->sm_count * 12);
        d2g::WeightedArgs wa{hvA, entA, pos, nu, n_valid, threshold, m, T, state, keys, ovf, ovf_count, ovf_cap, error, wts, id_shift, nullptr, bmh_next};
        d2g::TexpConsts tc{};
        if (p->mode == D2G_MODE_PROBMINHASH) {   // bmh.h:490-502
            const long double lambda = log1pl(1.L / (m - 1));
            const long double c1 = (expl(lambda) - 1.L) / lambda, c2 = logl(2.L / (1.L + expl(-lambda))) / lambda, c3 = (1.L - expl(-lambda)) / lambda;
            tc = d2g::TexpConsts{(double)lambda, (double)c1, (double)c2, (double)c3, (double)(c1 * lambda)};
        }
        uint64_t nslots = std::min<uint64_t>(4096, (256ULL << 20) / ((uint64_t)m * 8));
        nslots = std::max<uint64_t>(32, nslots / 32 * 32);
        if (p->mode == D2G_MODE_PROBMINHASH) { if (int rc = c->aux2.reserve(nslots * 2ULL * m * 4)) return rc; }
        for (int round = 0; round < 40; ++round) {
            CU(cudaMemsetAsync(n_redo, 0, 4, st));
            CU(cudaMemsetAsync(ovf_count, 0, 8, st));
            {
                KernelTimer kt(c, D2G_T_SKETCH_BOOT);
                if (p->mode == D2G_MODE_PROBMINHASH) {
                    d2g::pmh_kernel<<<gu, 128, 0, st>>>(wa, tc);
                    CU(cudaMemsetAsync(c->aux2.p, 0, nslots * 2ULL * m * 4, st));
                    d2g::pmh_longwalk_kernel<<<(unsigned)(nslots / 32), 32, 0, st>>>(wa, tc, c->aux2.as<uint32_t>());
                    c->launches += 2;
                } else {
                    CU(cudaMemsetAsync(bmh_next, 0, 8, st)); d2g::bmh_kernel<<<bmh_grid, 128, 0, st>>>(wa);
                    c->launches++;
                }
            }
            d2g::weighted_verify_kernel<<<n_ent, 256, 0, st>>>(keys, m, T, state, n_redo);
            c->launches++;
            unsigned int h[2] = {0, 0};
            CU(cudaMemcpyAsync(h, n_redo, 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            CU(cudaGetLastError());
            if (h[1]) return fail(D2G_EUNSUPPORTED, "weighted sketch: device work queue overflow (code %u); split the batch", h[1]);
            if (!h[0]) break;
            if (round == 39) return fail(D2G_ECUDA, "weighted sketch: bound verification did not converge");
        }
        if (ids_d) {   // --save-kmers (bmh.h: ids_[idx] = id where a register is lowered): replay every element once more against the final
                       // registers with the verified bounds; the element whose point equals a register is the one that set it
            CU(cudaMemsetAsync(ids_d, 0, nreg * 8, st));
            CU(cudaMemsetAsync(ovf_count, 0, 8, st));
            wa.ids = ids_d;
            if (p->mode == D2G_MODE_PROBMINHASH) {
                d2g::pmh_kernel<<<gu, 128, 0, st>>>(wa, tc);
                CU(cudaMemsetAsync(c->aux2.p, 0, nslots * 2ULL * m * 4, st));
                d2g::pmh_longwalk_kernel<<<(unsigned)(nslots / 32), 32, 0, st>>>(wa, tc, c->aux2.as<uint32_t>());
                c->launches += 2;
            } else {
                CU(cudaMemsetAsync(bmh_next, 0, 8, st)); d2g::bmh_kernel<<<bmh_grid, 128, 0, st>>>(wa);
                c->launches++;
            }
            unsigned int h[2] = {0, 0};
            CU(cudaMemcpyAsync(h, n_redo, 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            CU(cudaGetLastError());
            if (h[1]) return fail(D2G_EUNSUPPORTED, "weighted sketch: device work queue overflow in the ids pass (code %u); split the batch", h[1]);
        }
    } else if (ids_d && nreg) CU(cudaMemsetAsync(ids_d, 0, nreg * 8, st));
    const uint64_t nthreads = std::max<uint64_t>(nreg, n_ent);
    weighted_finalize_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(keys, wsum, n_ent, m, sig_d, card_d);
    c->launches++;
    CU(cudaGetLastError());
    return D2G_OK;
}

// ---- One-permutation MinHash with --count-threshold c > 1 (oph.h:188-205) ----------------------------
// The reference promotes a candidate of a bucket once it has been seen c times while it is below the bucket's register, and then
// drops the candidates above it; candidates below survive.  The register therefore ends as the minimum over the ids of the bucket seen
// at least c times, whatever the order -- only the multiplicity field (not part of the signature) depends on the order.  Device:
// emit every k-mer / window (as the counting sketches do), sort by (entity, value), and let the head of every run that is at least c
// long update its bucket.
namespace {
__global__ void opmh_mincount_kernel(const uint64_t *hv, const uint32_t *ent, uint64_t n, uint32_t c, uint64_t *regs, d2g::FastMod32 fm, uint32_t m) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t e = ent[i];
    if (e == 0xFFFFFFFFu) return;
    const uint64_t v = hv[i];
    if (i && ent[i - 1] == e && hv[i - 1] == v) return;              // not a run head
    const uint64_t last = i + c - 1;                                  // sorted: the run is at least c long iff element i+c-1 still belongs to it
    if (last >= n || ent[last] != e || hv[last] != v) return;
    const uint64_t id = d2g::dhash(v);                                // oph.h:178
    const uint32_t idx = d2g::fastmod32((uint32_t)id, fm);            // oph.h:184
    atomicMin(reinterpret_cast<unsigned long long *>(regs + (uint64_t)e * m + idx), (unsigned long long)id);
}
} // namespace

int launch_opmh_mincount(d2g_ctx *c, const d2g_sketch_params *p, const d2g::PackedSeq &seq_d, const uint64_t *rec_off_d, const uint32_t *rec_ent_d,
                         uint64_t n_rec, uint32_t n_ent, uint64_t total_len, uint64_t *regs_d) {
    const uint32_t m = d2g_opmh_m(p->sketchsize);
    const uint64_t n = total_len, nreg = (uint64_t)n_ent * m;
    cudaStream_t st = c->stream;
    if (nreg) { fill_u64_kernel<<<(unsigned)std::min<uint64_t>((nreg + 255) / 256, 4096), 256, 0, st>>>(regs_d, nreg, ~0ULL); c->launches++; }
    if (!n || !n_rec) return D2G_OK;
    if (n >= 0xFFFFFFF0ULL) return fail(D2G_EINVAL, "--count-threshold: at most 2^32 bases per batch (got %llu)", (unsigned long long)n);
    auto al = [](uint64_t b) { return (b + 255) / 256 * 256; };
    uint64_t off = 0;
    const uint64_t o_hvA = off; off += al(n * 8 + 8);
    const uint64_t o_hvB = off; off += al(n * 8 + 8);
    const uint64_t o_entA = off; off += al(n * 4 + 4);
    const uint64_t o_entB = off; off += al(n * 4 + 4);
    if (int rc = c->wbuf.reserve(off)) return rc;
    unsigned char *B = c->wbuf.as<unsigned char>();
    uint64_t *hvA = (uint64_t *)(B + o_hvA), *hvB = (uint64_t *)(B + o_hvB);
    uint32_t *entA = (uint32_t *)(B + o_entA), *entB = (uint32_t *)(B + o_entB);
    CU(cudaMemsetAsync(hvA, 0xFF, n * 8 + 8, st));
    CU(cudaMemsetAsync(entA, 0xFF, n * 4 + 4, st));
    d2g::SketchArgs a = make_sketch_args(c, p, seq_d, rec_off_d, rec_ent_d, n_rec, total_len, 0, SketchRange{0, total_len, 0});
    if (a.span >= 0xFFFFFFFFULL) return fail(D2G_EINVAL, "span too large");
    d2g::EmitConsumer::Params ep{hvA, entA, a.span};
    if (int rc = launch_sketch<d2g::EmitConsumer>(c, a, ep, p->w > p->k, D2G_T_SKETCH_MAIN)) return rc;
    const int ebits = ent_sort_bits(n_ent);   // the entity pass sorts only the bits entities use (the 0xFFFFFFFF sentinel still sorts last)
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, hvA, hvB, entA, entB, n, 0, 64, st);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, entB, entA, hvB, hvA, n, 0, ebits, st);
    const size_t tb = std::max(t1, t2);
    if (int rc = c->wtmp.reserve(tb + 256)) return rc;
    size_t tbytes = tb;
    CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, hvA, hvB, entA, entB, n, 0, 64, st));
    tbytes = tb;
    CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, entB, entA, hvB, hvA, n, 0, ebits, st));
    c->launches += 2 * 9;
    opmh_mincount_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hvA, entA, n, p->count_threshold, regs_d, d2g::make_fastmod32(m), m);
    c->launches++;
    CU(cudaGetLastError());
    return D2G_OK;
}

// ---- exact distinct k-mers per entity (the --parse-by-seq small-cardinality fallback) -----------------
namespace {
// sorted (entity, value) stream: one count per run head, aggregated per warp and entity
__global__ void distinct_count_kernel(const uint64_t *hv, const uint32_t *ent, uint64_t n, unsigned long long *cnt) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t e = 0xFFFFFFFFu; bool head = false;
    if (i < n) { e = ent[i]; head = e != 0xFFFFFFFFu && (i == 0 || ent[i - 1] != e || hv[i - 1] != hv[i]); }
    unsigned todo = __ballot_sync(0xffffffffu, head);
    while (todo) {
        const int leader = __ffs((int)todo) - 1;
        const uint32_t le = __shfl_sync(0xffffffffu, e, leader);
        const unsigned same = __ballot_sync(0xffffffffu, head && e == le);
        if ((int)(threadIdx.x & 31) == leader) atomicAdd(cnt + le, (unsigned long long)__popc(same));
        todo &= ~same;
    }
}
}

extern "C" int d2g_distinct_kmers(d2g_ctx *c, const d2g_sketch_params *p, const char *seq, const uint64_t *rec_off,
                                  const uint32_t *rec_entity, uint64_t n_rec, uint32_t n_entities, uint64_t *distinct_out) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_sketch_params(p)) return rc;
    if (!distinct_out && n_entities) return fail(D2G_EINVAL, "null output");
    if (n_rec && (!rec_off || !rec_entity)) return fail(D2G_EINVAL, "null record tables");
    CU(cudaSetDevice(c->device));
    const uint64_t n_bases = n_rec ? rec_off[n_rec] : 0;
    if (n_bases && !seq) return fail(D2G_EINVAL, "null sequence buffer");
    if (n_rec && rec_off[0] != 0) return fail(D2G_EINVAL, "rec_off[0] must be 0");
    if (n_bases >= 0xFFFFFFF0ULL) return fail(D2G_EINVAL, "distinct k-mers: at most 2^32 bases per call (got %llu)", (unsigned long long)n_bases);
    for (uint64_t r = 0; r < n_rec; ++r) {
        if (rec_off[r + 1] < rec_off[r]) return fail(D2G_EINVAL, "rec_off not monotone at %llu", (unsigned long long)r);
        if (rec_entity[r] >= n_entities) return fail(D2G_EINVAL, "rec_entity[%llu]=%u >= n_entities", (unsigned long long)r, rec_entity[r]);
        if (r && rec_entity[r] < rec_entity[r - 1]) return fail(D2G_EINVAL, "rec_entity must be non-decreasing");
    }
    for (uint32_t e = 0; e < n_entities; ++e) distinct_out[e] = 0;
    if (!n_bases || !n_rec || !n_entities) return D2G_OK;
    if (int rc = c->seq.reserve(n_bases + 64)) return rc;
    if (int rc = c->recoff.reserve((n_rec + 1) * 8)) return rc;
    if (int rc = c->recent.reserve((n_rec + 1) * 4)) return rc;
    const uint64_t nw = d2g::packed_words(n_bases);
    if (int rc = c->pcodes.reserve(nw * 8)) return rc;
    if (int rc = c->pmask.reserve(nw * 4)) return rc;
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->seq.p, seq, n_bases, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->recoff.p, rec_off, (n_rec + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->recent.p, rec_entity, n_rec * 4, cudaMemcpyHostToDevice, st));
    if (int rc = d2g_pack_dev(c, c->seq.as<char>(), n_bases, c->pcodes.as<uint64_t>(), c->pmask.as<uint32_t>())) return rc;
    d2g::PackedSeq seq_d{c->pcodes.as<uint64_t>(), c->pmask.as<uint32_t>()};
    const uint64_t *off_d = c->recoff.as<uint64_t>();
    uint64_t n = n_bases;
    StreamView sv;
    if (is_stream_mode(p)) {                                        // element streams: slots are item regions (stream_kernels.cuh)
        if (int rc = prepare_stream(c, p, seq_d, c->seq.as<uint8_t>(), off_d, n_rec, n_bases, &sv)) return rc;
        p = &sv.p; seq_d = sv.seq; off_d = sv.rec_off_d; n = sv.total_len;
        if (n >= 0xFFFFFFF0ULL) return fail(D2G_EINVAL, "distinct k-mers: at most 2^31 bases per call for this k-mer stream");
    }
    auto al = [](uint64_t b) { return (b + 255) / 256 * 256; };
    uint64_t off = 0;
    const uint64_t o_hvA = off; off += al(n * 8 + 8);
    const uint64_t o_hvB = off; off += al(n * 8 + 8);
    const uint64_t o_entA = off; off += al(n * 4 + 4);
    const uint64_t o_entB = off; off += al(n * 4 + 4);
    const uint64_t o_cnt = off; off += al((uint64_t)n_entities * 8);
    if (int rc = c->wbuf.reserve(off)) return rc;
    unsigned char *B = c->wbuf.as<unsigned char>();
    uint64_t *hvA = (uint64_t *)(B + o_hvA), *hvB = (uint64_t *)(B + o_hvB);
    uint32_t *entA = (uint32_t *)(B + o_entA), *entB = (uint32_t *)(B + o_entB);
    unsigned long long *cnt = (unsigned long long *)(B + o_cnt);
    CU(cudaMemsetAsync(hvA, 0xFF, n * 8 + 8, st));
    CU(cudaMemsetAsync(entA, 0xFF, n * 4 + 4, st));
    CU(cudaMemsetAsync(cnt, 0, (uint64_t)n_entities * 8, st));
    d2g::SketchArgs a = make_sketch_args(c, p, seq_d, off_d, c->recent.as<uint32_t>(), n_rec, n, 0, SketchRange{0, n, 0});
    if (a.span >= 0xFFFFFFFFULL) return fail(D2G_EINVAL, "span too large");
    d2g::EmitConsumer::Params ep{hvA, entA, a.span};
    if (int rc = launch_sketch<d2g::EmitConsumer>(c, a, ep, p->w > p->k, D2G_T_SKETCH_MAIN)) return rc;
    const int ebits = ent_sort_bits(n_entities);   // the entity pass sorts only the bits entities use (the 0xFFFFFFFF sentinel still sorts last)
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, hvA, hvB, entA, entB, n, 0, 64, st);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, entB, entA, hvB, hvA, n, 0, ebits, st);
    const size_t tb = std::max(t1, t2);
    if (int rc = c->wtmp.reserve(tb + 256)) return rc;
    size_t tbytes = tb;
    CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, hvA, hvB, entA, entB, n, 0, 64, st));
    tbytes = tb;
    CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, entB, entA, hvB, hvA, n, 0, ebits, st));
    c->launches += 2 * 9;
    distinct_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hvA, entA, n, cnt);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(distinct_out, cnt, (uint64_t)n_entities * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return D2G_OK;
}

// ---- --filterset (src/d2.cpp:45-98, src/filterset.h): the sorted set of hashed k-mers that later sketch calls skip -----------------
namespace {
__global__ void count_valid_kernel(const uint32_t *ent_sorted, uint64_t n, unsigned long long *n_valid) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (ent_sorted[i] != 0xFFFFFFFFu && (i + 1 == n || ent_sorted[i + 1] == 0xFFFFFFFFu)) *n_valid = i + 1;
}
}

extern "C" int d2g_clear_filterset(d2g_ctx *c) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    c->filter_n = 0;
    return D2G_OK;
}

extern "C" int d2g_set_filterset_values(d2g_ctx *c, const uint64_t *values, uint64_t n) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (n && !values) return fail(D2G_EINVAL, "null values");
    CU(cudaSetDevice(c->device));
    c->filter_n = 0;
    if (!n) return D2G_OK;
    std::vector<uint64_t> v(values, values + n);
    std::sort(v.begin(), v.end());
    if (int rc = c->filter.reserve(n * 8)) return rc;
    CU(cudaMemcpyAsync(c->filter.p, v.data(), n * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->filter_n = n;
    return D2G_OK;
}

extern "C" int d2g_set_filterset(d2g_ctx *c, const d2g_sketch_params *p, const char *seq, const uint64_t *rec_off, uint64_t n_rec, uint64_t *n_out) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_sketch_params(p)) return rc;
    if (n_rec && !rec_off) return fail(D2G_EINVAL, "null record table");
    CU(cudaSetDevice(c->device));
    c->filter_n = 0;                                                 // the set is built from unfiltered k-mers
    if (n_out) *n_out = 0;
    const uint64_t n_bases = n_rec ? rec_off[n_rec] : 0;
    if (!n_bases) return D2G_OK;
    if (!seq) return fail(D2G_EINVAL, "null sequence buffer");
    if (rec_off[0] != 0) return fail(D2G_EINVAL, "rec_off[0] must be 0");
    if (n_bases >= 0x7FFFFFF0ULL) return fail(D2G_EINVAL, "filter set: at most 2^31 bases per call (got %llu)", (unsigned long long)n_bases);
    for (uint64_t r = 0; r < n_rec; ++r) if (rec_off[r + 1] < rec_off[r]) return fail(D2G_EINVAL, "rec_off not monotone at %llu", (unsigned long long)r);
    std::vector<uint32_t> ent(n_rec, 0u);
    if (int rc = c->seq.reserve(n_bases + 64)) return rc;
    if (int rc = c->recoff.reserve((n_rec + 1) * 8)) return rc;
    if (int rc = c->recent.reserve((n_rec + 1) * 4)) return rc;
    const uint64_t nw = d2g::packed_words(n_bases);
    if (int rc = c->pcodes.reserve(nw * 8)) return rc;
    if (int rc = c->pmask.reserve(nw * 4)) return rc;
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->seq.p, seq, n_bases, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->recoff.p, rec_off, (n_rec + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->recent.p, ent.data(), n_rec * 4, cudaMemcpyHostToDevice, st));
    if (int rc = d2g_pack_dev(c, c->seq.as<char>(), n_bases, c->pcodes.as<uint64_t>(), c->pmask.as<uint32_t>())) return rc;
    d2g::PackedSeq seq_d{c->pcodes.as<uint64_t>(), c->pmask.as<uint32_t>()};
    const uint64_t *off_d = c->recoff.as<uint64_t>();
    uint64_t n = n_bases;
    StreamView sv;
    if (is_stream_mode(p)) {
        if (int rc = prepare_stream(c, p, seq_d, c->seq.as<uint8_t>(), off_d, n_rec, n_bases, &sv)) return rc;
        p = &sv.p; seq_d = sv.seq; off_d = sv.rec_off_d; n = sv.total_len;
    }
    auto al = [](uint64_t b) { return (b + 255) / 256 * 256; };
    uint64_t off = 0;
    const uint64_t o_hvA = off; off += al(n * 8 + 8);
    const uint64_t o_hvB = off; off += al(n * 8 + 8);
    const uint64_t o_entA = off; off += al(n * 4 + 4);
    const uint64_t o_entB = off; off += al(n * 4 + 4);
    const uint64_t o_cnt = off; off += 256;
    if (int rc = c->wbuf.reserve(off)) return rc;
    unsigned char *B = c->wbuf.as<unsigned char>();
    uint64_t *hvA = (uint64_t *)(B + o_hvA), *hvB = (uint64_t *)(B + o_hvB);
    uint32_t *entA = (uint32_t *)(B + o_entA), *entB = (uint32_t *)(B + o_entB);
    unsigned long long *cnt = (unsigned long long *)(B + o_cnt);
    CU(cudaMemsetAsync(hvA, 0xFF, n * 8 + 8, st));
    CU(cudaMemsetAsync(entA, 0xFF, n * 4 + 4, st));
    CU(cudaMemsetAsync(cnt, 0, 8, st));
    d2g::SketchArgs a = make_sketch_args(c, p, seq_d, off_d, c->recent.as<uint32_t>(), n_rec, n, 0, SketchRange{0, n, 0});
    if (a.span >= 0xFFFFFFFFULL) return fail(D2G_EINVAL, "span too large");
    d2g::EmitConsumer::Params ep{hvA, entA, a.span};
    if (int rc = launch_sketch<d2g::EmitConsumer>(c, a, ep, p->w > p->k, D2G_T_SKETCH_MAIN)) return rc;
    const int ebits = ent_sort_bits(1u);   // the entity pass sorts only the bits entities use (the 0xFFFFFFFF sentinel still sorts last)
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, hvA, hvB, entA, entB, n, 0, 64, st);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, entB, entA, hvB, hvA, n, 0, ebits, st);
    const size_t tb = std::max(t1, t2);
    if (int rc = c->wtmp.reserve(tb + 256)) return rc;
    size_t tbytes = tb;
    CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, hvA, hvB, entA, entB, n, 0, 64, st));
    tbytes = tb;
    CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, entB, entA, hvB, hvA, n, 0, ebits, st));
    count_valid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(entA, n, cnt);
    c->launches += 2 * 9 + 1;
    unsigned long long h_n = 0;
    CU(cudaMemcpyAsync(&h_n, cnt, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (h_n) {
        if (int rc = c->filter.reserve(h_n * 8)) return rc;
        CU(cudaMemcpyAsync(c->filter.p, hvA, h_n * 8, cudaMemcpyDeviceToDevice, st));   // sorted by value; duplicates do not disturb the search
        CU(cudaStreamSynchronize(st));
    }
    c->filter_n = h_n;
    if (n_out) *n_out = h_n;
    return D2G_OK;
}

// ---- --save-kmercounts (-N): multiplicity of the element that owns each register ---------------------------------------
// The reference keeps, beside every register, how often the element that set it was seen (one-permutation sketch: counts_ bumped when
// an update equals the register, src/oph.h:206-209; Full SetSketch: setsketch.h:405-406; BagMinHash / ProbMinHash: the element's weight)
// and writes them as float32 to FILE.kmercounts.f64 (src/sketch_core.cpp:162-171).  That count is a function of the register's id alone:
// its multiplicity in the stream of hashed k-mers (one per window when w > k) that fed the entity.  Device: the emit pass of the counting
// sketches, a sort by (entity, value), and two binary searches per (entity, register).
namespace {
__global__ void kmer_count_lookup_kernel(const uint64_t *hv, const uint32_t *ent, uint64_t n, const uint64_t *ids, uint64_t n_ids, uint32_t S, float *out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_ids) return;
    const uint32_t e = (uint32_t)(i / S);
    const uint64_t id = ids[i];
    uint64_t lo = 0, hi = n;                                         // first index with (ent, hv) >= (e, id)
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; const uint32_t me = ent[mid]; if (me < e || (me == e && hv[mid] < id)) lo = mid + 1; else hi = mid; }
    const uint64_t lb = lo;
    hi = n;                                                          // first index with (ent, hv) > (e, id)
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; const uint32_t me = ent[mid]; if (me < e || (me == e && hv[mid] <= id)) lo = mid + 1; else hi = mid; }
    out[i] = (float)(lo - lb);
}
}

extern "C" int d2g_kmer_counts(d2g_ctx *c, const d2g_sketch_params *p, const uint64_t *codes, const uint32_t *mask, const uint64_t *rec_off,
                               const uint32_t *rec_entity, uint64_t n_rec, uint32_t n_entities, const uint64_t *ids, float *counts_out) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_sketch_params(p)) return rc;
    if (p->countsketch_size) return fail(D2G_EUNSUPPORTED, "--save-kmercounts with a count sketch (--countsketch-size) is not implemented on the GPU");
    if (n_entities && (!ids || !counts_out)) return fail(D2G_EINVAL, "null ids / output");
    if (n_rec && (!rec_off || !rec_entity)) return fail(D2G_EINVAL, "null record tables");
    CU(cudaSetDevice(c->device));
    const uint32_t S = p->sketchsize;
    const uint64_t n_bases = n_rec ? rec_off[n_rec] : 0, n_ids = (uint64_t)n_entities * S;
    for (uint64_t i = 0; i < n_ids; ++i) counts_out[i] = 0.f;
    if (!n_bases || !n_rec || !n_entities) return D2G_OK;
    if (!codes) return fail(D2G_EINVAL, "null packed sequence");
    if (rec_off[0] != 0) return fail(D2G_EINVAL, "rec_off[0] must be 0");
    if (n_bases >= 0xFFFFFFF0ULL) return fail(D2G_EINVAL, "k-mer counts: at most 2^32 bases per call (got %llu)", (unsigned long long)n_bases);
    for (uint64_t r = 0; r < n_rec; ++r) {
        if (rec_off[r + 1] < rec_off[r]) return fail(D2G_EINVAL, "rec_off not monotone at %llu", (unsigned long long)r);
        if (rec_entity[r] >= n_entities) return fail(D2G_EINVAL, "rec_entity[%llu]=%u >= n_entities", (unsigned long long)r, rec_entity[r]);
        if (r && rec_entity[r] < rec_entity[r - 1]) return fail(D2G_EINVAL, "rec_entity must be non-decreasing");
    }
    const uint64_t nw = d2g::packed_words(n_bases);
    if (int rc = c->pcodes.reserve(nw * 8)) return rc;
    if (int rc = c->pmask.reserve(nw * 4)) return rc;
    if (int rc = c->recoff.reserve((n_rec + 1) * 8)) return rc;
    if (int rc = c->recent.reserve((n_rec + 1) * 4)) return rc;
    if (int rc = c->ids.reserve(n_ids * 8)) return rc;
    if (int rc = c->sig.reserve(n_ids * 4)) return rc;
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->pcodes.p, codes, nw * 8, cudaMemcpyHostToDevice, st));
    if (mask) CU(cudaMemcpyAsync(c->pmask.p, mask, nw * 4, cudaMemcpyHostToDevice, st)); else CU(cudaMemsetAsync(c->pmask.p, 0, nw * 4, st));
    CU(cudaMemcpyAsync(c->recoff.p, rec_off, (n_rec + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->recent.p, rec_entity, n_rec * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->ids.p, ids, n_ids * 8, cudaMemcpyHostToDevice, st));
    d2g::PackedSeq seq_d{c->pcodes.as<uint64_t>(), c->pmask.as<uint32_t>()};
    const uint64_t *off_d = c->recoff.as<uint64_t>();
    uint64_t n = n_bases;
    StreamView sv;
    if (is_stream_mode(p)) {                                        // element streams: slots are item regions (stream_kernels.cuh)
        if (int rc = prepare_stream(c, p, seq_d, nullptr, off_d, n_rec, n_bases, &sv)) return rc;
        p = &sv.p; seq_d = sv.seq; off_d = sv.rec_off_d; n = sv.total_len;
        if (n >= 0xFFFFFFF0ULL) return fail(D2G_EINVAL, "k-mer counts: at most 2^31 bases per call for this k-mer stream");
    }
    auto al = [](uint64_t b) { return (b + 255) / 256 * 256; };
    uint64_t off = 0;
    const uint64_t o_hvA = off; off += al(n * 8 + 8);
    const uint64_t o_hvB = off; off += al(n * 8 + 8);
    const uint64_t o_entA = off; off += al(n * 4 + 4);
    const uint64_t o_entB = off; off += al(n * 4 + 4);
    if (int rc = c->wbuf.reserve(off)) return rc;
    unsigned char *B = c->wbuf.as<unsigned char>();
    uint64_t *hvA = (uint64_t *)(B + o_hvA), *hvB = (uint64_t *)(B + o_hvB);
    uint32_t *entA = (uint32_t *)(B + o_entA), *entB = (uint32_t *)(B + o_entB);
    CU(cudaMemsetAsync(hvA, 0xFF, n * 8 + 8, st));
    CU(cudaMemsetAsync(entA, 0xFF, n * 4 + 4, st));
    d2g::SketchArgs a = make_sketch_args(c, p, seq_d, off_d, c->recent.as<uint32_t>(), n_rec, n, 0, SketchRange{0, n, 0});
    if (a.span >= 0xFFFFFFFFULL) return fail(D2G_EINVAL, "span too large");
    d2g::EmitConsumer::Params ep{hvA, entA, a.span};
    if (int rc = launch_sketch<d2g::EmitConsumer>(c, a, ep, p->w > p->k, D2G_T_SKETCH_MAIN)) return rc;
    const int ebits = ent_sort_bits(n_entities);   // the entity pass sorts only the bits entities use (the 0xFFFFFFFF sentinel still sorts last)
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, hvA, hvB, entA, entB, n, 0, 64, st);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, entB, entA, hvB, hvA, n, 0, ebits, st);
    const size_t tb = std::max(t1, t2);
    if (int rc = c->wtmp.reserve(tb + 256)) return rc;
    size_t tbytes = tb;
    CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, hvA, hvB, entA, entB, n, 0, 64, st));
    tbytes = tb;
    CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, entB, entA, hvB, hvA, n, 0, ebits, st));
    c->launches += 2 * 9;
    kmer_count_lookup_kernel<<<(unsigned)((n_ids + 255) / 256), 256, 0, st>>>(hvA, entA, n, c->ids.as<uint64_t>(), n_ids, S, c->sig.as<float>());
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(counts_out, c->sig.p, n_ids * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return D2G_OK;
}
