// api_sketch.cu -- sketch path of the C ABI (include/d2gpu.h): one-permutation MinHash and Full SetSketch launchers,
// host<->device staging of d2g_sketch_batch, and the parts of the reference's finalisation that are x87 long-double
// arithmetic on the host in the reference too (/root/reference/src/oph.h:240-263).
#include "api_sketch_launch.h"
#include "fss_kernels.cuh"
#include "host/pack_host.h"

namespace {
__global__ void opmh_ids_kernel(const uint64_t *regs, uint64_t *ids, uint64_t n_ent, uint32_t m, uint32_t S) {
    const uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (e >= n_ent * S) return;
    const uint64_t g = e / S, i = e % S;
    ids[e] = d2g::dhash_inv(regs[g * m + i]);   // src/oph.h:264-271
}
} // namespace

int check_sketch_params(const d2g_sketch_params *p) {
    if (!p) return fail(D2G_EINVAL, "null params");
    if (p->k < 1 || p->k > 32) return fail(D2G_EUNSUPPORTED, "k=%d: only 1..32 (exact 2-bit encoding) is implemented; k>32 rolling hash is out of scope", p->k);
    if (p->sketchsize == 0) return fail(D2G_EINVAL, "sketchsize must be > 0");
    if (p->w > p->k) {
        if (p->w > d2g::SK_MAX_W) return fail(D2G_EUNSUPPORTED, "window %d > %d not supported", p->w, d2g::SK_MAX_W);
        if (!p->canon) return fail(D2G_EUNSUPPORTED, "windowed minimizers without canonicalisation (-C -w) are not implemented on the GPU");
    }
    if (p->mode < D2G_MODE_OPMH || p->mode > D2G_MODE_PROBMINHASH) return fail(D2G_EINVAL, "bad sketch mode %d", p->mode);
    if ((p->mode == D2G_MODE_BAGMINHASH || p->mode == D2G_MODE_PROBMINHASH) && p->sketchsize < 2) return fail(D2G_EINVAL, "weighted sketches need sketchsize >= 2");
    if (p->count_threshold > 1 && p->mode == D2G_MODE_FULL_SETSKETCH)
        return fail(D2G_EUNSUPPORTED, "--count-threshold > 1 with --full-setsketch (CountFilteredCSetSketch, src/setsketch.h:1000-1132: the result depends on "
                                      "the order of the k-mers) is not implemented on the GPU; one-permutation and the counting sketches are");
    if (p->countsketch_size && p->mode != D2G_MODE_BAGMINHASH && p->mode != D2G_MODE_PROBMINHASH)
        return fail(D2G_EINVAL, "--countsketch-size applies to the counting sketches (--multiset / --prob) only");
    if (p->countsketch_size >> 40) return fail(D2G_EINVAL, "--countsketch-size too large");
    return D2G_OK;
}

namespace {
// regs_d: [n_ent][m] u64 for OPMH.  Launches the OPMH sketch kernels on the ctx stream.
int launch_opmh(d2g_ctx *c, const d2g_sketch_params *p, const d2g::PackedSeq &seq_d, const uint64_t *rec_off_d,
                const uint32_t *rec_ent_d, uint64_t n_rec, uint32_t n_ent, uint64_t total_len, uint64_t *regs_d,
                const SketchRange *range = nullptr) {
    const SketchRange rg = range ? *range : SketchRange{0, total_len, 0};
    const uint32_t m = d2g_opmh_m(p->sketchsize);
    const uint64_t nreg = (uint64_t)n_ent * m;
    if (nreg) { fill_u64_kernel<<<(unsigned)std::min<uint64_t>((nreg + 255) / 256, 4096), 256, 0, c->stream>>>(regs_d, nreg, ~0ULL); c->launches++; }
    if (rg.pos_end <= rg.pos_base || n_rec == 0) return D2G_OK;
    const bool windowed = p->w > p->k;
    d2g::SketchArgs a = make_sketch_args(c, p, seq_d, rec_off_d, rec_ent_d, n_rec, total_len, m, rg);
    d2g::OpmhConsumer::Params cp{regs_d, d2g::make_fastmod32(m), m};
    return windowed ? launch_sketch_windowed_set<d2g::OpmhConsumer>(c, a, cp) : launch_sketch<d2g::OpmhConsumer>(c, a, cp, false);
}

// Full SetSketch (see fss_kernels.cuh): boot -> threshold -> main -> long walks -> finalize.
// sig_d [n_ent][S] / card_d [n_ent] may be null.  Synchronises the stream to check the long-walk queue.
int launch_fss(d2g_ctx *c, const d2g_sketch_params *p, const d2g::PackedSeq &seq_d, const uint64_t *rec_off_d, const uint32_t *rec_ent_d,
               uint64_t n_rec, uint32_t n_ent, uint64_t total_len, double *sig_d, double *card_d, uint64_t *ids_d,
               const SketchRange *range = nullptr) {
    if (n_ent == 0) return D2G_OK;
    const SketchRange rg = range ? *range : SketchRange{0, total_len, 0};
    const uint64_t work_len = rg.pos_end > rg.pos_base ? rg.pos_end - rg.pos_base : 0;
    const uint32_t m = p->sketchsize;
    const uint64_t nreg = (uint64_t)n_ent * m;
    const uint64_t ovf_cap = 1ULL << 20;
    // aux layout: maxrv[nreg] | keys[nreg] | T[n_ent] | Tguess[n_ent] | npos[n_ent] | state[n_ent] (u32, padded) | ovf_count, n_redo | ovf[2*ovf_cap]
    const size_t aux_bytes = (nreg * 2 + (uint64_t)n_ent * 4 + 4 + 2 * ovf_cap) * 8;
    if (int rc = c->aux.reserve(aux_bytes)) return rc;
    uint64_t *maxrv = c->aux.as<uint64_t>(), *keys = maxrv + nreg;
    double *T = reinterpret_cast<double *>(keys + nreg), *Tguess = T + n_ent;
    unsigned long long *npos = reinterpret_cast<unsigned long long *>(Tguess + n_ent);
    uint32_t *state = reinterpret_cast<uint32_t *>(npos + n_ent);
    unsigned long long *ovf_count = reinterpret_cast<unsigned long long *>(npos + 2 * (uint64_t)n_ent);
    unsigned int *n_redo = reinterpret_cast<unsigned int *>(ovf_count + 1);
    uint64_t *ovf = reinterpret_cast<uint64_t *>(ovf_count + 4);
    CU(cudaMemsetAsync(npos, 0, (uint64_t)n_ent * 8, c->stream));
    CU(cudaMemsetAsync(ovf_count, 0, 32, c->stream));
    fill_u64_kernel<<<(unsigned)std::min<uint64_t>((nreg + 255) / 256, 4096), 256, 0, c->stream>>>(keys, nreg, d2g::FSS_KEY_EMPTY);
    c->launches++;
    const bool windowed = p->w > p->k;
    if (work_len && n_rec) {
        d2g::SketchArgs a = make_sketch_args(c, p, seq_d, rec_off_d, rec_ent_d, n_rec, total_len, m, rg);
        const int wsz = windowed ? p->w - p->k + 1 : 1;
        d2g::FssMainConsumer::Params mp{keys, T, ovf, ovf_count, ovf_cap, m};
        // Pass A: bound guessed from the sequence length, verified afterwards (fss_kernels.cuh); only inputs with many elements per register
        d2g::fss_entity_positions_kernel<<<(unsigned)((n_rec + 255) / 256), 256, 0, c->stream>>>(rec_off_d, rec_ent_d, n_rec, rg.ent_base, windowed ? p->w : p->k, npos);
        d2g::fss_guess_kernel<<<(n_ent + 255) / 256, 256, 0, c->stream>>>(npos, n_ent, m, wsz, getenv("D2G_FSS_NO_GUESS") ? 0 : 1, T, Tguess, state);
        c->launches += 2;
        a.ent_state = state; a.want_state = 0;
        if (int rc = windowed ? launch_sketch_windowed_set<d2g::FssMainConsumer>(c, a, mp) : launch_sketch<d2g::FssMainConsumer>(c, a, mp, false)) return rc;
        d2g::fss_verify_kernel<<<n_ent, 256, 0, c->stream>>>(keys, m, Tguess, state, n_redo);
        c->launches++;
        unsigned int h_redo = 0;
        CU(cudaMemcpyAsync(&h_redo, n_redo, 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (getenv("D2G_DEBUG")) fprintf(stderr, "[d2g] fss: %u of %u entities take the boot pass\n", h_redo, n_ent);
        if (h_redo) {
            // Pass B (small inputs, failed guesses): boot on every stride-th tile gives a first bound T per entity so the first
            // walks of the main pass are short; the main kernel keeps tightening it.  n_eff = elements fed to the sketch per
            // entity (with minimizer windows only ~2/(window+1) of the positions emit).  Cost model per position: 1/stride for
            // the boot pass plus the extra walkers a looser threshold admits => stride ~ sqrt(n_eff / (2 m ln m)).
            CU(cudaMemsetAsync(maxrv, 0, nreg * 8, c->stream));
            const double per_ent = (double)work_len / std::max(1u, n_ent);
            const double n_eff = windowed ? per_ent * 2. / (p->w - p->k + 2) : per_ent;
            const double mlnm = (double)m * std::log((double)m + 2.);
            uint32_t stride = 1;
            while (stride < 64 && (double)(stride * 2) <= std::sqrt(n_eff / (2. * mlnm)) * 4.) stride *= 2;
            // ... but every register must be hit by the sample (an unhit register leaves T infinite): >= 24 sampled elements per register
            while (stride > 1 && n_eff / stride < 24. * m) stride /= 2;
            if (const char *ev = getenv("D2G_FSS_BOOT_STRIDE")) stride = (uint32_t)std::max(1, atoi(ev));   // tuning knob
            a.tile_stride = stride; a.want_state = 1;
            d2g::FssBootConsumer::Params bp{maxrv, d2g::make_fastmod32(m), m};
            if (int rc = launch_sketch<d2g::FssBootConsumer>(c, a, bp, windowed, D2G_T_SKETCH_BOOT)) return rc;
            d2g::fss_threshold_kernel<<<n_ent, 256, 0, c->stream>>>(maxrv, m, T, state);
            c->launches++;
            a.tile_stride = 1;
            if (int rc = windowed ? launch_sketch_windowed_set<d2g::FssMainConsumer>(c, a, mp) : launch_sketch<d2g::FssMainConsumer>(c, a, mp, false)) return rc;
        }
        // long walks (normally none): dense permutation state per thread slot
        uint64_t nslots = std::min<uint64_t>(4096, (256ULL << 20) / ((uint64_t)m * 8));
        nslots = std::max<uint64_t>(32, nslots / 32 * 32);
        if (int rc = c->aux2.reserve(nslots * 2ULL * m * 4)) return rc;
        CU(cudaMemsetAsync(c->aux2.p, 0, nslots * 2ULL * m * 4, c->stream));
        d2g::fss_longwalk_kernel<<<(unsigned)(nslots / 32), 32, 0, c->stream>>>(ovf, ovf_count, ovf_cap, m, T, keys, c->aux2.as<uint32_t>());
        c->launches++;
        if (ids_d) {   // --save-kmers: second pass over the final registers (FssIdsConsumer, fss_kernels.cuh)
            unsigned long long h1 = 0;
            CU(cudaMemcpyAsync(&h1, ovf_count, 8, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            if (h1 > ovf_cap) return fail(D2G_EUNSUPPORTED, "Full SetSketch: %llu elements needed a long register walk (queue holds %llu)", h1, (unsigned long long)ovf_cap);
            // the bound of the ids pass: the largest final register of the entity (every point at or below it is replayed)
            d2g::fss_final_bound_kernel<<<n_ent, 256, 0, c->stream>>>(keys, m, T);
            CU(cudaMemsetAsync(ids_d, 0, nreg * 8, c->stream));
            CU(cudaMemsetAsync(ovf_count, 0, 8, c->stream));
            a.ent_state = nullptr; a.tile_stride = 1;
            d2g::FssIdsConsumer::Params ip{keys, T, ids_d, ovf, ovf_count, ovf_cap, m};
            if (int rc = launch_sketch<d2g::FssIdsConsumer>(c, a, ip, windowed, D2G_T_SKETCH_BOOT)) return rc;
            CU(cudaMemsetAsync(c->aux2.p, 0, nslots * 2ULL * m * 4, c->stream));
            d2g::fss_longwalk_ids_kernel<<<(unsigned)(nslots / 32), 32, 0, c->stream>>>(ovf, ovf_count, ovf_cap, m, T, keys, ids_d, c->aux2.as<uint32_t>());
            c->launches += 2;
        }
    } else if (ids_d && nreg) CU(cudaMemsetAsync(ids_d, 0, nreg * 8, c->stream));
    const uint64_t nthreads = std::max<uint64_t>(nreg, n_ent);
    d2g::fss_finalize_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, c->stream>>>(keys, n_ent, m, sig_d, card_d);
    c->launches++;
    unsigned long long h_ovf = 0;
    CU(cudaMemcpyAsync(&h_ovf, ovf_count, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    if (getenv("D2G_DEBUG")) {
        std::vector<double> hT(n_ent);
        cudaMemcpy(hT.data(), T, n_ent * 8, cudaMemcpyDeviceToHost);
        uint32_t ninf = 0; double tmax = 0, tmin = 1e308;
        for (double t : hT) { if (t > 1e300) ++ninf; else { tmax = std::max(tmax, t); tmin = std::min(tmin, t); } }
        fprintf(stderr, "[d2g] fss: n_ent=%u m=%u long-walk queue=%llu boot T: inf=%u min=%g max=%g\n", n_ent, m, h_ovf, ninf, tmin, tmax);
    }
    if (h_ovf > ovf_cap) return fail(D2G_EUNSUPPORTED, "Full SetSketch: %llu elements needed a long register walk (queue holds %llu); "
                                     "inputs this small relative to the sketch size are not supported in one batch", h_ovf, (unsigned long long)ovf_cap);
    return D2G_OK;
}

// Host finalisation of one-permutation registers: x87 long double, as the reference does on the host
// (src/oph.h:240-263).  Threads over entities.
void opmh_finalize_host(const uint64_t *regs, uint32_t n_ent, uint32_t m, uint32_t S, double *sig, double *card) {
    auto work = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t g = lo; g < hi; ++g) {
            const uint64_t *r = regs + (uint64_t)g * m;
            if (card) {
                long double sum = 0.L;
                for (uint32_t i = 0; i < m; ++i) sum = sum + (long double)r[i] * 0x1p-64L;
                card[g] = sum ? (double)((long double)m * ((long double)m / sum)) : (double)INFINITY;
            }
            if (sig) {
                uint64_t nempty = 0;
                for (uint32_t i = 0; i < m; ++i) nempty += r[i] == ~0ULL;
                const long double mul = -1.0 / (double)((uint64_t)m - nempty);
                double *o = sig + (uint64_t)g * S;
                for (uint32_t i = 0; i < S; ++i) {
                    const uint64_t x = r[i];
                    o[i] = (x == ~0ULL || x == 0) ? 0. : (double)(mul * logl(0x1p-64L * (long double)(~0ULL - x + 1)));
                }
            }
        }
    };
    unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 64u));
    nt = std::min<unsigned>(nt, std::max(1u, n_ent / 4));
    if (nt <= 1) { work(0, n_ent); return; }
    std::vector<std::thread> th;
    const uint32_t per = (n_ent + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) { const uint32_t lo = t * per, hi = std::min(n_ent, lo + per); if (lo < hi) th.emplace_back(work, lo, hi); }
    for (auto &t : th) t.join();
}

} // namespace

extern "C" int d2g_opmh_finalize(const uint64_t *regs_u64, uint32_t n_entities, uint32_t sketchsize, double *sig_out, double *card_out) {
    if (!regs_u64) return fail(D2G_EINVAL, "null registers");
    opmh_finalize_host(regs_u64, n_entities, d2g_opmh_m(sketchsize), sketchsize, sig_out, card_out);
    return D2G_OK;
}

// ---- device-resident entry points ---------------------------------------------------------------------------------------
namespace {
// Sketch kernels over a packed batch that is resident on the device; outputs on the device.  Asynchronous on the ctx stream
// except where a launcher has to read a counter back (Full SetSketch, counting sketches).
int sketch_packed_dev(d2g_ctx *c, const d2g_sketch_params *p, const d2g::PackedSeq &seq_d, const uint64_t *rec_off_d,
                      const uint32_t *rec_entity_d, uint64_t n_rec, uint32_t n_entities, uint64_t total_len,
                      uint64_t *regs_u64_out_d, double *sig_out_d, double *card_out_d, uint64_t *ids_out_d) {
    if (p->mode == D2G_MODE_OPMH) {
        if (sig_out_d || card_out_d)
            return fail(D2G_EINVAL, "OPMH signatures/cardinalities are x87 long-double transforms of the u64 minima (src/oph.h:240-263): "
                                    "take regs_u64_out_d and call d2g_opmh_finalize on the host");
        if (!regs_u64_out_d) return fail(D2G_EINVAL, "regs_u64_out_d required for OPMH");
        if (int rc = p->count_threshold > 1 ? launch_opmh_mincount(c, p, seq_d, rec_off_d, rec_entity_d, n_rec, n_entities, total_len, regs_u64_out_d)
                                             : launch_opmh(c, p, seq_d, rec_off_d, rec_entity_d, n_rec, n_entities, total_len, regs_u64_out_d)) return rc;
        if (ids_out_d) {
            const uint64_t n = (uint64_t)n_entities * p->sketchsize;
            opmh_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(regs_u64_out_d, ids_out_d, n_entities, d2g_opmh_m(p->sketchsize), p->sketchsize);
            c->launches++;
        }
        return D2G_OK;
    }
    if (p->mode == D2G_MODE_FULL_SETSKETCH)
        return launch_fss(c, p, seq_d, rec_off_d, rec_entity_d, n_rec, n_entities, total_len, sig_out_d, card_out_d, ids_out_d);
    return launch_weighted(c, p, seq_d, rec_off_d, rec_entity_d, n_rec, n_entities, total_len, sig_out_d, card_out_d, ids_out_d);
}

int check_records(const uint64_t *rec_off, const uint32_t *rec_entity, uint64_t n_rec, uint32_t n_entities) {
    if (n_rec && (!rec_off || !rec_entity)) return fail(D2G_EINVAL, "null record tables");
    for (uint64_t r = 0; r < n_rec; ++r) {
        if (rec_off[r + 1] < rec_off[r]) return fail(D2G_EINVAL, "rec_off not monotone at %llu", (unsigned long long)r);
        if (rec_entity[r] >= n_entities) return fail(D2G_EINVAL, "rec_entity[%llu]=%u >= n_entities", (unsigned long long)r, rec_entity[r]);
        if (r && rec_entity[r] < rec_entity[r - 1]) return fail(D2G_EINVAL, "rec_entity must be non-decreasing");
    }
    return D2G_OK;
}
} // namespace

extern "C" uint64_t d2g_packed_words(uint64_t n_bases) { return d2g::packed_words(n_bases); }

extern "C" int d2g_pack_sequences(const char *const *pieces, const uint64_t *piece_len, uint64_t n_pieces, uint64_t *codes, uint32_t *mask,
                                  uint64_t *n_invalid_words) {
    if (n_pieces && (!pieces || !piece_len)) return fail(D2G_EINVAL, "null pieces");
    if (!codes || !mask) return fail(D2G_EINVAL, "null output");
    uint64_t total = 0;
    for (uint64_t i = 0; i < n_pieces; ++i) { if (piece_len[i] && !pieces[i]) return fail(D2G_EINVAL, "null piece %llu", (unsigned long long)i); total += piece_len[i]; }
    const uint64_t nz = d2g_host::pack_pieces(pieces, piece_len, n_pieces, d2g::packed_words(total), codes, mask);
    if (n_invalid_words) *n_invalid_words = nz;
    return D2G_OK;
}

extern "C" int d2g_pack_dev(d2g_ctx *c, const char *seq_d, uint64_t total_len, uint64_t *codes_d, uint32_t *mask_d) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (!codes_d || !mask_d || (total_len && !seq_d)) return fail(D2G_EINVAL, "null buffer");
    CU(cudaSetDevice(c->device));
    const uint64_t nw = d2g::packed_words(total_len);
    KernelTimer kt(c, D2G_T_PACK);
    d2g::pack_ascii_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const uint8_t *>(seq_d), total_len, nw, codes_d, mask_d);
    c->launches++;
    CU(cudaGetLastError());
    return D2G_OK;
}

extern "C" int d2g_sketch_batch_packed_dev(d2g_ctx *c, const d2g_sketch_params *p, const uint64_t *codes_d, const uint32_t *mask_d,
                                           const uint64_t *rec_off_d, const uint32_t *rec_entity_d, uint64_t n_rec, uint32_t n_entities,
                                           uint64_t total_len, uint64_t *regs_u64_out_d, double *sig_out_d, double *card_out_d, uint64_t *ids_out_d) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_sketch_params(p)) return rc;
    if (total_len && (!codes_d || !mask_d)) return fail(D2G_EINVAL, "null packed sequence");
    CU(cudaSetDevice(c->device));
    return sketch_packed_dev(c, p, d2g::PackedSeq{codes_d, mask_d}, rec_off_d, rec_entity_d, n_rec, n_entities, total_len,
                             regs_u64_out_d, sig_out_d, card_out_d, ids_out_d);
}

extern "C" int d2g_sketch_batch_dev(d2g_ctx *c, const d2g_sketch_params *p, const char *seq_d, const uint64_t *rec_off_d,
                                    const uint32_t *rec_entity_d, uint64_t n_rec, uint32_t n_entities, uint64_t total_len,
                                    uint64_t *regs_u64_out_d, double *sig_out_d, double *card_out_d, uint64_t *ids_out_d) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_sketch_params(p)) return rc;
    CU(cudaSetDevice(c->device));
    const uint64_t nw = d2g::packed_words(total_len);
    if (int rc = c->pcodes.reserve(nw * 8)) return rc;
    if (int rc = c->pmask.reserve(nw * 4)) return rc;
    if (int rc = d2g_pack_dev(c, seq_d, total_len, c->pcodes.as<uint64_t>(), c->pmask.as<uint32_t>())) return rc;
    return sketch_packed_dev(c, p, d2g::PackedSeq{c->pcodes.as<uint64_t>(), c->pmask.as<uint32_t>()}, rec_off_d, rec_entity_d, n_rec, n_entities,
                             total_len, regs_u64_out_d, sig_out_d, card_out_d, ids_out_d);
}

// ---- host entry points ------------------------------------------------------------------------------------------------
namespace {
// Whole-entity chunks of ~target bases: the upload of chunk i+1 overlaps the kernels of chunk i.
struct Chunk { uint64_t r0, r1; uint32_t e0, e1; };
std::vector<Chunk> make_chunks(const uint64_t *rec_off, const uint32_t *rec_entity, uint64_t n_rec, uint32_t n_entities, uint64_t target) {
    std::vector<Chunk> chunks;
    uint64_t r0 = 0; uint32_t e0 = 0;
    for (uint64_t r = 0; r < n_rec; ++r) {
        const bool last = r + 1 == n_rec;
        if (last || (rec_entity[r + 1] != rec_entity[r] && rec_off[r + 1] - rec_off[r0] >= target)) {
            const uint32_t e1 = last ? n_entities : rec_entity[r + 1];
            chunks.push_back({r0, r + 1, e0, e1});
            r0 = r + 1; e0 = e1;
        }
    }
    if (chunks.empty()) chunks.push_back({0, 0, 0, n_entities});
    return chunks;
}

// What feeds the device: either ASCII the library packs itself (host threads, pinned staging ring), or arrays the caller packed.
struct HostSeq {
    const char *ascii = nullptr;                      // concatenated record bytes
    const uint64_t *codes = nullptr; const uint32_t *mask = nullptr;   // packed by the caller (mask may be null: no invalid base)
};

// Shared body of d2g_sketch_batch / d2g_sketch_batch_packed: upload (packing on the way when the input is ASCII), kernels per chunk,
// results back to the host.
int sketch_batch_host(d2g_ctx *c, const d2g_sketch_params *p, const HostSeq &hs, const uint64_t *rec_off, const uint32_t *rec_entity,
                      uint64_t n_rec, uint32_t n_entities, uint64_t *regs_u64_out, double *sig_out, double *card_out, uint64_t *ids_out,
                      uint64_t *n_kmers_hashed) {
    if (int rc = check_records(rec_off, rec_entity, n_rec, n_entities)) return rc;
    CU(cudaSetDevice(c->device));
    const uint64_t total_len = n_rec ? rec_off[n_rec] : 0;
    if (n_kmers_hashed) *n_kmers_hashed = d2g_count_kmers(rec_off, n_rec, p->k);
    const uint32_t S = p->sketchsize, m = d2g_opmh_m(S);
    const uint64_t nw_total = d2g::packed_words(total_len);
    if (int rc = c->pcodes.reserve(nw_total * 8)) return rc;
    if (int rc = c->pmask.reserve(nw_total * 4)) return rc;
    if (int rc = c->recoff.reserve((n_rec + 1) * 8)) return rc;
    if (int rc = c->recent.reserve((n_rec + 1) * 4)) return rc;
    if (n_rec) {
        CU(cudaMemcpyAsync(c->recoff.p, rec_off, (n_rec + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->recent.p, rec_entity, n_rec * 4, cudaMemcpyHostToDevice, c->stream));
    }
    uint64_t *codes_d = c->pcodes.as<uint64_t>(); uint32_t *mask_d = c->pmask.as<uint32_t>();
    const d2g::PackedSeq seq_d{codes_d, mask_d};
    const uint64_t *off_d = c->recoff.as<uint64_t>();
    const uint32_t *ent_d = c->recent.as<uint32_t>();
    const bool opmh_mincount = p->mode == D2G_MODE_OPMH && p->count_threshold > 1;   // counts need the whole batch sorted at once
    const bool chunked = (p->mode == D2G_MODE_OPMH || p->mode == D2G_MODE_FULL_SETSKETCH) && !opmh_mincount;
    auto off_at = [&](uint64_t r) -> uint64_t { return n_rec ? rec_off[r] : 0; };
    uint64_t target = 128ULL << 20;                   // bases per chunk
    if (const char *ev = getenv("D2G_CHUNK_BYTES")) target = std::max<uint64_t>(1, strtoull(ev, nullptr, 10));
    if (!chunked) target = ~0ULL;
    const std::vector<Chunk> chunks = make_chunks(rec_off, rec_entity, n_rec, n_entities, target);
    const size_t nch = chunks.size();
    // word range of a chunk: whole 128-base blocks (16-byte granules of both arrays); a block that straddles two chunks goes up with both
    auto w_lo = [&](size_t i) { return off_at(chunks[i].r0) / 128 * 4; };
    auto w_hi = [&](size_t i) { return i + 1 == nch ? nw_total : (off_at(chunks[i].r1) + 127) / 128 * 4; };

    std::vector<cudaEvent_t> up(nch, nullptr);        // upload of chunk i complete
    auto free_evs = [&]() { for (auto e : up) if (e) cudaEventDestroy(e); };
    for (size_t i = 0; i < nch; ++i) if (cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming) != cudaSuccess) { free_evs(); return fail(D2G_ECUDA, "cudaEventCreate failed"); }

    // Producer: runs on its own host thread so that packing and uploading chunk i+1 overlap the kernels of chunk i, whose launcher
    // may block on the stream (Full SetSketch reads a counter back per chunk).
    std::atomic<size_t> uploaded{0}; std::atomic<int> prc{0};
    std::string perr;
    auto producer = [&]() {
        cudaSetDevice(c->device);
        for (size_t i = 0; i < nch; ++i) {
            const uint64_t a = w_lo(i), b = w_hi(i);
            cudaError_t e = cudaSuccess;
            if (b > a) {
                if (hs.ascii) {
                    const int slot = (int)(i % 3);
                    if (!c->stage_free[slot]) cudaEventCreateWithFlags(&c->stage_free[slot], cudaEventDisableTiming);
                    else cudaEventSynchronize(c->stage_free[slot]);          // the copy that last used this slot has left it
                    if (c->stage[slot].reserve((b - a) * 12) != D2G_OK) { perr = d2g_last_error(); prc = D2G_ENOMEM; break; }
                    uint64_t *sc = reinterpret_cast<uint64_t *>(c->stage[slot].p);
                    uint32_t *sm = reinterpret_cast<uint32_t *>(sc + (b - a));
                    const uint64_t nz = d2g_host::pack_contiguous(hs.ascii, total_len, a, b, sc, sm);
                    e = cudaMemcpyAsync(codes_d + a, sc, (b - a) * 8, cudaMemcpyHostToDevice, c->copy_stream);
                    // mask words only travel when the chunk holds an invalid base at all
                    if (e == cudaSuccess) e = nz ? cudaMemcpyAsync(mask_d + a, sm, (b - a) * 4, cudaMemcpyHostToDevice, c->copy_stream)
                                                 : cudaMemsetAsync(mask_d + a, 0, (b - a) * 4, c->copy_stream);
                    cudaEventRecord(c->stage_free[slot], c->copy_stream);
                } else {
                    e = cudaMemcpyAsync(codes_d + a, hs.codes + a, (b - a) * 8, cudaMemcpyHostToDevice, c->copy_stream);
                    if (e == cudaSuccess) e = hs.mask ? cudaMemcpyAsync(mask_d + a, hs.mask + a, (b - a) * 4, cudaMemcpyHostToDevice, c->copy_stream)
                                                      : cudaMemsetAsync(mask_d + a, 0, (b - a) * 4, c->copy_stream);
                }
            }
            if (e != cudaSuccess) { perr = std::string("sequence upload failed: ") + cudaGetErrorString(e); prc = D2G_ECUDA; break; }
            cudaEventRecord(up[i], c->copy_stream);
            uploaded.store(i + 1, std::memory_order_release);
        }
    };
    std::thread prod;
    if (nch > 1) prod = std::thread(producer); else producer();
    auto finish = [&](int rc) { if (prod.joinable()) prod.join(); cudaStreamSynchronize(c->copy_stream); free_evs(); return rc; };

    if (p->mode == D2G_MODE_OPMH) {
        if (int rc = c->regs.reserve((uint64_t)n_entities * m * 8)) return finish(rc);
    } else {
        if (int rc = c->sig.reserve((uint64_t)n_entities * S * 8)) return finish(rc);
        if (int rc = c->card.reserve((uint64_t)n_entities * 8)) return finish(rc);
        if (ids_out) if (int rc = c->ids.reserve((uint64_t)n_entities * S * 8)) return finish(rc);
    }
    for (size_t i = 0; i < nch; ++i) {
        const Chunk &ch = chunks[i];
        while (uploaded.load(std::memory_order_acquire) <= i) {
            if (prc.load()) { const int rc = prc.load(); if (prod.joinable()) prod.join(); free_evs(); return fail(rc, "%s", perr.c_str()); }
            std::this_thread::yield();
        }
        cudaStreamWaitEvent(c->stream, up[i], 0);
        const uint64_t nr = ch.r1 - ch.r0; const uint32_t ne = ch.e1 - ch.e0;
        const SketchRange rg{off_at(ch.r0), off_at(ch.r1), ch.e0};
        int rc;
        if (opmh_mincount)
            rc = launch_opmh_mincount(c, p, seq_d, off_d, ent_d, n_rec, n_entities, total_len, c->regs.as<uint64_t>());
        else if (p->mode == D2G_MODE_OPMH)
            rc = launch_opmh(c, p, seq_d, off_d + ch.r0, ent_d + ch.r0, nr, ne, rg.pos_end, c->regs.as<uint64_t>() + (uint64_t)ch.e0 * m, &rg);
        else if (p->mode == D2G_MODE_FULL_SETSKETCH)
            rc = launch_fss(c, p, seq_d, off_d + ch.r0, ent_d + ch.r0, nr, ne, rg.pos_end, c->sig.as<double>() + (uint64_t)ch.e0 * S,
                            c->card.as<double>() + ch.e0, ids_out ? c->ids.as<uint64_t>() + (uint64_t)ch.e0 * S : nullptr, &rg);
        else
            rc = launch_weighted(c, p, seq_d, off_d, ent_d, n_rec, n_entities, total_len, c->sig.as<double>(), c->card.as<double>(),
                                 ids_out ? c->ids.as<uint64_t>() : nullptr);
        if (rc) return finish(rc);
    }
    if (prod.joinable()) prod.join();
    free_evs();
    if (p->mode == D2G_MODE_OPMH) {
        std::vector<uint64_t> tmp;
        uint64_t *hregs = regs_u64_out;
        if (!hregs) { tmp.resize((uint64_t)n_entities * m); hregs = tmp.data(); }
        if (ids_out) {
            if (int rc = c->ids.reserve((uint64_t)n_entities * S * 8)) return rc;
            const uint64_t n = (uint64_t)n_entities * S;
            if (n) { opmh_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->regs.as<uint64_t>(), c->ids.as<uint64_t>(), n_entities, m, S); c->launches++; }
            CU(cudaMemcpyAsync(ids_out, c->ids.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
        }
        CU(cudaMemcpyAsync(hregs, c->regs.p, (uint64_t)n_entities * m * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (sig_out || card_out) opmh_finalize_host(hregs, n_entities, m, S, sig_out, card_out);
        return D2G_OK;
    }
    if (sig_out) CU(cudaMemcpyAsync(sig_out, c->sig.p, (uint64_t)n_entities * S * 8, cudaMemcpyDeviceToHost, c->stream));
    if (card_out) CU(cudaMemcpyAsync(card_out, c->card.p, (uint64_t)n_entities * 8, cudaMemcpyDeviceToHost, c->stream));
    if (ids_out) CU(cudaMemcpyAsync(ids_out, c->ids.p, (uint64_t)n_entities * S * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    (void)regs_u64_out;
    return D2G_OK;
}
} // namespace

extern "C" int d2g_sketch_batch(d2g_ctx *c, const d2g_sketch_params *p, const char *seq, const uint64_t *rec_off,
                                const uint32_t *rec_entity, uint64_t n_rec, uint32_t n_entities, uint64_t *regs_u64_out,
                                double *sig_out, double *card_out, uint64_t *ids_out, uint64_t *n_kmers_hashed) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_sketch_params(p)) return rc;
    if (n_rec && rec_off && rec_off[n_rec] && !seq) return fail(D2G_EINVAL, "null sequence buffer");
    HostSeq hs; hs.ascii = seq;
    return sketch_batch_host(c, p, hs, rec_off, rec_entity, n_rec, n_entities, regs_u64_out, sig_out, card_out, ids_out, n_kmers_hashed);
}

extern "C" int d2g_sketch_batch_packed(d2g_ctx *c, const d2g_sketch_params *p, const uint64_t *codes, const uint32_t *mask,
                                       const uint64_t *rec_off, const uint32_t *rec_entity, uint64_t n_rec, uint32_t n_entities,
                                       uint64_t *regs_u64_out, double *sig_out, double *card_out, uint64_t *ids_out, uint64_t *n_kmers_hashed) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_sketch_params(p)) return rc;
    if (n_rec && rec_off && rec_off[n_rec] && !codes) return fail(D2G_EINVAL, "null packed sequence");
    HostSeq hs; hs.codes = codes; hs.mask = mask;
    return sketch_batch_host(c, p, hs, rec_off, rec_entity, n_rec, n_entities, regs_u64_out, sig_out, card_out, ids_out, n_kmers_hashed);
}
