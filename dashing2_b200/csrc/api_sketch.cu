// api_sketch.cu -- sketch path of the C ABI (include/d2gpu.h): one-permutation MinHash and Full SetSketch launchers,
// host<->device staging of d2g_sketch_batch, and the parts of the reference's finalisation that are x87 long-double
// arithmetic on the host in the reference too (/root/reference/src/oph.h:240-263).
#include "api_sketch_launch.h"
#include "fss_kernels.cuh"
#include "host/pack_host.h"
#include <chrono>

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
    const int A = p->alphabet;
    if (A != 0 && A != 4 && A != 20 && A != 14 && A != 6 && A != 8) return fail(D2G_EINVAL, "alphabet %d: 0/4 (DNA), 20, 14, 6 or 8", A);
    if (p->sketchsize == 0) return fail(D2G_EINVAL, "sketchsize must be > 0");
    if (A != 0 && A != 4) {
        const int kmax = A == 20 ? 14 : A == 14 ? 16 : A == 6 ? 24 : 22;     // rhtraits.h nper64
        if (p->k < 1 || p->k > kmax) return fail(D2G_EUNSUPPORTED, "k=%d with the %d-letter protein alphabet: 1..%d (exact encoding) is implemented; the rolling "
                                                 "hash over protein alphabets is not", p->k, A, kmax);
        if (p->canon) return fail(D2G_EINVAL, "protein k-mers are never canonical (src/options.h:328-331): pass canon = 0");
    } else if (p->k < 1 || p->k > 2048) return fail(D2G_EUNSUPPORTED, "k=%d: 1..32 (exact encoding) and 33..2048 (rolling hash) are implemented", p->k);
    if (p->w > p->k) {
        if (is_stream_mode(p) ? p->w - p->k + 1 > d2g::SK_MAX_W : p->w > d2g::SK_MAX_W) return fail(D2G_EUNSUPPORTED, "window %d with k=%d not supported (at most %d)", p->w, p->k, d2g::SK_MAX_W);
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

// Full SetSketch (see fss_kernels.cuh): guessed bound -> main -> verify; entities that fail (or are too small to guess for) take
// boot -> threshold -> main; walks that outrun the sparse permutation state go through the long-walk queue.
// sig_d [n_ent][S] / card_d [n_ent] may be null.  Synchronises the stream (it reads the verify / queue counters back).
//
// The long-walk queue holds at most one entry per k-mer position, so it is sized from the positions of the entities that can
// reach it, and those entities are processed in groups when that is more than the queue budget (D2G_FSS_QUEUE_ELEMS, default
// 2^27 entries = 2 GiB): read sets sketched with a large S (every element of a record with fewer than ~m ln m elements walks all m
// registers, in the reference too: src/setsketch.h:369-423) are slow, not refused.
int launch_fss(d2g_ctx *c, const d2g_sketch_params *p, const d2g::PackedSeq &seq_d, const uint64_t *rec_off_d, const uint32_t *rec_ent_d,
               uint64_t n_rec, uint32_t n_ent, uint64_t total_len, double *sig_d, double *card_d, uint64_t *ids_d,
               const SketchRange *range = nullptr) {
    if (n_ent == 0) return D2G_OK;
    const SketchRange rg = range ? *range : SketchRange{0, total_len, 0};
    const uint64_t work_len = rg.pos_end > rg.pos_base ? rg.pos_end - rg.pos_base : 0;
    const uint32_t m = p->sketchsize;
    const uint64_t nreg = (uint64_t)n_ent * m;
    const uint64_t small_cap = 1ULL << 16;            // queue of the guessed pass: its walks are short by construction
    // aux layout: maxrv[nreg] | keys[nreg] | T[n_ent] | Tguess[n_ent] | npos[n_ent] | state[n_ent] (u32) | gstate[n_ent] (u32) | ovf_count, n_redo | ovf[2*small_cap]
    const uint64_t n_ent2 = ((uint64_t)n_ent + 1) / 2;   // u64 slots for n_ent u32
    const size_t aux_bytes = (nreg * 2 + (uint64_t)n_ent * 3 + 2 * n_ent2 + 4 + 2 * small_cap) * 8;
    if (int rc = c->aux.reserve(aux_bytes)) return rc;
    uint64_t *maxrv = c->aux.as<uint64_t>(), *keys = maxrv + nreg;
    double *T = reinterpret_cast<double *>(keys + nreg), *Tguess = T + n_ent;
    unsigned long long *npos = reinterpret_cast<unsigned long long *>(Tguess + n_ent);
    uint32_t *state = reinterpret_cast<uint32_t *>(npos + n_ent);
    uint32_t *gstate = reinterpret_cast<uint32_t *>(npos + n_ent + n_ent2);
    unsigned long long *ovf_count = reinterpret_cast<unsigned long long *>(npos + n_ent + 2 * n_ent2);
    unsigned int *n_redo = reinterpret_cast<unsigned int *>(ovf_count + 1);
    uint64_t *ovf_small = reinterpret_cast<uint64_t *>(ovf_count + 4);
    cudaStream_t st = c->stream;
    CU(cudaMemsetAsync(npos, 0, (uint64_t)n_ent * 8, st));
    CU(cudaMemsetAsync(ovf_count, 0, 32, st));
    fill_u64_kernel<<<(unsigned)std::min<uint64_t>((nreg + 255) / 256, 4096), 256, 0, st>>>(keys, nreg, d2g::FSS_KEY_EMPTY);
    c->launches++;
    const bool windowed = p->w > p->k;
    const bool debug = getenv("D2G_DEBUG") != nullptr;
    uint64_t total_long = 0;

    // the queued walks: one thread per element with a dense permutation state in scratch (2 m u32 per slot)
    auto run_longwalk = [&](const uint64_t *ovf, uint64_t cap, unsigned long long h_q, bool ids) -> int {
        if (h_q > cap) return fail(D2G_ECUDA, "Full SetSketch: long-walk queue overflow (%llu entries, room for %llu)", h_q, (unsigned long long)cap);
        if (!h_q) return D2G_OK;
        total_long += h_q;
        const uint64_t max_slots = std::max<uint64_t>(32, std::min<uint64_t>(65536, (2048ULL << 20) / ((uint64_t)m * 8)) / 32 * 32);
        const uint64_t nslots = std::min<uint64_t>(max_slots, (h_q + 31) / 32 * 32);
        if (int rc = c->aux2.reserve(nslots * 2ULL * m * 4)) return rc;
        CU(cudaMemsetAsync(c->aux2.p, 0, nslots * 2ULL * m * 4, st));
        if (ids) d2g::fss_longwalk_ids_kernel<<<(unsigned)(nslots / 32), 32, 0, st>>>(ovf, ovf_count, cap, m, T, keys, ids_d, c->aux2.as<uint32_t>());
        else d2g::fss_longwalk_kernel<<<(unsigned)(nslots / 32), 32, 0, st>>>(ovf, ovf_count, cap, m, T, keys, c->aux2.as<uint32_t>());
        c->launches++;
        CU(cudaGetLastError());
        return D2G_OK;
    };
    auto read_queue = [&](unsigned long long *h_q) -> int {
        CU(cudaMemcpyAsync(h_q, ovf_count, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return D2G_OK;
    };
    // Consecutive entities selected by `take`, in groups whose positions fit the queue budget; for each group gstate = 1 exactly on its
    // entities and body(queue, cap) runs the passes that may fill the queue.
    uint64_t budget = 1ULL << 27;
    if (const char *ev = getenv("D2G_FSS_QUEUE_ELEMS")) budget = std::max<uint64_t>(1024, strtoull(ev, nullptr, 10));
    std::vector<unsigned long long> h_npos;
    auto for_groups = [&](const std::vector<char> &take, auto &&body) -> int {
        if (h_npos.empty()) {
            h_npos.resize(n_ent);
            CU(cudaMemcpyAsync(h_npos.data(), npos, (uint64_t)n_ent * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        std::vector<uint32_t> hg(n_ent);
        uint32_t e = 0;
        while (e < n_ent) {
            while (e < n_ent && !take[e]) ++e;
            if (e >= n_ent) break;
            std::fill(hg.begin(), hg.end(), 0u);
            uint64_t sum = 0; uint32_t cnt = 0;
            for (; e < n_ent; ++e) {
                if (!take[e]) continue;
                if (cnt && sum + h_npos[e] > budget) break;
                hg[e] = 1; sum += h_npos[e]; ++cnt;
            }
            // an element is queued at most once per pass over its tile (the fast windowed kernel may hand a tile to the exact kernel: twice)
            const uint64_t cap = std::min<uint64_t>(2 * sum + 1024, 2 * budget + 1024);
            if (int rc = c->ovfq.reserve(2 * cap * 8)) return rc;
            CU(cudaMemcpyAsync(gstate, hg.data(), (uint64_t)n_ent * 4, cudaMemcpyHostToDevice, st));
            CU(cudaMemsetAsync(ovf_count, 0, 8, st));
            if (int rc = body(c->ovfq.as<uint64_t>(), cap)) return rc;
            CU(cudaStreamSynchronize(st));                // hg is reused by the next group
        }
        return D2G_OK;
    };

    if (work_len && n_rec) {
        d2g::SketchArgs a = make_sketch_args(c, p, seq_d, rec_off_d, rec_ent_d, n_rec, total_len, m, rg);
        const int wsz = windowed ? p->w - p->k + 1 : 1;
        // Pass A: bound guessed from the sequence length, verified afterwards (fss_kernels.cuh); only inputs with many elements per register
        d2g::fss_entity_positions_kernel<<<(unsigned)((n_rec + 255) / 256), 256, 0, st>>>(rec_off_d, rec_ent_d, n_rec, rg.ent_base, windowed ? p->w : p->k, npos);
        d2g::fss_guess_kernel<<<(n_ent + 255) / 256, 256, 0, st>>>(npos, n_ent, m, wsz, getenv("D2G_FSS_NO_GUESS") ? 0 : 1, T, Tguess, state);
        c->launches += 2;
        {
            d2g::FssMainConsumer::Params mp{keys, T, ovf_small, ovf_count, small_cap, m};
            a.ent_state = state; a.want_state = 0;
            if (int rc = windowed ? launch_sketch_windowed_set<d2g::FssMainConsumer>(c, a, mp) : launch_sketch<d2g::FssMainConsumer>(c, a, mp, false)) return rc;
        }
        unsigned long long h_cnt[2] = {0, 0};            // queue entries of the guessed pass, entities that failed
        CU(cudaMemcpyAsync(&h_cnt[0], ovf_count, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (h_cnt[0] > small_cap) {
            // more long walks than the guessed pass has room for (the length-based estimate was far off): fail every guess, the
            // entities are redone below with a queue sized for them
            CU(cudaMemsetAsync(Tguess, 0xFF, (uint64_t)n_ent * 8, st));
        } else if (int rc = run_longwalk(ovf_small, small_cap, h_cnt[0], false)) return rc;
        d2g::fss_verify_kernel<<<n_ent, 256, 0, st>>>(keys, m, Tguess, state, n_redo);
        c->launches++;
        unsigned int h_redo = 0;
        CU(cudaMemcpyAsync(&h_redo, n_redo, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (debug) fprintf(stderr, "[d2g] fss: %u of %u entities take the boot pass\n", h_redo, n_ent);
        if (h_redo) {
            // Pass B (small inputs, failed guesses): boot on every stride-th tile gives a first bound T per entity so the first
            // walks of the main pass are short; the main kernel keeps tightening it.  n_eff = elements fed to the sketch per
            // entity (with minimizer windows only ~2/(window+1) of the positions emit).  Cost model per position: 1/stride for
            // the boot pass plus the extra walkers a looser threshold admits => stride ~ sqrt(n_eff / (2 m ln m)).
            std::vector<uint32_t> h_state(n_ent);
            CU(cudaMemcpyAsync(h_state.data(), state, (uint64_t)n_ent * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            std::vector<char> take(n_ent);
            for (uint32_t e = 0; e < n_ent; ++e) take[e] = h_state[e] == 1;
            CU(cudaMemsetAsync(maxrv, 0, nreg * 8, st));
            const double per_ent = (double)work_len / std::max(1u, n_ent);
            const double n_eff = windowed ? per_ent * 2. / (p->w - p->k + 2) : per_ent;
            const double mlnm = (double)m * std::log((double)m + 2.);
            uint32_t stride = 1;
            while (stride < 64 && (double)(stride * 2) <= std::sqrt(n_eff / (2. * mlnm)) * 4.) stride *= 2;
            // ... but every register must be hit by the sample (an unhit register leaves T infinite): >= 24 sampled elements per register
            while (stride > 1 && n_eff / stride < 24. * m) stride /= 2;
            if (const char *ev = getenv("D2G_FSS_BOOT_STRIDE")) stride = (uint32_t)std::max(1, atoi(ev));   // tuning knob
            if (int rc = for_groups(take, [&](uint64_t *q, uint64_t cap) -> int {
                    a.ent_state = gstate; a.want_state = 1; a.tile_stride = stride;
                    d2g::FssBootConsumer::Params bp{maxrv, d2g::make_fastmod32(m), m};
                    if (int rc = launch_sketch<d2g::FssBootConsumer>(c, a, bp, windowed, D2G_T_SKETCH_BOOT)) return rc;
                    d2g::fss_threshold_kernel<<<n_ent, 256, 0, st>>>(maxrv, m, T, gstate);
                    c->launches++;
                    a.tile_stride = 1;
                    d2g::FssMainConsumer::Params mp{keys, T, q, ovf_count, cap, m};
                    if (int rc = windowed ? launch_sketch_windowed_set<d2g::FssMainConsumer>(c, a, mp) : launch_sketch<d2g::FssMainConsumer>(c, a, mp, false)) return rc;
                    unsigned long long h_q = 0;
                    if (int rc = read_queue(&h_q)) return rc;
                    return run_longwalk(q, cap, h_q, false);
                })) return rc;
        }
        if (ids_d) {   // --save-kmers: second pass over the final registers (FssIdsConsumer, fss_kernels.cuh)
            // the bound of the ids pass: the largest final register of the entity (every point at or below it is replayed)
            d2g::fss_final_bound_kernel<<<n_ent, 256, 0, st>>>(keys, m, T);
            c->launches++;
            CU(cudaMemsetAsync(ids_d, 0, nreg * 8, st));
            std::vector<char> take(n_ent, 1);
            if (int rc = for_groups(take, [&](uint64_t *q, uint64_t cap) -> int {
                    a.ent_state = gstate; a.want_state = 1; a.tile_stride = 1;
                    d2g::FssIdsConsumer::Params ip{keys, T, ids_d, q, ovf_count, cap, m};
                    if (int rc = launch_sketch<d2g::FssIdsConsumer>(c, a, ip, windowed, D2G_T_SKETCH_BOOT)) return rc;
                    unsigned long long h_q = 0;
                    if (int rc = read_queue(&h_q)) return rc;
                    return run_longwalk(q, cap, h_q, true);
                })) return rc;
        }
    } else if (ids_d && nreg) CU(cudaMemsetAsync(ids_d, 0, nreg * 8, st));
    const uint64_t nthreads = std::max<uint64_t>(nreg, n_ent);
    d2g::fss_finalize_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(keys, n_ent, m, sig_d, card_d);
    c->launches++;
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    if (debug) {
        std::vector<double> hT(n_ent);
        cudaMemcpy(hT.data(), T, n_ent * 8, cudaMemcpyDeviceToHost);
        uint32_t ninf = 0; double tmax = 0, tmin = 1e308;
        for (double t : hT) { if (t > 1e300) ++ninf; else { tmax = std::max(tmax, t); tmin = std::min(tmin, t); } }
        fprintf(stderr, "[d2g] fss: n_ent=%u m=%u long walks=%llu boot T: inf=%u min=%g max=%g\n", n_ent, m, (unsigned long long)total_long, ninf, tmin, tmax);
    }
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
                      uint64_t *regs_u64_out_d, double *sig_out_d, double *card_out_d, uint64_t *ids_out_d, const uint8_t *ascii_d = nullptr) {
    if (is_stream_mode(p)) {       // k > 32, -C with a window, protein: elements first, then the same launchers over item regions
        StreamView sv;
        if (int rc = prepare_stream(c, p, seq_d, ascii_d, rec_off_d, n_rec, total_len, &sv)) return rc;
        return sketch_packed_dev(c, &sv.p, sv.seq, sv.rec_off_d, rec_entity_d, n_rec, n_entities, sv.total_len, regs_u64_out_d, sig_out_d, card_out_d, ids_out_d);
    }
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
    if (p->alphabet != 0 && p->alphabet != 4) return fail(D2G_EINVAL, "protein alphabets need the record bytes, not packed DNA");
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
                             total_len, regs_u64_out_d, sig_out_d, card_out_d, ids_out_d, reinterpret_cast<const uint8_t *>(seq_d));
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
    const bool stream_mode = is_stream_mode(p), protein = p->alphabet != 0 && p->alphabet != 4;
    if (protein && !hs.ascii) return fail(D2G_EINVAL, "protein alphabets need the record bytes, not packed DNA");
    const bool chunked = (p->mode == D2G_MODE_OPMH || p->mode == D2G_MODE_FULL_SETSKETCH) && !opmh_mincount && !stream_mode;
    auto off_at = [&](uint64_t r) -> uint64_t { return n_rec ? rec_off[r] : 0; };
    uint64_t target = 384ULL << 20;                   // bases per chunk (per-chunk launches and read-backs cost ~0.5 ms: amortised over >= 1.5 ms of kernel)
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
    bool ascii_pinned = false, fixed_f = false;
    // Share of a chunk the host threads pack; the rest goes up as ASCII by DMA and is packed on the device.  Balance of f / P (host) against
    // ((1 - f) + f / 4) / B (link) for P = 4.6 G bases/s per packing thread (the measured best split on 16 threads is 0.7) and B = 50 GB/s: 0.70 with
    // 16 threads, 0.5 with 8, 0.31 with 4 -- processes that share a host (one per GPU) set D2G_HOST_THREADS to their share of the cores.
    double hybrid_f = [] { const double P = 4.6e9 * d2g_host::host_threads(), B = 50e9; return std::max(0.1, std::min(0.9, (1. / B) / (1. / P + 0.75 / B))); }();
    if (hs.ascii && total_len) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, hs.ascii) == cudaSuccess) ascii_pinned = at.type == cudaMemoryTypeHost;
        else cudaGetLastError();
        // The split is fixed per call (from the packing threads available, above); D2G_HYBRID_ADAPT=1 lets it follow
        // the packing rate measured per chunk (noisy: host threads and the DMA engine compete for the same memory bandwidth).
        fixed_f = !getenv("D2G_HYBRID_ADAPT");
        if (const char *ev = getenv("D2G_HYBRID_F")) { hybrid_f = std::max(0., std::min(1., atof(ev))); fixed_f = true; }   // 1 = pack everything on the host
        if (protein) { if (int rc = c->seq.reserve(total_len + 64)) { free_evs(); return rc; } }
        else if (ascii_pinned) {                           // device staging of the ASCII tails: sized once for the largest chunk
            uint64_t mw = 0;
            for (size_t i = 0; i < nch; ++i) mw = std::max(mw, w_hi(i) - w_lo(i));
            if (int rc = c->seq.reserve(mw * 32 + 4096)) { free_evs(); return rc; }
        }
    }
    std::string perr;
    auto producer = [&]() {
        cudaSetDevice(c->device);
        for (size_t i = 0; i < nch; ++i) {
            const uint64_t a = w_lo(i), b = w_hi(i);
            cudaError_t e = cudaSuccess;
            if (b > a) {
                if (protein) {                                       // residues go up as they are (one chunk): protein_kernel reads bytes
                    e = cudaMemcpyAsync(c->seq.p, hs.ascii, total_len, cudaMemcpyHostToDevice, c->copy_stream);
                } else if (hs.ascii) {
                    // Hybrid: the tail [ws, b) of the chunk goes up as ASCII (DMA from page-locked memory costs no host cycles) and is packed by
                    // the device; the head [a, ws) is packed by the host threads meanwhile and goes up at a quarter of the bytes.  The split
                    // balances host packing against the link: f / P = ((1 - f) + f / 4) / B for packing rate P and link rate B (bases/s, bytes/s).
                    const int slot = (int)(i % 3);
                    if (!c->stage_free[slot]) cudaEventCreateWithFlags(&c->stage_free[slot], cudaEventDisableTiming);
                    else cudaEventSynchronize(c->stage_free[slot]);          // the copy that last used this slot has left it
                    uint64_t ws = b;
                    if (ascii_pinned && hybrid_f < 1.) ws = std::min(b, a + (uint64_t)((double)(b - a) * hybrid_f) / 4 * 4);
                    if (c->stage[slot].reserve((ws - a) * 12 + 64) != D2G_OK) { perr = d2g_last_error(); prc = D2G_ENOMEM; break; }
                    if (ws < b) {
                        const uint64_t b0 = ws * 32, b1 = std::min(total_len, b * 32);
                        if (b1 > b0) {
                            e = cudaMemcpyAsync(c->seq.p, hs.ascii + b0, b1 - b0, cudaMemcpyHostToDevice, c->copy_stream);
                        }
                        if (e == cudaSuccess) {
                            d2g::pack_ascii_kernel<<<(unsigned)((b - ws + 255) / 256), 256, 0, c->copy_stream>>>(c->seq.as<uint8_t>(), b1 > b0 ? b1 - b0 : 0, b - ws, codes_d + ws, mask_d + ws);
                            c->launches++;
                            e = cudaGetLastError();
                        }
                    }
                    uint64_t *sc = reinterpret_cast<uint64_t *>(c->stage[slot].p);
                    uint32_t *sm = reinterpret_cast<uint32_t *>(sc + (ws - a));
                    const auto t0 = std::chrono::steady_clock::now();
                    const uint64_t nz = ws > a ? d2g_host::pack_contiguous(hs.ascii, total_len, a, ws, sc, sm) : 0;
                    const double th = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    if (ws > a && e == cudaSuccess) {
                        e = cudaMemcpyAsync(codes_d + a, sc, (ws - a) * 8, cudaMemcpyHostToDevice, c->copy_stream);
                        // mask words only travel when the chunk holds an invalid base at all
                        if (e == cudaSuccess) e = nz ? cudaMemcpyAsync(mask_d + a, sm, (ws - a) * 4, cudaMemcpyHostToDevice, c->copy_stream)
                                                     : cudaMemsetAsync(mask_d + a, 0, (ws - a) * 4, c->copy_stream);
                    }
                    cudaEventRecord(c->stage_free[slot], c->copy_stream);
                    if (ascii_pinned && !fixed_f && ws - a >= (1u << 16) && th > 0.) {   // re-balance from the packing rate just measured
                        const double P = (double)(ws - a) * 32. / th, B = 50e9;
                        const double want = (1. / B) / (1. / P + 0.75 / B);
                        hybrid_f = std::max(0.5, std::min(1., 0.5 * hybrid_f + 0.5 * want));   // damped: one slow chunk must not swing the split
                    }
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
        if (stream_mode)                                             // one chunk: the whole batch
            rc = sketch_packed_dev(c, p, seq_d, off_d, ent_d, n_rec, n_entities, total_len, p->mode == D2G_MODE_OPMH ? c->regs.as<uint64_t>() : nullptr,
                                   p->mode == D2G_MODE_OPMH ? nullptr : c->sig.as<double>(), p->mode == D2G_MODE_OPMH ? nullptr : c->card.as<double>(),
                                   (ids_out && p->mode != D2G_MODE_OPMH) ? c->ids.as<uint64_t>() : nullptr, protein ? c->seq.as<uint8_t>() : nullptr);
        else if (opmh_mincount)
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
