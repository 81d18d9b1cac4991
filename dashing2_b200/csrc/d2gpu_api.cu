// d2gpu_api.cu -- implementation of the C ABI declared in include/d2gpu.h.
//
// Host-side runtime around the kernels: context (device, stream, grow-only scratch), launch
// geometry, host<->device staging, and the parts of the reference's finalisation that are x87
// long-double arithmetic on the host in the reference too (one-permutation signature transform,
// /root/reference/src/oph.h:240-263).  There is deliberately no CPU implementation of the hot paths
// here: without a device every entry point fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/d2gpu.h"
#include "cmp_kernels.cuh"
#include "cmp16_kernels.cuh"
#include "sketch_kernels.cuh"
#include "fss_kernels.cuh"
#include "weighted_kernels.cuh"
#include "lsh_kernels.cuh"
#include <cub/device/device_segmented_radix_sort.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(D2G_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return D2G_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return fail(D2G_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); }
        cap = want; return D2G_OK;
    }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
    ~DevBuf() { if (p) cudaFree(p); }
};
struct PinBuf {
    void *p = nullptr; size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return D2G_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, n);
        if (e != cudaSuccess) { p = nullptr; return fail(D2G_ENOMEM, "cudaMallocHost(%zu) failed: %s", n, cudaGetErrorString(e)); }
        cap = n; return D2G_OK;
    }
    ~PinBuf() { if (p) cudaFreeHost(p); }
};

} // namespace

struct d2g_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    std::atomic<uint64_t> launches{0};
    DevBuf seq, recoff, recent, regs, sig, card, ids, aux, aux2;   // sketch scratch
    DevBuf wbuf, wtmp, lbuf;                                              // counting scratch (BagMinHash / ProbMinHash)
    DevBuf cregs, ccards, cout, clut, clut80, ctmp, cktmp;         // compare scratch
    DevBuf c16buf, c16codes, c16grank, c16flag;                    // order-code compare scratch (keys, sort buffers, codes, global ranks)
    struct { bool valid = false; const double *regs = nullptr; uint64_t g0 = 0, N = 0; uint32_t S = 0; int kind = 0; } c16g;   // global ranks built earlier in the same API call
    struct { bool valid = false; const double *regs = nullptr; uint64_t lo1 = 0, hi1 = 0, lo2 = 0, hi2 = 0; uint32_t S = 0; int kind = 0, mode = 0; } c16cache; // codes built earlier in the same API call
    PinBuf pin[2];
    cudaEvent_t ev[2] = {nullptr, nullptr}, evd[2] = {nullptr, nullptr};
    uint32_t lut_S = 0; int lut_k = -1;
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tev[D2G_T_NCLASSES];
};

namespace {
// RAII bracket: records an event pair around a kernel launch when ctx timing is on
struct KernelTimer {
    d2g_ctx *c; int cls; cudaEvent_t e0 = nullptr, e1 = nullptr;
    KernelTimer(d2g_ctx *ctx, int k) : c(ctx), cls(k) {
        if (!c->timing) return;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, c->stream);
    }
    ~KernelTimer() {
        if (!e0) return;
        cudaEventRecord(e1, c->stream);
        c->tev[cls].emplace_back(e0, e1);
    }
};
} // namespace

// -------------------------------------------------------------------------------------------------
extern "C" {

const char *d2g_last_error(void) { return g_err.c_str(); }
const char *d2g_version(void) { return "d2gpu 0.1 (sm_100a)"; }

int d2g_init(d2g_ctx **out, int device) {
    if (!out) return fail(D2G_EINVAL, "null ctx pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(D2G_ENODEVICE, "no CUDA device available (%s); libd2gpu has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(D2G_EINVAL, "device %d out of range (have %d)", device, n);
    CU(cudaSetDevice(device));
    d2g_ctx *c = new d2g_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev[1], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->evd[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->evd[1], cudaEventDisableTiming));
    *out = c;
    return D2G_OK;
}

void d2g_destroy(d2g_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->ev[0]) cudaEventDestroy(c->ev[0]);
    if (c->ev[1]) cudaEventDestroy(c->ev[1]);
    if (c->evd[0]) cudaEventDestroy(c->evd[0]);
    if (c->evd[1]) cudaEventDestroy(c->evd[1]);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

void *d2g_stream(d2g_ctx *c) { return c ? (void *)c->stream : nullptr; }
int d2g_sync(d2g_ctx *c) { if (!c) return fail(D2G_EINVAL, "null ctx"); CU(cudaSetDevice(c->device)); CU(cudaStreamSynchronize(c->stream)); return D2G_OK; }
uint64_t d2g_launch_count(const d2g_ctx *c) { return c ? c->launches.load() : 0; }
int d2g_set_timing(d2g_ctx *c, int on) { if (!c) return fail(D2G_EINVAL, "null ctx"); c->timing = on != 0; return D2G_OK; }
int d2g_get_timing(d2g_ctx *c, int cls, double *ms_total, uint64_t *n) {
    if (!c || cls < 0 || cls >= D2G_T_NCLASSES) return fail(D2G_EINVAL, "bad timing class");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    double tot = 0.;
    for (auto &pr : c->tev[cls]) { float ms = 0.f; cudaEventElapsedTime(&ms, pr.first, pr.second); tot += ms; cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    if (ms_total) *ms_total = tot;
    if (n) *n = c->tev[cls].size();
    c->tev[cls].clear();
    return D2G_OK;
}
void d2g_free(void *p) { free(p); }

uint32_t d2g_opmh_m(uint32_t S) { return S + (S & 1u); }

uint64_t d2g_count_kmers(const uint64_t *rec_off, uint64_t n_rec, int32_t k) {
    uint64_t t = 0;
    for (uint64_t r = 0; r < n_rec; ++r) { const uint64_t l = rec_off[r + 1] - rec_off[r]; if (l >= (uint64_t)k) t += l - k + 1; }
    return t;
}

} // extern "C"

// -------------------------------------------------------------------------------------------------
// sketch path
// -------------------------------------------------------------------------------------------------
namespace {

__global__ void fill_u64_kernel(uint64_t *p, uint64_t n, uint64_t v) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void opmh_ids_kernel(const uint64_t *regs, uint64_t *ids, uint64_t n_ent, uint32_t m, uint32_t S) {
    const uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (e >= n_ent * S) return;
    const uint64_t g = e / S, i = e % S;
    ids[e] = d2g::dhash_inv(regs[g * m + i]);   // src/oph.h:264-271
}

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
struct SketchRange { uint64_t pos_base, pos_end; uint32_t ent_base; };

d2g::SketchArgs make_sketch_args(const d2g_ctx *c, const d2g_sketch_params *p, const char *seq_d, const uint64_t *rec_off_d,
                                 const uint32_t *rec_ent_d, uint64_t n_rec, uint64_t total_len, uint32_t m, const SketchRange &rg) {
    d2g::SketchArgs a;
    a.seq = reinterpret_cast<const uint8_t *>(seq_d); a.rec_off = rec_off_d; a.rec_entity = rec_ent_d;
    a.n_rec = n_rec; a.total_len = total_len; a.k = p->k; a.w = p->w; a.canon = p->canon; a.xormask = p->xormask;
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

// regs_d: [n_ent][m] u64 for OPMH.  Launches the OPMH sketch kernels on the ctx stream.
int launch_opmh(d2g_ctx *c, const d2g_sketch_params *p, const char *seq_d, const uint64_t *rec_off_d,
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
    return launch_sketch<d2g::OpmhConsumer>(c, a, cp, windowed);
}

// Full SetSketch (see fss_kernels.cuh): boot -> threshold -> main -> long walks -> finalize.
// sig_d [n_ent][S] / card_d [n_ent] may be null.  Synchronises the stream to check the long-walk queue.
int launch_fss(d2g_ctx *c, const d2g_sketch_params *p, const char *seq_d, const uint64_t *rec_off_d, const uint32_t *rec_ent_d,
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
        if (int rc = launch_sketch<d2g::FssMainConsumer>(c, a, mp, windowed)) return rc;
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
            if (int rc = launch_sketch<d2g::FssMainConsumer>(c, a, mp, windowed)) return rc;
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

// BagMinHash / ProbMinHash (see weighted_kernels.cuh): emit -> sort -> run-length encode -> sketch -> verify loop.
// sig_d [n_ent][S], card_d [n_ent].  Synchronises (needs the number of distinct elements and the redo count).
__global__ void weighted_finalize_kernel(const uint64_t *keys, const unsigned long long *wsum, uint32_t n_ent, uint32_t m, double *sig, double *card) {
    const uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (sig && e < (uint64_t)n_ent * m) sig[e] = d2g::dunkey(keys[e]);
    if (card && e < n_ent) card[e] = (double)wsum[e];
}

int launch_weighted(d2g_ctx *c, const d2g_sketch_params *p, const char *seq_d, const uint64_t *rec_off_d, const uint32_t *rec_ent_d,
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
        size_t t1 = 0, t2 = 0, t3 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, t1, hvA, hvB, entA, entB, n, 0, 64, st);
        cub::DeviceRadixSort::SortPairs(nullptr, t2, entB, entA, hvB, hvA, n, 0, 32, st);
        cub::DeviceScan::ExclusiveSum(nullptr, t3, flag, excl, n, st);
        const size_t tb = std::max(t1, std::max(t2, t3));
        if (int rc = c->wtmp.reserve(tb + 256)) return rc;
        size_t tbytes = tb;
        CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, hvA, hvB, entA, entB, n, 0, 64, st));
        tbytes = tb;
        CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, entB, entA, hvB, hvA, n, 0, 32, st));
        c->launches += 2 * 9;
        const unsigned gb = (unsigned)((n + 255) / 256);
        d2g::rle_flag_kernel<<<gb, 256, 0, st>>>(hvA, entA, n, flag, id_shift);
        tbytes = tb;
        CU(cub::DeviceScan::ExclusiveSum(c->wtmp.p, tbytes, flag, excl, n, st));
        d2g::rle_scatter_kernel<<<gb, 256, 0, st>>>(flag, excl, entA, n, pos, n_valid);
        c->launches += 3;
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
        d2g::WeightedArgs wa{hvA, entA, pos, nu, n_valid, threshold, m, T, state, keys, ovf, ovf_count, ovf_cap, error, wts, id_shift, nullptr};
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
                    d2g::bmh_kernel<<<gu, 128, 0, st>>>(wa);
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
                d2g::bmh_kernel<<<gu, 128, 0, st>>>(wa);
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

int launch_opmh_mincount(d2g_ctx *c, const d2g_sketch_params *p, const char *seq_d, const uint64_t *rec_off_d, const uint32_t *rec_ent_d,
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
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, hvA, hvB, entA, entB, n, 0, 64, st);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, entB, entA, hvB, hvA, n, 0, 32, st);
    const size_t tb = std::max(t1, t2);
    if (int rc = c->wtmp.reserve(tb + 256)) return rc;
    size_t tbytes = tb;
    CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, hvA, hvB, entA, entB, n, 0, 64, st));
    tbytes = tb;
    CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, entB, entA, hvB, hvA, n, 0, 32, st));
    c->launches += 2 * 9;
    opmh_mincount_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hvA, entA, n, p->count_threshold, regs_d, d2g::make_fastmod32(m), m);
    c->launches++;
    CU(cudaGetLastError());
    return D2G_OK;
}
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
    const uint64_t n = n_rec ? rec_off[n_rec] : 0;
    if (n && !seq) return fail(D2G_EINVAL, "null sequence buffer");
    if (n_rec && rec_off[0] != 0) return fail(D2G_EINVAL, "rec_off[0] must be 0");
    if (n >= 0xFFFFFFF0ULL) return fail(D2G_EINVAL, "distinct k-mers: at most 2^32 bases per call (got %llu)", (unsigned long long)n);
    for (uint64_t r = 0; r < n_rec; ++r) {
        if (rec_off[r + 1] < rec_off[r]) return fail(D2G_EINVAL, "rec_off not monotone at %llu", (unsigned long long)r);
        if (rec_entity[r] >= n_entities) return fail(D2G_EINVAL, "rec_entity[%llu]=%u >= n_entities", (unsigned long long)r, rec_entity[r]);
        if (r && rec_entity[r] < rec_entity[r - 1]) return fail(D2G_EINVAL, "rec_entity must be non-decreasing");
    }
    for (uint32_t e = 0; e < n_entities; ++e) distinct_out[e] = 0;
    if (!n || !n_rec || !n_entities) return D2G_OK;
    if (int rc = c->seq.reserve(n + 64)) return rc;
    if (int rc = c->recoff.reserve((n_rec + 1) * 8)) return rc;
    if (int rc = c->recent.reserve((n_rec + 1) * 4)) return rc;
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->seq.p, seq, n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->recoff.p, rec_off, (n_rec + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->recent.p, rec_entity, n_rec * 4, cudaMemcpyHostToDevice, st));
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
    d2g::SketchArgs a = make_sketch_args(c, p, c->seq.as<char>(), c->recoff.as<uint64_t>(), c->recent.as<uint32_t>(), n_rec, n, 0, SketchRange{0, n, 0});
    if (a.span >= 0xFFFFFFFFULL) return fail(D2G_EINVAL, "span too large");
    d2g::EmitConsumer::Params ep{hvA, entA, a.span};
    if (int rc = launch_sketch<d2g::EmitConsumer>(c, a, ep, p->w > p->k, D2G_T_SKETCH_MAIN)) return rc;
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, hvA, hvB, entA, entB, n, 0, 64, st);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, entB, entA, hvB, hvA, n, 0, 32, st);
    const size_t tb = std::max(t1, t2);
    if (int rc = c->wtmp.reserve(tb + 256)) return rc;
    size_t tbytes = tb;
    CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, hvA, hvB, entA, entB, n, 0, 64, st));
    tbytes = tb;
    CU(cub::DeviceRadixSort::SortPairs(c->wtmp.p, tbytes, entB, entA, hvB, hvA, n, 0, 32, st));
    c->launches += 2 * 9;
    distinct_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hvA, entA, n, cnt);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(distinct_out, cnt, (uint64_t)n_entities * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return D2G_OK;
}

extern "C" int d2g_opmh_finalize(const uint64_t *regs_u64, uint32_t n_entities, uint32_t sketchsize, double *sig_out, double *card_out) {
    if (!regs_u64) return fail(D2G_EINVAL, "null registers");
    opmh_finalize_host(regs_u64, n_entities, d2g_opmh_m(sketchsize), sketchsize, sig_out, card_out);
    return D2G_OK;
}

extern "C" int d2g_sketch_batch_dev(d2g_ctx *c, const d2g_sketch_params *p, const char *seq_d, const uint64_t *rec_off_d,
                                    const uint32_t *rec_entity_d, uint64_t n_rec, uint32_t n_entities, uint64_t total_len,
                                    uint64_t *regs_u64_out_d, double *sig_out_d, double *card_out_d, uint64_t *ids_out_d) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_sketch_params(p)) return rc;
    CU(cudaSetDevice(c->device));
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

extern "C" int d2g_sketch_batch(d2g_ctx *c, const d2g_sketch_params *p, const char *seq, const uint64_t *rec_off,
                                const uint32_t *rec_entity, uint64_t n_rec, uint32_t n_entities, uint64_t *regs_u64_out,
                                double *sig_out, double *card_out, uint64_t *ids_out, uint64_t *n_kmers_hashed) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_sketch_params(p)) return rc;
    if (n_rec && (!rec_off || !rec_entity)) return fail(D2G_EINVAL, "null record tables");
    CU(cudaSetDevice(c->device));
    const uint64_t total_len = n_rec ? rec_off[n_rec] : 0;
    if (total_len && !seq) return fail(D2G_EINVAL, "null sequence buffer");
    for (uint64_t r = 0; r < n_rec; ++r) {
        if (rec_off[r + 1] < rec_off[r]) return fail(D2G_EINVAL, "rec_off not monotone at %llu", (unsigned long long)r);
        if (rec_entity[r] >= n_entities) return fail(D2G_EINVAL, "rec_entity[%llu]=%u >= n_entities", (unsigned long long)r, rec_entity[r]);
        if (r && rec_entity[r] < rec_entity[r - 1]) return fail(D2G_EINVAL, "rec_entity must be non-decreasing");
    }
    if (n_kmers_hashed) *n_kmers_hashed = d2g_count_kmers(rec_off, n_rec, p->k);
    const uint32_t S = p->sketchsize, m = d2g_opmh_m(S);
    if (int rc = c->seq.reserve(total_len + 64)) return rc;
    if (int rc = c->recoff.reserve((n_rec + 1) * 8)) return rc;
    if (int rc = c->recent.reserve((n_rec + 1) * 4)) return rc;
    if (n_rec) {
        CU(cudaMemcpyAsync(c->recoff.p, rec_off, (n_rec + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->recent.p, rec_entity, n_rec * 4, cudaMemcpyHostToDevice, c->stream));
    }
    const char *seq_d = c->seq.as<char>();
    const uint64_t *off_d = c->recoff.as<uint64_t>();
    const uint32_t *ent_d = c->recent.as<uint32_t>();
    const bool opmh_mincount = p->mode == D2G_MODE_OPMH && p->count_threshold > 1;   // counts need the whole batch sorted at once
    const bool chunked = (p->mode == D2G_MODE_OPMH || p->mode == D2G_MODE_FULL_SETSKETCH) && !opmh_mincount;
    auto off_at = [&](uint64_t r) -> uint64_t { return n_rec ? rec_off[r] : 0; };
    // Chunks of whole entities (~D2G_CHUNK_BYTES of sequence each, default 256 MiB): the upload of chunk i+1 runs on the copy
    // stream while chunk i is sketched, so a large batch moves at PCIe speed instead of copy + compute.
    struct Chunk { uint64_t r0, r1; uint32_t e0, e1; };
    std::vector<Chunk> chunks;
    {
        uint64_t target = 256ULL << 20;
        if (const char *ev = getenv("D2G_CHUNK_BYTES")) target = std::max<uint64_t>(1, strtoull(ev, nullptr, 10));
        if (!chunked) target = ~0ULL;
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
    }
    std::vector<cudaEvent_t> evs(chunks.size(), nullptr);
    auto free_evs = [&]() { for (auto e : evs) if (e) cudaEventDestroy(e); };
    for (size_t i = 0; i < chunks.size(); ++i) {
        const uint64_t b0 = off_at(chunks[i].r0), b1 = off_at(chunks[i].r1);
        if (b1 > b0) {
            cudaError_t e1 = cudaMemcpyAsync(c->seq.as<char>() + b0, seq + b0, b1 - b0, cudaMemcpyHostToDevice, chunks.size() > 1 ? c->copy_stream : c->stream);
            if (e1 != cudaSuccess) { free_evs(); return fail(D2G_ECUDA, "sequence upload failed: %s", cudaGetErrorString(e1)); }
        }
        if (chunks.size() > 1) {
            cudaEventCreateWithFlags(&evs[i], cudaEventDisableTiming);
            cudaEventRecord(evs[i], c->copy_stream);
        }
    }
    if (p->mode == D2G_MODE_OPMH) {
        if (int rc = c->regs.reserve((uint64_t)n_entities * m * 8)) { free_evs(); return rc; }
    } else {
        if (int rc = c->sig.reserve((uint64_t)n_entities * S * 8)) { free_evs(); return rc; }
        if (int rc = c->card.reserve((uint64_t)n_entities * 8)) { free_evs(); return rc; }
        if (ids_out) if (int rc = c->ids.reserve((uint64_t)n_entities * S * 8)) { free_evs(); return rc; }
    }
    for (size_t i = 0; i < chunks.size(); ++i) {
        const Chunk &ch = chunks[i];
        if (evs[i]) cudaStreamWaitEvent(c->stream, evs[i], 0);
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
        if (rc) { cudaStreamSynchronize(c->copy_stream); free_evs(); return rc; }
    }
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

// -------------------------------------------------------------------------------------------------
// compare path
// -------------------------------------------------------------------------------------------------
namespace {

int check_cmp_params(const d2g_cmp_params *p) {
    if (!p) return fail(D2G_EINVAL, "null params");
    if (p->sketchsize == 0) return fail(D2G_EINVAL, "sketchsize must be > 0");
    if (p->cmp_kind < D2G_CMP_GTLT || p->cmp_kind > D2G_CMP_BBIT) return fail(D2G_EINVAL, "bad cmp_kind %d", p->cmp_kind);
    if (p->cmp_kind >= D2G_CMP_SS_COMPRESSED) {
        if (p->regbytes != 1. && p->regbytes != 2. && p->regbytes != 4.) return fail(D2G_EINVAL, "compressed registers: regbytes must be 1, 2 or 4 (got %g)", p->regbytes);
        if (p->cmp_kind == D2G_CMP_SS_COMPRESSED && !(p->compressed_b > 1.L)) return fail(D2G_EINVAL, "compressed registers: base b must be > 1");
        if (p->sketchsize > 65535) return fail(D2G_EUNSUPPORTED, "compressed registers: sketchsize > 65535 not supported");
    }
    if (p->measure < 0 || p->measure > D2G_UNION_SIZE) return fail(D2G_EINVAL, "bad measure %d", p->measure);
    if (p->shape < 0 || p->shape > D2G_PANEL) return fail(D2G_EINVAL, "bad shape %d", p->shape);
    if (p->shape == D2G_PANEL && p->nq > p->n) return fail(D2G_EINVAL, "nq > n");
    return D2G_OK;
}
uint64_t n_rows(const d2g_cmp_params *p) { return p->shape == D2G_PANEL ? p->n - p->nq : p->n; }
uint64_t n_cols(const d2g_cmp_params *p) { return p->shape == D2G_PANEL ? p->nq : p->n; }
uint64_t rows_size(const d2g_cmp_params *p, uint64_t r0, uint64_t r1) {
    if (p->shape == D2G_SYMMETRIC) {
        auto tri = [&](uint64_t i) { return i * p->n - i * (i + 1) / 2; };
        return tri(r1) - tri(r0);
    }
    return (r1 - r0) * n_cols(p);
}

int make_consts(d2g_ctx *c, const d2g_cmp_params *p, d2g::CmpConsts *k) {
    const uint32_t S = p->sketchsize;
    k->invdenom = xf::from_long_double(1.L / S);
    k->eps = xf::from_long_double(1e-15L);
    k->poisson_mult = -1. / std::max(1, p->k);
    k->S = S; k->measure = p->measure; k->cmp_kind = p->cmp_kind;
    k->fast_sim = (p->measure == D2G_SIMILARITY && p->cmp_kind == D2G_CMP_GTLT && (S & (S - 1)) == 0) ? 1 : 0;
    k->eq_llr_lut = nullptr; k->lut80 = nullptr;
    if (p->cmp_kind >= D2G_CMP_SS_COMPRESSED) {
        // Everything that depends on the integer counts alone is x87 long-double arithmetic on the host in the reference
        // (powl in g_b, fmal, logl); tabulate it over the S + 1 possible counts (cmp_core.cpp:323-325,406-432).
        const long double invdenom = 1.L / S;
        std::vector<xf::f80> l80(S + 1);
        std::vector<long double> lv(S + 1);
        if (p->cmp_kind == D2G_CMP_BBIT) {
            const long double b2pow = -ldexpl(1.L, -(int)(p->regbytes * 8.));
            for (uint32_t e = 0; e <= S; ++e) lv[e] = std::max(0.L, fmal((long double)(uint64_t)e, invdenom, b2pow) / (1.L + b2pow));
        } else {
            const long double b = p->compressed_b;
            for (uint32_t e = 0; e <= S; ++e) lv[e] = (1.L - powl(b, -((uint64_t)e * invdenom))) / (1.L - 1.L / b);
        }
        for (uint32_t e = 0; e <= S; ++e) l80[e] = xf::from_long_double(lv[e]);
        if (int rc = c->clut80.reserve((S + 1) * sizeof(xf::f80))) return rc;
        CU(cudaMemcpyAsync(c->clut80.p, l80.data(), (S + 1) * sizeof(xf::f80), cudaMemcpyHostToDevice, c->stream));
        k->lut80 = c->clut80.as<xf::f80>();
        if (p->measure == D2G_POISSON_LLR) {
            auto llr = [&](long double ret) -> float {   // sim2dist on a long double argument (cmp_core.cpp:361) + :573
                ret = ret ? (long double)(double)(logl(2. * ret / (1. + ret)) * k->poisson_mult) : (long double)INFINITY;
                if (isnan(ret) || isinf(ret)) ret = __LDBL_MAX__;
                return (float)ret;
            };
            std::vector<float> lut;
            if (p->cmp_kind == D2G_CMP_BBIT) {
                lut.resize(S + 1);
                for (uint32_t e = 0; e <= S; ++e) lut[e] = llr(lv[e]);
            } else {
                lut.resize((size_t)(S + 1) * (S + 2) / 2);
                auto work = [&](uint32_t g0, uint32_t g1) {
                    for (uint32_t g = g0; g < g1; ++g)
                        for (uint32_t l = 0; l + g <= S; ++l)
                            lut[d2g::tri_index(g, l, S)] = llr(std::max(1.L - (lv[g] + lv[l]), 0.L));
                };
                const unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 32u));
                std::vector<std::thread> th;
                for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t]() { for (uint32_t g = t; g <= S; g += nt) work(g, g + 1); });
                for (auto &x : th) x.join();
            }
            if (int rc = c->clut.reserve(lut.size() * 4)) return rc;
            CU(cudaMemcpyAsync(c->clut.p, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice, c->stream));
            c->lut_S = 0; c->lut_k = -1;              // the equality-branch cache below no longer describes clut
            k->eq_llr_lut = c->clut.as<float>();
        }
        CU(cudaStreamSynchronize(c->stream));           // the host vectors go out of scope
        return D2G_OK;
    }
    if (p->cmp_kind == D2G_CMP_EQ && p->measure == D2G_POISSON_LLR) {
        if (c->lut_S != S || c->lut_k != p->k) {
            // equality branch of the Mash distance is long-double logl on the host in the reference
            // (cmp_core.cpp:361,509); it only depends on the integer count, so tabulate it here.
            std::vector<float> lut(S + 1);
            const long double invdenom = 1.L / S;
            for (uint32_t e = 0; e <= S; ++e) {
                long double ret = invdenom * e;
                ret = ret ? (long double)(double)(logl(2. * ret / (1. + ret)) * k->poisson_mult) : (long double)INFINITY;
                if (isnan(ret) || isinf(ret)) ret = __LDBL_MAX__;
                lut[e] = (float)ret;
            }
            if (int rc = c->clut.reserve((S + 1) * 4)) return rc;
            CU(cudaMemcpyAsync(c->clut.p, lut.data(), (S + 1) * 4, cudaMemcpyHostToDevice, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            c->lut_S = S; c->lut_k = p->k;
        }
        k->eq_llr_lut = c->clut.as<float>();
    }
    return D2G_OK;
}

// kinds 0 and 2 count (a > b, a < b); kinds 1 and 3 count bitwise-equal registers
inline bool counts_gtlt(int cmp_kind) { return cmp_kind == D2G_CMP_GTLT || cmp_kind == D2G_CMP_SS_COMPRESSED; }

// f64 tile kernel over rows [r0,r1) x columns [c0,c1) (global sketch ids); `base` carries the output mapping.
int launch_cmp_f64(d2g_ctx *c, const d2g_cmp_params *p, d2g::CmpArgs a, uint64_t r0, uint64_t r1, uint64_t c0, uint64_t c1,
                   const int *use_flag, int want) {
    if (r1 <= r0 || c1 <= c0) return D2G_OK;
    a.row0 = r0; a.row1 = r1; a.col0 = c0; a.col1 = c1; a.use_flag = use_flag; a.want = want;
    const uint64_t tiles_i = (r1 - r0 + d2g::CMP_T - 1) / d2g::CMP_T;
    a.tiles_j = (c1 - c0 + d2g::CMP_T - 1) / d2g::CMP_T;
    const uint64_t grid = tiles_i * a.tiles_j;
    if (grid > 0x7fffffffULL) return fail(D2G_EINVAL, "row block too large for one launch");
    KernelTimer kt(c, use_flag ? D2G_T_CMP_PREP : D2G_T_CMP);
    if (counts_gtlt(p->cmp_kind)) d2g::cmp_tile_kernel<0><<<(unsigned)grid, d2g::CMP_THREADS, 0, c->stream>>>(a);
    else d2g::cmp_tile_kernel<1><<<(unsigned)grid, d2g::CMP_THREADS, 0, c->stream>>>(a);
    c->launches++;
    CU(cudaGetLastError());
    return D2G_OK;
}

__global__ void fill_offsets_kernel(int64_t *offs, uint32_t nseg, uint64_t stride) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nseg) offs[i] = (int64_t)((uint64_t)i * stride);
}

// Per-register-position sort of the job's sketches + dense ranks, in groups of register positions so that one segmented
// sort holds < 2^31 items.  Writes half codes into codes16 (blocked layout) or, when grank != nullptr, the u32 ranks.
int c16_sort_rank(d2g_ctx *c, const d2g_cmp_params *p, d2g::C16Job j, uint16_t *codes16, uint32_t *grank, int *flag) {
    using namespace d2g;
    const uint32_t S = j.S;
    const uint64_t U = (uint64_t)j.nA + j.nB;
    const uint32_t group = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(S, 0x7fffffffULL / std::max<uint64_t>(1, U)));
    const uint64_t items_max = U * group;
    auto al = [](uint64_t b) { return (b + 255) / 256 * 256; };
    uint64_t off = 0;
    const uint64_t o_kA = off; off += al(items_max * 8); const uint64_t o_kB = off; off += al(items_max * 8);
    const uint64_t o_iA = off; off += al(items_max * 4); const uint64_t o_iB = off; off += al(items_max * 4);
    const uint64_t o_offs = off; off += al(((uint64_t)group + 1) * 8);
    if (int rc = c->c16buf.reserve(off)) return rc;
    unsigned char *B = c->c16buf.as<unsigned char>();
    uint64_t *kA = (uint64_t *)(B + o_kA), *kB = (uint64_t *)(B + o_kB);
    uint32_t *iA = (uint32_t *)(B + o_iA), *iB = (uint32_t *)(B + o_iB);
    int64_t *offs = (int64_t *)(B + o_offs);
    cudaStream_t st = c->stream;
    fill_offsets_kernel<<<(group + 1 + 255) / 256, 256, 0, st>>>(offs, group, U);
    c->launches++;
    for (uint32_t s0 = 0; s0 < S; s0 += group) {
        j.s_begin = s0; j.s_count = std::min(group, S - s0);
        const uint64_t items = U * j.s_count;
        const dim3 gk((unsigned)((U + 31) / 32), (j.s_count + 31) / 32);
        if (counts_gtlt(p->cmp_kind)) c16_keys_kernel<0><<<gk, 256, 0, st>>>(j, kA, iA, flag);
        else c16_keys_kernel<1><<<gk, 256, 0, st>>>(j, kA, iA, flag);
        size_t need = 0;
        cub::DeviceSegmentedRadixSort::SortPairs(nullptr, need, kA, kB, iA, iB, (int)items, (int)j.s_count, offs, offs + 1, 0, 64, st);
        if (int rc = c->wtmp.reserve(need + 256)) return rc;
        size_t tbytes = c->wtmp.cap;
        CU(cub::DeviceSegmentedRadixSort::SortPairs(c->wtmp.p, tbytes, kA, kB, iA, iB, (int)items, (int)j.s_count, offs, offs + 1, 0, 64, st));
        c16_rank_kernel<<<j.s_count, 256, 0, st>>>(j, kB, iB, codes16, grank, flag);
        c->launches += 2 + 11;
        CU(cudaGetLastError());
    }
    return D2G_OK;
}

// Multi-job comparisons: rank ALL sketches [g0, g0 + N) once per register position (u32 ranks in HBM); each job then
// derives its own dense codes from them with c16_local_codes_kernel instead of sorting its sketches again.
int c16_build_global(d2g_ctx *c, const d2g_cmp_params *p, const double *regs_d, uint64_t g0, uint64_t N) {
    using namespace d2g;
    auto &g = c->c16g;
    const uint32_t S = p->sketchsize;
    if (g.valid && g.regs == regs_d && g.g0 == g0 && g.N == N && g.S == S && g.kind == p->cmp_kind) return D2G_OK;
    g.valid = false;
    if (int rc = c->c16grank.reserve(N * S * 4)) return rc;
    if (int rc = c->c16flag.reserve(256)) return rc;
    KernelTimer kt(c, D2G_T_CMP_PREP);
    CU(cudaMemsetAsync(c->c16flag.p, 0, 4, c->stream));
    C16Job j{};
    j.regs = regs_d; j.S = S; j.gA0 = g0; j.nA = (uint32_t)N; j.nB = 0; j.posB0 = 0; j.KP = 0;
    if (int rc = c16_sort_rank(c, p, j, nullptr, c->c16grank.as<uint32_t>(), c->c16flag.as<int>())) return rc;
    g.valid = true; g.regs = regs_d; g.g0 = g0; g.N = N; g.S = S; g.kind = p->cmp_kind;
    return D2G_OK;
}

// One comparison job on 16-bit order codes (cmp16_kernels.cuh): the sketches [lo1,hi1) (and [lo2,hi2) when
// hi2 > lo2) are ranked per register position, coded, and rows [r0,r1) x columns [c0,c1) are compared.
// [r0,r1) must lie inside range 1; [c0,c1) inside range 2 when it exists, else inside range 1.
int run_cmp16_job(d2g_ctx *c, const d2g_cmp_params *p, const d2g::CmpArgs &base, uint64_t lo1, uint64_t hi1, uint64_t lo2, uint64_t hi2,
                  uint64_t r0, uint64_t r1, uint64_t c0, uint64_t c1) {
    using namespace d2g;
    if (r1 <= r0 || c1 <= c0) return D2G_OK;
    const uint32_t S = p->sketchsize;
    const bool two = hi2 > lo2;
    C16Job j{};
    j.regs = base.regs; j.S = S; j.gA0 = lo1; j.nA = (uint32_t)(hi1 - lo1); j.gB0 = two ? lo2 : 0; j.nB = two ? (uint32_t)(hi2 - lo2) : 0;
    j.posB0 = (j.nA + C16_BLK - 1) / C16_BLK * C16_BLK;
    j.KP = ((S + 1) / 2 + C16_KC - 1) / C16_KC * C16_KC;
    j.s_begin = 0; j.s_count = S;
    const uint64_t U = (uint64_t)j.nA + j.nB;
    const uint64_t nblocks = (uint64_t)j.posB0 / C16_BLK + (j.nB + C16_BLK - 1) / C16_BLK + 2;   // +2: a row tile reads two blocks
    const uint64_t code_bytes = nblocks * j.KP * C16_BLK * 4;
    if (int rc = c->c16codes.reserve(code_bytes)) return rc;
    if (int rc = c->c16flag.reserve(256)) return rc;
    int *flag = c->c16flag.as<int>();
    cudaStream_t st = c->stream;
    auto &cc = c->c16cache;
    auto &gl = c->c16g;
    // gt/lt registers with power-of-two S and no raw counts wanted: the != count alone determines the result
    const bool pow2 = (S & (S - 1)) == 0;
    const int mode = (!counts_gtlt(p->cmp_kind) || (p->cmp_kind == D2G_CMP_GTLT && pow2 && !base.c0_out && !getenv("D2G_C16_NO_NE"))) ? 1 : 0;
    const bool cached = cc.valid && cc.regs == base.regs && cc.lo1 == lo1 && cc.hi1 == hi1 && cc.lo2 == lo2 && cc.hi2 == hi2 && cc.S == S && cc.kind == p->cmp_kind && cc.mode == mode;
    if (!cached) {
        KernelTimer kt(c, D2G_T_CMP_PREP);
        cc.valid = true; cc.regs = base.regs; cc.lo1 = lo1; cc.hi1 = hi1; cc.lo2 = lo2; cc.hi2 = hi2; cc.S = S; cc.kind = p->cmp_kind; cc.mode = mode;
        CU(cudaMemsetAsync(c->c16codes.p, 0, code_bytes, st));
        const bool use_hash = mode == 1 && U <= C16_HASH_MAX_SKETCHES && !getenv("D2G_C16_NO_HASH");
        if (use_hash) {
            // != only: injective codes suffice -> open-addressing table per register position, no sort
            const uint32_t TS = (uint32_t)std::max<uint64_t>(64, U + U / 2);
            const size_t smem = (size_t)TS * 8;
            CU(cudaMemsetAsync(flag, 0, 4, st));
            if (counts_gtlt(p->cmp_kind)) {
                CU(cudaFuncSetAttribute(c16_hash_codes_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                c16_hash_codes_kernel<0><<<(S + 3) / 4, 512, smem, st>>>(j, TS, c->c16codes.as<uint16_t>(), flag);
            } else {
                CU(cudaFuncSetAttribute(c16_hash_codes_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                c16_hash_codes_kernel<1><<<(S + 3) / 4, 512, smem, st>>>(j, TS, c->c16codes.as<uint16_t>(), flag);
            }
            c->launches++;
            CU(cudaGetLastError());
        }
        const bool use_global = !use_hash && gl.valid && gl.regs == base.regs && gl.S == S && gl.kind == p->cmp_kind && lo1 >= gl.g0 && hi1 <= gl.g0 + gl.N &&
                                (!two || (lo2 >= gl.g0 && hi2 <= gl.g0 + gl.N));
        if (use_global) {
            const size_t smem = 2 * ((gl.N + 31) / 32) * 4;
            CU(cudaFuncSetAttribute(c16_local_codes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            c16_local_codes_kernel<<<S, 256, smem, st>>>(j, c->c16grank.as<uint32_t>(), gl.g0, (uint32_t)gl.N, c->c16codes.as<uint16_t>());
            c->launches++;
            CU(cudaGetLastError());
        } else if (!use_hash) {
            CU(cudaMemsetAsync(flag, 0, 4, st));
            if (int rc = c16_sort_rank(c, p, j, c->c16codes.as<uint16_t>(), nullptr, flag)) return rc;
        }
    }
    (void)U;
    C16Args a;
    a.codes = c->c16codes.as<uint32_t>(); a.KP = j.KP;
    auto view = [&](uint64_t lo, uint64_t hi, uint64_t rlo, uint32_t rpos0, uint32_t &blk0, uint32_t &n, uint64_t &g0) {
        const uint64_t blk = (lo - rlo) / C16_BLK;
        g0 = rlo + blk * C16_BLK; blk0 = rpos0 / C16_BLK + (uint32_t)blk; n = (uint32_t)(hi - g0);
    };
    view(r0, r1, lo1, 0, a.a_blk0, a.n_a, a.gi0);
    if (two) view(c0, c1, lo2, j.posB0, a.b_blk0, a.n_b, a.gj0);
    else view(c0, c1, lo1, 0, a.b_blk0, a.n_b, a.gj0);
    a.o = base; a.o.row0 = r0; a.o.row1 = r1; a.o.col0 = c0; a.o.col1 = c1; a.o.use_flag = nullptr; a.o.want = 0;
    a.use_flag = flag; a.want = 0;
    const uint64_t tiles_i = (a.n_a + C16_TM - 1) / C16_TM;
    a.tiles_j = (a.n_b + C16_TN - 1) / C16_TN;
    const uint64_t grid = tiles_i * a.tiles_j;
    if (grid > 0x7fffffffULL) return fail(D2G_EINVAL, "comparison job too large for one launch");
    {
        a.ne_is_gt = (mode == 1 && p->cmp_kind == D2G_CMP_GTLT) ? 1 : 0;
        a.one = 1;
        int acc = 1;
        if (const char *ev = getenv("D2G_C16_ACC")) acc = atoi(ev);
        KernelTimer kt(c, D2G_T_CMP);
        auto go = [&](auto kern) -> int {
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C16_SMEM));
            kern<<<(unsigned)grid, C16_THREADS, C16_SMEM, st>>>(a);
            return D2G_OK;
        };
        int rc;
        if (mode == 0) rc = acc == 0 ? go(cmp16_tile_kernel<0, 0>) : go(cmp16_tile_kernel<0, 1>);
        else rc = acc == 0 ? go(cmp16_tile_kernel<1, 0>) : go(cmp16_tile_kernel<1, 1>);
        if (rc) return rc;
        c->launches++;
        CU(cudaGetLastError());
    }
    // registers the codes cannot express (NaN): the f64 kernel recomputes the job, gated on the device flag
    return launch_cmp_f64(c, p, base, r0, r1, c0, c1, flag, 1);
}

// rows [r0,r1) of the output into out_d / counts (packed from row r0)
int launch_cmp(d2g_ctx *c, const d2g_cmp_params *p, const d2g::CmpConsts &k, const double *regs_d, const double *cards_d,
               uint64_t r0, uint64_t r1, float *out_d, uint32_t *c0_d, uint32_t *c1_d, uint64_t sym_lo = ~0ULL) {
    if (r1 <= r0) return D2G_OK;
    d2g::CmpArgs a{};
    a.regs = regs_d; a.cards = cards_d; a.n = p->n; a.out_row0 = r0;
    a.col_base = p->shape == D2G_PANEL ? p->n - p->nq : 0;
    a.ncols = n_cols(p); a.shape = p->shape; a.out = out_d; a.c0_out = c0_d; a.c1_out = c1_d; a.c = k;
    if (a.ncols == 0) return D2G_OK;
    const uint32_t S = p->sketchsize;
    // columns this row range needs; sym_lo lets successive row blocks of one call share one code space
    const uint64_t cb = p->shape == D2G_SYMMETRIC ? std::min(r0, sym_lo) : a.col_base, ce = a.col_base + a.ncols;
    const uint64_t nR = r1 - r0, nC = ce - cb;
    // path choice: codes pay a per-job sort of the registers, worth it from ~4e9 register comparisons on; jobs that only
    // count != and fit the shared-memory table (run_cmp16_job) build their codes in one cheap pass: from ~2e8 on
    const bool ne_only = !counts_gtlt(p->cmp_kind) || (p->cmp_kind == D2G_CMP_GTLT && (S & (S - 1)) == 0 && !c0_d);
    const bool hashable = ne_only && ((cb <= r0 && r1 <= ce) ? nC : nR + nC) <= d2g::C16_HASH_MAX_SKETCHES;
    int path = (double)nR * (double)nC * (double)S >= (hashable ? 2.0e8 : 4.0e9) ? 1 : 0;
    if (const char *ev = getenv("D2G_CMP_PATH")) path = !strcmp(ev, "codes") ? 1 : (!strcmp(ev, "f64") ? 0 : path);
    uint64_t M = 63232;                                                 // sketches per job: <= 63487 ranks, multiple of 128
    if (const char *ev = getenv("D2G_C16_MAXJOB")) M = std::max<uint64_t>(256, std::min<uint64_t>(M, strtoull(ev, nullptr, 10) / 128 * 128));  // test knob
    M = std::min<uint64_t>(M, (0x7fffffffULL / S) / 128 * 128);          // one segmented sort holds < 2^31 items
    if (S > 65535 || M < 256) path = 0;                                  // 16-bit counters / degenerate blocks
    if (!path) return launch_cmp_f64(c, p, a, r0, r1, cb, ce, nullptr, 0);
    auto up64 = [](uint64_t x) { return (x + 63) / 64 * 64; };
    // Decomposition into jobs of at most Mj sketches: one range when the rows lie inside the columns, two ranges
    // otherwise, block pairs when that is too many sketches (diagonal blocks of a symmetric comparison share one range).
    struct Job { uint64_t lo1, hi1, lo2, hi2, r0, r1, c0, c1; };
    auto plan = [&](uint64_t Mj, std::vector<Job> &jobs) {
        jobs.clear();
        if (cb <= r0 && r1 <= ce && nC <= Mj) { jobs.push_back({cb, ce, 0, 0, r0, r1, cb, ce}); return; }
        if (up64(nR) + nC <= Mj) { jobs.push_back({r0, r1, cb, ce, r0, r1, cb, ce}); return; }
        uint64_t BR, BC;
        if (p->shape == D2G_SYMMETRIC) BR = BC = Mj / 2;
        else if (up64(nR) <= Mj / 2) { BR = nR; BC = (Mj - up64(nR)) / 64 * 64; }
        else if (nC <= Mj / 2) { BC = nC; BR = (Mj - nC) / 128 * 128; }
        else BR = BC = Mj / 2;
        for (uint64_t rb = r0; rb < r1; rb += BR) {
            const uint64_t re = std::min(r1, rb + BR);
            for (uint64_t cc = cb; cc < ce; cc += BC) {
                const uint64_t cf = std::min(ce, cc + BC);
                if (p->shape == D2G_SYMMETRIC && cf <= rb + 1) continue;     // block entirely on/below the diagonal
                if (cc == rb && cf == re) jobs.push_back({rb, re, 0, 0, rb, re, cc, cf});
                else jobs.push_back({rb, re, cc, cf, rb, re, cc, cf});
            }
        }
    };
    auto job_sketches = [](const Job &j) { return (j.hi1 - j.lo1) + (j.hi2 - j.lo2); };
    std::vector<Job> jobs, hjobs;
    plan(M, jobs);
    const uint64_t g0 = std::min(r0, cb), g1 = std::max(r1, ce), N = g1 - g0;
    bool use_global = jobs.size() > 1 && 2 * ((N + 31) / 32) * 4 <= 200 * 1024 && N < 0xFFFFFFF0ULL && !getenv("D2G_C16_NO_GLOBAL");   // bitmap + prefix in shared memory
    if (ne_only && !getenv("D2G_C16_NO_HASH") && M > d2g::C16_HASH_MAX_SKETCHES) {
        // != only: jobs of <= 16 384 sketches take their codes from a hash table (no sort).  Pick the cheaper plan with measured
        // per-element costs (B200, S=4096): segmented sort + ranks 1.2e-10 s, local codes from global ranks 0.8e-11 s, hash codes 2.8e-11 s.
        plan(d2g::C16_HASH_MAX_SKETCHES, hjobs);
        double cost_sorted = 0, cost_hashed = 0;
        for (const Job &j : jobs) cost_sorted += (double)job_sketches(j) * S * (use_global ? 0.8e-11 : 1.2e-10);
        if (use_global) cost_sorted += (double)N * S * 1.2e-10;
        if (jobs.size() == 1 && job_sketches(jobs[0]) <= d2g::C16_HASH_MAX_SKETCHES) cost_sorted = 1e30;   // hashed inside run_cmp16_job anyway
        for (const Job &j : hjobs) cost_hashed += (double)job_sketches(j) * S * 2.8e-11 + 15e-6;
        if (cost_hashed < cost_sorted) { jobs.swap(hjobs); use_global = false; }
    }
    if (use_global) { if (int rc = c16_build_global(c, p, regs_d, g0, N)) return rc; }
    for (const Job &j : jobs)
        if (int rc = run_cmp16_job(c, p, a, j.lo1, j.hi1, j.lo2, j.hi2, j.r0, j.r1, j.c0, j.c1)) return rc;
    return D2G_OK;
}

} // namespace

extern "C" int d2g_make_compressed(const double *regs, const uint64_t *kmers, uint64_t n, uint32_t S, double fd, int32_t bbit,
                                   long double *a_io, long double *b_io, double *out, int32_t *bbit_used) {
    if (!regs || !out || !a_io || !b_io) return fail(D2G_EINVAL, "null argument");
    if (fd != 1. && fd != 2. && fd != 4.) return fail(D2G_EINVAL, "regbytes must be 1, 2 or 4 (got %g)", fd);
    const uint64_t nsigs = n * S;
    long double a = *a_io, b = *b_io;
    auto parallel = [&](auto fn) {
        unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 64u));
        if (nsigs < (1u << 16)) nt = 1;
        std::vector<std::thread> th;
        const uint64_t per = (nsigs + nt - 1) / nt;
        for (unsigned t = 0; t < nt; ++t) { const uint64_t lo = t * per, hi = std::min(nsigs, lo + per); if (lo < hi) th.emplace_back(fn, lo, hi); }
        for (auto &x : th) x.join();
    };
    if (!bbit) {
        const long double q = fd == 1. ? 254.3 : fd == 2. ? 65534 : 4294967294;   // double literals, as in the reference (cmp_core.cpp:249)
        if (a <= 0. || b <= 0.) {                       // cmp_core.cpp:250-266
            double minreg = __DBL_MAX__, maxreg = -__DBL_MAX__;
            for (uint64_t i = 0; i < nsigs; ++i) {
                const double v = regs[i];
                if (v <= 0 || v == __DBL_MAX__) continue;
                minreg = std::min(minreg, v); maxreg = std::max(maxreg, v);
            }
            long double mx = minreg, mn = maxreg;       // optimal_parameters(minreg, maxreg, q), src/setsketch.h:563-566
            if (mx < mn) std::swap(mx, mn);
            b = expl(logl(mx / mn) / q);                // src/setsketch.cpp:7-10
            a = mx / b;
        }
        if (a == 0. || isinf(b)) bbit = 1;              // cmp_core.cpp:267-270
        else {
            *a_io = a; *b_io = b;
            const long double logbinv = 1.L / log1pl(b - 1.L);
            const int64_t top = (int64_t)(q + 1);
            parallel([&](uint64_t lo, uint64_t hi) {
                for (uint64_t i = lo; i < hi; ++i) {
                    const long double sub = 1.L - logl((long double)regs[i] / a) * logbinv;
                    // static_cast<int64_t>(long double) as x86 executes it: out of range / NaN -> INT64_MIN
                    int64_t isub = (sub > -9223372036854775809.0L && sub < 9223372036854775808.0L) ? (int64_t)sub : INT64_MIN;
                    out[i] = (double)std::max<int64_t>(0, std::min(top, isub));
                }
            });
            if (bbit_used) *bbit_used = 0;
            return D2G_OK;
        }
    }
    const int shift = fd == 1. ? 58 : fd == 2. ? 48 : 32;   // cmp_core.cpp:306-320
    parallel([&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            uint64_t v;
            if (kmers) v = d2g::wang64(kmers[i]);
            else { memcpy(&v, regs + i, 8); v = d2g::wang64(v ^ 0xa3407fb23cd20efULL); }   // reg2sig, cmp_core.cpp:19-37
            out[i] = (double)(v >> shift);
        }
    });
    if (bbit_used) *bbit_used = 1;
    return D2G_OK;
}

extern "C" {

uint64_t d2g_cmp_output_size(const d2g_cmp_params *p) { return p ? rows_size(p, 0, n_rows(p)) : 0; }

int d2g_cmp_rows_size(const d2g_cmp_params *p, uint64_t r0, uint64_t r1, uint64_t *n_vals) {
    if (int rc = check_cmp_params(p)) return rc;
    if (r0 > r1 || r1 > n_rows(p)) return fail(D2G_EINVAL, "bad row range");
    if (n_vals) *n_vals = rows_size(p, r0, r1);
    return D2G_OK;
}

int d2g_densify_dev(d2g_ctx *c, double *sig_d, uint64_t *kmers_d, uint64_t n, uint32_t S) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (!n || !S) return D2G_OK;
    CU(cudaSetDevice(c->device));
    const uint64_t tot = n * S;
    if (int rc = c->ctmp.reserve(tot * 8)) return rc;
    if (kmers_d) if (int rc = c->cktmp.reserve(tot * 8)) return rc;
    d2g::densify_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(sig_d, kmers_d, n, S, c->ctmp.as<double>(), c->cktmp.as<uint64_t>());
    c->launches++;
    CU(cudaMemcpyAsync(sig_d, c->ctmp.p, tot * 8, cudaMemcpyDeviceToDevice, c->stream));
    if (kmers_d) CU(cudaMemcpyAsync(kmers_d, c->cktmp.p, tot * 8, cudaMemcpyDeviceToDevice, c->stream));
    return D2G_OK;
}

int d2g_densify(d2g_ctx *c, double *sig, uint64_t *kmers, uint64_t n, uint32_t S) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (!n || !S) return D2G_OK;
    CU(cudaSetDevice(c->device));
    const uint64_t tot = n * S;
    if (int rc = c->cregs.reserve(tot * 8)) return rc;
    CU(cudaMemcpyAsync(c->cregs.p, sig, tot * 8, cudaMemcpyHostToDevice, c->stream));
    if (kmers) { if (int rc = c->ids.reserve(tot * 8)) return rc; CU(cudaMemcpyAsync(c->ids.p, kmers, tot * 8, cudaMemcpyHostToDevice, c->stream)); }
    if (int rc = d2g_densify_dev(c, c->cregs.as<double>(), kmers ? c->ids.as<uint64_t>() : nullptr, n, S)) return rc;
    CU(cudaMemcpyAsync(sig, c->cregs.p, tot * 8, cudaMemcpyDeviceToHost, c->stream));
    if (kmers) CU(cudaMemcpyAsync(kmers, c->ids.p, tot * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return D2G_OK;
}

int d2g_cmp_rows_dev(d2g_ctx *c, const d2g_cmp_params *p, const double *regs_d, const double *cards_d,
                     uint64_t r0, uint64_t r1, float *out_d) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_cmp_params(p)) return rc;
    if (r0 > r1 || r1 > n_rows(p)) return fail(D2G_EINVAL, "bad row range");
    CU(cudaSetDevice(c->device));
    c->c16cache.valid = false; c->c16g.valid = false;
    d2g::CmpConsts k;
    if (int rc = make_consts(c, p, &k)) return rc;
    return launch_cmp(c, p, k, regs_d, cards_d, r0, r1, out_d, nullptr, nullptr);
}

// Rows [r0,r1) in row blocks of <= ~64M values.  Kernels run on the ctx stream, the device->host copies on the copy
// stream (block b+1 computes while block b drains).  direct_out != nullptr: results go straight into the caller's
// buffer (full speed when it is pinned); otherwise through two pinned staging buffers to the sink, in row order.
static int cmp_blocks(d2g_ctx *c, const d2g_cmp_params *p, const double *regs, const double *cards,
                      uint64_t r0, uint64_t r1, d2g_sink_fn sink, void *user, float *direct_out) {
    CU(cudaSetDevice(c->device));
    c->c16cache.valid = false; c->c16g.valid = false;
    const uint32_t S = p->sketchsize;
    if (int rc = c->cregs.reserve(p->n * S * 8)) return rc;
    if (int rc = c->ccards.reserve(p->n * 8)) return rc;
    CU(cudaMemcpyAsync(c->cregs.p, regs, p->n * S * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->ccards.p, cards, p->n * 8, cudaMemcpyHostToDevice, c->stream));
    d2g::CmpConsts k;
    if (int rc = make_consts(c, p, &k)) return rc;
    const uint64_t max_vals = 64ULL << 20;
    const uint64_t ncol = n_cols(p);
    uint64_t rows_per = std::max<uint64_t>(d2g::CMP_T, max_vals / std::max<uint64_t>(1, ncol) / d2g::CMP_T * d2g::CMP_T);
    uint64_t cap_vals = 0;
    for (uint64_t b = r0; b < r1; b += rows_per) cap_vals = std::max(cap_vals, rows_size(p, b, std::min(r1, b + rows_per)));
    if (int rc = c->cout.reserve(2 * cap_vals * 4)) return rc;
    if (!direct_out) {
        if (int rc = c->pin[0].reserve(cap_vals * 4)) return rc;
        if (int rc = c->pin[1].reserve(cap_vals * 4)) return rc;
    }
    struct Pending { uint64_t b0, b1, nv; int slot; bool live; } pend{0, 0, 0, 0, false};
    uint64_t done_vals = 0, iblk = 0;
    for (uint64_t b = r0; b < r1; b += rows_per, ++iblk) {
        const int slot = (int)(iblk & 1);
        const uint64_t e = std::min(r1, b + rows_per), nv = rows_size(p, b, e);
        float *out_d = c->cout.as<float>() + (uint64_t)slot * cap_vals;
        if (iblk >= 2) CU(cudaStreamWaitEvent(c->stream, c->evd[slot], 0));       // the copy of block b-2 has left this slot
        if (int rc = launch_cmp(c, p, k, c->cregs.as<double>(), c->ccards.as<double>(), b, e, out_d, nullptr, nullptr, r0)) return rc;
        CU(cudaEventRecord(c->ev[slot], c->stream));
        CU(cudaStreamWaitEvent(c->copy_stream, c->ev[slot], 0));
        float *dst = direct_out ? direct_out + done_vals : (float *)c->pin[slot].p;
        CU(cudaMemcpyAsync(dst, out_d, nv * 4, cudaMemcpyDeviceToHost, c->copy_stream));
        CU(cudaEventRecord(c->evd[slot], c->copy_stream));
        done_vals += nv;
        if (!direct_out) {
            if (pend.live) {
                CU(cudaEventSynchronize(c->evd[pend.slot]));
                if (sink(user, (const float *)c->pin[pend.slot].p, pend.b0, pend.b1 - pend.b0, pend.nv)) return fail(D2G_EIO, "sink aborted");
            }
            pend = {b, e, nv, slot, true};
        }
    }
    if (pend.live) {
        CU(cudaEventSynchronize(c->evd[pend.slot]));
        if (sink(user, (const float *)c->pin[pend.slot].p, pend.b0, pend.b1 - pend.b0, pend.nv)) return fail(D2G_EIO, "sink aborted");
    }
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->stream));
    return D2G_OK;
}

int d2g_cmp_stream(d2g_ctx *c, const d2g_cmp_params *p, const double *regs, const double *cards,
                   uint64_t r0, uint64_t r1, d2g_sink_fn sink, void *user) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_cmp_params(p)) return rc;
    if (r0 > r1 || r1 > n_rows(p)) return fail(D2G_EINVAL, "bad row range");
    if (!sink) return fail(D2G_EINVAL, "null sink");
    if (r0 == r1 || p->n == 0) return D2G_OK;
    return cmp_blocks(c, p, regs, cards, r0, r1, sink, user, nullptr);
}

int d2g_cmp_rows(d2g_ctx *c, const d2g_cmp_params *p, const double *regs, const double *cards,
                 uint64_t r0, uint64_t r1, float *out) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_cmp_params(p)) return rc;
    if (r0 > r1 || r1 > n_rows(p)) return fail(D2G_EINVAL, "bad row range");
    if (r0 == r1 || p->n == 0) return D2G_OK;
    if (!out) return fail(D2G_EINVAL, "null output");
    return cmp_blocks(c, p, regs, cards, r0, r1, nullptr, nullptr, out);
}

int d2g_cmp_matrix(d2g_ctx *c, const d2g_cmp_params *p, const double *regs, const double *cards, float *out) {
    if (int rc = check_cmp_params(p)) return rc;
    return d2g_cmp_rows(c, p, regs, cards, 0, n_rows(p), out);
}

int d2g_cmp_counts(d2g_ctx *c, uint32_t S, int32_t cmp_kind, const double *rows, uint64_t nr, const double *cols, uint64_t nc,
                   uint32_t *c0_out, uint32_t *c1_out) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (!nr || !nc) return D2G_OK;
    CU(cudaSetDevice(c->device));
    d2g_cmp_params p{};
    p.sketchsize = S; p.cmp_kind = cmp_kind; p.measure = D2G_SIMILARITY; p.k = 31; p.shape = D2G_PANEL; p.n = nr + nc; p.nq = nc;
    if (int rc = check_cmp_params(&p)) return rc;
    c->c16cache.valid = false; c->c16g.valid = false;
    if (int rc = c->cregs.reserve(p.n * S * 8)) return rc;
    CU(cudaMemcpyAsync(c->cregs.p, rows, nr * S * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->cregs.as<double>() + nr * S, cols, nc * S * 8, cudaMemcpyHostToDevice, c->stream));
    if (int rc = c->cout.reserve(2 * nr * nc * 4)) return rc;
    uint32_t *c0_d = c->cout.as<uint32_t>(), *c1_d = c0_d + nr * nc;
    d2g::CmpConsts k;
    if (int rc = make_consts(c, &p, &k)) return rc;
    if (int rc = launch_cmp(c, &p, k, c->cregs.as<double>(), nullptr, 0, nr, nullptr, c0_d, cmp_kind == D2G_CMP_GTLT ? c1_d : nullptr)) return rc;
    if (c0_out) CU(cudaMemcpyAsync(c0_out, c0_d, nr * nc * 4, cudaMemcpyDeviceToHost, c->stream));
    if (c1_out && cmp_kind == D2G_CMP_GTLT) CU(cudaMemcpyAsync(c1_out, c1_d, nr * nc * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return D2G_OK;
}

} // extern "C"

// -------------------------------------------------------------------------------------------------
// LSH top-k
// -------------------------------------------------------------------------------------------------
extern "C" int d2g_lsh_topk(d2g_ctx *c, const d2g_cmp_params *p, const double *regs, const double *cards, int32_t topk,
                            uint64_t *indptr_out, uint32_t **idx_out, float **val_out) {
    return d2g_lsh_topk_rows(c, p, regs, cards, topk, 0, p ? p->n : 0, indptr_out, idx_out, val_out);
}

extern "C" int d2g_lsh_topk_rows(d2g_ctx *c, const d2g_cmp_params *p, const double *regs, const double *cards, int32_t topk,
                                 uint64_t x0, uint64_t x1, uint64_t *indptr_out, uint32_t **idx_out, float **val_out) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_cmp_params(p)) return rc;
    if (x0 > x1 || x1 > p->n) return fail(D2G_EINVAL, "bad row range");
    if (topk <= 0) return fail(D2G_EINVAL, "topk must be > 0 (similarity-threshold graphs are not implemented)");
    if (p->cmp_kind >= D2G_CMP_SS_COMPRESSED) return fail(D2G_EUNSUPPORTED, "top-k over compressed registers (--fastcmp with --topk) is not implemented on the GPU");
    if (!indptr_out || !idx_out || !val_out) return fail(D2G_EINVAL, "null output");
    const uint64_t n = p->n; const uint32_t S = p->sketchsize;
    if (n < 2 || x0 == x1) { for (uint64_t i = 0; i <= x1 - x0; ++i) indptr_out[i] = 0; *idx_out = (uint32_t *)malloc(4); *val_out = (float *)malloc(4); return D2G_OK; }
    if (n >= 0x7FFFFFFFULL) return fail(D2G_EINVAL, "too many sketches for 32-bit LSH ids");
    if (S < 2) return fail(D2G_EINVAL, "sketchsize must be >= 2 for the default two LSH table types");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    if (p->nlsh < 0 || p->nlsh > 3)
        return fail(D2G_EUNSUPPORTED, "--nLSH %d: 1, 2 (the default) and 3 are implemented; 4 and up add six-register tables keyed by XXH3 (src/ssi.h:345-352)", p->nlsh);
    if (p->nlsh == 3 && S < 4) return fail(D2G_EINVAL, "--nLSH 3 needs at least four registers");
    const uint32_t n1 = p->nlsh == 1 ? 0 : S / 2;                      // two-register tables
    const uint32_t n2 = p->nlsh == 3 ? (uint32_t)((uint64_t)S * 8 / 4) : 0;   // four-register tables: 8S / 4 (cmp_core.cpp:767)
    const uint32_t ntab = S + n1 + n2;                                 // cmp_core.cpp:757-770
    uint64_t ntoquery = (uint64_t)((float)topk * 3.5f);                // index_build.cpp:57-60
    ntoquery = std::min<uint64_t>(ntoquery, n - 1);
    CU(cudaFuncSetAttribute(d2g::lsh_trim_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, d2g::LSH_TRIM_BIG_CAP * 8));
    if (ntoquery == 0 || ntoquery > 4096) return fail(D2G_EUNSUPPORTED, "topk %d out of the supported range", topk);
    const uint32_t maxcand = (uint32_t)ntoquery;
    if (int rc = c->cregs.reserve(n * S * 8)) return rc;
    if (int rc = c->ccards.reserve(n * 8)) return rc;
    CU(cudaMemcpyAsync(c->cregs.p, regs, n * S * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->ccards.p, cards, n * 8, cudaMemcpyHostToDevice, st));
    const double *regs_d = c->cregs.as<double>(), *cards_d = c->ccards.as<double>();
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
        d2g::lsh_keys_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(regs_d, n, S, t0, nt, kA + (uint64_t)t0 * n, iA + (uint64_t)t0 * n, n1);
        size_t need = 0;
        cub::DeviceSegmentedRadixSort::SortPairs(nullptr, need, kA, kB, iA, iB, (int)items, (int)nt, offs, offs + 1, 0, 32, st);
        if (need > tb) { tb = need; if (int rc = c->wtmp.reserve(tb + 256)) return rc; }
        size_t tbytes = tb;
        CU(cub::DeviceSegmentedRadixSort::SortPairs(c->wtmp.p, tbytes, kA + (uint64_t)t0 * n, kB + (uint64_t)t0 * n, iA + (uint64_t)t0 * n, iB + (uint64_t)t0 * n,
                                                   (int)items, (int)nt, offs, offs + 1, 0, 32, st));
        c->launches += 6;
    }
    // 3. ordered candidate scan, one warp per query
    {
        const int wpb = 4;
        const size_t smem = (size_t)wpb * 2 * maxcand * 4;
        KernelTimer kt(c, D2G_T_CMP);
        d2g::lsh_query_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, smem, st>>>(regs_d, n, S, kB, iB, maxcand, cand, cnt, ncand, n1, n2);
        c->launches++;
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
    {
        const bool warp_ok = maxcand <= (uint32_t)d2g::LSH_REPLAY_DCAP;
        if (warp_ok) d2g::lsh_replay_warp_kernel<<<(unsigned)((n + d2g::LSH_REPLAY_WARPS - 1) / d2g::LSH_REPLAY_WARPS), d2g::LSH_REPLAY_WARPS * 32, 0, st>>>(apB, seg, n, maxcand, lst, lsz);
        else d2g::lsh_replay_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(apB, seg, n, maxcand, 0u, lst, dset, lsz);
    }
    c->launches += 8;
    uint32_t h_total = 0;
    CU(cudaMemcpyAsync(&h_total, seg + n, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
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
    if (h_total) {
        const uint64_t threads = (uint64_t)h_total * 32;
        if (p->cmp_kind == D2G_CMP_GTLT) d2g::lsh_refine_kernel<0><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(regs_d, cards_d, n, seg, lsz, lst, k, mult);
        else d2g::lsh_refine_kernel<1><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(regs_d, cards_d, n, seg, lsz, lst, k, mult);
        c->launches++;
    }
    // short lists first: a list the second kernel has trimmed (> LSH_TRIM_CAP entries before) must not be seen as short afterwards
    d2g::lsh_trim_kernel<<<(unsigned)((n + d2g::LSH_TRIM_WARPS - 1) / d2g::LSH_TRIM_WARPS), d2g::LSH_TRIM_WARPS * 32, 0, st>>>(seg, n, (uint32_t)topk, is_dist, lst, lsz);
    d2g::lsh_trim_big_kernel<<<(unsigned)n, d2g::LSH_TRIM_BIG_THREADS, d2g::LSH_TRIM_BIG_CAP * 8, st>>>(seg, n, (uint32_t)topk, is_dist, lst, lsz);
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
