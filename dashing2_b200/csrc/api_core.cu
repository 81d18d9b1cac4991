// api_core.cu -- context lifecycle, error reporting and kernel timing of the C ABI declared in include/d2gpu.h.
//
// There is deliberately no CPU implementation of the hot paths in this library: without a device
// every entry point fails.
#include "api_internal.h"

namespace { thread_local std::string g_err; }

int d2g_fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

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
    d2g_comm_destroy(c);
    if (c->ev[0]) cudaEventDestroy(c->ev[0]);
    if (c->ev[1]) cudaEventDestroy(c->ev[1]);
    if (c->evd[0]) cudaEventDestroy(c->evd[0]);
    if (c->evd[1]) cudaEventDestroy(c->evd[1]);
    for (auto &e : c->stage_free) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

void *d2g_stream(d2g_ctx *c) { return c ? (void *)c->stream : nullptr; }
int d2g_sync(d2g_ctx *c) { if (!c) return fail(D2G_EINVAL, "null ctx"); CU(cudaSetDevice(c->device)); CU(cudaStreamSynchronize(c->stream)); return D2G_OK; }
uint64_t d2g_launch_count(const d2g_ctx *c) { return c ? c->launches.load() : 0; }
uint64_t d2g_stat(const d2g_ctx *c, int which) { return (c && which >= 0 && which < D2G_STAT_NSTATS) ? c->stats[which] : 0; }
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
