// api_internal.h -- host-side runtime shared by the translation units that implement include/d2gpu.h:
// context (device, streams, grow-only scratch), error reporting, kernel timing.  Not part of the ABI.
#pragma once
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

int d2g_fail(int code, const char *fmt, ...);
#define fail d2g_fail
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

struct d2g_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    std::atomic<uint64_t> launches{0};
    DevBuf seq, pcodes, pmask, recoff, recent, regs, sig, card, ids, aux, aux2, redo, ovfq;   // sketch scratch (pcodes / pmask: the packed batch)
    PinBuf stage[3]; cudaEvent_t stage_free[3] = {nullptr, nullptr, nullptr};        // pinned staging ring of the host packer
    DevBuf items, itemcnt, itemoff, segs, nseg, sflag, sexcl, svmask, slut, stmp;   // element streams (api_stream.cu)
    DevBuf filter; uint64_t filter_n = 0;                                // --filterset: sorted hashed k-mers (d2g_set_filterset*)
    DevBuf lregs;                                                        // top-k over compressed registers: what refinement compares
    DevBuf wbuf, wtmp, lbuf;                                              // counting scratch (BagMinHash / ProbMinHash)
    DevBuf cregs, ccards, cout, clut, clut80, ctmp, cktmp;         // compare scratch
    DevBuf c16buf, c16codes, c16grank, c16flag;                    // order-code compare scratch (keys, sort buffers, codes, global ranks)
    struct { bool valid = false; const double *regs = nullptr; uint64_t g0 = 0, N = 0; uint32_t S = 0; int kind = 0; } c16g;   // global ranks built earlier in the same API call
    struct { bool valid = false; const double *regs = nullptr; uint64_t lo1 = 0, hi1 = 0, lo2 = 0, hi2 = 0; uint32_t S = 0; int kind = 0, mode = 0; } c16cache; // codes built earlier in the same API call
    PinBuf pin[2];
    cudaEvent_t ev[2] = {nullptr, nullptr}, evd[2] = {nullptr, nullptr};
    uint32_t lut_S = 0; int lut_k = -1;
    bool timing = false;
    uint64_t stats[D2G_STAT_NSTATS] = {0};
    void *nccl_comm = nullptr; int nranks = 1, rank = 0;        // communicator this context owns (api_comm.cu), null = single device
    bool c16_sharded = false;                                    // compare launcher: order codes come from gathered global ranks only (api_cmp.cu)
    DevBuf xsend, xrecv, xrank, xcards;                          // sharded compare: register exchange, local rank slice, gathered cardinalities
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tev[D2G_T_NCLASSES];
};

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

// ---- shared between the sketch translation units ----
#include "common.cuh"
#include "pack_kernels.cuh"
namespace d2g { struct SketchArgs; }
struct SketchRange { uint64_t pos_base, pos_end; uint32_t ent_base; };
int check_sketch_params(const d2g_sketch_params *p);
int launch_weighted(d2g_ctx *c, const d2g_sketch_params *p, const d2g::PackedSeq &seq_d, const uint64_t *rec_off_d, const uint32_t *rec_ent_d,
                    uint64_t n_rec, uint32_t n_ent, uint64_t total_len, double *sig_d, double *card_d, uint64_t *ids_d);
int launch_opmh_mincount(d2g_ctx *c, const d2g_sketch_params *p, const d2g::PackedSeq &seq_d, const uint64_t *rec_off_d, const uint32_t *rec_ent_d,
                         uint64_t n_rec, uint32_t n_ent, uint64_t total_len, uint64_t *regs_d);
// Element streams (stream_kernels.cuh: k > 32, -C with a window, protein alphabets).  When p selects one, prepare_stream builds the
// items on the ctx stream and returns the view the ordinary launchers run over: PackedSeq with items set, item-region offsets in
// place of rec_off, (k, w) = (1, w-k+1).  ascii_d: the record bytes on the device (protein only; the packed arrays otherwise).
struct StreamView { d2g::PackedSeq seq; const uint64_t *rec_off_d; uint64_t total_len; d2g_sketch_params p; };
bool is_stream_mode(const d2g_sketch_params *p);
int prepare_stream(d2g_ctx *c, const d2g_sketch_params *p, const d2g::PackedSeq &seq_d, const uint8_t *ascii_d, const uint64_t *rec_off_d,
                   uint64_t n_rec, uint64_t total_len, StreamView *out);
static __global__ void fill_u64_kernel(uint64_t *p, uint64_t n, uint64_t v) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---- shared between the compare translation units ----
namespace d2g { struct CmpConsts; }
int check_cmp_params(const d2g_cmp_params *p);
int make_consts(d2g_ctx *c, const d2g_cmp_params *p, d2g::CmpConsts *k);
