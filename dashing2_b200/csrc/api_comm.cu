// api_comm.cu -- the communicator a context owns (NCCL over NVLink / NVSwitch) and the entry points that create it.
//
// The reference is a single-process OpenMP program; its all-pairs phase shards over output rows (src/emitrect.cpp:198-326).  Here the
// rows shard over GPUs and the one exchange step of the path -- every GPU needs an order code for every register of every sketch --
// runs over NCCL inside the library (api_cmp.cu: d2g_cmp_rows_sharded_dev).  libnccl.so.2 is resolved with dlopen when a communicator
// is first asked for, so the library has no link-time dependency on it and shares the copy a host program (PyTorch) may have loaded.
#include "api_internal.h"
#include "nccl_dl.h"
#include <dlfcn.h>
#include <mutex>

namespace d2g_nccl {
Api g_api;
namespace {
std::once_flag g_once;
int g_rc = D2G_OK;
std::string g_msg;
template <class F> bool sym(void *h, const char *name, F &out) { out = reinterpret_cast<F>(dlsym(h, name)); return out != nullptr; }
}
int load() {
    std::call_once(g_once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { g_rc = D2G_EUNSUPPORTED; g_msg = std::string("libnccl.so.2 not found: ") + dlerror(); return; }
        Api &a = g_api;
        const bool ok = sym(h, "ncclGetUniqueId", a.GetUniqueId) && sym(h, "ncclCommInitRank", a.CommInitRank) && sym(h, "ncclCommInitAll", a.CommInitAll) &&
                        sym(h, "ncclCommDestroy", a.CommDestroy) && sym(h, "ncclAllGather", a.AllGather) && sym(h, "ncclAllReduce", a.AllReduce) &&
                        sym(h, "ncclSend", a.Send) && sym(h, "ncclRecv", a.Recv) && sym(h, "ncclGroupStart", a.GroupStart) && sym(h, "ncclGroupEnd", a.GroupEnd) &&
                        sym(h, "ncclGetErrorString", a.GetErrorString);
        if (!ok) { g_rc = D2G_EUNSUPPORTED; g_msg = "libnccl.so.2 lacks a required symbol"; }
    });
    if (g_rc) return fail(g_rc, "%s", g_msg.c_str());
    return D2G_OK;
}
} // namespace d2g_nccl

#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(D2G_ECUDA, "%s failed: %s", #call, d2g_nccl::g_api.GetErrorString(r_)); } while (0)

extern "C" {

int d2g_comm_unique_id(void *id_out) {
    if (!id_out) return fail(D2G_EINVAL, "null id buffer");
    if (int rc = d2g_nccl::load()) return rc;
    ncclUniqueId id;
    NC(d2g_nccl::g_api.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    static_assert(sizeof(ncclUniqueId) == D2G_COMM_ID_BYTES, "unique id size");
    return D2G_OK;
}

int d2g_comm_init_rank(d2g_ctx *c, int nranks, int rank, const void *id_bytes) {
    if (!c || !id_bytes) return fail(D2G_EINVAL, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(D2G_EINVAL, "bad rank %d of %d", rank, nranks);
    if (c->nccl_comm) return fail(D2G_EINVAL, "this context already owns a communicator");
    if (int rc = d2g_nccl::load()) return rc;
    CU(cudaSetDevice(c->device));
    ncclUniqueId id; memcpy(&id, id_bytes, sizeof id);
    ncclComm_t comm = nullptr;
    NC(d2g_nccl::g_api.CommInitRank(&comm, nranks, id, rank));
    c->nccl_comm = comm; c->nranks = nranks; c->rank = rank;
    return D2G_OK;
}

// `d2g_init(ctx**, const int *devices, int ndev)` -- one process that owns several devices and their communicator.
int d2g_init_devices(d2g_ctx **ctxs, const int *devices, int ndev) {
    if (!ctxs || !devices || ndev < 1) return fail(D2G_EINVAL, "bad device list");
    for (int i = 0; i < ndev; ++i) ctxs[i] = nullptr;
    for (int i = 0; i < ndev; ++i)
        if (int rc = d2g_init(&ctxs[i], devices[i])) { for (int j = 0; j < i; ++j) { d2g_destroy(ctxs[j]); ctxs[j] = nullptr; } return rc; }
    if (ndev == 1) return D2G_OK;
    return d2g_comm_init_all(ctxs, ndev);
}

int d2g_comm_init_all(d2g_ctx **ctxs, int n) {
    if (!ctxs || n < 1) return fail(D2G_EINVAL, "bad context list");
    if (int rc = d2g_nccl::load()) return rc;
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i]) return fail(D2G_EINVAL, "null context %d", i);
        if (ctxs[i]->nccl_comm) return fail(D2G_EINVAL, "context %d already owns a communicator", i);
        devs[i] = ctxs[i]->device;
        for (int j = 0; j < i; ++j) if (devs[j] == devs[i]) return fail(D2G_EINVAL, "contexts %d and %d share device %d: a communicator needs distinct devices", j, i, devs[i]);
    }
    std::vector<ncclComm_t> comms(n, nullptr);
    NC(d2g_nccl::g_api.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) { ctxs[i]->nccl_comm = comms[i]; ctxs[i]->nranks = n; ctxs[i]->rank = i; }
    return D2G_OK;
}

int d2g_comm_size(const d2g_ctx *c) { return c ? c->nranks : 0; }
int d2g_comm_rank(const d2g_ctx *c) { return c ? c->rank : 0; }

int d2g_comm_destroy(d2g_ctx *c) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (c->nccl_comm) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        d2g_nccl::g_api.CommDestroy(static_cast<ncclComm_t>(c->nccl_comm));
        c->nccl_comm = nullptr; c->nranks = 1; c->rank = 0;
    }
    return D2G_OK;
}

} // extern "C"
