// dpe_comm.cu -- multi-GPU plumbing of libdpe_b200: one NCCL communicator per context, every
// collective enqueued on the context's stream between its kernels (include/dpe_b200.h, "multi-GPU").
//
// NCCL is not linked: libnccl.so.2 is dlopen()ed on the first dpe_comm_* call, so single-GPU use
// (and the CPU-side symbol checks) need no NCCL at all, and a process that already carries an NCCL
// (PyTorch bundles one under the same soname) shares that copy instead of loading a second one.
// The reference has no multi-GPU path to cite (no cudaSetDevice, no collective anywhere in it):
// this extends BatchCorrManifold::Update (batchcorrmanifold.cu:2501-2635) over the GPUs of a box.
#include <dlfcn.h>
#include <nccl.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "dpe_internal.cuh"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;   // optional (>= 2.14)
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

template <typename F>
bool sym(void* h, const char* name, F* out) {
    *out = reinterpret_cast<F>(dlsym(h, name));
    return *out != nullptr;
}

// returns nullptr on success, else the reason
const char* nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.ok) return nullptr;
    static char why[256];
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        snprintf(why, sizeof(why), "dlopen(libnccl.so.2) failed: %s", dlerror());
        return why;
    }
    NcclApi a;
    a.handle = h;
    if (!(sym(h, "ncclGetVersion", &a.GetVersion) && sym(h, "ncclGetUniqueId", &a.GetUniqueId) &&
          sym(h, "ncclCommInitRank", &a.CommInitRank) && sym(h, "ncclCommDestroy", &a.CommDestroy) &&
          sym(h, "ncclBroadcast", &a.Broadcast) && sym(h, "ncclAllGather", &a.AllGather) &&
          sym(h, "ncclGetErrorString", &a.GetErrorString))) {
        snprintf(why, sizeof(why), "libnccl.so.2 lacks a required symbol");
        return why;
    }
    sym(h, "ncclCommInitRankConfig", &a.CommInitRankConfig);
    a.ok = true;
    g_nccl = a;
    return nullptr;
}

#define DPE_NCCL(call)                                                                       \
    do {                                                                                     \
        ncclResult_t r__ = (call);                                                           \
        if (r__ != ncclSuccess) {                                                            \
            dpe::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
            return DPE_ECOMM;                                                                \
        }                                                                                    \
    } while (0)

}  // namespace

namespace dpe {

int comm_broadcast(dpe_ctx* c, void* buf, size_t bytes, cudaStream_t s) {
    DPE_NCCL(g_nccl.Broadcast(buf, buf, bytes, ncclChar, 0, (ncclComm_t)c->comm, s));
    c->launches++;
    return DPE_OK;
}

int comm_allgather(dpe_ctx* c, const double* send, double* recv, size_t count, cudaStream_t s) {
    DPE_NCCL(g_nccl.AllGather(send, recv, count, ncclDouble, (ncclComm_t)c->comm, s));
    c->launches++;
    return DPE_OK;
}

}  // namespace dpe

extern "C" {

int dpe_comm_get_unique_id(void* id) {
    if (!id) { dpe::set_error("dpe_comm_get_unique_id: null argument"); return DPE_EINVAL; }
    if (const char* why = nccl_load()) { dpe::set_error("%s", why); return DPE_ECOMM; }
    static_assert(sizeof(ncclUniqueId) == DPE_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    DPE_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return DPE_OK;
}

int dpe_comm_init(dpe_ctx* c, int nranks, int rank, const void* id) {
    if (!c || !id) { dpe::set_error("dpe_comm_init: null argument"); return DPE_EINVAL; }
    if (nranks < 1 || rank < 0 || rank >= nranks) {
        dpe::set_error("dpe_comm_init: rank %d of %d", rank, nranks);
        return DPE_EINVAL;
    }
    if (c->comm) { dpe::set_error("dpe_comm_init: the context already has a communicator"); return DPE_ESTATE; }
    if (c->inflight) { dpe::set_error("dpe_comm_init: an epoch is in flight"); return DPE_ESTATE; }
    if (const char* why = nccl_load()) { dpe::set_error("%s", why); return DPE_ECOMM; }
    dpe::DevGuard g(c->cfg.device);
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t comm = nullptr;
    // one CTA per collective: the packet (200 kB - 800 kB) and the 128-byte partials need no more, and a collective
    // must fit on the single SM k_brute leaves free (comm_reserve_sms below) to run UNDER it
    const char* mc = getenv("DPE_COMM_MAX_CTAS");
    const int max_ctas = mc ? atoi(mc) : 1;
    if (g_nccl.CommInitRankConfig && max_ctas > 0) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.minCTAs = 1;
        cfg.maxCTAs = max_ctas;
        DPE_NCCL(g_nccl.CommInitRankConfig(&comm, nranks, u, rank, &cfg));
    } else {
        DPE_NCCL(g_nccl.CommInitRank(&comm, nranks, u, rank));
    }
    double* gathered = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&gathered), sizeof(double) * dpe::kPartialLen * nranks);
    if (e != cudaSuccess) {
        g_nccl.CommDestroy(comm);
        dpe::set_error("cudaMalloc(gathered) -> %s", cudaGetErrorString(e));
        return DPE_ENOMEM;
    }
    cudaMemset(gathered, 0, sizeof(double) * dpe::kPartialLen * nranks);
    const char* rs = getenv("DPE_COMM_RESERVE_SMS");
    c->comm_reserve_sms = rs ? atoi(rs) : 1;
    if (c->comm_reserve_sms < 0 || c->comm_reserve_sms > c->sm_count / 2) c->comm_reserve_sms = 1;
    c->comm = comm;
    c->nranks = nranks;
    c->rank = rank;
    c->gathered = gathered;
    return DPE_OK;
}

int dpe_comm_destroy(dpe_ctx* c) {
    if (!c) { dpe::set_error("null context"); return DPE_EINVAL; }
    if (!c->comm) return DPE_OK;
    dpe::DevGuard g(c->cfg.device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    cudaDeviceSynchronize();
    g_nccl.CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
    if (c->gathered) { cudaFree(c->gathered); c->gathered = nullptr; }
    c->nranks = 1;
    c->rank = 0;
    return DPE_OK;
}

int dpe_comm_info(dpe_ctx* c, int* nranks, int* rank, int* nccl_version) {
    if (!c) { dpe::set_error("null context"); return DPE_EINVAL; }
    if (nranks) *nranks = c->comm ? c->nranks : 1;
    if (rank) *rank = c->comm ? c->rank : 0;
    if (nccl_version) {
        *nccl_version = 0;
        if (g_nccl.ok) g_nccl.GetVersion(nccl_version);
    }
    return DPE_OK;
}

}  // extern "C"
