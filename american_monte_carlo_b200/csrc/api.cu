// C ABI of libamc (see include/amc.h): contexts, device-resident path sets, the backward sweep driver and the
// small array ops kept for API parity with /root/reference/american_monte_carlo.py.
//
// One process per GPU.  The sweep is a chain of (fused decide+moments launch) -> (solve launch) pairs on one
// stream with no host synchronisation inside; multi-GPU inserts one NCCL all-reduce of <= 31 doubles between
// the partial-sum reduction and the solve of every step (NCCL is dlopen'ed lazily: single-GPU use needs no
// NCCL at all, and under torch the already-loaded libnccl.so.2 is reused).
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>      // types and enums only: the library is resolved at run time
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <string>
#include <vector>

#include "../../include/amc.h"
#include "kernels.h"

using namespace amc;

// ---------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(AMC_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

extern "C" const char* amc_last_error(void) { return g_err; }
extern "C" int amc_version(void) { return 100; }

// ---------------------------------------------------------------------------------------------------------
// NCCL, resolved lazily
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return AMC_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // torch's copy, if the process has one
    const char* env = getenv("AMC_NCCL_LIB");
    if (!h && env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(AMC_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                                          \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                                    \
    if (!g_nccl.field) return fail(AMC_ERR_NCCL, "libnccl.so.2 lacks symbol %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(GetErrorString, "ncclGetErrorString")
    SYM(CommGetAsyncError, "ncclCommGetAsyncError")
    SYM(CommAbort, "ncclCommAbort")
#undef SYM
    g_nccl.handle = h;
    return AMC_OK;
}

#define NC(call)                                                                                          \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess)                                                                            \
            return fail(AMC_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct amc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t total_mem = 0;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    // fused all-reduce over NVLink peer memory (PeerArgs in kernels.h): own mailbox + the peers' mailboxes mapped
    // through CUDA IPC.  transport: 0 = single GPU, 1 = NCCL all-reduce between two solve launches, 2 = peer memory
    int transport = 0;
    void* mailbox = nullptr;
    void* peer_mailbox[kPeerMax] = {};
    int* peer_err = nullptr;
    // grow-only scratch (one pricing call at a time per context)
    DevBuf U, tau, first_hit, partials, sums, diag, stage, misc, batch_tab, ccr, tabs, syncbuf, lstate, regx;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;      // around the persistent sweep kernel
    int sweep_grid_cache[4][AMC_MAX_K];
    // one-cluster sweep of small path sets (lsm_cluster.cuh): path capacity per storage combination and degree
    // (-1: not asked yet, 0: unavailable), and the tables last uploaded for it (same contract again -> no copy)
    int64_t cluster_cap[3][AMC_MAX_K];
    std::vector<unsigned char> tabs_host;
    uint32_t peer_seq = 0;          // sequence numbers of the peer exchange are handed out per sweep by the host: every
                                    // rank advances by the same amount per sharded sweep, also when a sweep failed
    std::vector<amc_paths*> live_paths;   // path sets not yet freed: invalidated (not leaked, not dangling) on destroy
    std::vector<cudaEvent_t> events;
    int grid_cache[3][AMC_MAX_K];
    // freed path matrices are kept for reuse (all work is ordered on one stream, so a recycled buffer is safe):
    // a pricing loop then never pays cudaMalloc/cudaFree (both synchronise the device) for multi-GB matrices
    std::vector<DevBuf> path_pool;
    // host -> device streaming of injected normals: a copy stream and two staging halves, so that the H2D copy of
    // chunk k+1 overlaps the path kernel of chunk k (and the staging buffer is 2 chunks, not the whole array)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_ready = nullptr;
    // CUDA graph of the last sweep's launch chain (replayed when the next sweep would enqueue identical launches)
    cudaGraphExec_t graph_exec = nullptr;
    std::vector<unsigned char> graph_key, graph_seen;
};
constexpr size_t kPathPoolMax = 2;

struct amc_paths {
    amc_ctx* ctx = nullptr;         // null once the context has been destroyed (the handle can still be freed)
    void* S = nullptr;              // null for a lean (path-free) set
    // lean set: terminal fixed-point log2-prices + what regenerates every earlier column (gbm_quad.cuh)
    bool lean = false, borrowed = false;
    int32_t* Ln = nullptr;
    QuadGen gen = {};
    int rounds = 10;
    int64_t path_offset = 0;
    uint64_t seed = 0;
    int64_t ld = 0, n_local = 0, n_global = 0;
    int n_steps = 0, dtype = 0;
    size_t bytes = 0;
    std::vector<double> mu, sigma;      // per-column affine maps
};

static int ensure(DevBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return AMC_OK;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes < 256 ? 256 : bytes;
    CU(cudaMalloc(&b.p, want));
    b.cap = want;
    return AMC_OK;
}

// Wait for the context stream while NCCL kernels may be in flight: a lost rank would otherwise leave the collective (and
// this host thread) waiting forever.  The stream is polled; every pass asks the communicator for asynchronous errors, and
// a wall-clock limit (AMC_NCCL_TIMEOUT_S, default 120) bounds the wait.  On either, the communicator is aborted -- which
// releases the stuck kernels -- and the call fails with AMC_ERR_NCCL; the context has no communicator afterwards.
static int sync_with_nccl_watchdog(amc_ctx* c) {
    if (!c->comm || !g_nccl.CommGetAsyncError) {
        CU(cudaStreamSynchronize(c->stream));
        return AMC_OK;
    }
    static const double limit_s = getenv("AMC_NCCL_TIMEOUT_S") ? atof(getenv("AMC_NCCL_TIMEOUT_S")) : 120.0;
    timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (unsigned spins = 0;; ++spins) {
        cudaError_t q = cudaStreamQuery(c->stream);
        if (q == cudaSuccess) return AMC_OK;
        if (q != cudaErrorNotReady) return fail(AMC_ERR_CUDA, "stream: %s", cudaGetErrorString(q));
        ncclResult_t async = ncclSuccess;
        ncclResult_t r = g_nccl.CommGetAsyncError(c->comm, &async);
        timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double waited = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        const bool failed = (r != ncclSuccess) || (async != ncclSuccess && async != ncclInProgress);
        if (failed || waited > limit_s) {
            const char* why = failed ? g_nccl.GetErrorString(r != ncclSuccess ? r : async) : "no progress within the time limit";
            g_nccl.CommAbort(c->comm);
            c->comm = nullptr;
            c->transport = 0;
            c->world = 1;
            cudaStreamSynchronize(c->stream);
            return fail(AMC_ERR_NCCL, "NCCL all-reduce of the moment sums did not complete (%s, waited %.1f s): a rank is gone; "
                                      "the communicator has been aborted", why, waited);
        }
        if (spins > 1000) {
            timespec nap = {0, 200000};           // 0.2 ms between polls once the first millisecond has passed
            nanosleep(&nap, nullptr);
        }
    }
}

static size_t elem_size(int dtype) { return dtype == AMC_F32 ? 4 : 8; }
static const void* column(const amc_paths* p, int t) {
    return (const char*)p->S + (size_t)t * (size_t)p->ld * elem_size(p->dtype);
}

extern "C" int amc_ctx_create(int device, void* stream, amc_ctx** out) {
    if (!out) return fail(AMC_ERR_VALUE, "amc_ctx_create: out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(AMC_ERR_CUDA, "amc_ctx_create: no CUDA device (%s); libamc has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(AMC_ERR_VALUE, "amc_ctx_create: device %d of %d", device, count);
    CU(cudaSetDevice(device));
    amc_ctx* c = new amc_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->total_mem = prop.totalGlobalMem;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < AMC_MAX_K; ++j) c->grid_cache[i][j] = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < AMC_MAX_K; ++j) c->sweep_grid_cache[i][j] = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < AMC_MAX_K; ++j) c->cluster_cap[i][j] = -1;
    *out = c;
    return AMC_OK;
}

extern "C" int amc_ctx_destroy(amc_ctx* c) {
    if (!c) return AMC_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int q = 0; q < kPeerMax; ++q)
        if (c->peer_mailbox[q] && q != c->rank) cudaIpcCloseMemHandle(c->peer_mailbox[q]);
    if (c->mailbox) cudaFree(c->mailbox);
    if (c->peer_err) cudaFree(c->peer_err);
    if (c->comm && g_nccl.handle) g_nccl.CommDestroy(c->comm);
    // path sets that outlive their context keep a valid handle: their device memory goes with the context
    for (amc_paths* p : c->live_paths) {
        if (p->S) cudaFree(p->S);
        if (p->Ln) cudaFree(p->Ln);
        p->S = nullptr;
        p->Ln = nullptr;
        p->ctx = nullptr;
    }
    c->live_paths.clear();
    for (cudaEvent_t ev : {c->ev_k0, c->ev_k1})
        if (ev) cudaEventDestroy(ev);
    DevBuf* bufs[] = {&c->U, &c->tau, &c->first_hit, &c->partials, &c->sums, &c->diag, &c->stage, &c->misc, &c->batch_tab, &c->ccr,
                      &c->tabs, &c->syncbuf, &c->lstate, &c->regx};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    for (DevBuf& b : c->path_pool)
        if (b.p) cudaFree(b.p);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
        if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]);
    }
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (cudaEvent_t ev : c->events) cudaEventDestroy(ev);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return AMC_OK;
}

extern "C" int amc_ctx_sync(amc_ctx* c) {
    if (!c) return fail(AMC_ERR_VALUE, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return AMC_OK;
}

extern "C" int amc_ctx_device_info(amc_ctx* c, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem) {
    if (!c) return fail(AMC_ERR_VALUE, "null context");
    if (sm_count) *sm_count = c->sm_count;
    if (cc_major) *cc_major = c->cc_major;
    if (cc_minor) *cc_minor = c->cc_minor;
    if (total_mem) *total_mem = (int64_t)c->total_mem;
    return AMC_OK;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int amc_comm_unique_id(char id[128]) {
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId uid;
    NC(g_nccl.GetUniqueId(&uid));
    memcpy(id, uid.internal, 128);
    return AMC_OK;
}

// Map every rank's mailbox into this process.  The 64-byte IPC handles travel through one NCCL all-gather; every
// rank then reports whether it could open all of them and the minimum over ranks decides (so either all ranks use
// peer memory or none does).
static int peer_setup(amc_ctx* c) {
    const int W = c->world;
    if (W > kPeerMax) return fail(AMC_ERR_NCCL, "peer-memory all-reduce supports up to %d ranks", kPeerMax);
    // Every rank takes part in BOTH agreement collectives below whatever happens locally (a rank that left early would
    // leave its peers blocked inside NCCL): local failures only clear `ok`.
    const size_t box_bytes = (size_t)kPeerRing * W * kAccStride * sizeof(uint4);
    int ok = 1;
    char why[200] = "";
    auto note = [&](const char* what, cudaError_t e) {
        if (e == cudaSuccess) return;
        if (ok) snprintf(why, sizeof(why), "%s: %s", what, cudaGetErrorString(e));
        ok = 0;
        cudaGetLastError();
    };
    note("cudaMalloc(mailbox)", cudaMalloc(&c->mailbox, box_bytes));
    note("cudaMalloc(flags)", cudaMalloc((void**)&c->peer_err, 256));
    if (ok) {
        note("memset", cudaMemsetAsync(c->mailbox, 0, box_bytes, c->stream));
        note("memset", cudaMemsetAsync(c->peer_err, 0, 256, c->stream));
        note("sync", cudaStreamSynchronize(c->stream));          // zeroed before any peer can learn the handle
    }
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok) note("cudaIpcGetMemHandle", cudaIpcGetMemHandle(&mine, c->mailbox));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    int rc = ensure(c->misc, 64 * (size_t)(W + 1) + 64);
    if (rc) return rc;                                          // nothing can be exchanged without this scratch
    char* send = (char*)c->misc.p;
    char* recv = send + 64;
    CU(cudaMemcpyAsync(send, &mine, 64, cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllGather(send, recv, 64, ncclChar, c->comm, c->stream));
    std::vector<cudaIpcMemHandle_t> all(W);
    CU(cudaMemcpyAsync(all.data(), recv, 64 * (size_t)W, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int q = 0; q < W && ok; ++q) {
        if (q == c->rank) { c->peer_mailbox[q] = c->mailbox; continue; }
        cudaError_t e = cudaIpcOpenMemHandle(&c->peer_mailbox[q], all[q], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            c->peer_mailbox[q] = nullptr;
            char what[64];
            snprintf(what, sizeof(what), "cudaIpcOpenMemHandle(rank %d)", q);
            note(what, e);
        }
    }
    // agreement: sum of ok flags must equal W
    double flag = ok ? 1.0 : 0.0;
    double* fdev = (double*)c->misc.p;
    CU(cudaMemcpyAsync(fdev, &flag, 8, cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllReduce(fdev, fdev, 1, ncclFloat64, ncclSum, c->comm, c->stream));
    CU(cudaMemcpyAsync(&flag, fdev, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if ((int)(flag + 0.5) != W)
        return fail(AMC_ERR_NCCL, "%d of %d ranks could map all mailboxes%s%s", (int)(flag + 0.5), W, why[0] ? "; " : "", why);
    c->transport = 2;
    return AMC_OK;
}

extern "C" int amc_comm_init(amc_ctx* c, int world_size, int rank, const char id[128]) {
    if (!c) return fail(AMC_ERR_VALUE, "null context");
    if (world_size < 1 || rank < 0 || rank >= world_size)
        return fail(AMC_ERR_VALUE, "amc_comm_init: rank %d of %d", rank, world_size);
    if (c->comm) return fail(AMC_ERR_STATE, "amc_comm_init: communicator already initialised");
    if (world_size == 1) { c->world = 1; c->rank = 0; return AMC_OK; }
    int rc = nccl_load();
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    NC(g_nccl.CommInitRank(&c->comm, world_size, uid, rank));
    c->world = world_size;
    c->rank = rank;
    c->transport = 1;
    // AMC_ALLREDUCE = p2p (must work) | nccl (do not try) | unset: peer memory when every rank can map every mailbox
    const char* mode = getenv("AMC_ALLREDUCE");
    if (mode && mode[0] == 'n') return AMC_OK;
    rc = peer_setup(c);
    if (rc != AMC_OK && mode && mode[0] == 'p') return rc;
    if (rc != AMC_OK) fprintf(stderr, "libamc: rank %d: peer-memory all-reduce unavailable (%s); using NCCL\n", rank, g_err);
    return AMC_OK;
}

extern "C" int amc_comm_transport(amc_ctx* c, int* transport) {
    if (!c || !transport) return fail(AMC_ERR_VALUE, "null argument");
    *transport = c->transport;
    return AMC_OK;
}

extern "C" int amc_comm_info(amc_ctx* c, int* world_size, int* rank) {
    if (!c) return fail(AMC_ERR_VALUE, "null context");
    if (world_size) *world_size = c->world;
    if (rank) *rank = c->rank;
    return AMC_OK;
}

extern "C" int amc_comm_allreduce_host(amc_ctx* c, double* buf, int n) {
    if (!c || !buf || n < 0) return fail(AMC_ERR_VALUE, "amc_comm_allreduce_host: bad argument");
    if (c->world == 1 || n == 0) return AMC_OK;
    CU(cudaSetDevice(c->device));
    int rc = ensure(c->misc, (size_t)n * 8);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->misc.p, buf, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllReduce(c->misc.p, c->misc.p, (size_t)n, ncclFloat64, ncclSum, c->comm, c->stream));
    CU(cudaMemcpyAsync(buf, c->misc.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMC_OK;
}

// ---------------------------------------------------------------------------------------------------------
// paths
enum PathKind { kPathMatrix = 0, kPathLean = 1, kPathBorrowed = 2 };

// kPathMatrix: the [n+1][ld] matrix (pooled); kPathLean: only the terminal log-price vector; kPathBorrowed: one column
// living in the context's regression scratch (regx) -- never pooled, never freed with the handle
static int paths_alloc(amc_ctx* c, int n_steps, int64_t n_local, int64_t n_global, int dtype, amc_paths** out,
                       int kind = kPathMatrix) {
    if (!c || !out) return fail(AMC_ERR_VALUE, "null argument");
    if (n_steps < 0 || n_local < 0 || n_global < n_local)
        return fail(AMC_ERR_VALUE, "bad path-set shape: n_time_steps=%d n_paths_local=%lld n_paths_global=%lld", n_steps,
                    (long long)n_local, (long long)n_global);
    if (dtype != AMC_F64 && dtype != AMC_F32) return fail(AMC_ERR_VALUE, "unknown dtype %d", dtype);
    CU(cudaSetDevice(c->device));
    amc_paths* p = new amc_paths();
    p->ctx = c;
    p->n_steps = n_steps;
    p->n_local = n_local;
    p->n_global = n_global;
    p->dtype = dtype;
    p->ld = padded_len(n_local > 0 ? n_local : 1);
    p->mu.assign(n_steps + 1, 0.0);
    p->sigma.assign(n_steps + 1, 1.0);
    if (kind == kPathLean) {
        p->lean = true;
        p->bytes = (size_t)p->ld * 4;
        for (size_t i = 0; i < c->path_pool.size(); ++i) {
            if (c->path_pool[i].cap == p->bytes) {
                p->Ln = (int32_t*)c->path_pool[i].p;
                c->path_pool.erase(c->path_pool.begin() + i);
                break;
            }
        }
        if (!p->Ln) {
            cudaError_t e = cudaMalloc((void**)&p->Ln, p->bytes);
            if (e != cudaSuccess) {
                delete p;
                return fail(AMC_ERR_CUDA, "cudaMalloc of the log-price vector failed: %s", cudaGetErrorString(e));
            }
        }
        c->live_paths.push_back(p);
        *out = p;
        return AMC_OK;
    }
    if (kind == kPathBorrowed) {
        p->borrowed = true;
        p->bytes = (size_t)p->ld * (size_t)(n_steps + 1) * elem_size(dtype);
        int rc = ensure(c->regx, p->bytes);
        if (rc) { delete p; return rc; }
        p->S = c->regx.p;
        *out = p;
        return AMC_OK;
    }
    p->bytes = (size_t)p->ld * (size_t)(n_steps + 1) * elem_size(dtype);
    for (size_t i = 0; i < c->path_pool.size(); ++i) {
        if (c->path_pool[i].cap == p->bytes) {
            p->S = c->path_pool[i].p;
            c->path_pool.erase(c->path_pool.begin() + i);
            break;
        }
    }
    if (!p->S) {
        cudaError_t e = cudaMalloc(&p->S, p->bytes);
        if (e != cudaSuccess && !c->path_pool.empty()) {          // give pooled memory back and retry once
            for (DevBuf& b : c->path_pool) cudaFree(b.p);
            c->path_pool.clear();
            cudaGetLastError();
            e = cudaMalloc(&p->S, p->bytes);
        }
        if (e != cudaSuccess) {
            const size_t bytes = p->bytes;
            delete p;
            return fail(AMC_ERR_CUDA, "cudaMalloc of %zu bytes for the path matrix failed: %s", bytes, cudaGetErrorString(e));
        }
    }
    c->live_paths.push_back(p);
    *out = p;
    return AMC_OK;
}

// GBM moments: E S_t = S0 e^{rt}, Var S_t = S0^2 e^{2rt} (e^{sigma^2 t} - 1).  Any centre/scale is valid for the
// internal standardisation; the analytic ones are deterministic and independent of the number of GPUs.
static void analytic_maps(amc_paths* p, double S0, double r, double sigma, double T) {
    const int n = p->n_steps;
    for (int t = 0; t <= n; ++t) {
        const double tt = n > 0 ? T * (double)t / (double)n : 0.0;
        const double m = S0 * exp(r * tt);
        const double v = expm1(sigma * sigma * tt);
        double sd = m * sqrt(v > 0.0 ? v : 0.0);
        p->mu[t] = m;
        p->sigma[t] = (sd > 1e-12 * fabs(m) && sd > 0.0 && isfinite(sd)) ? sd : 1.0;
    }
}

static GbmParams gbm_params(double S0, double r, double sigma, double T, int n) {
    GbmParams g;
    const double dt = T / (double)n;                       // amc.py:73
    g.S0 = S0;
    g.drift = (r - 0.5 * sigma * sigma) * dt;              // amc.py:75
    g.vol = sigma * sqrt(dt);
    return g;
}

extern "C" int amc_paths_generate(amc_ctx* c, double S0, double r, double sigma, double T, int n_time_steps,
                                  int64_t n_paths_local, int64_t path_offset, int64_t n_paths_global, int dtype,
                                  uint64_t seed, amc_paths** out) {
    if (n_time_steps < 1) return fail(AMC_ERR_VALUE, "n_time_steps must be >= 1");
    amc_paths* p = nullptr;
    int rc = paths_alloc(c, n_time_steps, n_paths_local, n_paths_global, dtype, &p);
    if (rc) return rc;
    if (n_paths_local > 0) {
        cudaError_t e = launch_generate_philox(dtype, p->S, nullptr, p->ld, n_time_steps, n_paths_local, path_offset,
                                               gbm_params(S0, r, sigma, T, n_time_steps), seed, c->sm_count, c->stream);
        if (e != cudaSuccess) {
            amc_paths_free(p);
            return fail(AMC_ERR_CUDA, "philox path kernel launch: %s", cudaGetErrorString(e));
        }
    }
    analytic_maps(p, S0, r, sigma, T);
    *out = p;
    return AMC_OK;
}

// Path-free ("lean") set: nothing but the terminal log2-prices is stored (4 bytes per path instead of 4 (n+1)); the
// backward sweep regenerates every earlier column from the Philox counters (lsm_sweep.cuh, LEAN).  Float generator only.
extern "C" int amc_paths_generate_lean(amc_ctx* c, double S0, double r, double sigma, double T, int n_time_steps,
                                       int64_t n_paths_local, int64_t path_offset, int64_t n_paths_global, uint64_t seed,
                                       amc_paths** out) {
    if (n_time_steps < 1) return fail(AMC_ERR_VALUE, "n_time_steps must be >= 1");
    if (path_offset & 3)
        return fail(AMC_ERR_VALUE, "amc_paths_generate_lean: path_offset %lld is not a multiple of 4 (one Philox call "
                                   "serves four adjacent paths; shard in units of 4 paths)", (long long)path_offset);
    amc_paths* p = nullptr;
    int rc = paths_alloc(c, n_time_steps, n_paths_local, n_paths_global, AMC_F32, &p, kPathLean);
    if (rc) return rc;
    const GbmParams g = gbm_params(S0, r, sigma, T, n_time_steps);
    p->gen = make_quad_gen(g, n_time_steps, seed);
    p->rounds = philox_rounds();
    p->path_offset = path_offset;
    p->seed = seed;
    if (n_paths_local > 0) {
        cudaError_t e = launch_generate_philox(AMC_F32, nullptr, p->Ln, p->ld, n_time_steps, n_paths_local, path_offset, g, seed,
                                               c->sm_count, c->stream);
        if (e != cudaSuccess) {
            amc_paths_free(p);
            return fail(AMC_ERR_CUDA, "philox path kernel launch: %s", cudaGetErrorString(e));
        }
    }
    analytic_maps(p, S0, r, sigma, T);
    *out = p;
    return AMC_OK;
}

// Host normals -> paths, pipelined: chunks of whole paths go through two staging halves; the copy stream moves chunk
// k+1 over PCIe while the path kernel turns chunk k into prices on the context stream.
// `row_doubles` doubles per path in the host array; `consume(staged_chunk, first_path, n_paths)` launches the kernel that
// turns a staged chunk into columns.
template <typename Consume>
static int stream_rows_from_host(amc_ctx* c, amc_paths* p, const double* Z, int row_doubles, Consume consume) {
    const int n = row_doubles;
    const int64_t P = p->n_local;
    if (!c->copy_stream) {
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    }
    const size_t row_bytes = (size_t)n * 8;
    const char* mb_env = getenv("AMC_STAGE_CHUNK_MB");                       // tests shrink it to force many chunks
    const size_t mb = (mb_env && atoi(mb_env) > 0) ? (size_t)atoi(mb_env) : 256;
    int64_t chunk = (int64_t)((mb << 20) / row_bytes) / 128 * 128;           // a multiple of the kernels' 128-path blocks
    if (chunk < 128) chunk = 128;
    if (chunk > P) chunk = P;
    const size_t chunk_bytes = ((size_t)chunk * row_bytes + 255) / 256 * 256;
    int rc = ensure(c->stage, 2 * chunk_bytes);
    if (rc) return rc;
    // the staging halves may still be read by kernels queued earlier on the context stream
    CU(cudaEventRecord(c->ev_ready, c->stream));
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_ready, 0));
    const size_t es = elem_size(p->dtype);
    int k = 0;
    for (int64_t p0 = 0; p0 < P; p0 += chunk, ++k) {
        const int buf = k & 1;
        const int64_t np = (P - p0 < chunk) ? (P - p0) : chunk;
        double* st = (double*)((char*)c->stage.p + (size_t)buf * chunk_bytes);
        if (k >= 2) CU(cudaStreamWaitEvent(c->copy_stream, c->ev_free[buf], 0));
        CU(cudaMemcpyAsync(st, Z + (size_t)p0 * n, (size_t)np * row_bytes, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(c->ev_copied[buf], c->copy_stream));
        CU(cudaStreamWaitEvent(c->stream, c->ev_copied[buf], 0));
        CU(consume(st, (char*)p->S + (size_t)p0 * es, np));
        CU(cudaEventRecord(c->ev_free[buf], c->stream));
    }
    return AMC_OK;
}

static int stream_normals_from_host(amc_ctx* c, amc_paths* p, const double* Z, GbmParams g) {
    return stream_rows_from_host(c, p, Z, p->n_steps, [&](const double* st, void* S0p, int64_t np) {
        return launch_from_normals(p->dtype, st, S0p, p->ld, p->n_steps, np, g, c->stream);
    });
}

static int from_normals_impl(amc_ctx* c, const double* Z, bool z_on_device, double S0, double r, double sigma, double T,
                             int n_time_steps, int64_t n_local, int64_t n_global, int dtype, amc_paths** out) {
    if (n_time_steps < 1) return fail(AMC_ERR_VALUE, "n_time_steps must be >= 1");
    if (!Z && n_local > 0) return fail(AMC_ERR_VALUE, "Z is null");
    amc_paths* p = nullptr;
    int rc = paths_alloc(c, n_time_steps, n_local, n_global, dtype, &p);
    if (rc) return rc;
    if (n_local > 0 && z_on_device) {
        cudaError_t e = launch_from_normals(dtype, Z, p->S, p->ld, n_time_steps, n_local,
                                            gbm_params(S0, r, sigma, T, n_time_steps), c->stream);
        if (e != cudaSuccess) { amc_paths_free(p); return fail(AMC_ERR_CUDA, "normals path kernel launch: %s", cudaGetErrorString(e)); }
    } else if (n_local > 0) {
        rc = stream_normals_from_host(c, p, Z, gbm_params(S0, r, sigma, T, n_time_steps));
        if (rc) { amc_paths_free(p); return rc; }
    }
    analytic_maps(p, S0, r, sigma, T);
    *out = p;
    return AMC_OK;
}

extern "C" int amc_paths_from_normals(amc_ctx* c, const double* Z, double S0, double r, double sigma, double T,
                                      int n_time_steps, int64_t n_paths_local, int64_t n_paths_global, int dtype,
                                      amc_paths** out) {
    return from_normals_impl(c, Z, false, S0, r, sigma, T, n_time_steps, n_paths_local, n_paths_global, dtype, out);
}

extern "C" int amc_paths_from_normals_dev(amc_ctx* c, const double* Z_dev, double S0, double r, double sigma, double T,
                                          int n_time_steps, int64_t n_paths_local, int64_t n_paths_global, int dtype,
                                          amc_paths** out) {
    return from_normals_impl(c, Z_dev, true, S0, r, sigma, T, n_time_steps, n_paths_local, n_paths_global, dtype, out);
}

// measured maps for adopted matrices: per-column (count, mean, M2), merged over chunks and ranks in fixed order
static int measured_maps(amc_ctx* c, amc_paths* p, bool across_ranks = true) {
    const int ncol = p->n_steps + 1;
    const int n_chunks = 64;
    std::vector<double> cnt(ncol, 0.0), mean(ncol, 0.0), m2(ncol, 0.0);
    if (p->n_local > 0) {
        int rc = ensure(c->misc, (size_t)ncol * (n_chunks * 2 + 1) * 8);
        if (rc) return rc;
        double* partial = (double*)c->misc.p;
        double* shift = partial + (size_t)ncol * n_chunks * 2;
        CU(launch_column_stats(p->dtype, p->S, p->ld, ncol, p->n_local, n_chunks, partial, shift, c->stream));
        std::vector<double> h((size_t)ncol * (n_chunks * 2 + 1));
        CU(cudaMemcpyAsync(h.data(), partial, h.size() * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (int t = 0; t < ncol; ++t) {
            double s1 = 0.0, s2 = 0.0;
            for (int k = 0; k < n_chunks; ++k) {
                s1 += h[((size_t)t * n_chunks + k) * 2];
                s2 += h[((size_t)t * n_chunks + k) * 2 + 1];
            }
            const double n = (double)p->n_local;
            const double sh = h[(size_t)ncol * n_chunks * 2 + t];
            const double m1 = s1 / n;
            double var = s2 / n - m1 * m1;
            if (var < 0.0) var = 0.0;
            cnt[t] = n;
            mean[t] = sh + m1;
            m2[t] = var * n;
        }
    }
    if (c->world > 1 && across_ranks) {
        // all-gather (count, mean, M2) triples and merge in rank order (Chan et al. pairwise update)
        const size_t per = (size_t)ncol * 3;
        int rc = ensure(c->misc, per * 8 * (size_t)(c->world + 1));
        if (rc) return rc;
        std::vector<double> mine(per);
        for (int t = 0; t < ncol; ++t) { mine[3 * t] = cnt[t]; mine[3 * t + 1] = mean[t]; mine[3 * t + 2] = m2[t]; }
        double* send = (double*)c->misc.p;
        double* recv = send + per;
        CU(cudaMemcpyAsync(send, mine.data(), per * 8, cudaMemcpyHostToDevice, c->stream));
        NC(g_nccl.AllGather(send, recv, per, ncclFloat64, c->comm, c->stream));
        std::vector<double> all(per * c->world);
        CU(cudaMemcpyAsync(all.data(), recv, all.size() * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (int t = 0; t < ncol; ++t) {
            double n = 0.0, m = 0.0, q = 0.0;
            for (int rk = 0; rk < c->world; ++rk) {
                const double nb = all[rk * per + 3 * t], mb = all[rk * per + 3 * t + 1], qb = all[rk * per + 3 * t + 2];
                if (nb <= 0.0) continue;
                const double tot = n + nb, dlt = mb - m;
                q = q + qb + dlt * dlt * n * nb / tot;
                m = m + dlt * nb / tot;
                n = tot;
            }
            cnt[t] = n; mean[t] = m; m2[t] = q;
        }
    }
    for (int t = 0; t < ncol; ++t) {
        const double sd = cnt[t] > 0.0 ? sqrt(m2[t] / cnt[t]) : 0.0;
        p->mu[t] = mean[t];
        p->sigma[t] = (sd > 1e-12 * fabs(mean[t]) && sd > 0.0 && isfinite(sd)) ? sd : 1.0;
    }
    return AMC_OK;
}

extern "C" int amc_paths_from_host(amc_ctx* c, const double* S, int n_time_steps, int64_t n_paths_local,
                                   int64_t n_paths_global, int dtype, amc_paths** out) {
    if (!S && n_paths_local > 0) return fail(AMC_ERR_VALUE, "S is null");
    amc_paths* p = nullptr;
    int rc = paths_alloc(c, n_time_steps, n_paths_local, n_paths_global, dtype, &p);
    if (rc) return rc;
    if (n_paths_local > 0) {
        // same pipeline as the injected normals: chunks of whole rows, copy stream || transpose kernel
        rc = stream_rows_from_host(c, p, S, n_time_steps + 1, [&](const double* st, void* S0p, int64_t np) {
            return launch_transpose_in(dtype, st, S0p, p->ld, n_time_steps + 1, np, c->stream);
        });
        if (rc) { amc_paths_free(p); return rc; }
    }
    // column statistics are merged over the ranks only where the path axis is sharded over them: a rank-local matrix
    // (n_global == n_local) involves no collective at all
    rc = measured_maps(c, p, n_paths_global != n_paths_local);
    if (rc) { amc_paths_free(p); return rc; }
    *out = p;
    return AMC_OK;
}

extern "C" int amc_paths_free(amc_paths* p) {
    if (!p) return AMC_OK;
    amc_ctx* c = p->ctx;
    if (!c || p->borrowed) {            // the context is gone (its destroy released the device memory), or a scratch view
        delete p;
        return AMC_OK;
    }
    cudaSetDevice(c->device);
    for (size_t i = 0; i < c->live_paths.size(); ++i)
        if (c->live_paths[i] == p) { c->live_paths.erase(c->live_paths.begin() + i); break; }
    void* mem = p->S ? p->S : (void*)p->Ln;
    p->Ln = nullptr;
    p->S = mem;
    if (p->S) {
        if (c->path_pool.size() < kPathPoolMax) {
            DevBuf b;
            b.p = p->S;
            b.cap = p->bytes;
            c->path_pool.push_back(b);
        } else {
            cudaStreamSynchronize(c->stream);
            cudaFree(p->S);
        }
    }
    delete p;
    return AMC_OK;
}

extern "C" int amc_paths_info(const amc_paths* p, int64_t* n_paths_local, int64_t* n_paths_global, int* n_time_steps,
                              int* dtype, int64_t* bytes_on_device) {
    if (!p) return fail(AMC_ERR_VALUE, "null paths");
    if (n_paths_local) *n_paths_local = p->n_local;
    if (n_paths_global) *n_paths_global = p->n_global;
    if (n_time_steps) *n_time_steps = p->n_steps;
    if (dtype) *dtype = p->dtype;
    if (bytes_on_device) *bytes_on_device = (int64_t)p->bytes;
    return AMC_OK;
}

extern "C" int amc_paths_column(const amc_paths* p, int t, double* out) {
    if (!p || !out) return fail(AMC_ERR_VALUE, "null argument");
    if (t < 0 || t > p->n_steps) return fail(AMC_ERR_VALUE, "column %d out of range 0..%d", t, p->n_steps);
    if (p->n_local == 0) return AMC_OK;
    amc_ctx* c = p->ctx;
    if (!c) return fail(AMC_ERR_STATE, "the path set's context has been destroyed");
    CU(cudaSetDevice(c->device));
    if (p->lean) {
        int rc = ensure(c->misc, (size_t)p->n_local * 8);
        if (rc) return rc;
        CU(launch_lean_walk(0, p->rounds, p->gen, p->path_offset >> 2, p->n_local, p->n_steps, t, 0, p->n_local, 0.0, nullptr,
                            (double*)c->misc.p, nullptr, c->sm_count, c->stream));
        CU(cudaMemcpyAsync(out, c->misc.p, (size_t)p->n_local * 8, cudaMemcpyDeviceToHost, c->stream));
    } else if (p->dtype == AMC_F64) {
        CU(cudaMemcpyAsync(out, column(p, t), (size_t)p->n_local * 8, cudaMemcpyDeviceToHost, c->stream));
    } else {
        int rc = ensure(c->misc, (size_t)p->n_local * 8);
        if (rc) return rc;
        CU(launch_column_to_f64(p->dtype, column(p, t), p->n_local, (double*)c->misc.p, c->stream));
        CU(cudaMemcpyAsync(out, c->misc.p, (size_t)p->n_local * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return AMC_OK;
}

extern "C" int amc_paths_rows(const amc_paths* p, int64_t p0, int64_t p1, double* out) {
    if (!p || !out) return fail(AMC_ERR_VALUE, "null argument");
    if (p0 < 0 || p1 < p0 || p1 > p->n_local)
        return fail(AMC_ERR_VALUE, "rows [%lld, %lld) out of range 0..%lld", (long long)p0, (long long)p1, (long long)p->n_local);
    if (p1 == p0) return AMC_OK;
    amc_ctx* c = p->ctx;
    if (!c) return fail(AMC_ERR_STATE, "the path set's context has been destroyed");
    CU(cudaSetDevice(c->device));
    const size_t bytes = (size_t)(p1 - p0) * (size_t)(p->n_steps + 1) * 8;
    int rc = ensure(c->misc, bytes);
    if (rc) return rc;
    if (p->lean)
        CU(launch_lean_walk(1, p->rounds, p->gen, p->path_offset >> 2, p->n_local, p->n_steps, p->n_steps, p0, p1, 0.0, nullptr,
                            (double*)c->misc.p, nullptr, c->sm_count, c->stream));
    else
        CU(launch_gather_rows(p->dtype, p->S, p->ld, p->n_steps + 1, p0, p1, (double*)c->misc.p, c->stream));
    CU(cudaMemcpyAsync(out, c->misc.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMC_OK;
}

extern "C" int amc_paths_column_maps(const amc_paths* p, double* mu, double* sigma) {
    if (!p) return fail(AMC_ERR_VALUE, "null paths");
    for (int t = 0; t <= p->n_steps; ++t) {
        if (mu) mu[t] = p->mu[t];
        if (sigma) sigma[t] = p->sigma[t];
    }
    return AMC_OK;
}

// ---------------------------------------------------------------------------------------------------------
// backward sweep
static int check_spec(const amc_lsm_spec* s) {
    if (!s) return fail(AMC_ERR_VALUE, "null spec");
    if (s->basis < 0 || s->basis > AMC_BASIS_LAGUERRE)
        return fail(AMC_ERR_VALUE, "Unknown basis type id %d. Use 'Power', 'Chebyshev', or 'Legendre'.", s->basis);
    if (s->degree < 0 || s->degree > AMC_MAX_DEGREE)
        return fail(AMC_ERR_VALUE, "degree %d outside 0..%d", s->degree, AMC_MAX_DEGREE);
    return AMC_OK;
}

static int step_grid(amc_ctx* c, int dtype, int state_f32, int degree, int64_t n_paths) {
    int& g = c->grid_cache[state_f32 ? 2 : dtype][degree];
    if (g == 0) g = step_grid_size(dtype, state_f32, degree, c->sm_count);
    // small path sets: no more blocks than there are tiles of 1024 paths to hand out
    int64_t need = (n_paths + 1023) / 1024;
    if (need < 1) need = 1;
    return (int)(need < g ? need : g);
}

struct EventPool {
    amc_ctx* c;
    size_t used = 0;
    int get(cudaEvent_t* ev) {
        if (used == c->events.size()) {
            cudaEvent_t e;
            CU(cudaEventCreate(&e));
            c->events.push_back(e);
        }
        *ev = c->events[used++];
        return AMC_OK;
    }
};

// One backward sweep for `C` contracts on one path set.  C == 1: the single-contract entry point with all its
// diagnostics.  C > 1 (amc_lsm_price_batch): the contracts share r, dt, basis, degree, scaling and barrier and differ
// in strike / payoff side / exercise style; every launch carries all of them (grid.y = contract), so the columns are
// read from HBM once per step for the whole batch and the per-step launch chain is paid once.
static int lsm_sweep(amc_ctx* c, const amc_paths* p, const amc_lsm_spec* specs, int C, double* price,
                     amc_lsm_steps* steps, int32_t* exercise_step_out, double* cashflow0_out, double* gamma_batch_out,
                     amc_lsm_timing* timing, int profile, const int32_t* first_hit_host = nullptr) {
    const amc_lsm_spec* spec = specs;
    int rc;
    CU(cudaSetDevice(c->device));

    const int n = p->n_steps, D = spec->degree, dtype = p->dtype;
    const int64_t P = p->n_local;
    const double Pg = (double)p->n_global;
    bool american = false;
    for (int i = 0; i < C; ++i) american = american || specs[i].is_american != 0;
    const bool regress = (american || spec->want_regression) && n >= 1;
    const bool barrier = !isnan(spec->barrier) || first_hit_host != nullptr;
    const int sf32 = spec->state_f32 ? 1 : 0;
    const size_t bU = sf32 ? 4 : 8;
    int grid = step_grid(c, dtype, sf32, D, P);
    if (C > 1) {                                      // all contracts' blocks co-resident: split the grid among them
        grid = grid / C;                              // floor: one block too many per contract would add a whole wave
        if (grid < 1) grid = 1;
    }
    const double rdt = spec->r * spec->dt;
    // the per-step exchange exists only where the path axis is really sharded over the ranks: a rank-local path set
    // (n_global == n_local, e.g. an ndarray adopted under a multi-rank context) is priced by this rank alone, and
    // batches are sharded by contract -- neither has a data-path collective
    const bool exchange = c->world > 1 && C == 1 && p->n_global != p->n_local;

    // scratch
    const int64_t ldp = padded_len(P > 0 ? P : 1);
    if ((rc = ensure(c->U, (size_t)ldp * bU * C))) return rc;
    if (spec->want_exercise_steps && (rc = ensure(c->tau, (size_t)ldp * 4))) return rc;
    if (barrier && (rc = ensure(c->first_hit, (size_t)ldp * 4))) return rc;
    if ((rc = ensure(c->partials, (size_t)grid * kAccStride * 8 * C))) return rc;
    if ((rc = ensure(c->sums, kAccStride * 8 * (size_t)C))) return rc;
    // diagnostics block: gamma (one [n+1][kMaxK] table per contract) | beta | sv | mean_std | price | rank
    const size_t nrow = (size_t)(n + 1);
    const size_t off_gamma = 0, off_beta = (size_t)C * nrow * kMaxK, off_sv = off_beta + nrow * kMaxK;
    const size_t off_ms = off_sv + nrow * kMaxK;
    const size_t off_price = off_ms + 3 * nrow, off_rank = off_price + (size_t)C + 1;   // rank: int32 after the doubles
    const size_t diag_bytes = off_rank * 8 + nrow * 4;
    if ((rc = ensure(c->diag, diag_bytes))) return rc;
    double* dg = (double*)c->diag.p;
    int* drank = (int*)(dg + off_rank);
    CU(cudaMemsetAsync(c->diag.p, 0, diag_bytes, c->stream));
    CU(cudaMemsetAsync(c->sums.p, 0, kAccStride * 8 * (size_t)C, c->stream));
    BatchContract* tab = nullptr;
    std::vector<BatchContract> tab_h;
    if (C > 1) {
        if ((rc = ensure(c->batch_tab, sizeof(BatchContract) * (size_t)C))) return rc;
        tab = (BatchContract*)c->batch_tab.p;
        tab_h.resize(C);
        for (int i = 0; i < C; ++i) {
            tab_h[i].K = specs[i].K;
            tab_h[i].is_put = specs[i].is_put;
            tab_h[i].is_american = specs[i].is_american;
        }
        CU(cudaMemcpyAsync(tab, tab_h.data(), sizeof(BatchContract) * (size_t)C, cudaMemcpyHostToDevice, c->stream));
    }

    int32_t* tau = spec->want_exercise_steps ? (int32_t*)c->tau.p : nullptr;
    int32_t* fh = barrier ? (int32_t*)c->first_hit.p : nullptr;
    if (first_hit_host && P > 0)
        CU(cudaMemcpyAsync(fh, first_hit_host, (size_t)P * 4, cudaMemcpyHostToDevice, c->stream));
    else if (barrier && P > 0 && p->lean)
        CU(launch_lean_walk(2, p->rounds, p->gen, p->path_offset >> 2, P, n, n, 0, P, spec->barrier, nullptr, nullptr, fh,
                            c->sm_count, c->stream));
    else if (barrier && P > 0)
        CU(launch_first_hit(dtype, p->S, p->ld, n + 1, P, spec->barrier, fh, c->stream));

    // L2 management knob (on by default; AMC_L2_REVERSE=0 for A/B measurements)
    static const int opt_reverse = getenv("AMC_L2_REVERSE") ? atoi(getenv("AMC_L2_REVERSE")) : 1;
    // programmatic dependent launch along the K3 -> K4 -> K3 chain (single GPU, not while profiling: the
    // per-launch events and the NCCL kernels are ordinary stream dependencies)
    static const int opt_pdl = getenv("AMC_PDL") ? atoi(getenv("AMC_PDL")) : 1;
    const bool pdl = opt_pdl && !profile && (!exchange || c->transport == 2);
    EventPool pool{c};
    cudaEvent_t ev_start, ev_stop;
    if ((rc = pool.get(&ev_start)) || (rc = pool.get(&ev_stop))) return rc;
    std::vector<cudaEvent_t> step_ev, solve_ev;
    int n_step = 0, n_solve = 0, n_other = barrier ? 1 : 0;
    auto bracket = [&](std::vector<cudaEvent_t>& v) -> int {
        if (!profile) return AMC_OK;
        cudaEvent_t e;
        int r2 = pool.get(&e);
        if (r2) return r2;
        v.push_back(e);
        CU(cudaEventRecord(e, c->stream));
        return AMC_OK;
    };

    CU(cudaEventRecord(ev_start, c->stream));

    SolveSpec sspec;
    sspec.degree = D;
    sspec.basis = spec->basis;
    sspec.scaling = spec->scaling;
    sspec.want_svd = spec->want_svd;
    sspec.scaling_factor = spec->scaling_factor;
    sspec.n_paths = Pg;
    sspec.inv_n_paths = Pg > 0.0 ? 1.0 / Pg : 0.0;
    static const int opt_warp_solve = getenv("AMC_WARP_SOLVE") ? atoi(getenv("AMC_WARP_SOLVE")) : -1;
    sspec.warp_solve = opt_warp_solve < 0 ? (D >= 6) : opt_warp_solve;    // scalar registers win up to k = 6

    // Launch-bound sweeps (small path sets: two launches per ~7 us step) are replayed as a CUDA graph: the launches of
    // a sweep are first written into a plan; if the plan is byte-identical to the previous sweep's (same buffers, same
    // contract, same path-set geometry) the instantiated graph of that sweep is launched instead of 2(n+1) kernels.
    // The first occurrence of a plan runs eagerly, the second is captured, later ones are replayed.
    struct PlanItem { int kind, pdl; StepArgs a; SolveArgs s; };
    static const int opt_graph = getenv("AMC_GRAPH") ? atoi(getenv("AMC_GRAPH")) : 1;
    const bool planned = opt_graph && !profile && (!exchange || c->transport == 2);
    std::vector<PlanItem> plan;
    auto emit_step = [&](const StepArgs& a, bool pdl_flag) -> int {
        if (planned) {
            plan.emplace_back();
            PlanItem& it = plan.back();
            memset(&it, 0, sizeof(it));
            it.kind = 0; it.pdl = pdl_flag; it.a = a;
            return AMC_OK;
        }
        CU(launch_step(dtype, sf32, D, grid, a, c->stream, pdl_flag, C));
        return AMC_OK;
    };
    auto emit_solve = [&](const SolveArgs& s, bool pdl_flag) -> int {
        if (planned) {
            plan.emplace_back();
            PlanItem& it = plan.back();
            memset(&it, 0, sizeof(it));
            it.kind = 1; it.pdl = pdl_flag; it.s = s;
            return AMC_OK;
        }
        CU(launch_solve(s, c->stream, pdl_flag));
        return AMC_OK;
    };

    auto run_step = [&](int t, int mode, bool moments) -> int {
        StepArgs a;
        memset(&a, 0, sizeof(a));
        a.x_dec = (mode != kObserve) ? column(p, t) : nullptr;
        a.x_reg = moments ? column(p, t - 1) : nullptr;
        a.U = c->U.p;
        a.tau = tau;
        a.first_hit = fh;
        a.coef = dg + off_gamma + (size_t)t * kMaxK;
        a.partials = (double*)c->partials.p;
        a.n_paths = P;
        a.t_dec = t;
        a.mode = mode;
        a.moments = moments ? 1 : 0;
        a.is_put = spec->is_put;
        a.reverse = opt_reverse ? (t & 1) : 0;
        a.K = spec->K;
        a.disc_dec = exp(-rdt * (double)t);
        a.mu_dec = p->mu[t];
        a.isg_dec = 1.0 / p->sigma[t];
        a.mu_reg = moments ? p->mu[t - 1] : 0.0;
        a.isg_reg = moments ? 1.0 / p->sigma[t - 1] : 1.0;
        a.batch = tab;
        a.u_stride = ldp;
        a.coef_stride = (int64_t)nrow * kMaxK;
        int r2;
        if ((r2 = bracket(step_ev))) return r2;
        if ((r2 = emit_step(a, pdl && n_step > 0))) return r2;
        if ((r2 = bracket(step_ev))) return r2;
        ++n_step;
        return AMC_OK;
    };

    auto run_solve = [&](int t_reg, bool final_price) -> int {
        SolveArgs s;
        memset(&s, 0, sizeof(s));
        s.partials = (const double*)c->partials.p;
        s.n_rows = grid;
        s.sums = (double*)c->sums.p;
        s.spec = sspec;
        s.final_price = 0;
        s.y_scale = final_price ? 1.0 : exp(rdt * (double)t_reg);
        s.mu_ref = final_price ? 0.0 : p->mu[t_reg];
        s.sigma_ref = final_price ? 1.0 : 1.0 / (1.0 / p->sigma[t_reg]);   // the scale the kernels effectively used
        const size_t row = final_price ? 0 : (size_t)t_reg;
        s.gamma = dg + off_gamma + row * kMaxK;
        s.beta = dg + off_beta + row * kMaxK;
        s.sv = dg + off_sv + row * kMaxK;
        s.mean_std = dg + off_ms + row * 3;
        s.rank = drank + row;
        s.price = dg + off_price;
        s.n_batch = C;
        s.gamma_stride = (int64_t)nrow * kMaxK;
        int r2;
        if ((r2 = bracket(solve_ev))) return r2;
        if (!exchange || c->transport == 2) {
            s.do_reduce = 1;
            s.do_solve = final_price ? 0 : 1;
            s.final_price = final_price ? 1 : 0;
            if (exchange) {
                for (int q = 0; q < c->world; ++q) s.peer.mailbox[q] = (uint4*)c->peer_mailbox[q];
                s.peer.world = c->world;
                s.peer.rank = c->rank;
                s.peer.seq_ctr = (uint32_t*)(c->peer_err + 16);        // same 256-byte block as the error flag
                s.peer.err = c->peer_err;
            }
            if ((r2 = emit_solve(s, pdl))) return r2;
            ++n_solve;
        } else {
            s.do_reduce = 1; s.do_solve = 0;
            CU(launch_solve(s, c->stream));
            NC(g_nccl.AllReduce(c->sums.p, c->sums.p, (size_t)(3 * D + 1), ncclFloat64, ncclSum, c->comm, c->stream));
            s.do_reduce = 0;
            s.do_solve = final_price ? 0 : 1;
            s.final_price = final_price ? 1 : 0;
            CU(launch_solve(s, c->stream));
            n_solve += 2;
            ++n_other;
        }
        if ((r2 = bracket(solve_ev))) return r2;
        return AMC_OK;
    };

    // The persistent sweep (lsm_sweep.cuh) -- ONE cooperative launch whose blocks stay resident for all passes; the block
    // that finishes a pass last reduces, exchanges and solves -- is what prices path-free sets (it regenerates the columns
    // from the counters).  For stored sets the launch chain below is the default: measured on B200 the chain is faster
    // on every workload (profiles/r2_persistent_vs_chain.md: the streaming loop needs the whole 64-register budget of 4
    // blocks per SM; sweep-level state and the in-kernel solve spill).  AMC_PERSISTENT=1 selects it for stored sets too.
    static const int opt_persistent = getenv("AMC_PERSISTENT") ? atoi(getenv("AMC_PERSISTENT")) : 0;
    // Small stored sets (everything fits the shared memory of one 16-CTA cluster) are priced by the one-cluster kernel
    // (lsm_cluster.cuh): ~4 us per step instead of the chain's 6-8.  AMC_CLUSTER=0 keeps them on the chain.
    static const int opt_cluster = getenv("AMC_CLUSTER") ? atoi(getenv("AMC_CLUSTER")) : 1;
    bool cluster = false;
    if (C == 1 && !exchange && !p->lean && !opt_persistent && opt_cluster && P >= 1) {
        int64_t& cap = c->cluster_cap[sf32 ? 2 : dtype][D];
        if (cap < 0) cap = cluster_sweep_capacity(dtype, sf32, D);
        // what the cluster's shared memory holds is more than what it prices faster than the chain: the 16 SMs of the
        // cluster stream a pass at FP64-pipe speed, the chain's 148 do not care (profiles/r2_cluster_vs_chain.md)
        static const long long opt_cluster_max = getenv("AMC_CLUSTER_MAX_PATHS") ? atoll(getenv("AMC_CLUSTER_MAX_PATHS")) : 147456;
        cluster = P <= cap && P <= opt_cluster_max;
    }
    const bool persistent = C == 1 && (opt_persistent || p->lean || cluster) && (!exchange || c->transport == 2);
    if (p->lean && !persistent)
        return fail(AMC_ERR_STATE, "path-free sets need the persistent sweep (peer-memory transport when sharded)");
    bool used_persistent = false, used_cluster = false, chain_fallback = false;
    if (exchange && c->transport == 2 && !persistent) {
        // the chain's solve launches draw their exchange sequence numbers from a device counter: seed it with this rank's
        // host-side count of exchanges so far -- the same on every rank however an earlier sweep ended
        if (c->peer_seq > 0xFFF00000u) c->peer_seq = 0;
        const uint32_t seed = c->peer_seq;                  // pageable source: staged before the call returns
        CU(cudaMemcpyAsync(c->peer_err + 16, &seed, 4, cudaMemcpyHostToDevice, c->stream));
        c->peer_seq += (uint32_t)(regress ? n + 1 : 1);
    }
    if (persistent) {
        used_persistent = !cluster;
        used_cluster = cluster;
        const int n_passes = regress ? n + 1 : 1;
        const int lean = p->lean ? 1 : 0;
        int wgrid = 1;
        if (!cluster) {
            int& gcache = c->sweep_grid_cache[lean ? 3 : (sf32 ? 2 : dtype)][D];
            if (gcache == 0) gcache = sweep_grid_size(dtype, sf32, D, lean, c->sm_count);
            int64_t need = (P + 1023) / 1024;
            if (need < 1) need = 1;
            wgrid = (int)(need < gcache ? need : gcache);
        }
        if (!cluster && getenv("AMC_SWEEP_DEBUG")) {
            const int gcache = c->sweep_grid_cache[lean ? 3 : (sf32 ? 2 : dtype)][D];
            static int told = 0;
            if (!told++) fprintf(stderr, "libamc sweep debug: cooperative grid %d (resident capacity %d on %d SMs), P=%lld\n", wgrid, gcache,
                                 c->sm_count, (long long)P);
        }
        if (!c->ev_k0) {
            CU(cudaEventCreate(&c->ev_k0));
            CU(cudaEventCreate(&c->ev_k1));
        }
        if ((rc = ensure(c->partials, (size_t)wgrid * kAccStride * 8))) return rc;
        const size_t sync_bytes = (size_t)(kSyncTickets + n_passes + 32) * 4;
        if (!cluster && (rc = ensure(c->syncbuf, sync_bytes))) return rc;
        const size_t tab_bytes = nrow * (sizeof(SweepTab) + sizeof(SolverTab));
        if ((rc = ensure(c->tabs, tab_bytes))) return rc;
        std::vector<unsigned char> tab_h(tab_bytes);
        SweepTab* wt = (SweepTab*)tab_h.data();
        SolverTab* st = (SolverTab*)(tab_h.data() + nrow * sizeof(SweepTab));
        for (int t = 0; t <= n; ++t) {
            wt[t].disc = exp(-rdt * (double)t);
            wt[t].mu = p->mu[t];
            wt[t].isg = 1.0 / p->sigma[t];
            wt[t].pad = 0.0;
            st[t].y_scale = exp(rdt * (double)t);
            st[t].mu = p->mu[t];
            st[t].sigma = 1.0 / (1.0 / p->sigma[t]);       // the scale the kernels effectively use
            st[t].pad = 0.0;
        }
        // the tables of a sweep depend on (r, dt, n) and the path set's column maps only: pricing the same set again (the
        // latency regime the cluster kernel exists for) finds them on the device already
        const bool tabs_cached = cluster && c->tabs_host.size() == tab_bytes + sizeof(void*) &&
                                 memcmp(c->tabs_host.data(), tab_h.data(), tab_bytes) == 0 &&
                                 memcmp(c->tabs_host.data() + tab_bytes, &c->tabs.p, sizeof(void*)) == 0;
        if (!tabs_cached) {
            CU(cudaMemcpyAsync(c->tabs.p, tab_h.data(), tab_bytes, cudaMemcpyHostToDevice, c->stream));
            c->tabs_host.assign(tab_h.begin(), tab_h.end());
            c->tabs_host.insert(c->tabs_host.end(), (const unsigned char*)&c->tabs.p, (const unsigned char*)&c->tabs.p + sizeof(void*));
        }
        if (!cluster) CU(cudaMemsetAsync(c->syncbuf.p, 0, sync_bytes, c->stream));
        int32_t* Lstate = nullptr;
        if (lean) {
            if ((rc = ensure(c->lstate, (size_t)ldp * 4))) return rc;
            Lstate = (int32_t*)c->lstate.p;
            CU(cudaMemcpyAsync(Lstate, p->Ln, (size_t)ldp * 4, cudaMemcpyDeviceToDevice, c->stream));
        }

        SweepArgs wa;
        memset(&wa, 0, sizeof(wa));
        wa.S = p->S;
        wa.ld = p->ld;
        wa.L = Lstate;
        wa.U = c->U.p;
        wa.tau = tau;
        wa.first_hit = fh;
        wa.tab = (const SweepTab*)c->tabs.p;
        wa.gamma = dg + off_gamma;
        wa.partials = (double*)c->partials.p;
        wa.sync = cluster ? nullptr : (uint32_t*)c->syncbuf.p;
        // AMC_CLUSTER_TRACE=1: the cluster kernel writes CTA 0's SM clock at 8 points of every pass (debug aid)
        static const int opt_trace = getenv("AMC_CLUSTER_TRACE") ? atoi(getenv("AMC_CLUSTER_TRACE")) : 0;
        if (cluster && opt_trace) {
            if ((rc = ensure(c->syncbuf, (size_t)n_passes * 64))) return rc;
            CU(cudaMemsetAsync(c->syncbuf.p, 0, (size_t)n_passes * 64, c->stream));
            wa.sync = (uint32_t*)c->syncbuf.p;
        }
        wa.n_paths = P;
        wa.n_steps = n;
        wa.n_passes = n_passes;
        wa.american = american ? 1 : 0;
        wa.is_put = spec->is_put;
        wa.reverse = opt_reverse;
        wa.K = spec->K;
        wa.gen = p->gen;
        wa.quad0 = p->path_offset >> 2;
        wa.rounds = p->rounds;

        SolveArgs& so = wa.solve;
        so.partials = (const double*)c->partials.p;
        so.n_rows = wgrid;
        so.sums = (double*)c->sums.p;
        so.spec = sspec;
        // inside the sweep kernel the scalar routine runs under the streaming loop's register cap (it spills): the
        // warp-cooperative routine takes every step it can certify, at every degree
        // (the cluster kernel runs one block per SM with the registers of the dedicated solve kernel: its rule applies)
        if (opt_warp_solve < 0 && !cluster) so.spec.warp_solve = 1;
        so.gamma = dg + off_gamma;
        so.beta = dg + off_beta;
        so.sv = dg + off_sv;
        so.mean_std = dg + off_ms;
        so.rank = drank;
        so.price = dg + off_price;
        so.n_batch = 1;
        so.gamma_stride = (int64_t)nrow * kMaxK;
        if (exchange) {
            for (int q = 0; q < c->world; ++q) so.peer.mailbox[q] = (uint4*)c->peer_mailbox[q];
            so.peer.world = c->world;
            so.peer.rank = c->rank;
            so.peer.err = c->peer_err;
            if (c->peer_seq > 0xFFF00000u) c->peer_seq = 0;        // same rule on every rank; a sequence number is never 0
            wa.seq_base = c->peer_seq;
            c->peer_seq += (uint32_t)n_passes;
        }
        wa.solve_tab = (const SolverTab*)((const char*)c->tabs.p + nrow * sizeof(SweepTab));

        CU(cudaEventRecord(c->ev_k0, c->stream));
        if (cluster) {
            const cudaError_t ce = launch_cluster_sweep(dtype, sf32, D, wa, c->stream);
            if (ce != cudaSuccess) {
                // not fatal: the launch chain prices the same set; this combination is not tried again
                cudaGetLastError();
                fprintf(stderr, "libamc: one-cluster sweep unavailable (%s); small path sets take the launch chain\n",
                        cudaGetErrorString(ce));
                c->cluster_cap[sf32 ? 2 : dtype][D] = 0;
                used_cluster = false;
                chain_fallback = true;
            }
        } else {
            CU(launch_sweep(dtype, sf32, D, lean, wgrid, wa, c->stream));
        }
        CU(cudaEventRecord(c->ev_k1, c->stream));
        n_step = chain_fallback ? 0 : 1;
        n_solve = 0;
    }
    if (!persistent || chain_fallback)
    {
    if (!regress) {
        // no early exercise and nobody wants continuation values: the price is the discounted mean payoff
        if ((rc = run_step(n, kMaturity, false))) return rc;
        if ((rc = run_solve(0, true))) return rc;
    } else {
        for (int t = n; t >= 0; --t) {
            const int mode = (t == n) ? kMaturity : (american ? kDecide : kObserve);
            if ((rc = run_step(t, mode, t > 0))) return rc;
            if ((rc = run_solve(t - 1, t == 0))) return rc;
        }
    }
    if (planned) {
        auto enqueue_plan = [&]() -> int {
            for (const PlanItem& it : plan) {
                if (it.kind == 0) CU(launch_step(dtype, sf32, D, grid, it.a, c->stream, it.pdl != 0, C));
                else CU(launch_solve(it.s, c->stream, it.pdl != 0));
            }
            return AMC_OK;
        };
        const int hdr[6] = {dtype, sf32, D, grid, C, (int)plan.size()};
        std::vector<unsigned char> key(sizeof(hdr) + plan.size() * sizeof(PlanItem));
        memcpy(key.data(), hdr, sizeof(hdr));
        if (!plan.empty()) memcpy(key.data() + sizeof(hdr), plan.data(), plan.size() * sizeof(PlanItem));
        bool done = false;
        if (c->graph_exec && key == c->graph_key) {
            CU(cudaGraphLaunch(c->graph_exec, c->stream));
            done = true;
        } else if (key == c->graph_seen) {
            // second identical sweep: capture it (a failed capture falls back to plain launches)
            cudaGraph_t g = nullptr;
            cudaGraphExec_t ge = nullptr;
            bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                ok = enqueue_plan() == AMC_OK;
                ok = (cudaStreamEndCapture(c->stream, &g) == cudaSuccess) && ok && g;
            }
            if (ok) ok = cudaGraphInstantiate(&ge, g, 0) == cudaSuccess;
            if (g) cudaGraphDestroy(g);
            if (ok) {
                if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
                c->graph_exec = ge;
                c->graph_key = key;
                CU(cudaGraphLaunch(c->graph_exec, c->stream));
                done = true;
            } else {
                cudaGetLastError();
                c->graph_seen.clear();
            }
        }
        if (!done) {
            c->graph_seen = key;
            if ((rc = enqueue_plan())) return rc;
        }
    }
    }
    CU(cudaEventRecord(ev_stop, c->stream));

    CU(cudaMemcpyAsync(price, dg + off_price, 8 * (size_t)C, cudaMemcpyDeviceToHost, c->stream));
    if (gamma_batch_out)
        CU(cudaMemcpyAsync(gamma_batch_out, dg + off_gamma, (size_t)C * nrow * kMaxK * 8, cudaMemcpyDeviceToHost, c->stream));
    std::vector<double> diag_h;
    std::vector<int> rank_h;
    if (steps) {
        diag_h.resize(off_price);
        rank_h.resize(nrow);
        CU(cudaMemcpyAsync(diag_h.data(), dg, off_price * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(rank_h.data(), drank, nrow * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    if (exercise_step_out && P > 0)
        CU(cudaMemcpyAsync(exercise_step_out, tau, (size_t)P * 4, cudaMemcpyDeviceToHost, c->stream));
    if (cashflow0_out && P > 0) {
        if (sf32) {
            if ((rc = ensure(c->misc, (size_t)P * 8))) return rc;
            CU(launch_column_to_f64(AMC_F32, c->U.p, P, (double*)c->misc.p, c->stream));
            CU(cudaMemcpyAsync(cashflow0_out, c->misc.p, (size_t)P * 8, cudaMemcpyDeviceToHost, c->stream));
        } else {
            CU(cudaMemcpyAsync(cashflow0_out, c->U.p, (size_t)P * 8, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    int peer_err_h = 0;
    uint32_t abort_h = 0;
    if (exchange && c->transport == 2) CU(cudaMemcpyAsync(&peer_err_h, c->peer_err, 4, cudaMemcpyDeviceToHost, c->stream));
    if (used_persistent)
        CU(cudaMemcpyAsync(&abort_h, (uint32_t*)c->syncbuf.p + kSyncAbort, 4, cudaMemcpyDeviceToHost, c->stream));
    if (exchange && c->transport == 1) {
        if ((rc = sync_with_nccl_watchdog(c))) return rc;
    } else {
        CU(cudaStreamSynchronize(c->stream));
    }
    if (used_cluster && getenv("AMC_CLUSTER_TRACE") && atoi(getenv("AMC_CLUSTER_TRACE")) && n >= 4) {
        static int traced = 0;
        if (++traced == 8) {                 // once, on a warm call
            const int np = regress ? n + 1 : 1;
            std::vector<long long> tr((size_t)np * 8);
            cudaMemcpy(tr.data(), c->syncbuf.p, tr.size() * 8, cudaMemcpyDeviceToHost);
            double d[8] = {};
            int cntp = 0;
            for (int q = 2; q + 2 < np; ++q, ++cntp) {
                for (int k = 0; k < 7; ++k) d[k] += (double)(tr[q * 8 + k + 1] - tr[q * 8 + k]);
                d[7] += (double)(tr[(q + 1) * 8] - tr[q * 8 + 7]);
            }
            fprintf(stderr, "libamc cluster trace (SM cycles per pass, mean over %d passes; P=%lld D=%d): top->waited %.0f | loop %.0f | "
                            "block reduce %.0f | cluster barrier %.0f | gather %.0f | solve %.0f | outputs %.0f | loop-back %.0f\n",
                    cntp, (long long)P, D, d[0] / cntp, d[1] / cntp, d[2] / cntp, d[3] / cntp, d[4] / cntp, d[5] / cntp, d[6] / cntp,
                    d[7] / cntp);
        }
    }
    if (abort_h && getenv("AMC_SWEEP_DEBUG")) {
        std::vector<uint32_t> w(kSyncTickets + 8);
        cudaMemcpy(w.data(), c->syncbuf.p, w.size() * 4, cudaMemcpyDeviceToHost);
        fprintf(stderr, "libamc sweep debug: published=%u abort=%u tickets=%u %u %u %u (P=%lld n=%d D=%d dtype=%d sf32=%d)\n",
                w[kSyncPublished], w[kSyncAbort], w[kSyncTickets], w[kSyncTickets + 1], w[kSyncTickets + 2],
                w[kSyncTickets + 3], (long long)P, n, D, dtype, sf32);
    }
    if (abort_h && !peer_err_h)
        return fail(AMC_ERR_CUDA, "the persistent sweep timed out waiting for one of its own passes; AMC_PERSISTENT=0 selects the "
                                  "per-step launch chain");
    if (peer_err_h) {
        cudaMemsetAsync(c->peer_err, 0, 4, c->stream);
        return fail(AMC_ERR_NCCL, "peer-memory all-reduce timed out: a rank did not reach the same step of the sweep");
    }

    if (steps) {
        if (steps->gamma) memcpy(steps->gamma, diag_h.data() + off_gamma, nrow * kMaxK * 8);
        if (steps->beta) memcpy(steps->beta, diag_h.data() + off_beta, nrow * kMaxK * 8);
        if (steps->sv) memcpy(steps->sv, diag_h.data() + off_sv, nrow * kMaxK * 8);
        for (size_t t = 0; t < nrow; ++t) {
            if (steps->mean_x) steps->mean_x[t] = diag_h[off_ms + 3 * t];
            if (steps->std_x) steps->std_x[t] = diag_h[off_ms + 3 * t + 1];
            if (steps->pivot_loss) steps->pivot_loss[t] = diag_h[off_ms + 3 * t + 2];
            if (steps->rank) steps->rank[t] = rank_h[t];
        }
    }
    if (timing) {
        memset(timing, 0, sizeof(*timing));
        CU(cudaEventElapsedTime(&timing->total_ms, ev_start, ev_stop));
        for (size_t i = 0; i + 1 < step_ev.size(); i += 2) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, step_ev[i], step_ev[i + 1]));
            timing->step_kernel_ms += ms;
        }
        for (size_t i = 0; i + 1 < solve_ev.size(); i += 2) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, solve_ev[i], solve_ev[i + 1]));
            timing->solve_kernel_ms += ms;
        }
        if (used_persistent || used_cluster) CU(cudaEventElapsedTime(&timing->step_kernel_ms, c->ev_k0, c->ev_k1));
        timing->step_launches = n_step;
        timing->solve_launches = n_solve;
        timing->other_launches = n_other;
        timing->sweep_kind = used_cluster ? 2 : (used_persistent ? 1 : 0);
    }
    return AMC_OK;
}

extern "C" int amc_lsm_price(amc_ctx* c, const amc_paths* p, const amc_lsm_spec* spec, double* price,
                             amc_lsm_steps* steps, int32_t* exercise_step_out, double* cashflow0_out,
                             amc_lsm_timing* timing, int profile) {
    if (!c || !p || !price) return fail(AMC_ERR_VALUE, "amc_lsm_price: null argument");
    if (p->ctx != c) return fail(AMC_ERR_STATE, "amc_lsm_price: path set belongs to another context");
    int rc = check_spec(spec);
    if (rc) return rc;
    if (exercise_step_out && !spec->want_exercise_steps)
        return fail(AMC_ERR_VALUE, "exercise_step_out needs spec.want_exercise_steps");
    if (spec->state_f32 && p->dtype != AMC_F32)
        return fail(AMC_ERR_VALUE, "state_f32 needs a float32 path set (a float state under float64 paths would cap the accuracy)");
    return lsm_sweep(c, p, spec, 1, price, steps, exercise_step_out, cashflow0_out, nullptr, timing, profile);
}

extern "C" int amc_lsm_price_with_hits(amc_ctx* c, const amc_paths* p, const amc_lsm_spec* spec,
                                       const int32_t* first_hit, double* price, amc_lsm_steps* steps,
                                       int32_t* exercise_step_out, double* cashflow0_out, amc_lsm_timing* timing,
                                       int profile) {
    if (!c || !p || !price) return fail(AMC_ERR_VALUE, "amc_lsm_price_with_hits: null argument");
    if (p->ctx != c) return fail(AMC_ERR_STATE, "amc_lsm_price_with_hits: path set belongs to another context");
    int rc = check_spec(spec);
    if (rc) return rc;
    if (exercise_step_out && !spec->want_exercise_steps)
        return fail(AMC_ERR_VALUE, "exercise_step_out needs spec.want_exercise_steps");
    if (spec->state_f32 && p->dtype != AMC_F32) return fail(AMC_ERR_VALUE, "state_f32 needs a float32 path set");
    return lsm_sweep(c, p, spec, 1, price, steps, exercise_step_out, cashflow0_out, nullptr, timing, profile, first_hit);
}

extern "C" int amc_paths_gather_steps(const amc_paths* p, const int32_t* steps, double* out) {
    if (!p || !steps || !out) return fail(AMC_ERR_VALUE, "amc_paths_gather_steps: null argument");
    if (p->n_local == 0) return AMC_OK;
    amc_ctx* c = p->ctx;
    if (!c) return fail(AMC_ERR_STATE, "the path set's context has been destroyed");
    CU(cudaSetDevice(c->device));
    int rc = ensure(c->misc, (size_t)p->n_local * 12);
    if (rc) return rc;
    double* out_dev = (double*)c->misc.p;
    int32_t* st_dev = (int32_t*)(out_dev + p->n_local);
    CU(cudaMemcpyAsync(st_dev, steps, (size_t)p->n_local * 4, cudaMemcpyHostToDevice, c->stream));
    if (p->lean)
        CU(launch_lean_walk(3, p->rounds, p->gen, p->path_offset >> 2, p->n_local, p->n_steps, p->n_steps, 0, p->n_local, 0.0,
                            st_dev, out_dev, nullptr, c->sm_count, c->stream));
    else
        CU(launch_gather_steps(p->dtype, p->S, p->ld, p->n_steps + 1, p->n_local, st_dev, out_dev, c->stream));
    CU(cudaMemcpyAsync(out, out_dev, (size_t)p->n_local * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMC_OK;
}

extern "C" int amc_lsm_price_batch(amc_ctx* c, const amc_paths* p, const amc_lsm_spec* specs, int n_contracts,
                                   double* prices, double* gamma_out, amc_lsm_timing* timing, int profile) {
    if (!c || !p || !prices || !specs) return fail(AMC_ERR_VALUE, "amc_lsm_price_batch: null argument");
    if (p->ctx != c) return fail(AMC_ERR_STATE, "amc_lsm_price_batch: path set belongs to another context");
    if (n_contracts < 1 || n_contracts > AMC_MAX_BATCH)
        return fail(AMC_ERR_VALUE, "amc_lsm_price_batch: n_contracts %d outside 1..%d", n_contracts, AMC_MAX_BATCH);
    if (p->lean) return fail(AMC_ERR_VALUE, "amc_lsm_price_batch: path-free sets are priced one contract at a time (amc_lsm_price)");
    if (p->n_global != p->n_local)
        return fail(AMC_ERR_VALUE, "amc_lsm_price_batch: the path set is sharded over ranks; batches are sharded by "
                                   "contract (every rank prices its own contracts on its own complete path sets)");
    for (int i = 0; i < n_contracts; ++i) {
        int rc = check_spec(&specs[i]);
        if (rc) return rc;
        const amc_lsm_spec &a = specs[0], &b = specs[i];
        const bool same_barrier = (isnan(a.barrier) && isnan(b.barrier)) || a.barrier == b.barrier;
        if (a.r != b.r || a.dt != b.dt || a.basis != b.basis || a.degree != b.degree || a.scaling != b.scaling ||
            a.scaling_factor != b.scaling_factor || !same_barrier)
            return fail(AMC_ERR_VALUE, "amc_lsm_price_batch: contract %d differs from contract 0 in r/dt/basis/degree/"
                                       "scaling/barrier; only strike, payoff side and exercise style may vary", i);
        if (a.state_f32 != b.state_f32) return fail(AMC_ERR_VALUE, "amc_lsm_price_batch: contract %d differs in state_f32", i);
        if (b.state_f32 && p->dtype != AMC_F32) return fail(AMC_ERR_VALUE, "state_f32 needs a float32 path set");
        if (b.want_exercise_steps) return fail(AMC_ERR_VALUE, "amc_lsm_price_batch: want_exercise_steps is per contract; use amc_lsm_price");
    }
    if (n_contracts == 1) {
        amc_lsm_spec one = specs[0];
        if (gamma_out) one.want_regression = 1;
        int rc = lsm_sweep(c, p, &one, 1, prices, nullptr, nullptr, nullptr, gamma_out, timing, profile);
        return rc;
    }
    return lsm_sweep(c, p, specs, n_contracts, prices, nullptr, nullptr, nullptr, gamma_out, timing, profile);
}

extern "C" int amc_continuation(amc_ctx* c, const amc_paths* p, int t, const double* gamma, int degree, double* out) {
    if (!c || !p || !gamma || !out) return fail(AMC_ERR_VALUE, "amc_continuation: null argument");
    if (t < 0 || t > p->n_steps) return fail(AMC_ERR_VALUE, "step %d out of range", t);
    if (degree < 0 || degree > AMC_MAX_DEGREE) return fail(AMC_ERR_VALUE, "degree %d outside 0..%d", degree, AMC_MAX_DEGREE);
    if (p->n_local == 0) return AMC_OK;
    CU(cudaSetDevice(c->device));
    int rc = ensure(c->misc, (size_t)p->n_local * 8 + kMaxK * 8);
    if (rc) return rc;
    double* out_dev = (double*)c->misc.p;
    double* gam_dev = out_dev + p->n_local;
    CU(cudaMemcpyAsync(gam_dev, gamma, (size_t)(degree + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    if (p->lean) {
        if ((rc = ensure(c->lstate, (size_t)p->n_local * 8))) return rc;
        CU(launch_lean_walk(0, p->rounds, p->gen, p->path_offset >> 2, p->n_local, p->n_steps, t, 0, p->n_local, 0.0, nullptr,
                            (double*)c->lstate.p, nullptr, c->sm_count, c->stream));
        CU(launch_continuation(AMC_F64, c->lstate.p, p->n_local, gam_dev, degree, p->mu[t], 1.0 / p->sigma[t], 1, out_dev,
                               c->stream));
    } else {
        CU(launch_continuation(p->dtype, column(p, t), p->n_local, gam_dev, degree, p->mu[t], 1.0 / p->sigma[t], 1, out_dev,
                               c->stream));
    }
    CU(cudaMemcpyAsync(out, out_dev, (size_t)p->n_local * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMC_OK;
}

// ---------------------------------------------------------------------------------------------------------
// exposures: compute_ccr_exposures, amc.py:400-414
static int ccr_run(amc_ctx* c, const CcrSource& src, int64_t n_local, bool exchange, double q_lo, double q_hi,
                   double* out3_dev, SelState* st, unsigned long long* hist, double* partials, int grid) {
    CU(cudaMemsetAsync(st, 0, sizeof(SelState), c->stream));
    for (int pass = 0; pass < kSelPasses; ++pass) {
        CU(launch_ccr_hist(src, n_local, st, pass, hist, partials, grid, c->stream));
        if (exchange) {        // sharded paths: the histograms (and, once, the partial sums) are global
            NC(g_nccl.AllReduce(hist, hist, (size_t)kSelTargets * kSelBins, ncclUint64, ncclSum, c->comm, c->stream));
            if (pass == 0) NC(g_nccl.AllReduce(partials, partials, (size_t)grid, ncclFloat64, ncclSum, c->comm, c->stream));
        }
        CU(launch_ccr_scan(st, hist, pass, partials, grid, q_lo, q_hi, out3_dev, c->stream));
    }
    return AMC_OK;
}

static int ccr_scratch(amc_ctx* c, int n_out, int64_t n_local, bool across_ranks, double** out_dev, SelState** st,
                       unsigned long long** hist, double** partials, int* grid) {
    int64_t g = (n_local + 255) / 256;
    const int64_t cap = (int64_t)c->sm_count * 4;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    if (c->world > 1 && across_ranks) g = cap;     // the same on every rank: the partial sums are all-reduced elementwise
    *grid = (int)g;
    const size_t hist_bytes = (size_t)kSelTargets * kSelBins * 8;
    const size_t bytes = hist_bytes + 256 + (size_t)g * 8 + (size_t)n_out * 3 * 8;
    int rc = ensure(c->ccr, bytes);
    if (rc) return rc;
    char* base = (char*)c->ccr.p;
    *hist = (unsigned long long*)base;
    *st = (SelState*)(base + hist_bytes);
    *partials = (double*)(base + hist_bytes + 256);
    *out_dev = (double*)(base + hist_bytes + 256 + (size_t)g * 8);
    static_assert(sizeof(SelState) <= 256, "SelState");
    CU(cudaMemsetAsync(base, 0, hist_bytes, c->stream));
    return AMC_OK;
}

extern "C" int amc_ccr_exposures(amc_ctx* c, const amc_paths* p, const double* gamma, int degree, double q_lo,
                                 double q_hi, double* pfe_lo, double* pfe_hi, double* epe) {
    if (!c || !p || !gamma || !pfe_lo || !pfe_hi || !epe) return fail(AMC_ERR_VALUE, "amc_ccr_exposures: null argument");
    if (p->ctx != c) return fail(AMC_ERR_STATE, "amc_ccr_exposures: path set belongs to another context");
    if (degree < 0 || degree > AMC_MAX_DEGREE) return fail(AMC_ERR_VALUE, "degree %d outside 0..%d", degree, AMC_MAX_DEGREE);
    if (!(q_lo >= 0.0 && q_lo <= 1.0 && q_hi >= 0.0 && q_hi <= 1.0))
        return fail(AMC_ERR_VALUE, "Percentiles must be in the range [0, 100]");
    CU(cudaSetDevice(c->device));
    const int n = p->n_steps;
    double* out_dev; SelState* st; unsigned long long* hist; double* partials; int grid;
    const bool exchange = c->world > 1 && p->n_global != p->n_local;
    int rc = ccr_scratch(c, n + 1, p->n_local, exchange, &out_dev, &st, &hist, &partials, &grid);
    if (rc) return rc;
    for (int t = 0; t <= n; ++t) {
        CcrSource src;
        memset(&src, 0, sizeof(src));
        if (p->lean) {                                  // materialise the column (walks the counters forward to step t)
            if ((rc = ensure(c->lstate, (size_t)p->n_local * 8))) return rc;
            CU(launch_lean_walk(0, p->rounds, p->gen, p->path_offset >> 2, p->n_local, n, t, 0, p->n_local, 0.0, nullptr,
                                (double*)c->lstate.p, nullptr, c->sm_count, c->stream));
            src.x = c->lstate.p;
            src.x_f32 = 0;
        } else {
            src.x = column(p, t);
            src.x_f32 = p->dtype == AMC_F32;
        }
        src.degree = degree;
        src.clamp = 1;                                  // np.maximum(fit, 0), amc.py:132
        src.zero = (t == n);                            // amc.py:145: zeros at maturity
        src.mu = p->mu[t];
        src.isg = 1.0 / p->sigma[t];
        for (int i = 0; i <= degree; ++i) src.gam[i] = gamma[(size_t)t * kMaxK + i];
        if ((rc = ccr_run(c, src, p->n_local, exchange, q_lo, q_hi, out_dev + 3 * t, st, hist, partials, grid))) return rc;
    }
    std::vector<double> h((size_t)(n + 1) * 3);
    CU(cudaMemcpyAsync(h.data(), out_dev, h.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int t = 0; t <= n; ++t) { pfe_lo[t] = h[3 * t]; pfe_hi[t] = h[3 * t + 1]; epe[t] = h[3 * t + 2]; }
    return AMC_OK;
}

extern "C" int amc_percentiles(amc_ctx* c, const double* values, int64_t n, double q_lo, double q_hi, double out3[3]) {
    if (!c || !out3 || (n > 0 && !values)) return fail(AMC_ERR_VALUE, "amc_percentiles: null argument");
    if (!(q_lo >= 0.0 && q_lo <= 1.0 && q_hi >= 0.0 && q_hi <= 1.0))
        return fail(AMC_ERR_VALUE, "Percentiles must be in the range [0, 100]");
    CU(cudaSetDevice(c->device));
    double* out_dev; SelState* st; unsigned long long* hist; double* partials; int grid;
    int rc = ccr_scratch(c, 1, n, false, &out_dev, &st, &hist, &partials, &grid);       // this array only
    if (rc) return rc;
    if ((rc = ensure(c->misc, (size_t)(n > 0 ? n : 1) * 8))) return rc;
    if (n > 0) CU(cudaMemcpyAsync(c->misc.p, values, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CcrSource src;
    memset(&src, 0, sizeof(src));
    src.vals = (const double*)c->misc.p;
    if ((rc = ccr_run(c, src, n, false, q_lo, q_hi, out_dev, st, hist, partials, grid))) return rc;
    CU(cudaMemcpyAsync(out3, out_dev, 24, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMC_OK;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int amc_intrinsic_value(amc_ctx* c, const double* S, int64_t n, double K, int is_put, double* out) {
    if (!c || (n > 0 && (!S || !out))) return fail(AMC_ERR_VALUE, "amc_intrinsic_value: null argument");
    if (n <= 0) return AMC_OK;
    CU(cudaSetDevice(c->device));
    int rc = ensure(c->misc, (size_t)n * 16);
    if (rc) return rc;
    double* in_dev = (double*)c->misc.p;
    double* out_dev = in_dev + n;
    CU(cudaMemcpyAsync(in_dev, S, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CU(launch_intrinsic(in_dev, n, K, is_put, out_dev, c->stream));
    CU(cudaMemcpyAsync(out, out_dev, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMC_OK;
}

extern "C" int amc_basis_matrix(amc_ctx* c, const double* X, int64_t n, int basis, int degree, double* out) {
    if (!c || (n > 0 && (!X || !out))) return fail(AMC_ERR_VALUE, "amc_basis_matrix: null argument");
    if (basis < 0 || basis > AMC_BASIS_LAGUERRE)
        return fail(AMC_ERR_VALUE, "Unknown basis type id %d. Use 'Power', 'Chebyshev', or 'Legendre'.", basis);
    if (degree < 0 || degree > AMC_MAX_DEGREE) return fail(AMC_ERR_VALUE, "degree %d outside 0..%d", degree, AMC_MAX_DEGREE);
    if (n <= 0) return AMC_OK;
    CU(cudaSetDevice(c->device));
    const size_t k = (size_t)degree + 1;
    int rc = ensure(c->misc, (size_t)n * 8 * (k + 1));
    if (rc) return rc;
    double* in_dev = (double*)c->misc.p;
    double* out_dev = in_dev + n;
    CU(cudaMemcpyAsync(in_dev, X, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CU(launch_basis_matrix(in_dev, n, basis, degree, out_dev, c->stream));
    CU(cudaMemcpyAsync(out, out_dev, (size_t)n * k * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMC_OK;
}

// Y either comes from the host (Y) or is produced on the device by `prepare_y` (writes n doubles at c->U.p)
template <typename PrepareY>
static int regression_fit_impl(amc_ctx* c, const double* X, const double* Y, PrepareY prepare_y, int64_t n, int basis,
                               int degree, int scaling, double scaling_factor, int clamp, double* fitted, double* beta,
                               int* rank) {
    amc_lsm_spec sp;
    memset(&sp, 0, sizeof(sp));
    sp.basis = basis;
    sp.degree = degree;
    int rc = check_spec(&sp);
    if (rc) return rc;
    if (n <= 0) return AMC_OK;
    CU(cudaSetDevice(c->device));
    // a one-column path set holding X; Y plays the role of the per-path state
    amc_paths* px = nullptr;
    if ((rc = paths_alloc(c, 0, n, n, AMC_F64, &px, kPathBorrowed))) return rc;
    auto cleanup = [&](int code) { amc_paths_free(px); return code; };
    cudaError_t e = cudaMemcpyAsync(px->S, X, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) return cleanup(fail(AMC_ERR_CUDA, "H2D X: %s", cudaGetErrorString(e)));
    rc = measured_maps(c, px, false);               // statistics of THIS array only
    if (rc) return cleanup(rc);
    const int grid = step_grid(c, AMC_F64, 0, degree, n);
    const int64_t ldp = padded_len(n);
    if ((rc = ensure(c->U, (size_t)ldp * 8)) || (rc = ensure(c->partials, (size_t)grid * kAccStride * 8)) ||
        (rc = ensure(c->sums, kAccStride * 8)) || (rc = ensure(c->diag, (4 * kMaxK + 8) * 8)))
        return cleanup(rc);
    if (Y) {
        e = cudaMemcpyAsync(c->U.p, Y, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) return cleanup(fail(AMC_ERR_CUDA, "H2D Y: %s", cudaGetErrorString(e)));
    } else if ((rc = prepare_y((double*)c->U.p))) {
        return cleanup(rc);
    }
    double* dg = (double*)c->diag.p;
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.x_reg = px->S;
    a.U = c->U.p;
    a.partials = (double*)c->partials.p;
    a.n_paths = n;
    a.t_dec = 1;
    a.mode = kObserve;
    a.moments = 1;
    a.mu_reg = px->mu[0];
    a.isg_reg = 1.0 / px->sigma[0];
    e = launch_step(AMC_F64, 0, degree, grid, a, c->stream);
    if (e != cudaSuccess) return cleanup(fail(AMC_ERR_CUDA, "moment kernel: %s", cudaGetErrorString(e)));
    SolveArgs s;
    memset(&s, 0, sizeof(s));
    s.partials = (const double*)c->partials.p;
    s.n_rows = grid;
    s.sums = (double*)c->sums.p;
    s.do_reduce = 1;
    s.do_solve = 1;
    s.spec.degree = degree;
    s.spec.basis = basis;
    s.spec.scaling = scaling;
    s.spec.want_svd = 0;
    s.spec.scaling_factor = scaling_factor;
    s.spec.n_paths = (double)n;
    s.spec.inv_n_paths = n > 0 ? 1.0 / (double)n : 0.0;
    s.y_scale = 1.0;
    s.mu_ref = px->mu[0];
    s.sigma_ref = 1.0 / a.isg_reg;
    s.gamma = dg;
    s.beta = dg + kMaxK;
    s.sv = dg + 2 * kMaxK;
    s.mean_std = dg + 3 * kMaxK;
    s.rank = (int*)(dg + 3 * kMaxK + 3);
    s.price = dg + 3 * kMaxK + 4;
    e = launch_solve(s, c->stream);
    if (e != cudaSuccess) return cleanup(fail(AMC_ERR_CUDA, "solve kernel: %s", cudaGetErrorString(e)));
    if ((rc = ensure(c->misc, (size_t)n * 8))) return cleanup(rc);
    e = launch_continuation(AMC_F64, px->S, n, dg, degree, px->mu[0], a.isg_reg, clamp, (double*)c->misc.p, c->stream);
    if (e != cudaSuccess) return cleanup(fail(AMC_ERR_CUDA, "fit kernel: %s", cudaGetErrorString(e)));
    double small[3 * kMaxK + 4];
    cudaMemcpyAsync(fitted, c->misc.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(small, dg, sizeof(small), cudaMemcpyDeviceToHost, c->stream);
    e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cleanup(fail(AMC_ERR_CUDA, "regression fit: %s", cudaGetErrorString(e)));
    if (beta) memcpy(beta, small + kMaxK, (size_t)(degree + 1) * 8);
    if (rank) memcpy(rank, small + 3 * kMaxK + 3, 4);
    return cleanup(AMC_OK);
}

extern "C" int amc_regression_fit(amc_ctx* c, const double* X, const double* Y, int64_t n, int basis, int degree,
                                  int scaling, double scaling_factor, double* fitted, double* beta, int* rank) {
    if (!c || (n > 0 && (!X || !Y || !fitted))) return fail(AMC_ERR_VALUE, "amc_regression_fit: null argument");
    return regression_fit_impl(c, X, Y, [](double*) { return (int)AMC_OK; }, n, basis, degree, scaling, scaling_factor, 0,
                               fitted, beta, rank);
}

extern "C" int amc_estimate_continuation(amc_ctx* c, const double* X, const double* cashflows,
                                         const int64_t* exercise_times, int64_t n, int64_t t, double r, double dt, int basis,
                                         int degree, int scaling, double scaling_factor, double* out) {
    if (!c || (n > 0 && (!X || !cashflows || !exercise_times || !out)))
        return fail(AMC_ERR_VALUE, "amc_estimate_continuation: null argument");
    auto prepare = [&](double* y_dev) -> int {
        // Y = cashflows * exp(-r dt (tau - t)), amc.py:128, formed on the device
        int rc = ensure(c->stage, (size_t)n * 16);
        if (rc) return rc;
        double* cf_dev = (double*)c->stage.p;
        int64_t* tau_dev = (int64_t*)(cf_dev + n);
        CU(cudaMemcpyAsync(cf_dev, cashflows, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(tau_dev, exercise_times, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(launch_discount(cf_dev, tau_dev, n, t, r, dt, y_dev, c->stream));
        return AMC_OK;
    };
    return regression_fit_impl(c, X, nullptr, prepare, n, basis, degree, scaling, scaling_factor, 1, out, nullptr, nullptr);
}

extern "C" int amc_apply_exercise(amc_ctx* c, double* cashflows, int64_t* exercise_times, int64_t n_total,
                                  const double* exercise_value, const double* continuation, const int64_t* indices,
                                  int64_t m, int64_t t) {
    if (!c || (n_total > 0 && (!cashflows || !exercise_times)) || (m > 0 && (!exercise_value || !continuation || !indices)))
        return fail(AMC_ERR_VALUE, "amc_apply_exercise: null argument");
    if (m <= 0 || n_total <= 0) return AMC_OK;
    for (int64_t i = 0; i < m; ++i)
        if (indices[i] < 0 || indices[i] >= n_total)
            return fail(AMC_ERR_VALUE, "index %lld is out of bounds for axis 0 with size %lld", (long long)indices[i], (long long)n_total);
    CU(cudaSetDevice(c->device));
    int rc = ensure(c->misc, (size_t)n_total * 16 + (size_t)m * 24);
    if (rc) return rc;
    double* cf_dev = (double*)c->misc.p;
    int64_t* tau_dev = (int64_t*)(cf_dev + n_total);
    double* ev_dev = (double*)(tau_dev + n_total);
    double* ce_dev = ev_dev + m;
    int64_t* idx_dev = (int64_t*)(ce_dev + m);
    CU(cudaMemcpyAsync(cf_dev, cashflows, (size_t)n_total * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(tau_dev, exercise_times, (size_t)n_total * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(ev_dev, exercise_value, (size_t)m * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(ce_dev, continuation, (size_t)m * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(idx_dev, indices, (size_t)m * 8, cudaMemcpyHostToDevice, c->stream));
    CU(launch_apply_exercise(cf_dev, tau_dev, ev_dev, ce_dev, idx_dev, m, t, c->stream));
    CU(cudaMemcpyAsync(cashflows, cf_dev, (size_t)n_total * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(exercise_times, tau_dev, (size_t)n_total * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMC_OK;
}

extern "C" int amc_barrier_hit_matrix(amc_ctx* c, const amc_paths* p, double barrier, uint8_t* out) {
    if (!c || !p || !out) return fail(AMC_ERR_VALUE, "amc_barrier_hit_matrix: null argument");
    if (p->n_local == 0) return AMC_OK;
    CU(cudaSetDevice(c->device));
    const size_t total = (size_t)p->n_local * (size_t)(p->n_steps + 1);
    if (isnan(barrier)) {                      // amc.py:175: no barrier -> all True
        memset(out, 1, total);
        return AMC_OK;
    }
    int rc = ensure(c->first_hit, (size_t)padded_len(p->n_local) * 4);
    if (rc) return rc;
    if ((rc = ensure(c->misc, total))) return rc;
    if (p->lean)
        CU(launch_lean_walk(2, p->rounds, p->gen, p->path_offset >> 2, p->n_local, p->n_steps, p->n_steps, 0, p->n_local,
                            barrier, nullptr, nullptr, (int32_t*)c->first_hit.p, c->sm_count, c->stream));
    else
        CU(launch_first_hit(p->dtype, p->S, p->ld, p->n_steps + 1, p->n_local, barrier, (int32_t*)c->first_hit.p, c->stream));
    CU(launch_hit_matrix((const int32_t*)c->first_hit.p, p->n_steps + 1, p->n_local, (uint8_t*)c->misc.p, c->stream));
    CU(cudaMemcpyAsync(out, c->misc.p, total, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMC_OK;
}

// ---------------------------------------------------------------------------------------------------------
// generator self-tests (see pathgen.cu)
extern "C" int amc_selftest_philox(amc_ctx* c, int rounds, const uint32_t* counters, const uint32_t key[2], int n,
                                   uint32_t* out) {
    if (!c || !counters || !key || !out || n < 0) return fail(AMC_ERR_VALUE, "amc_selftest_philox: bad argument");
    if (rounds != 10 && rounds != 7) return fail(AMC_ERR_VALUE, "amc_selftest_philox: rounds must be 10 or 7");
    if (n == 0) return AMC_OK;
    CU(cudaSetDevice(c->device));
    int rc = ensure(c->misc, (size_t)n * 32);
    if (rc) return rc;
    uint32_t* ctr_dev = (uint32_t*)c->misc.p;
    uint32_t* out_dev = ctr_dev + (size_t)n * 4;
    CU(cudaMemcpyAsync(ctr_dev, counters, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
    CU(launch_philox_kat(rounds, ctr_dev, key[0], key[1], n, out_dev, c->stream));
    CU(cudaMemcpyAsync(out, out_dev, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMC_OK;
}

extern "C" int amc_selftest_normals(amc_ctx* c, int rounds, uint64_t seed, int64_t n_quads, int n_steps, int n_bins,
                                    double lo, double hi, uint64_t* hist, double stats[6]) {
    if (!c || !hist || !stats || n_quads < 0 || n_steps < 1)
        return fail(AMC_ERR_VALUE, "amc_selftest_normals: bad argument");
    if (rounds != 10 && rounds != 7) return fail(AMC_ERR_VALUE, "amc_selftest_normals: rounds must be 10 or 7");
    if (n_bins < 1 || n_bins > 4096 || !(hi > lo)) return fail(AMC_ERR_VALUE, "amc_selftest_normals: 1..4096 bins on [lo, hi)");
    CU(cudaSetDevice(c->device));
    const int grid = c->sm_count * 4;
    const size_t hist_bytes = (size_t)(n_bins + 2) * 8;
    int rc = ensure(c->misc, hist_bytes + (size_t)grid * 6 * 8);
    if (rc) return rc;
    unsigned long long* hist_dev = (unsigned long long*)c->misc.p;
    double* stats_dev = (double*)((char*)c->misc.p + hist_bytes);
    CU(cudaMemsetAsync(c->misc.p, 0, hist_bytes + (size_t)grid * 6 * 8, c->stream));
    CU(launch_normals_hist(rounds, seed, n_quads, n_steps, n_bins, lo, hi, hist_dev, stats_dev, grid, c->stream));
    std::vector<double> st((size_t)grid * 6);
    CU(cudaMemcpyAsync(hist, hist_dev, hist_bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(st.data(), stats_dev, st.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int j = 0; j < 6; ++j) stats[j] = 0.0;
    for (int b = 0; b < grid; ++b) {
        for (int j = 0; j < 5; ++j) stats[j] += st[(size_t)b * 6 + j];
        if (st[(size_t)b * 6 + 5] > stats[5]) stats[5] = st[(size_t)b * 6 + 5];
    }
    return AMC_OK;
}
