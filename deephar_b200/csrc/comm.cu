// The one exchange step of the data-parallel forward path: an all-gather of the
// per-rank outputs -- action probabilities (B_local, n_pred, n_act) and, optionally, poses -- over NCCL
// (NVLink 5 / NVSwitch).  The reference has no multi-GPU code path at all (one process, one GPU:
// exp/*/eval_*.py); this is what a data-parallel evaluator would call after Model.predict on its shard.
//
// NCCL is bound at RUN time (dlopen of the libnccl.so.2 already loaded by the host process -- torch's bundled
// copy -- else the system one): libdeephar_b200.so has no link-time dependency on it and single-GPU users never
// touch it.  One communicator per dh_ctx; every call is asynchronous on the caller's stream.
#include <dlfcn.h>
#include <string.h>
#include "common.cuh"

namespace {

typedef struct { char internal[128]; } nccl_uid_t;
typedef void* nccl_comm_t;
typedef int (*fn_get_uid)(nccl_uid_t*);
typedef int (*fn_init_rank)(nccl_comm_t*, int, nccl_uid_t, int);
typedef int (*fn_allgather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_destroy)(nccl_comm_t);
typedef const char* (*fn_errstr)(int);
typedef int (*fn_version)(int*);

struct Nccl {
    void* lib = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_allgather allgather = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
    fn_version version = nullptr;
};

Nccl* nccl() {
    static Nccl n;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the copy the process already uses
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            n.get_uid = (fn_get_uid)dlsym(h, "ncclGetUniqueId");
            n.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
            n.allgather = (fn_allgather)dlsym(h, "ncclAllGather");
            n.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
            n.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
            n.version = (fn_version)dlsym(h, "ncclGetVersion");
            if (n.get_uid && n.init_rank && n.allgather && n.destroy) n.lib = h;
        }
    }
    return n.lib ? &n : nullptr;
}

int nccl_fail(const char* who, int rc) {
    Nccl* n = nccl();
    dh_set_error("%s: NCCL error %d (%s)", who, rc, (n && n->errstr) ? n->errstr(rc) : "?");
    return 1000 + rc;      // > 0: library error (cudaError_t / ncclResult_t + 1000), see deephar_b200.h
}

}  // namespace

extern "C" int dh_comm_unique_id(void* out128) {
    DH_CHECK_ARG(out128 != nullptr, "dh_comm_unique_id: NULL buffer");
    Nccl* n = nccl();
    DH_CHECK_ARG(n != nullptr, "dh_comm_unique_id: libnccl.so.2 could not be loaded (%s)", dlerror());
    nccl_uid_t id;
    int rc = n->get_uid(&id);
    if (rc) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(out128, &id, sizeof(id));
    return 0;
}

extern "C" int dh_comm_init(dh_ctx* ctx, int rank, int world, const void* unique_id128) {
    DH_CHECK_ARG(ctx && unique_id128, "dh_comm_init: NULL argument");
    DH_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "dh_comm_init: rank %d of %d", rank, world);
    DH_CHECK_ARG(ctx->comm == nullptr, "dh_comm_init: this context already has a communicator");
    Nccl* n = nccl();
    DH_CHECK_ARG(n != nullptr, "dh_comm_init: libnccl.so.2 could not be loaded");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) { dh_set_error("dh_comm_init: cudaSetDevice: %s", cudaGetErrorString(e)); return (int)e; }
    nccl_uid_t id;
    memcpy(&id, unique_id128, sizeof(id));
    nccl_comm_t comm = nullptr;
    int rc = n->init_rank(&comm, world, id, rank);
    if (rc) return nccl_fail("ncclCommInitRank", rc);
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    return 0;
}

extern "C" int dh_comm_destroy(dh_ctx* ctx) {
    DH_CHECK_ARG(ctx != nullptr, "dh_comm_destroy: NULL ctx");
    if (ctx->comm) {
        Nccl* n = nccl();
        if (n) n->destroy((nccl_comm_t)ctx->comm);
        ctx->comm = nullptr;
    }
    return 0;
}

// recv[r * count .. (r+1) * count) = rank r's send[0 .. count)   (fp32, device pointers, caller's stream)
extern "C" int dh_allgather_f32(dh_ctx* ctx, const float* send, float* recv, int64_t count, void* stream) {
    DH_CHECK_ARG(ctx && send && recv, "dh_allgather_f32: NULL argument");
    DH_CHECK_ARG(count >= 0, "dh_allgather_f32: negative count");
    DH_CHECK_ARG(ctx->comm != nullptr, "dh_allgather_f32: no communicator (call dh_comm_init first)");
    Nccl* n = nccl();
    int rc = n->allgather(send, recv, (size_t)count, /*ncclFloat32*/ 7, (nccl_comm_t)ctx->comm, (cudaStream_t)stream);
    if (rc) return nccl_fail("ncclAllGather", rc);
    return 0;
}

extern "C" int dh_comm_info(dh_ctx* ctx, int* rank, int* world, int* nccl_version) {
    DH_CHECK_ARG(ctx != nullptr, "dh_comm_info: NULL ctx");
    if (rank) *rank = ctx->comm ? ctx->comm_rank : -1;
    if (world) *world = ctx->comm ? ctx->comm_world : 0;
    if (nccl_version) {
        *nccl_version = 0;
        Nccl* n = nccl();
        if (n && n->version) n->version(nccl_version);
    }
    return 0;
}
