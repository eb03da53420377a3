// comm.cu -- the one collective of the path behind the C ABI: the per-model sufficient statistics of a Baum-Welch iteration
// summed over the ranks (custom_hmm.py:417-419, :434-439 accumulate them in one process; with utterances sharded over GPUs
// the same block is all-reduced once per iteration).  NCCL is resolved at run time (dlopen of libnccl.so.2: the copy the
// host process already has, e.g. the one PyTorch loads, or any on the library path), so the library links against cudart
// only and loads on a box without NCCL; every entry point fails loudly when NCCL is missing.
#include <dlfcn.h>

#include "common.cuh"

namespace {
typedef struct ncclComm *ncclComm_t;
struct nccl_uid { char internal[128]; };
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(ncclComm_t *, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_destroy)(ncclComm_t);
typedef const char *(*fn_errstr)(int);
struct NcclApi {
    void *lib = nullptr;
    fn_get_uid get_uid = nullptr; fn_init_rank init_rank = nullptr; fn_allreduce allreduce = nullptr;
    fn_destroy destroy = nullptr; fn_errstr errstr = nullptr;
    bool ok = false;
};
NcclApi &nccl() {
    static NcclApi a;
    if (a.lib) return a;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return a;
    a.get_uid = (fn_get_uid)dlsym(a.lib, "ncclGetUniqueId");
    a.init_rank = (fn_init_rank)dlsym(a.lib, "ncclCommInitRank");
    a.allreduce = (fn_allreduce)dlsym(a.lib, "ncclAllReduce");
    a.destroy = (fn_destroy)dlsym(a.lib, "ncclCommDestroy");
    a.errstr = (fn_errstr)dlsym(a.lib, "ncclGetErrorString");
    a.ok = a.get_uid && a.init_rank && a.allreduce && a.destroy;
    return a;
}
constexpr int kNcclFloat64 = 8, kNcclSum = 0;     // ncclDataType_t / ncclRedOp_t values (nccl.h)
}   // namespace

struct sapr_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

static std::string nccl_err(int rc) {
    NcclApi &a = nccl();
    return std::string(a.errstr ? a.errstr(rc) : "NCCL error") + " (" + std::to_string(rc) + ")";
}

extern "C" int sapr_comm_unique_id(void *uid128) {
    if (!uid128) return SAPR_E_INVALID;
    NcclApi &a = nccl();
    if (!a.ok) return SAPR_E_CUDA;
    nccl_uid id;
    if (a.get_uid(&id) != 0) return SAPR_E_CUDA;
    memcpy(uid128, &id, sizeof(id));
    return SAPR_OK;
}

extern "C" int sapr_comm_init_rank(sapr_ctx *ctx, const void *uid128, int rank, int world, sapr_comm **out) {
    if (!ctx || !uid128 || !out || world < 1 || rank < 0 || rank >= world) return SAPR_E_INVALID;
    *out = nullptr;
    NcclApi &a = nccl();
    if (!a.ok) SAPR_FAIL(ctx, SAPR_E_CUDA, "comm_init_rank: libnccl.so.2 not found (dlopen)");
    SAPR_CUDA(ctx, cudaSetDevice(ctx->device));
    nccl_uid id;
    memcpy(&id, uid128, sizeof(id));
    sapr_comm *c = new sapr_comm();
    c->rank = rank; c->world = world;
    const int rc = a.init_rank(&c->comm, world, id, rank);
    if (rc != 0) { delete c; SAPR_FAIL(ctx, SAPR_E_CUDA, "ncclCommInitRank: " + nccl_err(rc)); }
    *out = c;
    return SAPR_OK;
}

// in place, float64 sum over the ranks, enqueued on the context's stream (no host synchronisation): E-step -> all-reduce ->
// M-step run back to back on the device
extern "C" int sapr_stats_allreduce(sapr_ctx *ctx, sapr_comm *comm, double *stats, int64_t n) {
    if (!ctx || !comm || !stats || n < 0) return SAPR_E_INVALID;
    if (comm->world == 1 || n == 0) return SAPR_OK;
    NcclApi &a = nccl();
    const int rc = a.allreduce(stats, stats, (size_t)n, kNcclFloat64, kNcclSum, comm->comm, ctx->stream);
    if (rc != 0) SAPR_FAIL(ctx, SAPR_E_CUDA, "ncclAllReduce: " + nccl_err(rc));
    ctx->launches++;
    return SAPR_OK;
}

extern "C" int sapr_comm_destroy(sapr_comm *comm) {
    if (!comm) return SAPR_OK;
    NcclApi &a = nccl();
    if (a.ok && comm->comm) a.destroy(comm->comm);
    delete comm;
    return SAPR_OK;
}
