// Gradient all-reduce of the data-parallel path as C-ABI entry points (allreduce_flat taking an
// ncclComm_t).  The reference is single-process (scripts/train-nets:159-164); here the batch is sharded over
// one process per GPU and the ONE collective of a step is a sum over the flat buffer [gradients | TALR moments].
//
// NCCL is resolved at run time (dlopen / dlsym): the process that hosts this library has normally loaded a
// libnccl already (torch.distributed's), and two copies of NCCL in one process must be avoided; a host without
// any NCCL can still load libmpnn_sm100.so and use every other entry point.  MPNN_NCCL_LIB names a library to
// load when none is resident.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "../../include/mpnn.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;            // NCCL_UNIQUE_ID_BYTES
enum { ncclSuccess = 0, ncclFloat32 = 7, ncclSum = 0 };         // nccl.h: ncclDataType_t / ncclRedOp_t

struct Api {
    int (*GetUniqueId)(ncclUniqueId*);
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(int);
    int (*GetVersion)(int*);
    bool ok;
};

Api* api() {
    static Api a = {};
    static bool tried = false;
    if (tried) return a.ok ? &a : nullptr;
    tried = true;
    void* h = nullptr;
    if (dlsym(RTLD_DEFAULT, "ncclAllReduce")) h = RTLD_DEFAULT;          // already resident (torch's copy)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h && getenv("MPNN_NCCL_LIB")) h = dlopen(getenv("MPNN_NCCL_LIB"), RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
    *(void**)&a.GetUniqueId = dlsym(h, "ncclGetUniqueId");
    *(void**)&a.CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void**)&a.CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void**)&a.AllReduce = dlsym(h, "ncclAllReduce");
    *(void**)&a.GetErrorString = dlsym(h, "ncclGetErrorString");
    *(void**)&a.GetVersion = dlsym(h, "ncclGetVersion");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString;
    return a.ok ? &a : nullptr;
}

int fail(Api* a, const char* what, int rc) {
    mpnn_set_error("%s: %s", what, a->GetErrorString(rc));
    return MPNN_ERR_CUDA;
}

}  // namespace

extern "C" int mpnn_nccl_version(void) {
    Api* a = api();
    int v = 0;
    if (!a || !a->GetVersion || a->GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

extern "C" int mpnn_comm_unique_id(void* id128) {
    Api* a = api();
    MPNN_REQUIRE(a, "comm_unique_id: no NCCL library in this process (set MPNN_NCCL_LIB)");
    MPNN_REQUIRE(id128, "comm_unique_id: id128 is NULL");
    ncclUniqueId id;
    const int rc = a->GetUniqueId(&id);
    if (rc != ncclSuccess) return fail(a, "ncclGetUniqueId", rc);
    memcpy(id128, &id, sizeof(id));
    return MPNN_OK;
}

extern "C" int mpnn_comm_init_rank(void** comm, int world, int rank, const void* id128) {
    Api* a = api();
    MPNN_REQUIRE(a, "comm_init_rank: no NCCL library in this process (set MPNN_NCCL_LIB)");
    MPNN_REQUIRE(comm && id128 && world >= 1 && rank >= 0 && rank < world, "comm_init_rank: world=%d rank=%d", world, rank);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    const int rc = a->CommInitRank(&c, world, id, rank);        // collective over all ranks; uses the current CUDA device
    if (rc != ncclSuccess) return fail(a, "ncclCommInitRank", rc);
    *comm = c;
    return MPNN_OK;
}

extern "C" int mpnn_comm_destroy(void* comm) {
    Api* a = api();
    MPNN_REQUIRE(a && comm, "comm_destroy: args");
    const int rc = a->CommDestroy((ncclComm_t)comm);
    return rc == ncclSuccess ? MPNN_OK : fail(a, "ncclCommDestroy", rc);
}

extern "C" int mpnn_allreduce_flat(void* comm, float* buf, long long n, void* stream) {
    Api* a = api();
    MPNN_REQUIRE(a && comm && buf && n >= 0, "allreduce_flat: args");
    const int rc = a->AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, (ncclComm_t)comm, (cudaStream_t)stream);
    return rc == ncclSuccess ? MPNN_OK : fail(a, "ncclAllReduce", rc);
}
