// lpi_comm_*: the two exchange steps of the multi-GPU path behind the C ABI, so that a binder which does not
// go through torch.distributed can still drive it: ONE all-gather of the per-shard top-k candidates (gallery-sharded Recall@K) or of the
// [b, 2E] feature rows (data-parallel training), and ONE all-reduce of the flat 5 284-float prompt gradient.
// Reference anchor: gather_features, retrieval/methods/sprompt.py:38-82 (torch.distributed / horovod there).
// NCCL is resolved at run time (dlopen) -- the library links against nothing but the CUDA runtime: first a copy already loaded into the
// process (torch ships its own libnccl.so.2), then $LPI_NCCL_LIB, then the system libnccl.so.2.
#include "lpi_internal.h"
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    bool ok;
};
NcclApi g_nccl{};

int load_nccl() {
    if (g_nccl.ok) return LPI_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy the host process already uses, if any
    if (!h) {
        const char* env = getenv("LPI_NCCL_LIB");
        if (env && env[0]) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return lpi::set_error(LPI_ERR_UNSUPPORTED, "lpi_comm: libnccl.so.2 not found (%s); set LPI_NCCL_LIB", dlerror());
#define LPI_SYM(field, name)                                                                                   \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));                                  \
    if (!g_nccl.field) return lpi::set_error(LPI_ERR_UNSUPPORTED, "lpi_comm: symbol %s missing in libnccl", name);
    LPI_SYM(GetUniqueId, "ncclGetUniqueId")
    LPI_SYM(CommInitRank, "ncclCommInitRank")
    LPI_SYM(CommDestroy, "ncclCommDestroy")
    LPI_SYM(AllGather, "ncclAllGather")
    LPI_SYM(AllReduce, "ncclAllReduce")
    LPI_SYM(GetErrorString, "ncclGetErrorString")
#undef LPI_SYM
    g_nccl.ok = true;
    return LPI_OK;
}

int nccl_check(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return LPI_OK;
    return lpi::set_error(LPI_ERR_CUDA, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
}

}  // namespace

extern "C" int lpi_comm_unique_id_bytes(void) { return int(sizeof(ncclUniqueId)); }

// rank 0 creates the id and ships its bytes to the other ranks by whatever channel the host has (file, socket, MPI, a torch store)
extern "C" int lpi_comm_unique_id(void* id_out) {
    if (!id_out) return lpi::set_error(LPI_ERR_ARG, "lpi_comm_unique_id: null buffer");
    if (int rc = load_nccl()) return rc;
    ncclUniqueId id;
    if (int rc = nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId")) return rc;
    memcpy(id_out, &id, sizeof(id));
    return LPI_OK;
}

// one communicator per process and GPU (the current CUDA device); collective call on all ranks
extern "C" int lpi_comm_init(void** comm_out, int n_ranks, int rank, const void* unique_id) {
    if (!comm_out || !unique_id || n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return lpi::set_error(LPI_ERR_ARG, "lpi_comm_init: bad arguments (n_ranks=%d rank=%d)", n_ranks, rank);
    if (int rc = load_nccl()) return rc;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclComm_t c = nullptr;
    if (int rc = nccl_check(g_nccl.CommInitRank(&c, n_ranks, id, rank), "ncclCommInitRank")) return rc;
    *comm_out = c;
    return LPI_OK;
}

// recv [n_ranks * bytes_per_rank] <- every rank's send [bytes_per_rank], rank-major (the candidate / feature exchange)
extern "C" int lpi_comm_allgather(void* comm, const void* send, void* recv, long long bytes_per_rank, void* stream) {
    if (!comm || !send || !recv || bytes_per_rank < 0) return lpi::set_error(LPI_ERR_ARG, "lpi_comm_allgather: bad arguments");
    if (!g_nccl.ok) return lpi::set_error(LPI_ERR_ARG, "lpi_comm_allgather: no communicator was created by this library");
    return nccl_check(g_nccl.AllGather(send, recv, size_t(bytes_per_rank), ncclChar, static_cast<ncclComm_t>(comm), static_cast<cudaStream_t>(stream)),
                      "ncclAllGather");
}

// recv[i] = sum over ranks of send[i] (the flat prompt gradient); in place when send == recv
extern "C" int lpi_comm_allreduce_sum_f32(void* comm, const float* send, float* recv, long long n, void* stream) {
    if (!comm || !send || !recv || n < 0) return lpi::set_error(LPI_ERR_ARG, "lpi_comm_allreduce: bad arguments");
    if (!g_nccl.ok) return lpi::set_error(LPI_ERR_ARG, "lpi_comm_allreduce: no communicator was created by this library");
    return nccl_check(g_nccl.AllReduce(send, recv, size_t(n), ncclFloat, ncclSum, static_cast<ncclComm_t>(comm), static_cast<cudaStream_t>(stream)),
                      "ncclAllReduce");
}

extern "C" int lpi_comm_destroy(void* comm) {
    if (!comm) return LPI_OK;
    if (!g_nccl.ok) return lpi::set_error(LPI_ERR_ARG, "lpi_comm_destroy: no communicator was created by this library");
    return nccl_check(g_nccl.CommDestroy(static_cast<ncclComm_t>(comm)), "ncclCommDestroy");
}
