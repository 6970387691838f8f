// comm.cu — the library's own NCCL communicator (SURVEY.md §8b "Library owns ... NCCL comm"; §8e).
//
// One process drives one GPU; a Rust (or any other) host has no torch.distributed, so the collectives of the two paths
// live here, behind the C ABI, and run on device buffers on the caller's stream with no host hop:
//   * BLS batch: ncclAllReduce(min) of the 1-byte device verdict      (lhb200_bls_batch_verify_collective)
//   * sharded state root: ncclAllGather of the per-rank subtree roots  (lhb200_state_root_sharded, merkle_host.cu)
// NCCL is loaded with dlopen at lhb200_comm_init (libnccl.so.2: the system library or the one a host process — e.g.
// PyTorch — has already mapped), so liblhb200.so has no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include "ctx.h"

namespace lhb200 {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_world = 1;
static std::mutex g_comm_mu;

static int32_t nccl_load() {
    if (g_nccl.handle) return LHB200_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { set_error("cannot load NCCL (libnccl.so.2): %s", dlerror()); return LHB200_ENODEV; }
#define LHB_SYM(field, name)                                                        \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));        \
    if (!g_nccl.field) { set_error("NCCL symbol %s missing", name); dlclose(h); return LHB200_ENODEV; }
    LHB_SYM(GetUniqueId, "ncclGetUniqueId");
    LHB_SYM(CommInitRank, "ncclCommInitRank");
    LHB_SYM(CommDestroy, "ncclCommDestroy");
    LHB_SYM(AllReduce, "ncclAllReduce");
    LHB_SYM(AllGather, "ncclAllGather");
    LHB_SYM(GetErrorString, "ncclGetErrorString");
#undef LHB_SYM
    g_nccl.handle = h;
    return LHB200_OK;
}
static int32_t nccl_fail(ncclResult_t r, const char* what) {
    set_error("%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
    return LHB200_ECUDA;
}

// ---- used by bls_host.cu / merkle_host.cu
bool comm_active() { return g_comm != nullptr && g_world > 1; }
int comm_world() { return g_world; }
int comm_rank() { return g_rank; }
int32_t comm_allreduce_min_u8(void* d_buf, size_t n, cudaStream_t s) {
    if (!comm_active()) return LHB200_OK;
    ncclResult_t r = g_nccl.AllReduce(d_buf, d_buf, n, ncclUint8, ncclMin, g_comm, s);
    return r == ncclSuccess ? LHB200_OK : nccl_fail(r, "ncclAllReduce(min)");
}
int32_t comm_allgather_bytes(const void* d_send, void* d_recv, size_t bytes_per_rank, cudaStream_t s) {
    if (!comm_active()) {
        if (d_send != d_recv) {
            cudaError_t e = cudaMemcpyAsync(d_recv, d_send, bytes_per_rank, cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) return cuda_fail(e, "allgather(1 rank) copy");
        }
        return LHB200_OK;
    }
    ncclResult_t r = g_nccl.AllGather(d_send, d_recv, bytes_per_rank, ncclUint8, g_comm, s);
    return r == ncclSuccess ? LHB200_OK : nccl_fail(r, "ncclAllGather");
}

}  // namespace lhb200

using namespace lhb200;

extern "C" {

// ncclGetUniqueId: rank 0 calls this and ships the 128 bytes to the other ranks by any means (file, socket, MPI, ...).
int32_t lhb200_comm_unique_id(uint8_t id[128]) {
    if (!id) return LHB200_EINVAL;
    int32_t rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId u;
    ncclResult_t r = g_nccl.GetUniqueId(&u);
    if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
    static_assert(sizeof(u) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, 128);
    return LHB200_OK;
}

// ncclCommInitRank on the device of lhb200_init (collective: every rank of the job calls it with the same id).
int32_t lhb200_comm_init(int32_t rank, int32_t world, const uint8_t id[128]) {
    LHB_REQUIRE_READY();
    if (!id || world < 1 || rank < 0 || rank >= world) return LHB200_EINVAL;
    std::lock_guard<std::mutex> g(g_comm_mu);
    if (g_comm) { set_error("communicator already initialised"); return LHB200_EINVAL; }
    int32_t rc = nccl_load();
    if (rc) return rc;
    LHB_CUDA(cudaSetDevice(ctx().device));
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclComm_t c = nullptr;
    ncclResult_t r = g_nccl.CommInitRank(&c, world, u, rank);
    if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
    g_comm = c;
    g_rank = rank;
    g_world = world;
    return LHB200_OK;
}

int32_t lhb200_comm_destroy(void) {
    std::lock_guard<std::mutex> g(g_comm_mu);
    if (g_comm) {
        if (ctx().ready) cudaDeviceSynchronize();
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
    }
    g_rank = 0;
    g_world = 1;
    return LHB200_OK;
}

int32_t lhb200_comm_info(int32_t* rank, int32_t* world) {
    if (rank) *rank = g_rank;
    if (world) *world = g_comm ? g_world : 1;
    return LHB200_OK;
}

}  // extern "C"
