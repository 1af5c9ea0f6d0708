// hb_dist.cu — one-process-per-GPU plumbing for the row-sharded local_laplacian.
//
// The reference has no distributed layer at all (SURVEY.md §2b "Collectives: none"); row-sharding is new work
// (SURVEY.md §8e).  Ranks are laid out top to bottom over the frame's rows.  A sharded call communicates twice
// (DESIGN.md §7): one exchange of input halo rows with the two row neighbours and one all-to-all gather of a coarse
// pyramid level — each ONE ncclGroup of ncclSend / ncclRecv on the compute stream (`exchange`).  The bands of all ranks
// are gathered once per geometry with `allgather_bytes`.
//
// NCCL is bound at run time (dlopen "libnccl.so.2") so the library has no link-time dependency and
// shares the NCCL instance torch already loaded when the host process is the Python bench/test.
// Bootstrap: rank 0 calls halide_b200_dist_unique_id, the 128-byte id travels over the host's own
// control plane (torch.distributed broadcast in halide_b200/dist.py), every rank calls
// halide_b200_dist_init.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "hb_common.h"
#include "hb_dist.h"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_size = 1;

int load_nccl() {
    if (g_nccl.handle) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return hb::fail(halide_error_code_generic_error, "dist: cannot load libnccl.so.2: %s", dlerror());
#define HB_SYM(field, name)                                                                             \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));                            \
    if (!g_nccl.field) return hb::fail(halide_error_code_generic_error, "dist: libnccl lacks %s", name);
    HB_SYM(GetUniqueId, "ncclGetUniqueId")
    HB_SYM(CommInitRank, "ncclCommInitRank")
    HB_SYM(CommDestroy, "ncclCommDestroy")
    HB_SYM(Send, "ncclSend")
    HB_SYM(Recv, "ncclRecv")
    HB_SYM(AllGather, "ncclAllGather")
    HB_SYM(GroupStart, "ncclGroupStart")
    HB_SYM(GroupEnd, "ncclGroupEnd")
    HB_SYM(GetErrorString, "ncclGetErrorString")
#undef HB_SYM
    g_nccl.handle = h;
    return 0;
}

int check(ncclResult_t r, const char *what) {
    if (r == ncclSuccess) return 0;
    return hb::fail(halide_error_code_generic_error, "dist: %s failed: %s", what,
                    g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
}

}  // namespace

namespace hbdist {

bool active() { return g_comm != nullptr; }  // a 1-rank communicator is legal (timing the sharded path without neighbours)
int rank() { return g_rank; }
int size() { return g_size; }

int exchange(const Msg *msgs, int n, cudaStream_t s) {
    if (!active()) return hb::fail(halide_error_code_generic_error, "dist: halo exchange without an initialised communicator");
    int r;
    if ((r = check(g_nccl.GroupStart(), "ncclGroupStart"))) return r;
    for (int i = 0; i < n; i++) {
        const Msg &m = msgs[i];
        if (m.bytes == 0) continue;
        if (m.send) r = check(g_nccl.Send(m.ptr, m.bytes, ncclInt8, m.peer, g_comm, s), "ncclSend");
        else r = check(g_nccl.Recv(m.ptr, m.bytes, ncclInt8, m.peer, g_comm, s), "ncclRecv");
        if (r) {
            g_nccl.GroupEnd();
            return r;
        }
    }
    return check(g_nccl.GroupEnd(), "ncclGroupEnd");
}

int allgather_bytes(const void *send, void *recv_all, size_t bytes) {
    if (!active()) return hb::fail(halide_error_code_generic_error, "dist: all-gather without an initialised communicator");
    void *dsend = nullptr, *drecv = nullptr;
    cudaStream_t s = hb::stream();
    if (cudaMalloc(&dsend, bytes) != cudaSuccess || cudaMalloc(&drecv, bytes * g_size) != cudaSuccess) {
        cudaGetLastError();
        return hb::fail(halide_error_code_device_malloc_failed, "dist: all-gather staging allocation failed");
    }
    cudaMemcpyAsync(dsend, send, bytes, cudaMemcpyHostToDevice, s);
    int r = check(g_nccl.AllGather(dsend, drecv, bytes, ncclInt8, g_comm, s), "ncclAllGather");
    if (!r) {
        cudaMemcpyAsync(recv_all, drecv, bytes * g_size, cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) r = hb::fail(halide_error_code_generic_error, "dist: all-gather failed");
    }
    cudaFree(dsend);
    cudaFree(drecv);
    return r;
}

}  // namespace hbdist

extern "C" {

int halide_b200_dist_unique_id(char *out128) {
    int r = load_nccl();
    if (r) return r;
    ncclUniqueId id;
    if ((r = check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId"))) return r;
    memcpy(out128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return 0;
}

int halide_b200_dist_init(int rank, int nranks, const char *id128) {
    int r = load_nccl();
    if (r) return r;
    if (g_comm) return hb::fail(halide_error_code_generic_error, "dist: already initialised");
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    if ((r = check(g_nccl.CommInitRank(&g_comm, nranks, id, rank), "ncclCommInitRank"))) return r;
    g_rank = rank;
    g_size = nranks;
    return 0;
}

int halide_b200_dist_shutdown(void) {
    if (g_comm) {
        cudaDeviceSynchronize();
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
    }
    g_rank = 0;
    g_size = 1;
    return 0;
}

int halide_b200_dist_rank(void) { return g_rank; }
int halide_b200_dist_size(void) { return g_size; }

}  // extern "C"
