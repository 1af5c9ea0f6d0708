// hb_dist.cu — one-process-per-GPU plumbing for the row-sharded pipelines.
//
// The reference has no distributed layer at all (SURVEY.md §2b "Collectives: none"); row-sharding with
// halo exchange is new work.  Ranks are laid out top to bottom over the frame's rows.
// Two transports:
//   * peer memory (default): producers store boundary rows straight into the neighbours' CUDA-IPC-mapped slabs and
//     handshake through release/acquire flags (PeerIO in ll_kernels.cuh); this file holds the two helper kernels that
//     are not fused into a pipeline kernel — peer_exchange_kernel (the caller-owned input rows) and peer_gather_kernel
//     (all-to-all gather of one coarse pyramid level) — both bounded-spin so a protocol bug reports instead of hanging;
//   * NCCL groups (HALIDE_B200_HALO=nccl): one ncclGroup{send up, recv up, send down, recv down} per level on the
//     compute stream.
// NCCL is also the bootstrap for the peer path (all-gather of the IPC handles and slab layouts).
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

namespace {

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) peer_exchange_kernel(PeerXchg x) {
    // 1. push: grid-stride copy of every segment into the neighbour's memory (16-byte or 2-byte elements)
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = 0; i < x.nseg; i++) {
        const PeerSeg &g = x.seg[i];
        if (g.elem == 16) {
            const uint4 *s = reinterpret_cast<const uint4 *>(g.src);
            uint4 *d = reinterpret_cast<uint4 *>(g.dst);
            for (unsigned k = tid; k < g.bytes / 16; k += nth) d[k] = s[k];
        } else {
            const unsigned short *s = reinterpret_cast<const unsigned short *>(g.src);
            unsigned short *d = reinterpret_cast<unsigned short *>(g.dst);
            for (unsigned k = tid; k < g.bytes / 2; k += nth) d[k] = s[k];
        }
    }
    // 2. announce: the last block to finish its slice releases the flags in the neighbours' memory
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned prev = atomicAdd(x.done_counter, 1u);
        if (prev == gridDim.x - 1) {
            *x.done_counter = 0u;
            __threadfence_system();
            if (x.peer_flag[0]) st_release_sys(x.peer_flag[0], x.epoch);
            if (x.peer_flag[1]) st_release_sys(x.peer_flag[1], x.epoch);
            for (int i = 0; i < x.nready; i++) st_release_sys(x.ready_flag[i], x.epoch);
        }
    }
    // 3. wait for the neighbours' rows of this step (block 0 only; bounded spin so a protocol bug cannot hang the GPU)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int d = 0; d < 2; d++) {
            if (!x.my_flag[d]) continue;
            long long t0 = clock64();
            while ((int)(ld_acquire_sys(x.my_flag[d]) - x.epoch) < 0) {  // epochs only grow
                if (clock64() - t0 > 4000000000LL) {  // ~2 s
                    *x.error_flag = 1u;
                    break;
                }
                __nanosleep(100);
            }
        }
        __threadfence_system();
    }
}

__device__ __forceinline__ void spin_until(const unsigned *flag, unsigned epoch, unsigned *error_flag) {
    long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {  // epochs only grow
        if (clock64() - t0 > 4000000000LL) {            // ~2 s: report instead of hanging the GPU
            *error_flag = 1u;
            break;
        }
        __nanosleep(100);
    }
}

__global__ void __launch_bounds__(256) peer_gather_kernel(PeerGather x) {
    // 0. no peer's copy may be overwritten before that peer has drained its previous call
    if ((int)threadIdx.x < x.npeer) spin_until(x.ready[threadIdx.x], x.epoch, x.error_flag);
    __syncthreads();
    // 1. push my rows into every peer's copy (16-byte elements; rows are 16-byte multiples)
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = 0; i < x.nseg; i++) {
        const PeerSeg &g = x.seg[i];
        const uint4 *s = reinterpret_cast<const uint4 *>(g.src);
        uint4 *d = reinterpret_cast<uint4 *>(g.dst);
        for (unsigned k = tid; k < g.bytes / 16; k += nth) d[k] = s[k];
    }
    // 2. the last block to finish announces the rows to every peer
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        unsigned prev = atomicAdd(x.done_counter, 1u);
        s_last = prev == gridDim.x - 1;
        if (s_last) {
            *x.done_counter = 0u;
            __threadfence_system();
        }
    }
    __syncthreads();
    if (s_last) {
        if ((int)threadIdx.x < x.npeer) st_release_sys(x.peer_flag[threadIdx.x], x.epoch);
        // 3. ... and waits for theirs; the kernel boundary then orders the gathered rows before the consumers
        if ((int)threadIdx.x < x.npeer) spin_until(x.my_flag[threadIdx.x], x.epoch, x.error_flag);
        __threadfence_system();
    }
}

}  // namespace

void launch_peer_exchange(const PeerXchg &x, cudaStream_t s) {
    HB_LAUNCH("peer_exchange", peer_exchange_kernel, 8, 256, 0, s, x);
}
void launch_peer_gather(const PeerGather &g, cudaStream_t s) {
    HB_LAUNCH("peer_gather", peer_gather_kernel, 64, 256, 0, s, g);
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
