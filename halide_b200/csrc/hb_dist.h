// hb_dist.h — halo-exchange primitive used by the row-sharded filters (implemented in hb_dist.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace hbdist {

struct Msg {
    void *ptr;     // device address
    size_t bytes;
    int peer;      // rank
    bool send;     // true: ncclSend, false: ncclRecv
};

bool active();
int rank();
int size();
// One ncclGroup of point-to-point messages enqueued on stream `s` (ordered with the kernels on it).
int exchange(const Msg *msgs, int n, cudaStream_t s);

// Host-side all-gather of a small POD blob (plan setup only, synchronous): recv_all holds size() blobs in rank order.
int allgather_bytes(const void *send, void *recv_all, size_t bytes_per_rank);

// ---- peer-memory halo exchange (NVLink stores + flags instead of ncclSend/ncclRecv) ------------------------
// One kernel per exchange step: every block copies a slice of this rank's boundary rows straight into the
// neighbours' halo rows (peer pointers obtained through CUDA IPC), the last block to finish releases a flag in the
// neighbour's memory, and one thread then spins (acquire, with a timeout) until both neighbours' flags for this
// step and epoch have arrived.  Stream order makes the halo rows visible to the next kernel.
struct PeerSeg {
    const void *src;  // local rows
    void *dst;        // same rows in the neighbour's buffer (peer address)
    unsigned bytes;
    unsigned elem;    // copy granularity in bytes: 16 or 2
};
struct PeerXchg {
    PeerSeg seg[12];
    int nseg;
    unsigned *peer_flag[2];      // where to announce completion: [0] up neighbour, [1] down neighbour (null if none)
    const unsigned *my_flag[2];  // where the neighbours announce theirs
    unsigned epoch;
    unsigned *done_counter;      // local scratch counter (self-resetting)
    unsigned *error_flag;        // set to 1 on wait timeout
    unsigned *ready_flag[8];     // "my previous call has drained" announcements to every other rank (see PeerGather)
    int nready;
};
void launch_peer_exchange(const PeerXchg &x, cudaStream_t s);

// ---- all-to-all gather of one (small) pyramid level --------------------------------------------------------
// Every rank stores its band of the level into the same rows of every other rank's copy, so that the levels
// below it can be computed redundantly on every rank with no further exchange.  Before overwriting a peer's copy
// the kernel waits for that peer's `ready` announcement of this epoch (made by the peer's first kernel of the call,
// i.e. after everything of its previous call has drained); afterwards it releases one flag per peer and waits for
// all of theirs.
struct PeerGather {
    PeerSeg seg[16];
    int nseg;
    int npeer;
    unsigned *peer_flag[8];       // my slot in each peer's gather-flag array
    const unsigned *my_flag[8];   // the peers' slots in mine
    const unsigned *ready[8];     // the peers' ready slots in mine
    unsigned epoch;
    unsigned *done_counter;
    unsigned *error_flag;
};
void launch_peer_gather(const PeerGather &g, cudaStream_t s);

}  // namespace hbdist
