// sn_comm.h -- the collectives of the multi-GPU path, issued by the C++ host.
//
// Two implementations behind one interface:
//   * NcclComm  : NCCL over NVLink / NVSwitch (ncclSend/ncclRecv groups for the alltoallv and the uneven allgather,
//                 ncclAllReduce), one rank per GPU.  libnccl.so.2 is resolved at run time (dlopen): the library
//                 stays loadable on a box without NCCL, and inside a torch process it binds to the copy torch loaded.
//   * LocalComm : N ranks = N contexts of ONE process on ONE device (a thread per rank): barrier + device-to-device
//                 copies.  Test infrastructure for the rank logic (bucket ownership, exchanges, stitching) on a
//                 single-GPU box; never used for measurements.
// All sizes are bytes.  Every call is collective: all ranks of the communicator must make it, in the same order.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <string>

namespace snc {

class Comm {
public:
    virtual ~Comm() {}
    int rank = 0, n = 1;
    std::string err;
    // rank r's `send_bytes[d]` bytes at send + send_off[d] go to rank d, which receives them at recv + recv_off[r]
    virtual int alltoallv(const void* send, const size_t* send_bytes, const size_t* send_off,
                          void* recv, const size_t* recv_bytes, const size_t* recv_off, cudaStream_t st) = 0;
    // every rank's `bytes[r]` bytes land at recv + off[r] on all ranks (send may alias recv + off[rank])
    virtual int allgatherv(const void* send, void* recv, const size_t* bytes, const size_t* off, cudaStream_t st) = 0;
    // element-wise sum over the ranks, in place; elem_size 1 (u8), 4 (u32) or 8 (u64)
    virtual int allreduce_sum(void* buf, size_t count, int elem_size, cudaStream_t st) = 0;
    virtual const char* kind() const = 0;
};

// NCCL communicator of `n` ranks; `unique_id` = the 128 bytes of ncclGetUniqueId made by rank 0.
Comm* make_nccl_comm(int rank, int n, const void* unique_id, std::string& err);
bool nccl_unique_id(void* out128, std::string& err);

// in-process group of n ranks on one device
struct LocalGroup;
LocalGroup* local_group_create(int n);
void local_group_destroy(LocalGroup* g);
void local_group_abort(LocalGroup* g);      // a rank failed: every pending and later collective of the group returns an error
Comm* make_local_comm(LocalGroup* g, int rank);

}  // namespace snc
