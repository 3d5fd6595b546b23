// sn_comm.cu -- see sn_comm.h
#include "sn_comm.h"
#include <nccl.h>          // types and prototypes only: the library is bound with dlopen/dlsym below
#include <dlfcn.h>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <vector>

namespace snc {
namespace {

// ---- NCCL, bound at run time -----------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    std::string err;
};
NcclApi* nccl_api()
{
    static NcclApi api; static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
        if (!api.lib) { api.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define SN_SYM(f) api.f = reinterpret_cast<decltype(api.f)>(dlsym(api.lib, "nccl" #f)); if (!api.f) { api.err = "libnccl lacks nccl" #f; return; }
        SN_SYM(GetUniqueId) SN_SYM(CommInitRank) SN_SYM(CommDestroy) SN_SYM(GetErrorString) SN_SYM(GroupStart) SN_SYM(GroupEnd)
        SN_SYM(Send) SN_SYM(Recv) SN_SYM(AllReduce) SN_SYM(AllGather)
#undef SN_SYM
    });
    return &api;
}

class NcclComm : public Comm {
public:
    ncclComm_t comm = nullptr;
    NcclApi* a = nullptr;
    ~NcclComm() override { if (comm) a->CommDestroy(comm); }
    const char* kind() const override { return "nccl"; }
    int ck(ncclResult_t r, const char* what) { if (r == ncclSuccess) return 0; err = std::string(what) + ": " + a->GetErrorString(r); return -1; }
    int alltoallv(const void* send, const size_t* send_bytes, const size_t* send_off, void* recv, const size_t* recv_bytes, const size_t* recv_off, cudaStream_t st) override
    {
        if (ck(a->GroupStart(), "ncclGroupStart")) return -1;
        for (int p = 0; p < n; ++p) {
            if (send_bytes[p] && ck(a->Send((const char*)send + send_off[p], send_bytes[p], ncclUint8, p, comm, st), "ncclSend")) return -1;
            if (recv_bytes[p] && ck(a->Recv((char*)recv + recv_off[p], recv_bytes[p], ncclUint8, p, comm, st), "ncclRecv")) return -1;
        }
        return ck(a->GroupEnd(), "ncclGroupEnd");
    }
    int allgatherv(const void* send, void* recv, const size_t* bytes, const size_t* off, cudaStream_t st) override
    {
        bool even = true;
        for (int p = 0; p < n; ++p) even = even && bytes[p] == bytes[0] && off[p] == (size_t)p * bytes[0];
        if (even) { if (!bytes[0]) return 0; return ck(a->AllGather(send, recv, bytes[0], ncclUint8, comm, st), "ncclAllGather"); }
        // uneven slices: one grouped send/recv round (every rank sends its slice to every peer)
        if (ck(a->GroupStart(), "ncclGroupStart")) return -1;
        for (int p = 0; p < n; ++p) {
            if (p != rank && bytes[rank] && ck(a->Send(send, bytes[rank], ncclUint8, p, comm, st), "ncclSend")) return -1;
            if (p != rank && bytes[p] && ck(a->Recv((char*)recv + off[p], bytes[p], ncclUint8, p, comm, st), "ncclRecv")) return -1;
        }
        if (ck(a->GroupEnd(), "ncclGroupEnd")) return -1;
        if (bytes[rank] && (const char*)send != (const char*)recv + off[rank])
            if (cudaMemcpyAsync((char*)recv + off[rank], send, bytes[rank], cudaMemcpyDeviceToDevice, st) != cudaSuccess) { err = "allgatherv: local copy failed"; return -1; }
        return 0;
    }
    int allreduce_sum(void* buf, size_t count, int elem_size, cudaStream_t st) override
    {
        if (!count) return 0;
        const ncclDataType_t t = elem_size == 1 ? ncclUint8 : (elem_size == 4 ? ncclUint32 : ncclUint64);
        return ck(a->AllReduce(buf, buf, count, t, ncclSum, comm, st), "ncclAllReduce");
    }
};

// ---- in-process ranks ------------------------------------------------------------------------------------
template <class T> __global__ void k_sum_ranks(T* __restrict__ out, const T* const* __restrict__ in, int n, size_t count)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        T s = 0;
        for (int r = 0; r < n; ++r) s += in[r][i];
        out[i] = s;
    }
}
}  // namespace

struct LocalGroup {
    int n = 0;
    // a barrier that can be broken: a rank that fails aborts the group instead of leaving the others waiting
    std::mutex m; std::condition_variable cv; int waiting = 0; unsigned generation = 0; bool aborted = false;
    struct Slot { const void* send = nullptr; const size_t* off = nullptr; const size_t* bytes = nullptr; void* buf = nullptr; };
    std::vector<Slot> slot;
    bool barrier()
    {
        std::unique_lock<std::mutex> lk(m);
        if (aborted) return false;
        const unsigned gen = generation;
        if (++waiting == n) { waiting = 0; ++generation; cv.notify_all(); return true; }
        cv.wait(lk, [&] { return generation != gen || aborted; });
        return !aborted;
    }
    void abort() { std::lock_guard<std::mutex> lk(m); aborted = true; cv.notify_all(); }
};
LocalGroup* local_group_create(int n)
{
    if (n < 1) return nullptr;
    LocalGroup* g = new LocalGroup();
    g->n = n; g->slot.resize(n);
    return g;
}
void local_group_destroy(LocalGroup* g) { delete g; }
void local_group_abort(LocalGroup* g) { if (g) g->abort(); }

namespace {
class LocalComm : public Comm {
public:
    LocalGroup* g = nullptr;
    const char* kind() const override { return "local"; }
    bool sync_all(cudaStream_t st) { cudaStreamSynchronize(st); if (g->barrier()) return true; err = "another rank of the local group failed"; return false; }
    int alltoallv(const void* send, const size_t* send_bytes, const size_t* send_off, void* recv, const size_t* recv_bytes, const size_t* recv_off, cudaStream_t st) override
    {
        g->slot[rank].send = send; g->slot[rank].off = send_off; g->slot[rank].bytes = send_bytes;
        if (!sync_all(st)) return -1;                       // everybody's send buffer is complete and published
        int rc = 0;
        for (int s = 0; s < n; ++s) {
            const LocalGroup::Slot& q = g->slot[s];
            if (q.bytes[rank] != recv_bytes[s]) { err = "alltoallv: the sizes of sender and receiver disagree"; rc = -1; continue; }
            if (recv_bytes[s] && cudaMemcpyAsync((char*)recv + recv_off[s], (const char*)q.send + q.off[rank], recv_bytes[s], cudaMemcpyDeviceToDevice, st) != cudaSuccess) { err = "alltoallv: copy failed"; rc = -1; }
        }
        if (!sync_all(st)) return -1;                       // nobody reuses a send buffer before every peer has read it
        return rc;
    }
    int allgatherv(const void* send, void* recv, const size_t* bytes, const size_t* off, cudaStream_t st) override
    {
        g->slot[rank].send = send;
        if (!sync_all(st)) return -1;
        int rc = 0;
        for (int s = 0; s < n; ++s) {
            if (!bytes[s]) continue;
            const char* src = (const char*)g->slot[s].send; char* dst = (char*)recv + off[s];
            if (src != dst && cudaMemcpyAsync(dst, src, bytes[s], cudaMemcpyDeviceToDevice, st) != cudaSuccess) { err = "allgatherv: copy failed"; rc = -1; }
        }
        if (!sync_all(st)) return -1;
        return rc;
    }
    int allreduce_sum(void* buf, size_t count, int elem_size, cudaStream_t st) override
    {
        g->slot[rank].buf = buf;
        if (!sync_all(st)) return -1;
        if (!count) return sync_all(st) ? 0 : -1;
        void* tmp = nullptr; const void** tab = nullptr;
        int rc = 0;
        if (cudaMalloc(&tmp, count * (size_t)elem_size) != cudaSuccess || cudaMalloc((void**)&tab, sizeof(void*) * n) != cudaSuccess) { err = "allreduce: out of memory"; rc = -1; }
        if (!rc) {
            std::vector<const void*> h(n);
            for (int s = 0; s < n; ++s) h[s] = g->slot[s].buf;
            cudaMemcpyAsync(tab, h.data(), sizeof(void*) * n, cudaMemcpyHostToDevice, st);
            const unsigned grid = (unsigned)std::min<size_t>((count + 255) / 256, 2048);
            if (elem_size == 1) k_sum_ranks<uint8_t><<<grid, 256, 0, st>>>((uint8_t*)tmp, (const uint8_t* const*)tab, n, count);
            else if (elem_size == 4) k_sum_ranks<uint32_t><<<grid, 256, 0, st>>>((uint32_t*)tmp, (const uint32_t* const*)tab, n, count);
            else k_sum_ranks<unsigned long long><<<grid, 256, 0, st>>>((unsigned long long*)tmp, (const unsigned long long* const*)tab, n, count);
            cudaStreamSynchronize(st);                       // (h must outlive the copy)
        }
        const bool alive = sync_all(st);                    // every rank has its sum in tmp: the inputs may now be overwritten
        if (!alive) rc = -1;
        if (!rc && cudaMemcpyAsync(buf, tmp, count * (size_t)elem_size, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { err = "allreduce: copy failed"; rc = -1; }
        cudaStreamSynchronize(st);
        if (tmp) cudaFree(tmp);
        if (tab) cudaFree(tab);
        return rc;
    }
};
}  // namespace

Comm* make_local_comm(LocalGroup* g, int rank)
{
    if (!g || rank < 0 || rank >= g->n) return nullptr;
    LocalComm* c = new LocalComm();
    c->g = g; c->rank = rank; c->n = g->n;
    return c;
}
bool nccl_unique_id(void* out128, std::string& err)
{
    NcclApi* a = nccl_api();
    if (!a->lib || !a->err.empty()) { err = a->err; return false; }
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclResult_t r = a->GetUniqueId(&id);
    if (r != ncclSuccess) { err = std::string("ncclGetUniqueId: ") + a->GetErrorString(r); return false; }
    memcpy(out128, &id, 128);
    return true;
}
Comm* make_nccl_comm(int rank, int n, const void* unique_id, std::string& err)
{
    NcclApi* a = nccl_api();
    if (!a->lib || !a->err.empty()) { err = a->err; return nullptr; }
    ncclUniqueId id; memcpy(&id, unique_id, 128);
    NcclComm* c = new NcclComm();
    c->a = a; c->rank = rank; c->n = n;
    ncclResult_t r = a->CommInitRank(&c->comm, n, id, rank);
    if (r != ncclSuccess) { err = std::string("ncclCommInitRank: ") + a->GetErrorString(r); c->comm = nullptr; delete c; return nullptr; }
    return c;
}

}  // namespace snc
