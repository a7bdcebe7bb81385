// dist.cu -- NCCL bootstrap and all-to-all (see dist.h).
#include "dist.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "plan.h"

namespace cpc {

namespace {
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void *ncclComm_tt;
typedef int ncclResult_tt;
enum { NCCL_UINT8 = 1, NCCL_FLOAT32 = 7, NCCL_SUM = 0 };

struct Api {
    void *handle = nullptr;
    ncclResult_tt (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
    ncclResult_tt (*CommInitRank)(ncclComm_tt *, int, ncclUniqueId_t, int) = nullptr;
    ncclResult_tt (*CommDestroy)(ncclComm_tt) = nullptr;
    ncclResult_tt (*GroupStart)() = nullptr;
    ncclResult_tt (*GroupEnd)() = nullptr;
    ncclResult_tt (*Send)(const void *, size_t, int, int, ncclComm_tt, cudaStream_t) = nullptr;
    ncclResult_tt (*Recv)(void *, size_t, int, int, ncclComm_tt, cudaStream_t) = nullptr;
    ncclResult_tt (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_tt, cudaStream_t) = nullptr;
    ncclResult_tt (*AllGather)(const void *, void *, size_t, int, ncclComm_tt, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_tt) = nullptr;
    bool ok = false;
};

Api g_api;
std::once_flag g_once;

void load_api()
{
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char *nm : names) {
        g_api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_api.handle) break;
    }
    if (!g_api.handle) return;
#define LOAD(field, sym)                                                   \
    *(void **)(&g_api.field) = dlsym(g_api.handle, sym);                   \
    if (!g_api.field) return;
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(AllGather, "ncclAllGather")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    g_api.ok = true;
}

int need_api()
{
    std::call_once(g_once, load_api);
    if (!g_api.ok) {
        set_error("NCCL not available: could not dlopen libnccl.so.2 (%s)", dlerror() ? dlerror() : "missing symbol");
        return CPC_ERR_NCCL;
    }
    return CPC_OK;
}

int nccl_fail(ncclResult_tt r, const char *what)
{
    set_error("NCCL error in %s: %s", what, g_api.GetErrorString ? g_api.GetErrorString(r) : "?");
    return CPC_ERR_NCCL;
}
#define CPC_NCCL(call)                                   \
    do {                                                 \
        ncclResult_tt _r = (call);                       \
        if (_r != 0) return nccl_fail(_r, #call);        \
    } while (0)
}  // namespace

int dist_unique_id(void *out128)
{
    int rc = need_api();
    if (rc) return rc;
    ncclUniqueId_t id;
    CPC_NCCL(g_api.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return CPC_OK;
}

int dist_init(DistState &d, int nranks, int rank, const void *unique_id128, int device)
{
    d.nranks = nranks;
    d.rank = rank;
    if (nranks == 1) return CPC_OK;
    if (!unique_id128) { set_error("multi-rank plan needs nccl_unique_id"); return CPC_ERR_ARG; }
    int rc = need_api();
    if (rc) return rc;
    CPC_CUDA(cudaSetDevice(device));
    ncclUniqueId_t id;
    memcpy(&id, unique_id128, sizeof(id));
    ncclComm_tt comm = nullptr;
    CPC_NCCL(g_api.CommInitRank(&comm, nranks, id, rank));
    d.comm = comm;
    return CPC_OK;
}

void dist_destroy(DistState &d)
{
    if (d.flag_barrier) {
        dist_unmap_peers(d, d.peer_flags);
        d.flag_barrier = false;
    }
    if (d.flags) cudaFree(d.flags);
    if (d.timeout_flag) cudaFree(d.timeout_flag);
    if (d.sync_counters) cudaFree(d.sync_counters);
    d.sync_counters = nullptr;
    d.flags = nullptr;
    d.timeout_flag = nullptr;
    if (d.barrier_buf) cudaFree(d.barrier_buf);
    d.barrier_buf = nullptr;
    if (d.comm && g_api.ok) g_api.CommDestroy((ncclComm_tt)d.comm);
    d.comm = nullptr;
}

int dist_map_peers(DistState &d, void *local, void **peers, int device, cudaStream_t stream)
{
    for (int q = 0; q < d.nranks; ++q) peers[q] = nullptr;
    peers[d.rank] = local;
    if (d.nranks == 1) return CPC_OK;
    CPC_CUDA(cudaSetDevice(device));
    // A local failure before the collectives must not leave the other ranks waiting in them: it is carried as
    // rc through the all-gather and reported by the agreement flag below.
    int rc = CPC_OK;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    {
        cudaError_t e = cudaIpcGetMemHandle(&mine, local);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
            rc = CPC_ERR_UNSUPPORTED;
        }
    }
    const size_t hs = sizeof(cudaIpcMemHandle_t);
    char *dev = nullptr;
    CPC_CUDA(cudaMalloc(&dev, hs * (size_t)(d.nranks + 1)));
    CPC_CUDA(cudaMemcpyAsync(dev + hs * d.nranks, &mine, hs, cudaMemcpyHostToDevice, stream));
    CPC_NCCL(g_api.AllGather(dev + hs * d.nranks, dev, hs, NCCL_UINT8, (ncclComm_tt)d.comm, stream));
    std::vector<cudaIpcMemHandle_t> all(d.nranks);
    CPC_CUDA(cudaMemcpyAsync(all.data(), dev, hs * d.nranks, cudaMemcpyDeviceToHost, stream));
    CPC_CUDA(cudaStreamSynchronize(stream));
    cudaFree(dev);
    for (int q = 0; q < d.nranks && rc == CPC_OK; ++q) {
        if (q == d.rank) continue;
        cudaError_t e = cudaIpcOpenMemHandle(&peers[q], all[q], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", q, cudaGetErrorString(e));
            peers[q] = nullptr;
            rc = CPC_ERR_UNSUPPORTED;
        }
    }
    // every rank must agree on the outcome, otherwise one would push while another waits for NCCL
    float flag = rc == CPC_OK ? 0.f : 1.f, *dflag = nullptr;
    CPC_CUDA(cudaMalloc(&dflag, sizeof(float)));
    CPC_CUDA(cudaMemcpyAsync(dflag, &flag, sizeof(float), cudaMemcpyHostToDevice, stream));
    CPC_NCCL(g_api.AllReduce(dflag, dflag, 1, NCCL_FLOAT32, NCCL_SUM, (ncclComm_tt)d.comm, stream));
    CPC_CUDA(cudaMemcpyAsync(&flag, dflag, sizeof(float), cudaMemcpyDeviceToHost, stream));
    CPC_CUDA(cudaStreamSynchronize(stream));
    cudaFree(dflag);
    if (flag != 0.f) {
        dist_unmap_peers(d, peers);
        if (rc == CPC_OK) set_error("a peer rank could not map IPC memory");
        return CPC_ERR_UNSUPPORTED;
    }
    return CPC_OK;
}

void dist_unmap_peers(DistState &d, void **peers)
{
    for (int q = 0; q < d.nranks; ++q) {
        if (q != d.rank && peers[q]) cudaIpcCloseMemHandle(peers[q]);
        peers[q] = nullptr;
    }
}

int dist_alltoall(DistState &d, const void *send, void *recv, size_t chunk_bytes, cudaStream_t stream)
{
    if (d.nranks == 1) {
        CPC_CUDA(cudaMemcpyAsync(recv, send, chunk_bytes, cudaMemcpyDeviceToDevice, stream));
        return CPC_OK;
    }
    CPC_NCCL(g_api.GroupStart());
    for (int p = 0; p < d.nranks; ++p) {
        CPC_NCCL(g_api.Send((const char *)send + (size_t)p * chunk_bytes, chunk_bytes, NCCL_UINT8, p,
                            (ncclComm_tt)d.comm, stream));
        CPC_NCCL(g_api.Recv((char *)recv + (size_t)p * chunk_bytes, chunk_bytes, NCCL_UINT8, p, (ncclComm_tt)d.comm,
                            stream));
    }
    CPC_NCCL(g_api.GroupEnd());
    return CPC_OK;
}

int dist_alltoall_group(DistState &d, const void *send, void *recv, size_t chunk_bytes, const int *peers, int npeers,
                        cudaStream_t stream)
{
    if (npeers == 1) {                     // a group of one: the "exchange" is a copy
        CPC_CUDA(cudaMemcpyAsync(recv, send, chunk_bytes, cudaMemcpyDeviceToDevice, stream));
        return CPC_OK;
    }
    if (!d.comm) { set_error("group all-to-all without an NCCL communicator"); return CPC_ERR_STATE; }
    CPC_NCCL(g_api.GroupStart());
    for (int q = 0; q < npeers; ++q) {
        CPC_NCCL(g_api.Send((const char *)send + (size_t)q * chunk_bytes, chunk_bytes, NCCL_UINT8, peers[q],
                            (ncclComm_tt)d.comm, stream));
        CPC_NCCL(g_api.Recv((char *)recv + (size_t)q * chunk_bytes, chunk_bytes, NCCL_UINT8, peers[q], (ncclComm_tt)d.comm,
                            stream));
    }
    CPC_NCCL(g_api.GroupEnd());
    return CPC_OK;
}

struct FlagPeers {
    unsigned long long *p[CPC_DIST_MAX_PEERS];
};

// Thread q: signal rank q (store this rank's epoch into slot [rank] of q's flag array, release at system scope so that
// everything this GPU wrote before -- the preceding kernels of the stream have completed -- is visible first), then
// wait for q's signal in the local array (acquire).  Epochs only grow, so a peer that is already one barrier ahead
// still satisfies the wait.
__global__ void flag_barrier_kernel(FlagPeers peers, unsigned long long *mine, int nranks, int rank,
                                    unsigned long long epoch, int *timeout_flag)
{
    const int q = threadIdx.x;
    if (q >= nranks) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.p[q] + rank), "l"(epoch) : "memory");
    const long long t0 = clock64();
    unsigned long long seen = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine + q) : "memory");
        if (seen >= epoch) break;
        if (clock64() - t0 > 4000000000ll) { *timeout_flag = 1; break; }       // ~2 s: a peer is gone
    }
    __threadfence_system();
}

int dist_flag_barrier_init(DistState &d, int device, cudaStream_t stream)
{
    if (d.nranks == 1 || d.nranks > CPC_DIST_MAX_PEERS) return CPC_OK;
    CPC_CUDA(cudaSetDevice(device));
    CPC_CUDA(cudaMalloc(&d.flags, sizeof(unsigned long long) * 2 * CPC_DIST_MAX_PEERS));
    CPC_CUDA(cudaMemset(d.flags, 0, sizeof(unsigned long long) * 2 * CPC_DIST_MAX_PEERS));
    CPC_CUDA(cudaMalloc(&d.timeout_flag, sizeof(int)));
    CPC_CUDA(cudaMemset(d.timeout_flag, 0, sizeof(int)));
    CPC_CUDA(cudaMalloc(&d.sync_counters, 2 * sizeof(int)));
    CPC_CUDA(cudaMemset(d.sync_counters, 0, 2 * sizeof(int)));
    int rc = dist_map_peers(d, d.flags, d.peer_flags, device, stream);
    if (rc == CPC_OK) d.flag_barrier = true;
    else if (rc != CPC_ERR_UNSUPPORTED) return rc;
    return CPC_OK;
}

int dist_barrier(DistState &d, cudaStream_t stream, int group)
{
    if (d.nranks == 1) return CPC_OK;
    if (d.flag_barrier) {
        FlagPeers fp{};
        const int go = (group ? 1 : 0) * CPC_DIST_MAX_PEERS;
        for (int q = 0; q < d.nranks; ++q) fp.p[q] = (unsigned long long *)d.peer_flags[q] + go;
        flag_barrier_kernel<<<1, 32, 0, stream>>>(fp, d.flags + go, d.nranks, d.rank, ++d.epoch[group ? 1 : 0],
                                                  d.timeout_flag);
        CPC_CUDA(cudaGetLastError());
        return CPC_OK;
    }
    if (!d.barrier_buf) {
        CPC_CUDA(cudaMalloc(&d.barrier_buf, sizeof(float)));
        CPC_CUDA(cudaMemsetAsync(d.barrier_buf, 0, sizeof(float), stream));
    }
    CPC_NCCL(g_api.AllReduce(d.barrier_buf, d.barrier_buf, 1, NCCL_FLOAT32, NCCL_SUM, (ncclComm_tt)d.comm, stream));
    return CPC_OK;
}

int dist_allreduce_sum_f32(DistState &d, float *buf, size_t count, cudaStream_t stream)
{
    if (d.nranks == 1) return CPC_OK;
    CPC_NCCL(g_api.AllReduce(buf, buf, count, NCCL_FLOAT32, NCCL_SUM, (ncclComm_tt)d.comm, stream));
    return CPC_OK;
}

// Every rank contributes `bytes` from send; recv holds the P contributions in rank order.
int dist_allgather(DistState &d, const void *send, void *recv, size_t bytes, cudaStream_t stream)
{
    if (d.nranks == 1) {
        if (send != recv) CPC_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, stream));
        return CPC_OK;
    }
    CPC_NCCL(g_api.AllGather(send, recv, bytes, NCCL_UINT8, (ncclComm_tt)d.comm, stream));
    return CPC_OK;
}

}  // namespace cpc
