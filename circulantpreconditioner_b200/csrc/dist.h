// dist.h -- multi-rank plumbing: NCCL (dlopen'ed, so single-GPU use never needs it) all-to-all over NVLink.
// The reference's counterpart is fftw-mpi's MPI_Alltoall inside PETSc MATFFTW when MatCreateFFT is given
// PETSC_COMM_WORLD with more than one rank (reference src/PCSHELLFft_3D.cxx:35).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace cpc {

#define CPC_DIST_MAX_PEERS 8

struct DistState {
    int nranks = 1, rank = 0;
    void *comm = nullptr;        // ncclComm_t
    float *barrier_buf = nullptr;
    // flag barrier over peer-mapped memory (dist_flag_barrier_init): flags[g][s] = last epoch rank s signalled to this
    // rank in barrier group g.  Two groups: barriers issued on two different streams must not share epochs.
    unsigned long long *flags = nullptr;
    void *peer_flags[CPC_DIST_MAX_PEERS] = {};
    unsigned long long epoch[2] = { 0, 0 };
    int *timeout_flag = nullptr; // device int, set by a barrier kernel that gave up waiting (a peer died)
    int *sync_counters = nullptr; // 2 device ints: finished-block counters of kernels that signal at their end (zsolve.cuh)
    bool flag_barrier = false;
};

// Loads libnccl.so.2 on first use.  Returns CPC_OK or CPC_ERR_NCCL (message via cpc_last_error()).
int dist_unique_id(void *out128);
int dist_init(DistState &d, int nranks, int rank, const void *unique_id128, int device);
void dist_destroy(DistState &d);
// Rank r sends bytes [q*chunk_bytes, (q+1)*chunk_bytes) of `send` to rank q and receives rank s's chunk into
// recv + s*chunk_bytes (grouped ncclSend/ncclRecv: NCCL 2.27 has no all-to-all entry point).
int dist_alltoall(DistState &d, const void *send, void *recv, size_t chunk_bytes, cudaStream_t stream);
// The same among a subset of the ranks (the row or column group of a pencil grid, pencil.h), on the world communicator:
// chunk q of `send` goes to rank peers[q], chunk q of `recv` comes from rank peers[q]; peers lists the group in the same
// order on every member and contains this rank.  Groups are disjoint, so all groups exchange at once.
int dist_alltoall_group(DistState &d, const void *send, void *recv, size_t chunk_bytes, const int *peers, int npeers,
                        cudaStream_t stream);
// ncclAllGather of `bytes` per rank (the z-slab carry exchange of the recurrence middle pass, zsolve.cuh).
int dist_allgather(DistState &d, const void *send, void *recv, size_t bytes, cudaStream_t stream);
// In-place sum over ranks of `count` floats (agreement flags of collective set-up steps).
int dist_allreduce_sum_f32(DistState &d, float *buf, size_t count, cudaStream_t stream);
// Stream-ordered barrier across ranks (peer flags, or a 1-element all-reduce).  Every rank must issue the barriers of
// one group in the same order; group 1 is for a second stream (NCCL fallback: group is ignored, so callers that use
// two streams must check d.flag_barrier first).
int dist_barrier(DistState &d, cudaStream_t stream, int group = 0);
// Barrier through peer-mapped flags instead of an NCCL all-reduce (a 1-CTA kernel: every rank writes its epoch into
// every peer's flag array over NVLink and waits for the peers' epochs in its own; ~5 us instead of ~20 at 8 GPUs).
// The wait gives up after ~2 s and raises timeout_flag, so a dead peer cannot hang the GPU.  dist_barrier() uses it
// once dist_flag_barrier_init() has succeeded on every rank (collective); otherwise it stays on NCCL.
int dist_flag_barrier_init(DistState &d, int device, cudaStream_t stream);
// Exchange CUDA IPC handles of `local` (a cudaMalloc'ed buffer) through an NCCL all-gather and map every peer's
// buffer: peers[q] = address of rank q's buffer in this process (peers[rank] = local).  CPC_ERR_UNSUPPORTED when a
// peer cannot be mapped (the caller then keeps the NCCL all-to-all path).
int dist_map_peers(DistState &d, void *local, void **peers, int device, cudaStream_t stream);
void dist_unmap_peers(DistState &d, void **peers);

}  // namespace cpc
