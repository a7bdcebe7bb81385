// plan_impl.cuh -- PlanT<T>: set-up, pass scheduling and launches for one dtype (double or float).
//
// Reference behaviour being reproduced (all in /root/reference/src):
//   set-up     FftLinearSolver_3D.c:80-164, 218-249 and PCSHELLFft_3D.cxx:26-84 (once; tables stay in HBM)
//   apply      FftLinearSolver_3D.c:166-190 (solve_3D): forward DFT, divide by Diag, backward DFT, scale 1/size
//   tear-down  PCSHELLFft_3D.cxx:86-99
// Schedule of one apply on one GPU (5 HBM passes):
//   Fx : b -> x      Fy : x -> x      [Fz . 1/(N Lambda) . Bz] : x -> x      By : x -> x      Bx : x -> x
// The bracketed middle pass is one kernel: the fused forward-FFT / division / backward-FFT form (fft_pass.cuh,
// fft_r2x.cuh), or -- for the transport symbol -- the equivalent cyclic recurrence along z (zsolve.cuh), which also
// lets multi-rank (z-slab) plans run without any transpose (apply_device_zslab).
// Every pass is tile-disjoint (a tile is read completely before it is written), so all passes run in place on x
// and b == x aliasing (tests/TransportEquationFFT_SphericalExplosion_impl_mpi.cxx:111) is safe.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <thread>
#include <tuple>

#include "dist.h"
#include "generic_pass.cuh"
#include "plan.h"
#include "registry.cuh"
#include "zsolve.cuh"

namespace cpc {

template <typename T> struct FastRegistry;
template <> struct FastRegistry<double> {
    static void fill(std::map<FastKey<double>, FastEntry<double>> &m)
    {
        fill_fast_f64_odd(m);
        fill_fast_f64_pow2a(m);
        fill_fast_f64_pow2b(m);
        fill_fast_f64_r2x(m);
    }
};
template <> struct FastRegistry<float> {
    static void fill(std::map<FastKey<float>, FastEntry<float>> &m)
    {
        fill_fast_f32_odd(m);
        fill_fast_f32_pow2a(m);
        fill_fast_f32_pow2b(m);
        fill_fast_f32_r2x(m);
    }
};

// ------------------------------------------------------------------------------------------------
// Small set-up kernels
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void invert_table_kernel(const double2 *__restrict__ diag, cplx_t<T> *__restrict__ inv, long long n,
                                    double scale)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double2 d = diag[i];
        const double q = scale / (d.x * d.x + d.y * d.y);
        inv[i] = mk<T>((T)(d.x * q), (T)(-d.y * q));
    }
}

template <typename T>
__global__ void invert_inplace_kernel(cplx_t<T> *__restrict__ tab, long long n, double scale)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const cplx_t<T> d = tab[i];
        const double dx = d.x, dy = d.y;
        const double q = scale / (dx * dx + dy * dy);
        tab[i] = mk<T>((T)(dx * q), (T)(-dy * q));
    }
}

// Diag[k,j,i] = ax[i] + ay[j] + az[k] from the double-precision 1-D tables (what ctx->Diag holds in the reference).
__global__ void diag_from_separable_kernel(double2 *__restrict__ diag, const double2 *__restrict__ ax,
                                           const double2 *__restrict__ ay, const double2 *__restrict__ az, int nx,
                                           int ny, long long n);
// max |Diag - (ax + ay + az)| and max |Diag| over the local slab, as bit patterns of non-negative doubles in out[0], out[1]
// (az already points at the slab's first plane)
__global__ void diag_check_separable_kernel(const double2 *__restrict__ diag, const double2 *__restrict__ ax,
                                            const double2 *__restrict__ ay, const double2 *__restrict__ az, int nx, int ny,
                                            long long n, unsigned long long *out);
// z-slab [z_loc][y][x] -> per-destination chunks [q][z_loc][y_loc][x] (the layout the all-to-all transposes)
template <typename C>
__global__ void slab_to_chunked_kernel(const C *__restrict__ in, C *__restrict__ out, int nx, int ny, int nyl, int nzl)
{
    const long long n = (long long)nx * ny * nzl;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % nx);
        const long long r = i / nx;
        const int y = (int)(r % ny), z = (int)(r / ny);
        const int q = y / nyl, yl = y - q * nyl;
        out[(((long long)q * nzl + z) * nyl + yl) * nx + x] = in[i];
    }
}

template <typename T>
__global__ void diag_from_invtable_kernel(double2 *__restrict__ diag, const cplx_t<T> *__restrict__ inv, long long n,
                                          double scale)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double dx = inv[i].x, dy = inv[i].y;
        const double q = scale / (dx * dx + dy * dy);
        diag[i] = make_double2(dx * q, -dy * q);
    }
}

template <typename T>
__global__ void real_to_complex_kernel(const T *__restrict__ in, cplx_t<T> *__restrict__ out, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = mk<T>(in[i], (T)0);
}
template <typename T>
__global__ void complex_to_real_kernel(const cplx_t<T> *__restrict__ in, T *__restrict__ out, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = in[i].x;
}

// y[i] = sum_p val[p] * x[col[p]] over CSR row i (rows of the projection have a handful of entries: one thread per row)
template <typename T>
__global__ void csr_spmv_kernel(long long rows, const long long *__restrict__ rowptr, const int *__restrict__ colidx,
                                const double *__restrict__ val, const cplx_t<T> *__restrict__ x, cplx_t<T> *__restrict__ y)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
        double sr = 0.0, si = 0.0;
        for (long long p = rowptr[i]; p < rowptr[i + 1]; ++p) {
            const cplx_t<T> v = x[colidx[p]];
            sr += val[p] * (double)v.x;
            si += val[p] * (double)v.y;
        }
        y[i] = mk<T>((T)sr, (T)si);
    }
}

// ------------------------------------------------------------------------------------------------
// PlanT
// ------------------------------------------------------------------------------------------------
template <typename T> struct PlanT : PlanBase {
    using C = cplx_t<T>;

    struct AxisCfg {
        bool fast = false;
        int variant = 0;          // enum Variant of the fast kernel
        int variant_local = -1;   // >= 0: multi-rank plans, variant for plain passes on the local slab (same tile width)
        int nfast = 0;            // transform length of the fast kernel (nx/2 for the r2c / c2r pass)
        int tx = 1;               // lanes per tile
        int threads = 0;          // generic kernel block size
        size_t smem_generic = 0;
        FactorList fl{};
    };

    int n[3] = { 1, 1, 1 };
    int nc = 1;
    // real-scalar plans (CPC_F64 / CPC_F32): b and x are real; the spectrum is the half spectrum [nz][ny][nxp],
    // nxp = nx/2 + 1 rounded up to a multiple of 8 complex (128-byte rows) -- 5 passes of half the bytes.
    bool real = false, real_promote = false;
    int n2 = 0, nxp = 0;
    long long wx = 0;             // contiguous complex elements per (y, z) row in the y and z passes
    C *work = nullptr;            // half spectrum (real plans) or the promoted complex copy
    int num_sms = 148;
    int pf_waves = 0;             // L2 prefetch distance in units of (SM count) CTAs; 0 = off
    // L2-chained schedule (off by default): the x and y passes run z-chunk by z-chunk (Fx then Fy on the same planes,
    // By then Bx), so that the second pass of a pair finds its input in the 126 MB L2.  Measured at 512^3
    // (profiles/r02_notes.md): with separate launches per chunk the kernel boundaries cost as much as the L2 hits
    // save -- 3.28 ms at best (28 MiB chunks on two streams) against 3.27 ms for whole-array passes.
    long long l2_chunk_bytes = 0;             // bytes of x per z-chunk; 0 = whole-array passes
    int chain_streams = 1;        // 2: alternate the chunks' chains between the plan stream and a side stream
    cudaStream_t side_stream = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    // transport symbol with the upwind z column: the middle pass is a cyclic first-order recurrence (zsolve.cuh)
    bool zrec = false;            // the symbol allows it
    bool zrec_off = false;        // CPC_ZSOLVE=0: keep the FFT-based fused pass (comparison / tuning hook)
    double zrec_lz = 0.0;
    int zrec_e = 0;               // points per thread (0: nz does not fit any compiled form)
    // Single-rank recurrence as two thread-per-line sweeps (carry-in from the planes that can still matter, then the
    // solve): ~1 + end_fraction passes, any nz.  Taken when the tile kernel does not fit nz, or when it is slow there
    // (see update_zrec_line) and the decay keeps the first sweep short.
    bool zrec_line = false;
    double end_fraction = 1.0;    // share of the array the first sweep reads (mean over lines of min(M, nz) / nz)
    int zline_mode = -1;          // tuning hook: 1 force the line form, 0 never, -1 choose
    double2 *zcarry = nullptr;    // [nx ny] carry into plane 0 (single-rank line form)
    int stagger = 0;              // start offset (cycles) of every SM's second resident CTA (tuning hook)
    int nzl = 1, z0 = 0;          // local z slab
    int nyl = 1, y0 = 0;          // local y range in the transposed distribution
    long long nloc = 0;           // local elements (slab distribution)
    long long ntot = 0;           // global number of cells * ncomp / ncomp (= nx*ny*nz)
    std::map<FastKey<T>, FastEntry<T>> reg;
    AxisCfg cfg[3];
    C *tw[3] = { nullptr, nullptr, nullptr };    // roots exp(-2 pi i m / n)
    C *stw[3] = { nullptr, nullptr, nullptr };   // per-stage twiddle tables of the fast kernel chosen for the axis
    C *stw_local[3] = { nullptr, nullptr, nullptr };   // stage tables of variant_local, if any

    // symbol state
    C *sym_tab[3] = { nullptr, nullptr, nullptr };       // ax, ay(+1), az in T
    double2 *sym_tab64[3] = { nullptr, nullptr, nullptr };
    C *inv_table = nullptr;
    double wave_c0 = 0, wave_mu[3] = { 0, 0, 0 };

    // projection (unstructured mesh <-> Cartesian grid): P and P^T in CSR
    long long proj_cols = 0;
    long long *p_rowptr = nullptr, *pt_rowptr = nullptr;
    int *p_colidx = nullptr, *pt_colidx = nullptr;
    double *p_val = nullptr, *pt_val = nullptr;
    C *pbuf = nullptr, *pio = nullptr;

    // host staging
    // Pageable host arrays (what a CPU-only PETSc Vec hands over) go through pinned bounce buffers filled / drained by
    // several host threads while the previous buffer is on the wire; pinned arrays are copied directly.
    static constexpr size_t kStageBytes = 64ull << 20;
    char *stage[4] = { nullptr, nullptr, nullptr, nullptr };      // [0,1] host -> device, [2,3] device -> host
    cudaEvent_t stage_ev[4] = {};
    int host_threads = 8;
    C *dbuf = nullptr;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> chunk_ev;
    // profiled applies: one event after every launch, durations summed per pass kind
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_kind;
    size_t prof_n = 0;
    bool prof_on = false;

    // multi-rank
    DistState dist;
    C *sendbuf = nullptr, *tbuf = nullptr;    // transposing schedule only; allocated (and IPC-mapped) on first use
    bool tbuf_tried = false;
    bool want_p2p = false;
    bool p2p = false;                         // peers' buffers are IPC-mapped: transposes are fused into the passes
    void *peer_t[CPC_MAX_PEERS] = {}, *peer_s[CPC_MAX_PEERS] = {};
    // z-slab recurrence (multi-rank transport symbol), all fp64: this rank's line-end values ebuf[nx ny]; the gather
    // buffer of the lines this rank owns, gbuf[P][lsub] (peers store into it), or every rank's end values [P][nx ny]
    // when peers cannot be mapped (ncclAllGather); zinbuf[nx ny] = the carry into this rank's first plane.
    double2 *ebuf = nullptr, *gbuf = nullptr, *zinbuf = nullptr;
    long long lsub = 0;           // lines owned per rank
    bool carry_p2p = false;
    void *peer_g[CPC_MAX_PEERS] = {}, *peer_z[CPC_MAX_PEERS] = {};
    // Optional (tuning hook CPC_XSPLIT=1, off by default): the carry exchange of one half of the columns (kx) runs on
    // xstream -- barrier, owner kernel, barrier: ~40 us of latency -- while the main stream does the forward y pass
    // and the end-value sweep of the other half, and then the solve of the first half.  Needs the flag barrier (two
    // barrier groups) and nx divisible by two tile widths.  Measured at 512^3 on 2 GPUs: 1.80 ms against 1.73 ms for
    // the serial schedule -- the three extra kernel tails cost more than the hidden exchange saves; kept for larger grids.
    bool xsplit = false;
    // Optional (tuning hook CPC_FUSED_SYNC=1, off by default): the kernels of the carry exchange signal and wait
    // themselves (zsolve.cuh FlagSync) instead of separate barrier launches between them: END -> owner -> solve is
    // three launches, not five.  Measured at 512^3: no gain on 2 GPUs (1.75 vs 1.73 ms) and a loss on 8 (0.536 vs
    // 0.508 ms): every block of the signalling kernels pays a system-scope fence behind its NVLink stores, which costs
    // more than the two 1-CTA barrier launches it replaces (profiles/r02_variants_n8.log).
    bool fused_sync = false;
    cudaStream_t xstream = nullptr;
    cudaEvent_t xs_ev[4] = {};
    int zslab_e = 0;              // points per thread for nz / P point lines (0: no tile form fits -> one thread per line)
    bool zslab_line = false;      // tuning hook: always take the thread-per-line form of the second sweep
    bool end_trunc = true;        // end values summed over the planes that can still matter (zs_end_accum_kernel)

    ~PlanT() override
    {
        cudaSetDevice(device);
        for (int a = 0; a < 3; ++a) {
            if (tw[a]) cudaFree(tw[a]);
            if (stw[a]) cudaFree(stw[a]);
            if (stw_local[a]) cudaFree(stw_local[a]);
            if (sym_tab[a]) cudaFree(sym_tab[a]);
            if (sym_tab64[a]) cudaFree(sym_tab64[a]);
        }
        if (inv_table) cudaFree(inv_table);
        if (work) cudaFree(work);
        free_projection();
        if (dbuf) cudaFree(dbuf);
        for (int i = 0; i < 4; ++i) {
            if (stage[i]) cudaFreeHost(stage[i]);
            if (stage_ev[i]) cudaEventDestroy(stage_ev[i]);
        }
        if (p2p || carry_p2p) {
            dist_barrier(dist, stream);
            cudaStreamSynchronize(stream);
            if (p2p) {
                dist_unmap_peers(dist, peer_t);
                dist_unmap_peers(dist, peer_s);
            }
            if (carry_p2p) {
                dist_unmap_peers(dist, peer_g);
                dist_unmap_peers(dist, peer_z);
            }
            dist_barrier(dist, stream);          // nobody frees a buffer a peer still has mapped
            cudaStreamSynchronize(stream);
        }
        if (sendbuf) cudaFree(sendbuf);
        if (ebuf) cudaFree(ebuf);
        if (zcarry) cudaFree(zcarry);
        if (gbuf) cudaFree(gbuf);
        if (zinbuf) cudaFree(zinbuf);
        if (tbuf) cudaFree(tbuf);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (side_stream) cudaStreamDestroy(side_stream);
        if (xstream) cudaStreamDestroy(xstream);
        for (auto e : xs_ev)
            if (e) cudaEventDestroy(e);
        if (fork_ev) cudaEventDestroy(fork_ev);
        if (join_ev) cudaEventDestroy(join_ev);
        for (auto e : chunk_ev) cudaEventDestroy(e);
        for (auto e : prof_ev) cudaEventDestroy(e);
        dist_destroy(dist);
    }

    // ---------------------------------------------------------------------------------------- init
    int init() override
    {
        const bool dbg = getenv("CPC_DEBUG") != nullptr;
#define CPC_TRACE(msg) do { if (dbg) { fprintf(stderr, "[cpc] %s\n", msg); fflush(stderr); } } while (0)
        CPC_TRACE("init begin");
        n[0] = desc.nx; n[1] = desc.ny; n[2] = desc.nz;
        nc = desc.ncomp;
        ntot = (long long)n[0] * n[1] * n[2];
        const SlabRange zr = slab_range(n[2], desc.nranks, desc.rank);
        const SlabRange yr = slab_range(n[1], desc.nranks, desc.rank);
        nzl = zr.count; z0 = zr.start;
        nyl = yr.count; y0 = yr.start;
        if (desc.nranks == 1) { nyl = n[1]; y0 = 0; }
        nloc = (long long)n[0] * n[1] * nzl * nc;
        wx = (long long)n[0] * nc;
        real = (desc.dtype == CPC_F64 || desc.dtype == CPC_F32);
        if (real) {
            if (nc != 1) { set_error("real-scalar plans need ncomp == 1"); return CPC_ERR_ARG; }
        }
        // z-slabs need nz divisible by the ranks; ny too only for the transposing schedule (checked when it is chosen)
        if (desc.nranks > 1 && n[2] % desc.nranks != 0) {
            set_error("multi-rank plans need nz divisible by nranks (nz=%d nranks=%d)", n[2], desc.nranks);
            return CPC_ERR_UNSUPPORTED;
        }
        FastRegistry<T>::fill(reg);
        CPC_TRACE("registry filled");
        if (real) {
            n2 = n[0] / 2;
            if (n[0] % 2 == 0 && reg.find(FastKey<T>(n2, VAR_XMAP, MODE_R2C)) != reg.end()) {
                nxp = (n2 + 1 + 7) / 8 * 8;
                wx = nxp;
                CPC_CUDA(cudaMalloc(&work, sizeof(C) * (size_t)nxp * n[1] * nzl));        // this rank's planes
                CPC_CUDA(cudaMemset(work, 0, sizeof(C) * (size_t)nxp * n[1] * nzl));
            } else {
                if (desc.nranks != 1) { set_error("multi-rank real-scalar plans need an even nx with an r2c kernel (nx=%d)", n[0]); return CPC_ERR_UNSUPPORTED; }
                real_promote = true;         // odd or unsupported nx: run the complex path on a promoted copy
                CPC_CUDA(cudaMalloc(&work, sizeof(C) * (size_t)nloc));
            }
        }

        int dev_smem = 0;
        CPC_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        CPC_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
        // Tuning hooks (tools/ only): read only when CPC_TUNING=1 is set, so that a stray variable in a production
        // environment cannot change which kernels run.  The supported switches are cpc_set_option().
        const bool tuning = getenv("CPC_TUNING") && atoi(getenv("CPC_TUNING")) != 0;
        auto tune = [&](const char *name) -> const char * { return tuning ? getenv(name) : nullptr; };
        if (const char *pf = tune("CPC_PREFETCH_WAVES")) pf_waves = atoi(pf);
        if (const char *zs = tune("CPC_ZSOLVE")) zrec_off = (atoi(zs) == 0);
        if (const char *lc = tune("CPC_L2_CHUNK_MB")) l2_chunk_bytes = (long long)(atof(lc) * 1048576.0);
        if (const char *cs = tune("CPC_CHAIN_STREAMS")) chain_streams = atoi(cs) >= 2 ? 2 : 1;
        zrec_e = zsolve_points_per_thread(n[2]);
        if (const char *ze = tune("CPC_ZSOLVE_E")) {                               // force E if it fits
            constexpr int ZTX = 128 / (int)sizeof(C), ZQW = 32 / ZTX;
            const int E = atoi(ze);
            if ((E == 4 || E == 5 || E == 8 || E == 10 || E == 16) && n[2] % E == 0 && (n[2] / E) % ZQW == 0 &&
                n[2] / E * ZTX <= (E >= 16 ? 512 : 1024))
                zrec_e = E;
        }
        if (const char *sg = tune("CPC_STAGGER")) stagger = atoi(sg);
        if (const char *zl = tune("CPC_ZSLAB_LINE")) zslab_line = atoi(zl) != 0;
        if (const char *et = tune("CPC_END_TRUNC")) end_trunc = atoi(et) != 0;
        if (const char *zm = tune("CPC_ZLINE")) zline_mode = atoi(zm);
        if (const char *xs = tune("CPC_XSPLIT")) xsplit = atoi(xs) != 0;
        if (const char *fs = tune("CPC_FUSED_SYNC")) fused_sync = atoi(fs) != 0;
        CPC_TRACE("got smem attribute");

        // multi-rank plans whose ny is not divisible by the ranks can only run the transpose-free z-slab schedule:
        // their y passes are always local, so the kernels are chosen as for a single rank
        const bool multi_tr = desc.nranks > 1 && n[1] % desc.nranks == 0;
        for (int a = 0; a < 3; ++a) {
            // root table
            std::vector<C> h(n[a]);
            for (int m = 0; m < n[a]; ++m) {
                double re, im;
                exact_root(m, n[a], &re, &im);
                h[m] = mk<T>((T)re, (T)im);
            }
            CPC_CUDA(cudaMalloc(&tw[a], sizeof(C) * n[a]));
            CPC_CUDA(cudaMemcpy(tw[a], h.data(), sizeof(C) * n[a], cudaMemcpyHostToDevice));

            // kernel choice.  Scalar x lines are contiguous in HBM -> XMAP; the wave x pass uses the 4 components of
            // a cell as lanes (narrow strided); y and z lines are strided with >= 128 contiguous bytes across lanes.
            AxisCfg &c = cfg[a];
            int var = (a == 0) ? (nc == 4 ? VAR_NARROW : VAR_XMAP) : VAR_WIDE;
            // the fused pass does two transforms per tile and is LSU / issue bound: 512 = 2 x (16 x 16) with a register
            // radix-2 level and one shared-memory exchange per transform (fft_r2x.cuh) measured 1.23 ms at 512^3, the
            // 8.8.8 kernel with two butterflies per thread 1.31 ms, with one butterfly per thread 1.98 ms
            if (a == 2 && (n[a] == 512 || n[a] == 256)) {
                if (reg.find(FastKey<T>(n[a], VAR_R2X, MODE_FUSED_SEP)) != reg.end()) var = VAR_R2X;
                else if (n[a] == 512 && reg.find(FastKey<T>(n[a], VAR_WIDE2, MODE_FUSED_SEP)) != reg.end()) var = VAR_WIDE2;
            }
            // contiguous 512-point x lines: one warp per line, no block barrier (0.62 vs 0.69 ms at 512^3)
            if (a == 0 && nc == 1 && !real && n[a] == 512 && !tune("CPC_VARIANT_X") &&
                reg.find(FastKey<T>(n[a], VAR_XR2X, MODE_FWD)) != reg.end()) var = VAR_XR2X;
            // the 2 x (16 x 16) kernel also for the plain y passes of 512-point lines: 0.64 ms against 0.70 ms (8.8.8)
            // (single-rank only: in the chunked multi-rank layout its paired loads k / k+256 sit exactly one chunk,
            // a large power of two, apart and collide in the L2 / DRAM address hash: 0.57 vs 0.35 ms per half slab)
            // 256-point y lines otherwise: radix 8.8.4 with 8 points per thread and 4 small CTAs per SM (0.090 vs 0.108 ms
            // at 256^3; the 2 x (16 x 8) kernel: 0.086 ms)
            if (a == 1 && n[a] == 256 && reg.find(FastKey<T>(n[a], VAR_SMALL, MODE_FWD)) != reg.end()) var = VAR_SMALL;
            // 1024-point y lines: 64 KB tiles (4 lanes) so that two CTAs share an SM: 7.1 vs 8.2 ms at 1024^3.  Not for z,
            // whose 16 MB line stride makes 64-byte segments slower than the one-CTA 128-byte tiles (14.2 vs 13.3 ms)
            if (a == 1 && n[a] == 1024 && !multi_tr && reg.find(FastKey<T>(n[a], VAR_SLIM, MODE_FWD)) != reg.end())
                var = VAR_SLIM;
            if (a == 1 && (n[a] == 512 || n[a] == 256) && !multi_tr &&
                reg.find(FastKey<T>(n[a], VAR_R2X, MODE_FWD)) != reg.end()) var = VAR_R2X;
            {
                const char *names[3] = { "CPC_VARIANT_X", "CPC_VARIANT_Y", "CPC_VARIANT_Z" };
                const char *ov = tune(names[a]);
                if (ov && *ov) var = atoi(ov);      // tuning hook (tools/), not part of the ABI
            }
            const int nfast = (real && !real_promote && a == 0) ? n2 : n[a];     // r2c: nx real points = nx/2 complex
            if (real && !real_promote && a == 0) var = VAR_XMAP;
            // contiguous 256-point lines (complex nx = 256, or the r2c / c2r pass of real nx = 512): half a warp per line
            if (a == 0 && nc == 1 && nfast == 256 && !tune("CPC_VARIANT_X") &&
                reg.find(FastKey<T>(256, VAR_XR2X, real ? MODE_R2C : MODE_FWD)) != reg.end()) var = VAR_XR2X;
            if (reg.find(FastKey<T>(nfast, var, MODE_FWD)) == reg.end() && var != VAR_XMAP) var = VAR_NARROW;
            auto it = reg.find(FastKey<T>(nfast, var, MODE_FWD));
            if (it != reg.end() && a == 2 && reg.find(FastKey<T>(n[a], var, MODE_FUSED_SEP)) == reg.end()) it = reg.end();
            // multi-rank fast kernels address chunks with shifts: ny/nranks and nz/nranks must be powers of two
            if (it != reg.end() && multi_tr && a >= 1 && ((nyl & (nyl - 1)) != 0 || (nzl & (nzl - 1)) != 0)) it = reg.end();
            // multi-rank plans push / chunk their y and z stores: only variants with a general-addressing build
            if (it != reg.end() && multi_tr && a >= 1 &&
                reg.find(FastKey<T>(nfast, var, MODE_FWD + GEN_BIT)) == reg.end()) {
                var = VAR_WIDE;
                it = reg.find(FastKey<T>(nfast, var, MODE_FWD + GEN_BIT)) != reg.end() ? reg.find(FastKey<T>(nfast, var, MODE_FWD))
                                                                                  : reg.end();
            }
            c.nfast = nfast;
            if (it != reg.end() && it->second.smem <= (size_t)dev_smem) {
                c.fast = true;
                c.variant = var;
                c.tx = it->second.tx;
                int rc_t = build_stage_table(it->second.radix, &stw[a]);
                if (rc_t) return rc_t;
                // multi-rank: the 2 x (R0 x R1) kernel for y passes that run on the local slab (no split)
                if (a == 1 && multi_tr && (n[a] == 512 || n[a] == 256) && var != VAR_R2X) {
                    auto il = reg.find(FastKey<T>(n[a], VAR_R2X, MODE_FWD));
                    if (il != reg.end() && il->second.tx == it->second.tx && il->second.smem <= (size_t)dev_smem) {
                        c.variant_local = VAR_R2X;
                        if ((rc_t = build_stage_table(il->second.radix, &stw_local[a]))) return rc_t;
                    }
                }
            } else {
                c.fast = false;
                c.fl.n = n[a];
                c.fl.nfac = 0;
                int rem = n[a];
                for (int p = 2; rem > 1;) {
                    if (rem % p == 0) {
                        if (c.fl.nfac >= CPC_MAX_FACTORS) { set_error("too many factors"); return CPC_ERR_UNSUPPORTED; }
                        c.fl.fac[c.fl.nfac++] = p;
                        rem /= p;
                    } else {
                        ++p;
                        if ((long long)p * p > rem) p = rem;
                    }
                }
                if (c.fl.nfac == 0) c.fl.fac[c.fl.nfac++] = 1;      // length-1 axis: a radix-1 "copy" stage
                // the 2s become radix 8 / 4 (fewer stages and barriers), largest radix first so that the cheap
                // twiddle-free first stage (p = 1) is the widest one
                {
                    int twos = 0, w = 0, rest[CPC_MAX_FACTORS];
                    for (int i = 0; i < c.fl.nfac; ++i) {
                        if (c.fl.fac[i] == 2) ++twos;
                        else rest[w++] = c.fl.fac[i];
                    }
                    int k = 0;
                    for (; twos >= 3 && twos != 4; twos -= 3) c.fl.fac[k++] = 8;
                    for (; twos >= 2; twos -= 2) c.fl.fac[k++] = 4;
                    if (twos) c.fl.fac[k++] = 2;
                    for (int i = 0; i < w; ++i) c.fl.fac[k++] = rest[i];
                    c.fl.nfac = k;
                    // a first stage that is an O(R) sum would read every input R times from HBM: stage the tile
                    // through a radix-1 copy instead
                    const int r0 = c.fl.fac[0];
                    if (!(r0 == 1 || r0 == 2 || r0 == 3 || r0 == 4 || r0 == 5 || r0 == 7 || r0 == 8)) {
                        if (c.fl.nfac >= CPC_MAX_FACTORS) { set_error("too many factors"); return CPC_ERR_UNSUPPORTED; }
                        for (int i = c.fl.nfac; i > 0; --i) c.fl.fac[i] = c.fl.fac[i - 1];
                        c.fl.fac[0] = 1;
                        ++c.fl.nfac;
                    }
                }
                int gtx = (a == 0 && nc == 4) ? 4 : 8;
                // two CTAs per SM when 4 lanes (64-byte segments) allow it; below that only to fit at all
                if (gtx == 8 && 2ull * n[a] * 8 * sizeof(C) > (size_t)dev_smem / 2 - 1024) gtx = 4;
                while (gtx > ((nc == 4) ? 4 : 1) && 2ull * n[a] * gtx * sizeof(C) > (size_t)dev_smem) gtx >>= 1;
                if (2ull * n[a] * gtx * sizeof(C) > (size_t)dev_smem) {
                    set_error("axis length %d too long for the generic kernel", n[a]);
                    return CPC_ERR_UNSUPPORTED;
                }
                c.tx = gtx;
                c.smem_generic = 2ull * n[a] * gtx * sizeof(C);
                // CTA size: the stage of radix R has (n / R) * lanes butterflies, one per thread per round.  Pick the
                // size whose rounds waste the fewest thread slots (320-point lines: 320 threads do the radix-8 stages
                // in one full round, 256 would need two rounds with the second one a quarter full).  <= 320 threads
                // keeps two CTAs of this 96-register kernel on an SM.
                {
                    long long best = -1;
                    int best_t = 256;
                    for (int t = 128; t <= 320; t += 32) {
                        long long cost = 2 * (((long long)n[a] * gtx + t - 1) / t) * t;        // load + store sweeps
                        for (int i = 0; i < c.fl.nfac; ++i) {
                            const int R = c.fl.fac[i];
                            const bool bf = (R == 2 || R == 3 || R == 4 || R == 5 || R == 7 || R == 8);
                            const long long items = bf ? (long long)(n[a] / R) * gtx : (long long)n[a] * gtx;
                            cost += ((items + t - 1) / t) * t * R;
                        }
                        if (best < 0 || cost < best || (cost == best && t > best_t)) { best = cost; best_t = t; }
                    }
                    const long long work = (long long)n[a] * gtx;
                    c.threads = work >= 128 ? best_t : (int)((work + 31) / 32 * 32);
                }
            }
        }
        CPC_TRACE("axes configured");
        // opt in to large dynamic shared memory for every kernel we may launch
        for (auto &kv : reg) {
            if (kv.second.smem > 48 * 1024 && kv.second.smem <= (size_t)dev_smem)
                CPC_CUDA(cudaFuncSetAttribute((const void *)kv.second.kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)kv.second.smem));
        }
        CPC_CUDA(cudaFuncSetAttribute((const void *)generic_pass_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      dev_smem));
        CPC_TRACE("func attributes set");
        CPC_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        CPC_CUDA(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
        CPC_CUDA(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
        CPC_CUDA(cudaEventCreateWithFlags(&join_ev, cudaEventDisableTiming));
        if (desc.nranks > 1) {
            int rc = dist_init(dist, desc.nranks, desc.rank, desc.nccl_unique_id, device);
            if (rc) return rc;
            // Peer mapping (CUDA IPC over NVLink) serves the carry exchange and the fused transposes; without it
            // (or with the tuning hook CPC_DIST_MODE=nccl) both fall back to NCCL collectives.
            const char *mode = tune("CPC_DIST_MODE");
            want_p2p = !(mode && strcmp(mode, "nccl") == 0) && desc.nranks <= CPC_MAX_PEERS;
            const char *fb = tune("CPC_FLAG_BARRIER");
            if (want_p2p && !(fb && atoi(fb) == 0) && (rc = dist_flag_barrier_init(dist, device, stream))) return rc;
            if (nc == 1) {
                zslab_e = zsolve_points_per_thread(nzl);
                const long long L = wx * n[1];        // (kx, ky) lines: nx ny, or the padded half spectrum of a real plan
                lsub = (L + desc.nranks - 1) / desc.nranks;
                CPC_CUDA(cudaMalloc(&ebuf, sizeof(double2) * (size_t)L));
                CPC_CUDA(cudaMalloc(&zinbuf, sizeof(double2) * (size_t)L));
                CPC_CUDA(cudaMalloc(&gbuf, sizeof(double2) * (size_t)lsub * desc.nranks));
                if (want_p2p) {
                    int r1 = dist_map_peers(dist, gbuf, peer_g, device, stream);
                    int r2 = r1 == CPC_OK ? dist_map_peers(dist, zinbuf, peer_z, device, stream) : r1;
                    if (r1 == CPC_OK && r2 != CPC_OK) dist_unmap_peers(dist, peer_g);
                    carry_p2p = (r1 == CPC_OK && r2 == CPC_OK);
                    if (!carry_p2p && r1 != CPC_ERR_UNSUPPORTED && r2 != CPC_ERR_UNSUPPORTED) return r1 ? r1 : r2;
                }
                if (carry_p2p && dist.flag_barrier) {
                    int lo = 0, hi = 0;
                    CPC_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                    CPC_CUDA(cudaStreamCreateWithPriority(&xstream, cudaStreamNonBlocking, hi));
                    for (auto &e : xs_ev) CPC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                }
                if (!carry_p2p) {                 // all-gather fallback: every rank holds every rank's end values
                    cudaFree(gbuf);
                    gbuf = nullptr;
                    CPC_CUDA(cudaMalloc(&gbuf, sizeof(double2) * (size_t)L * desc.nranks));
                }
            }
        }
        return CPC_OK;
    }

    // Buffers of the transposing schedule (two slabs, IPC-mapped into every peer when possible): allocated the first
    // time a symbol needs that schedule.  Collective: every rank reaches it in the same call (the symbol is the same
    // on all ranks).
    int ensure_transpose_buffers()
    {
        if (tbuf_tried) return (sendbuf && tbuf) ? CPC_OK : CPC_ERR_NOMEM;
        tbuf_tried = true;
        if (n[1] % desc.nranks != 0) {
            set_error("the transposing multi-rank schedule needs ny divisible by nranks (ny=%d nranks=%d)", n[1], desc.nranks);
            return CPC_ERR_UNSUPPORTED;
        }
        CPC_CUDA(cudaMalloc(&sendbuf, sizeof(C) * nloc));
        CPC_CUDA(cudaMalloc(&tbuf, sizeof(C) * nloc));
        if (want_p2p) {
            int r1 = dist_map_peers(dist, tbuf, peer_t, device, stream);
            int r2 = r1 == CPC_OK ? dist_map_peers(dist, sendbuf, peer_s, device, stream) : r1;
            if (r1 == CPC_OK && r2 != CPC_OK) dist_unmap_peers(dist, peer_t);
            p2p = (r1 == CPC_OK && r2 == CPC_OK);
            if (!p2p && r1 != CPC_ERR_UNSUPPORTED && r2 != CPC_ERR_UNSUPPORTED) return r1 ? r1 : r2;
        }
        return CPC_OK;
    }

    // per-stage twiddle tables [r-1][k]: exp(-2 pi i r k / (P R)) for the 2nd and 3rd radix of a kernel
    int build_stage_table(const int *R, C **dst)
    {
        std::vector<C> st;
        int P = R[0];
        for (int sidx = 1; sidx < 3; ++sidx) {
            if (R[sidx] <= 1) break;
            for (int r = 1; r < R[sidx]; ++r)
                for (int k = 0; k < P; ++k) {
                    double re, im;
                    exact_root((long long)r * k, (long long)P * R[sidx], &re, &im);
                    st.push_back(mk<T>((T)re, (T)im));
                }
            P *= R[sidx];
        }
        if (!st.empty()) {
            CPC_CUDA(cudaMalloc(dst, sizeof(C) * st.size()));
            CPC_CUDA(cudaMemcpy(*dst, st.data(), sizeof(C) * st.size(), cudaMemcpyHostToDevice));
        }
        return CPC_OK;
    }

    // ------------------------------------------------------------------------------------ geometry
    // Geometry of the pass along `axis` over z planes [zb, zb+zc) of an array with ny_ rows per plane.
    // layout: 0 = slab [z][y][x][c] with ny rows; 1 = transposed [z_glob][y_loc][x][c] (axis 2 only).
    PassGeom make_geom(int axis, int tx, int zb, int zc, int layout, long long *ptr_off) const
    {
        PassGeom g{};
        const long long W = wx;
        const int rows = (layout == 1) ? nyl : n[1];
        g.ncomp = nc; g.nx = (int)(wx / nc); g.ny = rows; g.y0 = (layout == 1) ? y0 : 0;
        *ptr_off = 0;
        if (axis == 0) {
            const long long lines = (long long)rows * zc;
            *ptr_off = (long long)zb * rows * W;
            if (real && !real_promote) {
                // r2c / c2r: the real line of nx points is read (written) as nx/2 complex points; the half spectrum
                // row has pitch nxp.  Which side is which is fixed up in run_real_x().
                g.SI = 1; g.SL = n2;
                g.tiles_inner = (int)((lines + tx - 1) / tx);
                g.B0 = (long long)tx * n2; g.B1 = 0;
                g.lines_inner = (int)lines;
                g.ntiles = g.tiles_inner;
                *ptr_off = 0;
            } else if (nc == 1) {
                g.SI = 1; g.SL = n[0];
                g.tiles_inner = (int)((lines + tx - 1) / tx);
                g.B0 = (long long)tx * n[0]; g.B1 = 0;
                g.lines_inner = (int)lines;
                g.ntiles = g.tiles_inner;
            } else {
                g.SI = nc; g.SL = 1;
                g.tiles_inner = 1; g.B0 = 0; g.B1 = W;
                g.lines_inner = nc;
                g.ntiles = (int)lines;
            }
        } else if (axis == 1) {
            *ptr_off = (long long)zb * rows * W;
            g.SI = W; g.SL = 1;
            g.tiles_inner = (int)((W + tx - 1) / tx);
            g.B0 = tx; g.B1 = W * rows;
            g.lines_inner = (int)W;
            g.ntiles = g.tiles_inner * zc;
        } else {
            const long long inner = W * rows;
            g.SI = inner; g.SL = 1;
            g.tiles_inner = (int)((inner + tx - 1) / tx);
            g.B0 = tx; g.B1 = 0;
            g.lines_inner = (int)inner;
            g.ntiles = g.tiles_inner;
        }
        g.SIo = g.SI; g.B0o = g.B0; g.B1o = g.B1; g.SLo = g.SL;
        g.SCi = g.SCo = 0; g.Di = g.Do = 0; g.shi = g.sho = -1;
        g.maski = g.masko = 0x7fffffff;
        g.pf_tiles = 0;
        g.stagger = stagger; g.num_sms = num_sms;
        g.npeer = 0;
        for (int q = 0; q < CPC_MAX_PEERS; ++q) g.peer[q] = nullptr;
        return g;
    }

    // Multi-rank y pass: one side is the z-slab [z_loc][y][x][c], the other the per-destination chunked layout
    // [q][z_loc][y_loc][x][c] (q = rank owning global y = q*nyl + y_loc after the transpose).
    void make_split(PassGeom &g, bool split_out) const
    {
        const long long W = (long long)n[0] * nc;
        int sh = -1;
        for (int b = 0; b < 31; ++b)
            if ((1 << b) == nyl) sh = b;
        if (split_out) {
            g.SIo = W; g.B0o = g.B0; g.B1o = (long long)nyl * W;
            g.Do = nyl; g.sho = sh; g.SCo = (long long)nzl * nyl * W;
            g.masko = nyl - 1;
        } else {
            g.maski = nyl - 1;
            g.B1o = g.B1; g.B0o = g.B0; g.SIo = g.SI; g.SLo = g.SL;
            g.B1 = (long long)nyl * W;
            g.Di = nyl; g.shi = sh; g.SCi = (long long)nzl * nyl * W;
        }
    }

    SymbolArgs<T> symbol_args() const
    {
        SymbolArgs<T> s{};
        s.ax = sym_tab[0]; s.ay = sym_tab[1]; s.az = sym_tab[2];
        s.inv_table = inv_table;
        s.rx = tw[0]; s.ry = tw[1]; s.rz = tw[2];
        s.c0 = (T)wave_c0; s.mux = (T)wave_mu[0]; s.muy = (T)wave_mu[1]; s.muz = (T)wave_mu[2];
        s.scale = (T)(1.0 / (double)ntot);
        return s;
    }

    // Launch one pass.  `in`/`out` point at the start of the local array.
    // split: 0 = none, 1 = store side chunked (forward y of a multi-rank plan), 2 = load side chunked (backward y)
    // split: 0 = none, 1 = store side chunked, 2 = load side chunked, 3 = stores pushed to the peers' transposed
    // buffers (forward y), 4 = stores pushed to the peers' chunked buffers (fused z)
    // xs0 / xsn (y passes only): restrict the pass to columns [xs0, xs0 + xsn) of every row, a multiple of the tile width
    int run_pass(int axis, int mode, const C *in, C *out, int zb, int zc, int layout, cudaStream_t st, int split = 0,
                 int xs0 = 0, int xsn = 0)
    {
        const AxisCfg &c = cfg[axis];
        long long off = 0;
        // middle pass of a transport symbol: cyclic recurrence along z (zsolve.cuh), 128-byte rows as lanes.
        // Multi-rank pushes need power-of-two chunks (shift / mask addressing), as for the FFT kernels.
        // Line form (single rank): carry into plane 0 from the planes that can still matter, then one thread per line
        // marches along z (zsolve.cuh); chosen by set_symbol_tables when the tile kernel does not fit or is slow.
        if (axis == 2 && mode == MODE_FUSED_SEP && zrec && !zrec_off && zrec_line && split == 0 && layout == 0 &&
            desc.nranks == 1 && zb == 0 && zc == n[2]) {
            const long long L = wx * n[1];       // wx = nx, or the padded half-spectrum row of a real plan
            if (!zcarry) CPC_CUDA(cudaMalloc(&zcarry, sizeof(double2) * (size_t)L));
            ZSolveArgs za = zsolve_args();
            za.nline = n[2];
            za.zin = zcarry;
            const int egrid = (int)((L + 255) / 256);
            zs_end_accum_kernel<T><<<egrid, 256, 0, st>>>(in, L, (int)wx, 0, n[2], 0, end_trunc ? 1 : 0, zcarry, za, -2, 1, ZCarryPeers{},
                                                          FlagSync{});
            zs_dist_line_kernel<T><<<egrid, 256, 0, st>>>(in, out, L, (int)wx, n[2], za, FlagSync{});
            launches += 2;
            CPC_CUDA(cudaGetLastError());
            return CPC_OK;
        }
        const bool zs = axis == 2 && mode == MODE_FUSED_SEP && zrec && !zrec_off && zrec_e > 0 && nc == 1 &&
                        (split == 0 || (nzl & (nzl - 1)) == 0);
        PassGeom g = make_geom(axis, zs ? 128 / (int)sizeof(C) : c.tx, zb, zc, layout, &off);
        if (axis == 1 && xsn > 0) {
            g.tiles_inner = xsn / c.tx;
            g.lines_inner = xsn;
            g.ntiles = g.tiles_inner * zc;
            off += xs0;
        }
        if (split == 1 || split == 2) make_split(g, split == 1);
        if (split == 3 || split == 4) {
            const long long W = (long long)n[0] * nc;
            int sh = -1;
            const int D = split == 3 ? nyl : nzl;
            for (int b = 0; b < 31; ++b)
                if ((1 << b) == D) sh = b;
            g.npeer = desc.nranks;
            g.Do = D; g.sho = sh; g.SCo = 0;
            g.masko = D - 1;
            if (split == 3) {
                // point ky of the y line at (z_loc, x) -> rank q = ky / nyl, element [(z0 + z_loc)][ky % nyl][x]
                g.SIo = W; g.B0o = g.B0; g.B1o = (long long)nyl * W;
                for (int q = 0; q < desc.nranks; ++q) g.peer[q] = (C *)peer_t[q] + (long long)z0 * nyl * W;
            } else {
                // point k of the z line at (y_loc, x) -> rank s = k / nzl, element [me][k % nzl][y_loc][x]
                g.SIo = (long long)nyl * W; g.B0o = g.B0; g.B1o = 0;
                for (int q = 0; q < desc.nranks; ++q) g.peer[q] = (C *)peer_s[q] + (long long)desc.rank * (nloc / desc.nranks);
            }
        }
        if (g.ntiles <= 0) return CPC_OK;
        SymbolArgs<T> s = symbol_args();
        s.rz = tw[axis];          // roots of the transformed axis (fft_r2x.cuh); the fused pass is always axis 2
        if (zs) {
            if (split != 0) {
                if (g.Di == 0) { g.shi = 31; g.maski = 0x7fffffff; g.SCi = 0; }
                if (g.Do == 0) { g.sho = 31; g.masko = 0x7fffffff; g.SCo = 0; }
                if (g.npeer == 0)
                    for (int q = 0; q < CPC_MAX_PEERS; ++q) g.peer[q] = (void *)(out + off);
            }
            launch_zsolve_e(zrec_e, ZS_CYCLIC, split != 0, n[2], g.ntiles, st, in + off, out + off, g);
        } else if (c.fast) {
            // multi-rank y / z passes need the general-addressing build (init() made sure it exists)
            if (split != 0) {
                // branch-free general addressing: a side without a split behaves as one chunk of 2^31 points, and
                // without peers every "peer" is the local output
                if (g.Di == 0) { g.shi = 31; g.maski = 0x7fffffff; g.SCi = 0; }
                if (g.Do == 0) { g.sho = 31; g.masko = 0x7fffffff; g.SCo = 0; }
                if (g.npeer == 0)
                    for (int q = 0; q < CPC_MAX_PEERS; ++q) g.peer[q] = (void *)(out + off);
                if (g.shi < 0 || g.sho < 0) { set_error("multi-rank fast path needs power-of-two ny/nranks and nz/nranks"); return CPC_ERR_UNSUPPORTED; }
            }
            // y passes that stay on the local z-slab (the z-slab recurrence path) need no chunked addressing
            const bool alt = c.variant_local >= 0 && split == 0 && layout == 0 && (mode == MODE_FWD || mode == MODE_INV);
            const FastEntry<T> &e = reg.at(FastKey<T>(c.nfast, alt ? c.variant_local : c.variant, mode + (split != 0 ? GEN_BIT : 0)));
            const int grid = (g.ntiles + e.g - 1) / e.g;
            g.pf_tiles = pf_waves > 0 ? pf_waves * num_sms * e.g : 0;
            e.kern<<<grid, e.threads, e.smem, st>>>(in + off, out + off, g, alt ? stw_local[axis] : stw[axis], s);
        } else {
            generic_pass_kernel<T><<<g.ntiles, c.threads, c.smem_generic, st>>>(in + off, out + off, g, tw[axis], s,
                                                                               c.fl, c.tx, mode);
        }
        ++launches;
        CPC_CUDA(cudaGetLastError());
        return CPC_OK;
    }

    // Real plans: r2c (forward, real array -> half spectrum) or c2r (backward) x pass over z planes [zb, zb+zc).
    int run_real_x(bool forward, const void *in, void *out, int zb, int zc, cudaStream_t st)
    {
        const AxisCfg &c = cfg[0];
        long long off = 0;
        PassGeom g = make_geom(0, c.tx, zb, zc, 0, &off);
        const long long line0 = (long long)zb * n[1];
        const C *pin;
        C *pout;
        if (forward) {       // in: real [lines][nx] viewed as complex pitch n2; out: half spectrum pitch nxp
            g.SLo = nxp; g.B0o = (long long)c.tx * nxp;
            pin = (const C *)in + line0 * n2;
            pout = (C *)out + line0 * nxp;
        } else {
            g.SL = nxp; g.B0 = (long long)c.tx * nxp;
            g.SLo = n2; g.B0o = (long long)c.tx * n2;
            pin = (const C *)in + line0 * nxp;
            pout = (C *)out + line0 * n2;
        }
        if (g.ntiles <= 0) return CPC_OK;
        const SymbolArgs<T> s = symbol_args();
        const FastEntry<T> &e = reg.at(FastKey<T>(c.nfast, c.variant, forward ? MODE_R2C : MODE_C2R));
        const int grid = (g.ntiles + e.g - 1) / e.g;
        e.kern<<<grid, e.threads, e.smem, st>>>(pin, pout, g, stw[0], s);
        ++launches;
        CPC_CUDA(cudaGetLastError());
        return CPC_OK;
    }

    int fused_mode() const
    {
        switch (symbol_kind) {
        case CPC_SYMBOL_SEPARABLE: return MODE_FUSED_SEP;
        case CPC_SYMBOL_TABLE: return MODE_FUSED_TABLE;
        case CPC_SYMBOL_WAVE: return MODE_FUSED_WAVE;
        default: return -1;
        }
    }

    // ------------------------------------------------------------------------------------- symbols
    int upload_tables(const std::vector<double2> (&h)[3])
    {
        for (int a = 0; a < 3; ++a) {
            // real plans index the x table with the padded half-spectrum column (< nxp <= nx for nx >= 16)
            const size_t len = (a == 0 && real && !real_promote && (size_t)nxp > (size_t)n[a]) ? (size_t)nxp : (size_t)n[a];
            std::vector<C> ht(len, mk<T>((T)1, (T)0));
            for (int m = 0; m < n[a]; ++m) ht[m] = mk<T>((T)h[a][m].x, (T)h[a][m].y);
            if (!sym_tab[a]) CPC_CUDA(cudaMalloc(&sym_tab[a], sizeof(C) * len));
            std::vector<double2> h64(len, make_double2(1.0, 0.0));
            for (int m = 0; m < n[a]; ++m) h64[m] = h[a][m];
            if (!sym_tab64[a]) CPC_CUDA(cudaMalloc(&sym_tab64[a], sizeof(double2) * len));
            CPC_CUDA(cudaMemcpyAsync(sym_tab[a], ht.data(), sizeof(C) * len, cudaMemcpyHostToDevice, stream));
            CPC_CUDA(cudaMemcpyAsync(sym_tab64[a], h64.data(), sizeof(double2) * len, cudaMemcpyHostToDevice, stream));
            CPC_CUDA(cudaStreamSynchronize(stream));
        }
        return CPC_OK;
    }

    // Separable symbol from the three 1-D tables h[a][m] (already multiplied by lambda; the "+1" rides on y).
    // Decides whether the middle pass may run as the cyclic recurrence of zsolve.cuh: the z table must be
    // lambda_z (1 - exp(-2 pi i k / nz)) -- the DFT of the reference's upwind column [1, -1, 0, ...]
    // (build_transport_col, FftLinearSolver_3D.c:80-90) -- for some 0 <= lambda_z <= 4096, and
    // Re(ax[i] + ay[j]) >= 1/2 everywhere, so that |lambda_z / (alpha + lambda_z)| < 1.
    int set_symbol_tables(const std::vector<double2> (&h)[3])
    {
        int rc = upload_tables(h);
        if (rc) return rc;
        symbol_kind = CPC_SYMBOL_SEPARABLE;
        double lz = 0.0;
        zrec = symbol_recurrence_lambda(n[0], n[1], n[2], h[0].data(), h[1].data(), h[2].data(), &lz) != 0;
        zrec_lz = lz;
        // how much of a line the end-value sweep has to read: planes whose weight |c|^m can reach 1e-17
        end_fraction = 1.0;
        if (zrec && n[2] > 1) {
            double acc = 0.0;
            for (int j = 0; j < n[1]; ++j)
                for (int i = 0; i < n[0]; ++i) {
                    const double ar = h[0][i].x + h[1][j].x + lz, ai = h[0][i].y + h[1][j].y;
                    const double c2 = lz * lz / (ar * ar + ai * ai);
                    double m = c2 > 0.0 ? -78.3 / std::log(c2) + 1.0 : 1.0;
                    if (!(m < (double)n[2])) m = (double)n[2];
                    acc += m;
                }
            end_fraction = acc / ((double)n[0] * n[1] * n[2]);
        }
        update_zrec_line();
        return CPC_OK;
    }

    void update_zrec_line()
    {
        // Measured (profiles/r02_notes.md): the tile kernel wins with 16 points per thread and two CTAs per SM (512^3:
        // 0.70 against 0.82 ms); the line form wins when the tile kernel runs one CTA per SM (1024^3: 8.2 -> 6.0 ms) or
        // with fewer points per thread (500^3: 1.57 -> 0.78 ms, 1000^3: 9.1 -> 5.5 ms).  It needs enough lines to fill
        // the GPU with one thread each (400^3, 160 k lines: 0.44 ms against 0.33 for the tile kernel).  complex64
        // storage: the tile kernel computes in fp64 on 16-lane rows, 512 threads at 128 registers -- one CTA per SM,
        // 0.70 ms at 512^3 where the line form streams the 1 GB array at full rate.
        const bool fills = wx * n[1] >= 200000;
        zrec_line = zrec && desc.nranks == 1 && nc == 1 && !real_promote && n[2] > 1 &&
                    (zline_mode == 1 ||
                     (zline_mode != 0 && (zrec_e == 0 || (end_trunc && end_fraction < 0.25 && fills &&
                                                         (zrec_e != 16 || n[2] >= 1024 || sizeof(T) == 4)))));
    }

    int set_symbol_separable(const double *cx, const double *cy, const double *cz, double lx, double ly,
                             double lz) override
    {
        if (nc != 1) { set_error("separable symbol needs ncomp == 1"); return CPC_ERR_ARG; }
        const double *c[3] = { cx, cy, cz };
        const double lam[3] = { lx, ly, lz };
        std::vector<double2> h[3];
        for (int a = 0; a < 3; ++a) {
            if (!c[a]) { set_error("null eigenvalue table"); return CPC_ERR_ARG; }
            h[a].resize(n[a]);
            for (int m = 0; m < n[a]; ++m) {
                h[a][m].x = lam[a] * c[a][2 * m] + (a == 1 ? 1.0 : 0.0);   // the "+1" (VecShift, :155) rides on y
                h[a][m].y = lam[a] * c[a][2 * m + 1];
            }
        }
        return set_symbol_tables(h);
    }

    // Points per thread of the recurrence kernel for nz-point lines: nz / E segments, a multiple of the 32 / TX
    // segments a warp holds, at most 32 warps and within the CTA size the kernel is compiled for.  0 = no fit.
    static int zsolve_points_per_thread(int nz)
    {
        constexpr int TX = 128 / (int)sizeof(C), QW = 32 / TX;
        const int cand[5] = { 16, 8, 10, 5, 4 };
        for (int i = 0; i < 5; ++i) {
            const int E = cand[i];
            if (nz % E) continue;
            const int S = nz / E;
            if (S % QW) continue;
            const int threads = S * TX;
            const int maxt = E >= 16 ? 512 : 1024;       // fp64 arithmetic whatever the storage type (zsolve.cuh)
            if (threads > maxt) continue;
            return E;
        }
        return 0;
    }

    ZSolveArgs zsolve_args() const
    {
        ZSolveArgs a{};
        a.ax = sym_tab64[0]; a.ay = sym_tab64[1];
        a.lz = zrec_lz;
        a.scale = (double)n[2] / (double)ntot;
        a.n = n[2];
        a.zin = zinbuf;
        a.xs0 = 0;
        a.xsn = (int)wx;
        return a;
    }

    // nline = points of the z line a tile holds (nz, or nz / P for the z-slab sweeps)
    FlagSync make_sync(int group, unsigned long long epoch) const
    {
        FlagSync f{};
        for (int q = 0; q < desc.nranks; ++q) f.peer[q] = (unsigned long long *)dist.peer_flags[q] + group * CPC_DIST_MAX_PEERS;
        f.mine = dist.flags + group * CPC_DIST_MAX_PEERS;
        f.epoch = epoch;
        f.counter = dist.sync_counters + group;
        f.timeout = dist.timeout_flag;
        f.nranks = desc.nranks;
        f.rank = desc.rank;
        return f;
    }

    template <int E> void launch_zsolve(int kind, bool gen, int nline, int grid, cudaStream_t st, const C *in, C *out,
                                        const PassGeom &g, const FlagSync &wait = FlagSync{})
    {
        constexpr int TX = 128 / (int)sizeof(C);
        ZSolveArgs a = zsolve_args();
        a.nline = nline;
        // CTAs of at least 128 threads: several tiles per CTA when the line is short
        const int tpt = nline / E * TX;
        const int gpc = tpt >= 128 ? 1 : 128 / tpt;
        const int threads = tpt * gpc;
        grid = (grid + gpc - 1) / gpc;
        if (kind == ZS_DIST) zsolve_kernel<T, E, false, ZS_DIST><<<grid, threads, 0, st>>>(in, out, g, a, wait);
        else if (gen) zsolve_kernel<T, E, true><<<grid, threads, 0, st>>>(in, out, g, a, wait);
        else zsolve_kernel<T, E, false><<<grid, threads, 0, st>>>(in, out, g, a, wait);
    }
    void launch_zsolve_e(int e, int kind, bool gen, int nline, int grid, cudaStream_t st, const C *in, C *out,
                         const PassGeom &g, const FlagSync &wait = FlagSync{})
    {
        switch (e) {
        case 16: launch_zsolve<16>(kind, gen, nline, grid, st, in, out, g, wait); break;
        case 8: launch_zsolve<8>(kind, gen, nline, grid, st, in, out, g, wait); break;
        case 10: launch_zsolve<10>(kind, gen, nline, grid, st, in, out, g, wait); break;
        case 5: launch_zsolve<5>(kind, gen, nline, grid, st, in, out, g, wait); break;
        default: launch_zsolve<4>(kind, gen, nline, grid, st, in, out, g, wait); break;
        }
    }

    // z-slab plans with a transport symbol: no transpose at all (zsolve.cuh)
    bool use_zslab() const
    {
        return desc.nranks > 1 && symbol_kind == CPC_SYMBOL_SEPARABLE && zrec && !zrec_off && ebuf && gbuf && zinbuf;
    }

    int set_symbol_transport(double lx, double ly, double lz) override
    {
        // c = [1,-1,0..] (FftLinearSolver_3D.c:80-90)  =>  c_hat[q] = 1 - exp(-2 pi i q / n); n == 1 => 0
        std::vector<double> ch[3];
        for (int a = 0; a < 3; ++a) {
            ch[a].assign(2 * (size_t)n[a], 0.0);
            if (n[a] > 1)
                for (int q = 0; q < n[a]; ++q) {
                    double re, im;
                    exact_root(q, n[a], &re, &im);
                    ch[a][2 * q] = 1.0 - re;
                    ch[a][2 * q + 1] = -im;
                }
        }
        return set_symbol_separable(ch[0].data(), ch[1].data(), ch[2].data(), lx, ly, lz);
    }

    int ensure_inv_table()
    {
        if (!inv_table) CPC_CUDA(cudaMalloc(&inv_table, sizeof(C) * nloc));
        return CPC_OK;
    }

    // Explicit eigenvalues (the Diag argument of solve_3D, FftLinearSolver_3D.c:166,174): this rank's z-slab of Diag.
    // The reference only ever builds separable tables (build_diag_mat_vec_3D, :136-164), so the table is first
    // tested for that structure -- Diag[k,j,i] = a[i] + b[j] + c[k], with a, b, c read off three of its lines --
    // over all N entries on the GPU.  If it holds (to 1e-13 relative) the plan keeps three 1-D tables instead of N
    // eigenvalues and the middle pass can take the recurrence form; otherwise 1 / (N Diag) is stored as a full table.
    int set_symbol_diag(const void *diag, int mem_kind) override
    {
        if (real) { set_error("cpc_set_symbol_diag: complex plans only (real plans take the transport / separable symbol)"); return CPC_ERR_UNSUPPORTED; }
        if (!diag) { set_error("null diag"); return CPC_ERR_ARG; }
        if (nc != 1) { set_error("cpc_set_symbol_diag needs ncomp == 1"); return CPC_ERR_ARG; }
        const long long cells = (long long)n[0] * n[1] * nzl;   // one eigenvalue per local cell
        const double2 *d = (const double2 *)diag;
        double2 *tmp = nullptr;
        int rc = CPC_OK;
        if (mem_kind == CPC_MEM_HOST) {
            CPC_CUDA(cudaMalloc(&tmp, sizeof(double2) * cells));
            CPC_CUDA(cudaMemcpyAsync(tmp, diag, sizeof(double2) * cells, cudaMemcpyHostToDevice, stream));
            h2d_bytes += sizeof(double2) * cells;
            d = tmp;
        }
        rc = diag_try_separable(d);
        if (rc == CPC_OK && symbol_kind != CPC_SYMBOL_SEPARABLE) rc = diag_to_table(d, cells);
        cudaStreamSynchronize(stream);
        if (tmp) cudaFree(tmp);
        return rc;
    }

    // On return symbol_kind == CPC_SYMBOL_SEPARABLE iff the device table d is separable (on every rank).
    int diag_try_separable(const double2 *d)
    {
        const int P = desc.nranks;
        const size_t plane = (size_t)n[0] * n[1];
        std::vector<double2> A(n[0]), B(n[1]), Cl(nzl), Call((size_t)n[2]);
        CPC_CUDA(cudaMemcpyAsync(A.data(), d, sizeof(double2) * n[0], cudaMemcpyDeviceToHost, stream));
        CPC_CUDA(cudaMemcpy2DAsync(B.data(), sizeof(double2), d, sizeof(double2) * n[0], sizeof(double2), n[1],
                                   cudaMemcpyDeviceToHost, stream));
        CPC_CUDA(cudaMemcpy2DAsync(Cl.data(), sizeof(double2), d, sizeof(double2) * plane, sizeof(double2), nzl,
                                   cudaMemcpyDeviceToHost, stream));
        CPC_CUDA(cudaStreamSynchronize(stream));
        if (P > 1) {            // the z line of Diag through (0, 0) from every rank's slab
            double2 *dz = nullptr;
            CPC_CUDA(cudaMalloc(&dz, sizeof(double2) * (size_t)n[2]));
            CPC_CUDA(cudaMemcpyAsync(dz + z0, Cl.data(), sizeof(double2) * nzl, cudaMemcpyHostToDevice, stream));
            int rc = dist_allgather(dist, dz + z0, dz, sizeof(double2) * nzl, stream);
            if (rc) { cudaFree(dz); return rc; }
            CPC_CUDA(cudaMemcpyAsync(Call.data(), dz, sizeof(double2) * n[2], cudaMemcpyDeviceToHost, stream));
            CPC_CUDA(cudaStreamSynchronize(stream));
            cudaFree(dz);
        } else {
            Call = Cl;
        }
        // a[i] - a[0], b[j] - b[0] + Diag[0,0,0], c[k] - c[0]: their sum is Diag wherever Diag is separable
        const double2 d00 = A[0], D000 = Call[0];
        std::vector<double2> h[3];
        h[0].resize(n[0]); h[1].resize(n[1]); h[2].resize(n[2]);
        for (int i = 0; i < n[0]; ++i) h[0][i] = make_double2(A[i].x - d00.x, A[i].y - d00.y);
        for (int j = 0; j < n[1]; ++j) h[1][j] = make_double2(B[j].x - d00.x + D000.x, B[j].y - d00.y + D000.y);
        for (int k = 0; k < n[2]; ++k) h[2][k] = make_double2(Call[k].x - D000.x, Call[k].y - D000.y);
        int rc = set_symbol_tables(h);
        if (rc) return rc;
        unsigned long long *res = nullptr, hres[2] = { 0, 0 };
        CPC_CUDA(cudaMalloc(&res, 2 * sizeof(unsigned long long) + sizeof(float)));
        CPC_CUDA(cudaMemsetAsync(res, 0, 2 * sizeof(unsigned long long) + sizeof(float), stream));
        diag_check_separable_kernel<<<1184, 256, 0, stream>>>(d, sym_tab64[0], sym_tab64[1], sym_tab64[2] + z0, n[0], n[1],
                                                              (long long)plane * nzl, res);
        ++launches;
        CPC_CUDA(cudaGetLastError());
        CPC_CUDA(cudaMemcpyAsync(hres, res, sizeof(hres), cudaMemcpyDeviceToHost, stream));
        CPC_CUDA(cudaStreamSynchronize(stream));
        double maxdiff, maxabs;
        memcpy(&maxdiff, &hres[0], 8);
        memcpy(&maxabs, &hres[1], 8);
        float bad = (maxdiff <= 1e-13 * (1.0 + maxabs)) ? 0.f : 1.f;
        if (P > 1) {            // every rank must take the same decision
            float *flag = (float *)(res + 2);
            CPC_CUDA(cudaMemcpyAsync(flag, &bad, sizeof(float), cudaMemcpyHostToDevice, stream));
            rc = dist_allreduce_sum_f32(dist, flag, 1, stream);
            if (rc) { cudaFree(res); return rc; }
            CPC_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(float), cudaMemcpyDeviceToHost, stream));
            CPC_CUDA(cudaStreamSynchronize(stream));
        }
        cudaFree(res);
        if (bad != 0.f) { symbol_kind = CPC_SYMBOL_NONE; zrec = false; }
        return CPC_OK;
    }

    // 1 / (N Diag) as a full table in the layout the fused z pass runs in (transposed for multi-rank plans)
    int diag_to_table(const double2 *d, long long cells)
    {
        int rc = ensure_inv_table();
        if (rc) return rc;
        if (desc.nranks == 1) {
            invert_table_kernel<T><<<1184, 256, 0, stream>>>(d, inv_table, cells, 1.0 / (double)ntot);
            ++launches;
            CPC_CUDA(cudaGetLastError());
        } else {
            if ((rc = ensure_transpose_buffers())) return rc;
            invert_table_kernel<T><<<1184, 256, 0, stream>>>(d, tbuf, cells, 1.0 / (double)ntot);
            slab_to_chunked_kernel<C><<<1184, 256, 0, stream>>>(tbuf, sendbuf, n[0], n[1], nyl, nzl);
            launches += 2;
            CPC_CUDA(cudaGetLastError());
            if ((rc = alltoall(sendbuf, inv_table))) return rc;
        }
        CPC_CUDA(cudaStreamSynchronize(stream));
        symbol_kind = CPC_SYMBOL_TABLE; zrec = false;
        return CPC_OK;
    }

    int set_symbol_first_column(const void *col, int mem_kind) override
    {
        if (nc != 1) { set_error("first-column symbol needs ncomp == 1"); return CPC_ERR_ARG; }
        if (real) { set_error("cpc_set_symbol_first_column: complex plans only"); return CPC_ERR_UNSUPPORTED; }
        if (!col) { set_error("null column"); return CPC_ERR_ARG; }
        int rc = ensure_inv_table();
        if (rc) return rc;
        // Lambda = FFT3(col) lands in the layout the fused pass runs in (transposed for multi-rank plans).
        rc = transform_impl((const C *)col, inv_table, mem_kind, -1, /*out_is_device=*/true);
        if (rc) return rc;
        // a column with entries on the axes only (the transport stencil) has a separable spectrum: keep three 1-D
        // tables then, and the recurrence form of the middle pass if it applies (fp64 single-rank plans)
        if (sizeof(T) == 8 && desc.nranks == 1) {
            if ((rc = diag_try_separable((const double2 *)inv_table))) return rc;
            if (symbol_kind == CPC_SYMBOL_SEPARABLE) return CPC_OK;
        }
        invert_inplace_kernel<T><<<1184, 256, 0, stream>>>(inv_table, nloc, 1.0 / (double)ntot);
        ++launches;
        CPC_CUDA(cudaGetLastError());
        CPC_CUDA(cudaStreamSynchronize(stream));
        symbol_kind = CPC_SYMBOL_TABLE; zrec = false;
        return CPC_OK;
    }

    int set_symbol_wave(double c0, double mx, double my, double mz) override
    {
        if (nc != 4) { set_error("wave symbol needs ncomp == 4"); return CPC_ERR_ARG; }
        wave_c0 = c0;
        // degenerate axes contribute nothing (their root table is [1]: sin = 0, 1 - cos = 0)
        wave_mu[0] = mx; wave_mu[1] = my; wave_mu[2] = mz;
        symbol_kind = CPC_SYMBOL_WAVE; zrec = false;
        return CPC_OK;
    }

    // a flag barrier that gave up waiting (dist.cu) leaves a mark: report it instead of returning wrong numbers
    int health() override
    {
        if (!dist.flag_barrier || !dist.timeout_flag) return CPC_OK;
        int t = 0;
        CPC_CUDA(cudaMemcpy(&t, dist.timeout_flag, sizeof(int), cudaMemcpyDeviceToHost));
        if (t) { set_error("a peer rank did not reach a barrier within 2 s (multi-rank apply aborted)"); return CPC_ERR_NCCL; }
        return CPC_OK;
    }

    int set_option(int option, long long value) override
    {
        switch (option) {
        case CPC_OPT_Z_RECURRENCE: zrec_off = (value == 0); return CPC_OK;
        case CPC_OPT_L2_CHUNK_BYTES: l2_chunk_bytes = value < 0 ? 0 : value; return CPC_OK;
        case CPC_OPT_CHAIN_STREAMS: chain_streams = value >= 2 ? 2 : 1; return CPC_OK;
        case CPC_OPT_Z_LINE_FORM: zline_mode = value < 0 ? -1 : (value ? 1 : 0); update_zrec_line(); return CPC_OK;
        default: set_error("cpc_set_option: unknown option %d", option); return CPC_ERR_ARG;
        }
    }

    int get_diag(void *diag, int mem_kind) override
    {
        if (desc.nranks != 1) { set_error("cpc_get_diag: single-rank plans only"); return CPC_ERR_UNSUPPORTED; }
        if (real) { set_error("cpc_get_diag: complex plans only"); return CPC_ERR_UNSUPPORTED; }
        if (symbol_kind != CPC_SYMBOL_SEPARABLE && symbol_kind != CPC_SYMBOL_TABLE) {
            set_error("cpc_get_diag: no scalar symbol set");
            return CPC_ERR_STATE;
        }
        double2 *d = (double2 *)diag, *tmp = nullptr;
        if (mem_kind == CPC_MEM_HOST) {
            CPC_CUDA(cudaMalloc(&tmp, sizeof(double2) * ntot));
            d = tmp;
        }
        if (symbol_kind == CPC_SYMBOL_SEPARABLE)
            diag_from_separable_kernel<<<1184, 256, 0, stream>>>(d, sym_tab64[0], sym_tab64[1], sym_tab64[2], n[0], n[1],
                                                                 ntot);
        else
            diag_from_invtable_kernel<T><<<1184, 256, 0, stream>>>(d, inv_table, ntot, 1.0 / (double)ntot);
        ++launches;
        CPC_CUDA(cudaGetLastError());
        if (tmp) {
            CPC_CUDA(cudaMemcpyAsync(diag, tmp, sizeof(double2) * ntot, cudaMemcpyDeviceToHost, stream));
            d2h_bytes += sizeof(double2) * ntot;
        }
        CPC_CUDA(cudaStreamSynchronize(stream));
        if (tmp) cudaFree(tmp);
        return CPC_OK;
    }

    // ------------------------------------------------------------------------------------ schedules
    int ensure_dbuf()
    {
        if (!dbuf) CPC_CUDA(cudaMalloc(&dbuf, sizeof(C) * nloc));
        return CPC_OK;
    }

    // ---- profiled applies: an event after every launch, durations summed per pass kind -------------------------
    int prof_begin(float *pass_ms)
    {
        prof_on = pass_ms != nullptr;
        prof_n = 0;
        prof_kind.clear();
        return prof_on ? prof_mark(-1) : CPC_OK;
    }
    int prof_mark(int kind)
    {
        if (!prof_on) return CPC_OK;
        if (prof_n == prof_ev.size()) {
            cudaEvent_t e;
            CPC_CUDA(cudaEventCreate(&e));
            prof_ev.push_back(e);
        }
        CPC_CUDA(cudaEventRecord(prof_ev[prof_n++], stream));
        prof_kind.push_back(kind);
        return CPC_OK;
    }
    int prof_end(float *pass_ms, int *npasses, int nkinds)
    {
        if (!prof_on) return CPC_OK;
        prof_on = false;
        CPC_CUDA(cudaEventSynchronize(prof_ev[prof_n - 1]));
        for (int i = 0; i < nkinds; ++i) pass_ms[i] = 0.f;
        for (size_t i = 1; i < prof_n; ++i) {
            float ms = 0.f;
            CPC_CUDA(cudaEventElapsedTime(&ms, prof_ev[i - 1], prof_ev[i]));
            if (prof_kind[i] >= 0 && prof_kind[i] < nkinds) pass_ms[prof_kind[i]] += ms;
        }
        if (npasses) *npasses = nkinds;
        return CPC_OK;
    }

    // ---- L2-chained x / y passes -----------------------------------------------------------------------------
    // planes per z-chunk; 0 = whole-array passes (nothing to chain, or the local array sits in L2 anyway)
    int chain_planes() const
    {
        if (l2_chunk_bytes <= 0 || n[0] == 1 || n[1] == 1) return 0;
        const long long plane = wx * n[1] * (long long)sizeof(C);
        if (plane * nzl <= 2 * l2_chunk_bytes) return 0;
        const long long p = l2_chunk_bytes / plane;
        return (int)(p < 1 ? 1 : p);
    }

    // Forward x and y passes (src -> x, then in place), or backward y and x passes (in place on x), over the local
    // planes.  With a chunk size the two passes of the pair run back to back on each z-chunk so that the second
    // finds the chunk in L2; `after(zb, zc)` (optional) is queued behind the pair of each chunk.
    // kinds: profile kinds of the two passes in launch order.
    template <typename After>
    int run_xy(bool forward, const C *src, C *x, const int kinds[2], After after)
    {
        int rc;
        const int cp = chain_planes();
        const int step = cp > 0 ? cp : nzl;
        const bool two = cp > 0 && chain_streams >= 2 && !prof_on;
        if (two) {
            CPC_CUDA(cudaEventRecord(fork_ev, stream));
            CPC_CUDA(cudaStreamWaitEvent(side_stream, fork_ev, 0));
        }
        int ci = 0;
        for (int zb = 0; zb < nzl; zb += step, ++ci) {
            const int zc = nzl - zb < step ? nzl - zb : step;
            cudaStream_t st = (two && (ci & 1)) ? side_stream : stream;
            if (forward) {
                const C *cur = src;
                for (int a = 0; a < 2; ++a) {
                    if (n[a] == 1) continue;
                    if ((rc = run_pass(a, MODE_FWD, cur, x, zb, zc, 0, st))) return rc;
                    cur = x;
                    if ((rc = prof_mark(kinds[a]))) return rc;
                }
                if (cur == src && src != x)        // both axes degenerate: the pair is a copy
                    CPC_CUDA(cudaMemcpyAsync(x + (long long)zb * n[1] * wx, src + (long long)zb * n[1] * wx,
                                             sizeof(C) * (size_t)zc * n[1] * wx, cudaMemcpyDeviceToDevice, st));
            } else {
                for (int a = 1; a >= 0; --a) {
                    if (n[a] == 1) continue;
                    if ((rc = run_pass(a, MODE_INV, x, x, zb, zc, 0, st))) return rc;
                    if ((rc = prof_mark(kinds[1 - a]))) return rc;
                }
            }
            if ((rc = after(zb, zc, st))) return rc;
        }
        if (two) {
            CPC_CUDA(cudaEventRecord(join_ev, side_stream));
            CPC_CUDA(cudaStreamWaitEvent(stream, join_ev, 0));
        }
        return CPC_OK;
    }
    static int no_after(int, int, cudaStream_t) { return CPC_OK; }

    // Single-rank apply on device pointers.  pass_ms != nullptr => per-pass durations (CUDA events).
    int apply_device_single(const C *b, C *x, float *pass_ms, int *npasses)
    {
        const int fm = fused_mode();
        int rc;
        // pass kinds in the order Fx, Fy, middle, By, Bx (degenerate axes have no pass)
        int k = 0, kf[2] = { -1, -1 }, kb[2] = { -1, -1 };
        for (int a = 0; a < 2; ++a)
            if (n[a] > 1) kf[a] = k++;
        const int kz = k++;
        for (int a = 1; a >= 0; --a)
            if (n[a] > 1) kb[1 - a] = k++;
        if ((rc = prof_begin(pass_ms))) return rc;
        const bool copy_first = (n[0] == 1 && n[1] == 1);
        if (!copy_first && (rc = run_xy(true, b, x, kf, no_after))) return rc;
        if ((rc = run_pass(2, fm, copy_first ? b : x, x, 0, nzl, 0, stream))) return rc;
        if ((rc = prof_mark(kz))) return rc;
        if (!copy_first && (rc = run_xy(false, x, x, kb, no_after))) return rc;
        return prof_end(pass_ms, npasses, k);
    }

    // Multi-rank (z-slab) apply: Fx, Fy(split store) | all-to-all | fused z | all-to-all | By(split load), Bx
    int alltoall(const C *send, C *recv)
    {
        const size_t chunk = sizeof(C) * (size_t)(nloc / desc.nranks);
        return dist_alltoall(dist, send, recv, chunk, stream);
    }

    // Multi-rank apply for a transport symbol: no transposes.  [Fx, Fy, end-value accumulation] z-chunk by z-chunk
    // (L2-chained), the carry exchange (zsolve.cuh), the z solve on the local slab from the exchanged carry-in,
    // [By, Bx] z-chunk by z-chunk.  Pass kinds: Fx, Fy, end values, carry exchange, z solve, By, Bx.
    // Can the carry exchange be pipelined over two halves of the columns?  (peer flags for two barrier groups, both
    // x and y transformed, the halves whole tiles of the y pass)
    bool xsplit_ok() const
    {
        return xsplit && xstream && !real && carry_p2p && dist.flag_barrier && n[0] > 1 && n[1] > 1 && l2_chunk_bytes <= 0 &&
               cfg[1].tx > 0 && n[0] % (2 * cfg[1].tx) == 0 && (long long)n[0] * n[1] >= 65536;
    }

    // The z-slab schedule with the carry exchange of one half of the columns hidden behind the work on the other:
    //   main stream:  Fx | Fy(A) END(A) | Fy(B) END(B) | wait A: solve(A) | wait B: solve(B) | By | Bx
    //   xstream:             wait END(A): barrier, owner(A), barrier | wait END(B): barrier, owner(B), barrier
    // A = columns [0, nx/2), B = the rest; END pushes its end values straight to the line owners.  The second sweep is
    // the thread-per-line kernel (it takes a column range as it is).  Barriers of xstream use their own flag group.
    int apply_device_zslab_xsplit(const C *b, C *x)
    {
        int rc;
        const long long L = (long long)n[0] * n[1];
        const int nxh = n[0] / 2;
        ZCarryPeers gp{}, zp{};
        for (int q = 0; q < desc.nranks; ++q) { gp.p[q] = (double2 *)peer_g[q]; zp.p[q] = (double2 *)peer_z[q]; }
        const int hgrid = (int)(((long long)nxh * n[1] + 255) / 256);
        const long long line0 = lsub * desc.rank;
        const long long cnt = line0 >= L ? 0 : (L - line0 < lsub ? L - line0 : lsub);
        if ((rc = run_pass(0, MODE_FWD, b, x, 0, nzl, 0, stream))) return rc;
        ZSolveArgs za[2];
        for (int h = 0; h < 2; ++h) {
            za[h] = zsolve_args();
            za[h].nline = nzl;
            za[h].xs0 = h * nxh;
            za[h].xsn = nxh;
            if ((rc = run_pass(1, MODE_FWD, x, x, 0, nzl, 0, stream, 0, h * nxh, nxh))) return rc;
            zs_end_accum_kernel<T><<<hgrid, 256, 0, stream>>>(x, L, n[0], 0, nzl, 0, end_trunc ? 1 : 0, ebuf, za[h], desc.rank, lsub, gp,
                                                              FlagSync{});
            ++launches;
            CPC_CUDA(cudaGetLastError());
            CPC_CUDA(cudaEventRecord(xs_ev[h], stream));
            CPC_CUDA(cudaStreamWaitEvent(xstream, xs_ev[h], 0));
            if ((rc = dist_barrier(dist, xstream, 1))) return rc;         // every rank's end values of this half have landed
            if (cnt > 0) {
                zs_carry_owner_kernel<<<(int)((cnt + 255) / 256), 256, 0, xstream>>>(gbuf, lsub, line0, cnt, n[0], nzl, desc.nranks,
                                                                                  desc.rank, 0, zp, za[h], FlagSync{}, FlagSync{});
                ++launches;
                CPC_CUDA(cudaGetLastError());
            }
            if ((rc = dist_barrier(dist, xstream, 1))) return rc;         // every carry-in of this half has landed
            CPC_CUDA(cudaEventRecord(xs_ev[2 + h], xstream));
        }
        for (int h = 0; h < 2; ++h) {
            CPC_CUDA(cudaStreamWaitEvent(stream, xs_ev[2 + h], 0));
            zs_dist_line_kernel<T><<<hgrid, 256, 0, stream>>>(x, x, L, n[0], nzl, za[h], FlagSync{});
            ++launches;
            CPC_CUDA(cudaGetLastError());
        }
        if ((rc = run_pass(1, MODE_INV, x, x, 0, nzl, 0, stream))) return rc;
        return run_pass(0, MODE_INV, x, x, 0, nzl, 0, stream);
    }

    // The z part of the z-slab schedule on array x (rows of wx complex numbers: nx, or the padded half spectrum of a
    // real plan), behind the forward y pass: end-value sweep, carry exchange, second sweep.  Profile kinds 2, 3, 4.
    int zslab_middle(C *x)
    {
        int rc;
        const long long L = wx * n[1];
        const int nxw = (int)wx;
        ZSolveArgs za = zsolve_args();
        za.nline = nzl;
        const int egrid = (int)((L + 255) / 256);
        ZCarryPeers gp{}, zp{};
        for (int q = 0; q < desc.nranks; ++q) { gp.p[q] = (double2 *)peer_g[q]; zp.p[q] = (double2 *)peer_z[q]; }
        // With peers mapped and CPC_FUSED_SYNC the exchange needs no separate barrier launches: the end-value sweep
        // signals "landed at the owners" when its last block finishes, the owner kernel waits for every rank's signal,
        // pushes the carry-ins and signals in turn, and the second sweep waits for that (FlagSync, zsolve.cuh).
        const bool fused = carry_p2p && dist.flag_barrier && fused_sync;
        const FlagSync none{};
        const FlagSync s_end = fused ? make_sync(0, ++dist.epoch[0]) : none;
        const FlagSync s_own = fused ? make_sync(1, ++dist.epoch[1]) : none;
        // the end values go straight to the ranks that own the lines (peer stores) when peers are mapped
        zs_end_accum_kernel<T><<<egrid, 256, 0, stream>>>(x, L, nxw, 0, nzl, 0, end_trunc ? 1 : 0, ebuf, za,
                                                          carry_p2p ? desc.rank : -1, lsub, gp, s_end);
        ++launches;
        CPC_CUDA(cudaGetLastError());
        if ((rc = prof_mark(2))) return rc;
        const int cgrid = 148 * 8;
        if (carry_p2p) {
            if (!fused && (rc = dist_barrier(dist, stream))) return rc;  // every rank's end values have landed
            const long long line0 = lsub * desc.rank;
            const long long cnt = line0 >= L ? 0 : (L - line0 < lsub ? L - line0 : lsub);
            // (launched even without lines of its own: with fused signals the peers wait for this rank's)
            zs_carry_owner_kernel<<<(int)((cnt + 255) / 256 > 0 ? (cnt + 255) / 256 : 1), 256, 0, stream>>>(
                gbuf, lsub, line0, cnt, nxw, nzl, desc.nranks, desc.rank, 0, zp, za, s_end, s_own);
            ++launches;
            CPC_CUDA(cudaGetLastError());
            if (!fused && (rc = dist_barrier(dist, stream))) return rc;  // every line's carry-in has landed
        } else {
            if ((rc = dist_allgather(dist, ebuf, gbuf, sizeof(double2) * (size_t)L, stream))) return rc;
            ZCarryPeers zl{};
            zl.p[0] = zinbuf;
            zs_carry_owner_kernel<<<cgrid, 256, 0, stream>>>(gbuf, L, 0, L, nxw, nzl, desc.nranks, desc.rank, 1, zl, za, none, none);
            ++launches;
            CPC_CUDA(cudaGetLastError());
        }
        if ((rc = prof_mark(3))) return rc;
        if (zslab_e > 0 && !zslab_line) {
            long long off = 0;
            const PassGeom g = make_geom(2, 128 / (int)sizeof(C), 0, nzl, 0, &off);
            launch_zsolve_e(zslab_e, ZS_DIST, false, nzl, g.ntiles, stream, x, x, g, s_own);
        } else {
            zs_dist_line_kernel<T><<<egrid, 256, 0, stream>>>(x, x, L, nxw, nzl, za, s_own);
        }
        ++launches;
        CPC_CUDA(cudaGetLastError());
        return prof_mark(4);
    }

    int apply_device_zslab(const C *b, C *x, float *pass_ms, int *npasses)
    {
        int rc;
        // (per-pass profiling keeps the serial schedule: its pass kinds would overlap in the pipelined one)
        if (!pass_ms && xsplit_ok()) return apply_device_zslab_xsplit(b, x);
        if ((rc = prof_begin(pass_ms))) return rc;
        const int kf[2] = { 0, 1 }, kb[2] = { 5, 6 };
        if ((rc = run_xy(true, b, x, kf, no_after))) return rc;
        if ((rc = zslab_middle(x))) return rc;
        if ((rc = run_xy(false, x, x, kb, no_after))) return rc;
        return prof_end(pass_ms, npasses, 7);
    }

    int apply_device_dist(const C *b, C *x, float *pass_ms, int *npasses)
    {
        if (use_zslab()) return apply_device_zslab(b, x, pass_ms, npasses);
        const int fm = fused_mode();
        int np = 0, rc;
        if ((rc = ensure_transpose_buffers())) return rc;
        if ((rc = prof_begin(pass_ms))) return rc;
        const C *cur = b;
        if (n[0] > 1) {
            if ((rc = run_pass(0, MODE_FWD, cur, x, 0, nzl, 0, stream))) return rc;
            cur = x;
            if ((rc = prof_mark(np++))) return rc;
        }
        if (p2p) {
            // transposes fused into the producing kernels: stores go straight to the owning rank over NVLink.
            // Leading barrier: no peer may still be reading its buffers from an earlier call when pushes start.
            if ((rc = dist_barrier(dist, stream))) return rc;
            if ((rc = run_pass(1, MODE_FWD, cur, tbuf, 0, nzl, 0, stream, 3))) return rc;    // Fy -> peers' y-slabs
            if ((rc = prof_mark(np++))) return rc;
            if ((rc = dist_barrier(dist, stream))) return rc;                                // all pushes have landed
            if ((rc = prof_mark(np++))) return rc;
            if ((rc = run_pass(2, fm, tbuf, sendbuf, 0, n[2], 1, stream, 4))) return rc;     // fused z -> peers' z-slabs
            if ((rc = prof_mark(np++))) return rc;
            if ((rc = dist_barrier(dist, stream))) return rc;
            if ((rc = prof_mark(np++))) return rc;
        } else {
            if ((rc = run_pass(1, MODE_FWD, cur, sendbuf, 0, nzl, 0, stream, 1))) return rc; // Fy, chunked store
            if ((rc = prof_mark(np++))) return rc;
            if ((rc = alltoall(sendbuf, tbuf))) return rc;                                   // z-slab -> y-slab
            if ((rc = prof_mark(np++))) return rc;
            if ((rc = run_pass(2, fm, tbuf, tbuf, 0, n[2], 1, stream))) return rc;           // Fz . 1/(N Lambda) . Bz
            if ((rc = prof_mark(np++))) return rc;
            if ((rc = alltoall(tbuf, sendbuf))) return rc;                                   // y-slab -> z-slab
            if ((rc = prof_mark(np++))) return rc;
        }
        if ((rc = run_pass(1, MODE_INV, sendbuf, x, 0, nzl, 0, stream, 2))) return rc;       // By, chunked load
        if ((rc = prof_mark(np++))) return rc;
        if (n[0] > 1) {
            if ((rc = run_pass(0, MODE_INV, x, x, 0, nzl, 0, stream))) return rc;
            if ((rc = prof_mark(np++))) return rc;
        }
        return prof_end(pass_ms, npasses, np);
    }

    // forward: in = z-slab, out = transposed (all z, local y range); backward: the reverse
    int transform_device_dist(const C *in, C *out, int dir)
    {
        int rc;
        if ((rc = ensure_transpose_buffers())) return rc;
        if (dir < 0) {
            const C *cur = in;
            if (n[0] > 1) {
                if ((rc = run_pass(0, MODE_FWD, cur, tbuf, 0, nzl, 0, stream))) return rc;
                cur = tbuf;
            }
            if ((rc = run_pass(1, MODE_FWD, cur, sendbuf, 0, nzl, 0, stream, 1))) return rc;
            if ((rc = alltoall(sendbuf, tbuf))) return rc;
            if ((rc = run_pass(2, MODE_FWD, tbuf, out, 0, n[2], 1, stream))) return rc;
        } else {
            if ((rc = run_pass(2, MODE_INV, in, tbuf, 0, n[2], 1, stream))) return rc;
            if ((rc = alltoall(tbuf, sendbuf))) return rc;
            if ((rc = run_pass(1, MODE_INV, sendbuf, out, 0, nzl, 0, stream, 2))) return rc;
            if (n[0] > 1 && (rc = run_pass(0, MODE_INV, out, out, 0, nzl, 0, stream))) return rc;
        }
        return CPC_OK;
    }

    // Host pointers, single rank: z-chunked pipeline so the PCIe copies overlap the x/y passes.
    static bool is_pageable(const void *p)
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
        return at.type == cudaMemoryTypeUnregistered;
    }

    int ensure_stage()
    {
        if (stage[0]) return CPC_OK;
        for (int i = 0; i < 4; ++i) {
            CPC_CUDA(cudaHostAlloc((void **)&stage[i], kStageBytes, cudaHostAllocDefault));
            CPC_CUDA(cudaEventCreateWithFlags(&stage_ev[i], cudaEventDisableTiming));
        }
        const unsigned hc = std::thread::hardware_concurrency();
        host_threads = hc >= 16 ? 8 : (hc >= 4 ? (int)hc / 2 : 1);
        return CPC_OK;
    }

    // memcpy split over the host threads (a single thread moves ~10 GB/s, PCIe 5 x16 takes 55)
    void par_memcpy(void *dst, const void *src, size_t bytes) const
    {
        const int nth = bytes >= (8u << 20) ? host_threads : 1;
        if (nth <= 1) { memcpy(dst, src, bytes); return; }
        std::vector<std::thread> th;
        const size_t per = ((bytes + nth - 1) / nth + 4095) & ~(size_t)4095;
        for (int i = 0; i < nth; ++i) {
            const size_t o = per * i;
            if (o >= bytes) break;
            const size_t nb = bytes - o < per ? bytes - o : per;
            th.emplace_back([=] { memcpy((char *)dst + o, (const char *)src + o, nb); });
        }
        for (auto &w : th) w.join();
    }

    // pageable host -> device through the two bounce buffers, queued on copy_stream
    int h2d_staged(C *dst, const C *src, size_t bytes)
    {
        for (size_t o = 0, i = 0; o < bytes; o += kStageBytes, ++i) {
            const size_t nb = bytes - o < kStageBytes ? bytes - o : kStageBytes;
            const int s = (int)(i & 1);
            CPC_CUDA(cudaEventSynchronize(stage_ev[s]));                  // the buffer's previous transfer is done
            par_memcpy(stage[s], (const char *)src + o, nb);
            CPC_CUDA(cudaMemcpyAsync((char *)dst + o, stage[s], nb, cudaMemcpyHostToDevice, copy_stream));
            CPC_CUDA(cudaEventRecord(stage_ev[s], copy_stream));
        }
        return CPC_OK;
    }

    // device -> pageable host: the transfer of piece i + 1 runs while the host threads drain piece i
    int d2h_staged(C *dst, const C *src, size_t bytes)
    {
        size_t prev_o = 0, prev_nb = 0;
        int prev_s = -1;
        for (size_t o = 0, i = 0; o < bytes; o += kStageBytes, ++i) {
            const size_t nb = bytes - o < kStageBytes ? bytes - o : kStageBytes;
            const int s = 2 + (int)(i & 1);
            CPC_CUDA(cudaMemcpyAsync(stage[s], (const char *)src + o, nb, cudaMemcpyDeviceToHost, copy_stream));
            CPC_CUDA(cudaEventRecord(stage_ev[s], copy_stream));
            if (prev_s >= 0) {
                CPC_CUDA(cudaEventSynchronize(stage_ev[prev_s]));
                par_memcpy((char *)dst + prev_o, stage[prev_s], prev_nb);
            }
            prev_o = o; prev_nb = nb; prev_s = s;
        }
        if (prev_s >= 0) {
            CPC_CUDA(cudaEventSynchronize(stage_ev[prev_s]));
            par_memcpy((char *)dst + prev_o, stage[prev_s], prev_nb);
        }
        return CPC_OK;
    }

    // Host pointers, single rank: z-chunked pipeline so the PCIe copies overlap the x/y passes.
    int apply_host_single(const C *b, C *x)
    {
        int rc = ensure_dbuf();
        if (rc) return rc;
        const int fm = fused_mode();
        const long long plane = (long long)n[0] * n[1] * nc;
        const bool big = plane * nzl * (long long)sizeof(C) >= (4ll << 20);
        const bool stage_b = big && is_pageable(b), stage_x = big && is_pageable(x);
        if ((stage_b || stage_x) && (rc = ensure_stage())) return rc;
        int nchunk = nzl < 8 ? nzl : 8;
        if (!big) nchunk = 1;
        while ((int)chunk_ev.size() < 2 * nchunk) {
            cudaEvent_t e;
            CPC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            chunk_ev.push_back(e);
        }
        // make the copy stream wait for whatever the caller queued on the plan stream before
        CPC_CUDA(cudaEventRecord(chunk_ev[0], stream));
        CPC_CUDA(cudaStreamWaitEvent(copy_stream, chunk_ev[0], 0));
        for (int c = 0; c < nchunk; ++c) {
            const int zb = (int)((long long)nzl * c / nchunk), ze = (int)((long long)nzl * (c + 1) / nchunk);
            const long long off = plane * zb, cnt = plane * (ze - zb);
            if (stage_b) {
                if ((rc = h2d_staged(dbuf + off, b + off, sizeof(C) * cnt))) return rc;
            } else {
                CPC_CUDA(cudaMemcpyAsync(dbuf + off, b + off, sizeof(C) * cnt, cudaMemcpyHostToDevice, copy_stream));
            }
            h2d_bytes += sizeof(C) * cnt;
            CPC_CUDA(cudaEventRecord(chunk_ev[c], copy_stream));
            CPC_CUDA(cudaStreamWaitEvent(stream, chunk_ev[c], 0));
            for (int a = 0; a < 2; ++a)
                if (n[a] > 1 && (rc = run_pass(a, MODE_FWD, dbuf, dbuf, zb, ze - zb, 0, stream))) return rc;
        }
        if ((rc = run_pass(2, fm, dbuf, dbuf, 0, nzl, 0, stream))) return rc;
        for (int c = 0; c < nchunk; ++c) {
            const int zb = (int)((long long)nzl * c / nchunk), ze = (int)((long long)nzl * (c + 1) / nchunk);
            const long long off = plane * zb, cnt = plane * (ze - zb);
            for (int a = 1; a >= 0; --a)
                if (n[a] > 1 && (rc = run_pass(a, MODE_INV, dbuf, dbuf, zb, ze - zb, 0, stream))) return rc;
            CPC_CUDA(cudaEventRecord(chunk_ev[nchunk + c], stream));
            CPC_CUDA(cudaStreamWaitEvent(copy_stream, chunk_ev[nchunk + c], 0));
            if (stage_x) {
                if ((rc = d2h_staged(x + off, dbuf + off, sizeof(C) * cnt))) return rc;
            } else {
                CPC_CUDA(cudaMemcpyAsync(x + off, dbuf + off, sizeof(C) * cnt, cudaMemcpyDeviceToHost, copy_stream));
            }
            d2h_bytes += sizeof(C) * cnt;
        }
        CPC_CUDA(cudaStreamSynchronize(copy_stream));
        return CPC_OK;
    }

    // Real plan, device pointers: r2c x | Fy | fused z | By | c2r x, the middle three on the half spectrum; the
    // (r2c, Fy) and (By, c2r) pairs run z-chunk by z-chunk like the complex plans' (L2-chained).
    int apply_device_real(const T *b, T *x, float *pass_ms, int *npasses)
    {
        int rc;
        if ((rc = prof_begin(pass_ms))) return rc;
        const long long N = (long long)n[0] * n[1] * nzl;
        if (desc.nranks > 1 && !use_zslab()) {
            set_error("multi-rank real-scalar plans run the transpose-free z-slab schedule only (transport / upwind-z separable symbol)");
            return CPC_ERR_UNSUPPORTED;
        }
        if (real_promote) {
            real_to_complex_kernel<T><<<1184, 256, 0, stream>>>(b, work, N);
            ++launches;
            const bool on = prof_on;
            prof_on = false;
            rc = apply_device_single(work, work, nullptr, nullptr);
            prof_on = on;
            if (rc) return rc;
            complex_to_real_kernel<T><<<1184, 256, 0, stream>>>(work, x, N);
            ++launches;
            CPC_CUDA(cudaGetLastError());
            if ((rc = prof_mark(0))) return rc;
            return prof_end(pass_ms, npasses, 1);
        }
        int k = 0;
        const bool multi = desc.nranks > 1;
        if (multi && prof_on) { prof_on = false; set_error("profiled apply: single-rank real plans only"); return CPC_ERR_UNSUPPORTED; }
        const int kx0 = k++, ky0 = n[1] > 1 ? k++ : -1, kz = k++, ky1 = n[1] > 1 ? k++ : -1, kx1 = k++;
        const int cp = chain_planes();
        const int step = cp > 0 ? cp : nzl;
        for (int zb = 0; zb < nzl; zb += step) {
            const int zc = nzl - zb < step ? nzl - zb : step;
            if ((rc = run_real_x(true, b, work, zb, zc, stream))) return rc;
            if ((rc = prof_mark(kx0))) return rc;
            if (n[1] > 1) {
                if ((rc = run_pass(1, MODE_FWD, work, work, zb, zc, 0, stream))) return rc;
                if ((rc = prof_mark(ky0))) return rc;
            }
        }
        if (multi) {
            if ((rc = zslab_middle(work))) return rc;        // half-spectrum lines: carries exchanged between the slabs
        } else {
            if ((rc = run_pass(2, fused_mode(), work, work, 0, n[2], 0, stream))) return rc;
            if ((rc = prof_mark(kz))) return rc;
        }
        for (int zb = 0; zb < nzl; zb += step) {
            const int zc = nzl - zb < step ? nzl - zb : step;
            if (n[1] > 1) {
                if ((rc = run_pass(1, MODE_INV, work, work, zb, zc, 0, stream))) return rc;
                if ((rc = prof_mark(ky1))) return rc;
            }
            if ((rc = run_real_x(false, work, x, zb, zc, stream))) return rc;
            if ((rc = prof_mark(kx1))) return rc;
        }
        return prof_end(pass_ms, npasses, k);
    }

    int apply_real(const void *b, void *x, int mem_kind, float *pass_ms, int *npasses)
    {
        const size_t bytes = sizeof(T) * (size_t)n[0] * n[1] * nzl;
        if (mem_kind == CPC_MEM_DEVICE) return apply_device_real((const T *)b, (T *)x, pass_ms, npasses);
        int rc = ensure_dbuf();        // nloc complex >= N reals
        if (rc) return rc;
        CPC_CUDA(cudaMemcpyAsync(dbuf, b, bytes, cudaMemcpyHostToDevice, stream));
        h2d_bytes += bytes;
        if ((rc = apply_device_real((const T *)dbuf, (T *)dbuf, nullptr, nullptr))) return rc;
        CPC_CUDA(cudaMemcpyAsync(x, dbuf, bytes, cudaMemcpyDeviceToHost, stream));
        d2h_bytes += bytes;
        CPC_CUDA(cudaStreamSynchronize(stream));
        return CPC_OK;
    }

    int apply(const void *b, void *x, int mem_kind, float *pass_ms, int *npasses) override
    {
        if (fused_mode() < 0) { set_error("cpc_apply: no symbol set (call cpc_set_symbol_* first)"); return CPC_ERR_STATE; }
        if (!b || !x) { set_error("cpc_apply: null pointer"); return CPC_ERR_ARG; }
        CPC_CUDA(cudaSetDevice(device));
        if (real) {
            if (mem_kind != CPC_MEM_DEVICE && mem_kind != CPC_MEM_HOST) { set_error("bad mem_kind %d", mem_kind); return CPC_ERR_ARG; }
            if (pass_ms && mem_kind != CPC_MEM_DEVICE) { set_error("profiled apply needs device pointers"); return CPC_ERR_ARG; }
            return apply_real(b, x, mem_kind, pass_ms, npasses);
        }
        if (mem_kind == CPC_MEM_DEVICE) {
            if (desc.nranks == 1) return apply_device_single((const C *)b, (C *)x, pass_ms, npasses);
            return apply_device_dist((const C *)b, (C *)x, pass_ms, npasses);
        }
        if (mem_kind != CPC_MEM_HOST) { set_error("bad mem_kind %d", mem_kind); return CPC_ERR_ARG; }
        if (pass_ms) { set_error("profiled apply needs device pointers"); return CPC_ERR_ARG; }
        if (desc.nranks == 1) return apply_host_single((const C *)b, (C *)x);
        // multi-rank host pointers: stage the local slab
        int rc = ensure_dbuf();
        if (rc) return rc;
        CPC_CUDA(cudaMemcpyAsync(dbuf, b, sizeof(C) * nloc, cudaMemcpyHostToDevice, stream));
        h2d_bytes += sizeof(C) * nloc;
        if ((rc = apply_device_dist(dbuf, dbuf, nullptr, nullptr))) return rc;
        CPC_CUDA(cudaMemcpyAsync(x, dbuf, sizeof(C) * nloc, cudaMemcpyDeviceToHost, stream));
        d2h_bytes += sizeof(C) * nloc;
        CPC_CUDA(cudaStreamSynchronize(stream));
        return CPC_OK;
    }

    // Plain transforms (MatMult / MatMultTranspose of the reference, FftLinearSolver_3D.c:170,180)
    int transform_impl(const C *in, C *out, int mem_kind, int dir, bool out_is_device)
    {
        CPC_CUDA(cudaSetDevice(device));
        const C *din = in;
        C *dout = out;
        int rc;
        if (mem_kind == CPC_MEM_HOST) {
            if ((rc = ensure_dbuf())) return rc;
            CPC_CUDA(cudaMemcpyAsync(dbuf, in, sizeof(C) * nloc, cudaMemcpyHostToDevice, stream));
            h2d_bytes += sizeof(C) * nloc;
            din = dbuf;
            if (!out_is_device) dout = dbuf;
        }
        if (desc.nranks > 1) {
            if ((rc = transform_device_dist(din, dout, dir))) return rc;
        } else {
            const int mode = dir < 0 ? MODE_FWD : MODE_INV;
            const C *cur = din;
            int done = 0;
            for (int a = 0; a < 3; ++a) {
                if (n[a] == 1) continue;
                if ((rc = run_pass(a, mode, cur, dout, 0, nzl, 0, stream))) return rc;
                cur = dout;
                ++done;
            }
            if (!done && cur != dout)
                CPC_CUDA(cudaMemcpyAsync(dout, cur, sizeof(C) * nloc, cudaMemcpyDeviceToDevice, stream));
        }
        if (mem_kind == CPC_MEM_HOST && !out_is_device) {
            CPC_CUDA(cudaMemcpyAsync(out, dbuf, sizeof(C) * nloc, cudaMemcpyDeviceToHost, stream));
            d2h_bytes += sizeof(C) * nloc;
            CPC_CUDA(cudaStreamSynchronize(stream));
        }
        return CPC_OK;
    }

    int transform(const void *in, void *out, int mem_kind, int dir) override
    {
        if (real) { set_error("cpc_forward / cpc_inverse: complex plans only (real plans expose cpc_apply)"); return CPC_ERR_UNSUPPORTED; }
        if (!in || !out) { set_error("null pointer"); return CPC_ERR_ARG; }
        if (mem_kind != CPC_MEM_DEVICE && mem_kind != CPC_MEM_HOST) { set_error("bad mem_kind"); return CPC_ERR_ARG; }
        return transform_impl((const C *)in, (C *)out, mem_kind, dir, false);
    }

    void free_projection()
    {
        void *ptrs[] = { p_rowptr, pt_rowptr, p_colidx, pt_colidx, p_val, pt_val, pbuf, pio };
        for (void *q : ptrs)
            if (q) cudaFree(q);
        p_rowptr = pt_rowptr = nullptr; p_colidx = pt_colidx = nullptr; p_val = pt_val = nullptr; pbuf = pio = nullptr;
        proj_cols = 0;
    }

    int set_projection(int64_t cols, const int64_t *rowptr, const int32_t *colidx, const double *val) override
    {
        if (real || nc != 1 || desc.nranks != 1) { set_error("cpc_set_projection: complex scalar single-rank plans only"); return CPC_ERR_UNSUPPORTED; }
        const long long rows = ntot;
        const long long nnz = rowptr[rows];
        if (rowptr[0] != 0 || nnz < 0) { set_error("cpc_set_projection: malformed rowptr"); return CPC_ERR_ARG; }
        for (long long p = 0; p < nnz; ++p)
            if (colidx[p] < 0 || colidx[p] >= cols) { set_error("cpc_set_projection: column index %d out of range", colidx[p]); return CPC_ERR_ARG; }
        // transpose on the host (counting sort by column)
        std::vector<long long> trp((size_t)cols + 1, 0);
        for (long long p = 0; p < nnz; ++p) ++trp[(size_t)colidx[p] + 1];
        for (long long c = 0; c < cols; ++c) trp[(size_t)c + 1] += trp[(size_t)c];
        std::vector<int> tci((size_t)(nnz > 0 ? nnz : 1));
        std::vector<double> tv((size_t)(nnz > 0 ? nnz : 1));
        std::vector<long long> fill(trp.begin(), trp.end() - 1);
        for (long long i = 0; i < rows; ++i)
            for (long long p = rowptr[i]; p < rowptr[i + 1]; ++p) {
                const long long q = fill[(size_t)colidx[p]]++;
                tci[(size_t)q] = (int)i;
                tv[(size_t)q] = val[p];
            }
        free_projection();
        const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
        CPC_CUDA(cudaMalloc(&p_rowptr, sizeof(long long) * (size_t)(rows + 1)));
        CPC_CUDA(cudaMalloc(&pt_rowptr, sizeof(long long) * (size_t)(cols + 1)));
        CPC_CUDA(cudaMalloc(&p_colidx, sizeof(int) * nz));
        CPC_CUDA(cudaMalloc(&pt_colidx, sizeof(int) * nz));
        CPC_CUDA(cudaMalloc(&p_val, sizeof(double) * nz));
        CPC_CUDA(cudaMalloc(&pt_val, sizeof(double) * nz));
        CPC_CUDA(cudaMalloc(&pbuf, sizeof(C) * (size_t)rows));
        CPC_CUDA(cudaMalloc(&pio, sizeof(C) * (size_t)cols));
        CPC_CUDA(cudaMemcpy(p_rowptr, rowptr, sizeof(long long) * (size_t)(rows + 1), cudaMemcpyHostToDevice));
        CPC_CUDA(cudaMemcpy(pt_rowptr, trp.data(), sizeof(long long) * (size_t)(cols + 1), cudaMemcpyHostToDevice));
        CPC_CUDA(cudaMemcpy(p_colidx, colidx, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice));
        CPC_CUDA(cudaMemcpy(pt_colidx, tci.data(), sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice));
        CPC_CUDA(cudaMemcpy(p_val, val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice));
        CPC_CUDA(cudaMemcpy(pt_val, tv.data(), sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice));
        proj_cols = cols;
        return CPC_OK;
    }

    // x = P^T solve_3D(P b): reference applyFFT3DPrecTransport (PCSHELLFft_3D.cxx:10-24) plus the back-projection
    int apply_projected(const void *b, void *x, int mem_kind) override
    {
        if (!proj_cols) { set_error("cpc_apply_projected: no projection set"); return CPC_ERR_STATE; }
        if (fused_mode() < 0) { set_error("cpc_apply_projected: no symbol set"); return CPC_ERR_STATE; }
        const C *bin = (const C *)b;
        C *xout = (C *)x;
        if (mem_kind == CPC_MEM_HOST) {
            CPC_CUDA(cudaMemcpyAsync(pio, b, sizeof(C) * (size_t)proj_cols, cudaMemcpyHostToDevice, stream));
            h2d_bytes += sizeof(C) * (size_t)proj_cols;
            bin = pio;
            xout = pio;
        }
        const int grid_r = (int)((ntot + 255) / 256 < 148 * 16 ? (ntot + 255) / 256 : 148 * 16);
        csr_spmv_kernel<T><<<grid_r > 0 ? grid_r : 1, 256, 0, stream>>>(ntot, p_rowptr, p_colidx, p_val, bin, pbuf);
        ++launches;
        CPC_CUDA(cudaGetLastError());
        int rc = apply_device_single(pbuf, pbuf, nullptr, nullptr);
        if (rc) return rc;
        const int grid_c = (int)((proj_cols + 255) / 256 < 148 * 16 ? (proj_cols + 255) / 256 : 148 * 16);
        csr_spmv_kernel<T><<<grid_c > 0 ? grid_c : 1, 256, 0, stream>>>(proj_cols, pt_rowptr, pt_colidx, pt_val, pbuf, xout);
        ++launches;
        CPC_CUDA(cudaGetLastError());
        if (mem_kind == CPC_MEM_HOST) {
            CPC_CUDA(cudaMemcpyAsync(x, pio, sizeof(C) * (size_t)proj_cols, cudaMemcpyDeviceToHost, stream));
            d2h_bytes += sizeof(C) * (size_t)proj_cols;
            CPC_CUDA(cudaStreamSynchronize(stream));
        }
        return CPC_OK;
    }

    int get_info(cpc_plan_info *info) override
    {
        memset(info, 0, sizeof(*info));
        info->nx = n[0]; info->ny = n[1]; info->nz = n[2];
        info->ncomp = nc; info->dtype = desc.dtype;
        info->nranks = desc.nranks; info->rank = desc.rank;
        info->symbol_kind = symbol_kind;
        info->passes_per_apply = 1 + 2 * ((n[0] > 1) + (n[1] > 1));
        // before the transposing schedule's buffers exist (first use) report what it will try
        info->dist_mode = desc.nranks == 1 ? 0 : (use_zslab() ? 3 : ((tbuf_tried ? p2p : want_p2p) ? 2 : 1));
        for (int a = 0; a < 3; ++a) info->fast_path[a] = cfg[a].fast ? 1 : 0;
        if (desc.nranks > 1 ? use_zslab()
                            : (symbol_kind == CPC_SYMBOL_SEPARABLE && zrec && !zrec_off && (zrec_e > 0 || zrec_line) && nc == 1))
            info->fast_path[2] = 2;
        info->local_elems = nloc;
        info->bytes_per_apply_alg = 5ll * 2 * nloc * (long long)(real ? sizeof(T) : sizeof(C));
        info->kernel_launches = launches;
        info->h2d_bytes = h2d_bytes;
        info->d2h_bytes = d2h_bytes;
        return CPC_OK;
    }
};

}  // namespace cpc
