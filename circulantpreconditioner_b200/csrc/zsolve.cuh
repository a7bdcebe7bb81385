// zsolve.cuh -- the middle pass for the transport symbol without any FFT along z.
//
// After the x and y transforms every z line (fixed kx, ky) is an independent 1-D circulant system.  With the
// reference's upwind column c_z = [1, -1, 0, ...] (build_transport_col, src/FftLinearSolver_3D.c:80-90) and
//   alpha = 1 + lambda_x c_x_hat[kx] + lambda_y c_y_hat[ky]            (build_diag_mat_vec_3D, :146-157)
// the z part of  F^H (F b ./ Diag)  (solve_3D, :170-184) is exactly the solution of the cyclic bidiagonal system
//   (alpha + lambda_z) x_k - lambda_z x_{k-1} = b_k ,   k = 0 .. nz-1,   x_{-1} = x_{nz-1},
// because Diag[k] = alpha + lambda_z (1 - exp(-2 pi i k / nz)) is that circulant's spectrum.  With
// r = 1 / (alpha + lambda_z) and c = lambda_z r (|c| < 1 since Re alpha >= 1 and lambda_z >= 0):
//   y_k = c y_{k-1} + b_k ,  x_k = r y_k ,  and the cyclic closure  y_k = sum_{m < nz} c^m b_{k-m} / (1 - c^nz).
// That is a first-order linear recurrence: ~20 flops per point instead of two 512-point FFTs plus a division
// (~70 fp64 instructions per point in fft_r2x_kernel's fused mode), so the pass becomes purely HBM-bound.
// It is the same linear operator as forward-z FFT, divide, backward-z FFT; the results agree to rounding
// (7e-15 relative at lambda = 55.56, nz = 512; the plan only takes this path for 0 <= lambda_z <= 4096).
//
// Parallel form: a tile is TX neighbouring z lines (TX lanes x 16 B = one 128-byte row, as in fft_pass.cuh).
// Thread (l, s) owns E consecutive points [sE, (s+1)E) of line l:
//   1. local recurrence from a zero carry-in, in independent blocks of 4 (or 5) points        (E complex FMAs)
//   2. carries between segments: an inclusive scan over the QW = 32/TX segments of a warp with two
//      shuffles, the warp aggregates through shared memory, and a Horner sum over the NW warps that also closes
//      the cycle (factor 1 / (1 - c^nz))                                       (one block barrier)
//   3. x_k = r (y_k + c^(k - sE + 1) carry) / (nx ny), stored straight to HBM (or pushed to a peer, GEN builds).
//
// z-slab plans (one process per GPU) need NO transpose for this pass: a recurrence only hands a carry from one slab to
// the next.  The value at the end of every local line from a zero carry-in is accumulated plane by plane over the
// planes whose weight can still matter (zs_end_accum_kernel, a read-only sweep right behind the forward y pass), the
// carries are exchanged -- each rank owns 1/P of the (kx, ky) lines, gathers their P end values through peer stores,
// closes the cycle over the ranks and pushes one carry-in per line back to every rank (zs_carry_owner_kernel; 2 x 3.5 MB
// over NVLink per rank at 512^2 and 8 ranks instead of two 235 MB all-to-alls) -- and the second sweep (ZS_DIST)
// solves the local lines from that carry-in.
//
// All arithmetic of this file is fp64 whatever the storage type: the decay 1 - c = 1 / (alpha + lambda_z) carries a
// relative rounding error of eps (1 + lambda_z), which in fp32 would cost complex64 / float32 plans 2-3 digits
// against the FFT form at lambda_z = 55 .. 4096.  The pass stays HBM-bound either way.
#pragma once
#include "fft_pass.cuh"

namespace cpc {

struct ZSolveArgs {
    const double2 *ax, *ay;       // lambda_x c_x_hat[kx]  and  1 + lambda_y c_y_hat[ky]   (the plan's fp64 symbol tables)
    double lz;                    // lambda_z
    double scale;                 // 1 / (nx ny): the x and y transforms are unnormalised, the z solve is exact
    int n;                        // nz
    int nline;                    // points of the line a tile holds: nz, or nz / P in the z-slab sweep
    const double2 *zin;           // ZS_DIST: carry into the first local plane of every (kx, ky) line, [nx ny]
    // thread-per-line kernels: the launch covers columns [xs0, xs0 + xsn) of every row (the whole row by default);
    // the z-slab schedule pipelines the carry exchange of one half of the columns behind the work on the other
    int xs0, xsn;
};

// line index (x + nx y) of the t-th line of a launch that covers columns [xs0, xs0 + xsn)
__device__ __forceinline__ long long zs_line_of(const ZSolveArgs &a, long long t, int nx, int &x, int &y)
{
    y = (int)(t / a.xsn);
    x = a.xs0 + (int)(t - (long long)y * a.xsn);
    return x + (long long)nx * y;
}

// Signals between the ranks of a z-slab plan, carried by the kernels themselves instead of separate barrier launches
// (dist.cu keeps the flag arrays: flags[s] = last epoch rank s signalled to this rank, written over NVLink).
//   zs_signal_when_grid_done: every block fences its (peer) stores and counts itself; the last one to finish stores the
//                             epoch into every peer's flag array -- "this kernel's output has landed everywhere".
//   zs_wait_for_peers:        the first warp of every block waits until every rank's flag has reached the epoch.
// A wait gives up after ~2 s and raises *timeout (a dead peer must not hang the GPU); nranks == 0 disables both.
struct FlagSync {
    unsigned long long *peer[CPC_MAX_PEERS];      // every rank's flag array (IPC-mapped), this group
    unsigned long long *mine;                     // this rank's flag array
    unsigned long long epoch;
    int *counter;                                 // blocks of the signalling kernel that have finished
    int *timeout;
    int nranks, rank;
};

__device__ __forceinline__ void zs_wait_for_peers(const FlagSync &f)
{
    if (f.nranks == 0) return;
    if (threadIdx.x < (unsigned)f.nranks) {
        const long long t0 = clock64();
        unsigned long long seen = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f.mine + threadIdx.x) : "memory");
            if (seen >= f.epoch) break;
            if (clock64() - t0 > 4000000000ll) { *f.timeout = 1; break; }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void zs_signal_when_grid_done(const FlagSync &f)
{
    if (f.nranks == 0) return;
    __threadfence_system();                       // this thread's stores (to peers too) before the block counts itself
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(f.counter, 1);
        if (done == (int)gridDim.x - 1) {
            __threadfence_system();               // ... and every other block's, before the flags
            *f.counter = 0;
            for (int q = 0; q < f.nranks; ++q)
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.peer[q] + f.rank), "l"(f.epoch) : "memory");
        }
    }
}

enum ZSolveKind {
    ZS_CYCLIC = 0,    // the whole z line is in the tile: close the cycle inside the kernel (single GPU, transposed slabs)
    ZS_DIST = 2       // z-slab plans: the local part of the line with the carry-in computed by zs_carry_owner_kernel
};

__device__ __forceinline__ double2 to_d2(double2 v) { return v; }
__device__ __forceinline__ double2 to_d2(float2 v) { return make_double2((double)v.x, (double)v.y); }
template <typename C> __device__ __forceinline__ C from_d2(double2 v);
template <> __device__ __forceinline__ double2 from_d2<double2>(double2 v) { return v; }
template <> __device__ __forceinline__ float2 from_d2<float2>(double2 v) { return make_float2((float)v.x, (float)v.y); }

// z^P by binary exponentiation, unrolled at compile time
template <int P, typename C> __device__ __forceinline__ C cpow(C z)
{
    if constexpr (P == 1) return z;
    else if constexpr (P % 2 == 0) { const C h = cpow<P / 2>(z); return cmul(h, h); }
    else return cmul(z, cpow<P - 1>(z));
}
// z^p, runtime exponent p >= 1
__device__ __forceinline__ double2 cpow_rt(double2 z, int p)
{
    double2 r = make_double2(1.0, 0.0);
    while (p > 0) {
        if (p & 1) r = cmul(r, z);
        z = cmul(z, z);
        p >>= 1;
    }
    return r;
}

// c = lambda_z / (alpha + lambda_z) and r = 1 / (alpha + lambda_z) of line (x, y)
__device__ __forceinline__ void zs_coeffs(const ZSolveArgs &a, int x, int y, double2 &r, double2 &c)
{
    const double2 alpha = cadd(a.ax[x], a.ay[y]);
    r = crecip_scaled<double>(make_double2(alpha.x + a.lz, alpha.y), 1.0);
    c = make_double2(a.lz * r.x, a.lz * r.y);
}

// largest CTA the kernel is compiled for: E = 16 points are 64 registers of fp64 data, so <= 128 registers
template <int E> struct ZSolveMaxThreads { static constexpr int v = (E >= 16) ? 512 : 1024; };

template <typename T, int E, bool GEN, int KIND = ZS_CYCLIC>
__global__ void __launch_bounds__((ZSolveMaxThreads<E>::v), 1)
zsolve_kernel(const cplx_t<T> *in, cplx_t<T> *out, const PassGeom g, const ZSolveArgs a, const FlagSync wait)
{
    using CS = cplx_t<T>;                           // storage type
    if constexpr (KIND == ZS_DIST) zs_wait_for_peers(wait);       // the carry-ins have landed (zs_carry_owner_kernel)
    using C = double2;                              // arithmetic type
    constexpr int TX = 128 / (int)sizeof(CS);       // lanes = lines per tile
    constexpr int QW = 32 / TX;                     // segments per warp
    __shared__ C agg_all[32 * TX];                  // warp aggregates [warp of the CTA][lane]

    // A CTA holds blockDim.x / (threads per tile) tiles: short lines (nz / P planes of a z-slab at 8 GPUs are 64
    // points = one warp per tile) would otherwise make 32-thread CTAs.  Tiles never exchange data with each other.
    const int tpt = a.nline / E * TX;               // threads per tile, a multiple of 32
    const int grp = threadIdx.x / tpt;              // tile within the CTA
    const int tid = threadIdx.x - grp * tpt;
    const int l = tid % TX;
    const int seg = tid / TX;
    const int q = seg % QW;                         // segment within the warp
    const int wrp = tid >> 5;
    const int nwarps = tpt >> 5;
    const int k0 = seg * E;
    C *agg = agg_all + grp * nwarps * TX;

    const int t = blockIdx.x * (blockDim.x / tpt) + grp;
    const bool tile_ok = t < g.ntiles;
    const int ti = tile_ok ? t % g.tiles_inner : 0, to = tile_ok ? t / g.tiles_inner : 0;
    const int w = ti * TX + l;
    const bool active = tile_ok && w < g.lines_inner;
    const long long gbase = (long long)to * g.B1 + (long long)ti * g.B0 + (long long)l * g.SL;
    const long long obase = (long long)to * g.B1o + (long long)ti * g.B0o + (long long)l * g.SLo;

    C v[E];
    if (active) {
        if (!GEN) {
            const CS *p = in + gbase + (long long)k0 * g.SI;
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = to_d2(p[m * g.SI]);
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = to_d2(in[gbase + gen_in_off(g, k0 + m)]);
        }
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = make_double2(0.0, 0.0);
    }

    // r = 1 / (alpha + lambda_z), c = lambda_z r
    const int wc = active ? w : 0;
    C r, c;
    zs_coeffs(a, wc % g.nx, wc / g.nx + g.y0, r, c);

    // 1. local recurrence from a zero carry-in, in NB independent blocks of B points so that the dependent chain is
    //    B - 1 + NB - 1 complex FMAs deep instead of E - 1; the blocks' carries t[] are folded into step 3.
    constexpr int B = (E % 4 == 0) ? 4 : 5, NB = E / B;
    C pw[B];                                        // c^(j+1)
    pw[0] = c;
#pragma unroll
    for (int j = 1; j < B; ++j) pw[j] = cmul(pw[(j - 1) / 2], pw[j / 2]);      // c^(j+1) = c^(floor((j+1)/2)) c^(ceil((j+1)/2))
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
        for (int j = 1; j < B; ++j) {
            v[b * B + j].x = fma(c.x, v[b * B + j - 1].x, fma(-c.y, v[b * B + j - 1].y, v[b * B + j].x));
            v[b * B + j].y = fma(c.x, v[b * B + j - 1].y, fma(c.y, v[b * B + j - 1].x, v[b * B + j].y));
        }
    }
    C tb[NB];                                       // value at the end of block b with a zero carry into the segment
    tb[0] = v[B - 1];
#pragma unroll
    for (int b = 1; b < NB; ++b) tb[b] = cadd(v[b * B + B - 1], cmul(pw[B - 1], tb[b - 1]));
    const C cE = cpow<NB>(pw[B - 1]);               // c^E: the carry factor across one segment

    // 2a. inclusive scan over the warp's QW segments: P_q = sum_{i <= q} cE^(q-i) e_i
    C P = tb[NB - 1];
    C cp = cE;
#pragma unroll
    for (int d = 1; d < QW; d <<= 1) {
        C up;
        up.x = __shfl_up_sync(0xffffffffu, P.x, d * TX);
        up.y = __shfl_up_sync(0xffffffffu, P.y, d * TX);
        if (q >= d) P = cadd(P, cmul(cp, up));
        cp = cmul(cp, cp);
    }
    const C D = cp;                                 // cE^QW: the carry factor across one warp
    C Pex;                                          // exclusive prefix: the carry from the segments before q in this warp
    Pex.x = __shfl_up_sync(0xffffffffu, P.x, TX);
    Pex.y = __shfl_up_sync(0xffffffffu, P.y, TX);
    if (q == 0) Pex = make_double2(0.0, 0.0);
    if (q == QW - 1) agg[wrp * TX + l] = P;
    __syncthreads();

    // 2b. value carried into this warp
    C Z = make_double2(0.0, 0.0);
    if constexpr (KIND == ZS_CYCLIC) {
        // cycle closed: Z_{w-1} = sum_{m < NW} D^m A_{w-1-m} / (1 - D^NW)   (Horner over A_w, A_{w+1}, ..., A_{w-1})
        C acc = make_double2(0.0, 0.0), Dn = make_double2(1.0, 0.0);
        int idx = wrp;
        for (int i = 0; i < nwarps; ++i) {
            acc = cadd(cmul(D, acc), agg[idx * TX + l]);
            Dn = cmul(Dn, D);
            idx = (idx + 1 == nwarps) ? 0 : idx + 1;
        }
        Z = cmul(acc, crecip_scaled<double>(make_double2(1.0 - Dn.x, -Dn.y), 1.0));
    } else {
        // the carry into the local line closes the cycle over the ranks (zs_carry_owner_kernel); then through the
        // warps before this one: Z = D^wrp Zin + sum_{i < wrp} D^(wrp-1-i) A_i
        Z = a.zin[wc];
        for (int i = 0; i < wrp; ++i) Z = cadd(cmul(D, Z), agg[i * TX + l]);
    }
    // carry into this segment: cE^q Z + Pex
    C cq = make_double2(1.0, 0.0);
#pragma unroll
    for (int i = 1; i < QW; ++i)
        if (i <= q) cq = cmul(cq, cE);
    const C carry = cadd(cmul(cq, Z), Pex);

    // 3. x_k = r scale (y_k + c^(j+1) H_b),  H_b = carry into block b = t_{b-1} + c^(bB) carry
    const C rs = make_double2(r.x * a.scale, r.y * a.scale);
    C G = carry;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const C H = (b == 0) ? carry : cadd(tb[b - 1], G);
#pragma unroll
        for (int j = 0; j < B; ++j) v[b * B + j] = cmul(cadd(v[b * B + j], cmul(pw[j], H)), rs);
        G = cmul(G, pw[B - 1]);
    }

    if (active) {
        if (!GEN) {
            CS *p = out + obase + (long long)k0 * g.SIo;
#pragma unroll
            for (int m = 0; m < E; ++m) p[m * g.SIo] = from_d2<CS>(v[m]);
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) *gen_out_ptr<CS>(g, obase, k0 + m) = from_d2<CS>(v[m]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// z-slab plans: end values and the carry exchange
// ---------------------------------------------------------------------------------------------------------------
struct ZCarryPeers {
    double2 *p[CPC_MAX_PEERS];
};

// e[line] = c^zc e[line] (if carry_in) + sum_{k < zc} c^(zc-1-k) v[zb + k][line]: the value at the end of planes
// [zb, zb + zc) of every local line, continuing the planes before zb.  One thread per (kx, ky) line, coalesced over kx.
// With `trunc` the sum starts at the first plane whose weight |c|^(zc-1-k) can still reach 1e-17 of the last plane's:
// earlier planes (and the carry-in) are below the rounding of the sum itself.  At lambda = 55.56 seven lines out of
// eight have |c| < 0.54 and need fewer than 64 planes -- the planes the forward y pass wrote last, still in L2.
// With push_rank >= 0 (the last chunk, peers mapped) the result is stored straight into the gather buffer of the
// rank q = line / lsub that owns the line, G_q[push_rank][line - q lsub], over NVLink.
// push_rank == -2 (single rank, the planes are the whole line): the cycle is closed on the spot and e receives the
// carry into plane 0,  e / (1 - c^nz)  -- what zs_carry_owner_kernel computes for one rank.
template <typename T>
__global__ void __launch_bounds__(256)
zs_end_accum_kernel(const cplx_t<T> *__restrict__ x, long long lines, int nx, int zb, int zc, int carry_in, int trunc,
                    double2 *__restrict__ e, const ZSolveArgs a, int push_rank, long long lsub, ZCarryPeers gpeer,
                    const FlagSync done)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < lines / nx * a.xsn) {
        int lx, ly;
        const long long line = zs_line_of(a, t, nx, lx, ly);
        double2 r, c;
        zs_coeffs(a, lx, ly, r, c);
        int k = 0;
        if (trunc) {
            const double c2 = c.x * c.x + c.y * c.y;                  // |c|^2 < 1
            // |c|^m <= 1e-17  <=>  m >= ln(1e-17) / ln|c| = -78.3 / ln(|c|^2)
            const double m = c2 > 0.0 ? -78.3 / log(c2) + 1.0 : 1.0;
            if (m < (double)zc) k = zc - (int)m;
            // whole groups of 16 planes (the extra planes only carry less weight): no one-plane-at-a-time remainder
            k = zc - ((zc - k + 15) & ~15);
            if (k < 0) k = 0;
        }
        double2 acc = (carry_in && k == 0) ? e[line] : make_double2(0.0, 0.0);
        const cplx_t<T> *p = x + (long long)zb * lines + line;
        // 16 planes in flight per thread: the kernel's duration is the longest line's (the few low-frequency lines that
        // need every plane), i.e. (planes / planes in flight) DRAM latencies (ncu: 221 us for 1024 planes with 8 in flight)
        for (; k + 16 <= zc; k += 16) {
            double2 v[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = to_d2(p[(long long)(k + m) * lines]);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const double2 tt = acc;
                acc.x = fma(c.x, tt.x, fma(-c.y, tt.y, v[m].x));
                acc.y = fma(c.x, tt.y, fma(c.y, tt.x, v[m].y));
            }
        }
        for (; k < zc; ++k) {
            const double2 v = to_d2(p[(long long)k * lines]), tt = acc;
            acc.x = fma(c.x, tt.x, fma(-c.y, tt.y, v.x));
            acc.y = fma(c.x, tt.y, fma(c.y, tt.x, v.y));
        }
        if (push_rank >= 0) {
            const int q = (int)(line / lsub);
            gpeer.p[q][(long long)push_rank * lsub + (line - (long long)q * lsub)] = acc;
        } else if (push_rank == -2) {
            const double2 cn = cpow_rt(c, zc);
            e[line] = cmul(acc, crecip_scaled<double>(make_double2(1.0 - cn.x, -cn.y), 1.0));
        } else {
            e[line] = acc;
        }
    }
    zs_signal_when_grid_done(done);                // "this rank's end values have landed at their owners"
}

// Second sweep for slabs whose plane count fits no tile form of zsolve_kernel (nz / P = 8 planes at 8 ranks of a 64^3
// grid, odd counts ...): one thread per (kx, ky) line marches over the local planes from the exchanged carry-in,
//   y_k = c y_{k-1} + b_k,  x_k = r y_k / (nx ny),   8 planes in flight per thread (in == x allowed: a thread only
// touches its own line and reads a plane before it writes it).  Also the second sweep of the single-rank line form.
template <typename T>
__global__ void __launch_bounds__(256)
zs_dist_line_kernel(const cplx_t<T> *in, cplx_t<T> *x, long long lines, int nx, int nzl, const ZSolveArgs a,
                    const FlagSync wait)
{
    using CS = cplx_t<T>;
    zs_wait_for_peers(wait);                       // the carry-ins have landed (zs_carry_owner_kernel)
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= lines / nx * a.xsn) return;
    int lx, ly;
    const long long line = zs_line_of(a, t, nx, lx, ly);
    double2 r, c;
    zs_coeffs(a, lx, ly, r, c);
    const double2 rs = make_double2(r.x * a.scale, r.y * a.scale);
    double2 acc = a.zin[line];
    const CS *pi = in + line;
    CS *p = x + line;
    int k = 0;
    for (; k + 8 <= nzl; k += 8) {
        double2 v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = to_d2(pi[(long long)(k + m) * lines]);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const double2 t = acc;
            acc.x = fma(c.x, t.x, fma(-c.y, t.y, v[m].x));
            acc.y = fma(c.x, t.y, fma(c.y, t.x, v[m].y));
            v[m] = cmul(acc, rs);
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) p[(long long)(k + m) * lines] = from_d2<CS>(v[m]);
    }
    for (; k < nzl; ++k) {
        const double2 v = to_d2(pi[(long long)k * lines]), t = acc;
        acc.x = fma(c.x, t.x, fma(-c.y, t.y, v.x));
        acc.y = fma(c.x, t.y, fma(c.y, t.x, v.y));
        p[(long long)k * lines] = from_d2<CS>(cmul(acc, rs));
    }
}

// Owner of lines [line0, line0 + count): from the P end values of each line, the carry into every rank's first plane
//   Zin_r = sum_{m < P} cL^m e_{r-1-m} / (1 - cL^P),  cL = c^(nz / P)       (the cycle closed over the ranks),
// computed for r = 0 by Horner and then Zin_{r+1} = e_r + cL Zin_r; stored into rank r's zin buffer (peer store), or
// only into this rank's when self_only (fallback without peer mapping: every rank owns all lines).
__global__ void __launch_bounds__(256)
zs_carry_owner_kernel(const double2 *__restrict__ gbuf, long long gstride, long long line0, long long count, int nx,
                      int nzl, int nranks, int rank, int self_only, ZCarryPeers zpeer, const ZSolveArgs a,
                      const FlagSync wait, const FlagSync done);

}  // namespace cpc
