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
// the next.  Sweep 1 (ZS_END, read-only) computes the value at the end of every local line from a zero carry-in,
// the P x nx x ny end values are all-gathered (4 MB per rank at 512^2 instead of two 2 GB all-to-alls per apply), and
// sweep 2 (ZS_DIST) solves the local lines with the carry that closes the cycle over the ranks.
#pragma once
#include "fft_pass.cuh"

namespace cpc {

template <typename T> struct ZSolveArgs {
    const cplx_t<T> *ax, *ay;     // lambda_x c_x_hat[kx]  and  1 + lambda_y c_y_hat[ky]   (the plan's symbol tables)
    T lz;                         // lambda_z
    T scale;                      // 1 / (nx ny): the x and y transforms are unnormalised, the z solve is exact
    int n;                        // nz
    int nline;                    // points of the line a tile holds: nz, or nz / P in the z-slab sweeps
    // z-slab plans (DIST builds): the line is this rank's nz / P planes; the carry into its first plane comes from the
    // end values of every rank's slab (zero carry-in), all-gathered into ecat[P][nx ny]
    const cplx_t<T> *ecat;
    cplx_t<T> *eout;              // END builds: where this rank's end values go, [nx ny]
    int nranks, rank;
};

enum ZSolveKind {
    ZS_CYCLIC = 0,    // the whole z line is in the tile: close the cycle inside the kernel (single GPU, transposed slabs)
    ZS_END = 1,       // z-slab plans, first sweep: only the value at the end of the local line, zero carry-in (read-only)
    ZS_DIST = 2       // z-slab plans, second sweep: the local line with the carry computed from ecat
};

// z^P by binary exponentiation, unrolled at compile time
template <int P, typename C> __device__ __forceinline__ C cpow(C z)
{
    if constexpr (P == 1) return z;
    else if constexpr (P % 2 == 0) { const C h = cpow<P / 2>(z); return cmul(h, h); }
    else return cmul(z, cpow<P - 1>(z));
}

// largest CTA the kernel is compiled for: E = 16 complex128 points are 64 registers of data, so <= 128 registers
template <typename T, int E> struct ZSolveMaxThreads { static constexpr int v = (sizeof(T) == 8 && E >= 16) ? 512 : 1024; };

template <typename T, int E, bool GEN, int KIND = ZS_CYCLIC>
__global__ void __launch_bounds__((ZSolveMaxThreads<T, E>::v), 1)
zsolve_kernel(const cplx_t<T> *__restrict__ in, cplx_t<T> *__restrict__ out, const PassGeom g, const ZSolveArgs<T> a)
{
    using C = cplx_t<T>;
    constexpr int TX = 128 / (int)sizeof(C);        // lanes = lines per tile
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
            const C *p = in + gbase + (long long)k0 * g.SI;
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = p[m * g.SI];
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = in[gbase + gen_in_off(g, k0 + m)];
        }
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = mk<T>((T)0, (T)0);
    }

    // r = 1 / (alpha + lambda_z), c = lambda_z r
    const int wc = active ? w : 0;
    const int x = wc % g.nx, y = wc / g.nx + g.y0;
    const C alpha = cadd(a.ax[x], a.ay[y]);
    const C r = crecip_scaled<T>(mk<T>(alpha.x + a.lz, alpha.y), (T)1);
    const C c = mk<T>(a.lz * r.x, a.lz * r.y);

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
    if (q == 0) Pex = mk<T>((T)0, (T)0);
    if (q == QW - 1) agg[wrp * TX + l] = P;
    __syncthreads();

    // 2b. value carried into this warp
    C Z = mk<T>((T)0, (T)0);
    if constexpr (KIND == ZS_CYCLIC) {
        // cycle closed: Z_{w-1} = sum_{m < NW} D^m A_{w-1-m} / (1 - D^NW)   (Horner over A_w, A_{w+1}, ..., A_{w-1})
        C acc = mk<T>((T)0, (T)0), Dn = mk<T>((T)1, (T)0);
        int idx = wrp;
        for (int i = 0; i < nwarps; ++i) {
            acc = cadd(cmul(D, acc), agg[idx * TX + l]);
            Dn = cmul(Dn, D);
            idx = (idx + 1 == nwarps) ? 0 : idx + 1;
        }
        Z = cmul(acc, crecip_scaled<T>(mk<T>((T)1 - Dn.x, -Dn.y), (T)1));
    } else if constexpr (KIND == ZS_END) {
        // value at the end of the local line with a zero carry-in: sum_w D^(NW-1-w) A_w; one 128-byte row per tile
        if (wrp == nwarps - 1 && q == QW - 1) {
            C acc = mk<T>((T)0, (T)0);
            for (int i = 0; i < nwarps; ++i) acc = cadd(cmul(D, acc), agg[i * TX + l]);
            if (active) a.eout[w] = acc;
        }
        return;
    } else {
        // carry into the local line: the other slabs' end values with the cycle closed over the P ranks,
        //   Zin = sum_{m < P} cL^m e_{rank-1-m} / (1 - cL^P),  cL = c^(nz / P) = D^NW;
        // then through the warps before this one: Z = D^wrp Zin + sum_{i < wrp} D^(wrp-1-i) A_i
        C cL = mk<T>((T)1, (T)0);
        for (int i = 0; i < nwarps; ++i) cL = cmul(cL, D);
        C acc = mk<T>((T)0, (T)0), cLp = mk<T>((T)1, (T)0);
        int idx = a.rank;
        const long long plane = (long long)g.lines_inner;
        for (int i = 0; i < a.nranks; ++i) {
            acc = cadd(cmul(cL, acc), a.ecat[idx * plane + wc]);
            cLp = cmul(cLp, cL);
            idx = (idx + 1 == a.nranks) ? 0 : idx + 1;
        }
        Z = cmul(acc, crecip_scaled<T>(mk<T>((T)1 - cLp.x, -cLp.y), (T)1));
        for (int i = 0; i < wrp; ++i) Z = cadd(cmul(D, Z), agg[i * TX + l]);
    }
    // carry into this segment: cE^q Z + Pex
    C cq = mk<T>((T)1, (T)0);
#pragma unroll
    for (int i = 1; i < QW; ++i)
        if (i <= q) cq = cmul(cq, cE);
    const C carry = cadd(cmul(cq, Z), Pex);

    // 3. x_k = r scale (y_k + c^(j+1) H_b),  H_b = carry into block b = t_{b-1} + c^(bB) carry
    const C rs = mk<T>(r.x * a.scale, r.y * a.scale);
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
            C *p = out + obase + (long long)k0 * g.SIo;
#pragma unroll
            for (int m = 0; m < E; ++m) p[m * g.SIo] = v[m];
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) *gen_out_ptr<C>(g, obase, k0 + m) = v[m];
        }
    }
}

}  // namespace cpc
