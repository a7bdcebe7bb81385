// fft_pass.cuh -- one HBM pass of the circulant-preconditioner apply: a batch of 1-D Stockham FFTs.
//
// A pass transforms every line of one axis of the [nz][ny][nx][ncomp] array.  A CTA owns G "tiles"; a tile is
// TX neighbouring lines (neighbouring in the direction that is contiguous in HBM, so that each warp-level
// load/store touches full 32-byte sectors) times the N points of the line.  Thread (l, j) of a tile -- l = lane
// within the TX lines, j = butterfly column in [0, N/E) -- keeps E = max radix points of its line in registers:
//
//   global --(coalesced vector loads)--> registers --radix-R0--> smem --radix-R1--> smem --radix-R2--> registers
//
// Stockham autosort (decimation in time): in the stage with radix R and p = product of the previous radices, the
// butterfly jb reads x[jb + r N/R], multiplies by exp(-/+ 2 pi i r (jb mod p) / (pR)), does an R-point DFT and
// writes y[(jb - jb mod p) R + (jb mod p) + r p].  With E/R butterflies jb = j + b N/E per thread, every stage
// reads exactly the indices {j + (N/E) m, m < E}, and the last stage writes exactly those indices, so
//   * the first stage loads straight from HBM and the last stage stores straight to HBM (no staging copy), and
//   * in the fused middle pass the forward transform's outputs are, in registers, already the inputs of the
//     backward transform's first stage: forward-z, the eigenvalue division and backward-z are ONE pass
//     (reference src/FftLinearSolver_3D.c:170-184 does them as four full-array sweeps).
//
// Thread/shared-memory mappings (template flag XMAP):
//   XMAP = false (lines strided in HBM: y, z passes, and the wave x pass whose 4 components are the lanes):
//       l = tid % TX fastest, shared memory [N][TX]: the 8 (fp64, TX=8) lanes of a quarter warp always hit 128
//       contiguous bytes of HBM and of shared memory -> coalesced and bank-conflict-free for any point index.
//   XMAP = true (lines contiguous in HBM: the scalar x pass):
//       j = tid % (N/E) fastest, so a warp reads 32 consecutive points (512 contiguous bytes) of ONE line; shared
//       memory is [TX][N + N/R0] with one pad element every R0 points, which makes the stride-R0 stores of the first
//       stage and every later (consecutive) access conflict-free.
// Twiddles come from per-stage tables laid out [r-1][k] (k = butterfly index mod p), so the lanes of a warp read
// consecutive entries: one LSU wavefront per load instead of one per distinct cache line.
#pragma once
#include "fft_core.cuh"

namespace cpc {

#define CPC_MAX_PEERS 8

// Geometry of one pass (all in units of complex elements).
struct PassGeom {
    long long SI;         // stride between consecutive points of a line
    long long SL;         // stride between the TX lines of a tile (input side)
    long long SLo;        // same on the output side (differs only in the r2c / c2r passes)
    long long B0, B1;     // tile t starts at (t / tiles_inner) * B1 + (t % tiles_inner) * B0
    int tiles_inner;
    int ntiles;
    int lines_inner;      // number of valid lines along the SL direction per outer index (for partial tiles)
    // Output side (equal to the input side except in the multi-rank y passes) and split addressing: with D > 0
    // point i of a line lives at (i / D) * SC + (i % D) * S, i.e. the line is cut into per-destination-rank
    // chunks (the "zero-pack" slab layout [q][z_loc][y_loc][x]).  sh = log2(D) or -1.
    long long SIo, B0o, B1o;
    long long SCi, SCo;
    int Di, Do, shi, sho;
    int maski, masko;     // Di - 1 / Do - 1 for power-of-two splits, 0x7fffffff (with sh = 31) when there is no split
    // symbol addressing (fused pass only): line index w = (t % tiles_inner) * TX + l decomposes as
    //   c = w % ncomp, x = (w / ncomp) % nx, y = w / (ncomp * nx)
    int ncomp, nx, ny;
    int y0;               // global y of local y index 0 (multi-rank transposed slab)
    int stagger;          // > 0: CTAs of the second resident slot of every SM start this many cycles late, so that
    int num_sms;          //      the two CTAs sharing an SM are not in the same (fp64 vs LSU) phase all the time
    int pf_tiles;         // > 0: prefetch into L2 the tile `pf_tiles` after this one (about one wave of CTAs ahead)
    // Fused transpose (multi-rank plans with peer access): with npeer > 0 the chunk index i / Do selects the rank
    // whose buffer receives the point, and the store goes straight to that rank's HBM over NVLink
    // (peer[q] is the IPC-mapped base of rank q's buffer, already offset to this rank's slot).
    int npeer;
    void *peer[CPC_MAX_PEERS];
};

enum PassMode {
    MODE_FWD = 0, MODE_INV = 1, MODE_FUSED_SEP = 2, MODE_FUSED_TABLE = 3, MODE_FUSED_WAVE = 4,
    // real-scalar plans (XMAP x pass only): a real line of 2N points is transformed as N complex points
    MODE_R2C = 5,     // forward: N-point FFT of z[m] = x[2m] + i x[2m+1], then untangle to X[0..N] (N+1 outputs)
    MODE_C2R = 6      // backward: tangle X[0..N] into N complex points, N-point backward FFT, store as 2N reals
};

template <typename T> struct SymbolArgs {
    // separable: Lambda = ax[x] + ay[y] + az[k]  (ay carries the "+1"), result scaled by `scale` = 1/N
    const cplx_t<T> *ax, *ay, *az;
    // table: precomputed scale / Lambda, same layout as the data
    const cplx_t<T> *inv_table;
    // wave: 1-D root tables exp(-2 pi i q / n) per axis
    const cplx_t<T> *rx, *ry, *rz;
    T c0, mux, muy, muz;
    T scale;
};

// Output address of point i of a line: local (split or plain) or a peer's buffer.
template <typename C>
__device__ __forceinline__ C *out_ptr(C *out, const PassGeom &g, long long obase, int i)
{
    if (g.npeer > 0) {
        const int q = g.sho >= 0 ? (i >> g.sho) : (i / g.Do);
        const int r = g.sho >= 0 ? (i & (g.Do - 1)) : (i % g.Do);
        return reinterpret_cast<C *>(g.peer[q]) + obase + (long long)r * g.SIo;
    }
    if (g.Do == 0) return out + obase + (long long)i * g.SIo;
    if (g.sho >= 0) return out + obase + (long long)(i >> g.sho) * g.SCo + (long long)(i & (g.Do - 1)) * g.SIo;
    return out + obase + (long long)(i / g.Do) * g.SCo + (long long)(i % g.Do) * g.SIo;
}

// Branch-free forms used by the general-addressing builds of the fast kernels (power-of-two splits only; the host
// sets sh = 31, mask = 0x7fffffff, SC = 0 and peer[q] = out when a side is a plain strided line).
__device__ __forceinline__ long long gen_in_off(const PassGeom &g, int i)
{
    return (long long)(i >> g.shi) * g.SCi + (long long)(i & g.maski) * g.SI;
}
template <typename C> __device__ __forceinline__ C *gen_out_ptr(const PassGeom &g, long long obase, int i)
{
    const int q = i >> g.sho;
    C *base = reinterpret_cast<C *>(g.peer[0]);                      // the local output unless peers are mapped
    if (g.npeer > 0) base = reinterpret_cast<C *>(g.peer[q]);        // (indexed kernel-parameter load only then)
    return base + obase + (long long)q * g.SCo + (long long)(i & g.masko) * g.SIo;
}

__device__ __forceinline__ long long point_off(int i, long long S, int D, int sh, long long SC)
{
    if (D == 0) return (long long)i * S;
    if (sh >= 0) return (long long)(i >> sh) * SC + (long long)(i & (D - 1)) * S;
    return (long long)(i / D) * SC + (long long)(i % D) * S;
}

template <int A, int B> struct CMax { static constexpr int v = A > B ? A : B; };

// 16-byte (8-byte for float2) read-only load that the compiler keeps where it is written
__device__ __forceinline__ double2 ld_nc_ordered(const double2 *p)
{
    double2 r;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float2 ld_nc_ordered(const float2 *p)
{
    float2 r;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p) : "memory");
    return r;
}


// ---------------------------------------------------------------------------------------------------------------
// Shared-memory index of point i of lane-line l.
// ---------------------------------------------------------------------------------------------------------------
template <int N, int TX, int PADSH, bool XMAP> __device__ __forceinline__ int sm_index(int i, int l)
{
    if (XMAP) return l * (N + (N >> PADSH)) + i + (i >> PADSH);
    return i * TX + l;
}
template <int N, int TX, int PADSH, bool XMAP> struct SmemTile {
    static constexpr int elems = XMAP ? TX * (N + (N >> PADSH)) : N * TX;
};
template <int R> struct Log2 { static constexpr int v = 1 + Log2<R / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

// ---------------------------------------------------------------------------------------------------------------
// One Stockham stage on the register file.  v[] is indexed by m with point index j + TPL*m.
// tw points at this stage's table: tw[(r-1)*P + k] = exp(-2 pi i r k / (P R)).
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int N, int E, int R, int P, int DIR, bool TO_SMEM, int TX, int PADSH, bool XMAP>
__device__ __forceinline__ void stockham_stage(cplx_t<T> (&v)[E], int j, int l, cplx_t<T> *sm,
                                               const cplx_t<T> *__restrict__ tw)
{
    using C = cplx_t<T>;
    constexpr int TPL = N / E;
    constexpr int NB = E / R;          // butterflies per thread in this stage
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int jb = j + b * TPL;
        C u[R];
#pragma unroll
        for (int r = 0; r < R; ++r) u[r] = v[b + r * NB];
        // butterfly index within the previous stages' period: a mask for the power-of-two kernels, a constant
        // modulo (multiply-high) when an earlier radix was 5, 10 or 20
        const int k = (P > 1) ? (((P & (P - 1)) == 0) ? (jb & (P - 1)) : (jb % P)) : 0;
        if constexpr (P > 1 && R == 8) {
            C w[7];
#pragma unroll
            for (int r = 1; r < 8; ++r) w[r - 1] = __ldg(&tw[(r - 1) * P + k]);
            butterfly8_twiddled<DIR>(u, w);
        } else {
            if (P > 1) {
#pragma unroll
                for (int r = 1; r < R; ++r) u[r] = twmul<DIR>(u[r], __ldg(&tw[(r - 1) * P + k]));
            }
            Butterfly<R, DIR, C>::run(u);
        }
        if (TO_SMEM) {
            const int j0 = (jb - k) * R + k;
#pragma unroll
            for (int r = 0; r < R; ++r) sm[sm_index<N, TX, PADSH, XMAP>(j0 + r * P, l)] = u[r];
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) v[b + r * NB] = u[r];
        }
    }
}

template <typename T, int N, int E, int TX, int PADSH, bool XMAP>
__device__ __forceinline__ void smem_gather(cplx_t<T> (&v)[E], int j, int l, const cplx_t<T> *sm)
{
    constexpr int TPL = N / E;
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = sm[sm_index<N, TX, PADSH, XMAP>(j + TPL * m, l)];
}

// Barrier over the threads that share lines.  With NG > 1 a tile's TX lines are split into NG independent
// sub-groups (lines never exchange data with each other), each with its own named barrier, so the sub-groups of a
// CTA drift apart and one group's butterfly (fp64) phase overlaps another group's shared-memory (LSU) phase.
template <int NTHREADS> __device__ __forceinline__ void group_sync(int bar_id)
{
    if (NTHREADS == 0) __syncthreads();      // whole CTA (too many groups for the 16 named barriers, or partial warps)
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NTHREADS) : "memory");
}

// Full 1-D transform of the thread's line: registers -> registers (through shared memory).
// Stage tables are concatenated: stage 1 at tw, stage 2 at tw + (R1-1)*R0.
template <typename T, int N, int R0, int R1, int R2, int E, int DIR, int TX, bool XMAP, int NT>
__device__ __forceinline__ void line_fft(cplx_t<T> (&v)[E], int j, int l, cplx_t<T> *sm,
                                         const cplx_t<T> *__restrict__ tw, int bar)
{
    constexpr int NST = (R1 > 1) + (R2 > 1) + 1;
    constexpr int PS = Log2<R0>::v;
    const cplx_t<T> *tw2 = tw + (R1 - 1) * R0;
    if (NST == 1) {
        stockham_stage<T, N, E, R0, 1, DIR, false, TX, PS, XMAP>(v, j, l, sm, tw);
    } else if (NST == 2) {
        stockham_stage<T, N, E, R0, 1, DIR, true, TX, PS, XMAP>(v, j, l, sm, tw);
        group_sync<NT>(bar);
        smem_gather<T, N, E, TX, PS, XMAP>(v, j, l, sm);
        stockham_stage<T, N, E, R1, R0, DIR, false, TX, PS, XMAP>(v, j, l, sm, tw);
    } else {
        stockham_stage<T, N, E, R0, 1, DIR, true, TX, PS, XMAP>(v, j, l, sm, tw);
        group_sync<NT>(bar);
        smem_gather<T, N, E, TX, PS, XMAP>(v, j, l, sm);
        group_sync<NT>(bar);
        stockham_stage<T, N, E, R1, R0, DIR, true, TX, PS, XMAP>(v, j, l, sm, tw);
        group_sync<NT>(bar);
        smem_gather<T, N, E, TX, PS, XMAP>(v, j, l, sm);
        stockham_stage<T, N, E, R2, R0 * R1, DIR, false, TX, PS, XMAP>(v, j, l, sm, tw2);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Eigenvalue division, applied on the registers between the forward and the backward z transform.
// ---------------------------------------------------------------------------------------------------------------
// Point m of the thread sits at frequency index k0 + kstep * m of the line.
template <typename T, int E, int MODE>
__device__ __forceinline__ void apply_symbol(cplx_t<T> (&v)[E], int k0, int kstep, int w, long long gbase_l, long long SI,
                                             const PassGeom &g, const SymbolArgs<T> &s)
{
    using C = cplx_t<T>;
    if (MODE == MODE_FUSED_SEP) {
        // Lambda[k,j,i] = 1 + lx cx[i] + ly cy[j] + lz cz[k]   (reference FftLinearSolver_3D.c:146-157)
        const int x = w % g.nx, y = w / g.nx + g.y0;
        const C a = cadd(s.ax[x], s.ay[y]);
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const C lam = cadd(a, s.az[k0 + kstep * m]);
            v[m] = cmul(v[m], crecip_scaled<T>(lam, s.scale));   // b_hat / Diag (:174) and 1/size (:184)
        }
    } else if (MODE == MODE_FUSED_TABLE) {
        // Four table entries at a time: with all E loads hoisted to the top (what ptxas does when left alone) the
        // fused kernel needs 2 E more registers than it has and spills 100-140 B per thread.  The rows were pulled
        // into L2 at the top of the kernel (prefetch_symbol_table), so each group waits for L2, not for HBM.
        constexpr int GRP = (E % 4 == 0) ? 4 : 1;
#pragma unroll
        for (int m0 = 0; m0 < E; m0 += GRP) {
            C t[GRP];
#pragma unroll
            for (int i = 0; i < GRP; ++i) t[i] = ld_nc_ordered(&s.inv_table[gbase_l + (long long)(k0 + kstep * (m0 + i)) * SI]);
#pragma unroll
            for (int i = 0; i < GRP; ++i) v[m0 + i] = cmul(v[m0 + i], t[i]);
            asm volatile("" ::: "memory");
        }
    } else if (MODE == MODE_FUSED_WAVE) {
        // 4 consecutive lanes hold (p, rho0 u, rho0 v, rho0 w) of one cell.  Arrow-matrix Schur solve
        // (blocks from reference src/WaveSystem.cxx:92-107):
        //   p  = (r0 - sum_d i c0^2 s_d r_d / D_d) / (M00 + c0^2 sum_d s_d^2 / D_d),  y_d = (r_d - i s_d p) / D_d
        const int c = w & 3;
        const int x = (w >> 2) % g.nx, y = (w >> 2) / g.nx + g.y0;
        const C rx = s.rx[x], ry = s.ry[y];
        const T sx = -rx.y * s.mux, sy = -ry.y * s.muy;                    // mu_d sin(theta_d)
        const T ox = ((T)1 - rx.x) * s.mux, oy = ((T)1 - ry.x) * s.muy;    // mu_d (1 - cos(theta_d))
        const T c0 = s.c0, c02 = s.c0 * s.c0;
        const T iDx = fast_rcp((T)1 + c0 * ox), iDy = fast_rcp((T)1 + c0 * oy);
        const T sxx = sx * sx * iDx, syy = sy * sy * iDy;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const C rz = s.rz[k0 + kstep * m];
            const T sz = -rz.y * s.muz, oz = ((T)1 - rz.x) * s.muz;
            const T Dz = (T)1 + c0 * oz, iDz = fast_rcp(Dz);
            const T m00 = (T)1 + c0 * (ox + oy + oz);
            const T den = m00 + c02 * (sxx + syy + sz * sz * iDz);
            const T sd = (c == 1) ? sx : (c == 2) ? sy : sz;
            const T iDd = (c == 1) ? iDx : (c == 2) ? iDy : iDz;
            // contribution of this lane to the numerator: c=0: r0 ; c=d: -i c0^2 s_d r_d / D_d
            C t;
            if (c == 0) t = v[m];
            else {
                const T f = c02 * sd * iDd;
                t = mk<T>(v[m].y * f, -v[m].x * f);
            }
            t.x += __shfl_xor_sync(0xffffffffu, t.x, 1);
            t.y += __shfl_xor_sync(0xffffffffu, t.y, 1);
            t.x += __shfl_xor_sync(0xffffffffu, t.x, 2);
            t.y += __shfl_xor_sync(0xffffffffu, t.y, 2);
            const T inv = s.scale * fast_rcp(den);
            const C p = mk<T>(t.x * inv, t.y * inv);                       // already scaled by 1/N
            if (c == 0) v[m] = p;
            else                                                           // (r_d / N - i s_d p) / D_d
                v[m] = mk<T>((v[m].x * s.scale + sd * p.y) * iDd, (v[m].y * s.scale - sd * p.x) * iDd);
        }
    }
}

// Table symbols: pull the thread's E table rows towards L2 at the top of the fused kernel (no registers held); one
// lane per 128-byte row issues the prefetch.
template <typename T, int E>
__device__ __forceinline__ void prefetch_symbol_table(const SymbolArgs<T> &s, long long gbase_l, long long SI, int k0, int kstep,
                                                      bool row_leader)
{
    if (!row_leader) return;
#pragma unroll
    for (int m = 0; m < E; ++m)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(s.inv_table + gbase_l + (long long)(k0 + kstep * m) * SI));
}

// ---------------------------------------------------------------------------------------------------------------
// The pass kernel.  E = points per thread (a multiple of every radix), TX = lines per tile, G = tiles per CTA.
// ---------------------------------------------------------------------------------------------------------------
// GEN = false compiles the chunked-layout / peer-push addressing out (single-rank plans): ~15 % less code, which
// matters for the fused kernel whose straight-line SASS otherwise exceeds the 32 KB instruction cache.
template <typename T, int N, int R0, int R1, int R2, int E, int TX, int G, int MODE, int MINB, bool XMAP, bool GEN = true,
          int NG = 1>
__global__ void __launch_bounds__((N / E) * TX * G, MINB)
fft_pass_kernel(const cplx_t<T> *in, cplx_t<T> *out, const PassGeom g,
                const cplx_t<T> *__restrict__ tw, const SymbolArgs<T> sym)
{
    using C = cplx_t<T>;
    constexpr int TPL = N / E;
    constexpr int NST = (R1 > 1) + (R2 > 1) + 1;
    static_assert(R0 * R1 * R2 == N, "radices must multiply to N");
    static_assert(E % R0 == 0 && E % R1 == 0 && E % R2 == 0, "each radix must divide the register tile");
    extern __shared__ __align__(16) unsigned char smem_raw[];

    // thread -> (tile of the CTA, sub-group of the tile, line, butterfly column)
    constexpr int TXG = TX / NG;                  // lines per sub-group
    constexpr int NTG = TXG * TPL;                // threads per sub-group
    // participants of the sub-group's named barrier; 0 = fall back to __syncthreads()
    constexpr int NT = (G * NG <= 15 && NTG % 32 == 0 && G * NG > 1) ? NTG : 0;
    static_assert(TX % NG == 0 && (NG == 1 || (!XMAP && NTG % 32 == 0)), "sub-groups must be whole warps");
    const int tid = threadIdx.x;
    const int grp = tid / (TX * TPL);
    const int sg = (tid / NTG) % NG;
    const int ts = tid % NTG;
    const int lg = XMAP ? (ts / TPL) % TXG : ts % TXG;      // line within the sub-group (shared-memory lane)
    const int l = sg * TXG + lg;                            // line within the tile
    const int j = XMAP ? ts % TPL : (ts / TXG) % TPL;
    const int bar = 1 + grp * NG + sg;                      // named barrier of this sub-group (0 is __syncthreads)
    C *sm = reinterpret_cast<C *>(smem_raw) +
            (size_t)(grp * NG + sg) * (NST > 1 ? SmemTile<N, TXG, Log2<R0>::v, XMAP>::elems : 0);

    const int t = blockIdx.x * G + grp;
    const bool tile_ok = t < g.ntiles;
    const int ti = tile_ok ? t % g.tiles_inner : 0;
    const int to = tile_ok ? t / g.tiles_inner : 0;
    const int w = ti * TX + l;                                   // line index along the SL direction
    const bool active = tile_ok && (w < g.lines_inner);
    const long long gbase = (long long)to * g.B1 + (long long)ti * g.B0 + (long long)l * g.SL;
    const long long obase = (long long)to * g.B1o + (long long)ti * g.B0o + (long long)l * g.SLo;

    // Pull the tile that a later wave of CTAs will read towards L2 while this one computes: HBM stays busy during
    // the butterfly phases although only a few CTAs fit on an SM.  One 128-byte row per prefetch instruction.
    if (g.pf_tiles > 0) {
        const int tp = t + g.pf_tiles;
        if (tp < g.ntiles) {
            const long long pbase = (long long)(tp / g.tiles_inner) * g.B1 + (long long)(tp % g.tiles_inner) * g.B0;
            if (XMAP) {
                // lines are contiguous: thread (l, j) covers 128 bytes = 128/sizeof(C) points at a time
                constexpr int PER = 128 / (int)sizeof(C);
                for (int i = j * PER; i < N; i += TPL * PER)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(in + pbase + (long long)l * g.SL + i));
            } else {
                for (int m = l; m < E; m += TX)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(in + pbase + point_off(j + TPL * m, g.SI, g.Di, g.shi, g.SCi)));
            }
        }
    }

    if (MODE == MODE_FUSED_TABLE && active) prefetch_symbol_table<T, E>(sym, gbase, g.SI, j, TPL, XMAP || (l % (128 / (int)sizeof(C))) == 0);

    C v[E];
    if (active) {
        if (!GEN) {      // plain strided line
            const C *p = in + gbase + (long long)j * g.SI;
            const long long step = (long long)TPL * g.SI;
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = p[m * step];
        } else {         // chunked layout on the load side (multi-rank backward y pass), branch-free
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = in[gbase + gen_in_off(g, j + TPL * m)];
        }
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = mk<T>((T)0, (T)0);
    }

    if (MODE == MODE_FWD) {
        line_fft<T, N, R0, R1, R2, E, -1, TXG, XMAP, NT>(v, j, lg, sm, tw, bar);
    } else if (MODE == MODE_INV) {
        line_fft<T, N, R0, R1, R2, E, +1, TXG, XMAP, NT>(v, j, lg, sm, tw, bar);
    } else if (MODE == MODE_R2C) {
        // Untangle (reference a2/a4 rows of SURVEY.md 8a: the real-scalar build's r2c transform):
        //   X[k] = E[k] + w^k O[k],  E[k] = (Z[k] + conj Z[N-k]) / 2,  O[k] = -i (Z[k] - conj Z[N-k]) / 2,  w = exp(-i pi / N)
        line_fft<T, N, R0, R1, R2, E, -1, TXG, XMAP, NT>(v, j, lg, sm, tw, bar);
        constexpr int PS = Log2<R0>::v;
        group_sync<NT>(bar);
#pragma unroll
        for (int m = 0; m < E; ++m) sm[sm_index<N, TXG, PS, XMAP>(j + TPL * m, lg)] = v[m];
        group_sync<NT>(bar);
        C nyq = mk<T>((T)0, (T)0);
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int k = j + TPL * m;
            const C a = v[m];
            const C bq = sm[sm_index<N, TXG, PS, XMAP>(k == 0 ? 0 : N - k, lg)];
            const C ev = mk<T>((T)0.5 * (a.x + bq.x), (T)0.5 * (a.y - bq.y));
            const C od = mk<T>((T)0.5 * (a.y + bq.y), (T)-0.5 * (a.x - bq.x));       // -i (a - conj b) / 2
            const C wk = sym.rx[k];                                                   // exp(-2 pi i k / (2N))
            v[m] = cadd(ev, cmul(od, wk));
            if (k == 0) nyq = mk<T>(a.x - a.y, (T)0);                                // X[N] = Re Z[0] - Im Z[0]
        }
        if (active && j == 0) out[obase + N] = nyq;
    } else if (MODE == MODE_C2R) {
        //   Z[k] = (X[k] + conj X[N-k]) + i conj(w^k) (X[k] - conj X[N-k])      (unnormalised, cf. MODE_R2C)
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int k = j + TPL * m;
            const C a = v[m];
            const C bq = active ? in[gbase + (N - k)] : mk<T>((T)0, (T)0);
            const C s1 = mk<T>(a.x + bq.x, a.y - bq.y);
            const C d1 = mk<T>(a.x - bq.x, a.y + bq.y);
            const C t1 = cmulc(d1, sym.rx[k]);                                        // conj(w^k) (a - conj b)
            v[m] = mk<T>(s1.x - t1.y, s1.y + t1.x);                                   // s1 + i t1
        }
        line_fft<T, N, R0, R1, R2, E, +1, TXG, XMAP, NT>(v, j, lg, sm, tw, bar);
    } else {
        line_fft<T, N, R0, R1, R2, E, -1, TXG, XMAP, NT>(v, j, lg, sm, tw, bar);
        apply_symbol<T, E, MODE>(v, j, TPL, active ? w : 0, gbase, g.SI, g, sym);
        if (NST > 1) group_sync<NT>(bar);
        line_fft<T, N, R0, R1, R2, E, +1, TXG, XMAP, NT>(v, j, lg, sm, tw, bar);
    }

    if (active) {
        if (!GEN) {      // plain strided line
            C *p = out + obase + (long long)j * g.SIo;
            const long long step = (long long)TPL * g.SIo;
#pragma unroll
            for (int m = 0; m < E; ++m) p[m * step] = v[m];
        } else {         // chunked layout or peer push (multi-rank plans), branch-free
#pragma unroll
            for (int m = 0; m < E; ++m) *gen_out_ptr<C>(g, obase, j + TPL * m) = v[m];
        }
    }
}

}  // namespace cpc
