// fft_r2x.cuh -- 2H-point strided-line pass as  2 x (R0 x R1)  (512 = 2 x 16 x 16, 256 = 2 x 16 x 8):  one radix-2
// level in registers, a warp shuffle, and a two-stage Stockham transform per half -- ONE shared-memory exchange per transform instead
// of the two that the 8.8.8 kernel needs.  The fused z pass is limited by LSU wavefronts and barrier-serialised
// phases, not by HBM (profiles/r01_notes.md); this variant cuts its shared-memory traffic by half.
//
// Decimation in frequency for the first level (forward):
//     a[k] = x[k] + x[k+H],   b[k] = (x[k] - x[k+H]) W_{2H}^k,   X[2q] = FFT_H(a)[q],   X[2q+1] = FFT_H(b)[q]
// and its mirror (decimation in time) for the backward transform:
//     x[n] = yA[n] + W_{2H}^{-n} yB[n],   x[n+H] = yA[n] - W_{2H}^{-n} yB[n],   yA = IFFT_H(X[2q]),  yB = IFFT_H(X[2q+1]).
// A line is handled by 32 threads: thread (s, j), s = which half-transform it runs after the exchange, j in [0,16).
// Thread (s, j) loads the pairs (k, k+H) for k = j + 16 m, m in [8s, 8s+8): both members of a pair sit in one thread,
// so the radix-2 level needs no exchange; the partner (1-s, j) is 8 lanes away in the same warp, and the two swap
// half of their points with shfl.xor.  After that each thread holds the 16 points {j + 16 m} of its half-line: the
// standard input of the 16 x 16 Stockham kernel code.
#pragma once
#include "fft_pass.cuh"

namespace cpc {

template <typename C> __device__ __forceinline__ C shfl_xor_c(C v, int mask)
{
    C r;
    r.x = __shfl_xor_sync(0xffffffffu, v.x, mask);
    r.y = __shfl_xor_sync(0xffffffffu, v.y, mask);
    return r;
}

// N = 2 H points per line, H = R0 * R1 (two radix stages, 16 points per thread): 512 = 2 x (16 x 16) and
// 256 = 2 x (16 x 8).  The half-line transform is the generic Stockham code of fft_pass.cuh run with the 16 "lanes"
// (s, l): shared memory [point][s][l], so the 8 lanes of a quarter warp still hit one 128-byte row.
template <typename T, int H, int R0, int R1, int MODE, bool GEN, int TX = 128 / (int)sizeof(cplx_t<T>)>
__global__ void __launch_bounds__(H * TX / 8, (TX == 16 ? 8192 : 4096) / (H * TX))
fft_r2x_kernel(const cplx_t<T> *in, cplx_t<T> *out, const PassGeom g,
               const cplx_t<T> *__restrict__ tw, const SymbolArgs<T> sym)
{
    using C = cplx_t<T>;
    // TX lanes x sizeof(C) = one 128-byte row: 8 lanes for complex128, 16 for complex64 (the partner lane of the
    // radix-2 swap is TX lanes away, still inside the warp)
    static_assert(TX == 8 || TX == 16, "TX must be 8 or 16");
    constexpr int LTX = TX == 8 ? 3 : 4;
    constexpr int TP = H / 16;            // threads per half-line = stride between a thread's points
    static_assert(R0 * R1 == H && 16 % R0 == 0 && 16 % R1 == 0, "half-line must be two radix stages of 16 points per thread");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *sm = reinterpret_cast<C *>(smem_raw);

    const int tid = threadIdx.x;
    const int l = tid & (TX - 1);
    const int s = (tid >> LTX) & 1;
    const int j = tid >> (LTX + 1);       // in [0, TP)
    const int ls = s * TX + l;            // the (s, l) "lane" of the half-line transform

    const int t = blockIdx.x;
    const int ti = t % g.tiles_inner, to = t / g.tiles_inner;
    const int w = ti * TX + l;
    const bool active = w < g.lines_inner;
    const long long gbase = (long long)to * g.B1 + (long long)ti * g.B0 + (long long)l * g.SL;
    const long long obase = (long long)to * g.B1o + (long long)ti * g.B0o + (long long)l * g.SLo;
    const C *rt = sym.rz;                 // roots exp(-2 pi i k / N) of the transformed axis

    // Tuning hook (CPC_STAGGER, default 0 = no delay): start every SM's second resident CTA late.  The delay itself
    // made no difference (profiles/r01_notes.md), but the loop is an instruction-scheduling fence at the top of the
    // kernel: with it ptxas keeps the fused kernel's spills at 8 B instead of 64 B and the pass runs 1.05 ms instead
    // of 1.22 ms.  It is compiled into the fused modes only (it costs the plain backward pass 0.09 ms).
    if (MODE >= MODE_FUSED_SEP && g.stagger > 0 && blockIdx.x >= (unsigned)g.num_sms && blockIdx.x < 2u * (unsigned)g.num_sms) {
        const long long t0 = clock64();
        while (clock64() - t0 < g.stagger) { }
    }

    // table symbols: the fused pass reads 1 / (N Lambda) at the points X[2 (j + TP m) + s]
    if (MODE == MODE_FUSED_TABLE && active) prefetch_symbol_table<T, 16>(sym, gbase, g.SI, 2 * j + s, 2 * TP, l == 0);

    C u[16];
    // Plain transforms (forward and backward) both use the decimation-in-frequency structure, with conjugated
    // twiddles for the backward one: it measured 0.64 ms for a 512^3 y pass against 0.81 ms for the decimation-in-time
    // structure, which is kept only where it is needed -- after the eigenvalue division of the fused pass.
    constexpr int D1 = (MODE == MODE_INV) ? +1 : -1;
    {
        // ---- load pairs, radix-2 level (decimation in frequency) -------------------------------------------------
        C a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = j + TP * (8 * s + i);
            C x0 = mk<T>((T)0, (T)0), x1 = x0;
            if (active) {
                x0 = in[gbase + (GEN ? gen_in_off(g, k) : (long long)k * g.SI)];
                x1 = in[gbase + (GEN ? gen_in_off(g, k + H) : (long long)(k + H) * g.SI)];
            }
            a[i] = cadd(x0, x1);
            b[i] = twmul<D1>(csub(x0, x1), __ldg(&rt[k]));
        }
        // ---- swap halves with the partner thread (lane ^ 8) ---------------------------------------------------------
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const C send = s ? a[i] : b[i];
            const C recv = shfl_xor_c(send, TX);
            u[i] = s ? recv : a[i];
            u[i + 8] = s ? b[i] : recv;
        }
        line_fft<T, H, R0, R1, 1, 16, D1, 2 * TX, false, 0>(u, j, ls, sm, tw, 0);   // u[m] = X[2 (j + TP m) + s]
    }

    if (MODE == MODE_FWD || MODE == MODE_INV) {
        if (active) {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int K = 2 * (j + TP * m) + s;
                if (!GEN) out[obase + (long long)K * g.SIo] = u[m];
                else *gen_out_ptr<C>(g, obase, K) = u[m];
            }
        }
        return;
    }

    apply_symbol<T, 16, MODE>(u, 2 * j + s, 2 * TP, active ? w : 0, gbase, g.SI, g, sym);
    __syncthreads();                                                 // shared memory is reused by the backward transform

    line_fft<T, H, R0, R1, 1, 16, +1, 2 * TX, false, 0>(u, j, ls, sm, tw, 0);       // u[m] = y_s[j + TP m]

    // ---- swap back and radix-2 level (decimation in time) ----------------------------------------------------------
    C rn[8];                                                         // the eight roots first: one latency, not eight
#pragma unroll
    for (int i = 0; i < 8; ++i) rn[i] = __ldg(&rt[j + TP * (8 * s + i)]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const C send = s ? u[i] : u[i + 8];                          // s = 0 gives away yA[8+i], s = 1 gives away yB[i]
        const C recv = shfl_xor_c(send, TX);
        const C ya = s ? recv : u[i];
        const C yb = s ? u[i + 8] : recv;
        const int n = j + TP * (8 * s + i);
        const C tt = cmulc(yb, rn[i]);                               // conj(W^n) yB[n]
        const C x0 = cadd(ya, tt), x1 = csub(ya, tt);
        if (active) {
            if (!GEN) {
                out[obase + (long long)n * g.SIo] = x0;
                out[obase + (long long)(n + H) * g.SIo] = x1;
            } else {
                *gen_out_ptr<C>(g, obase, n) = x0;
                *gen_out_ptr<C>(g, obase, n + H) = x1;
            }
        }
    }
}

}  // namespace cpc

namespace cpc {

// ---------------------------------------------------------------------------------------------------------------
// Contiguous 512-point lines (the scalar x pass): ONE WARP PER LINE, no block-wide barrier at all.
// Lane L loads the pairs (k, k+256) for k = L + 32 i (a warp reads 512 contiguous bytes per instruction), does the
// radix-2 level, swaps half of its points with lane L ^ 16, and then lanes 0-15 transform half-line A, lanes 16-31
// half-line B with the 16 x 16 Stockham code; the exchange between its two stages goes through a warp-private
// 8 KB shared-memory slab (XOR-swizzled 16-byte chunks, conflict-free for both access patterns) fenced by
// __syncwarp only.  Plain transforms only (MODE_FWD / MODE_INV), decimation in frequency in both directions.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int MODE, int LINES>
__global__ void __launch_bounds__(32 * LINES, 2)
fft_r2x512_line_kernel(const cplx_t<T> *in, cplx_t<T> *out, const PassGeom g,
                       const cplx_t<T> *__restrict__ tw, const SymbolArgs<T> sym)
{
    using C = cplx_t<T>;
    constexpr int H = 256;
    constexpr int D1 = (MODE == MODE_INV) ? +1 : -1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    C *sm = reinterpret_cast<C *>(smem_raw) + (size_t)wrp * 512;
    const int sg = lane >> 4;             // half-line this lane transforms after the swap
    const int j = lane & 15;

    const int t = blockIdx.x;
    const int ti = t % g.tiles_inner, to = t / g.tiles_inner;
    const int w = ti * LINES + wrp;
    const bool active = w < g.lines_inner;
    const long long gbase = (long long)to * g.B1 + (long long)ti * g.B0 + (long long)wrp * g.SL;
    const long long obase = (long long)to * g.B1o + (long long)ti * g.B0o + (long long)wrp * g.SLo;
    const C *rt = sym.rz;

    C u[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = lane + 32 * i;
        C x0 = mk<T>((T)0, (T)0), x1 = x0;
        if (active) {
            x0 = in[gbase + (long long)k * g.SI];
            x1 = in[gbase + (long long)(k + H) * g.SI];
        }
        const C a = cadd(x0, x1);
        const C b = twmul<D1>(csub(x0, x1), __ldg(&rt[k]));
        // lanes 0-15 keep a (point k = j + 16 (2i)) and get the partner's a (k = j + 16 (2i+1)); lanes 16-31 keep b
        // (k = j + 16 (2i+1)) and get the partner's b (k = j + 16 (2i))
        const C recv = shfl_xor_c(sg ? a : b, 16);
        u[2 * i] = sg ? recv : a;
        u[2 * i + 1] = sg ? b : recv;
    }

    Butterfly<16, D1, C>::run(u);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int idx = j * 16 + r;
        sm[sg * H + (idx ^ (j & 7))] = u[r];                      // (idx >> 4) & 7 == j & 7
    }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int idx = j + 16 * m;
        u[m] = sm[sg * H + (idx ^ (m & 7))];                      // (idx >> 4) & 7 == m & 7
    }
#pragma unroll
    for (int r = 1; r < 16; ++r) u[r] = twmul<D1>(u[r], __ldg(&tw[(r - 1) * 16 + j]));
    Butterfly<16, D1, C>::run(u);

    if (active) {
#pragma unroll
        for (int m = 0; m < 16; ++m) out[obase + (long long)(2 * (j + 16 * m) + sg) * g.SIo] = u[m];
    }
}

}  // namespace cpc

namespace cpc {

// ---------------------------------------------------------------------------------------------------------------
// Contiguous 256-point lines (16 x 16), HALF A WARP PER LINE, no block-wide barrier: the x pass of 256-wide grids
// and -- through MODE_R2C / MODE_C2R -- the r2c / c2r x pass of real-scalar plans with nx = 512 (a real line of 512
// points is transformed as 256 complex points and untangled / tangled, see fft_pass.cuh).  Each half warp owns a
// 4 KB XOR-swizzled shared-memory slab used for the exchange between the two radix-16 stages and, in r2c mode, for
// fetching the mirror point Z[256 - k]; everything is fenced by __syncwarp.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int MODE, int LINES>
__global__ void __launch_bounds__(16 * LINES, 2)
fft_line256_kernel(const cplx_t<T> *in, cplx_t<T> *out, const PassGeom g,
                   const cplx_t<T> *__restrict__ tw, const SymbolArgs<T> sym)
{
    using C = cplx_t<T>;
    constexpr int N = 256;
    constexpr int DIR = (MODE == MODE_INV || MODE == MODE_C2R) ? +1 : -1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lw = threadIdx.x >> 4;                 // line within the tile
    const int j = threadIdx.x & 15;
    C *sm = reinterpret_cast<C *>(smem_raw) + (size_t)lw * N;

    const int t = blockIdx.x;
    const int ti = t % g.tiles_inner, to = t / g.tiles_inner;
    const int w = ti * LINES + lw;
    const bool active = w < g.lines_inner;
    const long long gbase = (long long)to * g.B1 + (long long)ti * g.B0 + (long long)lw * g.SL;
    const long long obase = (long long)to * g.B1o + (long long)ti * g.B0o + (long long)lw * g.SLo;

    C u[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) u[m] = active ? in[gbase + (j + 16 * m)] : mk<T>((T)0, (T)0);

    if (MODE == MODE_C2R) {
        // Z[k] = (X[k] + conj X[N-k]) + i conj(w^k) (X[k] - conj X[N-k]),  w = exp(-i pi / N)   (cf. fft_pass.cuh)
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int k = j + 16 * m;
            const C a = u[m];
            const C bq = active ? in[gbase + (N - k)] : mk<T>((T)0, (T)0);
            const C s1 = mk<T>(a.x + bq.x, a.y - bq.y);
            const C d1 = mk<T>(a.x - bq.x, a.y + bq.y);
            const C t1 = cmulc(d1, __ldg(&sym.rx[k]));
            u[m] = mk<T>(s1.x - t1.y, s1.y + t1.x);
        }
    }

    Butterfly<16, DIR, C>::run(u);
#pragma unroll
    for (int r = 0; r < 16; ++r) sm[(j * 16 + r) ^ (j & 7)] = u[r];
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 16; ++m) u[m] = sm[(j + 16 * m) ^ (m & 7)];
#pragma unroll
    for (int r = 1; r < 16; ++r) u[r] = twmul<DIR>(u[r], __ldg(&tw[(r - 1) * 16 + j]));
    Butterfly<16, DIR, C>::run(u);                   // u[m] = point j + 16 m

    if (MODE == MODE_R2C) {
        // X[k] = E[k] + w^k O[k],  E = (Z[k] + conj Z[N-k]) / 2,  O = -i (Z[k] - conj Z[N-k]) / 2;  X[N] = Re Z[0] - Im Z[0]
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 16; ++m) sm[(j + 16 * m) ^ (m & 7)] = u[m];
        __syncwarp();
        C nyq = mk<T>((T)0, (T)0);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int k = j + 16 * m;
            const int km = (N - k) & (N - 1);
            const C a = u[m];
            const C bq = sm[km ^ ((km >> 4) & 7)];
            const C ev = mk<T>((T)0.5 * (a.x + bq.x), (T)0.5 * (a.y - bq.y));
            const C od = mk<T>((T)0.5 * (a.y + bq.y), (T)-0.5 * (a.x - bq.x));
            u[m] = cadd(ev, cmul(od, __ldg(&sym.rx[k])));
            if (k == 0) nyq = mk<T>(a.x - a.y, (T)0);
        }
        if (active && j == 0) out[obase + N] = nyq;
    }

    if (active) {
#pragma unroll
        for (int m = 0; m < 16; ++m) out[obase + (j + 16 * m)] = u[m];
    }
}

}  // namespace cpc
