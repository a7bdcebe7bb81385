// generic_pass.cuh -- any-length fallback pass (same PassGeom / SymbolArgs contract as fft_pass.cuh).
//
// Used for axis lengths the templated Stockham kernels do not cover (non powers of two such as the reference's
// own 10 x 25 x 40 and 50 x 200 test grids, tests/FFTDirectSolver/testFftSolver_3D.py:82-93, and n < 16).
// Same Stockham recurrence with a runtime factor list (prime factors, the 2s merged into radix 8 / 4), ping-pong in
// shared memory.  Stages of radix 2, 3, 4, 5, 7, 8 run the in-register butterflies of fft_core.cuh, one butterfly per
// thread at a time; a prime factor >= 11 falls back to an O(R) sum per output with the root exponent stepped mod n
// in integers.  Coverage path: lengths 2^a and 2^a * 3 up to 1024 never get here.
#pragma once
#include "fft_pass.cuh"

namespace cpc {

#define CPC_MAX_FACTORS 24
struct FactorList {
    int n;
    int nfac;
    int fac[CPC_MAX_FACTORS];
};

// a / d and a % d for 0 <= a < 2^24 with a precomputed float reciprocal (exact after one correction step);
// the hardware has no integer divider and the generic kernel does two of these per output.
__device__ __forceinline__ int fdivmod(int a, int d, float invd, int &rem)
{
    int q = (int)((float)a * invd);
    int r = a - q * d;
    if (r < 0) { --q; r += d; }
    else if (r >= d) { ++q; r -= d; }
    rem = r;
    return q;
}

// Where a stage reads its inputs / writes its outputs: the shared-memory ping-pong buffers [point][lane], or --
// first stage of a transform that starts from HBM, last stage of one that ends there -- the global array itself
// (same addressing as fft_pass.cuh: strided lines, the chunked multi-rank layouts, peer pushes).
template <typename C> struct SmemSrc {
    const C *p; int txsh;
    __device__ __forceinline__ C operator()(int i, int l) const { return p[(i << txsh) + l]; }
};
template <typename C> struct SmemDst {
    C *p; int txsh;
    __device__ __forceinline__ void operator()(int i, int l, C v) const { p[(i << txsh) + l] = v; }
};
template <typename C> struct GlobalSrc {
    const C *in; const PassGeom *g; long long tbase; int lanes_ok;       // lanes l < lanes_ok hold a real line
    __device__ __forceinline__ C operator()(int i, int l) const
    {
        C z; z.x = 0; z.y = 0;
        return l < lanes_ok ? in[tbase + (long long)l * g->SL + point_off(i, g->SI, g->Di, g->shi, g->SCi)] : z;
    }
};
template <typename C> struct GlobalDst {
    C *out; const PassGeom *g; long long obase; int lanes_ok;
    __device__ __forceinline__ void operator()(int i, int l, C v) const
    {
        if (l < lanes_ok) *out_ptr<C>(out, *g, obase + (long long)l * g->SLo, i) = v;
    }
};

// One Stockham stage with a compile-time radix: each thread does whole butterflies (R loads, R - 1 twiddles,
// the in-register butterfly of fft_core.cuh, R stores).  Butterfly jb = m p + k reads x[jb + r n/R], multiplies by
// exp(-/+ 2 pi i r k / (p R)) and writes y[(m R + q) p + k].
template <typename T, int DIR, int R, typename Src, typename Dst>
__device__ __forceinline__ void generic_stage_butterfly(const Src &src, const Dst &dst, int n, int p, int txsh,
                                                        const cplx_t<T> *__restrict__ tw)
{
    using C = cplx_t<T>;
    const int TX = 1 << txsh;
    const int nR = n / R;
    const int tws = nR / p;                      // n / (p R): table stride of this stage's roots
    const float inv_p = 1.0f / (float)p;
    for (int idx = threadIdx.x; idx < nR * TX; idx += blockDim.x) {
        const int jb = idx >> txsh, l = idx & (TX - 1);
        int k;
        const int m = fdivmod(jb, p, inv_p, k);
        C u[R];
#pragma unroll
        for (int r = 0; r < R; ++r) u[r] = src(jb + r * nR, l);
        if (p > 1) {
            const int e1 = k * tws;              // < n / R, so r * e1 < n: no reduction mod n
#pragma unroll
            for (int r = 1; r < R; ++r) u[r] = twmul<DIR>(u[r], __ldg(&tw[r * e1]));
        }
        Butterfly<R, DIR, C>::run(u);
        const int o0 = (m * R) * p + k;
#pragma unroll
        for (int q = 0; q < R; ++q) dst(o0 + q * p, l, u[q]);
    }
}

// Any other radix (primes >= 11, and the radix-1 "copy" of an axis of length 1): each thread produces one output
// with an O(R) sum.
template <typename T, int DIR, typename Src, typename Dst>
__device__ __forceinline__ void generic_stage_sum(const Src &src, const Dst &dst, int n, int p, int R, int txsh,
                                                  const cplx_t<T> *__restrict__ tw)
{
    using C = cplx_t<T>;
    const int TX = 1 << txsh;
    const int nR = n / R;
    const int tws = nR / p;
    const float inv_p = 1.0f / (float)p, inv_R = 1.0f / (float)R;
    for (int idx = threadIdx.x; idx < n * TX; idx += blockDim.x) {
        // output o = (m R + q) p + k of butterfly jb = m p + k
        const int o = idx >> txsh, l = idx & (TX - 1);
        int k, q;
        const int t = fdivmod(o, p, inv_p, k);
        const int m = fdivmod(t, R, inv_R, q);
        const int jb = m * p + k;
        // root exponent of term r is r (k tws + q nR) mod n: one add and a conditional subtract per term
        int estep = k * tws + q * nR;
        if (estep >= n) estep -= n;
        int e = 0;
        C acc = src(jb, l);
        for (int r = 1; r < R; ++r) {
            e += estep;
            if (e >= n) e -= n;
            acc = cadd(acc, twmul<DIR>(src(jb + r * nR, l), tw[e]));
        }
        dst(o, l, acc);
    }
}

template <typename T, int DIR, typename Src, typename Dst>
__device__ __forceinline__ void generic_stage(int R, const Src &src, const Dst &dst, int n, int p, int txsh,
                                              const cplx_t<T> *__restrict__ tw)
{
    switch (R) {
    case 2: generic_stage_butterfly<T, DIR, 2>(src, dst, n, p, txsh, tw); break;
    case 3: generic_stage_butterfly<T, DIR, 3>(src, dst, n, p, txsh, tw); break;
    case 4: generic_stage_butterfly<T, DIR, 4>(src, dst, n, p, txsh, tw); break;
    case 5: generic_stage_butterfly<T, DIR, 5>(src, dst, n, p, txsh, tw); break;
    case 7: generic_stage_butterfly<T, DIR, 7>(src, dst, n, p, txsh, tw); break;
    case 8: generic_stage_butterfly<T, DIR, 8>(src, dst, n, p, txsh, tw); break;
    default: generic_stage_sum<T, DIR>(src, dst, n, p, R, txsh, tw); break;
    }
}

// Full transform of the tile's lines.  from_global: the first stage reads HBM (through gs) instead of `cur`;
// to_global: the last stage writes HBM (through gd).  Otherwise the data starts in `cur` / ends in the returned
// buffer.  The factor list is never empty (an axis of length 1 carries the single factor 1).
template <typename T, int DIR>
__device__ __forceinline__ cplx_t<T> *generic_line_fft(cplx_t<T> *cur, cplx_t<T> *nxt, const FactorList &f, int txsh,
                                                       const cplx_t<T> *__restrict__ tw, bool from_global,
                                                       const GlobalSrc<cplx_t<T>> &gs, bool to_global,
                                                       const GlobalDst<cplx_t<T>> &gd)
{
    using C = cplx_t<T>;
    const int n = f.n;
    int p = 1;
    for (int s = 0; s < f.nfac; ++s) {
        const int R = f.fac[s];
        const bool first = from_global && s == 0, last = to_global && s == f.nfac - 1;
        const SmemSrc<C> ss{ cur, txsh };
        const SmemDst<C> sd{ first ? cur : nxt, txsh };
        if (first && last) generic_stage<T, DIR>(R, gs, gd, n, p, txsh, tw);
        else if (first) generic_stage<T, DIR>(R, gs, sd, n, p, txsh, tw);
        else if (last) generic_stage<T, DIR>(R, ss, gd, n, p, txsh, tw);
        else generic_stage<T, DIR>(R, ss, sd, n, p, txsh, tw);
        if (!last) __syncthreads();
        if (!first) { C *tmp = cur; cur = nxt; nxt = tmp; }
        p *= R;
    }
    return cur;   // buffer holding the result (meaningless after a to_global transform)
}

template <typename T>
__global__ void generic_pass_kernel(const cplx_t<T> *in, cplx_t<T> *out, const PassGeom g,
                                    const cplx_t<T> *__restrict__ tw, const SymbolArgs<T> s, const FactorList f,
                                    const int TX, const int mode)
{
    using C = cplx_t<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = f.n;
    C *A = reinterpret_cast<C *>(smem_raw);
    C *B = A + (size_t)n * TX;

    const int txsh = __ffs(TX) - 1;          // TX is a power of two (plan_impl.cuh)
    const int t = blockIdx.x;
    const int ti = t % g.tiles_inner, to = t / g.tiles_inner;
    const long long tbase = (long long)to * g.B1 + (long long)ti * g.B0;
    const long long obase = (long long)to * g.B1o + (long long)ti * g.B0o;
    const int lanes_ok = g.lines_inner - ti * TX;            // >= TX for a full tile
    const GlobalSrc<C> gs{ in, &g, tbase, lanes_ok };
    const GlobalDst<C> gd{ out, &g, obase, lanes_ok };

    if (mode == MODE_FWD) {
        generic_line_fft<T, -1>(A, B, f, txsh, tw, true, gs, true, gd);
    } else if (mode == MODE_INV) {
        generic_line_fft<T, +1>(A, B, f, txsh, tw, true, gs, true, gd);
    } else {
        C *res = generic_line_fft<T, -1>(A, B, f, txsh, tw, true, gs, false, gd);
        C *oth = (res == A) ? B : A;
        for (int idx = threadIdx.x; idx < n * TX; idx += blockDim.x) {
            const int k = idx >> txsh, l = idx & (TX - 1);
            const int wr = ti * TX + l;
            const int w = wr < g.lines_inner ? wr : g.lines_inner - 1;   // clamp: masked lanes must not index tables
            C v = res[idx];
            if (mode == MODE_FUSED_SEP) {
                const int x = w % g.nx, y = w / g.nx + g.y0;
                const C lam = cadd(cadd(s.ax[x], s.ay[y]), s.az[k]);
                v = cmul(v, crecip_scaled<T>(lam, s.scale));
            } else if (mode == MODE_FUSED_TABLE) {
                const bool ok = wr < g.lines_inner;
                if (ok) v = cmul(v, s.inv_table[tbase + (long long)l * g.SL + (long long)k * g.SI]);
            } else {   // MODE_FUSED_WAVE: same closed form as fft_pass.cuh, operands read from shared memory
                const int c = w & 3;
                const int cell = w >> 2;
                const int x = cell % g.nx, y = cell / g.nx + g.y0;
                const C rx = s.rx[x], ry = s.ry[y], rz = s.rz[k];
                const T sd[3] = { -rx.y * s.mux, -ry.y * s.muy, -rz.y * s.muz };
                const T od[3] = { ((T)1 - rx.x) * s.mux, ((T)1 - ry.x) * s.muy, ((T)1 - rz.x) * s.muz };
                const T c0 = s.c0, c02 = s.c0 * s.c0;
                T D[3], den = (T)1 + c0 * (od[0] + od[1] + od[2]);
                for (int d = 0; d < 3; ++d) { D[d] = (T)1 + c0 * od[d]; den += c02 * sd[d] * sd[d] / D[d]; }
                const int l0 = l & ~3;
                C num = res[k * TX + l0];
                for (int d = 0; d < 3; ++d) {
                    const C rd = res[k * TX + l0 + 1 + d];
                    const T fct = c02 * sd[d] / D[d];
                    num.x += rd.y * fct;       // -i * fct * rd
                    num.y -= rd.x * fct;
                }
                const T inv = s.scale / den;
                const C p = mk<T>(num.x * inv, num.y * inv);
                if (c == 0) v = p;
                else {
                    const T qd = (T)1 / D[c - 1];
                    v = mk<T>((v.x * s.scale + sd[c - 1] * p.y) * qd, (v.y * s.scale - sd[c - 1] * p.x) * qd);
                }
            }
            oth[idx] = v;
        }
        __syncthreads();
        generic_line_fft<T, +1>(oth, res, f, txsh, tw, false, gs, true, gd);
    }
}

}  // namespace cpc
