// generic_pass.cuh -- any-length fallback pass (same PassGeom / SymbolArgs contract as fft_pass.cuh).
//
// Used for axis lengths the templated Stockham kernels do not cover (non powers of two such as the reference's
// own 10 x 25 x 40 and 50 x 200 test grids, tests/FFTDirectSolver/testFftSolver_3D.py:82-93, and n < 16).
// Same Stockham recurrence with a runtime factor list (prime factors, pairs of 2 merged into radix 4); each thread
// produces one output of one butterfly with an O(R) sum, twiddle exponents stepped mod n in integers, ping-pong in
// shared memory.  Coverage path: O(n * sum of factors) per line; lengths 2^a and 2^a * 3 up to 1024 never get here.
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

template <typename T, int DIR>
__device__ __forceinline__ cplx_t<T> *generic_line_fft(cplx_t<T> *src, cplx_t<T> *dst, const FactorList &f, int txsh,
                                                       const cplx_t<T> *__restrict__ tw)
{
    using C = cplx_t<T>;
    const int n = f.n;
    const int TX = 1 << txsh;
    int p = 1;
    for (int s = 0; s < f.nfac; ++s) {
        const int R = f.fac[s];
        const int nR = n / R;
        const int tws = nR / p;                  // n / (p R)
        const float inv_p = 1.0f / (float)p, inv_R = 1.0f / (float)R;
        for (int idx = threadIdx.x; idx < n * TX; idx += blockDim.x) {
            // output o = (m R + q) p + k of butterfly jb = m p + k
            const int o = idx >> txsh, l = idx & (TX - 1);
            int k, q;
            const int t = fdivmod(o, p, inv_p, k);
            const int m = fdivmod(t, R, inv_R, q);
            const C *sp = src + (m * p + k) * TX + l;
            // root exponent of term r is r (k tws + q nR) mod n: one add and a conditional subtract per term
            int estep = k * tws + q * nR;
            if (estep >= n) estep -= n;
            int e = 0;
            C acc = sp[0];
            for (int r = 1; r < R; ++r) {
                e += estep;
                if (e >= n) e -= n;
                acc = cadd(acc, twmul<DIR>(sp[(size_t)r * nR * TX], tw[e]));
            }
            dst[idx] = acc;
        }
        __syncthreads();
        C *tmp = src; src = dst; dst = tmp;
        p *= R;
    }
    return src;   // buffer holding the result
}

template <typename T>
__global__ void generic_pass_kernel(const cplx_t<T> *__restrict__ in, cplx_t<T> *__restrict__ out, const PassGeom g,
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

    for (int idx = threadIdx.x; idx < n * TX; idx += blockDim.x) {
        const int i = idx >> txsh, l = idx & (TX - 1);
        const bool ok = (ti * TX + l) < g.lines_inner;
        A[idx] = ok ? in[tbase + (long long)l * g.SL + point_off(i, g.SI, g.Di, g.shi, g.SCi)] : mk<T>((T)0, (T)0);
    }
    __syncthreads();

    C *res;
    if (mode == MODE_FWD) {
        res = generic_line_fft<T, -1>(A, B, f, txsh, tw);
    } else if (mode == MODE_INV) {
        res = generic_line_fft<T, +1>(A, B, f, txsh, tw);
    } else {
        res = generic_line_fft<T, -1>(A, B, f, txsh, tw);
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
        res = generic_line_fft<T, +1>(oth, res, f, txsh, tw);
    }

    for (int idx = threadIdx.x; idx < n * TX; idx += blockDim.x) {
        const int i = idx >> txsh, l = idx & (TX - 1);
        if ((ti * TX + l) < g.lines_inner)
            *out_ptr<C>(out, g, obase + (long long)l * g.SLo, i) = res[idx];
    }
}

}  // namespace cpc
