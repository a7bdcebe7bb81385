// generic_pass.cuh -- any-length fallback pass (same PassGeom / SymbolArgs contract as fft_pass.cuh).
//
// Used for axis lengths the templated Stockham kernels do not cover (non powers of two such as the reference's
// own 10 x 25 x 40 and 50 x 200 test grids, tests/FFTDirectSolver/testFftSolver_3D.py:82-93, and n < 16).
// Same Stockham recurrence with a runtime factor list (smallest-prime-factor order); each thread produces one
// output of one butterfly with an O(R) sum, twiddle exponents reduced mod n in integers, ping-pong in shared
// memory.  Correctness path, not a performance path: O(n * sum of factors) per line.
#pragma once
#include "fft_pass.cuh"

namespace cpc {

#define CPC_MAX_FACTORS 24
struct FactorList {
    int n;
    int nfac;
    int fac[CPC_MAX_FACTORS];
};

template <typename T, int DIR>
__device__ __forceinline__ cplx_t<T> *generic_line_fft(cplx_t<T> *src, cplx_t<T> *dst, const FactorList &f, int TX,
                                                       const cplx_t<T> *__restrict__ tw)
{
    using C = cplx_t<T>;
    const int n = f.n;
    int p = 1;
    for (int s = 0; s < f.nfac; ++s) {
        const int R = f.fac[s];
        const int nR = n / R;
        const int tws = n / (p * R);
        for (int idx = threadIdx.x; idx < n * TX; idx += blockDim.x) {
            const int o = idx / TX, l = idx - o * TX;
            const int k = o % p;
            const int q = (o / p) % R;
            const int jb = (o / (p * R)) * p + k;
            C acc = mk<T>((T)0, (T)0);
            for (int r = 0; r < R; ++r) {
                const long long e = ((long long)r * k * tws + (long long)r * q * nR) % n;
                acc = cadd(acc, twmul<DIR>(src[(jb + r * nR) * TX + l], tw[e]));
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

    const int t = blockIdx.x;
    const int ti = t % g.tiles_inner, to = t / g.tiles_inner;
    const long long tbase = (long long)to * g.B1 + (long long)ti * g.B0;
    const long long obase = (long long)to * g.B1o + (long long)ti * g.B0o;

    for (int idx = threadIdx.x; idx < n * TX; idx += blockDim.x) {
        const int i = idx / TX, l = idx - i * TX;
        const bool ok = (ti * TX + l) < g.lines_inner;
        A[idx] = ok ? in[tbase + (long long)l * g.SL + point_off(i, g.SI, g.Di, g.shi, g.SCi)] : mk<T>((T)0, (T)0);
    }
    __syncthreads();

    C *res;
    if (mode == MODE_FWD) {
        res = generic_line_fft<T, -1>(A, B, f, TX, tw);
    } else if (mode == MODE_INV) {
        res = generic_line_fft<T, +1>(A, B, f, TX, tw);
    } else {
        res = generic_line_fft<T, -1>(A, B, f, TX, tw);
        C *oth = (res == A) ? B : A;
        for (int idx = threadIdx.x; idx < n * TX; idx += blockDim.x) {
            const int k = idx / TX, l = idx - k * TX;
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
        res = generic_line_fft<T, +1>(oth, res, f, TX, tw);
    }

    for (int idx = threadIdx.x; idx < n * TX; idx += blockDim.x) {
        const int i = idx / TX, l = idx - i * TX;
        if ((ti * TX + l) < g.lines_inner)
            *out_ptr<C>(out, g, obase + (long long)l * g.SLo, i) = res[idx];
    }
}

}  // namespace cpc
