// fft_core.cuh -- in-register butterflies and complex helpers for the sm_100a Stockham passes.
//
// All transforms are the unnormalised DFT of PETSc MATFFTW / FFTW3 that the reference calls through
// MatMult / MatMultTranspose (reference src/FftLinearSolver_3D.c:170,180):
//   DIR = -1 : X[q] = sum_r x[r] exp(-2 pi i r q / R)   (FFTW_FORWARD)
//   DIR = +1 : X[q] = sum_r x[r] exp(+2 pi i r q / R)   (FFTW_BACKWARD)
// Outputs are in natural order.  Arithmetic is plain fp64 (or fp32) FMA; no tensor cores: the apply is
// HBM-bound, not a dense contraction (BASELINE.json north_star).
#pragma once
#include <cuda_runtime.h>

namespace cpc {

template <typename T> struct cplx_of;
template <> struct cplx_of<double> { using type = double2; };
template <> struct cplx_of<float> { using type = float2; };
template <typename T> using cplx_t = typename cplx_of<T>::type;

template <typename T> __host__ __device__ __forceinline__ cplx_t<T> mk(T a, T b) { cplx_t<T> r; r.x = a; r.y = b; return r; }
template <typename C> __host__ __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __host__ __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
// a * b
template <typename C> __host__ __device__ __forceinline__ C cmul(C a, C b)
{
    C r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
// a * conj(b)
template <typename C> __host__ __device__ __forceinline__ C cmulc(C a, C b)
{
    C r;
    r.x = a.x * b.x + a.y * b.y;
    r.y = a.y * b.x - a.x * b.y;
    return r;
}
// multiply by the table root w = exp(-2 pi i e / n): forward uses w, backward uses conj(w)
template <int DIR, typename C> __host__ __device__ __forceinline__ C twmul(C a, C w)
{
    return DIR < 0 ? cmul(a, w) : cmulc(a, w);
}
// multiply by DIR * i  (forward: -i, backward: +i)
template <int DIR, typename C> __host__ __device__ __forceinline__ C mul_dir_i(C a)
{
    C r;
    if (DIR < 0) { r.x = a.y; r.y = -a.x; }
    else         { r.x = -a.y; r.y = a.x; }
    return r;
}

template <int R, int DIR, typename C> struct Butterfly;

template <int DIR, typename C> struct Butterfly<1, DIR, C> {
    __host__ __device__ __forceinline__ static void run(C *) {}
};

template <int DIR, typename C> struct Butterfly<2, DIR, C> {
    __host__ __device__ __forceinline__ static void run(C *u)
    {
        C a = u[0], b = u[1];
        u[0] = cadd(a, b);
        u[1] = csub(a, b);
    }
};

template <int DIR, typename C> struct Butterfly<4, DIR, C> {
    __host__ __device__ __forceinline__ static void run(C *u)
    {
        C t0 = cadd(u[0], u[2]), t1 = csub(u[0], u[2]);
        C t2 = cadd(u[1], u[3]), t3 = mul_dir_i<DIR>(csub(u[1], u[3]));
        u[0] = cadd(t0, t2);
        u[2] = csub(t0, t2);
        u[1] = cadd(t1, t3);
        u[3] = csub(t1, t3);
    }
};

template <int DIR, typename C> struct Butterfly<8, DIR, C> {
    __host__ __device__ __forceinline__ static void run(C *u)
    {
        using T = decltype(u[0].x);
        const T h = (T)0.70710678118654752440084436210484903928;
        C e[4] = { u[0], u[2], u[4], u[6] };
        C o[4] = { u[1], u[3], u[5], u[7] };
        Butterfly<4, DIR, C>::run(e);
        Butterfly<4, DIR, C>::run(o);
        // X[k] = E[k] + W8^k O[k], X[k+4] = E[k] - W8^k O[k];  W8^1 = (1 + DIR i)/sqrt2, W8^2 = DIR i,
        // W8^3 = (-1 + DIR i)/sqrt2.  The 1/sqrt2 factor is folded into the final add as an FMA:
        // (a +- b) h + e  ->  fma(a +- b, h, e): 2 adds + 4 FMAs per pair instead of 2 adds + 2 muls + 4 adds.
        T p1, q1, p3, q3;     // W8^1 O[1] = (p1, q1) h,  W8^3 O[3] = (p3, q3) h
        if (DIR < 0) {
            p1 = o[1].x + o[1].y;  q1 = o[1].y - o[1].x;
            p3 = o[3].y - o[3].x;  q3 = -(o[3].x + o[3].y);
        } else {
            p1 = o[1].x - o[1].y;  q1 = o[1].x + o[1].y;
            p3 = -(o[3].x + o[3].y); q3 = o[3].x - o[3].y;
        }
        const C w2 = mul_dir_i<DIR>(o[2]);
        u[0] = cadd(e[0], o[0]); u[4] = csub(e[0], o[0]);
        u[1].x = fma(p1, h, e[1].x);  u[1].y = fma(q1, h, e[1].y);
        u[5].x = fma(-p1, h, e[1].x); u[5].y = fma(-q1, h, e[1].y);
        u[2] = cadd(e[2], w2);   u[6] = csub(e[2], w2);
        u[3].x = fma(p3, h, e[3].x);  u[3].y = fma(q3, h, e[3].y);
        u[7].x = fma(-p3, h, e[3].x); u[7].y = fma(-q3, h, e[3].y);
    }
};

// Odd-radix butterflies for line lengths 2^a * 3 (96, 192, 384, 768).  Radix 6 = 2 x 3 and radix 12 = 3 x 4 use the
// prime-factor (Good-Thomas) index maps, so there are no internal twiddles.
template <int DIR, typename C> struct Butterfly<3, DIR, C> {
    __host__ __device__ __forceinline__ static void run(C *u)
    {
        using T = decltype(u[0].x);
        const T s = (T)0.86602540378443864676372317075293618347;      // sin(2 pi / 3)
        const C t1 = cadd(u[1], u[2]);
        const C d = mul_dir_i<DIR>(csub(u[1], u[2]));                  // DIR i (x1 - x2)
        C t2;
        t2.x = fma((T)-0.5, t1.x, u[0].x);
        t2.y = fma((T)-0.5, t1.y, u[0].y);
        u[0] = cadd(u[0], t1);
        u[1].x = fma(s, d.x, t2.x);  u[1].y = fma(s, d.y, t2.y);
        u[2].x = fma(-s, d.x, t2.x); u[2].y = fma(-s, d.y, t2.y);
    }
};

// n = (3 n1 + 2 n2) mod 6, k = (3 k1 + 4 k2) mod 6
template <int DIR, typename C> struct Butterfly<6, DIR, C> {
    __host__ __device__ __forceinline__ static void run(C *u)
    {
        C a[3] = { cadd(u[0], u[3]), cadd(u[2], u[5]), cadd(u[4], u[1]) };
        C b[3] = { csub(u[0], u[3]), csub(u[2], u[5]), csub(u[4], u[1]) };
        Butterfly<3, DIR, C>::run(a);
        Butterfly<3, DIR, C>::run(b);
        u[0] = a[0]; u[4] = a[1]; u[2] = a[2];
        u[3] = b[0]; u[1] = b[1]; u[5] = b[2];
    }
};

// n = (4 n1 + 3 n2) mod 12, k = (4 k1 + 9 k2) mod 12
template <int DIR, typename C> struct Butterfly<12, DIR, C> {
    __host__ __device__ __forceinline__ static void run(C *u)
    {
        C s0[4] = { u[0], u[3], u[6], u[9] };
        C s1[4] = { u[4], u[7], u[10], u[1] };
        C s2[4] = { u[8], u[11], u[2], u[5] };
        Butterfly<4, DIR, C>::run(s0);
        Butterfly<4, DIR, C>::run(s1);
        Butterfly<4, DIR, C>::run(s2);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
            C t[3] = { s0[k2], s1[k2], s2[k2] };
            Butterfly<3, DIR, C>::run(t);
#pragma unroll
            for (int k1 = 0; k1 < 3; ++k1) u[(4 * k1 + 9 * k2) % 12] = t[k1];
        }
    }
};

// Odd primes 5 and 7 (generic any-length kernel): with t_j = x_j + x_{R-j}, d_j = DIR i (x_j - x_{R-j}),
//   X[q] = x_0 + sum_j cos(2 pi j q / R) t_j + sum_j sin(2 pi j q / R) d_j,   X[R-q] = the same with -sin.
template <int R> struct PrimeRoots;
template <> struct PrimeRoots<5> {
    __host__ __device__ static constexpr double c(int m) { return m == 1 ? 0.30901699437494742410 : -0.80901699437494742410; }
    __host__ __device__ static constexpr double s(int m) { return m == 1 ? 0.95105651629515357212 : 0.58778525229247312917; }
};
template <> struct PrimeRoots<7> {
    __host__ __device__ static constexpr double c(int m)
    {
        return m == 1 ? 0.62348980185873353053 : m == 2 ? -0.22252093395631440429 : -0.90096886790241912624;
    }
    __host__ __device__ static constexpr double s(int m)
    {
        return m == 1 ? 0.78183148246802980871 : m == 2 ? 0.97492791218182360702 : 0.43388373911755812048;
    }
};
template <int R, int DIR, typename C> struct OddPrimeButterfly {
    __host__ __device__ __forceinline__ static void run(C *u)
    {
        using T = decltype(u[0].x);
        constexpr int H = (R - 1) / 2;
        C t[H], d[H];
#pragma unroll
        for (int j = 1; j <= H; ++j) {
            t[j - 1] = cadd(u[j], u[R - j]);
            d[j - 1] = mul_dir_i<DIR>(csub(u[j], u[R - j]));
        }
        const C x0 = u[0];
        C sum = x0;
#pragma unroll
        for (int j = 0; j < H; ++j) sum = cadd(sum, t[j]);
        u[0] = sum;
#pragma unroll
        for (int q = 1; q <= H; ++q) {
            C a = x0, b;
            b.x = (T)0; b.y = (T)0;
#pragma unroll
            for (int j = 1; j <= H; ++j) {
                const int m = (j * q) % R;
                const T cm = (T)(m <= H ? PrimeRoots<R>::c(m) : PrimeRoots<R>::c(R - m));
                const T sm = (T)(m <= H ? PrimeRoots<R>::s(m) : -PrimeRoots<R>::s(R - m));
                a.x = fma(cm, t[j - 1].x, a.x); a.y = fma(cm, t[j - 1].y, a.y);
                b.x = fma(sm, d[j - 1].x, b.x); b.y = fma(sm, d[j - 1].y, b.y);
            }
            u[q] = cadd(a, b);
            u[R - q] = csub(a, b);
        }
    }
};
template <int DIR, typename C> struct Butterfly<5, DIR, C> : OddPrimeButterfly<5, DIR, C> {};
template <int DIR, typename C> struct Butterfly<7, DIR, C> : OddPrimeButterfly<7, DIR, C> {};

// Prime-factor (Good-Thomas) butterfly for R = A * B with gcd(A, B) = 1: no internal twiddles.
//   inputs  n = (B n1 + A n2) mod R,   outputs k = (cb k1 + ca k2) mod R,   cb = B (B^-1 mod A), ca = A (A^-1 mod B)
__host__ __device__ constexpr int mod_inverse(int a, int m)
{
    int r = 1;
    for (int i = 1; i < m; ++i)
        if ((a * i) % m == 1) r = i;
    return r;
}
template <int A, int B, int DIR, typename C> struct PfaButterfly {
    __host__ __device__ __forceinline__ static void run(C *u)
    {
        constexpr int R = A * B;
        constexpr int cb = B * mod_inverse(B % A, A), ca = A * mod_inverse(A % B, B);
        C s[A][B];
#pragma unroll
        for (int n1 = 0; n1 < A; ++n1) {
#pragma unroll
            for (int n2 = 0; n2 < B; ++n2) s[n1][n2] = u[(B * n1 + A * n2) % R];
            Butterfly<B, DIR, C>::run(s[n1]);
        }
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) {
            C t[A];
#pragma unroll
            for (int n1 = 0; n1 < A; ++n1) t[n1] = s[n1][k2];
            Butterfly<A, DIR, C>::run(t);
#pragma unroll
            for (int k1 = 0; k1 < A; ++k1) u[(cb * k1 + ca * k2) % R] = t[k1];
        }
    }
};
// line lengths 2^a * 5^b (100, 160, 200, 250, 320, 400, 500, 800, 1000)
template <int DIR, typename C> struct Butterfly<10, DIR, C> : PfaButterfly<2, 5, DIR, C> {};
template <int DIR, typename C> struct Butterfly<20, DIR, C> : PfaButterfly<4, 5, DIR, C> {};

// plus = a + w' b, minus = a - w' b with w' = w (forward) or conj(w) (backward): 6 FMAs instead of a complex
// multiply (4) plus an add and a subtract (4).  minus = 2a - plus.
template <int DIR, typename C> __host__ __device__ __forceinline__ void cfma_pm(C a, C w, C b, C &plus, C &minus)
{
    using T = decltype(a.x);
    const T wy = DIR < 0 ? w.y : -w.y;
    plus.x = fma(w.x, b.x, fma(-wy, b.y, a.x));
    plus.y = fma(w.x, b.y, fma(wy, b.x, a.y));
    minus.x = fma((T)2, a.x, -plus.x);
    minus.y = fma((T)2, a.y, -plus.y);
}

// Radix-8 butterfly of the twiddled inputs u[r] * w[r-1] (r = 1..7): the twiddles are folded into the first add /
// subtract level (36 instead of 44 fp64 instructions for that level).
template <int DIR, typename C> __host__ __device__ __forceinline__ void butterfly8_twiddled(C *u, const C *w)
{
    using T = decltype(u[0].x);
    const T h = (T)0.70710678118654752440084436210484903928;
    C t0, t1, t2, t3, s0, s1, s2, s3;
    cfma_pm<DIR>(u[0], w[3], u[4], t0, t1);                          // u0 +- w4 u4
    cfma_pm<DIR>(twmul<DIR>(u[2], w[1]), w[5], u[6], t2, t3);        // w2 u2 +- w6 u6
    cfma_pm<DIR>(twmul<DIR>(u[1], w[0]), w[4], u[5], s0, s1);        // w1 u1 +- w5 u5
    cfma_pm<DIR>(twmul<DIR>(u[3], w[2]), w[6], u[7], s2, s3);        // w3 u3 +- w7 u7
    t3 = mul_dir_i<DIR>(t3);
    s3 = mul_dir_i<DIR>(s3);
    const C e0 = cadd(t0, t2), e2 = csub(t0, t2), e1 = cadd(t1, t3), e3 = csub(t1, t3);
    const C o0 = cadd(s0, s2), o2 = csub(s0, s2), o1 = cadd(s1, s3), o3 = csub(s1, s3);
    T p1, q1, p3, q3;
    if (DIR < 0) {
        p1 = o1.x + o1.y;  q1 = o1.y - o1.x;
        p3 = o3.y - o3.x;  q3 = -(o3.x + o3.y);
    } else {
        p1 = o1.x - o1.y;  q1 = o1.x + o1.y;
        p3 = -(o3.x + o3.y); q3 = o3.x - o3.y;
    }
    const C w2 = mul_dir_i<DIR>(o2);
    u[0] = cadd(e0, o0); u[4] = csub(e0, o0);
    u[1].x = fma(p1, h, e1.x);  u[1].y = fma(q1, h, e1.y);
    u[5].x = fma(-p1, h, e1.x); u[5].y = fma(-q1, h, e1.y);
    u[2] = cadd(e2, w2);     u[6] = csub(e2, w2);
    u[3].x = fma(p3, h, e3.x);  u[3].y = fma(q3, h, e3.y);
    u[7].x = fma(-p3, h, e3.x); u[7].y = fma(-q3, h, e3.y);
}

template <int DIR, typename C> struct Butterfly<16, DIR, C> {
    __host__ __device__ __forceinline__ static void run(C *u)
    {
        using T = decltype(u[0].x);
        const T c1 = (T)0.92387953251128675612818318939678828682;  // cos(pi/8)
        const T s1 = (T)0.38268343236508977172845998403039886676;  // sin(pi/8)
        const T h = (T)0.70710678118654752440084436210484903928;
        C a[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int m = 0; m < 4; ++m) a[r][m] = u[r + 4 * m];
            Butterfly<4, DIR, C>::run(a[r]);
        }
        // twiddle a[r][k] *= W16^{r k}; W16^m = exp(DIR * 2 pi i m / 16)
        const C w1 = mk<T>(c1, (T)DIR * s1), w2 = mk<T>(h, (T)DIR * h), w3 = mk<T>(s1, (T)DIR * c1);
        a[1][1] = cmul(a[1][1], w1);
        a[1][2] = cmul(a[1][2], w2);
        a[1][3] = cmul(a[1][3], w3);
        a[2][1] = cmul(a[2][1], w2);
        a[2][2] = mul_dir_i<DIR>(a[2][2]);                                   // W16^4
        a[2][3] = cmul(a[2][3], mk<T>(-h, (T)DIR * h));                      // W16^6
        a[3][1] = cmul(a[3][1], w3);
        a[3][2] = cmul(a[3][2], mk<T>(-h, (T)DIR * h));                      // W16^6
        a[3][3] = cmul(a[3][3], mk<T>(-c1, (T)(-DIR) * s1));                 // W16^9
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            C b[4] = { a[0][k], a[1][k], a[2][k], a[3][k] };
            Butterfly<4, DIR, C>::run(b);
#pragma unroll
            for (int q = 0; q < 4; ++q) u[k + 4 * q] = b[q];
        }
    }
};

// 1/d for a positive, normal d.  fp64: hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps = 4 DFMA, instead
// of the IEEE division sequence with its special-case branch; the result is within 1 ulp, which is all the
// eigenvalue division needs (|Lambda|^2 >= 1 for the transport symbol, reference FftLinearSolver_3D.c:146-157).
__device__ __forceinline__ double fast_rcp(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ float fast_rcp(float d) { return 1.0f / d; }

// complex reciprocal scaled: s / z
template <typename T> __device__ __forceinline__ cplx_t<T> crecip_scaled(cplx_t<T> z, T s)
{
    T d = z.x * z.x + z.y * z.y;
    T inv = s * fast_rcp(d);
    return mk<T>(z.x * inv, -z.y * inv);
}

}  // namespace cpc
