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
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
// a * b
template <typename C> __device__ __forceinline__ C cmul(C a, C b)
{
    C r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
// a * conj(b)
template <typename C> __device__ __forceinline__ C cmulc(C a, C b)
{
    C r;
    r.x = a.x * b.x + a.y * b.y;
    r.y = a.y * b.x - a.x * b.y;
    return r;
}
// multiply by the table root w = exp(-2 pi i e / n): forward uses w, backward uses conj(w)
template <int DIR, typename C> __device__ __forceinline__ C twmul(C a, C w)
{
    return DIR < 0 ? cmul(a, w) : cmulc(a, w);
}
// multiply by DIR * i  (forward: -i, backward: +i)
template <int DIR, typename C> __device__ __forceinline__ C mul_dir_i(C a)
{
    C r;
    if (DIR < 0) { r.x = a.y; r.y = -a.x; }
    else         { r.x = -a.y; r.y = a.x; }
    return r;
}

template <int R, int DIR, typename C> struct Butterfly;

template <int DIR, typename C> struct Butterfly<1, DIR, C> {
    __device__ __forceinline__ static void run(C *) {}
};

template <int DIR, typename C> struct Butterfly<2, DIR, C> {
    __device__ __forceinline__ static void run(C *u)
    {
        C a = u[0], b = u[1];
        u[0] = cadd(a, b);
        u[1] = csub(a, b);
    }
};

template <int DIR, typename C> struct Butterfly<4, DIR, C> {
    __device__ __forceinline__ static void run(C *u)
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
    __device__ __forceinline__ static void run(C *u)
    {
        using T = decltype(u[0].x);
        const T h = (T)0.70710678118654752440084436210484903928;
        C e[4] = { u[0], u[2], u[4], u[6] };
        C o[4] = { u[1], u[3], u[5], u[7] };
        Butterfly<4, DIR, C>::run(e);
        Butterfly<4, DIR, C>::run(o);
        // W8^1 = (1 + DIR*i)/sqrt2, W8^2 = DIR*i, W8^3 = (-1 + DIR*i)/sqrt2
        C w1, w3;
        if (DIR < 0) {
            w1.x = (o[1].x + o[1].y) * h;  w1.y = (o[1].y - o[1].x) * h;
            w3.x = (o[3].y - o[3].x) * h;  w3.y = -(o[3].x + o[3].y) * h;
        } else {
            w1.x = (o[1].x - o[1].y) * h;  w1.y = (o[1].x + o[1].y) * h;
            w3.x = -(o[3].x + o[3].y) * h; w3.y = (o[3].x - o[3].y) * h;
        }
        C w2 = mul_dir_i<DIR>(o[2]);
        u[0] = cadd(e[0], o[0]); u[4] = csub(e[0], o[0]);
        u[1] = cadd(e[1], w1);   u[5] = csub(e[1], w1);
        u[2] = cadd(e[2], w2);   u[6] = csub(e[2], w2);
        u[3] = cadd(e[3], w3);   u[7] = csub(e[3], w3);
    }
};

template <int DIR, typename C> struct Butterfly<16, DIR, C> {
    __device__ __forceinline__ static void run(C *u)
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
