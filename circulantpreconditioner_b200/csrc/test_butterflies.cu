// test_butterflies.cu -- host-side check of the in-register butterflies of fft_core.cuh against the DFT definition
// (reference convention: MatMult on MATFFTW = unnormalised forward DFT, src/FftLinearSolver_3D.c:170,180).
// Built and run by tests/test_butterflies.py; needs no GPU (the butterflies are __host__ __device__).
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "fft_core.cuh"

using namespace cpc;

template <int R, int DIR, typename T> static double check_plain()
{
    using C = cplx_t<T>;
    C u[R], x[R];
    for (int r = 0; r < R; ++r) x[r] = u[r] = mk<T>((T)(rand() / (double)RAND_MAX - 0.5), (T)(rand() / (double)RAND_MAX - 0.5));
    Butterfly<R, DIR, C>::run(u);
    double worst = 0;
    for (int q = 0; q < R; ++q) {
        double re = 0, im = 0;
        for (int r = 0; r < R; ++r) {
            const double a = DIR * 2.0 * M_PI * ((r * q) % R) / R;
            re += x[r].x * cos(a) - x[r].y * sin(a);
            im += x[r].x * sin(a) + x[r].y * cos(a);
        }
        worst = fmax(worst, hypot(u[q].x - re, u[q].y - im));
    }
    return worst;
}

template <int DIR, typename T> static double check_twiddled8()
{
    using C = cplx_t<T>;
    C u[8], x[8], w[7];
    for (int r = 0; r < 8; ++r) x[r] = u[r] = mk<T>((T)(rand() / (double)RAND_MAX - 0.5), (T)(rand() / (double)RAND_MAX - 0.5));
    for (int r = 1; r < 8; ++r) { const double a = -2.0 * M_PI * r * 5 / 64; w[r - 1] = mk<T>((T)cos(a), (T)sin(a)); }
    butterfly8_twiddled<DIR>(u, w);
    C v[8];
    v[0] = x[0];
    for (int r = 1; r < 8; ++r) v[r] = twmul<DIR>(x[r], w[r - 1]);
    Butterfly<8, DIR, C>::run(v);
    double worst = 0;
    for (int q = 0; q < 8; ++q) worst = fmax(worst, hypot(u[q].x - v[q].x, u[q].y - v[q].y));
    return worst;
}

#define RUN(R)                                                                                               \
    do {                                                                                                     \
        const double e1 = check_plain<R, -1, double>(), e2 = check_plain<R, +1, double>();                    \
        const double f1 = check_plain<R, -1, float>(), f2 = check_plain<R, +1, float>();                      \
        printf("radix %2d: fp64 %.1e %.1e  fp32 %.1e %.1e\n", R, e1, e2, f1, f2);                           \
        if (e1 > 2e-14 || e2 > 2e-14 || f1 > 5e-6 || f2 > 5e-6) fail = 1;                                     \
    } while (0)

int main()
{
    int fail = 0;
    srand(7);
    RUN(2); RUN(3); RUN(4); RUN(5); RUN(6); RUN(7); RUN(8); RUN(10); RUN(12); RUN(16); RUN(20);
    const double t1 = check_twiddled8<-1, double>(), t2 = check_twiddled8<+1, double>();
    printf("twiddled radix 8: %.1e %.1e\n", t1, t2);
    if (t1 > 1e-14 || t2 > 1e-14) fail = 1;
    printf(fail ? "FAILED\n" : "ALL PASSED\n");
    return fail;
}
