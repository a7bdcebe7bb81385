/* CPU oracle (plain C) for the circulant preconditioner apply.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into libcirculantpc.so.  Built by
 * oracle/Makefile into oracle/liboracle_c.so and loaded with ctypes by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg as a *checker*.
 *
 * It restates /root/reference/src/FftLinearSolver_3D.c function by function
 * with PETSc Vec replaced by interleaved (re,im) double arrays.  The 3-D DFT
 * that the reference delegates to PETSc MATFFTW -> FFTW 3.3.x (not vendored,
 * not installable here) is restated as a separable mixed-radix
 * decimation-in-time DFT (any n: radix = smallest prime factor, O(n*sum p));
 * it is the same mathematical transform (forward sign exp(-2 pi i ..),
 * unnormalised, dims = {nz, ny, nx}, x fastest).
 *
 * Pinned by tests/test_oracle.py against the fixtures generated from the
 * reference's own Python tests (tests/golden/make_golden.py) and against the
 * integer known-answer vectors of tests/FFTDirectSolver/testFftSolver_*.c.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cplx;

static cplx cmul(cplx a, cplx b) { cplx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static cplx cadd(cplx a, cplx b) { cplx r = { a.re + b.re, a.im + b.im }; return r; }

/* exp(sign * 2 pi i m / n), m reduced mod n, octant-exact via long double */
static cplx root(long m, long n, int sign)
{
    m %= n; if (m < 0) m += n;
    long double a = 2.0L * 3.141592653589793238462643383279502884L * (long double)m / (long double)n;
    cplx r = { (double)cosl(a), (double)(sign * sinl(a)) };
    return r;
}

static int smallest_factor(int n)
{
    for (int p = 2; (long)p * p <= n; ++p) if (n % p == 0) return p;
    return n;
}

/* out[k], k<n, of in[0], in[stride], ...; tw = table of n_total-th roots, tw_step = n_total / n */
static void dft_rec(const cplx *in, cplx *out, int n, int stride, const cplx *tw, int tw_step, cplx *scratch)
{
    if (n == 1) { out[0] = in[0]; return; }
    int p = smallest_factor(n), m = n / p;
    /* p sub-transforms of length m over the decimated inputs */
    for (int r = 0; r < p; ++r)
        dft_rec(in + (size_t)r * stride, scratch + (size_t)r * m, m, stride * p, tw, tw_step * p, out);
    /* combine: X[k + q m] = sum_r W_n^{r (k + q m)} Y_r[k] */
    for (int k = 0; k < m; ++k)
        for (int q = 0; q < p; ++q) {
            cplx acc = { 0.0, 0.0 };
            int kk = k + q * m;
            for (int r = 0; r < p; ++r) {
                long e = ((long)r * kk) % n;
                acc = cadd(acc, cmul(tw[(size_t)e * tw_step], scratch[(size_t)r * m + k]));
            }
            out[kk] = acc;
        }
    /* scratch of the children is `out` of the parent and vice versa: copy children results were
       consumed above, so nothing else to do */
}

/* In-place unnormalised DFT of `count` lines of length n, element stride `stride`,
 * line starts given by two nested loops (n0 x n1 with strides s0, s1). */
static void dft_axis(cplx *a, int n, long stride, long n0, long s0, long n1, long s1, int sign)
{
    if (n == 1) return;
    cplx *tw = (cplx *)malloc(sizeof(cplx) * (size_t)n);
    for (int m = 0; m < n; ++m) tw[m] = root(m, n, sign);
#pragma omp parallel
    {
        cplx *buf = (cplx *)malloc(sizeof(cplx) * (size_t)n * 3);
        cplx *in = buf, *out = buf + n, *scr = buf + 2 * (size_t)n;
#pragma omp for collapse(2) schedule(static)
        for (long i0 = 0; i0 < n0; ++i0)
            for (long i1 = 0; i1 < n1; ++i1) {
                cplx *base = a + i0 * s0 + i1 * s1;
                for (int k = 0; k < n; ++k) in[k] = base[(long)k * stride];
                dft_rec(in, out, n, 1, tw, 1, scr);
                for (int k = 0; k < n; ++k) base[(long)k * stride] = out[k];
            }
        free(buf);
    }
    free(tw);
}

/* MatMult / MatMultTranspose on MATFFTW (FftLinearSolver_3D.c:170,180): sign=-1 forward, +1 backward */
void oracle_dft3(double *data, int nx, int ny, int nz, int sign)
{
    cplx *a = (cplx *)data;
    long sxy = (long)nx * ny;
    dft_axis(a, nx, 1, nz, sxy, ny, nx, sign);       /* x lines */
    dft_axis(a, ny, nx, nz, sxy, nx, 1, sign);       /* y lines */
    dft_axis(a, nz, sxy, ny, nx, nx, 1, sign);       /* z lines */
}

/* FftLinearSolver_3D.c:80-90 */
void oracle_build_transport_col(double *c, int size)
{
    memset(c, 0, sizeof(double) * 2 * (size_t)size);
    if (size > 1) { c[0] = 1.0; c[2] = -1.0; }
}

/* FftLinearSolver_3D.c:92-112 : res[j*c_size + i] = lambda * c[i] */
static void kron_identity_left(const cplx *c, cplx *res, long c_size, long id_size, double lambda)
{
    for (long i = 0; i < c_size; ++i) {
        cplx cur = { c[i].re * lambda, c[i].im * lambda };
        for (long j = 0; j < id_size; ++j) res[j * c_size + i] = cur;
    }
}

/* FftLinearSolver_3D.c:114-134 : res[i*id_size + j] = lambda * c[i] */
static void kron_identity_right(const cplx *c, cplx *res, long c_size, long id_size, double lambda)
{
    for (long i = 0; i < c_size; ++i) {
        cplx cur = { c[i].re * lambda, c[i].im * lambda };
        for (long j = 0; j < id_size; ++j) res[i * id_size + j] = cur;
    }
}

/* FftLinearSolver_3D.c:136-164 */
void oracle_build_diag_mat_vec_3D(double *Diag, const double *cxh, const double *cyh, const double *czh,
                                  int nx, int ny, int nz, double lx, double ly, double lz)
{
    size_t N = (size_t)nx * ny * nz;
    cplx *kx = (cplx *)malloc(sizeof(cplx) * N), *ky = (cplx *)malloc(sizeof(cplx) * N);
    cplx *kyi = (cplx *)malloc(sizeof(cplx) * (size_t)ny * nz), *kz = (cplx *)malloc(sizeof(cplx) * N);
    kron_identity_left((const cplx *)cxh, kx, nx, (long)ny * nz, lx);
    kron_identity_left((const cplx *)cyh, kyi, ny, nz, ly);
    kron_identity_right(kyi, ky, (long)ny * nz, nx, 1.0);
    kron_identity_right((const cplx *)czh, kz, nz, (long)nx * ny, lz);
    cplx *D = (cplx *)Diag;
    for (size_t m = 0; m < N; ++m) {
        D[m].re = kx[m].re + ky[m].re + kz[m].re + 1.0;      /* three AXPY + VecShift(1) */
        D[m].im = kx[m].im + ky[m].im + kz[m].im;
    }
    free(kx); free(ky); free(kyi); free(kz);
}

/* set-up path of FftTransportSolver (FftLinearSolver_3D.c:218-249) */
void oracle_transport_diag(double *Diag, int nx, int ny, int nz, double lx, double ly, double lz)
{
    int n[3] = { nx, ny, nz };
    double *ch[3];
    for (int d = 0; d < 3; ++d) {
        ch[d] = (double *)malloc(sizeof(double) * 2 * (size_t)n[d]);
        oracle_build_transport_col(ch[d], n[d]);
        oracle_dft3(ch[d], n[d], 1, 1, -1);
    }
    oracle_build_diag_mat_vec_3D(Diag, ch[0], ch[1], ch[2], nx, ny, nz, lx, ly, lz);
    for (int d = 0; d < 3; ++d) free(ch[d]);
}

/* FftLinearSolver_3D.c:166-190, complex-scalar branch.  b may alias X. */
void oracle_solve_3D(double *X, const double *Diag, const double *b, int nx, int ny, int nz)
{
    size_t N = (size_t)nx * ny * nz;
    cplx *bh = (cplx *)malloc(sizeof(cplx) * N);
    memcpy(bh, b, sizeof(cplx) * N);
    oracle_dft3((double *)bh, nx, ny, nz, -1);                       /* :170 */
    const cplx *D = (const cplx *)Diag;
    for (size_t m = 0; m < N; ++m) {                                 /* :174 */
        double d2 = D[m].re * D[m].re + D[m].im * D[m].im;
        cplx q = { (bh[m].re * D[m].re + bh[m].im * D[m].im) / d2, (bh[m].im * D[m].re - bh[m].re * D[m].im) / d2 };
        bh[m] = q;
    }
    oracle_dft3((double *)bh, nx, ny, nz, +1);                       /* :180 */
    double s = 1.0 / (double)N;                                      /* :184 */
    cplx *x = (cplx *)X;
    for (size_t m = 0; m < N; ++m) { x[m].re = bh[m].re * s; x[m].im = bh[m].im * s; }
    free(bh);
}

/* FftLinearSolver_3D.c:266-281 then :218-264 then :192-216 */
void oracle_Fft3DTransportSolver(int nx, int ny, int nz, double ax, double ay, double az, double dt,
                                 double dx, double dy, double dz, double *X, const double *b)
{
    size_t N = (size_t)nx * ny * nz;
    double *Diag = (double *)malloc(sizeof(double) * 2 * N);
    oracle_transport_diag(Diag, nx, ny, nz, ax * dt / dx, ay * dt / dy, az * dt / dz);
    oracle_solve_3D(X, Diag, b, nx, ny, nz);
    free(Diag);
}

/* The same solve with the z factor evaluated WITHOUT FFTs: after the x and y transforms every z line is the cyclic
 * bidiagonal system (alpha + lz) x_k - lz x_{k-1} = b_k, alpha = 1 + lx c_x_hat[kx] + ly c_y_hat[ky] -- its spectrum is
 * exactly Diag[k] of FftLinearSolver_3D.c:146-157 -- i.e. the recurrence y_k = c y_{k-1} + b_k, c = lz / (alpha + lz),
 * closed by y_{-1} = y_{n-1}^{(0)} / (1 - c^n), and x = y / (alpha + lz).  This is the form the CUDA middle pass takes
 * for the transport symbol (circulantpreconditioner_b200/csrc/zsolve.cuh); restated here so that its equality with
 * oracle_solve_3D is pinned on the CPU.  Needs lz >= 0 and lx, ly >= 0 (|c| < 1).  Returns 0, or -1 on bad lambdas. */
int oracle_transport_solve_z_recurrence(double *X, const double *b, int nx, int ny, int nz, double lx, double ly,
                                        double lz)
{
    if (lx < 0 || ly < 0 || lz < 0) return -1;
    size_t N = (size_t)nx * ny * nz;
    long sxy = (long)nx * ny;
    cplx *a = (cplx *)malloc(sizeof(cplx) * N);
    memcpy(a, b, sizeof(cplx) * N);
    dft_axis(a, nx, 1, nz, sxy, ny, nx, -1);          /* Fx */
    dft_axis(a, ny, nx, nz, sxy, nx, 1, -1);          /* Fy */
#pragma omp parallel for schedule(static)
    for (long line = 0; line < sxy; ++line) {
        const int i = (int)(line % nx), j = (int)(line / nx);
        cplx wx = root(i, nx, -1), wy = root(j, ny, -1);
        /* alpha + lz, with c_hat[q] = 1 - exp(-2 pi i q / n) (0 on an axis of length 1, :80-90) */
        double are = 1.0 + (nx > 1 ? lx * (1.0 - wx.re) : 0.0) + (ny > 1 ? ly * (1.0 - wy.re) : 0.0) + lz;
        double aim = (nx > 1 ? -lx * wx.im : 0.0) + (ny > 1 ? -ly * wy.im : 0.0);
        if (nz == 1) are -= lz;                        /* c_z_hat = 0: no z coupling at all */
        double d2 = are * are + aim * aim;
        cplx r = { are / d2, -aim / d2 };
        cplx c = { (nz > 1 ? lz : 0.0) * r.re, (nz > 1 ? lz : 0.0) * r.im };
        cplx acc = { 0.0, 0.0 };
        for (int k = 0; k < nz; ++k) {                 /* zero carry-in */
            acc = cadd(cmul(c, acc), a[line + (long)k * sxy]);
            a[line + (long)k * sxy] = acc;
        }
        cplx cn = { 1.0, 0.0 };
        for (int k = 0; k < nz; ++k) cn = cmul(cn, c);
        double e2 = (1.0 - cn.re) * (1.0 - cn.re) + cn.im * cn.im;
        cplx inv = { (1.0 - cn.re) / e2, cn.im / e2 };  /* 1 / (1 - c^n) */
        cplx carry = cmul(acc, inv);
        cplx cp = c;
        for (int k = 0; k < nz; ++k) {
            cplx y = cadd(a[line + (long)k * sxy], cmul(cp, carry));
            a[line + (long)k * sxy] = cmul(y, r);
            cp = cmul(cp, c);
        }
    }
    dft_axis(a, ny, nx, nz, sxy, nx, 1, +1);          /* By */
    dft_axis(a, nx, 1, nz, sxy, ny, nx, +1);          /* Bx */
    double s = 1.0 / ((double)nx * ny);
    cplx *x = (cplx *)X;
    for (size_t m = 0; m < N; ++m) { x[m].re = a[m].re * s; x[m].im = a[m].im * s; }
    free(a);
    return 0;
}

/* The recurrence form as the CUDA line kernels and the multi-rank owner scheme evaluate it (csrc/zsolve.cuh:
 * zs_end_accum_kernel, zs_carry_owner_kernel, zs_dist_line_kernel): the z line is cut into `slabs` equal parts; the end
 * value of every part from a zero carry-in is summed only over the planes whose weight |c|^m can still reach
 * `weight_floor` (whole groups of 16 planes, as the kernel does); the line's owner closes the cycle over the parts --
 * Zin_0 by Horner, Zin_{r+1} = e_r + c^(nz/slabs) Zin_r -- and every part is solved from its carry-in.  slabs = 1 is the
 * single-GPU line form.  Same operator as oracle_solve_3D (FftLinearSolver_3D.c:166-190) to rounding.
 * Returns 0, -1 on bad lambdas, -2 if nz is not divisible by slabs (or slabs > 64). */
int oracle_transport_solve_z_line_form(double *X, const double *b, int nx, int ny, int nz, double lx, double ly, double lz,
                                       int slabs, double weight_floor)
{
    if (lx < 0 || ly < 0 || lz < 0) return -1;
    if (slabs < 1 || slabs > 64 || nz % slabs) return -2;
    size_t N = (size_t)nx * ny * nz;
    long sxy = (long)nx * ny;
    const int nzl = nz / slabs;
    cplx *a = (cplx *)malloc(sizeof(cplx) * N);
    memcpy(a, b, sizeof(cplx) * N);
    dft_axis(a, nx, 1, nz, sxy, ny, nx, -1);          /* Fx */
    dft_axis(a, ny, nx, nz, sxy, nx, 1, -1);          /* Fy */
#pragma omp parallel for schedule(static)
    for (long line = 0; line < sxy; ++line) {
        const int i = (int)(line % nx), j = (int)(line / nx);
        cplx wx = root(i, nx, -1), wy = root(j, ny, -1);
        double are = 1.0 + (nx > 1 ? lx * (1.0 - wx.re) : 0.0) + (ny > 1 ? ly * (1.0 - wy.re) : 0.0) + lz;
        double aim = (nx > 1 ? -lx * wx.im : 0.0) + (ny > 1 ? -ly * wy.im : 0.0);
        if (nz == 1) are -= lz;
        double d2 = are * are + aim * aim;
        cplx r = { are / d2, -aim / d2 };
        cplx c = { (nz > 1 ? lz : 0.0) * r.re, (nz > 1 ? lz : 0.0) * r.im };
        /* first plane of a part that still matters */
        double c2 = c.re * c.re + c.im * c.im;
        double m = c2 > 0.0 ? log(weight_floor * weight_floor) / log(c2) + 1.0 : 1.0;
        int start = (m < (double)nzl) ? nzl - (int)m : 0;
        start = nzl - ((nzl - start + 15) / 16) * 16;
        if (start < 0) start = 0;
        cplx e[64];
        for (int s_ = 0; s_ < slabs; ++s_) {           /* end value of part s_, zero carry-in, truncated sum */
            cplx acc = { 0.0, 0.0 };
            for (int k = start; k < nzl; ++k) acc = cadd(cmul(c, acc), a[line + (long)(s_ * nzl + k) * sxy]);
            e[s_] = acc;
        }
        cplx cL = { 1.0, 0.0 };
        for (int k = 0; k < nzl; ++k) cL = cmul(cL, c);
        cplx acc = { 0.0, 0.0 }, cLp = { 1.0, 0.0 };
        for (int s_ = 0; s_ < slabs; ++s_) {           /* owner: Horner over e_0 .. e_{P-1}, closed cyclically */
            acc = cadd(cmul(cL, acc), e[s_]);
            cLp = cmul(cLp, cL);
        }
        double e2 = (1.0 - cLp.re) * (1.0 - cLp.re) + cLp.im * cLp.im;
        cplx inv = { (1.0 - cLp.re) / e2, cLp.im / e2 };
        cplx Z = cmul(acc, inv);
        for (int s_ = 0; s_ < slabs; ++s_) {           /* second sweep of every part from its carry-in */
            cplx y = Z;
            for (int k = 0; k < nzl; ++k) {
                long idx = line + (long)(s_ * nzl + k) * sxy;
                y = cadd(cmul(c, y), a[idx]);
                a[idx] = cmul(y, r);
            }
            Z = cadd(e[s_], cmul(cL, Z));              /* Zin_{r+1} = e_r + cL Zin_r */
        }
    }
    dft_axis(a, ny, nx, nz, sxy, nx, 1, +1);          /* By */
    dft_axis(a, nx, 1, nz, sxy, ny, nx, +1);          /* Bx */
    double s = 1.0 / ((double)nx * ny);
    cplx *x = (cplx *)X;
    for (size_t mm = 0; mm < N; ++mm) { x[mm].re = a[mm].re * s; x[mm].im = a[mm].im * s; }
    free(a);
    return 0;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
