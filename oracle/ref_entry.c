/* ref_entry.c -- TEST INFRASTRUCTURE ONLY.  Plain-pointer entry points into the reference's own
 * src/FftLinearSolver_3D.c (compiled unmodified against petsc_standin/, see its header): every function below only
 * wraps arrays into Vecs and calls the reference function of the same name.  Built into _ref/libreference_fftsolver.so
 * by the Makefile in this directory when /root/reference is present; loaded by oracle/ref_c.py for tests/ only. */
#include <string.h>

#include "FftLinearSolver_3D.h"

/* The library is built with -fvisibility=hidden -Wl,-Bsymbolic: the reference's function names (solve_3D, ...) and the
 * stand-in's PETSc names also exist in the product's glue library, which a test process may have loaded globally; only
 * the ref_* entry points are exported, and every internal call binds inside this library. */
#define REF_API __attribute__((visibility("default")))

/* not declared in the reference's header, defined in its .c */
PetscErrorCode Fft3DSolver(PetscInt n_x, PetscInt n_y, PetscInt n_z, PetscScalar lambda_x, PetscScalar lambda_y,
                           PetscScalar lambda_z, Vec X, Vec b, Mat FFT_MAT, Vec c_x_hat, Vec c_y_hat, Vec c_z_hat);

static PetscErrorCode fft_mat(int nx, int ny, int nz, Mat *A)
{
    PetscInt dims[3] = { nz, ny, nx };       /* as src/PCSHELLFft_3D.cxx:34 and the reference's drivers */
    return MatCreateFFT(PETSC_COMM_WORLD, 3, dims, MATFFTW, A);
}
static PetscErrorCode load(Vec v, const double *src, int n)
{
    PetscScalar *a;
    PetscCall(VecGetArray(v, &a));
    memcpy(a, src, sizeof(PetscScalar) * (size_t)n);
    return VecRestoreArray(v, &a);
}
static PetscErrorCode store(Vec v, double *dst, int n)
{
    PetscScalar *a;
    PetscCall(VecGetArray(v, &a));
    memcpy(dst, a, sizeof(PetscScalar) * (size_t)n);
    return VecRestoreArray(v, &a);
}

/* solver = 0: FftTransportSolver(lambdas = p[0..2]); 1: Fft3DTransportSolver(a = p[0..2], dt = p[3], delta = p[4..6]);
 * 2: Fft2DTransportSolver(a_x, a_y, dt, delta_x, delta_y = p[0], p[1], p[3], p[4], p[5]); 3: Fft1DTransportSolver(a_x, dt,
 * delta_x = p[0], p[3], p[4]); 4: PetscFft3DTransportSolver with the by-value context.  b, x: interleaved complex128.
 * (Fft3DSolver destroys the FFT Mat it is handed, :213, so it is not destroyed again here.) */
REF_API int ref_transport_solve(int solver, int nx, int ny, int nz, const double *p, const double *b, double *x)
{
    const int N = nx * ny * nz;
    Mat A;
    Vec B, X;
    PetscCall(fft_mat(nx, ny, nz, &A));
    PetscCall(MatCreateVecsFFTW(A, &B, &X, NULL));
    PetscCall(load(B, b, N));
    PetscErrorCode rc;
    if (solver == 0) rc = FftTransportSolver(nx, ny, nz, p[0], p[1], p[2], X, B, A);
    else if (solver == 1) rc = Fft3DTransportSolver(nx, ny, nz, p[0], p[1], p[2], p[3], p[4], p[5], p[6], X, B, A);
    else if (solver == 2) rc = Fft2DTransportSolver(nx, ny, p[0], p[1], p[3], p[4], p[5], X, B, A);
    else if (solver == 3) rc = Fft1DTransportSolver(nx, p[0], p[3], p[4], X, B, A);
    else {
        struct StructuredTransportContext c;
        c.n_x = nx; c.n_y = ny; c.n_z = nz;
        c.a_x = p[0]; c.a_y = p[1]; c.a_z = p[2]; c.dt = p[3];
        c.delta_x = p[4]; c.delta_y = p[5]; c.delta_z = p[6];
        c.FFT_MAT = A;
        rc = PetscFft3DTransportSolver(c, B, X);
    }
    if (!rc) rc = store(X, x, N);
    VecDestroy(&B);
    VecDestroy(&X);
    return rc;
}

/* Diag as the reference's set-up builds it: build_transport_col, three 1-D MatMults, build_diag_mat_vec_3D
 * (the prologue of FftTransportSolver :218-249 followed by :136-164). */
REF_API int ref_build_diag(int nx, int ny, int nz, double lx, double ly, double lz, double *diag)
{
    const int n[3] = { nx, ny, nz };
    Mat F[3], A;
    Vec c[3], ch[3], D;
    for (int a = 0; a < 3; ++a) {
        PetscInt d1[1] = { n[a] };
        PetscCall(MatCreateFFT(PETSC_COMM_WORLD, 1, d1, MATFFTW, &F[a]));
        PetscCall(MatCreateVecsFFTW(F[a], &c[a], &ch[a], NULL));
        PetscCall(build_transport_col(c[a], n[a]));
        PetscCall(MatMult(F[a], c[a], ch[a]));
    }
    PetscCall(fft_mat(nx, ny, nz, &A));
    PetscCall(MatCreateVecsFFTW(A, NULL, &D, NULL));
    PetscCall(build_diag_mat_vec_3D(D, ch[0], ch[1], ch[2], nx, ny, nz, lx, ly, lz));
    PetscCall(store(D, diag, nx * ny * nz));
    for (int a = 0; a < 3; ++a) { VecDestroy(&c[a]); VecDestroy(&ch[a]); MatDestroy(&F[a]); }
    VecDestroy(&D);
    MatDestroy(&A);
    return 0;
}

/* solve_3D (:166-190) with caller-supplied eigenvalues */
REF_API int ref_solve_3D(int nx, int ny, int nz, const double *diag, const double *b, double *x)
{
    const int N = nx * ny * nz;
    Mat A;
    Vec B, X, D, Bh;
    PetscCall(fft_mat(nx, ny, nz, &A));
    PetscCall(MatCreateVecsFFTW(A, &B, &X, &D));
    PetscCall(MatCreateVecsFFTW(A, NULL, &Bh, NULL));
    PetscCall(load(B, b, N));
    PetscCall(load(D, diag, N));
    PetscCall(solve_3D(A, X, D, B, Bh, N));
    PetscCall(store(X, x, N));
    VecDestroy(&B); VecDestroy(&X); VecDestroy(&D); VecDestroy(&Bh);
    MatDestroy(&A);
    return 0;
}

/* b == x aliasing as the reference's driver uses it (tests/TransportEquationFFT_SphericalExplosion_impl_mpi.cxx:111) */
REF_API int ref_transport_solve_in_place(int nx, int ny, int nz, double lx, double ly, double lz, double *bx)
{
    const int N = nx * ny * nz;
    Mat A;
    Vec U;
    PetscCall(fft_mat(nx, ny, nz, &A));
    PetscCall(MatCreateVecsFFTW(A, &U, NULL, NULL));
    PetscCall(load(U, bx, N));
    PetscCall(FftTransportSolver(nx, ny, nz, lx, ly, lz, U, U, A));
    PetscCall(store(U, bx, N));
    VecDestroy(&U);
    return 0;
}
