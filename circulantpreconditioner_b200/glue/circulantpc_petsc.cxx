// circulantpc_petsc.cxx -- reference-named solver entry points over the libcirculantpc C ABI.
// Each function cites the reference lines whose behaviour it reproduces; none of the arithmetic is done here.
#include "circulantpc_petsc.h"

#include <cmath>

namespace {

// The plan that stands behind an FFT Mat (MatCreateFFT in petsc_shim.cxx; with real PETSc: a plan composed onto
// the Mat with PetscObjectCompose).
PetscErrorCode plan_of(Mat FFT_MAT, PetscInt n_x, PetscInt n_y, PetscInt n_z, cpc_plan *plan)
{
    PetscCheck(FFT_MAT && FFT_MAT->kind == SHIM_MAT_FFT && FFT_MAT->plan, PETSC_COMM_WORLD, PETSC_ERR_ARG_WRONG,
               "FFT_MAT was not created by MatCreateFFT");
    cpc_plan_info info;
    PetscCallCPC(cpc_get_info(FFT_MAT->plan, &info));
    PetscCheck(n_x < 0 || (info.nx == n_x && info.ny == n_y && info.nz == n_z), PETSC_COMM_WORLD, PETSC_ERR_ARG_WRONG,
               "FFT_MAT is %d x %d x %d but the call says %d x %d x %d", info.nx, info.ny, info.nz, n_x, n_y, n_z);
    *plan = FFT_MAT->plan;
    return PETSC_SUCCESS;
}

}  // namespace

extern "C" {

// reference FftLinearSolver_3D.c:80-90: zero vector, then c[0] = 1, c[1] = -1 when size > 1
PetscErrorCode build_transport_col(Vec c, PetscInt size)
{
    PetscFunctionBeginUser;
    PetscCall(VecSet(c, 0.0));
    if (size > 1) {
        PetscCall(VecSetValue(c, 0, 1.0, INSERT_VALUES));
        PetscCall(VecSetValue(c, 1, -1.0, INSERT_VALUES));
    }
    PetscFunctionReturn(PETSC_SUCCESS);
}

// reference FftLinearSolver_3D.c:92-112: res = (1_{id_size} (x) lambda c), i.e. res[j*c_size + i] = lambda*c[i]
// ("tile").  Kept for callers that assemble Diag by hand; build_diag_mat_vec_3D below does not need it.
// One pass over the arrays instead of the reference's c_size*id_size VecSetValue calls.
PetscErrorCode vec_kronecker_product_identity_left(Vec c, Vec res, PetscInt c_size, PetscInt id_size, PetscScalar lambda)
{
    PetscFunctionBeginUser;
    PetscInt sc, sr;
    PetscCall(VecGetSize(c, &sc));
    PetscCall(VecGetSize(res, &sr));
    PetscCheck(sc == c_size && sr == c_size * id_size, PETSC_COMM_WORLD, PETSC_ERR_ARG_WRONG,
               "vec_kronecker_product_identity_left: c has %d entries, res %d, expected %d and %d", sc, sr, c_size,
               c_size * id_size);
    const PetscScalar *cc;
    PetscScalar *rr;
    PetscCall(VecGetArrayRead(c, &cc));
    PetscCall(VecGetArray(res, &rr));
    for (PetscInt j = 0; j < id_size; ++j)
        for (PetscInt i = 0; i < c_size; ++i) rr[(size_t)j * c_size + i] = lambda * cc[i];
    PetscCall(VecRestoreArray(res, &rr));
    PetscCall(VecRestoreArrayRead(c, &cc));
    PetscFunctionReturn(PETSC_SUCCESS);
}

// reference FftLinearSolver_3D.c:114-134: res = (lambda c (x) 1_{id_size}), i.e. res[i*id_size + j] = lambda*c[i]
// ("repeat").
PetscErrorCode vec_kronecker_product_identity_right(Vec c, Vec res, PetscInt c_size, PetscInt id_size, PetscScalar lambda)
{
    PetscFunctionBeginUser;
    PetscInt sc, sr;
    PetscCall(VecGetSize(c, &sc));
    PetscCall(VecGetSize(res, &sr));
    PetscCheck(sc == c_size && sr == c_size * id_size, PETSC_COMM_WORLD, PETSC_ERR_ARG_WRONG,
               "vec_kronecker_product_identity_right: c has %d entries, res %d, expected %d and %d", sc, sr, c_size,
               c_size * id_size);
    const PetscScalar *cc;
    PetscScalar *rr;
    PetscCall(VecGetArrayRead(c, &cc));
    PetscCall(VecGetArray(res, &rr));
    for (PetscInt i = 0; i < c_size; ++i) {
        const PetscScalar v = lambda * cc[i];
        for (PetscInt j = 0; j < id_size; ++j) rr[(size_t)i * id_size + j] = v;
    }
    PetscCall(VecRestoreArray(res, &rr));
    PetscCall(VecRestoreArrayRead(c, &cc));
    PetscFunctionReturn(PETSC_SUCCESS);
}

// reference FftLinearSolver_3D.c:136-164: Diag[k,j,i] = 1 + lx cx[i] + ly cy[j] + lz cz[k].
// The three 1-D tables go to the GPU (cpc_set_symbol_separable) and the N eigenvalues are produced there
// (cpc_get_diag) -- no per-element VecSetValue loops (reference :92-134 does 3N of them).
PetscErrorCode build_diag_mat_vec_3D(Vec Diag, Vec c_x_hat, Vec c_y_hat, Vec c_z_hat, PetscInt n_x, PetscInt n_y,
                                     PetscInt n_z, PetscScalar lambda_x, PetscScalar lambda_y, PetscScalar lambda_z)
{
    PetscFunctionBeginUser;
    PetscInt sx, sy, sz, sd;
    PetscCall(VecGetSize(c_x_hat, &sx));
    PetscCall(VecGetSize(c_y_hat, &sy));
    PetscCall(VecGetSize(c_z_hat, &sz));
    PetscCall(VecGetSize(Diag, &sd));
    PetscCheck(sx == n_x && sy == n_y && sz == n_z && sd == n_x * n_y * n_z, PETSC_COMM_WORLD, PETSC_ERR_ARG_WRONG,
               "build_diag_mat_vec_3D: vector sizes do not match %d x %d x %d", n_x, n_y, n_z);
    PetscCheck(std::imag(lambda_x) == 0 && std::imag(lambda_y) == 0 && std::imag(lambda_z) == 0, PETSC_COMM_WORLD,
               PETSC_ERR_ARG_WRONG, "build_diag_mat_vec_3D: lambdas must be real");
    // a scratch plan of the right shape turns the tables into the N eigenvalues on the GPU
    cpc_plan_desc d = { n_x, n_y, n_z, 1, CPC_C128, 1, 0, nullptr, nullptr, -1 };
    cpc_plan plan = nullptr;
    PetscCallCPC(cpc_plan_create(&plan, &d));
    const PetscScalar *cx, *cy, *cz;
    PetscCall(VecGetArrayRead(c_x_hat, &cx));
    PetscCall(VecGetArrayRead(c_y_hat, &cy));
    PetscCall(VecGetArrayRead(c_z_hat, &cz));
    int st = cpc_set_symbol_separable(plan, (const double *)cx, (const double *)cy, (const double *)cz,
                                      std::real(lambda_x), std::real(lambda_y), std::real(lambda_z));
    PetscScalar *dd;
    PetscCall(VecGetArray(Diag, &dd));
    if (!st) st = cpc_get_diag(plan, dd, CPC_MEM_HOST);
    PetscCall(VecRestoreArray(Diag, &dd));
    cpc_destroy(plan);
    if (st) return ShimError(PETSC_ERR_LIB, "libcirculantpc: %s", cpc_last_error());
    PetscFunctionReturn(PETSC_SUCCESS);
}

// reference FftLinearSolver_3D.c:166-190 (complex-scalar branch):
//   MatMult(FFT_MAT, b, b_hat); b_hat ./= Diag; MatMultTranspose(FFT_MAT, b_hat, X); X *= 1/size
// as ONE cpc_apply (5 HBM passes).  b may alias X.  b_hat is accepted for signature compatibility and not touched.
PetscErrorCode solve_3D(Mat FFT_MAT, Vec X, Vec Diag, Vec b, Vec b_hat, PetscInt size)
{
    PetscFunctionBeginUser;
    (void)b_hat;
    cpc_plan plan = nullptr;
    PetscCall(plan_of(FFT_MAT, -1, -1, -1, &plan));
    PetscInt nb, nx_, nd;
    PetscCall(VecGetSize(b, &nb));
    PetscCall(VecGetSize(X, &nx_));
    PetscCall(VecGetSize(Diag, &nd));
    PetscCheck(nb == size && nx_ == size && nd == size, PETSC_COMM_WORLD, PETSC_ERR_ARG_WRONG,
               "solve_3D: vectors must have %d entries (b %d, X %d, Diag %d)", size, nb, nx_, nd);
    const PetscScalar *dd;
    PetscCall(VecGetArrayRead(Diag, &dd));
    if (FFT_MAT->diag_seen != (const void *)dd || FFT_MAT->diag_state != Diag->state) {
        PetscCallCPC(cpc_set_symbol_diag(plan, dd, CPC_MEM_HOST));      // once per Diag, then it lives in HBM
        FFT_MAT->diag_seen = dd;
        FFT_MAT->diag_state = Diag->state;
    }
    PetscCall(VecRestoreArrayRead(Diag, &dd));
    const PetscScalar *bb;
    PetscScalar *xx;
    PetscCall(VecGetArrayRead(b, &bb));
    if (b == X) {
        xx = const_cast<PetscScalar *>(bb);
        ++X->state;
    } else {
        PetscCall(VecGetArray(X, &xx));
    }
    PetscCallCPC(cpc_apply(plan, bb, xx, CPC_MEM_HOST));
    if (b != X) PetscCall(VecRestoreArray(X, &xx));
    PetscCall(VecRestoreArrayRead(b, &bb));
    PetscFunctionReturn(PETSC_SUCCESS);
}

// reference FftLinearSolver_3D.c:192-216: allocate Diag and b_hat, build_diag_mat_vec_3D, solve_3D, free.
// Not destroying FFT_MAT is deliberate (the reference's MatDestroy at :213 leaves the caller with a dangling Mat).
PetscErrorCode Fft3DSolver(PetscInt n_x, PetscInt n_y, PetscInt n_z, PetscScalar lambda_x, PetscScalar lambda_y,
                           PetscScalar lambda_z, Vec X, Vec b, Mat FFT_MAT, Vec c_x_hat, Vec c_y_hat, Vec c_z_hat)
{
    PetscFunctionBeginUser;
    cpc_plan plan;
    PetscCall(plan_of(FFT_MAT, n_x, n_y, n_z, &plan));
    PetscInt size;
    PetscCall(VecGetSize(X, &size));
    PetscCheck(size == n_x * n_y * n_z, PETSC_COMM_WORLD, PETSC_ERR_ARG_WRONG, "Fft3DSolver: X has %d entries", size);
    // the separable tables go straight to the plan: no N-element Diag round trip
    const PetscScalar *cx, *cy, *cz;
    PetscCall(VecGetArrayRead(c_x_hat, &cx));
    PetscCall(VecGetArrayRead(c_y_hat, &cy));
    PetscCall(VecGetArrayRead(c_z_hat, &cz));
    PetscCallCPC(cpc_set_symbol_separable(plan, (const double *)cx, (const double *)cy, (const double *)cz,
                                          std::real(lambda_x), std::real(lambda_y), std::real(lambda_z)));
    FFT_MAT->diag_seen = nullptr;
    const PetscScalar *bb;
    PetscCall(VecGetArrayRead(b, &bb));
    PetscScalar *xx = (b == X) ? const_cast<PetscScalar *>(bb) : X->array;
    ++X->state;
    PetscCallCPC(cpc_apply(plan, bb, xx, CPC_MEM_HOST));
    PetscFunctionReturn(PETSC_SUCCESS);
}

// reference FftLinearSolver_3D.c:218-264: three columns [1,-1,0..], their 1-D DFTs, then Fft3DSolver.
// The column DFT has the closed form 1 - exp(-2 pi i q / n), which cpc_set_symbol_transport tabulates exactly.
PetscErrorCode FftTransportSolver(PetscInt n_x, PetscInt n_y, PetscInt n_z, PetscScalar lambda_x, PetscScalar lambda_y,
                                  PetscScalar lambda_z, Vec X, Vec b, Mat FFT_MAT)
{
    PetscFunctionBeginUser;
    cpc_plan plan;
    PetscCall(plan_of(FFT_MAT, n_x, n_y, n_z, &plan));
    PetscInt nb, nxx;
    PetscCall(VecGetSize(b, &nb));
    PetscCall(VecGetSize(X, &nxx));
    PetscCheck(nb == n_x * n_y * n_z && nxx == nb, PETSC_COMM_WORLD, PETSC_ERR_ARG_WRONG,
               "FftTransportSolver: vectors must have %d entries", n_x * n_y * n_z);
    PetscCheck(std::imag(lambda_x) == 0 && std::imag(lambda_y) == 0 && std::imag(lambda_z) == 0, PETSC_COMM_WORLD,
               PETSC_ERR_ARG_WRONG, "FftTransportSolver: lambdas must be real");
    PetscCallCPC(cpc_set_symbol_transport(plan, std::real(lambda_x), std::real(lambda_y), std::real(lambda_z)));
    FFT_MAT->diag_seen = nullptr;
    const PetscScalar *bb;
    PetscCall(VecGetArrayRead(b, &bb));
    PetscScalar *xx = (b == X) ? const_cast<PetscScalar *>(bb) : X->array;
    ++X->state;
    PetscCallCPC(cpc_apply(plan, bb, xx, CPC_MEM_HOST));
    PetscFunctionReturn(PETSC_SUCCESS);
}

// reference FftLinearSolver_3D.c:266-281: lambda_d = a_d dt / delta_d
PetscErrorCode Fft3DTransportSolver(PetscInt n_x, PetscInt n_y, PetscInt n_z, PetscScalar a_x, PetscScalar a_y,
                                    PetscScalar a_z, PetscScalar dt, PetscScalar delta_x, PetscScalar delta_y,
                                    PetscScalar delta_z, Vec X, Vec b, Mat FFT_MAT)
{
    PetscFunctionBeginUser;
    PetscCall(FftTransportSolver(n_x, n_y, n_z, a_x * dt / delta_x, a_y * dt / delta_y, a_z * dt / delta_z, X, b, FFT_MAT));
    PetscFunctionReturn(PETSC_SUCCESS);
}

// reference FftLinearSolver_3D.c:283-293: n_z = 1, a_z = 0, delta_z = 1
PetscErrorCode Fft2DTransportSolver(PetscInt n_x, PetscInt n_y, PetscScalar a_x, PetscScalar a_y, PetscScalar dt,
                                    PetscScalar delta_x, PetscScalar delta_y, Vec X, Vec b, Mat FFT_MAT)
{
    PetscFunctionBeginUser;
    PetscCall(Fft3DTransportSolver(n_x, n_y, 1, a_x, a_y, 0, dt, delta_x, delta_y, 1, X, b, FFT_MAT));
    PetscFunctionReturn(PETSC_SUCCESS);
}

// reference FftLinearSolver_3D.c:295-301
PetscErrorCode Fft1DTransportSolver(PetscInt n_x, PetscScalar a_x, PetscScalar dt, PetscScalar delta_x, Vec X, Vec b,
                                    Mat FFT_MAT)
{
    PetscFunctionBeginUser;
    PetscCall(Fft3DTransportSolver(n_x, 1, 1, a_x, 0, 0, dt, delta_x, 1, 1, X, b, FFT_MAT));
    PetscFunctionReturn(PETSC_SUCCESS);
}

// reference FftLinearSolver_3D.c:303-312: MatShell-style entry, context by value, note the (b, x) order
PetscErrorCode PetscFft3DTransportSolver(struct StructuredTransportContext c, Vec b, Vec x)
{
    PetscFunctionBeginUser;
    PetscCall(Fft3DTransportSolver(c.n_x, c.n_y, c.n_z, c.a_x, c.a_y, c.a_z, c.dt, c.delta_x, c.delta_y, c.delta_z, x, b,
                                   c.FFT_MAT));
    PetscFunctionReturn(PETSC_SUCCESS);
}

}  // extern "C"
