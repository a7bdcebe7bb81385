/* circulantpc_petsc.h -- the reference's operator / plugin interface for the preconditioner-apply path, served by
 * libcirculantpc (B200).  Same names, argument meaning and error behaviour (PetscErrorCode, 0 = success) as
 *   /root/reference/src/FftLinearSolver_3D.h:7-43   (C linkage: direct solver + solve_3D + eigenvalue set-up)
 *   /root/reference/src/PCSHELLFft_3D.hxx:8-45      (PCShell context and callbacks)
 * so a driver written against the reference links against libfftpreconditioner_b200.so unchanged
 * (FftLinearSolver_3D.h and PCSHELLFft_3D.hxx in this directory simply forward here).
 *
 * Deliberate differences, all fixes of defects listed in SURVEY.md section 0 / Appendix C:
 *   - the direct-solver wrappers do not MatDestroy the caller's FFT_MAT and do not rebuild the eigenvalues when
 *     called again with the same lambdas (reference FftLinearSolver_3D.c:213, 192-264): repeated time steps work;
 *   - PCShellGetContext is called with &ctx; getFFTPrec3DContext returns the context it fills (through
 *     getFFTPrec3DContextCreate) and uses lambda = a dt n / (Xmax - Xmin) (reference PCSHELLFft_3D.cxx:15,117,146-148);
 *   - PCShellFFT3DAttach registers the three callbacks, which the reference never does;
 *   - the projection Mat is optional (NULL = the mesh is the Cartesian grid) and, when present, the result is
 *     projected back with its transpose (reference PCSHELLFft_3D.cxx:17-21 leaves x on the Cartesian grid).
 */
#ifndef CIRCULANTPC_PETSC_H
#define CIRCULANTPC_PETSC_H

#if defined(CPC_WITH_PETSC) && defined(CPC_PETSC_STUB)
#include "petsc_opaque_stub.h"      /* compile check only: opaque PETSc types + the public prototypes this glue calls */
#elif defined(CPC_WITH_PETSC)
#include <petscksp.h>
#else
#include "petsc_shim.h"
#endif
#include "../../include/circulantpc.h"

/* a libcirculantpc status becomes a PETSc error carrying cpc_last_error() */
#define PetscCallCPC(call)                                                                                         \
    do {                                                                                                           \
        const int _st = (call);                                                                                    \
        PetscCheck(_st == 0, PETSC_COMM_SELF, _st == CPC_ERR_ARG ? PETSC_ERR_ARG_WRONG : PETSC_ERR_LIB,            \
                   "libcirculantpc: %s", cpc_last_error());                                                        \
    } while (0)

#ifndef CPC_WITH_SOLVERLAB
/* SOLVERLAB's Mesh is only an unused by-value argument of getFFTPrec3DContext (reference PCSHELLFft_3D.hxx:40). */
struct Mesh { int unused; };
#endif

extern "C" {

/* reference src/FftLinearSolver_3D.h:7-19 */
struct StructuredTransportContext {
    PetscInt n_x, n_y, n_z;
    PetscScalar a_x, a_y, a_z;
    PetscScalar dt;
    PetscScalar delta_x, delta_y, delta_z;
    Mat FFT_MAT;
};

/* ---- eigenvalue set-up (reference FftLinearSolver_3D.c:80-164) ---- */
PetscErrorCode build_transport_col(Vec c, PetscInt size);
PetscErrorCode vec_kronecker_product_identity_left(Vec c, Vec res, PetscInt c_size, PetscInt id_size, PetscScalar lambda);
PetscErrorCode vec_kronecker_product_identity_right(Vec c, Vec res, PetscInt c_size, PetscInt id_size, PetscScalar lambda);
PetscErrorCode build_diag_mat_vec_3D(Vec Diag, Vec c_x_hat, Vec c_y_hat, Vec c_z_hat, PetscInt n_x, PetscInt n_y,
                                     PetscInt n_z, PetscScalar lambda_x, PetscScalar lambda_y, PetscScalar lambda_z);

/* ---- the hot path (reference FftLinearSolver_3D.c:166-190) ---- */
PetscErrorCode solve_3D(Mat FFT_MAT, Vec X, Vec Diag, Vec b, Vec b_hat, PetscInt size);

/* ---- direct-solver wrappers (reference FftLinearSolver_3D.c:192-312) ---- */
PetscErrorCode Fft3DSolver(PetscInt n_x, PetscInt n_y, PetscInt n_z, PetscScalar lambda_x, PetscScalar lambda_y,
                           PetscScalar lambda_z, Vec X, Vec b, Mat FFT_MAT, Vec c_x_hat, Vec c_y_hat, Vec c_z_hat);
PetscErrorCode FftTransportSolver(PetscInt n_x, PetscInt n_y, PetscInt n_z, PetscScalar lambda_x, PetscScalar lambda_y,
                                  PetscScalar lambda_z, Vec X, Vec b, Mat FFT_MAT);
PetscErrorCode Fft3DTransportSolver(PetscInt n_x, PetscInt n_y, PetscInt n_z, PetscScalar a_x, PetscScalar a_y,
                                    PetscScalar a_z, PetscScalar dt, PetscScalar delta_x, PetscScalar delta_y,
                                    PetscScalar delta_z, Vec X, Vec b, Mat FFT_MAT);
PetscErrorCode Fft2DTransportSolver(PetscInt n_x, PetscInt n_y, PetscScalar a_x, PetscScalar a_y, PetscScalar dt,
                                    PetscScalar delta_x, PetscScalar delta_y, Vec X, Vec b, Mat FFT_MAT);
PetscErrorCode Fft1DTransportSolver(PetscInt n_x, PetscScalar a_x, PetscScalar dt, PetscScalar delta_x, Vec X, Vec b,
                                    Mat FFT_MAT);
PetscErrorCode PetscFft3DTransportSolver(struct StructuredTransportContext customCtx, Vec b, Vec x);

}  /* extern "C" */

/* reference src/PCSHELLFft_3D.hxx:8-21 */
struct FFTPrecTransportContext {
    PetscInt spaceDim;
    PetscInt n_x, n_y, n_z;
    PetscScalar lambda_x, lambda_y, lambda_z;
    Mat FFT_MAT;
    Mat intersectionMatrix;
    Vec Diag;
    Vec b_hat;
    Vec b_cartesien;
};

/* PCShell callbacks, reference src/PCSHELLFft_3D.hxx:23-25 / PCSHELLFft_3D.cxx:10-99 (C++ linkage there too) */
PetscErrorCode applyFFT3DPrecTransport(PC pc, Vec b, Vec x);
PetscErrorCode setupFFTPrec3D(PC pc);
PetscErrorCode destroyFFTPrec3D(PC pc);
/* reference src/PCSHELLFft_3D.hxx:27-41; fills the context kept by the library (see getFFTPrec3DContextCreate) */
PetscErrorCode getFFTPrec3DContext(PetscInt ndim, PetscScalar dt, PetscInt nbCells, PetscScalar a_x, PetscScalar a_y,
                                   PetscScalar a_z, PetscScalar Xmin, PetscScalar Ymin, PetscScalar Zmin,
                                   PetscScalar Xmax, PetscScalar Ymax, PetscScalar Zmax, Mesh srcMesh);

extern "C" {
/* additions (not in the reference): the plan that replaces the FFTW plan rides on the FFT Mat.
 *   CPCMatAttachPlan  creates the libcirculantpc plan for an n_x x n_y x n_z grid (z-slabs over comm's ranks) and
 *                     composes it onto A (PetscObjectCompose of a PetscContainer; destroyed with the Mat).  The shim's
 *                     MatCreateFFT calls it; with a real PETSc use MatCreateFFT_CPC below, or call it on the Mat that
 *                     MatCreateFFT returned.
 *   CPCMatGetPlan     the composed plan (error if none).                                                          */
PetscErrorCode CPCMatAttachPlan(Mat A, MPI_Comm comm, PetscInt n_x, PetscInt n_y, PetscInt n_z);
PetscErrorCode CPCMatGetPlan(Mat A, cpc_plan *plan);
#ifdef CPC_WITH_PETSC
/* MatCreateFFT(comm, ndim, dims, MATFFTW, A) without FFTW: a MATSHELL whose MatMult / MatMultTranspose are cpc_forward /
 * cpc_inverse, with the plan composed onto it (reference call sites: src/PCSHELLFft_3D.cxx:34-35,
 * tests/TransportEquationFFT_SphericalExplosion_impl_mpi.cxx:97-100; dims slowest first). */
PetscErrorCode MatCreateFFT_CPC(MPI_Comm comm, PetscInt ndim, const PetscInt dims[], Mat *A);
#endif
/* x = P^T solve_3D(P b) on the GPU in one call (used by applyFFT3DPrecTransport when ctx->intersectionMatrix is set) */
PetscErrorCode CPCApplyProjected(Mat FFT_MAT, Mat P, Vec Diag, Vec b, Vec x, PetscInt N);
/* additions (not in the reference): the missing wiring */
PetscErrorCode getFFTPrec3DContextCreate(PetscInt ndim, PetscScalar dt, PetscInt nbCells, PetscScalar a_x,
                                         PetscScalar a_y, PetscScalar a_z, PetscScalar Xmin, PetscScalar Ymin,
                                         PetscScalar Zmin, PetscScalar Xmax, PetscScalar Ymax, PetscScalar Zmax,
                                         struct FFTPrecTransportContext **ctx);
struct FFTPrecTransportContext *getFFTPrec3DLastContext(void);
PetscErrorCode FFTPrec3DContextFree(struct FFTPrecTransportContext **ctx);
PetscErrorCode PCShellFFT3DAttach(PC pc, struct FFTPrecTransportContext *ctx);
/* C-linkage aliases of the three callbacks, for dlsym / ctypes users */
PetscErrorCode cpc_glue_applyFFT3DPrecTransport(PC pc, Vec b, Vec x);
PetscErrorCode cpc_glue_setupFFTPrec3D(PC pc);
PetscErrorCode cpc_glue_destroyFFTPrec3D(PC pc);
}

#endif
