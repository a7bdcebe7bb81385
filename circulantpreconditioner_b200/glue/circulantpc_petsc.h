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

#ifdef CPC_WITH_PETSC
#include <petscksp.h>
#else
#include "petsc_shim.h"
#endif

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
