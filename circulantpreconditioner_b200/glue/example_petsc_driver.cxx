// example_petsc_driver.cxx -- what a maintainer of the reference changes in its two drivers to run on this library,
// written against the real PETSc API and compile-checked (syntax only) against petsc_opaque_stub.h by
// `make check-petsc-clean`.  Not part of any library; never linked or run here (PETSc is not installable).
//
//   direct solver   tests/TransportEquationFFT_SphericalExplosion_impl_mpi.cxx:96-112
//                   one line changes: MatCreateFFT(..., MATFFTW, &FFT_MAT)  ->  MatCreateFFT_CPC(..., &FFT_MAT)
//   preconditioner  tests/TransportEquation_SphericalExplosion_impl_mpi.cxx:119-126
//                   PCSetType(pc, PCNONE)  ->  PCSetType(pc, PCSHELL) + PCShellFFT3DAttach(pc, ctx)
//                   (the wiring the reference's ToDo.md item 1 asks for)
#if !defined(CPC_WITH_PETSC)
#error "this example is for a real-PETSc (or stub) build: -DCPC_WITH_PETSC"
#endif
#include "PCSHELLFft_3D.hxx"
#include "FftLinearSolver_3D.h"

// the time loop of TransportEquationFFT_impl_mpi (:96-112), Un advanced in place
PetscErrorCode transport_fft_time_loop(PetscInt nx, PetscInt ny, PetscInt nz, const double a[3], double dt, double delta_x,
                                       double delta_y, double delta_z, Vec Un, int ntmax)
{
    PetscFunctionBeginUser;
    Mat FFT_MAT;
    PetscInt ndim = 3;
    PetscInt dims[3] = { nz, ny, nx };
    PetscCall(MatCreateFFT_CPC(PETSC_COMM_WORLD, ndim, dims, &FFT_MAT));          // was: MatCreateFFT(..., MATFFTW, &FFT_MAT)
    StructuredTransportContext ctx = { nx, ny, nz, a[0], a[1], a[2], dt, delta_x, delta_y, delta_z, FFT_MAT };
    for (int it = 0; it < ntmax; ++it) PetscCall(PetscFft3DTransportSolver(ctx, Un, Un));    // unchanged (:111)
    PetscCall(MatDestroy(&FFT_MAT));       // the caller's Mat stays valid across steps (the reference destroys it, F7)
    PetscFunctionReturn(PETSC_SUCCESS);
}

// the KSP set-up of TransportEquation_impl_mpi (:119-126) with the circulant preconditioner plugged in
PetscErrorCode transport_ksp_solve(Mat A, Vec Un, PetscInt dim, double dt, PetscInt nbCells, const double a[3],
                                   const double lo[3], const double hi[3], Mat intersectionMatrix, double precision,
                                   PetscInt maxPetscIts)
{
    PetscFunctionBeginUser;
    KSP ksp;
    PC pc;
    PetscCall(KSPCreate(PETSC_COMM_WORLD, &ksp));
    PetscCall(KSPSetType(ksp, KSPGMRES));
    PetscCall(KSPSetTolerances(ksp, precision, precision, PETSC_DEFAULT, maxPetscIts));
    PetscCall(KSPGetPC(ksp, &pc));
    PetscCall(PCSetType(pc, PCSHELL));                                              // was: PCNONE
    FFTPrecTransportContext *ctx = nullptr;
    PetscCall(getFFTPrec3DContextCreate(dim, dt, nbCells, a[0], a[1], a[2], lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], &ctx));
    ctx->intersectionMatrix = intersectionMatrix;                                   // NULL on a Cartesian mesh
    PetscCall(PCShellFFT3DAttach(pc, ctx));     // SetContext + SetSetUp(setupFFTPrec3D) + SetApply(applyFFT3DPrecTransport) + SetDestroy
    PetscCall(KSPSetOperators(ksp, A, A));
    PetscCall(KSPSolve(ksp, Un, Un));                                               // unchanged (:136)
    PetscCall(KSPDestroy(&ksp));                                                    // -> destroyFFTPrec3D through the PC
    PetscCall(FFTPrec3DContextFree(&ctx));
    PetscFunctionReturn(PETSC_SUCCESS);
}
