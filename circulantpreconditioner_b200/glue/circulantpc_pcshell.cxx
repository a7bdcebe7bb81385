// circulantpc_pcshell.cxx -- the PCShell adapter of the reference (src/PCSHELLFft_3D.cxx) over libcirculantpc.
// Public PETSc API only (see circulantpc_petsc.cxx).
#include <cmath>
#include <cstdlib>

#include "circulantpc_petsc.h"

// MatCreateFFT(comm, ndim, dims, MATFFTW, &A) of the reference (PCSHELLFft_3D.cxx:34-35): with a real PETSc the FFT Mat
// is this library's MATSHELL (no FFTW involved); the shim's MatCreateFFT does the same thing under the reference's name.
static PetscErrorCode create_fft_mat(MPI_Comm comm, PetscInt ndim, const PetscInt dims[], Mat *A)
{
#ifdef CPC_WITH_PETSC
    return MatCreateFFT_CPC(comm, ndim, dims, A);
#else
    return MatCreateFFT(comm, ndim, dims, MATFFTW, A);
#endif
}

// reference PCSHELLFft_3D.cxx:10-24: project b onto the Cartesian grid (when a projection exists), then solve_3D.
PetscErrorCode applyFFT3DPrecTransport(PC pc, Vec b, Vec x)
{
    PetscFunctionBeginUser;
    FFTPrecTransportContext *ctx = nullptr;
    PetscCall(PCShellGetContext(pc, &ctx));
    PetscCheck(ctx && ctx->FFT_MAT, PETSC_COMM_SELF, PETSC_ERR_ORDER, "applyFFT3DPrecTransport: PC not set up");
    const PetscInt N = ctx->n_x * ctx->n_y * ctx->n_z;
    if (ctx->intersectionMatrix) {
        // unstructured -> Cartesian, solve, Cartesian -> unstructured (the transpose; the reference stops half way):
        // both SpMVs and the five passes run on the GPU in one cpc_apply_projected call
        PetscCall(CPCApplyProjected(ctx->FFT_MAT, ctx->intersectionMatrix, ctx->Diag, b, x, N));
    } else {
        PetscCall(solve_3D(ctx->FFT_MAT, x, ctx->Diag, b, ctx->b_hat, N));
    }
    PetscFunctionReturn(PETSC_SUCCESS);
}

// reference PCSHELLFft_3D.cxx:26-84: FFT Mat + work Vecs + eigenvalues, once.
PetscErrorCode setupFFTPrec3D(PC pc)
{
    PetscFunctionBeginUser;
    FFTPrecTransportContext *ctx = nullptr;
    PetscCall(PCShellGetContext(pc, &ctx));
    PetscCheck(ctx, PETSC_COMM_SELF, PETSC_ERR_ORDER, "setupFFTPrec3D: no context (PCShellSetContext / PCShellFFT3DAttach)");
    PetscCheck(ctx->spaceDim >= 1 && ctx->spaceDim <= 3, PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE,
               "Dimension should be 1, 2 or 3");
    // always three extents, 1 on unused axes (the reference passes ndim = spaceDim with a 3-entry dims array, :34-35)
    PetscInt dims[3] = { ctx->spaceDim > 2 ? ctx->n_z : 1, ctx->spaceDim > 1 ? ctx->n_y : 1, ctx->n_x };
    PetscCall(create_fft_mat(PETSC_COMM_WORLD, 3, dims, &ctx->FFT_MAT));
    PetscCall(MatCreateVecs(ctx->FFT_MAT, &ctx->Diag, NULL));                    // reference :36 (MatCreateVecsFFTW)
    PetscCall(MatCreateVecs(ctx->FFT_MAT, &ctx->b_cartesien, &ctx->b_hat));      // reference :37

    // 1-D columns and their DFTs (reference :51-66), through the same library (1-D plans on this rank alone)
    Vec c[3], ch[3];
    const PetscInt n[3] = { dims[2], dims[1], dims[0] };
    for (int a = 0; a < 3; ++a) {
        Mat F1;
        PetscInt d1[1] = { n[a] };
        PetscCall(create_fft_mat(PETSC_COMM_SELF, 1, d1, &F1));
        PetscCall(MatCreateVecs(F1, &c[a], &ch[a]));
        PetscCall(build_transport_col(c[a], n[a]));
        PetscCall(MatMult(F1, c[a], ch[a]));
        PetscCall(MatDestroy(&F1));
    }
    // Diag, distributed like b and X (this rank's z planes); solve_3D hands it to the plan on the first apply, where
    // it is recognised as separable: three 1-D tables and -- for this upwind column -- the recurrence middle pass
    PetscCall(build_diag_mat_vec_3D(ctx->Diag, ch[0], ch[1], ch[2], n[0], n[1], n[2], ctx->lambda_x, ctx->lambda_y,
                                    ctx->lambda_z));
    for (int a = 0; a < 3; ++a) {
        PetscCall(VecDestroy(&c[a]));
        PetscCall(VecDestroy(&ch[a]));
    }
    PetscFunctionReturn(PETSC_SUCCESS);
}

// reference PCSHELLFft_3D.cxx:86-99
PetscErrorCode destroyFFTPrec3D(PC pc)
{
    PetscFunctionBeginUser;
    FFTPrecTransportContext *ctx = nullptr;
    PetscCall(PCShellGetContext(pc, &ctx));
    if (!ctx) PetscFunctionReturn(PETSC_SUCCESS);
    PetscCall(VecDestroy(&ctx->Diag));
    PetscCall(VecDestroy(&ctx->b_cartesien));
    PetscCall(VecDestroy(&ctx->b_hat));
    PetscCall(MatDestroy(&ctx->FFT_MAT));
    PetscFunctionReturn(PETSC_SUCCESS);
}

static FFTPrecTransportContext *g_last_ctx = nullptr;

extern "C" {

// reference PCSHELLFft_3D.cxx:101-151 with its defects fixed (allocated context, lambda = a dt / delta).
PetscErrorCode getFFTPrec3DContextCreate(PetscInt ndim, PetscScalar dt, PetscInt nbCells, PetscScalar a_x,
                                         PetscScalar a_y, PetscScalar a_z, PetscScalar Xmin, PetscScalar Ymin,
                                         PetscScalar Zmin, PetscScalar Xmax, PetscScalar Ymax, PetscScalar Zmax,
                                         struct FFTPrecTransportContext **out)
{
    PetscFunctionBeginUser;
    PetscCheck(ndim > 0 && ndim < 4, PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "Dimension should be 1, 2 or 3");
    PetscCheck(nbCells > 0 && out, PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "nbCells must be positive");
    FFTPrecTransportContext *ctx = (FFTPrecTransportContext *)calloc(1, sizeof(FFTPrecTransportContext));
    PetscInt n = nbCells;
    if (ndim == 3) {                                   // floor(cbrt(nbCells)), robust to cbrt rounding just below
        n = (PetscInt)std::floor(std::cbrt((double)nbCells));
        while ((long long)(n + 1) * (n + 1) * (n + 1) <= nbCells) ++n;
    } else if (ndim == 2) {
        n = (PetscInt)std::floor(std::sqrt((double)nbCells));
        while ((long long)(n + 1) * (n + 1) <= nbCells) ++n;
    }
    ctx->spaceDim = ndim;
    ctx->n_x = n;
    ctx->n_y = ndim > 1 ? n : 1;
    ctx->n_z = ndim > 2 ? n : 1;
    ctx->lambda_x = a_x * dt * (double)ctx->n_x / (Xmax - Xmin);
    ctx->lambda_y = ndim > 1 ? a_y * dt * (double)ctx->n_y / (Ymax - Ymin) : PetscScalar(0);
    ctx->lambda_z = ndim > 2 ? a_z * dt * (double)ctx->n_z / (Zmax - Zmin) : PetscScalar(0);
    *out = ctx;
    g_last_ctx = ctx;
    PetscFunctionReturn(PETSC_SUCCESS);
}

struct FFTPrecTransportContext *getFFTPrec3DLastContext(void) { return g_last_ctx; }

PetscErrorCode FFTPrec3DContextFree(struct FFTPrecTransportContext **ctx)
{
    if (ctx && *ctx) {
        if (g_last_ctx == *ctx) g_last_ctx = nullptr;
        free(*ctx);
        *ctx = nullptr;
    }
    return PETSC_SUCCESS;
}

// what the reference's ToDo.md item 1 asks for: plug the three callbacks into a PCSHELL
PetscErrorCode PCShellFFT3DAttach(PC pc, struct FFTPrecTransportContext *ctx)
{
    PetscFunctionBeginUser;
    PetscCheck(pc && ctx, PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "PCShellFFT3DAttach: null argument");
    PetscCall(PCShellSetContext(pc, ctx));
    PetscCall(PCShellSetSetUp(pc, setupFFTPrec3D));
    PetscCall(PCShellSetApply(pc, applyFFT3DPrecTransport));
    PetscCall(PCShellSetDestroy(pc, destroyFFTPrec3D));
    PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode cpc_glue_applyFFT3DPrecTransport(PC pc, Vec b, Vec x) { return applyFFT3DPrecTransport(pc, b, x); }
PetscErrorCode cpc_glue_setupFFTPrec3D(PC pc) { return setupFFTPrec3D(pc); }
PetscErrorCode cpc_glue_destroyFFTPrec3D(PC pc) { return destroyFFTPrec3D(pc); }

}  // extern "C"

// reference signature (PCSHELLFft_3D.hxx:27-41): no way to hand the context back, so it is kept and can be fetched
// with getFFTPrec3DLastContext().
PetscErrorCode getFFTPrec3DContext(PetscInt ndim, PetscScalar dt, PetscInt nbCells, PetscScalar a_x, PetscScalar a_y,
                                   PetscScalar a_z, PetscScalar Xmin, PetscScalar Ymin, PetscScalar Zmin,
                                   PetscScalar Xmax, PetscScalar Ymax, PetscScalar Zmax, Mesh)
{
    FFTPrecTransportContext *ctx = nullptr;
    return getFFTPrec3DContextCreate(ndim, dt, nbCells, a_x, a_y, a_z, Xmin, Ymin, Zmin, Xmax, Ymax, Zmax, &ctx);
}
