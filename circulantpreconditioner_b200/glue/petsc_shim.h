/* petsc_shim.h -- the ~40 PETSc names the reference's hot path touches, implemented over plain host arrays.
 *
 * PETSc is not installable in this image (no network), so the reference-named glue (FftLinearSolver_3D.cxx,
 * PCSHELLFft_3D.cxx) is compiled and exercised against this shim.  With a real PETSc the same glue sources
 * compile against <petscksp.h> instead (define CPC_WITH_PETSC; INTEGRATION.md shows the four lines that differ:
 * how a Vec's array is obtained and how the cpc plan is attached to the FFT Mat).
 *
 * Semantics follow a complex-scalar PETSc build (PetscScalar = complex128), the branch of the reference whose
 * semantics are well defined (reference src/FftLinearSolver_3D.c:173-175,183-184).
 */
#ifndef CPC_PETSC_SHIM_H
#define CPC_PETSC_SHIM_H

#include <complex>
#include <cstddef>

#include "../../include/circulantpc.h"

typedef int PetscInt;
typedef int PetscErrorCode;
typedef int PetscMPIInt;
typedef double PetscReal;
typedef std::complex<double> PetscScalar;
typedef bool PetscBool;
typedef int MPI_Comm;
typedef const char *MatType;

#define PETSC_USE_COMPLEX 1
#define PETSC_SUCCESS 0
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_ARG_OUTOFRANGE 63
#define PETSC_ERR_LIB 76
#define PETSC_ERR_ORDER 58
#define PETSC_COMM_WORLD 0
#define PETSC_DECIDE (-1)
#define MATFFTW "fftw"
enum InsertMode { INSERT_VALUES = 1, ADD_VALUES = 2 };

struct _p_Vec {
    PetscInt n;
    PetscScalar *array;
    unsigned long state;        /* bumped by every write access (PetscObjectStateGet in real PETSc) */
};
typedef struct _p_Vec *Vec;

enum ShimMatKind { SHIM_MAT_FFT = 1, SHIM_MAT_CSR = 2 };
struct _p_Mat {
    int kind;
    /* SHIM_MAT_FFT: the B200 plan that replaces the FFTW plan behind MATFFTW */
    PetscInt ndim;
    PetscInt dims[3];           /* as given to MatCreateFFT: slowest first ({nz, ny, nx}) */
    cpc_plan plan;
    const void *diag_seen;      /* Diag array + state last uploaded through solve_3D */
    unsigned long diag_state;
    const void *proj_seen;      /* projection Mat last handed to cpc_set_projection */
    /* SHIM_MAT_CSR: projection matrix (intersectionMatrix) */
    PetscInt rows, cols;
    PetscInt *rowptr, *colidx;
    PetscScalar *val;
};
typedef struct _p_Mat *Mat;

struct _p_PC {
    void *ctx;
    PetscErrorCode (*apply)(struct _p_PC *, Vec, Vec);
    PetscErrorCode (*setup)(struct _p_PC *);
    PetscErrorCode (*destroy)(struct _p_PC *);
    bool is_setup;
};
typedef struct _p_PC *PC;

extern "C" {
const char *ShimLastError(void);
PetscErrorCode ShimError(PetscErrorCode code, const char *fmt, ...);

PetscErrorCode VecCreateSeq(MPI_Comm, PetscInt n, Vec *v);
PetscErrorCode VecDuplicate(Vec v, Vec *w);
PetscErrorCode VecDestroy(Vec *v);
PetscErrorCode VecGetSize(Vec v, PetscInt *n);
PetscErrorCode VecSet(Vec v, PetscScalar a);
PetscErrorCode VecSetValue(Vec v, PetscInt i, PetscScalar a, InsertMode mode);
PetscErrorCode VecAssemblyBegin(Vec v);
PetscErrorCode VecAssemblyEnd(Vec v);
PetscErrorCode VecGetArray(Vec v, PetscScalar **a);
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a);
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a);
PetscErrorCode VecRestoreArrayRead(Vec v, const PetscScalar **a);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecScale(Vec v, PetscScalar a);
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x);
PetscErrorCode VecShift(Vec v, PetscScalar a);
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y);
PetscErrorCode VecNorm2(Vec v, PetscReal *nrm);

PetscErrorCode MatCreateFFT(MPI_Comm, PetscInt ndim, const PetscInt dims[], MatType, Mat *A);
PetscErrorCode MatCreateVecsFFTW(Mat A, Vec *x, Vec *y, Vec *z);
PetscErrorCode MatCreateSeqAIJFromCSR(PetscInt rows, PetscInt cols, const PetscInt *rowptr, const PetscInt *colidx,
                                      const PetscScalar *val, Mat *A);
PetscErrorCode MatDestroy(Mat *A);
PetscErrorCode MatMult(Mat A, Vec x, Vec y);
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y);

PetscErrorCode PCCreate(MPI_Comm, PC *pc);
PetscErrorCode PCShellSetContext(PC pc, void *ctx);
PetscErrorCode PCShellGetContext(PC pc, void *ctx_out);   /* void** semantics, as in PETSc */
PetscErrorCode PCShellSetApply(PC pc, PetscErrorCode (*f)(PC, Vec, Vec));
PetscErrorCode PCShellSetSetUp(PC pc, PetscErrorCode (*f)(PC));
PetscErrorCode PCShellSetDestroy(PC pc, PetscErrorCode (*f)(PC));
PetscErrorCode PCSetUp(PC pc);
PetscErrorCode PCApply(PC pc, Vec b, Vec x);
PetscErrorCode PCDestroy(PC *pc);
}

#define PetscFunctionBeginUser do { } while (0)
#define PetscFunctionReturn(x) return (x)
#define PetscCall(call)                                       \
    do {                                                      \
        PetscErrorCode _ierr = (call);                        \
        if (_ierr) return _ierr;                              \
    } while (0)
#define PetscCheck(cond, comm, code, ...)                     \
    do {                                                      \
        if (!(cond)) return ShimError((code), __VA_ARGS__);   \
    } while (0)
/* a libcirculantpc status becomes a PETSc error carrying cpc_last_error() */
#define PetscCallCPC(call)                                                                  \
    do {                                                                                    \
        int _st = (call);                                                                   \
        if (_st) return ShimError(_st == CPC_ERR_ARG ? PETSC_ERR_ARG_WRONG : PETSC_ERR_LIB, \
                                  "libcirculantpc: %s", cpc_last_error());                  \
    } while (0)

#endif
