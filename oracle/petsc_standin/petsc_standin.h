/* petsc_standin.h -- TEST INFRASTRUCTURE ONLY.  A minimal CPU stand-in for the ~30 PETSc calls that the reference's
 * src/FftLinearSolver_3D.c makes, so that THAT FILE, unmodified and compiled from where it lies under /root/reference,
 * can run in this image (no PETSc, FFTW or MPI here) and pin the oracle: oracle/Makefile target
 * _ref/libreference_fftsolver.so, used by tests/test_reference_c.py only.
 *
 * What is the reference's and what is ours: everything FftLinearSolver_3D.c does itself -- the transport column
 * (:80-90), the Kronecker layout of Diag (:92-164), the order forward transform / VecPointwiseDivide / backward
 * transform / VecScale(1/size) (:166-190), the wrappers with their lambdas and degenerate axes (:192-312) -- runs as
 * written.  The stand-in supplies what PETSc and FFTW would: sequential complex Vecs, and MATFFTW as a plain O(n^2)-per-
 * line DFT with FFTW's conventions (unnormalised, exp(-2 pi i ..) for MatMult, exp(+..) for MatMultTranspose, dims slowest
 * first).  Complex build (PETSC_USE_COMPLEX), one rank.  Semantics follow the PETSc manual pages of each call; where the
 * reference leans on an implementation detail it is said here: VecDuplicate returns zeroed storage (PETSc's sequential
 * Vecs are calloc'ed; build_diag_mat_vec_3D :151-152 accumulates into a fresh duplicate without VecSet).
 */
#ifndef PETSC_STANDIN_H
#define PETSC_STANDIN_H
#include <complex.h>
#include <stddef.h>

#define PETSC_USE_COMPLEX 1
typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscReal;
typedef double _Complex PetscScalar;
typedef int PetscBool;
typedef int MPI_Comm;
typedef const char *MatType;
#define PETSC_SUCCESS 0
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_ARG_OUTOFRANGE 63
#define PETSC_COMM_WORLD 1
#define PETSC_COMM_SELF 2
#define MATFFTW "fftw"
typedef enum { NOT_SET_VALUES = 0, INSERT_VALUES = 1, ADD_VALUES = 2 } InsertMode;

#define PetscFunctionBeginUser do { } while (0)
#define PetscFunctionReturn(v) return (v)
#define PetscCall(call) do { PetscErrorCode ierr_standin_ = (call); if (ierr_standin_) return ierr_standin_; } while (0)
#define PetscCheck(cond, comm, code, ...) do { if (!(cond)) return (code); } while (0)

typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;

PetscErrorCode VecCreateSeq(MPI_Comm comm, PetscInt n, Vec *v);
PetscErrorCode VecDuplicate(Vec v, Vec *out);
PetscErrorCode VecDestroy(Vec *v);
PetscErrorCode VecSet(Vec v, PetscScalar a);
PetscErrorCode VecSetValue(Vec v, PetscInt i, PetscScalar a, InsertMode mode);
PetscErrorCode VecSetValues(Vec v, PetscInt n, const PetscInt *idx, const PetscScalar *a, InsertMode mode);
PetscErrorCode VecGetValues(Vec v, PetscInt n, const PetscInt *idx, PetscScalar *a);
PetscErrorCode VecAssemblyBegin(Vec v);
PetscErrorCode VecAssemblyEnd(Vec v);
PetscErrorCode VecGetArray(Vec v, PetscScalar **a);
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a);
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a);
PetscErrorCode VecRestoreArrayRead(Vec v, const PetscScalar **a);
PetscErrorCode VecGetSize(Vec v, PetscInt *n);
PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n);
PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt *lo, PetscInt *hi);
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x);          /* y += a x */
PetscErrorCode VecShift(Vec v, PetscScalar s);
PetscErrorCode VecCopy(Vec x, Vec y);                          /* y = x */
PetscErrorCode VecScale(Vec v, PetscScalar a);
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y);        /* w = x ./ y */

PetscErrorCode MatCreateFFT(MPI_Comm comm, PetscInt ndim, const PetscInt dims[], MatType type, Mat *A);
PetscErrorCode MatCreateVecsFFTW(Mat A, Vec *x, Vec *y, Vec *z);
PetscErrorCode MatMult(Mat A, Vec x, Vec y);                   /* unnormalised forward DFT */
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y);          /* unnormalised backward DFT */
PetscErrorCode MatDestroy(Mat *A);
#endif
