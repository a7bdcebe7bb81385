/* petsc_standin.h -- TEST INFRASTRUCTURE ONLY.  A minimal CPU stand-in for the ~30 PETSc calls that the reference's
 * src/FftLinearSolver_3D.c makes, so that THAT FILE, unmodified and compiled from where it lies under /root/reference,
 * can run in this image (no PETSc, FFTW or MPI here) and pin the oracle: oracle/Makefile target
 * _ref/libreference_fftsolver.so, used by tests/test_reference_c.py only.  The reference's own C test programs
 * (tests/FFTDirectSolver/testFftSolver_{1D,2D,3D}.c) build against it too (_ref/testFftSolver_*).
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
#include <assert.h>      /* the reference's C tests call assert / strcmp / printf-style output without including these: */
#include <complex.h>     /* with a real PETSc they arrive through petscsys.h */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

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
#define MATAIJ "aij"
#define PETSC_DECIDE (-1)
#define PETSC_VIEWER_STDOUT_WORLD ((PetscViewer)0)
typedef void *PetscViewer;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_INFINITY = 3 } NormType;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { MAT_INITIAL_MATRIX = 0, MAT_REUSE_MATRIX = 1 } MatReuse;
typedef enum { MAT_DO_NOT_COPY_VALUES = 0, MAT_COPY_VALUES = 1 } MatDuplicateOption;
typedef enum { DIFFERENT_NONZERO_PATTERN = 0, SUBSET_NONZERO_PATTERN = 1, SAME_NONZERO_PATTERN = 2 } MatStructure;
typedef enum { NOT_SET_VALUES = 0, INSERT_VALUES = 1, ADD_VALUES = 2 } InsertMode;

#define PetscFunctionBeginUser do { } while (0)
#define PetscFunctionReturn(v) return (v)
#define PetscCall(call) do { PetscErrorCode ierr_standin_ = (call); if (ierr_standin_) return ierr_standin_; } while (0)
#define PetscCheck(cond, comm, code, ...) do { if (!(cond)) return (code); } while (0)
#define SETERRQ(comm, code, ...) return (code)

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

/* ---- what the reference's C test programs (tests/FFTDirectSolver/testFftSolver_{1D,2D,3D}.c) use on top: small AIJ
 * matrices (held dense here: the tests build 4 x 4 .. 24 x 24 circulant matrices by Kronecker products), norms, output */
PetscErrorCode PetscInitialize(int *argc, char ***argv, const char *file, const char *help);
PetscErrorCode PetscFinalize(void);
PetscErrorCode PetscPrintf(MPI_Comm comm, const char *fmt, ...);
PetscErrorCode VecNorm(Vec v, NormType type, PetscReal *nrm);
PetscErrorCode VecView(Vec v, PetscViewer viewer);
PetscErrorCode MatCreateAIJ(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt M, PetscInt N, PetscInt d_nz, const PetscInt *d_nnz,
                            PetscInt o_nz, const PetscInt *o_nnz, Mat *A);
PetscErrorCode MatSetValue(Mat A, PetscInt i, PetscInt j, PetscScalar v, InsertMode mode);
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t);
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t);
PetscErrorCode MatShift(Mat A, PetscScalar a);
PetscErrorCode MatSeqAIJKron(Mat A, Mat B, MatReuse reuse, Mat *C);      /* C = A (x) B */
PetscErrorCode MatDuplicate(Mat A, MatDuplicateOption op, Mat *B);
PetscErrorCode MatAXPY(Mat Y, PetscScalar a, Mat X, MatStructure str);   /* Y += a X */
PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left);
PetscErrorCode MatGetType(Mat A, MatType *type);
PetscErrorCode MatView(Mat A, PetscViewer viewer);
#endif
