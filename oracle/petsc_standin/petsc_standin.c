/* petsc_standin.c -- TEST INFRASTRUCTURE ONLY: see petsc_standin.h. */
#include "petsc_standin.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

struct _p_Vec { PetscInt n; PetscScalar *a; };
struct _p_Mat { PetscInt ndim; PetscInt dims[3]; PetscInt N; };

PetscErrorCode VecCreateSeq(MPI_Comm comm, PetscInt n, Vec *v)
{
    (void)comm;
    if (n < 0 || !v) return PETSC_ERR_ARG_OUTOFRANGE;
    Vec w = (Vec)malloc(sizeof(*w));
    w->n = n;
    w->a = (PetscScalar *)calloc((size_t)(n > 0 ? n : 1), sizeof(PetscScalar));
    *v = w;
    return PETSC_SUCCESS;
}
PetscErrorCode VecDuplicate(Vec v, Vec *out) { return VecCreateSeq(PETSC_COMM_SELF, v->n, out); }
PetscErrorCode VecDestroy(Vec *v)
{
    if (v && *v) { free((*v)->a); free(*v); *v = NULL; }
    return PETSC_SUCCESS;
}
PetscErrorCode VecSet(Vec v, PetscScalar a)
{
    for (PetscInt i = 0; i < v->n; ++i) v->a[i] = a;
    return PETSC_SUCCESS;
}
PetscErrorCode VecSetValue(Vec v, PetscInt i, PetscScalar a, InsertMode mode)
{
    if (i < 0 || i >= v->n) return PETSC_ERR_ARG_OUTOFRANGE;
    if (mode == ADD_VALUES) v->a[i] += a; else v->a[i] = a;
    return PETSC_SUCCESS;
}
PetscErrorCode VecSetValues(Vec v, PetscInt n, const PetscInt *idx, const PetscScalar *a, InsertMode mode)
{
    for (PetscInt k = 0; k < n; ++k) PetscCall(VecSetValue(v, idx[k], a[k], mode));
    return PETSC_SUCCESS;
}
PetscErrorCode VecGetValues(Vec v, PetscInt n, const PetscInt *idx, PetscScalar *a)
{
    for (PetscInt k = 0; k < n; ++k) {
        if (idx[k] < 0 || idx[k] >= v->n) return PETSC_ERR_ARG_OUTOFRANGE;
        a[k] = v->a[idx[k]];
    }
    return PETSC_SUCCESS;
}
PetscErrorCode VecAssemblyBegin(Vec v) { (void)v; return PETSC_SUCCESS; }
PetscErrorCode VecAssemblyEnd(Vec v) { (void)v; return PETSC_SUCCESS; }
PetscErrorCode VecGetArray(Vec v, PetscScalar **a) { *a = v->a; return PETSC_SUCCESS; }
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a) { (void)v; if (a) *a = NULL; return PETSC_SUCCESS; }
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a) { *a = v->a; return PETSC_SUCCESS; }
PetscErrorCode VecRestoreArrayRead(Vec v, const PetscScalar **a) { (void)v; if (a) *a = NULL; return PETSC_SUCCESS; }
PetscErrorCode VecGetSize(Vec v, PetscInt *n) { *n = v->n; return PETSC_SUCCESS; }
PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n) { *n = v->n; return PETSC_SUCCESS; }
PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt *lo, PetscInt *hi)
{
    if (lo) *lo = 0;
    if (hi) *hi = v->n;
    return PETSC_SUCCESS;
}
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x)
{
    if (x->n != y->n) return PETSC_ERR_ARG_WRONG;
    for (PetscInt i = 0; i < y->n; ++i) y->a[i] += a * x->a[i];
    return PETSC_SUCCESS;
}
PetscErrorCode VecShift(Vec v, PetscScalar s)
{
    for (PetscInt i = 0; i < v->n; ++i) v->a[i] += s;
    return PETSC_SUCCESS;
}
PetscErrorCode VecCopy(Vec x, Vec y)
{
    if (x->n != y->n) return PETSC_ERR_ARG_WRONG;
    if (x != y) memcpy(y->a, x->a, sizeof(PetscScalar) * (size_t)x->n);
    return PETSC_SUCCESS;
}
PetscErrorCode VecScale(Vec v, PetscScalar a)
{
    for (PetscInt i = 0; i < v->n; ++i) v->a[i] *= a;
    return PETSC_SUCCESS;
}
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y)
{
    if (w->n != x->n || w->n != y->n) return PETSC_ERR_ARG_WRONG;
    for (PetscInt i = 0; i < w->n; ++i) w->a[i] = x->a[i] / y->a[i];
    return PETSC_SUCCESS;
}

/* ---- MATFFTW: unnormalised multi-dimensional DFT, dims slowest first (fftw_plan_dft's row-major convention) ---- */
PetscErrorCode MatCreateFFT(MPI_Comm comm, PetscInt ndim, const PetscInt dims[], MatType type, Mat *A)
{
    (void)comm; (void)type;
    if (ndim < 1 || ndim > 3 || !dims || !A) return PETSC_ERR_ARG_OUTOFRANGE;
    Mat M = (Mat)malloc(sizeof(*M));
    M->ndim = ndim;
    M->N = 1;
    for (PetscInt d = 0; d < ndim; ++d) {
        if (dims[d] < 1) { free(M); return PETSC_ERR_ARG_OUTOFRANGE; }
        M->dims[d] = dims[d];
        M->N *= dims[d];
    }
    *A = M;
    return PETSC_SUCCESS;
}
PetscErrorCode MatCreateVecsFFTW(Mat A, Vec *x, Vec *y, Vec *z)
{
    if (x) PetscCall(VecCreateSeq(PETSC_COMM_SELF, A->N, x));
    if (y) PetscCall(VecCreateSeq(PETSC_COMM_SELF, A->N, y));
    if (z) PetscCall(VecCreateSeq(PETSC_COMM_SELF, A->N, z));
    return PETSC_SUCCESS;
}
PetscErrorCode MatDestroy(Mat *A)
{
    if (A && *A) { free(*A); *A = NULL; }
    return PETSC_SUCCESS;
}

/* DFT of length n along an axis with the given stride, for every line; sign -1 forward, +1 backward.  Direct O(n^2) sum
 * with a root table: slow and obviously the definition. */
static void dft_axis(PetscScalar *a, PetscInt N, PetscInt n, PetscInt stride, int sign)
{
    if (n == 1) return;
    PetscScalar *w = (PetscScalar *)malloc(sizeof(PetscScalar) * (size_t)n);
    PetscScalar *t = (PetscScalar *)malloc(sizeof(PetscScalar) * (size_t)n);
    for (PetscInt m = 0; m < n; ++m) {
        const long double ang = 2.0L * 3.141592653589793238462643383279502884L * (long double)m / (long double)n;
        w[m] = (double)cosl(ang) + sign * (double)sinl(ang) * I;
    }
    const PetscInt outer = N / (n * stride);
    for (PetscInt o = 0; o < outer; ++o)
        for (PetscInt s = 0; s < stride; ++s) {
            PetscScalar *line = a + (size_t)o * n * stride + s;
            for (PetscInt k = 0; k < n; ++k) {
                PetscScalar acc = 0.0;
                for (PetscInt j = 0; j < n; ++j) acc += line[(size_t)j * stride] * w[(PetscInt)(((long long)j * k) % n)];
                t[k] = acc;
            }
            for (PetscInt k = 0; k < n; ++k) line[(size_t)k * stride] = t[k];
        }
    free(w);
    free(t);
}
static PetscErrorCode transform(Mat A, Vec x, Vec y, int sign)
{
    if (x->n != A->N || y->n != A->N) return PETSC_ERR_ARG_WRONG;
    if (x != y) memcpy(y->a, x->a, sizeof(PetscScalar) * (size_t)A->N);
    PetscInt stride = 1;
    for (PetscInt d = A->ndim - 1; d >= 0; --d) {       /* the last dimension is the contiguous one */
        dft_axis(y->a, A->N, A->dims[d], stride, sign);
        stride *= A->dims[d];
    }
    return PETSC_SUCCESS;
}
PetscErrorCode MatMult(Mat A, Vec x, Vec y) { return transform(A, x, y, -1); }
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y) { return transform(A, x, y, +1); }
