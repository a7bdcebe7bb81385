/* petsc_standin.c -- TEST INFRASTRUCTURE ONLY: see petsc_standin.h. */
#include "petsc_standin.h"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

struct _p_Vec { PetscInt n; PetscScalar *a; };
/* ndim > 0: MATFFTW over dims (N points); ndim == 0: a dense rows x cols matrix standing in for MATAIJ */
struct _p_Mat { PetscInt ndim; PetscInt dims[3]; PetscInt N; PetscInt rows, cols; PetscScalar *v; };

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
    Mat M = (Mat)calloc(1, sizeof(*M));
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
    if (A->ndim == 0) return PETSC_ERR_ARG_WRONG;
    if (x) PetscCall(VecCreateSeq(PETSC_COMM_SELF, A->N, x));
    if (y) PetscCall(VecCreateSeq(PETSC_COMM_SELF, A->N, y));
    if (z) PetscCall(VecCreateSeq(PETSC_COMM_SELF, A->N, z));
    return PETSC_SUCCESS;
}
PetscErrorCode MatDestroy(Mat *A)
{
    if (A && *A) { free((*A)->v); free(*A); *A = NULL; }
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
static PetscErrorCode dense_mult(Mat A, Vec x, Vec y, int transpose)
{
    const PetscInt m = transpose ? A->cols : A->rows, n = transpose ? A->rows : A->cols;
    if (x->n != n || y->n != m || x == y) return PETSC_ERR_ARG_WRONG;
    for (PetscInt i = 0; i < m; ++i) {
        PetscScalar acc = 0.0;
        for (PetscInt j = 0; j < n; ++j) acc += (transpose ? A->v[(size_t)j * A->cols + i] : A->v[(size_t)i * A->cols + j]) * x->a[j];
        y->a[i] = acc;
    }
    return PETSC_SUCCESS;
}
PetscErrorCode MatMult(Mat A, Vec x, Vec y) { return A->ndim ? transform(A, x, y, -1) : dense_mult(A, x, y, 0); }
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y) { return A->ndim ? transform(A, x, y, +1) : dense_mult(A, x, y, 1); }

/* ---- the rest: what the reference's C test programs need ---------------------------------------------------------- */
PetscErrorCode PetscInitialize(int *argc, char ***argv, const char *file, const char *help)
{
    (void)argc; (void)argv; (void)file; (void)help;
    return PETSC_SUCCESS;
}
PetscErrorCode PetscFinalize(void) { fflush(stdout); return PETSC_SUCCESS; }
PetscErrorCode PetscPrintf(MPI_Comm comm, const char *fmt, ...)
{
    (void)comm;
    va_list ap;
    va_start(ap, fmt);
    vprintf(fmt, ap);
    va_end(ap);
    return PETSC_SUCCESS;
}
PetscErrorCode VecNorm(Vec v, NormType type, PetscReal *nrm)
{
    long double s = 0.0L;
    for (PetscInt i = 0; i < v->n; ++i) {
        const long double m = cabs(v->a[i]);
        if (type == NORM_1) s += m;
        else if (type == NORM_2) s += m * m;
        else if (m > s) s = m;
    }
    *nrm = (PetscReal)(type == NORM_2 ? sqrtl(s) : s);
    return PETSC_SUCCESS;
}
static void print_scalar(PetscScalar z)
{
    if (cimag(z) == 0.0) printf("%.15g", creal(z));
    else printf("%.15g %c %.15g i", creal(z), cimag(z) < 0 ? '-' : '+', fabs(cimag(z)));
}
PetscErrorCode VecView(Vec v, PetscViewer viewer)
{
    (void)viewer;
    printf("Vec Object: 1 MPI process\n  type: seq\n");
    for (PetscInt i = 0; i < v->n; ++i) { print_scalar(v->a[i]); printf("\n"); }
    return PETSC_SUCCESS;
}
PetscErrorCode MatCreateAIJ(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt M, PetscInt N, PetscInt d_nz, const PetscInt *d_nnz,
                            PetscInt o_nz, const PetscInt *o_nnz, Mat *A)
{
    (void)comm; (void)d_nz; (void)d_nnz; (void)o_nz; (void)o_nnz;
    if (M < 0) M = m;
    if (N < 0) N = n;
    if (M < 1 || N < 1 || !A) return PETSC_ERR_ARG_OUTOFRANGE;
    Mat B = (Mat)calloc(1, sizeof(*B));
    B->rows = M; B->cols = N;
    B->v = (PetscScalar *)calloc((size_t)M * N, sizeof(PetscScalar));
    *A = B;
    return PETSC_SUCCESS;
}
PetscErrorCode MatSetValue(Mat A, PetscInt i, PetscInt j, PetscScalar v, InsertMode mode)
{
    if (A->ndim || i < 0 || i >= A->rows || j < 0 || j >= A->cols) return PETSC_ERR_ARG_OUTOFRANGE;
    if (mode == ADD_VALUES) A->v[(size_t)i * A->cols + j] += v; else A->v[(size_t)i * A->cols + j] = v;
    return PETSC_SUCCESS;
}
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t) { (void)A; (void)t; return PETSC_SUCCESS; }
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t) { (void)A; (void)t; return PETSC_SUCCESS; }
PetscErrorCode MatShift(Mat A, PetscScalar a)
{
    if (A->ndim) return PETSC_ERR_ARG_WRONG;
    for (PetscInt i = 0; i < A->rows && i < A->cols; ++i) A->v[(size_t)i * A->cols + i] += a;
    return PETSC_SUCCESS;
}
PetscErrorCode MatSeqAIJKron(Mat A, Mat B, MatReuse reuse, Mat *C)
{
    (void)reuse;
    if (A->ndim || B->ndim) return PETSC_ERR_ARG_WRONG;
    PetscCall(MatCreateAIJ(PETSC_COMM_SELF, A->rows * B->rows, A->cols * B->cols, A->rows * B->rows, A->cols * B->cols, 0, NULL, 0,
                           NULL, C));
    Mat K = *C;
    for (PetscInt ia = 0; ia < A->rows; ++ia)
        for (PetscInt ja = 0; ja < A->cols; ++ja) {
            const PetscScalar a = A->v[(size_t)ia * A->cols + ja];
            if (a == 0.0) continue;
            for (PetscInt ib = 0; ib < B->rows; ++ib)
                for (PetscInt jb = 0; jb < B->cols; ++jb)
                    K->v[(size_t)(ia * B->rows + ib) * K->cols + (ja * B->cols + jb)] = a * B->v[(size_t)ib * B->cols + jb];
        }
    return PETSC_SUCCESS;
}
PetscErrorCode MatDuplicate(Mat A, MatDuplicateOption op, Mat *B)
{
    if (A->ndim) return PETSC_ERR_ARG_WRONG;
    PetscCall(MatCreateAIJ(PETSC_COMM_SELF, A->rows, A->cols, A->rows, A->cols, 0, NULL, 0, NULL, B));
    if (op == MAT_COPY_VALUES) memcpy((*B)->v, A->v, sizeof(PetscScalar) * (size_t)A->rows * A->cols);
    return PETSC_SUCCESS;
}
PetscErrorCode MatAXPY(Mat Y, PetscScalar a, Mat X, MatStructure str)
{
    (void)str;
    if (Y->ndim || X->ndim || Y->rows != X->rows || Y->cols != X->cols) return PETSC_ERR_ARG_WRONG;
    for (size_t k = 0; k < (size_t)Y->rows * Y->cols; ++k) Y->v[k] += a * X->v[k];
    return PETSC_SUCCESS;
}
PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left)
{
    if (right) PetscCall(VecCreateSeq(PETSC_COMM_SELF, A->ndim ? A->N : A->cols, right));
    if (left) PetscCall(VecCreateSeq(PETSC_COMM_SELF, A->ndim ? A->N : A->rows, left));
    return PETSC_SUCCESS;
}
PetscErrorCode MatGetType(Mat A, MatType *type) { *type = A->ndim ? MATFFTW : MATAIJ; return PETSC_SUCCESS; }
PetscErrorCode MatView(Mat A, PetscViewer viewer)
{
    (void)viewer;
    if (A->ndim) { printf("Mat Object: type fftw, %d points\n", (int)A->N); return PETSC_SUCCESS; }
    printf("Mat Object: 1 MPI process\n  type: seqaij\n");
    for (PetscInt i = 0; i < A->rows; ++i) {
        printf("row %d:", (int)i);
        for (PetscInt j = 0; j < A->cols; ++j)
            if (A->v[(size_t)i * A->cols + j] != 0.0) { printf(" (%d, ", (int)j); print_scalar(A->v[(size_t)i * A->cols + j]); printf(")"); }
        printf("\n");
    }
    return PETSC_SUCCESS;
}
