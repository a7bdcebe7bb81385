// petsc_shim.cxx -- host-array implementation of the PETSc subset declared in petsc_shim.h (test scaffolding for
// the glue; the arithmetic of the hot path is NOT here: MatMult/MatMultTranspose on an FFT Mat call libcirculantpc).
#include "petsc_shim.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static thread_local char g_msg[512] = "";

extern "C" {

const char *ShimLastError(void) { return g_msg; }

PetscErrorCode ShimError(PetscErrorCode code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_msg, sizeof(g_msg), fmt, ap);
    va_end(ap);
    return code;
}

PetscErrorCode VecCreateSeq(MPI_Comm, PetscInt n, Vec *v)
{
    if (n < 0 || !v) return ShimError(PETSC_ERR_ARG_OUTOFRANGE, "VecCreateSeq: bad size");
    Vec w = (Vec)calloc(1, sizeof(_p_Vec));
    w->n = n;
    w->array = (PetscScalar *)calloc((size_t)(n > 0 ? n : 1), sizeof(PetscScalar));
    *v = w;
    return PETSC_SUCCESS;
}
PetscErrorCode VecDuplicate(Vec v, Vec *w) { return VecCreateSeq(0, v->n, w); }
PetscErrorCode VecDestroy(Vec *v)
{
    if (v && *v) { free((*v)->array); free(*v); *v = nullptr; }
    return PETSC_SUCCESS;
}
PetscErrorCode VecGetSize(Vec v, PetscInt *n) { *n = v->n; return PETSC_SUCCESS; }
PetscErrorCode VecSet(Vec v, PetscScalar a)
{
    for (PetscInt i = 0; i < v->n; ++i) v->array[i] = a;
    ++v->state;
    return PETSC_SUCCESS;
}
PetscErrorCode VecSetValue(Vec v, PetscInt i, PetscScalar a, InsertMode mode)
{
    if (i < 0 || i >= v->n) return ShimError(PETSC_ERR_ARG_OUTOFRANGE, "VecSetValue: index %d out of range", i);
    if (mode == ADD_VALUES) v->array[i] += a; else v->array[i] = a;
    ++v->state;
    return PETSC_SUCCESS;
}
PetscErrorCode VecAssemblyBegin(Vec) { return PETSC_SUCCESS; }
PetscErrorCode VecAssemblyEnd(Vec) { return PETSC_SUCCESS; }
PetscErrorCode VecGetArray(Vec v, PetscScalar **a) { *a = v->array; ++v->state; return PETSC_SUCCESS; }
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a) { if (a) *a = nullptr; ++v->state; return PETSC_SUCCESS; }
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a) { *a = v->array; return PETSC_SUCCESS; }
PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar **a) { if (a) *a = nullptr; return PETSC_SUCCESS; }
PetscErrorCode VecCopy(Vec x, Vec y)
{
    if (x->n != y->n) return ShimError(PETSC_ERR_ARG_WRONG, "VecCopy: size mismatch");
    if (x != y) memcpy(y->array, x->array, sizeof(PetscScalar) * (size_t)x->n);
    ++y->state;
    return PETSC_SUCCESS;
}
PetscErrorCode VecScale(Vec v, PetscScalar a)
{
    for (PetscInt i = 0; i < v->n; ++i) v->array[i] *= a;
    ++v->state;
    return PETSC_SUCCESS;
}
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x)
{
    if (x->n != y->n) return ShimError(PETSC_ERR_ARG_WRONG, "VecAXPY: size mismatch");
    for (PetscInt i = 0; i < y->n; ++i) y->array[i] += a * x->array[i];
    ++y->state;
    return PETSC_SUCCESS;
}
PetscErrorCode VecShift(Vec v, PetscScalar a)
{
    for (PetscInt i = 0; i < v->n; ++i) v->array[i] += a;
    ++v->state;
    return PETSC_SUCCESS;
}
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y)
{
    for (PetscInt i = 0; i < w->n; ++i) w->array[i] = x->array[i] / y->array[i];
    ++w->state;
    return PETSC_SUCCESS;
}
PetscErrorCode VecNorm2(Vec v, PetscReal *nrm)
{
    long double s = 0;
    for (PetscInt i = 0; i < v->n; ++i) s += std::norm(v->array[i]);
    *nrm = (PetscReal)sqrtl(s);
    return PETSC_SUCCESS;
}

// MatCreateFFT(comm, ndim, dims, MATFFTW, &A): the FFTW plan behind MATFFTW becomes a libcirculantpc plan
// (reference call sites: src/PCSHELLFft_3D.cxx:34-35, tests/TransportEquationFFT_...:97-100; dims slowest first).
PetscErrorCode MatCreateFFT(MPI_Comm, PetscInt ndim, const PetscInt dims[], MatType, Mat *A)
{
    if (ndim < 1 || ndim > 3 || !dims || !A) return ShimError(PETSC_ERR_ARG_OUTOFRANGE, "MatCreateFFT: ndim must be 1..3");
    Mat M = (Mat)calloc(1, sizeof(_p_Mat));
    M->kind = SHIM_MAT_FFT;
    M->ndim = ndim;
    PetscInt n[3] = { 1, 1, 1 };     // nx, ny, nz (x fastest = last entry of dims)
    for (PetscInt d = 0; d < ndim; ++d) { M->dims[d] = dims[d]; n[ndim - 1 - d] = dims[d]; }
    cpc_plan_desc desc = { n[0], n[1], n[2], 1, CPC_C128, 1, 0, nullptr, nullptr, -1 };
    int st = cpc_plan_create(&M->plan, &desc);
    if (st) { free(M); return ShimError(PETSC_ERR_LIB, "libcirculantpc: %s", cpc_last_error()); }
    *A = M;
    return PETSC_SUCCESS;
}
PetscErrorCode MatCreateVecsFFTW(Mat A, Vec *x, Vec *y, Vec *z)
{
    if (!A || A->kind != SHIM_MAT_FFT) return ShimError(PETSC_ERR_ARG_WRONG, "MatCreateVecsFFTW: not an FFT matrix");
    PetscInt N = 1;
    for (PetscInt d = 0; d < A->ndim; ++d) N *= A->dims[d];
    if (x) PetscCall(VecCreateSeq(0, N, x));
    if (y) PetscCall(VecCreateSeq(0, N, y));
    if (z) PetscCall(VecCreateSeq(0, N, z));
    return PETSC_SUCCESS;
}
PetscErrorCode MatCreateSeqAIJFromCSR(PetscInt rows, PetscInt cols, const PetscInt *rowptr, const PetscInt *colidx,
                                      const PetscScalar *val, Mat *A)
{
    Mat M = (Mat)calloc(1, sizeof(_p_Mat));
    M->kind = SHIM_MAT_CSR;
    M->rows = rows; M->cols = cols;
    const PetscInt nnz = rowptr[rows];
    M->rowptr = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(rows + 1));
    M->colidx = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(nnz > 0 ? nnz : 1));
    M->val = (PetscScalar *)malloc(sizeof(PetscScalar) * (size_t)(nnz > 0 ? nnz : 1));
    memcpy(M->rowptr, rowptr, sizeof(PetscInt) * (size_t)(rows + 1));
    memcpy(M->colidx, colidx, sizeof(PetscInt) * (size_t)nnz);
    memcpy(M->val, val, sizeof(PetscScalar) * (size_t)nnz);
    *A = M;
    return PETSC_SUCCESS;
}
PetscErrorCode MatDestroy(Mat *A)
{
    if (A && *A) {
        if ((*A)->plan) cpc_destroy((*A)->plan);
        free((*A)->rowptr); free((*A)->colidx); free((*A)->val);
        free(*A);
        *A = nullptr;
    }
    return PETSC_SUCCESS;
}
PetscErrorCode MatMult(Mat A, Vec x, Vec y)
{
    if (A->kind == SHIM_MAT_FFT) {          // unnormalised forward DFT (reference FftLinearSolver_3D.c:170)
        PetscCallCPC(cpc_forward(A->plan, x->array, y->array, CPC_MEM_HOST));
        ++y->state;
        return PETSC_SUCCESS;
    }
    if (x->n != A->cols || y->n != A->rows) return ShimError(PETSC_ERR_ARG_WRONG, "MatMult: size mismatch");
    for (PetscInt i = 0; i < A->rows; ++i) {
        PetscScalar s = 0;
        for (PetscInt p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) s += A->val[p] * x->array[A->colidx[p]];
        y->array[i] = s;
    }
    ++y->state;
    return PETSC_SUCCESS;
}
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y)
{
    if (A->kind == SHIM_MAT_FFT) {          // unnormalised backward DFT (reference FftLinearSolver_3D.c:180)
        PetscCallCPC(cpc_inverse(A->plan, x->array, y->array, CPC_MEM_HOST));
        ++y->state;
        return PETSC_SUCCESS;
    }
    if (x->n != A->rows || y->n != A->cols) return ShimError(PETSC_ERR_ARG_WRONG, "MatMultTranspose: size mismatch");
    for (PetscInt i = 0; i < A->cols; ++i) y->array[i] = 0;
    for (PetscInt i = 0; i < A->rows; ++i)
        for (PetscInt p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) y->array[A->colidx[p]] += A->val[p] * x->array[i];
    ++y->state;
    return PETSC_SUCCESS;
}

PetscErrorCode PCCreate(MPI_Comm, PC *pc) { *pc = (PC)calloc(1, sizeof(_p_PC)); return PETSC_SUCCESS; }
PetscErrorCode PCShellSetContext(PC pc, void *ctx) { pc->ctx = ctx; return PETSC_SUCCESS; }
PetscErrorCode PCShellGetContext(PC pc, void *ctx_out) { *(void **)ctx_out = pc->ctx; return PETSC_SUCCESS; }
PetscErrorCode PCShellSetApply(PC pc, PetscErrorCode (*f)(PC, Vec, Vec)) { pc->apply = f; return PETSC_SUCCESS; }
PetscErrorCode PCShellSetSetUp(PC pc, PetscErrorCode (*f)(PC)) { pc->setup = f; return PETSC_SUCCESS; }
PetscErrorCode PCShellSetDestroy(PC pc, PetscErrorCode (*f)(PC)) { pc->destroy = f; return PETSC_SUCCESS; }
PetscErrorCode PCSetUp(PC pc)
{
    if (!pc->is_setup && pc->setup) PetscCall(pc->setup(pc));
    pc->is_setup = true;
    return PETSC_SUCCESS;
}
PetscErrorCode PCApply(PC pc, Vec b, Vec x)
{
    if (!pc->apply) return ShimError(PETSC_ERR_ORDER, "PCApply: no apply callback registered");
    PetscCall(PCSetUp(pc));
    return pc->apply(pc, b, x);
}
PetscErrorCode PCDestroy(PC *pc)
{
    if (pc && *pc) {
        PetscErrorCode ierr = ((*pc)->destroy && (*pc)->is_setup) ? (*pc)->destroy(*pc) : 0;
        free(*pc);
        *pc = nullptr;
        return ierr;
    }
    return PETSC_SUCCESS;
}

}  // extern "C"
