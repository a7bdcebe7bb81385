// petsc_shim.cxx -- host / CUDA array implementation of the PETSc subset declared in petsc_shim.h (test scaffolding
// for the glue; the arithmetic of the hot path is NOT here: MatMult/MatMultTranspose on an FFT Mat call libcirculantpc).
#include "petsc_shim.h"

#include <cuda_runtime_api.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

// the glue's own additions (circulantpc_petsc.cxx, same shared library): the plan rides on the FFT Mat
extern "C" PetscErrorCode CPCMatAttachPlan(Mat A, MPI_Comm comm, PetscInt n_x, PetscInt n_y, PetscInt n_z);
extern "C" PetscErrorCode CPCMatGetPlan(Mat A, cpc_plan *plan);

static thread_local char g_msg[512] = "";
static int g_size = 1, g_rank = 0, g_default_cuda = 0;
static unsigned char g_nccl_id[CPC_NCCL_UNIQUE_ID_BYTES];
static bool g_have_id = false;
static PetscObjectId g_next_id = 1;

struct ShimComposed {
    char name[64];
    PetscObject obj;
    ShimComposed *next;
};

static void hdr_init(_p_PetscObject *h, int classid)
{
    h->classid = classid;
    h->id = g_next_id++;
    h->state = 0;
    h->refct = 1;
    h->composed = nullptr;
}

// destroying an object destroys the containers composed onto it (PETSc drops the reference it holds)
static void hdr_release(_p_PetscObject *h)
{
    for (ShimComposed *c = h->composed; c;) {
        ShimComposed *nx = c->next;
        if (c->obj && c->obj->classid == 4) {
            PetscContainer ct = (PetscContainer)c->obj;
            PetscContainerDestroy(&ct);
        }
        free(c);
        c = nx;
    }
    h->composed = nullptr;
}

#define CUDA_OK(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t _e = (call);                                                                           \
        if (_e != cudaSuccess) return ShimError(PETSC_ERR_LIB, "CUDA: %s in %s", cudaGetErrorString(_e), #call); \
    } while (0)

// offload mask of a CUDA Vec: bring the host mirror / the device array up to date
static PetscErrorCode to_host(Vec v)
{
    if (!v->darray) return PETSC_SUCCESS;
    if (!v->array) v->array = (PetscScalar *)calloc((size_t)(v->n > 0 ? v->n : 1), sizeof(PetscScalar));
    if (!(v->valid & 1)) {
        CUDA_OK(cudaMemcpy(v->array, v->darray, sizeof(PetscScalar) * (size_t)v->n, cudaMemcpyDeviceToHost));
        v->valid |= 1;
    }
    return PETSC_SUCCESS;
}
static PetscErrorCode to_device(Vec v)
{
    if (!v->darray || (v->valid & 2)) return PETSC_SUCCESS;
    CUDA_OK(cudaMemcpy(v->darray, v->array, sizeof(PetscScalar) * (size_t)v->n, cudaMemcpyHostToDevice));
    v->valid |= 2;
    return PETSC_SUCCESS;
}

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

PetscErrorCode ShimWorldSet(int size, int rank, const void *nccl_id128)
{
    if (size < 1 || rank < 0 || rank >= size) return ShimError(PETSC_ERR_ARG_OUTOFRANGE, "ShimWorldSet: bad rank %d of %d", rank, size);
    if (size > 1 && !nccl_id128) return ShimError(PETSC_ERR_ARG_WRONG, "ShimWorldSet: more than one rank needs the NCCL id");
    g_size = size;
    g_rank = rank;
    g_have_id = nccl_id128 != nullptr;
    if (nccl_id128) memcpy(g_nccl_id, nccl_id128, sizeof(g_nccl_id));
    return PETSC_SUCCESS;
}
const void *ShimWorldNcclId(void) { return g_have_id ? g_nccl_id : nullptr; }
PetscErrorCode ShimSetDefaultVecCUDA(int on) { g_default_cuda = on; return PETSC_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { *size = comm == PETSC_COMM_SELF ? 1 : g_size; return 0; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { *rank = comm == PETSC_COMM_SELF ? 0 : g_rank; return 0; }

// ---- PetscObject / PetscContainer ---------------------------------------------------------------------------
PetscErrorCode PetscObjectStateGet(PetscObject obj, PetscObjectState *state) { *state = obj->state; return PETSC_SUCCESS; }
PetscErrorCode PetscObjectGetId(PetscObject obj, PetscObjectId *id) { *id = obj->id; return PETSC_SUCCESS; }
PetscErrorCode PetscObjectCompose(PetscObject obj, const char name[], PetscObject ptr)
{
    if (ptr) ++ptr->refct;                                   // the object it is composed onto holds a reference
    for (ShimComposed *c = obj->composed; c; c = c->next)
        if (!strcmp(c->name, name)) {
            if (c->obj && c->obj->classid == 4) { PetscContainer old = (PetscContainer)c->obj; PetscContainerDestroy(&old); }
            c->obj = ptr;
            return PETSC_SUCCESS;
        }
    ShimComposed *c = (ShimComposed *)calloc(1, sizeof(ShimComposed));
    snprintf(c->name, sizeof(c->name), "%s", name);
    c->obj = ptr;
    c->next = obj->composed;
    obj->composed = c;
    return PETSC_SUCCESS;
}
PetscErrorCode PetscObjectQuery(PetscObject obj, const char name[], PetscObject *ptr)
{
    *ptr = nullptr;
    for (ShimComposed *c = obj->composed; c; c = c->next)
        if (!strcmp(c->name, name)) *ptr = c->obj;
    return PETSC_SUCCESS;
}
PetscErrorCode PetscContainerCreate(MPI_Comm, PetscContainer *container)
{
    PetscContainer c = (PetscContainer)calloc(1, sizeof(_p_PetscContainer));
    hdr_init(&c->hdr, 4);
    *container = c;
    return PETSC_SUCCESS;
}
PetscErrorCode PetscContainerSetPointer(PetscContainer c, void *ptr) { c->ptr = ptr; return PETSC_SUCCESS; }
PetscErrorCode PetscContainerGetPointer(PetscContainer c, void **ptr) { *ptr = c->ptr; return PETSC_SUCCESS; }
PetscErrorCode PetscContainerSetUserDestroy(PetscContainer c, PetscErrorCode (*destroy)(void *)) { c->destroy = destroy; return PETSC_SUCCESS; }
PetscErrorCode PetscContainerDestroy(PetscContainer *c)
{
    if (c && *c) {
        // reference counted, as in PETSc: create / compose / destroy leaves the composed-onto object as the owner
        if (--(*c)->hdr.refct > 0) { *c = nullptr; return PETSC_SUCCESS; }
        if ((*c)->destroy && (*c)->ptr) (*c)->destroy((*c)->ptr);
        free(*c);
        *c = nullptr;
    }
    return PETSC_SUCCESS;
}

// ---- Vec ----------------------------------------------------------------------------------------------------
static PetscErrorCode vec_new(PetscInt n, PetscInt N, PetscInt lo, bool cuda, const PetscScalar *darray, Vec *v)
{
    if (n < 0 || !v) return ShimError(PETSC_ERR_ARG_OUTOFRANGE, "Vec create: bad size");
    Vec w = (Vec)calloc(1, sizeof(_p_Vec));
    hdr_init(&w->hdr, 1);
    w->n = n; w->N = N; w->lo = lo;
    if (cuda) {
        if (darray) {
            w->darray = const_cast<PetscScalar *>(darray);
        } else {
            CUDA_OK(cudaMalloc((void **)&w->darray, sizeof(PetscScalar) * (size_t)(n > 0 ? n : 1)));
            CUDA_OK(cudaMemset(w->darray, 0, sizeof(PetscScalar) * (size_t)(n > 0 ? n : 1)));
            w->own_darray = true;
        }
        w->valid = 2;
    } else {
        w->array = (PetscScalar *)calloc((size_t)(n > 0 ? n : 1), sizeof(PetscScalar));
        w->valid = 1;
    }
    *v = w;
    return PETSC_SUCCESS;
}
static PetscErrorCode mpi_layout(PetscInt n, PetscInt N, PetscInt *lo)
{
    // contiguous ownership in rank order; the shim only knows its own n, so equal shares are assumed (that is what
    // MatCreateVecsFFTW hands out for z-slabs with nz divisible by the ranks)
    if (n * g_size != N) return ShimError(PETSC_ERR_SUP, "shim VecCreateMPI: equal local sizes only (n=%d N=%d ranks=%d)", n, N, g_size);
    *lo = g_rank * n;
    return PETSC_SUCCESS;
}
PetscErrorCode VecCreateSeq(MPI_Comm, PetscInt n, Vec *v) { return vec_new(n, n, 0, false, nullptr, v); }
PetscErrorCode VecCreateMPI(MPI_Comm, PetscInt n, PetscInt N, Vec *v)
{
    PetscInt lo = 0;
    PetscCall(mpi_layout(n, N, &lo));
    return vec_new(n, N, lo, false, nullptr, v);
}
PetscErrorCode VecCreateSeqCUDA(MPI_Comm, PetscInt n, Vec *v) { return vec_new(n, n, 0, true, nullptr, v); }
PetscErrorCode VecCreateSeqCUDAWithArray(MPI_Comm, PetscInt, PetscInt n, const PetscScalar *darray, Vec *v)
{
    return vec_new(n, n, 0, true, darray, v);
}
PetscErrorCode VecCreateMPICUDAWithArray(MPI_Comm, PetscInt, PetscInt n, PetscInt N, const PetscScalar *darray, Vec *v)
{
    PetscInt lo = 0;
    PetscCall(mpi_layout(n, N, &lo));
    return vec_new(n, N, lo, true, darray, v);
}
PetscErrorCode VecDuplicate(Vec v, Vec *w) { return vec_new(v->n, v->N, v->lo, v->darray != nullptr, nullptr, w); }
PetscErrorCode VecDestroy(Vec *v)
{
    if (v && *v) {
        hdr_release(&(*v)->hdr);
        free((*v)->array);
        if ((*v)->own_darray) cudaFree((*v)->darray);
        free(*v);
        *v = nullptr;
    }
    return PETSC_SUCCESS;
}
PetscErrorCode VecGetSize(Vec v, PetscInt *N) { *N = v->N; return PETSC_SUCCESS; }
PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n) { *n = v->n; return PETSC_SUCCESS; }
PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt *lo, PetscInt *hi)
{
    if (lo) *lo = v->lo;
    if (hi) *hi = v->lo + v->n;
    return PETSC_SUCCESS;
}
PetscErrorCode VecGetArray(Vec v, PetscScalar **a)
{
    PetscCall(to_host(v));
    *a = v->array;
    if (v->darray) v->valid = 1;
    ++v->hdr.state;
    return PETSC_SUCCESS;
}
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a) { if (a) *a = nullptr; ++v->hdr.state; return PETSC_SUCCESS; }
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a) { PetscCall(to_host(v)); *a = v->array; return PETSC_SUCCESS; }
PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar **a) { if (a) *a = nullptr; return PETSC_SUCCESS; }
PetscErrorCode VecGetArrayAndMemType(Vec v, PetscScalar **a, PetscMemType *mtype)
{
    if (v->darray) {
        PetscCall(to_device(v));
        *a = v->darray;
        v->valid = 2;
        if (mtype) *mtype = PETSC_MEMTYPE_CUDA;
    } else {
        *a = v->array;
        if (mtype) *mtype = PETSC_MEMTYPE_HOST;
    }
    ++v->hdr.state;
    return PETSC_SUCCESS;
}
PetscErrorCode VecRestoreArrayAndMemType(Vec v, PetscScalar **a) { if (a) *a = nullptr; ++v->hdr.state; return PETSC_SUCCESS; }
PetscErrorCode VecGetArrayReadAndMemType(Vec v, const PetscScalar **a, PetscMemType *mtype)
{
    if (v->darray) {
        PetscCall(to_device(v));
        *a = v->darray;
        if (mtype) *mtype = PETSC_MEMTYPE_CUDA;
    } else {
        *a = v->array;
        if (mtype) *mtype = PETSC_MEMTYPE_HOST;
    }
    return PETSC_SUCCESS;
}
PetscErrorCode VecRestoreArrayReadAndMemType(Vec, const PetscScalar **a) { if (a) *a = nullptr; return PETSC_SUCCESS; }

// the element-wise helpers work on the host array (set-up and test code only; nothing on the hot path calls them)
PetscErrorCode VecSet(Vec v, PetscScalar a)
{
    PetscScalar *p;
    PetscCall(VecGetArray(v, &p));
    for (PetscInt i = 0; i < v->n; ++i) p[i] = a;
    return VecRestoreArray(v, &p);
}
PetscErrorCode VecSetValue(Vec v, PetscInt i, PetscScalar a, InsertMode mode)
{
    if (i < v->lo || i >= v->lo + v->n) return ShimError(PETSC_ERR_ARG_OUTOFRANGE, "VecSetValue: index %d not owned", i);
    PetscScalar *p;
    PetscCall(VecGetArray(v, &p));
    if (mode == ADD_VALUES) p[i - v->lo] += a; else p[i - v->lo] = a;
    return VecRestoreArray(v, &p);
}
PetscErrorCode VecAssemblyBegin(Vec) { return PETSC_SUCCESS; }
PetscErrorCode VecAssemblyEnd(Vec) { return PETSC_SUCCESS; }
PetscErrorCode VecCopy(Vec x, Vec y)
{
    if (x->n != y->n) return ShimError(PETSC_ERR_ARG_WRONG, "VecCopy: size mismatch");
    if (x == y) return PETSC_SUCCESS;
    const PetscScalar *px;
    PetscScalar *py;
    PetscCall(VecGetArrayRead(x, &px));
    PetscCall(VecGetArray(y, &py));
    memcpy(py, px, sizeof(PetscScalar) * (size_t)x->n);
    return VecRestoreArray(y, &py);
}
PetscErrorCode VecScale(Vec v, PetscScalar a)
{
    PetscScalar *p;
    PetscCall(VecGetArray(v, &p));
    for (PetscInt i = 0; i < v->n; ++i) p[i] *= a;
    return VecRestoreArray(v, &p);
}
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x)
{
    if (x->n != y->n) return ShimError(PETSC_ERR_ARG_WRONG, "VecAXPY: size mismatch");
    const PetscScalar *px;
    PetscScalar *py;
    PetscCall(VecGetArrayRead(x, &px));
    PetscCall(VecGetArray(y, &py));
    for (PetscInt i = 0; i < y->n; ++i) py[i] += a * px[i];
    return VecRestoreArray(y, &py);
}
PetscErrorCode VecShift(Vec v, PetscScalar a)
{
    PetscScalar *p;
    PetscCall(VecGetArray(v, &p));
    for (PetscInt i = 0; i < v->n; ++i) p[i] += a;
    return VecRestoreArray(v, &p);
}
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y)
{
    const PetscScalar *px, *py;
    PetscScalar *pw;
    PetscCall(VecGetArrayRead(x, &px));
    PetscCall(VecGetArrayRead(y, &py));
    PetscCall(VecGetArray(w, &pw));
    for (PetscInt i = 0; i < w->n; ++i) pw[i] = px[i] / py[i];
    return VecRestoreArray(w, &pw);
}
PetscErrorCode VecNorm2(Vec v, PetscReal *nrm)
{
    const PetscScalar *p;
    PetscCall(VecGetArrayRead(v, &p));
    long double s = 0;
    for (PetscInt i = 0; i < v->n; ++i) s += std::norm(p[i]);
    *nrm = (PetscReal)sqrtl(s);
    return PETSC_SUCCESS;
}

// ---- Mat ----------------------------------------------------------------------------------------------------
// MatCreateFFT(comm, ndim, dims, MATFFTW, &A): the FFTW plan behind MATFFTW becomes a libcirculantpc plan composed
// onto the Mat (reference call sites: src/PCSHELLFft_3D.cxx:34-35, tests/TransportEquationFFT_...:97-100; dims
// slowest first).  With PETSC_COMM_WORLD and more than one rank the plan is a z-slab plan (fftw-mpi's layout).
PetscErrorCode MatCreateFFT(MPI_Comm comm, PetscInt ndim, const PetscInt dims[], MatType, Mat *A)
{
    if (ndim < 1 || ndim > 3 || !dims || !A) return ShimError(PETSC_ERR_ARG_OUTOFRANGE, "MatCreateFFT: ndim must be 1..3");
    Mat M = (Mat)calloc(1, sizeof(_p_Mat));
    hdr_init(&M->hdr, 2);
    M->kind = SHIM_MAT_FFT;
    M->ndim = ndim;
    M->comm = comm;
    PetscInt n[3] = { 1, 1, 1 };     // nx, ny, nz (x fastest = last entry of dims)
    for (PetscInt d = 0; d < ndim; ++d) { M->dims[d] = dims[d]; n[ndim - 1 - d] = dims[d]; }
    PetscErrorCode ierr = CPCMatAttachPlan(M, comm, n[0], n[1], n[2]);
    if (ierr) { free(M); return ierr; }
    *A = M;
    return PETSC_SUCCESS;
}
PetscErrorCode MatCreateVecsFFTW(Mat A, Vec *x, Vec *y, Vec *z)
{
    if (!A || A->kind != SHIM_MAT_FFT) return ShimError(PETSC_ERR_ARG_WRONG, "MatCreateVecsFFTW: not an FFT matrix");
    PetscInt N = 1;
    for (PetscInt d = 0; d < A->ndim; ++d) N *= A->dims[d];
    int size = 1, rank = 0;
    MPI_Comm_size(A->comm, &size);
    MPI_Comm_rank(A->comm, &rank);
    const PetscInt nloc = N / size;
    Vec *out[3] = { x, y, z };
    for (Vec **o = out; o < out + 3; ++o)
        if (*o) PetscCall(vec_new(nloc, N, rank * nloc, g_default_cuda != 0, nullptr, *o));
    return PETSC_SUCCESS;
}
PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left)
{
    if (A && A->kind == SHIM_MAT_FFT) return MatCreateVecsFFTW(A, right, left, nullptr);
    if (right) PetscCall(VecCreateSeq(PETSC_COMM_SELF, A->cols, right));
    if (left) PetscCall(VecCreateSeq(PETSC_COMM_SELF, A->rows, left));
    return PETSC_SUCCESS;
}
PetscErrorCode MatCreateSeqAIJFromCSR(PetscInt rows, PetscInt cols, const PetscInt *rowptr, const PetscInt *colidx,
                                      const PetscScalar *val, Mat *A)
{
    Mat M = (Mat)calloc(1, sizeof(_p_Mat));
    hdr_init(&M->hdr, 2);
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
PetscErrorCode MatGetSize(Mat A, PetscInt *rows, PetscInt *cols)
{
    PetscInt N = 1;
    if (A->kind == SHIM_MAT_FFT)
        for (PetscInt d = 0; d < A->ndim; ++d) N *= A->dims[d];
    if (rows) *rows = A->kind == SHIM_MAT_FFT ? N : A->rows;
    if (cols) *cols = A->kind == SHIM_MAT_FFT ? N : A->cols;
    return PETSC_SUCCESS;
}
PetscErrorCode MatGetRowIJ(Mat A, PetscInt shift, PetscBool, PetscBool, PetscInt *n, const PetscInt *ia[],
                           const PetscInt *ja[], PetscBool *done)
{
    if (A->kind != SHIM_MAT_CSR || shift != 0) { if (done) *done = PETSC_FALSE; return PETSC_SUCCESS; }
    *n = A->rows; *ia = A->rowptr; *ja = A->colidx;
    if (done) *done = PETSC_TRUE;
    return PETSC_SUCCESS;
}
PetscErrorCode MatRestoreRowIJ(Mat, PetscInt, PetscBool, PetscBool, PetscInt *, const PetscInt *ia[], const PetscInt *ja[],
                               PetscBool *)
{
    if (ia) *ia = nullptr;
    if (ja) *ja = nullptr;
    return PETSC_SUCCESS;
}
PetscErrorCode MatSeqAIJGetArrayRead(Mat A, const PetscScalar **array)
{
    if (A->kind != SHIM_MAT_CSR) return ShimError(PETSC_ERR_ARG_WRONG, "MatSeqAIJGetArrayRead: not an AIJ matrix");
    *array = A->val;
    return PETSC_SUCCESS;
}
PetscErrorCode MatSeqAIJRestoreArrayRead(Mat, const PetscScalar **array) { if (array) *array = nullptr; return PETSC_SUCCESS; }
PetscErrorCode MatDestroy(Mat *A)
{
    if (A && *A) {
        hdr_release(&(*A)->hdr);          // destroys the composed plan container (cpc_destroy)
        free((*A)->rowptr); free((*A)->colidx); free((*A)->val);
        free(*A);
        *A = nullptr;
    }
    return PETSC_SUCCESS;
}

static PetscErrorCode fft_mult(Mat A, Vec x, Vec y, int dir)
{
    cpc_plan plan = nullptr;
    PetscCall(CPCMatGetPlan(A, &plan));
    const PetscScalar *px;
    PetscScalar *py;
    PetscMemType mx, my;
    PetscCall(VecGetArrayReadAndMemType(x, &px, &mx));
    PetscCall(VecGetArrayAndMemType(y, &py, &my));
    if (PetscMemTypeDevice(mx) != PetscMemTypeDevice(my)) return ShimError(PETSC_ERR_SUP, "MatMult(FFT): x and y must live in the same memory");
    const int kind = PetscMemTypeDevice(mx) ? CPC_MEM_DEVICE : CPC_MEM_HOST;
    const int st = dir < 0 ? cpc_forward(plan, px, py, kind) : cpc_inverse(plan, px, py, kind);
    if (!st && kind == CPC_MEM_DEVICE) cpc_sync(plan);
    PetscCall(VecRestoreArrayAndMemType(y, &py));
    PetscCall(VecRestoreArrayReadAndMemType(x, &px));
    if (st) return ShimError(PETSC_ERR_LIB, "libcirculantpc: %s", cpc_last_error());
    return PETSC_SUCCESS;
}
PetscErrorCode MatMult(Mat A, Vec x, Vec y)
{
    if (A->kind == SHIM_MAT_FFT) return fft_mult(A, x, y, -1);     // unnormalised forward DFT (FftLinearSolver_3D.c:170)
    if (x->n != A->cols || y->n != A->rows) return ShimError(PETSC_ERR_ARG_WRONG, "MatMult: size mismatch");
    const PetscScalar *px;
    PetscScalar *py;
    PetscCall(VecGetArrayRead(x, &px));
    PetscCall(VecGetArray(y, &py));
    for (PetscInt i = 0; i < A->rows; ++i) {
        PetscScalar s = 0;
        for (PetscInt p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) s += A->val[p] * px[A->colidx[p]];
        py[i] = s;
    }
    return VecRestoreArray(y, &py);
}
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y)
{
    if (A->kind == SHIM_MAT_FFT) return fft_mult(A, x, y, +1);     // unnormalised backward DFT (FftLinearSolver_3D.c:180)
    if (x->n != A->rows || y->n != A->cols) return ShimError(PETSC_ERR_ARG_WRONG, "MatMultTranspose: size mismatch");
    const PetscScalar *px;
    PetscScalar *py;
    PetscCall(VecGetArrayRead(x, &px));
    PetscCall(VecGetArray(y, &py));
    for (PetscInt i = 0; i < A->cols; ++i) py[i] = 0;
    for (PetscInt i = 0; i < A->rows; ++i)
        for (PetscInt p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) py[A->colidx[p]] += A->val[p] * px[i];
    return VecRestoreArray(y, &py);
}

// ---- PC -----------------------------------------------------------------------------------------------------
PetscErrorCode PCCreate(MPI_Comm, PC *pc)
{
    *pc = (PC)calloc(1, sizeof(_p_PC));
    hdr_init(&(*pc)->hdr, 3);
    return PETSC_SUCCESS;
}
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
        hdr_release(&(*pc)->hdr);
        free(*pc);
        *pc = nullptr;
        return ierr;
    }
    return PETSC_SUCCESS;
}

}  // extern "C"
