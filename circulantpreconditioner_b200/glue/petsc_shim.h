/* petsc_shim.h -- the PETSc names the reference's hot path touches, implemented over plain host / CUDA arrays.
 *
 * PETSc is not installable in this image (no network), so the reference-named glue (circulantpc_petsc.cxx,
 * circulantpc_pcshell.cxx) is compiled and exercised against this stand-in.  The glue only ever uses PUBLIC PETSc
 * functions (it never looks inside a Vec, Mat or PC): the same two sources also compile with -DCPC_WITH_PETSC against
 * <petscksp.h> -- which `make check-petsc-clean` proves against petsc_opaque_stub.h, a header that declares the PETSc
 * types as opaque pointers with only the public prototypes the glue calls.  The structs below are transparent only
 * for petsc_shim.cxx itself and for the C++ test driver.
 *
 * Semantics follow a complex-scalar PETSc build (PetscScalar = complex128), the branch of the reference whose
 * semantics are well defined (reference src/FftLinearSolver_3D.c:173-175,183-184).
 *
 * Vec flavours: VecCreateSeq / VecCreateMPI (host array, like VECSEQ / VECMPI) and VecCreateSeqCUDA /
 * VecCreateSeqCUDAWithArray / VecCreateMPICUDAWithArray (device array, like VECSEQCUDA / VECMPICUDA):
 * Vec{Get,Restore}Array[Read]AndMemType hand back the device pointer of a CUDA Vec without a copy, the plain
 * VecGetArray family hands back a host mirror (copied on demand), exactly as PETSc's offload mask does.
 * "MPI": ShimWorldSet(size, rank, nccl_id) describes the communicator PETSC_COMM_WORLD stands for (one process per
 * GPU, launched by the test harness); MPI_Comm_size / MPI_Comm_rank report it.
 */
#ifndef CPC_PETSC_SHIM_H
#define CPC_PETSC_SHIM_H

#include <complex>
#include <cstddef>
#include <cstdint>

#include "../../include/circulantpc.h"

typedef int PetscInt;
typedef int PetscErrorCode;
typedef int PetscMPIInt;
typedef double PetscReal;
typedef std::complex<double> PetscScalar;
typedef bool PetscBool;
typedef int MPI_Comm;
typedef const char *MatType;
typedef int64_t PetscObjectState;
typedef int64_t PetscObjectId;

#define PETSC_USE_COMPLEX 1
#define PETSC_SUCCESS 0
#define PETSC_TRUE true
#define PETSC_FALSE false
#define PETSC_ERR_SUP 56
#define PETSC_ERR_ORDER 58
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_ARG_OUTOFRANGE 63
#define PETSC_ERR_LIB 76
#define PETSC_COMM_WORLD 0
#define PETSC_COMM_SELF 1
#define PETSC_DECIDE (-1)
#define MATFFTW "fftw"
enum InsertMode { INSERT_VALUES = 1, ADD_VALUES = 2 };
typedef enum { PETSC_MEMTYPE_HOST = 0, PETSC_MEMTYPE_DEVICE = 1, PETSC_MEMTYPE_CUDA = 1 } PetscMemType;
#define PetscMemTypeDevice(m) (((m) & 0x1) != 0)
#define PetscMemTypeHost(m) (((m) & 0x1) == 0)
#define PetscRealPart(a) (std::real(a))
#define PetscImaginaryPart(a) (std::imag(a))

/* every shim object starts with this header (PetscObject in PETSc) */
struct _p_PetscObject {
    int classid;                /* 1 Vec, 2 Mat, 3 PC, 4 PetscContainer */
    PetscObjectId id;
    PetscObjectState state;     /* bumped by every write access */
    int refct;                  /* references held (the creator's, plus one per object it is composed onto) */
    struct ShimComposed *composed;
};
typedef struct _p_PetscObject *PetscObject;

struct _p_PetscContainer {
    _p_PetscObject hdr;
    void *ptr;
    PetscErrorCode (*destroy)(void *);
};
typedef struct _p_PetscContainer *PetscContainer;

struct _p_Vec {
    _p_PetscObject hdr;
    PetscInt n;                 /* local entries */
    PetscInt N;                 /* global entries */
    PetscInt lo;                /* first global index owned */
    PetscScalar *array;         /* host array (the mirror of a CUDA Vec, allocated on demand) */
    PetscScalar *darray;        /* device array of a CUDA Vec, else NULL */
    bool own_darray;
    int valid;                  /* CUDA Vec offload mask: 1 host mirror valid, 2 device valid, 3 both */
};
typedef struct _p_Vec *Vec;

enum ShimMatKind { SHIM_MAT_FFT = 1, SHIM_MAT_CSR = 2 };
struct _p_Mat {
    _p_PetscObject hdr;
    int kind;
    /* SHIM_MAT_FFT: created by MatCreateFFT; the B200 plan that replaces the FFTW plan rides on the object as a
       composed container (CPCMatAttachPlan, circulantpc_petsc.cxx) */
    PetscInt ndim;
    PetscInt dims[3];           /* as given to MatCreateFFT: slowest first ({nz, ny, nx}) */
    MPI_Comm comm;
    /* SHIM_MAT_CSR: projection matrix (intersectionMatrix) */
    PetscInt rows, cols;
    PetscInt *rowptr, *colidx;
    PetscScalar *val;
};
typedef struct _p_Mat *Mat;

struct _p_PC {
    _p_PetscObject hdr;
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
/* what PETSC_COMM_WORLD stands for: size ranks, this process is `rank`, nccl_id = the 128 bytes every rank shares
   (cpc_nccl_unique_id on rank 0, broadcast by the launcher); size 1 needs no id */
PetscErrorCode ShimWorldSet(int size, int rank, const void *nccl_id128);
const void *ShimWorldNcclId(void);
/* MatCreateVecsFFTW makes CUDA Vecs when set (the stand-in for -vec_type cuda) */
PetscErrorCode ShimSetDefaultVecCUDA(int on);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);

PetscErrorCode PetscObjectStateGet(PetscObject obj, PetscObjectState *state);
PetscErrorCode PetscObjectGetId(PetscObject obj, PetscObjectId *id);
PetscErrorCode PetscObjectCompose(PetscObject obj, const char name[], PetscObject ptr);
PetscErrorCode PetscObjectQuery(PetscObject obj, const char name[], PetscObject *ptr);
PetscErrorCode PetscContainerCreate(MPI_Comm comm, PetscContainer *container);
PetscErrorCode PetscContainerSetPointer(PetscContainer container, void *ptr);
PetscErrorCode PetscContainerGetPointer(PetscContainer container, void **ptr);
PetscErrorCode PetscContainerSetUserDestroy(PetscContainer container, PetscErrorCode (*destroy)(void *));
PetscErrorCode PetscContainerDestroy(PetscContainer *container);

PetscErrorCode VecCreateSeq(MPI_Comm, PetscInt n, Vec *v);
PetscErrorCode VecCreateMPI(MPI_Comm, PetscInt n, PetscInt N, Vec *v);
PetscErrorCode VecCreateSeqCUDA(MPI_Comm, PetscInt n, Vec *v);
PetscErrorCode VecCreateSeqCUDAWithArray(MPI_Comm, PetscInt bs, PetscInt n, const PetscScalar *darray, Vec *v);
PetscErrorCode VecCreateMPICUDAWithArray(MPI_Comm, PetscInt bs, PetscInt n, PetscInt N, const PetscScalar *darray, Vec *v);
PetscErrorCode VecDuplicate(Vec v, Vec *w);
PetscErrorCode VecDestroy(Vec *v);
PetscErrorCode VecGetSize(Vec v, PetscInt *N);
PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n);
PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt *lo, PetscInt *hi);
PetscErrorCode VecSet(Vec v, PetscScalar a);
PetscErrorCode VecSetValue(Vec v, PetscInt i, PetscScalar a, InsertMode mode);
PetscErrorCode VecAssemblyBegin(Vec v);
PetscErrorCode VecAssemblyEnd(Vec v);
PetscErrorCode VecGetArray(Vec v, PetscScalar **a);
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a);
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a);
PetscErrorCode VecRestoreArrayRead(Vec v, const PetscScalar **a);
PetscErrorCode VecGetArrayAndMemType(Vec v, PetscScalar **a, PetscMemType *mtype);
PetscErrorCode VecRestoreArrayAndMemType(Vec v, PetscScalar **a);
PetscErrorCode VecGetArrayReadAndMemType(Vec v, const PetscScalar **a, PetscMemType *mtype);
PetscErrorCode VecRestoreArrayReadAndMemType(Vec v, const PetscScalar **a);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecScale(Vec v, PetscScalar a);
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x);
PetscErrorCode VecShift(Vec v, PetscScalar a);
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y);
PetscErrorCode VecNorm2(Vec v, PetscReal *nrm);

PetscErrorCode MatCreateFFT(MPI_Comm, PetscInt ndim, const PetscInt dims[], MatType, Mat *A);
PetscErrorCode MatCreateVecsFFTW(Mat A, Vec *x, Vec *y, Vec *z);
PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left);
PetscErrorCode MatCreateSeqAIJFromCSR(PetscInt rows, PetscInt cols, const PetscInt *rowptr, const PetscInt *colidx,
                                      const PetscScalar *val, Mat *A);
PetscErrorCode MatGetSize(Mat A, PetscInt *rows, PetscInt *cols);
PetscErrorCode MatGetRowIJ(Mat A, PetscInt shift, PetscBool symmetric, PetscBool inodecompressed, PetscInt *n,
                           const PetscInt *ia[], const PetscInt *ja[], PetscBool *done);
PetscErrorCode MatRestoreRowIJ(Mat A, PetscInt shift, PetscBool symmetric, PetscBool inodecompressed, PetscInt *n,
                               const PetscInt *ia[], const PetscInt *ja[], PetscBool *done);
PetscErrorCode MatSeqAIJGetArrayRead(Mat A, const PetscScalar **array);
PetscErrorCode MatSeqAIJRestoreArrayRead(Mat A, const PetscScalar **array);
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

#endif
