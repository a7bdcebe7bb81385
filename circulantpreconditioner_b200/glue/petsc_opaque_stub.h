/* petsc_opaque_stub.h -- COMPILE CHECK ONLY (make check-petsc-clean): the PETSc types as the opaque pointers they are
 * in <petscksp.h>, plus exactly the public PETSc prototypes and macros the glue sources use.  If circulantpc_petsc.cxx
 * or circulantpc_pcshell.cxx ever touched the inside of a Vec / Mat / PC (as round 1's did through the shim's
 * structs) they would not compile against this header.  Signatures follow PETSc 3.19-3.22 (complex scalars, 32-bit
 * indices, C++): petscsys.h, petscvec.h, petscmat.h, petscpc.h.  Nothing here is linked or run.
 */
#ifndef CPC_PETSC_OPAQUE_STUB_H
#define CPC_PETSC_OPAQUE_STUB_H

#include <complex>
#include <cstdint>

typedef int PetscInt;
typedef int PetscErrorCode;
typedef int PetscMPIInt;
typedef double PetscReal;
typedef std::complex<double> PetscScalar;      /* PetscComplex of a C++ complex build */
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef int64_t PetscObjectState;
typedef int64_t PetscObjectId;
typedef const char *MatType;

typedef struct ompi_communicator_t *MPI_Comm;  /* whatever mpi.h says; opaque here */
typedef struct ompi_datatype_t *MPI_Datatype;
extern MPI_Comm PETSC_COMM_WORLD, PETSC_COMM_SELF;
extern MPI_Datatype MPI_BYTE;
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Bcast(void *buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm);

typedef struct _p_PetscObject *PetscObject;
typedef struct _p_PetscContainer *PetscContainer;
typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;
typedef struct _p_PC *PC;
typedef struct _p_KSP *KSP;
typedef const char *KSPType;
typedef const char *PCType;
#define KSPGMRES "gmres"
#define PCSHELL "shell"
#define PETSC_DEFAULT (-2)

#define PETSC_USE_COMPLEX 1
#define PETSC_SUCCESS 0
#define PETSC_ERR_SUP 56
#define PETSC_ERR_ORDER 58
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_ARG_OUTOFRANGE 63
#define PETSC_ERR_LIB 76
#define MATFFTW "fftw"
typedef enum { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES } InsertMode;
typedef enum { PETSC_MEMTYPE_HOST = 0, PETSC_MEMTYPE_DEVICE = 0x01, PETSC_MEMTYPE_CUDA = 0x01 } PetscMemType;
#define PetscMemTypeDevice(m) (((m) & 0x1) == 0x1)
#define PetscRealPart(a) (std::real(a))
#define PetscImaginaryPart(a) (std::imag(a))
typedef enum { MATOP_MULT = 3, MATOP_MULT_TRANSPOSE = 5 } MatOperation;

PetscErrorCode PetscError(MPI_Comm, int, const char *, const char *, PetscErrorCode, int, const char *, ...);
#define PetscFunctionBeginUser do { } while (0)
#define PetscFunctionReturn(x) return (x)
#define PetscCall(...)                                 \
    do {                                               \
        PetscErrorCode ierr_petsc_call_ = __VA_ARGS__; \
        if (ierr_petsc_call_) return ierr_petsc_call_; \
    } while (0)
#define PetscCheck(cond, comm, ierr, ...)                                                               \
    do {                                                                                                \
        if (!(cond)) return PetscError(comm, __LINE__, __func__, __FILE__, ierr, 0, __VA_ARGS__);       \
    } while (0)

PetscErrorCode PetscObjectStateGet(PetscObject, PetscObjectState *);
PetscErrorCode PetscObjectGetId(PetscObject, PetscObjectId *);
PetscErrorCode PetscObjectCompose(PetscObject, const char[], PetscObject);
PetscErrorCode PetscObjectQuery(PetscObject, const char[], PetscObject *);
PetscErrorCode PetscContainerCreate(MPI_Comm, PetscContainer *);
PetscErrorCode PetscContainerSetPointer(PetscContainer, void *);
PetscErrorCode PetscContainerGetPointer(PetscContainer, void **);
PetscErrorCode PetscContainerSetUserDestroy(PetscContainer, PetscErrorCode (*)(void *));
PetscErrorCode PetscContainerDestroy(PetscContainer *);

PetscErrorCode VecDestroy(Vec *);
PetscErrorCode VecGetSize(Vec, PetscInt *);
PetscErrorCode VecGetLocalSize(Vec, PetscInt *);
PetscErrorCode VecGetOwnershipRange(Vec, PetscInt *, PetscInt *);
PetscErrorCode VecSet(Vec, PetscScalar);
PetscErrorCode VecSetValue(Vec, PetscInt, PetscScalar, InsertMode);
PetscErrorCode VecAssemblyBegin(Vec);
PetscErrorCode VecAssemblyEnd(Vec);
PetscErrorCode VecGetArray(Vec, PetscScalar **);
PetscErrorCode VecRestoreArray(Vec, PetscScalar **);
PetscErrorCode VecGetArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecGetArrayAndMemType(Vec, PetscScalar **, PetscMemType *);
PetscErrorCode VecRestoreArrayAndMemType(Vec, PetscScalar **);
PetscErrorCode VecGetArrayReadAndMemType(Vec, const PetscScalar **, PetscMemType *);
PetscErrorCode VecRestoreArrayReadAndMemType(Vec, const PetscScalar **);

PetscErrorCode MatCreateShell(MPI_Comm, PetscInt, PetscInt, PetscInt, PetscInt, void *, Mat *);
PetscErrorCode MatShellSetOperation(Mat, MatOperation, void (*)(void));
PetscErrorCode MatCreateVecs(Mat, Vec *, Vec *);
PetscErrorCode MatGetSize(Mat, PetscInt *, PetscInt *);
PetscErrorCode MatGetRowIJ(Mat, PetscInt, PetscBool, PetscBool, PetscInt *, const PetscInt *[], const PetscInt *[], PetscBool *);
PetscErrorCode MatRestoreRowIJ(Mat, PetscInt, PetscBool, PetscBool, PetscInt *, const PetscInt *[], const PetscInt *[], PetscBool *);
PetscErrorCode MatSeqAIJGetArrayRead(Mat, const PetscScalar **);
PetscErrorCode MatSeqAIJRestoreArrayRead(Mat, const PetscScalar **);
PetscErrorCode MatDestroy(Mat *);
PetscErrorCode MatMult(Mat, Vec, Vec);

PetscErrorCode KSPCreate(MPI_Comm, KSP *);
PetscErrorCode KSPSetType(KSP, KSPType);
PetscErrorCode KSPSetTolerances(KSP, PetscReal, PetscReal, PetscReal, PetscInt);
PetscErrorCode KSPGetPC(KSP, PC *);
PetscErrorCode KSPSetOperators(KSP, Mat, Mat);
PetscErrorCode KSPSolve(KSP, Vec, Vec);
PetscErrorCode KSPDestroy(KSP *);
PetscErrorCode PCSetType(PC, PCType);
PetscErrorCode PCShellSetContext(PC, void *);
PetscErrorCode PCShellGetContext(PC, void *);
PetscErrorCode PCShellSetApply(PC, PetscErrorCode (*)(PC, Vec, Vec));
PetscErrorCode PCShellSetSetUp(PC, PetscErrorCode (*)(PC));
PetscErrorCode PCShellSetDestroy(PC, PetscErrorCode (*)(PC));

#endif
