/* petscksp.h -- TEST INFRASTRUCTURE ONLY: the C++ flavour of the PETSc stand-in, for the reference's
 * src/TransportEquation.cxx and src/WaveSystem.cxx, which touch PETSc through Mat, PetscScalar, MatSetValue(s) and
 * ADD_VALUES only.  PetscScalar is std::complex<double>, as in a complex C++ build of PETSc (and <complex.h>'s macro `I`
 * stays away from WaveSystem.cxx's `int I, J;`).  The matrix is dense: see ref_assembly.cxx. */
#ifndef PETSCKSP_STANDIN_CXX_H
#define PETSCKSP_STANDIN_CXX_H
#include <complex>
#include <vector>
typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscReal;
typedef std::complex<double> PetscScalar;
enum InsertMode { NOT_SET_VALUES = 0, INSERT_VALUES = 1, ADD_VALUES = 2 };
struct _p_Mat { PetscInt rows, cols; std::vector<PetscScalar> v; };
typedef _p_Mat *Mat;
inline PetscErrorCode MatSetValues(Mat A, PetscInt m, const PetscInt *im, PetscInt n, const PetscInt *in, const PetscScalar *v,
                                   InsertMode mode)
{
    for (PetscInt a = 0; a < m; ++a)
        for (PetscInt b = 0; b < n; ++b) {
            if (im[a] < 0 || im[a] >= A->rows || in[b] < 0 || in[b] >= A->cols) return 63;
            PetscScalar &e = A->v[(size_t)im[a] * A->cols + in[b]];
            if (mode == ADD_VALUES) e += v[(size_t)a * n + b]; else e = v[(size_t)a * n + b];
        }
    return 0;
}
inline PetscErrorCode MatSetValue(Mat A, PetscInt i, PetscInt j, PetscScalar v, InsertMode mode)
{
    return MatSetValues(A, 1, &i, 1, &j, &v, mode);
}
#endif
