/* ref_assembly.cxx -- TEST INFRASTRUCTURE ONLY.  Plain-pointer entry points into the reference's own
 * computeDivergenceMatrix functions (src/TransportEquation.cxx:75-133, src/WaveSystem.cxx:109-176), compiled unmodified
 * against solverlab_standin/ (see its header).  The mesh arrives as arrays, the matrix leaves as a dense array; the
 * MatShift(A, 1) of the reference's drivers (tests/TransportEquation_SphericalExplosion_impl_mpi.cxx:117,
 * tests/WaveSystem_SphericalExplosion_impl_mpi.cxx:127) is applied when `shift` is set.  Built into
 * _ref/libreference_assembly.so by the Makefile in this directory; loaded by oracle/ref_assembly.py for tests/ only. */
#include <cstring>
#include <memory>

#include <petscksp.h>

#include "Mesh.hxx"

#define REF_API extern "C" __attribute__((visibility("default")))

/* the reference's functions (declared in its headers TransportEquation2.hxx / WaveSystem.hxx) and the sound speed its
 * header defines as a global */
void initial_conditions_shock(Mesh my_mesh, Field &Temperature_field);
void initial_conditions_shock(Mesh my_mesh, Field &pressure_field, Field &velocity_field);
void computeDivergenceMatrix(Mesh my_mesh, Mat *implMat, double dt, Vector vitesseTransport);
void computeDivergenceMatrix(Mesh my_mesh, Mat *implMat, double dt);
Matrix jacobianMatrices(Vector normal, double coeff);
extern double c0;

static const char *group_name(int code)
{
    switch (code) {
    case 1: return "Neumann";
    case 2: return "Periodic";
    case 3: return "Wall";
    default: return "";
    }
}

static Mesh make_mesh(int dim, int ncells, int nfaces, const int *cell_face_ptr, const int *cell_face_idx,
                      const double *cell_face_normal, const double *cell_measure, const double *cell_centre,
                      const double *face_measure, const int *face_cells, const int *face_group, const int *face_twin,
                      const double *bbox = nullptr)
{
    auto d = std::make_shared<StandinMeshData>();
    d->dim = dim;
    d->cell_faces.resize((size_t)ncells);
    d->cell_normals.resize((size_t)ncells);
    for (int j = 0; j < ncells; ++j) {
        for (int p = cell_face_ptr[j]; p < cell_face_ptr[j + 1]; ++p) {
            d->cell_faces[(size_t)j].push_back(cell_face_idx[p]);
            for (int i = 0; i < dim; ++i) d->cell_normals[(size_t)j].push_back(cell_face_normal[(size_t)p * dim + i]);
        }
    }
    d->cell_measure.assign(cell_measure, cell_measure + ncells);
    d->cell_centre.assign(cell_centre, cell_centre + 3 * (size_t)ncells);
    d->face_measure.assign(face_measure, face_measure + nfaces);
    d->face_cells.assign(face_cells, face_cells + 2 * (size_t)nfaces);
    d->face_twin.assign(face_twin, face_twin + nfaces);
    for (int f = 0; f < nfaces; ++f) d->face_group.push_back(group_name(face_group[f]));
    for (int a = 0; a < 3; ++a) { d->lo[a] = bbox ? bbox[a] : 1e300; d->hi[a] = bbox ? bbox[3 + a] : -1e300; }
    return Mesh(d);
}

static int finish(_p_Mat &A, int shift, double *dense)
{
    for (PetscInt i = 0; i < A.rows; ++i)
        for (PetscInt j = 0; j < A.cols; ++j) {
            const PetscScalar e = A.v[(size_t)i * A.cols + j];
            if (e.imag() != 0.0) return 2;
            dense[(size_t)i * A.cols + j] = e.real() + ((shift && i == j) ? 1.0 : 0.0);
        }
    return 0;
}

/* kind 0: TransportEquation.cxx (velocity a[dim]); 1: WaveSystem.cxx (sound speed c0_value, 4 or dim+1 unknowns per cell).
 * dense: (ncells * ncomp)^2 doubles, row major.  Returns 0; 1 if the reference threw; 2 on a complex entry. */
REF_API int ref_assemble(int kind, int dim, int ncells, int nfaces, const int *cell_face_ptr, const int *cell_face_idx,
                         const double *cell_face_normal, const double *cell_measure, const double *cell_centre,
                         const double *face_measure, const int *face_cells, const int *face_group, const int *face_twin,
                         double dt, const double *a, double c0_value, int shift, double *dense)
{
    try {
        Mesh m = make_mesh(dim, ncells, nfaces, cell_face_ptr, cell_face_idx, cell_face_normal, cell_measure, cell_centre,
                           face_measure, face_cells, face_group, face_twin);
        const int ncomp = kind == 0 ? 1 : dim + 1;
        _p_Mat A;
        A.rows = A.cols = ncells * ncomp;
        A.v.assign((size_t)A.rows * A.cols, PetscScalar(0.0, 0.0));
        Mat pA = &A;
        if (kind == 0) {
            Vector v(dim);
            for (int i = 0; i < dim; ++i) v[i] = a[i];
            computeDivergenceMatrix(m, &pA, dt, v);
        } else {
            c0 = c0_value;
            computeDivergenceMatrix(m, &pA, dt);
        }
        return finish(A, shift, dense);
    } catch (const std::exception &) {
        return 1;
    }
}

/* jacobianMatrices (src/WaveSystem.cxx:92-107): out = (dim+1)^2 doubles */
REF_API int ref_jacobian_minus(int dim, const double *normal, double coeff, double c0_value, double *out)
{
    c0 = c0_value;
    Vector n(dim);
    for (int i = 0; i < dim; ++i) n[i] = normal[i];
    const Matrix M = jacobianMatrices(n, coeff);
    for (int i = 0; i <= dim; ++i)
        for (int j = 0; j <= dim; ++j) out[i * (dim + 1) + j] = M(i, j);
    return 0;
}

/* initial_conditions_shock of both files (src/TransportEquation.cxx:25-73, src/WaveSystem.cxx:25-82): the spherical step
 * around the centre of the bounding box (bbox = xmin, ymin, zmin, xmax, ymax, zmax).  kind 0: out[ncells] = temperature;
 * kind 1: out[ncells * (1 + dim)] = pressure, then the velocity components, per cell. */
REF_API int ref_initial_conditions(int kind, int dim, int ncells, const double *cell_centre, const double *bbox, double *out)
{
    try {
        auto d = std::make_shared<StandinMeshData>();
        d->dim = dim;
        d->cell_faces.resize((size_t)ncells);
        d->cell_normals.resize((size_t)ncells);
        d->cell_measure.assign((size_t)ncells, 1.0);
        d->cell_centre.assign(cell_centre, cell_centre + 3 * (size_t)ncells);
        for (int a = 0; a < 3; ++a) { d->lo[a] = bbox[a]; d->hi[a] = bbox[3 + a]; }
        Mesh m(d);
        if (kind == 0) {
            Field T(ncells);
            initial_conditions_shock(m, T);
            for (int j = 0; j < ncells; ++j) out[j] = T(j);
        } else {
            /* (fields start from zero, as SOLVERLAB's do; the reference's `velocity_field[j,idim]=0` is a comma expression
             * and only ever touches the first dim entries, src/WaveSystem.cxx:69) */
            Field p(ncells), v(ncells, dim);
            initial_conditions_shock(m, p, v);
            for (int j = 0; j < ncells; ++j) {
                out[(size_t)j * (dim + 1)] = p(j);
                for (int c = 0; c < dim; ++c) out[(size_t)j * (dim + 1) + 1 + c] = v(j, c);
            }
        }
        return 0;
    } catch (const std::exception &) {
        return 1;
    }
}
