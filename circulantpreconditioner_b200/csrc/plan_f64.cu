// plan_f64.cu -- complex128 instantiation of the plan and its kernels (PetscScalar of a complex PETSc build).
#include "plan_impl.cuh"

namespace cpc {

__global__ void diag_from_separable_kernel(double2 *__restrict__ diag, const double2 *__restrict__ ax,
                                           const double2 *__restrict__ ay, const double2 *__restrict__ az, int nx,
                                           int ny, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % nx);
        const long long r = i / nx;
        const int y = (int)(r % ny);
        const int z = (int)(r / ny);
        // same summation order as build_diag_mat_vec_3D (reference FftLinearSolver_3D.c:151-155):
        // ((kpi_x + kpi_y) + kpi_z) + 1 ; the "+1" is carried by the y table, which is exact for the real part
        diag[i] = make_double2(ax[x].x + ay[y].x + az[z].x, ax[x].y + ay[y].y + az[z].y);
    }
}

PlanBase *make_plan_f64() { return new PlanT<double>(); }

}  // namespace cpc
