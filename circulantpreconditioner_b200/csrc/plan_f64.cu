// plan_f64.cu -- complex128 instantiation of the plan and its kernels (PetscScalar of a complex PETSc build).
#include "pencil_impl.cuh"

#include <vector>

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

__global__ void diag_check_separable_kernel(const double2 *__restrict__ diag, const double2 *__restrict__ ax,
                                            const double2 *__restrict__ ay, const double2 *__restrict__ az, int nx, int ny,
                                            long long n, unsigned long long *out)
{
    double md = 0.0, ma = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % nx);
        const long long r = i / nx;
        const int y = (int)(r % ny);
        const long long z = r / ny;
        const double2 d = diag[i];
        const double er = fabs(d.x - (ax[x].x + ay[y].x + az[z].x)), ei = fabs(d.y - (ax[x].y + ay[y].y + az[z].y));
        md = fmax(md, fmax(er, ei));
        ma = fmax(ma, fmax(fabs(d.x), fabs(d.y)));
        if (!(er == er) || !(ei == ei)) md = 1e300;                   // NaN entries are never "separable"
    }
    for (int o = 16; o > 0; o >>= 1) {
        md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o));
        ma = fmax(ma, __shfl_xor_sync(0xffffffffu, ma, o));
    }
    if ((threadIdx.x & 31) == 0) {                                    // non-negative doubles order like their bit patterns
        atomicMax(out, (unsigned long long)__double_as_longlong(md));
        atomicMax(out + 1, (unsigned long long)__double_as_longlong(ma));
    }
}

__global__ void __launch_bounds__(256)
zs_carry_owner_kernel(const double2 *__restrict__ gbuf, long long gstride, long long line0, long long count, int nx,
                      int nzl, int nranks, int rank, int self_only, ZCarryPeers zpeer, const ZSolveArgs a,
                      const FlagSync wait, const FlagSync done)
{
    zs_wait_for_peers(wait);                       // every rank's end values of the lines owned here have landed
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < count;
         j += (long long)gridDim.x * blockDim.x) {
        const long long line = line0 + j;
        const int lx = (int)(line % nx);
        if (lx < a.xs0 || lx >= a.xs0 + a.xsn) continue;          // this launch serves the other half of the columns
        double2 r, c;
        zs_coeffs(a, lx, (int)(line / nx), r, c);
        const double2 cL = cpow_rt(c, nzl);
        double2 e[CPC_MAX_PEERS];
#pragma unroll
        for (int s = 0; s < CPC_MAX_PEERS; ++s)
            e[s] = s < nranks ? gbuf[(long long)s * gstride + j] : make_double2(0.0, 0.0);
        // Zin_0 = sum_m cL^m e_{P-1-m} / (1 - cL^P): Horner over e_0, e_1, ..., e_{P-1}
        double2 acc = make_double2(0.0, 0.0), cLp = make_double2(1.0, 0.0);
#pragma unroll
        for (int s = 0; s < CPC_MAX_PEERS; ++s)
            if (s < nranks) {
                acc = cadd(cmul(cL, acc), e[s]);
                cLp = cmul(cLp, cL);
            }
        double2 Z = cmul(acc, crecip_scaled<double>(make_double2(1.0 - cLp.x, -cLp.y), 1.0));
#pragma unroll
        for (int q = 0; q < CPC_MAX_PEERS; ++q)
            if (q < nranks) {
                if (!self_only || q == rank) zpeer.p[self_only ? 0 : q][line] = Z;
                Z = cadd(e[q], cmul(cL, Z));                        // Zin_{q+1} = e_q + cL Zin_q
            }
    }
    zs_signal_when_grid_done(done);                // "the carry-ins computed here have landed on every rank"
}

int build_diag_separable(int nx, int ny, int nz, const double *cx, const double *cy, const double *cz, double lx, double ly,
                         double lz, int z0, int nzl, void *diag, int mem_kind)
{
    const int n[3] = { nx, ny, nz };
    const double *c[3] = { cx, cy, cz };
    const double lam[3] = { lx, ly, lz };
    const long long cells = (long long)nx * ny * nzl;
    double2 *tab = nullptr, *out = (double2 *)diag, *tmp = nullptr;
    std::vector<double2> h((size_t)nx + ny + nz);
    size_t o = 0;
    for (int a = 0; a < 3; ++a)
        for (int m = 0; m < n[a]; ++m, ++o)     // the "+1" (VecShift, :155) rides on y, as in the plan's tables
            h[o] = make_double2(lam[a] * c[a][2 * m] + (a == 1 ? 1.0 : 0.0), lam[a] * c[a][2 * m + 1]);
    CPC_CUDA(cudaMalloc(&tab, sizeof(double2) * h.size()));
    int rc = CPC_OK;
    auto fail = [&](cudaError_t e, const char *what) { if (e != cudaSuccess && rc == CPC_OK) rc = cuda_fail(e, what); };
    fail(cudaMemcpy(tab, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice), "cudaMemcpy(tables)");
    if (rc == CPC_OK && mem_kind == CPC_MEM_HOST) {
        fail(cudaMalloc(&tmp, sizeof(double2) * (size_t)(cells > 0 ? cells : 1)), "cudaMalloc(diag)");
        out = tmp;
    }
    if (rc == CPC_OK && cells > 0) {
        diag_from_separable_kernel<<<1184, 256>>>(out, tab, tab + nx, tab + nx + ny + z0, nx, ny, cells);
        fail(cudaGetLastError(), "diag_from_separable_kernel");
        if (tmp) fail(cudaMemcpy(diag, tmp, sizeof(double2) * (size_t)cells, cudaMemcpyDeviceToHost), "cudaMemcpy(diag)");
        fail(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    }
    if (tmp) cudaFree(tmp);
    cudaFree(tab);
    return rc;
}

PlanBase *make_plan_f64() { return new PlanT<double>(); }
PlanBase *make_pencil_plan_f64(int p_rows, int p_cols) { return new PencilPlanT<double>(p_rows, p_cols); }

}  // namespace cpc
