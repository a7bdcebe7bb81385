// plan.h -- host-side plan object behind the C ABI (include/circulantpc.h).  Internal header.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/circulantpc.h"

namespace cpc {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define CPC_CUDA(call)                                                  \
    do {                                                                \
        cudaError_t _e = (call);                                        \
        if (_e != cudaSuccess) return ::cpc::cuda_fail(_e, #call);      \
    } while (0)

#include <cmath>

// exp(-2 pi i m / n) rounded from long double
inline void exact_root(long long m, long long n, double *re, double *im)
{
    m %= n;
    const long double a = 2.0L * 3.141592653589793238462643383279502884L * (long double)m / (long double)n;
    // use symmetries so that the classic points are exact
    if (4 * m == n) { *re = 0.0; *im = -1.0; return; }
    if (2 * m == n) { *re = -1.0; *im = 0.0; return; }
    if (4 * m == 3 * n) { *re = 0.0; *im = 1.0; return; }
    *re = (double)cosl(a);
    *im = (double)(-sinl(a));
}


// Does a separable symbol (three 1-D tables already multiplied by their lambdas, the "+1" riding on y) admit the
// recurrence form of the middle pass (zsolve.cuh)?  The z table must be lambda_z (1 - exp(-2 pi i k / nz)) -- the DFT of
// the reference's upwind column [1, -1, 0, ...] (build_transport_col, FftLinearSolver_3D.c:80-90) -- for some
// 0 <= lambda_z <= 4096, entry by entry to 1e-13 max(1, lambda_z), and Re(ax[i] + ay[j]) >= 1/2 everywhere, so that
// |lambda_z / (alpha + lambda_z)| < 1.  Pure host code (cpc_symbol_recurrence_lambda exposes it).
int symbol_recurrence_lambda(int nx, int ny, int nz, const double2 *ax, const double2 *ay, const double2 *az, double *lambda_z);

// Slab bookkeeping shared by the multi-rank code and the pure-host ABI helpers.
struct SlabRange { int start, count; };
SlabRange slab_range(int n, int nranks, int rank);

struct NcclApi;   // dlopen'ed NCCL entry points (dist.cu)

// Type-erased plan; PlanT<T> (plan_impl.cuh) implements it for double / float.
struct PlanBase {
    cpc_plan_desc desc{};
    cudaStream_t stream = nullptr;
    int device = 0;
    int symbol_kind = CPC_SYMBOL_NONE;
    uint64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0;
    virtual ~PlanBase() {}
    virtual int init() = 0;
    virtual int set_symbol_separable(const double *cx, const double *cy, const double *cz, double lx, double ly,
                                     double lz) = 0;
    virtual int set_symbol_transport(double lx, double ly, double lz) = 0;
    virtual int set_symbol_diag(const void *diag, int mem_kind) = 0;
    virtual int set_symbol_first_column(const void *col, int mem_kind) = 0;
    virtual int set_symbol_wave(double c0, double mx, double my, double mz) = 0;
    virtual int set_option(int option, long long value) = 0;
    virtual int health() { return CPC_OK; }
    virtual int get_diag(void *diag, int mem_kind) = 0;
    virtual int apply(const void *b, void *x, int mem_kind, float *pass_ms, int *npasses) = 0;
    virtual int transform(const void *in, void *out, int mem_kind, int dir) = 0;
    virtual int get_info(cpc_plan_info *info) = 0;
    virtual int set_projection(int64_t cols, const int64_t *rowptr, const int32_t *colidx, const double *val) = 0;
    virtual int apply_projected(const void *b, void *x, int mem_kind) = 0;
};

int build_diag_separable(int nx, int ny, int nz, const double *cx, const double *cy, const double *cz, double lx, double ly,
                         double lz, int z0, int nzl, void *diag, int mem_kind);
PlanBase *make_plan_f64();
PlanBase *make_plan_f32();
// pencil (p_rows x p_cols) plans, pencil_impl.cuh
PlanBase *make_pencil_plan_f64(int p_rows, int p_cols);
PlanBase *make_pencil_plan_f32(int p_rows, int p_cols);

}  // namespace cpc

struct cpc_plan_s {
    cpc::PlanBase *impl;
};
