// cpc_api.cu -- the C ABI of include/circulantpc.h: argument checking, error strings, plan life cycle and the
// pure-host slab helpers.  No compute here; kernels live in fft_pass.cuh / generic_pass.cuh.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "dist.h"
#include "plan.h"

namespace cpc {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    cudaGetLastError();   // clear the sticky-less error state
    return e == cudaErrorMemoryAllocation ? CPC_ERR_NOMEM : CPC_ERR_CUDA;
}

int symbol_recurrence_lambda(int nx, int ny, int nz, const double2 *ax, const double2 *ay, const double2 *az, double *lambda_z)
{
    double lz = 0.0;
    if (nz > 1) {
        double re, im;
        exact_root(1, nz, &re, &im);
        lz = az[1].x / (1.0 - re);
    }
    if (lambda_z) *lambda_z = lz;
    bool ok = std::isfinite(lz) && lz >= 0.0 && lz <= 4096.0;
    const double tol = 1e-13 * (lz > 1.0 ? lz : 1.0);
    for (int m = 0; ok && m < nz; ++m) {
        double re = 1.0, im = 0.0;
        if (nz > 1) exact_root(m, nz, &re, &im);
        const double wr = nz > 1 ? lz * (1.0 - re) : 0.0, wi = nz > 1 ? -lz * im : 0.0;
        if (std::fabs(az[m].x - wr) > tol || std::fabs(az[m].y - wi) > tol) ok = false;
    }
    double mnx = ax[0].x, mny = ay[0].x;
    for (int m = 1; m < nx; ++m) mnx = ax[m].x < mnx ? ax[m].x : mnx;
    for (int m = 1; m < ny; ++m) mny = ay[m].x < mny ? ay[m].x : mny;
    if (!(mnx + mny >= 0.5)) ok = false;
    return ok ? 1 : 0;
}

SlabRange slab_range(int n, int nranks, int rank)
{
    const int base = n / nranks, rem = n % nranks;
    SlabRange r;
    r.count = base + (rank < rem ? 1 : 0);
    r.start = rank * base + (rank < rem ? rank : rem);
    return r;
}

}  // namespace cpc

using namespace cpc;

#define CHECK_PLAN(p)                                                       \
    if (!(p) || !(p)->impl) { set_error("null plan"); return CPC_ERR_ARG; }

extern "C" {

const char *cpc_last_error(void) { return cpc::g_err; }
int cpc_version(void) { return CPC_VERSION_MAJOR * 1000 + CPC_VERSION_MINOR; }

int cpc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int cpc_plan_create(cpc_plan *plan, const cpc_plan_desc *d)
{
    if (!plan || !d) { set_error("cpc_plan_create: null argument"); return CPC_ERR_ARG; }
    *plan = nullptr;
    if (d->nx < 1 || d->ny < 1 || d->nz < 1) {
        set_error("cpc_plan_create: extents must be >= 1 (got %d %d %d)", d->nx, d->ny, d->nz);
        return CPC_ERR_ARG;
    }
    if (d->ncomp != 1 && d->ncomp != 4) { set_error("cpc_plan_create: ncomp must be 1 or 4 (got %d)", d->ncomp); return CPC_ERR_ARG; }
    if (d->dtype < CPC_C128 || d->dtype > CPC_F32) { set_error("cpc_plan_create: unknown dtype %d", d->dtype); return CPC_ERR_ARG; }
    if (d->nranks < 1 || d->rank < 0 || d->rank >= d->nranks) {
        set_error("cpc_plan_create: bad rank %d of %d", d->rank, d->nranks);
        return CPC_ERR_ARG;
    }
    if ((long long)d->nx * d->ncomp * (long long)d->ny >= (1ll << 31)) {
        set_error("cpc_plan_create: nx*ncomp*ny must be < 2^31");
        return CPC_ERR_UNSUPPORTED;
    }
    int ndev = cpc_device_count();
    if (ndev <= 0) { set_error("cpc_plan_create: no CUDA device (this library has no CPU fallback)"); return CPC_ERR_CUDA; }
    int dev = d->device;
    if (dev < 0) CPC_CUDA(cudaGetDevice(&dev));
    if (dev >= ndev) { set_error("cpc_plan_create: device %d out of range (%d visible)", dev, ndev); return CPC_ERR_ARG; }
    CPC_CUDA(cudaSetDevice(dev));
    PlanBase *impl = (d->dtype == CPC_C128 || d->dtype == CPC_F64) ? make_plan_f64() : make_plan_f32();
    if (!impl) { set_error("out of host memory"); return CPC_ERR_NOMEM; }
    impl->desc = *d;
    impl->device = dev;
    impl->stream = (cudaStream_t)d->stream;
    int rc = impl->init();
    if (rc) { delete impl; return rc; }
    cpc_plan p = new (std::nothrow) cpc_plan_s;
    if (!p) { delete impl; set_error("out of host memory"); return CPC_ERR_NOMEM; }
    p->impl = impl;
    *plan = p;
    return CPC_OK;
}

int cpc_destroy(cpc_plan plan)
{
    if (!plan) return CPC_OK;
    if (plan->impl) {
        cudaSetDevice(plan->impl->device);
        cudaStreamSynchronize(plan->impl->stream);
        delete plan->impl;
    }
    delete plan;
    return CPC_OK;
}

int cpc_set_stream(cpc_plan plan, void *stream)
{
    CHECK_PLAN(plan);
    plan->impl->stream = (cudaStream_t)stream;
    return CPC_OK;
}

int cpc_sync(cpc_plan plan)
{
    CHECK_PLAN(plan);
    CPC_CUDA(cudaSetDevice(plan->impl->device));
    CPC_CUDA(cudaStreamSynchronize(plan->impl->stream));
    return plan->impl->health();
}

int cpc_set_symbol_transport(cpc_plan plan, double lx, double ly, double lz)
{
    CHECK_PLAN(plan);
    CPC_CUDA(cudaSetDevice(plan->impl->device));
    return plan->impl->set_symbol_transport(lx, ly, lz);
}

int cpc_set_symbol_separable(cpc_plan plan, const double *cx, const double *cy, const double *cz, double lx, double ly,
                             double lz)
{
    CHECK_PLAN(plan);
    CPC_CUDA(cudaSetDevice(plan->impl->device));
    return plan->impl->set_symbol_separable(cx, cy, cz, lx, ly, lz);
}

int cpc_set_symbol_diag(cpc_plan plan, const void *diag, int mem_kind)
{
    CHECK_PLAN(plan);
    if (mem_kind != CPC_MEM_DEVICE && mem_kind != CPC_MEM_HOST) { set_error("bad mem_kind %d", mem_kind); return CPC_ERR_ARG; }
    CPC_CUDA(cudaSetDevice(plan->impl->device));
    return plan->impl->set_symbol_diag(diag, mem_kind);
}

int cpc_set_symbol_first_column(cpc_plan plan, const void *col, int mem_kind)
{
    CHECK_PLAN(plan);
    if (mem_kind != CPC_MEM_DEVICE && mem_kind != CPC_MEM_HOST) { set_error("bad mem_kind %d", mem_kind); return CPC_ERR_ARG; }
    CPC_CUDA(cudaSetDevice(plan->impl->device));
    return plan->impl->set_symbol_first_column(col, mem_kind);
}

int cpc_set_symbol_wave(cpc_plan plan, double c0, double mx, double my, double mz)
{
    CHECK_PLAN(plan);
    return plan->impl->set_symbol_wave(c0, mx, my, mz);
}

int cpc_set_option(cpc_plan plan, int option, long long value)
{
    CHECK_PLAN(plan);
    return plan->impl->set_option(option, value);
}

int cpc_get_diag(cpc_plan plan, void *diag, int mem_kind)
{
    CHECK_PLAN(plan);
    if (!diag) { set_error("null diag"); return CPC_ERR_ARG; }
    CPC_CUDA(cudaSetDevice(plan->impl->device));
    return plan->impl->get_diag(diag, mem_kind);
}

int cpc_build_diag_separable(int nx, int ny, int nz, const double *cx, const double *cy, const double *cz, double lx, double ly,
                             double lz, int z0, int nzl, void *diag, int mem_kind)
{
    if (nx < 1 || ny < 1 || nz < 1 || !cx || !cy || !cz || !diag || z0 < 0 || nzl < 0 || z0 + nzl > nz) {
        set_error("cpc_build_diag_separable: bad argument");
        return CPC_ERR_ARG;
    }
    if (mem_kind != CPC_MEM_DEVICE && mem_kind != CPC_MEM_HOST) { set_error("bad mem_kind %d", mem_kind); return CPC_ERR_ARG; }
    if (cpc_device_count() <= 0) { set_error("cpc_build_diag_separable: no CUDA device (this library has no CPU fallback)"); return CPC_ERR_CUDA; }
    return build_diag_separable(nx, ny, nz, cx, cy, cz, lx, ly, lz, z0, nzl, diag, mem_kind);
}

// device arrays are accessed with 16-byte vector loads / stores (complex128; two complex64, two float64, four float32)
static int check_alignment(const void *a, const void *b, int mem_kind, const char *who)
{
    if (mem_kind == CPC_MEM_DEVICE && ((((uintptr_t)a) | ((uintptr_t)b)) & 15u) != 0) {
        set_error("%s: device pointers must be 16-byte aligned", who);
        return CPC_ERR_ARG;
    }
    return CPC_OK;
}

int cpc_apply(cpc_plan plan, const void *b, void *x, int mem_kind)
{
    CHECK_PLAN(plan);
    if (int rc = check_alignment(b, x, mem_kind, "cpc_apply")) return rc;
    return plan->impl->apply(b, x, mem_kind, nullptr, nullptr);
}

int cpc_apply_profiled(cpc_plan plan, const void *b, void *x, float *pass_ms, int *npasses)
{
    CHECK_PLAN(plan);
    if (!pass_ms || !npasses) { set_error("null output"); return CPC_ERR_ARG; }
    if (int rc = check_alignment(b, x, CPC_MEM_DEVICE, "cpc_apply_profiled")) return rc;
    return plan->impl->apply(b, x, CPC_MEM_DEVICE, pass_ms, npasses);
}

int cpc_forward(cpc_plan plan, const void *in, void *out, int mem_kind)
{
    CHECK_PLAN(plan);
    if (int rc = check_alignment(in, out, mem_kind, "cpc_forward")) return rc;
    return plan->impl->transform(in, out, mem_kind, -1);
}

int cpc_inverse(cpc_plan plan, const void *in, void *out, int mem_kind)
{
    CHECK_PLAN(plan);
    if (int rc = check_alignment(in, out, mem_kind, "cpc_inverse")) return rc;
    return plan->impl->transform(in, out, mem_kind, +1);
}

int cpc_set_projection(cpc_plan plan, int64_t cols, const int64_t *rowptr, const int32_t *colidx, const double *val)
{
    CHECK_PLAN(plan);
    if (cols < 1 || !rowptr || !colidx || !val) { set_error("cpc_set_projection: bad argument"); return CPC_ERR_ARG; }
    CPC_CUDA(cudaSetDevice(plan->impl->device));
    return plan->impl->set_projection(cols, rowptr, colidx, val);
}

int cpc_apply_projected(cpc_plan plan, const void *b, void *x, int mem_kind)
{
    CHECK_PLAN(plan);
    if (!b || !x) { set_error("cpc_apply_projected: null pointer"); return CPC_ERR_ARG; }
    if (mem_kind != CPC_MEM_DEVICE && mem_kind != CPC_MEM_HOST) { set_error("bad mem_kind %d", mem_kind); return CPC_ERR_ARG; }
    if (int rc = check_alignment(b, x, mem_kind, "cpc_apply_projected")) return rc;
    CPC_CUDA(cudaSetDevice(plan->impl->device));
    return plan->impl->apply_projected(b, x, mem_kind);
}

int cpc_get_info(cpc_plan plan, cpc_plan_info *info)
{
    CHECK_PLAN(plan);
    if (!info) { set_error("null info"); return CPC_ERR_ARG; }
    return plan->impl->get_info(info);
}

int cpc_slab_range(int n, int nranks, int rank, int *start, int *count)
{
    if (n < 1 || nranks < 1 || rank < 0 || rank >= nranks || !start || !count) { set_error("cpc_slab_range: bad argument"); return CPC_ERR_ARG; }
    const SlabRange r = slab_range(n, nranks, rank);
    *start = r.start;
    *count = r.count;
    return CPC_OK;
}

int cpc_slab_send_chunk(int nx, int ny, int nz, int ncomp, int nranks, int rank, int q, int64_t *offset, int64_t *count)
{
    if (nx < 1 || ny < 1 || nz < 1 || ncomp < 1 || nranks < 1 || rank < 0 || rank >= nranks || q < 0 || q >= nranks ||
        !offset || !count) {
        set_error("cpc_slab_send_chunk: bad argument");
        return CPC_ERR_ARG;
    }
    const long long W = (long long)nx * ncomp;
    const SlabRange zr = slab_range(nz, nranks, rank);
    long long off = 0;
    for (int p = 0; p < q; ++p) off += (long long)zr.count * slab_range(ny, nranks, p).count * W;
    *offset = off;
    *count = (long long)zr.count * slab_range(ny, nranks, q).count * W;
    return CPC_OK;
}

int cpc_slab_recv_chunk(int nx, int ny, int nz, int ncomp, int nranks, int rank, int s, int64_t *offset, int64_t *count)
{
    if (nx < 1 || ny < 1 || nz < 1 || ncomp < 1 || nranks < 1 || rank < 0 || rank >= nranks || s < 0 || s >= nranks ||
        !offset || !count) {
        set_error("cpc_slab_recv_chunk: bad argument");
        return CPC_ERR_ARG;
    }
    const long long W = (long long)nx * ncomp;
    const SlabRange yr = slab_range(ny, nranks, rank);
    const SlabRange zs = slab_range(nz, nranks, s);
    *offset = (long long)zs.start * yr.count * W;
    *count = (long long)zs.count * yr.count * W;
    return CPC_OK;
}

int cpc_symbol_recurrence_lambda(int nx, int ny, int nz, const double *ax, const double *ay, const double *az, double *lambda_z)
{
    if (nx < 1 || ny < 1 || nz < 1 || !ax || !ay || !az) return -1;
    return symbol_recurrence_lambda(nx, ny, nz, (const double2 *)ax, (const double2 *)ay, (const double2 *)az, lambda_z);
}

int cpc_nccl_unique_id(void *out_bytes)
{
    if (!out_bytes) { set_error("null output"); return CPC_ERR_ARG; }
    return dist_unique_id(out_bytes);
}

}  // extern "C"
