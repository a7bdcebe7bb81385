// cpc_api.cu -- the C ABI of include/circulantpc.h: argument checking, error strings, plan life cycle and the
// pure-host slab helpers.  No compute here; kernels live in fft_pass.cuh / generic_pass.cuh.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include <vector>

#include "dist.h"
#include "pencil.h"
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

static int check_alignment(const void *a, const void *b, int mem_kind, const char *who);

// p_rows == 0: the z-slab / single-rank plan of cpc_plan_create; otherwise a pencil plan on a p_rows x p_cols grid
static int create_plan(cpc_plan *plan, const cpc_plan_desc *d, int p_rows, int p_cols)
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
    const bool f64 = (d->dtype == CPC_C128 || d->dtype == CPC_F64);
    PlanBase *impl = p_rows > 0 ? (f64 ? make_pencil_plan_f64(p_rows, p_cols) : make_pencil_plan_f32(p_rows, p_cols))
                                : (f64 ? make_plan_f64() : make_plan_f32());
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

int cpc_plan_create(cpc_plan *plan, const cpc_plan_desc *d) { return create_plan(plan, d, 0, 0); }

int cpc_plan_create_pencil(cpc_plan *plan, const cpc_plan_desc *d, int p_rows, int p_cols)
{
    if (p_rows < 1 || p_cols < 1) { set_error("cpc_plan_create_pencil: the grid must be at least 1 x 1 (got %d x %d)", p_rows, p_cols); return CPC_ERR_ARG; }
    return create_plan(plan, d, p_rows, p_cols);
}

// One apply on all ranks of a pencil grid whose plans live in this process (created without an NCCL id): the local steps
// run on every plan's own device and stream, the all-to-alls are peer copies between the plans' buffers.
int cpc_pencil_apply_lockstep(cpc_plan *plans, int nplans, const void *const *b, void *const *x, int mem_kind)
{
    if (!plans || nplans < 1 || !b || !x) { set_error("cpc_pencil_apply_lockstep: null argument"); return CPC_ERR_ARG; }
    std::vector<PencilIface *> pp((size_t)nplans);
    for (int i = 0; i < nplans; ++i) {
        CHECK_PLAN(plans[i]);
        pp[i] = dynamic_cast<PencilIface *>(plans[i]->impl);
        if (!pp[i]) { set_error("cpc_pencil_apply_lockstep: plan %d is not a pencil plan", i); return CPC_ERR_ARG; }
        const PencilLayout &L = pp[i]->layout(), &L0 = pp[0]->layout();
        if (!pp[i]->in_process() || L.pr * L.pc != nplans || L.rank != i || L.nx != L0.nx || L.ny != L0.ny || L.nz != L0.nz ||
            L.pr != L0.pr || L.pc != L0.pc || pp[i]->elem_bytes() != pp[0]->elem_bytes()) {
            set_error("cpc_pencil_apply_lockstep: plans[i] must be rank i of one p_rows x p_cols grid of %d in-process plans", nplans);
            return CPC_ERR_ARG;
        }
    }
    auto sync_all = [&]() -> int {
        for (int i = 0; i < nplans; ++i) {
            CPC_CUDA(cudaSetDevice(plans[i]->impl->device));
            CPC_CUDA(cudaStreamSynchronize(plans[i]->impl->stream));
        }
        return CPC_OK;
    };
    int rc;
    for (int i = 0; i < nplans; ++i) {
        if ((rc = check_alignment(b[i], x[i], mem_kind, "cpc_pencil_apply_lockstep"))) return rc;
        if ((rc = pp[i]->begin(b[i], x[i], mem_kind))) return rc;
    }
    const std::vector<PencilStep> &steps = pp[0]->steps();
    std::vector<const void *> send((size_t)nplans);
    std::vector<void *> recv((size_t)nplans);
    std::vector<int> peers((size_t)nplans);
    for (size_t k = 0; k < steps.size(); ++k) {
        const int kind = steps[k].kind;
        if (kind != PSTEP_A2A_ROW && kind != PSTEP_A2A_COL) {
            for (int i = 0; i < nplans; ++i)
                if ((rc = pp[i]->local_step(k))) return rc;
            continue;
        }
        for (int i = 0; i < nplans; ++i)
            if ((rc = pp[i]->exchange_buffers(k, &send[i], &recv[i]))) return rc;
        if ((rc = sync_all())) return rc;                    // every rank's send buffer is complete
        for (int i = 0; i < nplans; ++i) {
            const PencilLayout &L = pp[i]->layout();
            const int np = pencil_group(L, kind, peers.data());
            const size_t chunk = pp[i]->elem_bytes() * (size_t)(L.nloc / np);
            const int me = kind == PSTEP_A2A_ROW ? L.r : L.c;            // this rank's place in its group
            CPC_CUDA(cudaSetDevice(plans[i]->impl->device));
            for (int q = 0; q < np; ++q)
                CPC_CUDA(cudaMemcpyAsync((char *)recv[peers[q]] + (size_t)me * chunk, (const char *)send[i] + (size_t)q * chunk,
                                         chunk, cudaMemcpyDefault, plans[i]->impl->stream));
        }
        if ((rc = sync_all())) return rc;                    // every chunk has landed
    }
    for (int i = 0; i < nplans; ++i)
        if ((rc = pp[i]->finish())) return rc;
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

int cpc_pencil_layout(int nx, int ny, int nz, int p_rows, int p_cols, int rank, cpc_pencil_layout_t *out)
{
    PencilLayout L;
    if (!out || pencil_make_layout(nx, ny, nz, p_rows, p_cols, rank, &L)) {
        set_error("cpc_pencil_layout: bad grid, rank or extents (nx, ny divisible by p_rows; ny, nz by p_cols)");
        return CPC_ERR_ARG;
    }
    out->r = L.r; out->c = L.c;
    out->nxl = L.nxl; out->x0 = L.x0; out->nyl = L.nyl; out->y0 = L.y0;
    out->nyl2 = L.nyl2; out->y02 = L.y02; out->nzl = L.nzl; out->z0 = L.z0;
    out->local_elems = L.nloc;
    return CPC_OK;
}

int cpc_pencil_steps(int nx, int ny, int nz, int p_rows, int p_cols, int rank, cpc_pencil_step_t *steps, int max_steps, int *nsteps)
{
    PencilLayout L;
    if (!nsteps || pencil_make_layout(nx, ny, nz, p_rows, p_cols, rank, &L)) {
        set_error("cpc_pencil_steps: bad grid, rank or extents");
        return CPC_ERR_ARG;
    }
    const std::vector<PencilStep> s = pencil_schedule(L);
    *nsteps = (int)s.size();
    if (!steps) return CPC_OK;
    if (max_steps < (int)s.size()) { set_error("cpc_pencil_steps: %d steps, room for %d", (int)s.size(), max_steps); return CPC_ERR_ARG; }
    const std::vector<PencilBuffers> dev = pencil_buffer_plan(s, false, nullptr), stg = pencil_buffer_plan(s, true, nullptr);
    for (size_t k = 0; k < s.size(); ++k) {
        steps[k].kind = s[k].kind; steps[k].dir = s[k].dir;
        steps[k].a = s[k].A; steps[k].b = s[k].B; steps[k].inner = s[k].inner;
        steps[k].scale = s[k].scale;
        steps[k].src_buf = dev[k].src; steps[k].dst_buf = dev[k].dst;
        steps[k].src_buf_staged = stg[k].src; steps[k].dst_buf_staged = stg[k].dst;
    }
    return CPC_OK;
}

int cpc_pencil_group(int nx, int ny, int nz, int p_rows, int p_cols, int rank, int step_kind, int *peers, int *npeers)
{
    PencilLayout L;
    if (!peers || !npeers || (step_kind != PSTEP_A2A_ROW && step_kind != PSTEP_A2A_COL) ||
        pencil_make_layout(nx, ny, nz, p_rows, p_cols, rank, &L)) {
        set_error("cpc_pencil_group: bad argument");
        return CPC_ERR_ARG;
    }
    *npeers = pencil_group(L, step_kind, peers);
    return CPC_OK;
}

int64_t cpc_pencil_swap_source(int64_t o, int64_t a, int64_t b, int64_t inner)
{
    if (a < 1 || b < 1 || inner < 1 || o < 0 || o >= a * b * inner) return -1;
    return pencil_swap_source(o, a, b, inner);
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
