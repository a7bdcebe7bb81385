// pencil_impl.cuh -- PencilPlanT<T>: the pencil (P_r x P_c) decomposition of the scalar circulant apply (pencil.h).
//
// Built from parts that single-rank plans already run: the three kinds of local pass are passes of three single-rank
// sub-plans whose extents are the local pencil (x lines of nx x nyl x nzl, y lines of nxl x ny x nzl, the fused middle
// pass -- FFT form or z recurrence, whichever the sub-plan picks for the symbol -- of nxl x nyl2 x nz with this rank's
// slices of the eigenvalue tables).  New here: one reordering kernel (SWAP) and the group all-to-all.  The exchanges go
// through NCCL (grouped send / recv among the members of the row or column group on the world communicator), or, for
// plans created without an NCCL id, through peer copies issued by cpc_pencil_apply_lockstep, which drives the plans of
// all ranks from one process.
//
// Scalar complex plans (CPC_C128 / CPC_C64) with the transport / separable symbol: what BASELINE config 4 sweeps.
// 5 transform passes + 6 reordering passes + 4 all-to-alls per apply, against 5 passes and one carry exchange for
// z-slabs: the pencil grid is for rank counts beyond nz / for comparison, z-slabs stay the default.
#pragma once
#include "pencil.h"
#include "plan_impl.cuh"

namespace cpc {

template <typename C>
__global__ void __launch_bounds__(256)
pencil_swap_kernel(const C *in, C *out, long long A, long long B, long long inner, long long total, double scale)
{
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        C v = in[pencil_swap_source(o, A, B, inner)];
        v.x = (decltype(v.x))(v.x * scale);
        v.y = (decltype(v.y))(v.y * scale);
        out[o] = v;
    }
}

template <typename T> struct PencilPlanT : PlanBase, PencilIface {
    using C = cplx_t<T>;
    PencilLayout L{};
    std::vector<PencilStep> sched;
    PlanT<T> *sub[3] = { nullptr, nullptr, nullptr };      // x lines, y lines, middle pass
    C *buf[2] = { nullptr, nullptr };
    DistState dist;
    bool loopback = false;         // no NCCL id: the exchanges are done by cpc_pencil_apply_lockstep
    int pr = 1, pc = 1;
    // which array every step reads / writes (pencil_buffer_plan), for device arrays and for staged host arrays
    std::vector<PencilBuffers> bufplan[2];
    int final_buf[2] = { PBUF_X, PBUF_W0 };
    // state of the apply in flight
    const void *arr[4] = { nullptr, nullptr, nullptr, nullptr };     // PBUF_B, PBUF_W0, PBUF_W1, PBUF_X
    int staged = 0;
    void *host_out = nullptr;

    PencilPlanT(int pr_, int pc_) : pr(pr_), pc(pc_) {}

    ~PencilPlanT() override
    {
        cudaSetDevice(device);
        for (auto *s : sub) delete s;
        for (auto *b : buf)
            if (b) cudaFree(b);
        dist_destroy(dist);
    }

    int make_sub(int k, int sx, int sy, int sz)
    {
        PlanT<T> *p = new PlanT<T>();
        p->desc = desc;
        p->desc.nx = sx; p->desc.ny = sy; p->desc.nz = sz;
        p->desc.nranks = 1; p->desc.rank = 0; p->desc.nccl_unique_id = nullptr;
        p->device = device;
        p->stream = stream;
        sub[k] = p;
        return p->init();
    }

    int init() override
    {
        if (desc.ncomp != 1) { set_error("pencil plans: ncomp == 1 only"); return CPC_ERR_UNSUPPORTED; }
        if (desc.dtype != CPC_C128 && desc.dtype != CPC_C64) { set_error("pencil plans: complex dtypes only"); return CPC_ERR_UNSUPPORTED; }
        if (pr * pc != desc.nranks) { set_error("pencil plans: p_rows * p_cols must equal nranks (%d x %d != %d)", pr, pc, desc.nranks); return CPC_ERR_ARG; }
        const int rc0 = pencil_make_layout(desc.nx, desc.ny, desc.nz, pr, pc, desc.rank, &L);
        if (rc0 == -2) {
            set_error("pencil plans need nx, ny divisible by p_rows and ny, nz by p_cols (%d %d %d on %d x %d)", desc.nx, desc.ny,
                      desc.nz, pr, pc);
            return CPC_ERR_UNSUPPORTED;
        }
        if (rc0) { set_error("pencil plans: bad grid or rank"); return CPC_ERR_ARG; }
        sched = pencil_schedule(L);
        for (int st = 0; st < 2; ++st) bufplan[st] = pencil_buffer_plan(sched, st != 0, &final_buf[st]);
        int rc;
        if ((rc = make_sub(0, L.nx, L.nyl, L.nzl))) return rc;
        if ((rc = make_sub(1, L.nxl, L.ny, L.nzl))) return rc;
        if ((rc = make_sub(2, L.nxl, L.nyl2, L.nz))) return rc;
        for (auto &b : buf) CPC_CUDA(cudaMalloc(&b, sizeof(C) * (size_t)L.nloc));
        loopback = desc.nranks > 1 && desc.nccl_unique_id == nullptr;
        if (desc.nranks > 1 && !loopback && (rc = dist_init(dist, desc.nranks, desc.rank, desc.nccl_unique_id, device))) return rc;
        return CPC_OK;
    }

    // ---- symbols: this rank's slices of the three 1-D tables go to the middle sub-plan ---------------------------
    int set_symbol_separable(const double *cx, const double *cy, const double *cz, double lx, double ly, double lz) override
    {
        if (!cx || !cy || !cz) { set_error("null eigenvalue table"); return CPC_ERR_ARG; }
        std::vector<double2> h[3];
        h[0].resize(L.nxl); h[1].resize(L.nyl2); h[2].resize(L.nz);
        for (int m = 0; m < L.nxl; ++m) h[0][m] = make_double2(lx * cx[2 * (L.x0 + m)], lx * cx[2 * (L.x0 + m) + 1]);
        for (int m = 0; m < L.nyl2; ++m)      // the "+1" (VecShift, FftLinearSolver_3D.c:155) rides on y, as in PlanT
            h[1][m] = make_double2(ly * cy[2 * (L.y02 + m)] + 1.0, ly * cy[2 * (L.y02 + m) + 1]);
        for (int m = 0; m < L.nz; ++m) h[2][m] = make_double2(lz * cz[2 * m], lz * cz[2 * m + 1]);
        sub[2]->stream = stream;
        int rc = sub[2]->set_symbol_tables(h);
        if (rc) return rc;
        symbol_kind = CPC_SYMBOL_SEPARABLE;
        return CPC_OK;
    }

    int set_symbol_transport(double lx, double ly, double lz) override
    {
        // c = [1,-1,0..] (FftLinearSolver_3D.c:80-90)  =>  c_hat[q] = 1 - exp(-2 pi i q / n); n == 1 => 0
        const int n[3] = { L.nx, L.ny, L.nz };
        std::vector<double> ch[3];
        for (int a = 0; a < 3; ++a) {
            ch[a].assign(2 * (size_t)n[a], 0.0);
            if (n[a] > 1)
                for (int q = 0; q < n[a]; ++q) {
                    double re, im;
                    exact_root(q, n[a], &re, &im);
                    ch[a][2 * q] = 1.0 - re;
                    ch[a][2 * q + 1] = -im;
                }
        }
        return set_symbol_separable(ch[0].data(), ch[1].data(), ch[2].data(), lx, ly, lz);
    }

    int unsupported(const char *what)
    {
        set_error("%s: not available on pencil plans (transport / separable symbols, cpc_apply)", what);
        return CPC_ERR_UNSUPPORTED;
    }
    int set_symbol_diag(const void *, int) override { return unsupported("cpc_set_symbol_diag"); }
    int set_symbol_first_column(const void *, int) override { return unsupported("cpc_set_symbol_first_column"); }
    int set_symbol_wave(double, double, double, double) override { return unsupported("cpc_set_symbol_wave"); }
    int get_diag(void *, int) override { return unsupported("cpc_get_diag"); }
    int transform(const void *, void *, int, int) override { return unsupported("cpc_forward / cpc_inverse"); }
    int set_projection(int64_t, const int64_t *, const int32_t *, const double *) override { return unsupported("cpc_set_projection"); }
    int apply_projected(const void *, void *, int) override { return unsupported("cpc_apply_projected"); }
    int set_option(int option, long long value) override { return sub[2]->set_option(option, value); }

    // ---- PencilIface ---------------------------------------------------------------------------------------------
    const PencilLayout &layout() const override { return L; }
    const std::vector<PencilStep> &steps() const override { return sched; }
    size_t elem_bytes() const override { return sizeof(C); }
    bool in_process() const override { return loopback || desc.nranks == 1; }

    int begin(const void *b, void *x, int mem_kind) override
    {
        if (symbol_kind != CPC_SYMBOL_SEPARABLE) { set_error("cpc_apply: no symbol set (call cpc_set_symbol_* first)"); return CPC_ERR_STATE; }
        if (!b || !x) { set_error("cpc_apply: null pointer"); return CPC_ERR_ARG; }
        if (mem_kind != CPC_MEM_DEVICE && mem_kind != CPC_MEM_HOST) { set_error("bad mem_kind %d", mem_kind); return CPC_ERR_ARG; }
        CPC_CUDA(cudaSetDevice(device));
        for (auto *s : sub) s->stream = stream;
        arr[PBUF_W0] = buf[0];
        arr[PBUF_W1] = buf[1];
        if (mem_kind == CPC_MEM_HOST) {
            CPC_CUDA(cudaMemcpyAsync(buf[0], b, sizeof(C) * (size_t)L.nloc, cudaMemcpyHostToDevice, stream));
            h2d_bytes += sizeof(C) * (size_t)L.nloc;
            staged = 1;
            arr[PBUF_B] = arr[PBUF_X] = nullptr;        // never touched by the staged plan
            host_out = x;
        } else {
            staged = 0;
            arr[PBUF_B] = b;
            arr[PBUF_X] = x;
            host_out = nullptr;
        }
        return CPC_OK;
    }

    const C *src_of(size_t k) const { return (const C *)arr[bufplan[staged][k].src]; }
    C *dst_of(size_t k) const { return (C *)arr[bufplan[staged][k].dst]; }

    int local_step(size_t k) override
    {
        const PencilStep &s = sched[k];
        int rc = CPC_OK;
        CPC_CUDA(cudaSetDevice(device));        // (cpc_pencil_apply_lockstep may drive plans on several devices)
        const C *src = src_of(k);
        C *out = dst_of(k);
        if (s.kind == PSTEP_SWAP) {
            const long long total = s.A * s.B * s.inner;
            const long long want = (total + 255) / 256;
            const int grid = (int)(want < 148ll * 16 ? (want > 0 ? want : 1) : 148ll * 16);
            pencil_swap_kernel<C><<<grid, 256, 0, stream>>>(src, out, s.A, s.B, s.inner, total, s.scale);
            ++launches;
            CPC_CUDA(cudaGetLastError());
            return CPC_OK;
        }
        if (s.kind == PSTEP_PASS_X) rc = sub[0]->run_pass(0, s.dir < 0 ? MODE_FWD : MODE_INV, src, out, 0, L.nzl, 0, stream);
        else if (s.kind == PSTEP_PASS_Y) rc = sub[1]->run_pass(1, s.dir < 0 ? MODE_FWD : MODE_INV, src, out, 0, L.nzl, 0, stream);
        else if (s.kind == PSTEP_MIDDLE) rc = sub[2]->run_pass(2, sub[2]->fused_mode(), src, out, 0, L.nz, 0, stream);
        else { set_error("pencil schedule: step %d is not a local step", (int)k); return CPC_ERR_STATE; }
        return rc;
    }

    int exchange_buffers(size_t k, const void **send, void **recv) override
    {
        *send = src_of(k);
        *recv = dst_of(k);
        return CPC_OK;
    }

    int finish() override
    {
        if (host_out) {
            CPC_CUDA(cudaMemcpyAsync(host_out, arr[final_buf[1]], sizeof(C) * (size_t)L.nloc, cudaMemcpyDeviceToHost, stream));
            d2h_bytes += sizeof(C) * (size_t)L.nloc;
            CPC_CUDA(cudaStreamSynchronize(stream));
        }
        return CPC_OK;
    }

    int apply(const void *b, void *x, int mem_kind, float *pass_ms, int *) override
    {
        if (pass_ms) return unsupported("cpc_apply_profiled");
        if (loopback) { set_error("cpc_apply: this pencil plan has no NCCL communicator; drive it with cpc_pencil_apply_lockstep"); return CPC_ERR_STATE; }
        int rc = begin(b, x, mem_kind);
        if (rc) return rc;
        std::vector<int> peers((size_t)(pr > pc ? pr : pc));
        for (size_t k = 0; k < sched.size(); ++k) {
            const int kind = sched[k].kind;
            if (kind == PSTEP_A2A_ROW || kind == PSTEP_A2A_COL) {
                const void *send;
                void *recv;
                if ((rc = exchange_buffers(k, &send, &recv))) return rc;
                const int np = pencil_group(L, kind, peers.data());
                if ((rc = dist_alltoall_group(dist, send, recv, sizeof(C) * (size_t)(L.nloc / np), peers.data(), np, stream))) return rc;
            } else if ((rc = local_step(k))) {
                return rc;
            }
        }
        return finish();
    }

    int get_info(cpc_plan_info *info) override
    {
        memset(info, 0, sizeof(*info));
        info->nx = L.nx; info->ny = L.ny; info->nz = L.nz;
        info->ncomp = 1; info->dtype = desc.dtype;
        info->nranks = desc.nranks; info->rank = desc.rank;
        info->symbol_kind = symbol_kind;
        int passes = 0;
        for (const PencilStep &s : sched)
            if (s.kind != PSTEP_A2A_ROW && s.kind != PSTEP_A2A_COL) ++passes;
        info->passes_per_apply = passes;
        info->dist_mode = 4;
        cpc_plan_info si;
        for (int a = 0; a < 3; ++a) {
            sub[a]->get_info(&si);
            info->fast_path[a] = si.fast_path[a];
        }
        info->local_elems = L.nloc;
        info->bytes_per_apply_alg = 5ll * 2 * L.nloc * (long long)sizeof(C);
        uint64_t ln = launches;
        for (auto *s : sub) ln += s->launches;
        info->kernel_launches = ln;
        info->h2d_bytes = h2d_bytes;
        info->d2h_bytes = d2h_bytes;
        return CPC_OK;
    }
};

}  // namespace cpc
