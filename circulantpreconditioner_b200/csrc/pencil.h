// pencil.h -- layout and schedule of the pencil (P_r x P_c) decomposition.  Pure host code, no CUDA: shared by the GPU
// plan (pencil_impl.cuh), the host ABI helpers (cpc_pencil_layout / cpc_pencil_steps) and, through them, the CPU tests
// that replay the schedule on virtual ranks and over gloo.
//
// Reference counterpart: MatCreateFFT(PETSC_COMM_WORLD, ...) (reference src/PCSHELLFft_3D.cxx:35) hands the distribution
// to fftw-mpi, which cuts z-slabs only; BASELINE config 4 asks for the sweep "pencil-sharded at 2/4/8 B200", which lifts
// the slab limit P <= nz.
//
// Rank (r, c) of the P_r x P_c grid has number c * P_r + r.  Three distributions of the nx x ny x nz grid, x fastest:
//   X pencils   [z in slab c of P_c][y in slab r of P_r][x]              b and x live here (index i + nx (jl + nyl kl))
//   Y pencils   [z in slab c of P_c][y][x in slab r of P_r]
//   Z pencils   [z][y in slab c of P_c][x in slab r of P_r]              the fused middle pass runs here
// X <-> Y exchanges within a row group (the P_r ranks of one c), Y <-> Z within a column group (the P_c ranks of one r).
// Every exchange is an all-to-all of equal contiguous chunks; the local reordering around it is one kernel of one
// kind: SWAP(A, B, inner) turns in[a][b][inner] into out[b][a][inner].
#pragma once
#include <cstddef>
#include <vector>

namespace cpc {

struct PencilLayout {
    int nx, ny, nz, pr, pc, rank;
    int r, c;                 // grid coordinates of `rank`
    int nxl, x0;              // x slab of the Y and Z pencils (over P_r)
    int nyl, y0;              // y slab of the X pencils (over P_r)
    int nyl2, y02;            // y slab of the Z pencils (over P_c)
    int nzl, z0;              // z slab of the X and Y pencils (over P_c)
    long long nloc;           // elements per rank, the same in all three distributions
};

enum PencilStepKind {
    PSTEP_PASS_X = 0,         // 1-D transforms along x on the X pencils   (local array nx x nyl x nzl, axis 0)
    PSTEP_PASS_Y = 1,         // 1-D transforms along y on the Y pencils   (local array nxl x ny x nzl, axis 1)
    PSTEP_MIDDLE = 2,         // forward z, division by the eigenvalues, backward z on the Z pencils (nxl x nyl2 x nz)
    PSTEP_SWAP = 3,           // in[a][b][inner] -> out[b][a][inner], optionally scaled
    PSTEP_A2A_ROW = 4,        // all-to-all of nloc / P_r element chunks within the row group
    PSTEP_A2A_COL = 5         // all-to-all of nloc / P_c element chunks within the column group
};

struct PencilStep {
    int kind;
    int dir;                  // passes: -1 forward, +1 backward
    long long A, B, inner;    // SWAP
    double scale;             // SWAP: factor applied on the way (1 except for the last one: 1 / (P_r P_c))
};

// 0 on success; -1 bad grid / rank, -2 extents not divisible (nx, ny by P_r; ny, nz by P_c)
inline int pencil_make_layout(int nx, int ny, int nz, int pr, int pc, int rank, PencilLayout *L)
{
    if (nx < 1 || ny < 1 || nz < 1 || pr < 1 || pc < 1 || rank < 0 || rank >= pr * pc) return -1;
    if (nx % pr || ny % pr || ny % pc || nz % pc) return -2;
    L->nx = nx; L->ny = ny; L->nz = nz; L->pr = pr; L->pc = pc; L->rank = rank;
    L->r = rank % pr; L->c = rank / pr;
    L->nxl = nx / pr;  L->x0 = L->r * L->nxl;
    L->nyl = ny / pr;  L->y0 = L->r * L->nyl;
    L->nyl2 = ny / pc; L->y02 = L->c * L->nyl2;
    L->nzl = nz / pc;  L->z0 = L->c * L->nzl;
    L->nloc = (long long)nx * ny * nz / ((long long)pr * pc);
    return 0;
}

// The ranks of the row group (kind PSTEP_A2A_ROW) or column group (PSTEP_A2A_COL) of L.rank, in chunk order: chunk q of
// the send buffer goes to peers[q], chunk q of the receive buffer comes from peers[q].  Returns the group size.
inline int pencil_group(const PencilLayout &L, int kind, int *peers)
{
    if (kind == PSTEP_A2A_ROW) {
        for (int q = 0; q < L.pr; ++q) peers[q] = L.c * L.pr + q;
        return L.pr;
    }
    for (int q = 0; q < L.pc; ++q) peers[q] = q * L.pr + L.r;
    return L.pc;
}

// One apply: b (X pencils) -> x (X pencils).  The steps alternate between two work buffers; a pass runs in place,
// a SWAP or an all-to-all moves the data to the other buffer.  Passes along an axis of length 1 are left out, as in
// the single-rank schedule; the middle pass always runs (it carries the division).
inline std::vector<PencilStep> pencil_schedule(const PencilLayout &L)
{
    std::vector<PencilStep> s;
    auto pass = [&](int kind, int dir) { s.push_back(PencilStep{ kind, dir, 0, 0, 0, 1.0 }); };
    auto swap = [&](long long A, long long B, long long inner, double scale = 1.0) {
        s.push_back(PencilStep{ PSTEP_SWAP, 0, A, B, inner, scale });
    };
    const long long lines = (long long)L.nzl * L.nyl;          // x lines of an X pencil
    const long long blk = (long long)L.nyl * L.nxl;            // one z plane of one row-group chunk
    const long long blk2 = (long long)L.nyl2 * L.nxl;          // one z plane of one column-group chunk
    if (L.nx > 1) pass(PSTEP_PASS_X, -1);
    // X -> Y pencils: [l][q][xl] -> [q][l][xl] | row all-to-all | [q'][zl][yl xl] -> [zl][q'][yl xl] = [zl][y][xl]
    swap(lines, L.pr, L.nxl);
    pass(PSTEP_A2A_ROW, 0);
    swap(L.pr, L.nzl, blk);
    if (L.ny > 1) pass(PSTEP_PASS_Y, -1);
    // Y -> Z pencils: [zl][q][yl2 xl] -> [q][zl][yl2 xl] | column all-to-all | [q'][zl][..] is [z][yl2][xl] already
    swap(L.nzl, L.pc, blk2);
    pass(PSTEP_A2A_COL, 0);
    pass(PSTEP_MIDDLE, 0);
    // Z -> Y pencils: chunk q = planes of slab q, contiguous | column all-to-all | [q'][zl][yl2 xl] -> [zl][q'][yl2 xl]
    pass(PSTEP_A2A_COL, 0);
    swap(L.pc, L.nzl, blk2);
    if (L.ny > 1) pass(PSTEP_PASS_Y, +1);
    // Y -> X pencils: [zl][q][yl xl] -> [q][zl][yl xl] | row all-to-all | [q'][l][xl] -> [l][q'][xl] = [zl][yl][x]
    swap(L.nzl, L.pr, blk);
    pass(PSTEP_A2A_ROW, 0);
    // the local middle pass normalises by its own 1 / (nxl nyl2 nz): the remaining 1 / (P_r P_c) rides on this SWAP
    swap(L.pr, lines, L.nxl, 1.0 / ((double)L.pr * L.pc));
    if (L.nx > 1) pass(PSTEP_PASS_X, +1);
    return s;
}

// Which array every step reads and writes.  Device arrays: the first step reads the caller's b, the last one writes the
// caller's x (b == x is fine: b is only read by the first step, x only written by the last).  Host arrays are staged:
// b is copied into W0 first, the result is copied out of *final_buf afterwards.  A pass runs in place on a work buffer;
// a SWAP or an all-to-all moves the data to the other work buffer.
enum PencilBufferId { PBUF_B = 0, PBUF_W0 = 1, PBUF_W1 = 2, PBUF_X = 3 };
struct PencilBuffers { int src, dst; };
inline std::vector<PencilBuffers> pencil_buffer_plan(const std::vector<PencilStep> &s, bool staged, int *final_buf)
{
    std::vector<PencilBuffers> out;
    int cur = staged ? PBUF_W0 : PBUF_B;
    for (size_t k = 0; k < s.size(); ++k) {
        const bool pass = s[k].kind == PSTEP_PASS_X || s[k].kind == PSTEP_PASS_Y || s[k].kind == PSTEP_MIDDLE;
        int dst;
        if (k + 1 == s.size() && !staged) dst = PBUF_X;
        else if (pass && cur != PBUF_B) dst = cur;
        else dst = cur == PBUF_W0 ? PBUF_W1 : PBUF_W0;
        out.push_back(PencilBuffers{ cur, dst });
        cur = dst;
    }
    if (final_buf) *final_buf = cur;
    return out;
}

// source element of output element o of SWAP(A, B, inner): out[b][a][t] = in[a][b][t]
#if defined(__CUDACC__)
__host__ __device__
#endif
inline long long pencil_swap_source(long long o, long long A, long long B, long long inner)
{
    const long long t = o % inner, ba = o / inner;
    const long long a = ba % A, b = ba / A;
    return (a * B + b) * inner + t;
}

// What cpc_pencil_apply_lockstep (cpc_api.cu) needs from a pencil plan, whatever its dtype (PencilPlanT, pencil_impl.cuh).
struct PencilIface {
    virtual ~PencilIface() {}
    virtual const PencilLayout &layout() const = 0;
    virtual const std::vector<PencilStep> &steps() const = 0;
    virtual size_t elem_bytes() const = 0;
    virtual bool in_process() const = 0;
    virtual int begin(const void *b, void *x, int mem_kind) = 0;
    virtual int local_step(size_t k) = 0;                           // a pass or a SWAP
    virtual int exchange_buffers(size_t k, const void **send, void **recv) = 0;    // an all-to-all: where from, where to
    virtual int finish() = 0;
};

}  // namespace cpc
