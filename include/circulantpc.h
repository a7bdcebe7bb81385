/* circulantpc.h -- C ABI of libcirculantpc.so (B200 / sm_100a circulant preconditioner apply).
 *
 * This is the drop-in boundary for the one hot path of ndjinga/CirculantPreconditioner:
 *     x = IFFT3( FFT3(b) ./ Lambda )        (reference: src/FftLinearSolver_3D.c:166-190, solve_3D)
 * plus its eigenvalue set-up (src/FftLinearSolver_3D.c:80-164) and the PCShell context life cycle
 * (src/PCSHELLFft_3D.cxx:10-99).  Plain pointers and sizes only; no PETSc, torch or C++ types.
 * All arrays use the reference's layout: index m = i + nx*(j + ny*k), x fastest
 * (dims = {nz, ny, nx}, src/PCSHELLFft_3D.cxx:34); complex numbers are interleaved (re, im).
 * For ncomp = 4 (wave system) unknown c of cell m is element 4*m + c
 * (tests/WaveSystem_SphericalExplosion_impl_mpi.cxx:104-115).
 *
 * Every function returns 0 (CPC_OK) on success, a cpc_status code otherwise; cpc_last_error()
 * returns a thread-local message.  There is no CPU fallback: without a CUDA device every compute
 * entry point fails with CPC_ERR_CUDA.
 *
 * Which reference interface each entry point replaces is noted per function; the reference-named
 * wrappers (solve_3D, build_diag_mat_vec_3D, setupFFTPrec3D, ...) that sit on top of this ABI are in
 * circulantpreconditioner_b200/glue/ and the binding a maintainer would add is in INTEGRATION.md.
 */
#ifndef CIRCULANTPC_H
#define CIRCULANTPC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPC_VERSION_MAJOR 0
#define CPC_VERSION_MINOR 1

typedef struct cpc_plan_s *cpc_plan;

enum cpc_status {
    CPC_OK = 0,
    CPC_ERR_ARG = 1,          /* bad argument (PETSC_ERR_ARG_WRONG in the glue) */
    CPC_ERR_CUDA = 2,         /* CUDA runtime error / no device */
    CPC_ERR_UNSUPPORTED = 3,  /* valid request the library does not implement */
    CPC_ERR_STATE = 4,        /* call order error (e.g. apply before a symbol was set) */
    CPC_ERR_NCCL = 5,         /* NCCL missing or failing (multi-rank plans only) */
    CPC_ERR_NOMEM = 6
};

enum cpc_dtype {
    CPC_C128 = 0,             /* PetscScalar of a complex PETSc build: complex128 in, complex128 out */
    CPC_C64 = 1,              /* fp32 option: complex64 in/out */
    CPC_F64 = 2,              /* PetscScalar of a real PETSc build: float64 in, float64 out; r2c / c2r inside, half
                                 the bytes per pass (the reference's unfinished real branch, FftLinearSolver_3D.c:7-78,
                                 176,186).  ncomp == 1, transport / separable symbol; multi-rank plans (z-slabs, even
                                 nx) run the transpose-free recurrence schedule only. */
    CPC_F32 = 3               /* float32 in/out */
};

enum cpc_mem {
    CPC_MEM_DEVICE = 0,       /* pointers as returned by VecCUDAGetArrayRead/Write */
    CPC_MEM_HOST = 1          /* pointers as returned by VecGetArrayRead/Write (staged through pinned memory) */
};

enum cpc_symbol_kind {
    CPC_SYMBOL_NONE = 0,
    CPC_SYMBOL_SEPARABLE = 1, /* Lambda[k,j,i] = 1 + lx*cx[i] + ly*cy[j] + lz*cz[k] from three 1-D tables */
    CPC_SYMBOL_TABLE = 2,     /* full table of N eigenvalues held in HBM */
    CPC_SYMBOL_WAVE = 3       /* 4x4 arrow matrix per frequency (ncomp = 4) */
};

typedef struct {
    int nx, ny, nz;           /* grid extents, x fastest; use 1 for unused axes (FftLinearSolver_3D.c:283-301) */
    int ncomp;                /* 1 = scalar circulant, 4 = wave block-circulant (p, rho0*u, rho0*v, rho0*w) */
    int dtype;                /* enum cpc_dtype */
    int nranks, rank;         /* slab decomposition over z (1, 0 for a single GPU) */
    const void *nccl_unique_id; /* CPC_NCCL_UNIQUE_ID_BYTES bytes made by cpc_nccl_unique_id on rank 0 and
                                   broadcast by the caller; NULL when nranks == 1 */
    void *stream;             /* cudaStream_t to run on; NULL = the legacy default stream */
    int device;               /* CUDA device ordinal; -1 = current device */
} cpc_plan_desc;

#define CPC_NCCL_UNIQUE_ID_BYTES 128

/* ---- life cycle ------------------------------------------------------------------------------
 * cpc_plan_create  <- setupFFTPrec3D: MatCreateFFT + MatCreateVecsFFTW (PCSHELLFft_3D.cxx:33-37)
 * cpc_destroy      <- destroyFFTPrec3D (PCSHELLFft_3D.cxx:86-99)
 * The plan owns twiddle tables, symbol tables, staging buffers and (multi-rank) the NCCL communicator. */
int cpc_plan_create(cpc_plan *plan, const cpc_plan_desc *desc);
/* The same on a p_rows x p_cols pencil grid (p_rows * p_cols == desc->nranks; rank = c * p_rows + r) instead of z-slabs
 * <- MatCreateFFT(PETSC_COMM_WORLD, ...) (PCSHELLFft_3D.cxx:35), whose fftw-mpi distribution is slab-only: the pencil
 * grid is what BASELINE config 4 names and what lifts the limit nranks <= nz.  b and x of rank (r, c) hold the planes
 * z in slab c (of p_cols) and the rows y in slab r (of p_rows), all of x: index i + nx*(jl + nyl*kl)
 * (cpc_pencil_layout below).  Needs nx, ny divisible by p_rows and ny, nz by p_cols.  Scalar complex plans with the
 * transport / separable symbol; cpc_apply only (5 transform passes, 6 reordering passes, 4 all-to-alls within the row /
 * column groups).  With desc->nccl_unique_id == NULL and nranks > 1 the plan has no communicator of its own: the plans
 * of all ranks then live in one process and are driven together by cpc_pencil_apply_lockstep. */
int cpc_plan_create_pencil(cpc_plan *plan, const cpc_plan_desc *desc, int p_rows, int p_cols);
/* One cpc_apply on every rank of a pencil grid from a single process: plans[i] is rank i (i < p_rows * p_cols, created
 * with nccl_unique_id == NULL, on any visible devices), b[i] / x[i] its local arrays.  The exchanges are peer copies
 * between the plans' buffers; returns after they have all been queued and the streams have been synchronised. */
int cpc_pencil_apply_lockstep(cpc_plan *plans, int nplans, const void *const *b, void *const *x, int mem_kind);
int cpc_destroy(cpc_plan plan);
int cpc_set_stream(cpc_plan plan, void *stream);
int cpc_sync(cpc_plan plan);

/* ---- eigenvalue set-up (done once, held in HBM) ------------------------------------------------
 * cpc_set_symbol_transport   <- build_transport_col x3 + 1-D MatMult x3 + build_diag_mat_vec_3D
 *                               (FftLinearSolver_3D.c:80-90,136-164,218-249; PCSHELLFft_3D.cxx:51-69)
 * cpc_set_symbol_separable   <- build_diag_mat_vec_3D with caller-supplied c_x_hat, c_y_hat, c_z_hat
 *                               (FftLinearSolver_3D.c:136-164); tables are HOST complex128 arrays of nx, ny, nz
 * cpc_set_symbol_diag        <- the Diag argument of solve_3D (FftLinearSolver_3D.c:166,174): N eigenvalues,
 *                               complex128 (always, whatever the plan dtype), host or device; multi-rank plans
 *                               pass the rank's z-slab (collective).  The table is tested on the GPU for the
 *                               separable structure a[i] + b[j] + c[k] that build_diag_mat_vec_3D produces; if it
 *                               has it, three 1-D tables are kept (and the recurrence middle pass applies),
 *                               otherwise the N reciprocals are held in HBM
 * cpc_set_symbol_first_column<- general circulant: Lambda = FFT3(first column), the 1-D form of which is
 *                               tests/FFTDirectSolver/testFftSolver_1D.c:144-177; column in the plan dtype
 * cpc_set_symbol_wave        <- (absent in the reference) arrow-matrix symbol derived from
 *                               jacobianMatrices (WaveSystem.cxx:92-107); mu_d = dt/delta_d                    */
int cpc_set_symbol_transport(cpc_plan plan, double lambda_x, double lambda_y, double lambda_z);
int cpc_set_symbol_separable(cpc_plan plan, const double *cx_hat, const double *cy_hat, const double *cz_hat,
                             double lambda_x, double lambda_y, double lambda_z);
int cpc_set_symbol_diag(cpc_plan plan, const void *diag_c128, int mem_kind);
int cpc_set_symbol_first_column(cpc_plan plan, const void *column, int mem_kind);
int cpc_set_symbol_wave(cpc_plan plan, double c0, double mu_x, double mu_y, double mu_z);
/* ---- options ------------------------------------------------------------------------------------
 * Schedule switches of a plan (no counterpart in the reference, whose FFTW plan is FFTW_ESTIMATE with no knobs).
 * Defaults are what the benchmarks run; the switches exist for comparison runs and tests. */
enum cpc_option {
    CPC_OPT_Z_RECURRENCE = 1,   /* 1 (default): a transport symbol's middle pass is the cyclic recurrence along z;
                                   0: keep the fused forward-FFT / division / backward-FFT form for every symbol */
    CPC_OPT_L2_CHUNK_BYTES = 2, /* > 0: run the x / y passes z-chunk by z-chunk with this many bytes of x per chunk (Fx and
                                   Fy, By and Bx back to back on each chunk, so the second pass reads it from L2);
                                   0 (default) = whole-array passes, which measured faster at 512^3 */
    CPC_OPT_CHAIN_STREAMS = 3,  /* 1 (default) or 2: alternate the chunks' chains between two streams */
    CPC_OPT_Z_LINE_FORM = 4     /* single-rank recurrence as two thread-per-line sweeps (carry-in from the planes whose
                                   weight |c|^m can reach 1e-17, then the solve) instead of the tile kernel:
                                   1 always, 0 never, -1 (default) when the tile kernel does not fit nz or nz >= 1024
                                   and the first sweep reads less than a quarter of the array */
};
int cpc_set_option(cpc_plan plan, int option, long long value);

/* Writes the N eigenvalues currently in force (complex128) -- what the reference keeps in ctx->Diag.  Single-rank
 * plans only; cpc_build_diag_separable below serves z-slabs. */
int cpc_get_diag(cpc_plan plan, void *diag_c128, int mem_kind);
/* build_diag_mat_vec_3D (FftLinearSolver_3D.c:136-164) without a plan: planes [z0, z0 + nzl) of
 *   Diag[k,j,i] = 1 + lambda_x cx_hat[i] + lambda_y cy_hat[j] + lambda_z cz_hat[k]
 * computed on the current CUDA device from the three HOST complex128 tables and written to diag (host or device,
 * nx*ny*nzl complex128): one launch instead of the reference's 3 N VecSetValue calls. */
int cpc_build_diag_separable(int nx, int ny, int nz, const double *cx_hat, const double *cy_hat, const double *cz_hat,
                             double lambda_x, double lambda_y, double lambda_z, int z0, int nzl, void *diag_c128,
                             int mem_kind);

/* ---- the hot path -----------------------------------------------------------------------------
 * cpc_apply   <- solve_3D (FftLinearSolver_3D.c:166-190): x = (1/N) F^H( F(b) ./ Lambda ).
 *                b is read-only, x fully overwritten, b == x allowed
 *                (tests/TransportEquationFFT_SphericalExplosion_impl_mpi.cxx:111 passes Un, Un).
 *                Asynchronous on the plan's stream for device pointers; host pointers return after the
 *                result has landed in x.
 *                With the transport symbol (z table = DFT of [1,-1,0..], 0 <= lambda_z <= 4096, lambda_x, lambda_y >= 0)
 *                the z factor of F^H diag(1/Lambda) F is evaluated as the equivalent cyclic first-order recurrence
 *                (alpha + lz) x_k - lz x_{k-1} = b_k instead of two z FFTs and a division: the same operator, equal
 *                to rounding (<= 1e-13 relative); multi-rank plans then exchange one carry per (kx, ky) line instead of
 *                transposing.  cpc_set_option(CPC_OPT_Z_RECURRENCE, 0) keeps the FFT form
 *                (cpc_plan_info.fast_path[2], dist_mode).
 * cpc_forward <- MatMult(FFT_MAT, in, out)          (unnormalised, exp(-2 pi i ..), :170)
 * cpc_inverse <- MatMultTranspose(FFT_MAT, in, out) (unnormalised, exp(+2 pi i ..), :180)
 * For multi-rank plans b/x/in are the rank's z-slab [cpc_slab_range over nz]; out of cpc_forward and in of
 * cpc_inverse are in the transposed distribution (all z, the rank's y-range; index i + nx*(jloc + nyloc*k)). */
int cpc_apply(cpc_plan plan, const void *b, void *x, int mem_kind);
int cpc_forward(cpc_plan plan, const void *in, void *out, int mem_kind);
int cpc_inverse(cpc_plan plan, const void *in, void *out, int mem_kind);
/* ---- unstructured meshes: projection onto the Cartesian grid ------------------------------------------
 * cpc_set_projection  <- ctx->intersectionMatrix (reference src/PCSHELLFft_3D.hxx:17; never built there, ToDo.md
 *                        last item): CSR matrix P with N = nx*ny*nz rows (Cartesian cells) and `cols` columns
 *                        (cells of the unstructured mesh), real weights; host arrays, copied to HBM with its transpose.
 * cpc_apply_projected <- applyFFT3DPrecTransport (src/PCSHELLFft_3D.cxx:10-24): x = P^T solve_3D( P b ); b and x
 *                        have `cols` entries.  (The reference stops after solve_3D and leaves x on the Cartesian
 *                        grid; the transpose brings it back to the mesh the Krylov vectors live on.)
 * Complex single-rank plans only. */
int cpc_set_projection(cpc_plan plan, int64_t cols, const int64_t *rowptr, const int32_t *colidx, const double *val);
int cpc_apply_projected(cpc_plan plan, const void *b, void *x, int mem_kind);

/* Same as cpc_apply on device pointers, but records CUDA events around each of the passes on the plan's
 * stream and returns their durations (ms).  pass_ms must hold CPC_MAX_PASSES floats; *npasses gets the count. */
#define CPC_MAX_PASSES 16
int cpc_apply_profiled(cpc_plan plan, const void *b, void *x, float *pass_ms, int *npasses);

/* ---- introspection -----------------------------------------------------------------------------*/
typedef struct {
    int nx, ny, nz, ncomp, dtype, nranks, rank;
    int symbol_kind;
    int passes_per_apply;          /* HBM passes (kernel launches) of one cpc_apply */
    int dist_mode;                 /* 4: pencil grid (cpc_plan_create_pencil).  Otherwise:
                                      0 single rank, 1 NCCL all-to-all transposes, 2 transposes fused into the passes
                                      (stores pushed to IPC-mapped peer buffers over NVLink), 3 no transposes: the
                                      current (transport) symbol's middle pass is a recurrence along z, the z-slabs
                                      only exchange one carry per (kx, ky) line */
    int fast_path[3];              /* 1 if axis x/y/z runs the templated Stockham kernel, 0 = generic kernel;
                                      [2] == 2: the middle pass of the current (transport) symbol is solved as a cyclic
                                      first-order recurrence along z instead of forward FFT, division, backward FFT */
    int64_t local_elems;           /* elements (of ncomp * cells) held by this rank */
    int64_t bytes_per_apply_alg;   /* 5 passes x 2 x local_elems x sizeof(elem): SURVEY.md 8(d) */
    uint64_t kernel_launches;      /* running count of kernels launched by this plan */
    uint64_t h2d_bytes, d2h_bytes; /* running totals of staged copies (host-pointer calls) */
} cpc_plan_info;
int cpc_get_info(cpc_plan plan, cpc_plan_info *info);

const char *cpc_last_error(void);
int cpc_version(void);             /* major * 1000 + minor */
int cpc_device_count(void);        /* number of CUDA devices visible, 0 if none / no driver */

/* ---- pure host helpers (no GPU needed; shared by the multi-rank code and its CPU tests) ----------
 * Slab decomposition: rank r of P owns indices [start, start+count) of an axis of length n,
 * count = n/P (+1 for the first n%P ranks).                                                        */
int cpc_slab_range(int n, int nranks, int rank, int *start, int *count);
/* Element offset (within the rank-local send buffer, laid out [dest q][z_loc][y_loc(q)][x]) and element count of
 * the chunk that rank `rank` sends to rank `q` in the forward transpose of an nx x ny x nz x ncomp grid. */
int cpc_slab_send_chunk(int nx, int ny, int nz, int ncomp, int nranks, int rank, int q,
                        int64_t *offset, int64_t *count);
/* Same for the chunk received from rank `s` into the transposed buffer [z_glob][y_loc][x]. */
int cpc_slab_recv_chunk(int nx, int ny, int nz, int ncomp, int nranks, int rank, int s,
                        int64_t *offset, int64_t *count);
/* Pencil grid (cpc_plan_create_pencil): what rank (r, c) = (rank % p_rows, rank / p_rows) holds in the three
 * distributions -- X pencils [z in slab c][y in slab r][x] (b and x), Y pencils [z in slab c][y][x in slab r],
 * Z pencils [z][y in slab c (of p_cols)][x in slab r] (the middle pass). */
typedef struct {
    int r, c;
    int nxl, x0;                   /* x slab of the Y and Z pencils (over p_rows) */
    int nyl, y0;                   /* y slab of the X pencils (over p_rows) */
    int nyl2, y02;                 /* y slab of the Z pencils (over p_cols) */
    int nzl, z0;                   /* z slab of the X and Y pencils (over p_cols) */
    int64_t local_elems;
} cpc_pencil_layout_t;
int cpc_pencil_layout(int nx, int ny, int nz, int p_rows, int p_cols, int rank, cpc_pencil_layout_t *out);
/* The steps of one apply on that grid, in order (the very list the GPU plan executes).  kind: 0 x pass, 1 y pass,
 * 2 middle pass (forward z, division, backward z, scaled by 1 / (nxl nyl2 nz)), 3 SWAP: in[a][b][inner] ->
 * out[b][a][inner] times scale, 4 / 5 all-to-all of local_elems / group size elements per peer within the row /
 * column group.  dir: -1 forward, +1 backward (passes).  steps == NULL: only *nsteps is written. */
typedef struct {
    int kind, dir;
    int64_t a, b, inner;
    double scale;
    int src_buf, dst_buf;          /* which array the step reads / writes when b and x are device arrays: 0 = the caller's b,
                                      1 / 2 = the plan's two work buffers, 3 = the caller's x (passes run in place on a work
                                      buffer; b is read by the first step only, x written by the last only, so b == x works) */
    int src_buf_staged, dst_buf_staged;   /* the same for host arrays: b is copied into buffer 1 first, the result is copied
                                      out of the last step's dst afterwards */
} cpc_pencil_step_t;
int cpc_pencil_steps(int nx, int ny, int nz, int p_rows, int p_cols, int rank, cpc_pencil_step_t *steps, int max_steps,
                     int *nsteps);
/* The ranks of the row (step_kind 4) or column (5) group of `rank`, in chunk order: chunk q of the send buffer goes to
 * peers[q], chunk q of the receive buffer comes from peers[q]. */
int cpc_pencil_group(int nx, int ny, int nz, int p_rows, int p_cols, int rank, int step_kind, int *peers, int *npeers);
/* Source element of output element o of SWAP(a, b, inner) -- the index map of the reordering kernel; -1 on bad arguments. */
int64_t cpc_pencil_swap_source(int64_t o, int64_t a, int64_t b, int64_t inner);
/* The test the library applies to a separable symbol before it takes the recurrence form of the middle pass (see
 * cpc_apply): tables ax[nx], ay[ny], az[nz] as HOST complex128 arrays, already multiplied by their lambdas, the "+1"
 * of build_diag_mat_vec_3D (FftLinearSolver_3D.c:155) riding on ay.  Returns 1 and *lambda_z when az is
 * lambda_z (1 - exp(-2 pi i k / nz)) -- the DFT of build_transport_col's [1, -1, 0, ...] (:80-90) -- for some
 * 0 <= lambda_z <= 4096 and Re(ax[i] + ay[j]) >= 1/2 everywhere; 0 otherwise (the FFT form runs); -1 on bad arguments. */
int cpc_symbol_recurrence_lambda(int nx, int ny, int nz, const double *ax, const double *ay, const double *az,
                                 double *lambda_z);
/* NCCL bootstrap: fills CPC_NCCL_UNIQUE_ID_BYTES bytes (call on rank 0, broadcast out of band). */
int cpc_nccl_unique_id(void *out_bytes);

#ifdef __cplusplus
}
#endif
#endif /* CIRCULANTPC_H */
