/* saddlepoint_b200.h -- C ABI of the B200-native Levenberg-Marquardt inner step.
 *
 * This is the drop-in boundary underneath SaddlePoint's header-only concept API.  The
 * reference has no FFI of its own (its plug-in mechanism is C++ template duck typing,
 * LMSolver.h:22); every entry point below names the reference call site it replaces.  The
 * concept-conforming C++ classes that sit on top of this ABI are in
 * include/saddlepoint_b200/ (LMSolver, DiagonalDamping, B200SolverWrapper).
 *
 * Conventions: plain pointers and sizes only; `mem` says whether value pointers are host
 * (SPB_HOST) or device (SPB_DEVICE) memory; pattern arrays are always host memory; indices
 * are int32 and values FP64 exactly like Eigen::SparseMatrix<double> (column-major CSC,
 * sorted inner indices, explicit zeros kept).  Every call returns an spb_status; no
 * exceptions cross the ABI.  One host thread per handle; all device work is ordered on the
 * handle's stream.  There is no CPU fallback: without a CUDA device spb_create fails.
 */
#ifndef SADDLEPOINT_B200_H
#define SADDLEPOINT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spb_context* spb_handle;

typedef enum {
  SPB_OK = 0,
  SPB_NOT_SPD = 1,        /* factorisation met a pivot <= 0  (EigenSolverWrapper.h:59 -> false) */
  SPB_ERR_ARG = 2,
  SPB_ERR_CUDA = 3,
  SPB_ERR_STATE = 4,      /* call order violated (e.g. assemble before analyze)            */
  SPB_NOT_CONVERGED = 5,  /* PCG hit maxit                                                 */
  SPB_ERR_NCCL = 6
} spb_status;

typedef enum { SPB_HOST = 0, SPB_DEVICE = 1 } spb_mem;

typedef enum {
  SPB_SOLVER_CHOLESKY = 0,   /* supernodal LL^T: the SimplicialLLT replacement            */
  SPB_SOLVER_PCG_JACOBI = 1, /* Jacobi-preconditioned CG on the same SpMV                 */
  SPB_SOLVER_PCG_IC0 = 2     /* IC(0)-preconditioned CG                                   */
} spb_solver;

/* ---- lifetime ------------------------------------------------------------------------ */
/* stream: a cudaStream_t (or NULL for a library-owned stream). */
spb_status spb_create(int device, void* stream, spb_handle* out);
spb_status spb_destroy(spb_handle h);
const char* spb_last_error(spb_handle h);
/* version / build info; callable without a GPU */
const char* spb_version(void);

/* ---- solver selection ------------------------------------------------------------------ */
spb_status spb_set_solver(spb_handle h, int solver /* spb_solver */, double pcg_rtol, int pcg_maxit);

/* ---- symbolic analysis (one-time per pattern) ------------------------------------------ */
/* Replaces EigenSolverWrapper::analyze(JPattern) (EigenSolverWrapper.h:30-52; note the
 * reference never calls it, LMSolver.h:69-73, and re-derives the pattern every iteration
 * inside the product at LMSolver.h:136).  Input: pattern of the m x n Jacobian in CSC.
 * Builds: the structural pattern of dampJ^T*dampJ (full symmetric, sorted, full diagonal),
 * the map from every H nonzero to its contributing J entry pairs in ascending residual-row
 * order, the SpMV layout, and (lazily, on first Cholesky use) ordering/etree/supernodes. */
spb_status spb_analyze(spb_handle h, int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx);
/* Cheap re-validation (the seamless-integration driver grows J between solves):
 * returns SPB_OK and *same=1 if (m,n,colptr,rowidx) matches the analysed pattern. */
spb_status spb_pattern_matches(spb_handle h, int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx, int* same);
/* Incremental re-analysis (benchmark/seamless_integration/main.cpp:69-80: the iterative-rounding driver re-inits
 * the LM solver after IterativeRoundingTraits.h:76-79 appended ONE constraint row, i.e. J gained one residual row
 * with a single nonzero and the rows below it moved down by one).  Same arguments as spb_analyze.  *how (may be
 * NULL): 0 = the pattern is the analysed one, nothing done; 1 = the analysed state was patched (H pattern, SpMV
 * layout, Cholesky / IC(0) symbolic state kept; J pattern, pair records and stage layouts updated); 2 = analysed
 * from scratch (any other change).  The resulting state is identical to spb_analyze's. */
spb_status spb_analyze_update(spb_handle h, int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx, int* how);
/* H pattern export (host), for parity tests: colptr[n+1], rowidx[nnz]. */
spb_status spb_h_nnz(spb_handle h, int64_t* nnz);
spb_status spb_h_pattern(spb_handle h, int32_t* colptr, int32_t* rowidx);
/* H values in the CSC order of spb_h_pattern. */
spb_status spb_h_values(spb_handle h, double* vals, int mem);

/* Host-only views of the symbolic phase (no GPU, no handle): the structural pattern, and a
 * self-check of the assembly plan (stats[8]: nslots, npairs, nchunks, rec_bits, padded SpMV
 * entries, nnz(H), ncolchunks, invariant violations).  Used by the CPU test tier. */
spb_status spb_symbolic_h_pattern(int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx, int64_t* nnz,
                                  int32_t* h_colptr, int32_t* h_rowidx);
spb_status spb_symbolic_plan_check(int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx, int64_t* stats);
/* Host execution of the packed "stream" form of the pair assembly (what assemble_pairs_stream_kernel reads), in the
 * kernel's own summation order: off-diagonal values of H in the CSC order of spb_symbolic_h_pattern (diagonal entries
 * 0).  stats[6] = form applicable, chunks, section bytes, runs, largest value window, largest section. */
spb_status spb_symbolic_pairs_v2(int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx, const double* Jvals, double* hvals,
                                 int64_t* stats);
/* Host-only view of the Cholesky symbolic phase (nested-dissection order, supernodes, fronts):
 * stats[8] = nsuper, nlevels, nnz(L), max pivot block, max front, flops, update-matrix entries,
 * structural violations; perm (optional, n) receives the fill-reducing order. */
spb_status spb_symbolic_chol_stats(int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx, int64_t* stats, int32_t* perm);

/* ---- numeric assembly ------------------------------------------------------------------- */
/* One call replaces, for the J / EVec of the last computeJacobian=true evaluation:
 *   rhs = -(J^T * EVec)                          LMSolver.h:117
 *   fooOptimality = rhs.lpNorm<Infinity>()       LMSolver.h:119
 *   dampVector(j) = sum_r (lambda*v)*v, sqrt, augmented rows   DiagonalDamping.h:31-43,73-85
 *   dampJ^T * dampJ                              LMSolver.h:136
 * Jvals: nnz(J) values in the CSC order given to spb_analyze; EVec: m values (may be NULL to
 * skip the gradient); foo: host out (may be NULL).  H and rhs stay on the device. */
spb_status spb_assemble(spb_handle h, const double* Jvals, const double* EVec, double lambda, int mem, double* foo);
/* copies of the assembled gradient rhs (n) and the damping vector d (n) */
spb_status spb_get_rhs(spb_handle h, double* rhs, int mem);
spb_status spb_get_damping(spb_handle h, double* d, int mem);

/* ---- LinearSolver concept ---------------------------------------------------------------- */
/* Load an explicit symmetric matrix (both triangles, CSC) instead of assembling it: this is
 * LinearSolver::factorize(const SparseMatrix<double>) (EigenSolverWrapper.h:54-60) for
 * callers that keep the reference's LMSolver.h unchanged and pass dampJ^T*dampJ themselves.
 * The pattern is analysed on first use and whenever it changes. */
spb_status spb_set_matrix(spb_handle h, int32_t n, const int32_t* colptr, const int32_t* rowidx, const double* vals, int mem);
/* Legacy LinearSolver concept used by GNSolver / SLSolver / EigenSingleSolveWrapper:
 * LS->analyze(HRows, HCols[, symmetric]) (GNSolver.h:131, SLSolver.h:123, EigenSolverWrapper.h:118)
 * and LS->factorize(HVals[, symmetric]) (GNSolver.h:169, SLSolver.h:159, EigenSolverWrapper.h:119).
 * The matrix is a triplet list (duplicates are summed in insertion order); symmetric != 0 means
 * only entries with row <= col are listed and are mirrored.  spb_analyze_triplets does all
 * pattern work once; spb_factorize_triplets takes one value per triplet, builds H on the
 * device and factorises it (SPB_NOT_SPD <=> the reference's `false`); spb_solve follows. */
spb_status spb_analyze_triplets(spb_handle h, int32_t n, int64_t nnz, const int32_t* rows, const int32_t* cols, int symmetric);
spb_status spb_factorize_triplets(spb_handle h, const double* vals, int mem);
/* Replaces solver.compute(A) + info()==Success (EigenSolverWrapper.h:58-59): numeric
 * factorisation of the current H.  SPB_NOT_SPD <=> the reference's `false`. */
spb_status spb_factorize(spb_handle h);
/* Replaces x = solver.solve(rhs) (EigenSolverWrapper.h:76; matrix RHS :62-70, nrhs columns,
 * column-major).  rhs==NULL means "use the assembled gradient". iters: PCG iterations or 0. */
spb_status spb_solve(spb_handle h, const double* rhs, double* x, int nrhs, int mem, int* iters);

/* ||H x - rhs|| / ||rhs|| with the assembled gradient and the current H, for checking a step returned by
 * spb_solve (LMSolver.h:143).  Subdomain (halo) mode: a global number, identical on every rank. */
spb_status spb_residual_check(spb_handle h, const double* x, int mem, double* relres);

/* Size of the current Cholesky factor (after the first spb_factorize with SPB_SOLVER_CHOLESKY):
 * nnz(L) incl. supernodal fill, factorisation flops, supernode and tree-level counts. */
spb_status spb_factor_info(spb_handle h, int64_t* nnzL, double* flops, int32_t* nsuper, int32_t* nlevels);

/* ---- vector reductions used by the LM loop (LMSolver.h:119,146,155,157,175) ------------- */
spb_status spb_squared_norm(spb_handle h, const double* v, int64_t len, int mem, double* out);
spb_status spb_inf_norm(spb_handle h, const double* v, int64_t len, int mem, double* out);
/* y = H * x with the current H (the SpMV the PCG path is built on); exported for tests. */
spb_status spb_spmv(spb_handle h, const double* x, double* y, int mem);

/* ---- whole LM loop ---------------------------------------------------------------------- */
/* SolverTraits concept (LMSolver.h:75-80,90,106,111,181) as C callbacks.  With
 * mem==SPB_DEVICE the x / EVec / Jvals pointers are device pointers and the callback must
 * enqueue its work on `stream`; with SPB_HOST they are host pointers. */
typedef struct {
  int32_t xSize, ESize;
  const int32_t* colptr; /* CSC pattern of J (host), fixed for the duration of a solve */
  const int32_t* rowidx;
  uint64_t pattern_version; /* 0: unknown (the library fingerprints the arrays on every solve);
                               nonzero: caller's promise that the pattern is unchanged while this
                               value and the colptr pointer are unchanged                         */
  int32_t row_begin, row_end; /* multi-GPU, one big problem: row_end>row_begin means this rank owns
                                 residual rows [row_begin,row_end) of the GLOBAL Jacobian described by
                                 ESize/colptr/rowidx; objective_jacobian then fills only the owned slice
                                 of EVec and the owned entries of Jvals (global CSC order restricted to
                                 the owned rows).  0,0: the whole problem (single GPU).            */
  int mem;
  void* user;
  void (*initial_solution)(void* user, double* x0_host);
  /* Jvals==NULL <=> computeJacobian=false */
  void (*objective_jacobian)(void* user, const double* x, double* EVec, double* Jvals, void* stream);
  void (*pre_iteration)(void* user, const double* x);  /* may be NULL */
  int (*post_iteration)(void* user, const double* x);  /* may be NULL; nonzero => stop   */
} spb_traits;

typedef struct {
  int32_t maxIterations;   /* LMSolver::init default 100 (LMSolver.h:58)                  */
  double funcTolerance;    /* default 10e-10 (LMSolver.h:59)                              */
  double fooTolerance;     /* default 10e-10 (LMSolver.h:60)                              */
  double lambda0;          /* DiagonalDamping ctor default 0.01 (DiagonalDamping.h:93)    */
  int32_t verbose;         /* also selects the first-order-optimality break (LMSolver.h:125) */
  int32_t dedupe_evaluations; /* 0: the reference's exact 5 objective calls per iteration;
                                 1: reuse the bit-identical residuals (2 calls)           */
  int32_t fixed_iterations;   /* >0: benchmark mode, run exactly this many loop bodies with
                                 the stop tests disabled                                   */
} spb_lm_options;

typedef struct {
  int32_t ret;             /* the bool LMSolver::solve returns                            */
  int32_t exit_path;       /* 1 foo 2 factorize failed 3 direction small 4 func-tol 5 traits stop 6 max-iter */
  int32_t currIter;
  int32_t factorizations;
  double energy, fooOptimality, lambda;
  int64_t linear_iterations; /* total PCG iterations                                      */
  double seconds_total, seconds_analyze;
  int64_t unconverged_solves; /* PCG solves that hit maxit (the inexact direction is still used) */
} spb_lm_result;

void spb_lm_default_options(spb_lm_options* o);
/* x_out: xSize values (host or device per x_mem).  trace (optional, host, 8 doubles per loop
 * body that reached the solve, at most trace_cap records): prevEnergy2,newEnergy2,lambda,
 * foo,dirnorm,accepted,lin_iters,linear solve not converged (0/1). */
spb_status spb_lm_solve(spb_handle h, const spb_traits* traits, const spb_lm_options* opt, spb_lm_result* res,
                        double* x_out, int x_mem, double* trace, int32_t trace_cap, int32_t* trace_len);

typedef struct spb_problem* spb_problem_t;
/* ---- many independent problems at once (config C4) ---------------------------------------------------------- */
/* `traits` describes ONE block-diagonal problem made of `nproblems` independent problems of equal size: problem p owns
 * unknowns [p*xSize/nproblems, (p+1)*xSize/nproblems) and residual rows [p*ESize/nproblems, ...); its Jacobian block may
 * have its own pattern.  Every problem runs LMSolver::solve's loop (LMSolver.h:86-191, with dedupe_evaluations semantics)
 * with its own lambda, acceptance, stop tests and exit path, but every stage is one launch over all problems: one fused
 * assembly, one persistent PCG kernel with per-problem dot products and convergence, a device-side controller; finished
 * problems are frozen while the others continue.  results: nproblems records; x_out: xSize values.  Jacobi-PCG only. */
spb_status spb_lm_solve_batch(spb_handle h, const spb_traits* traits, int32_t nproblems, const spb_lm_options* opt, spb_lm_result* results,
                              double* x_out, int x_mem);
/* nproblems LF problems (seeds seed, seed+seed_stride, ...) as one block-diagonal problem for spb_lm_solve_batch */
spb_status spb_problem_lf_batch(spb_handle h, int32_t g, int32_t rows_mult, uint64_t seed, int32_t seed_stride, int32_t nproblems, spb_problem_t* out);

/* ---- check_traits (check_traits.h:19-83) on the GPU -------------------------------------- */
/* Compares the traits' Jacobian at x with central finite differences (step 10e-5, entries with
 * |FE| <= 10e-7 dropped, discrepancies counted above 10e-7 -- the reference's constants when the
 * arguments are 0).  The reference perturbs one unknown at a time (2*xSize evaluations); here
 * columns that share no residual row are perturbed together, 2*ncolors evaluations.  A response in
 * a row that has no entry in any perturbed column is a derivative missing from the declared
 * pattern and is reported separately (outside_*). */
typedef struct {
  double max_diff;        /* max |FE - J| over the pattern ("Maximum gradient difference")      */
  int32_t max_row, max_col; /* its location (first in column-major order)                        */
  int64_t discrepancies;  /* entries with |FE - J| > tol                                        */
  double outside_max;     /* largest |FE| response outside the declared pattern                 */
  int64_t outside_count;  /* such responses above tol (row, colour) pairs                       */
  int32_t ncolors;
  int32_t evaluations;    /* objective evaluations used                                          */
} spb_check_result;
/* x: xSize values (mem); fe_vals (optional, host): nnz(J) finite-difference values in CSC order. */
spb_status spb_check_traits(spb_handle h, const spb_traits* traits, const double* x, int mem, double step, double tol,
                            spb_check_result* out, double* fe_vals);

/* ---- built-in device-resident benchmark traits --------------------------- */
/* "LF": sparse analogue of benchmark/linear_function: g x g torus, m=rows_mult*n,
 * 8 nnz/row, counter-based splitmix64 data, f = A x - 1, x0 = 1. */
spb_status spb_problem_lf(spb_handle h, int32_t g, int32_t rows_mult, uint64_t seed, spb_problem_t* out);
/* Extended Rosenbrock of benchmark/rosenbrock/main.cpp:11,47-62 over nblocks blocks;
 * x0: 2*nblocks host values or NULL for (-100,100) repeated (main.cpp:43). */
spb_status spb_problem_rosenbrock(spb_handle h, int32_t nblocks, const double* x0, spb_problem_t* out);
/* The LOCAL part of the LF problem for spb_dist_ring_halo: unknowns [col_begin, col_end) of the torus plus a
 * halo of 2g on either side, and the residual rows based at the owned unknowns (same data as spb_problem_lf). */
spb_status spb_problem_lf_cols(spb_handle h, int32_t g, int32_t rows_mult, uint64_t seed, int32_t col_begin, int32_t col_end,
                               spb_problem_t* out);
/* Block structure of the seamless-integration Jacobian (SIInitialSolutionTraits.h:115-248):
 * rows [0,m_lin): E = A x - b with a CONSTANT sparse A (n columns, CSC, host; the integration, closeness
 * and constraint blocks with their weights folded in); rows [m_lin, m_lin+nbar): the injectivity barrier
 * w_bar*(1/bar(t)-1), t = ((c.x n.y - c.y n.x)/vol)/s, of the four unknowns bar_idx[4r..4r+3] =
 * (c.x, c.y, n.x, n.y) (:148-172), whose Jacobian entries (:216-233; 0, finite or inf) are the only values
 * recomputed per evaluation.  x0: n start values.  The Jacobian pattern is the row-wise stack, CSC. */
spb_status spb_problem_blocks(spb_handle h, int32_t n, int32_t m_lin, const int32_t* A_colptr, const int32_t* A_rowidx, const double* A_vals,
                              const double* b, int32_t nbar, const int32_t* bar_idx, const double* bar_vol, double s, double w_bar,
                              const double* x0, spb_problem_t* out);
spb_status spb_problem_traits(spb_problem_t p, spb_traits* out);
spb_status spb_problem_destroy(spb_problem_t p);

/* ---- multi-GPU: one large problem, residual rows partitioned across ranks (config C5) ------ */
/* One process per GPU.  The reference is single-process; this is where the path shards
 * naturally: J^T J = sum_g J_g^T J_g, J^T f = sum_g J_g^T f_g.  spb_comm_unique_id is called on
 * one rank and the 128 bytes are distributed by the caller (e.g. torch.distributed broadcast). */
spb_status spb_comm_unique_id(void* id128);
spb_status spb_comm_init(spb_handle h, const void* id128, int rank, int nranks);
spb_status spb_comm_destroy(spb_handle h);
/* Subdomain (halo) mode -- the scalable form: the handle holds only the LOCAL problem of its rank (analyse /
 * assemble / solve with the local Jacobian whose columns are [halo ghosts of the left neighbour | owned
 * unknowns | halo ghosts of the right neighbour], ranks on a ring; suits banded problems partitioned by
 * unknown).  Nothing of size n is communicated: per PCG iteration one halo slab of p forward, one of q back
 * (added on the owners in a fixed order) and two scalar allreduces; assembly exchanges the halo parts of
 * rhs, d and diag(H).  own_begin must equal halo, own_end - own_begin >= 2*halo; halo = 0 switches it off.
 * spb_assemble / spb_factorize / spb_solve / spb_lm_solve then work on local vectors (ghost entries of
 * results are kept consistent); only SPB_SOLVER_PCG_JACOBI is supported in this mode. */
spb_status spb_dist_ring_halo(spb_handle h, int32_t own_begin, int32_t own_end, int32_t halo, int32_t left_rank, int32_t right_rank);
/* *ready = 1 if the halo mode's peer-memory exchange is in use: every rank mapped every rank's window (CUDA IPC over
 * NVLink), so spb_solve runs the PCG as one persistent kernel per rank with the halo slabs and the dot products
 * exchanged inside the kernel; 0: the ncclSend/ncclRecv + ncclAllReduce path (no peer access, or SPB_HALO_PEER=0). */
spb_status spb_dist_peer_ready(spb_handle h, int* ready);
/* Same GLOBAL Jacobian pattern on every rank; this rank owns residual rows [row_begin,row_end).
 * Afterwards spb_assemble takes the owned values only (global CSC order restricted to the owned
 * rows; see spb_local_pattern) and the owned slice of EVec; the partial H / rhs / d of all ranks
 * are combined with one grouped ncclAllReduce, and spb_solve (PCG) runs the distributed
 * conjugate gradient (NCCL allreduce of the dot products, allgather of the search direction). */
spb_status spb_analyze_partitioned(spb_handle h, int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx,
                                   int32_t row_begin, int32_t row_end);
spb_status spb_local_nnz(spb_handle h, int64_t* nnz);
spb_status spb_local_pattern(spb_handle h, int32_t* colptr, int32_t* rowidx); /* rows renumbered from 0 */
int64_t spb_collective_count(spb_handle h);
/* LF benchmark traits restricted to residual rows [row_begin,row_end) (global pattern, local values) */
spb_status spb_problem_lf_rows(spb_handle h, int32_t g, int32_t rows_mult, uint64_t seed, int32_t row_begin, int32_t row_end,
                               spb_problem_t* out);

/* ---- instrumentation --------------------------------------------------------------------- */
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t spb_launch_count(spb_handle h);
/* which pair-assembly kernel the LAST spb_assemble on this handle ran: 3 = stream form (assemble_pairs_stream_kernel:
 * warp-specialised, TMA-fed, packed sections), 2 = register-pipelined general kernel, 1 = one-CTA-per-chunk general
 * kernel, 0 = none yet / no off-diagonal entries.  The general kernels take over when the packed form does not apply
 * (J columns longer than 128 entries, H rows longer than 256, a misaligned value pointer, after spb_analyze_update). */
int spb_assembly_form(spb_handle h);
/* per-stage device time accumulated since the last reset, milliseconds (CUDA events):
 * stage 0 assembly, 1 spmv, 2 pcg-vector, 3 factor, 4 trisolve, 5 reductions, 6 traits,
 * 7 fused persistent PCG solve */
spb_status spb_stage_ms(spb_handle h, int stage, double* ms, int64_t* launches);
spb_status spb_stage_reset(spb_handle h, int enable_timing);

#ifdef __cplusplus
}
#endif
#endif
