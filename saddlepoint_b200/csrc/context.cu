// context.cu -- handle lifetime, error plumbing, per-stage CUDA-event timing, and the
// granular C ABI (analyze / assemble / factorize / solve / reductions).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "spb_internal.h"

namespace spb {

static cudaEvent_t get_event(spb_context* h) {
  if (!h->timing.pool.empty()) {
    cudaEvent_t e = h->timing.pool.back();
    h->timing.pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  SPB_CUDA(cudaEventCreate(&e));
  return e;
}

void timing_flush(spb_context* h) {
  if (h->timing.pending.empty()) return;
  SPB_CUDA(cudaStreamSynchronize(h->stream));
  for (auto& p : h->timing.pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) h->timing.ms[p.stage] += ms;
    h->timing.pool.push_back(p.a);
    h->timing.pool.push_back(p.b);
  }
  h->timing.pending.clear();
}

void stage_begin(spb_context* h, int stage) {
  if (!h->timing.enabled) return;
  if (h->timing.pending.size() >= 2048) timing_flush(h);
  Timing::Pending p;
  p.a = get_event(h);
  p.b = get_event(h);
  p.stage = stage;
  SPB_CUDA(cudaEventRecord(p.a, h->stream));
  h->timing.pending.push_back(p);
}

void stage_end(spb_context* h, int stage, int nlaunch) {
  h->launches += nlaunch;
  h->timing.count[stage] += nlaunch;
  if (!h->timing.enabled) return;
  SPB_CUDA(cudaEventRecord(h->timing.pending.back().b, h->stream));
}

}  // namespace spb

using namespace spb;

// Translates internal failures to status codes; nothing throws across the ABI.
template <class F>
static spb_status guarded(spb_handle h, F&& f) {
  try {
    cudaGetLastError();  // drop a stale (non-sticky) error left by an unrelated earlier call
    return f();
  } catch (const CudaFail& e) {
    if (h) h->err = std::string("CUDA error: ") + cudaGetErrorString(e.e) + " at " + e.where;
    return SPB_ERR_CUDA;
  } catch (const NcclFail& e) {
    if (h) h->err = std::string("NCCL error ") + std::to_string(e.code) + " at " + e.where;
    return SPB_ERR_NCCL;
  } catch (const ArgFail& e) {
    if (h) h->err = e.msg;
    return SPB_ERR_ARG;
  } catch (const StateFail& e) {
    if (h) h->err = e.msg;
    return SPB_ERR_STATE;
  } catch (const std::bad_alloc&) {
    if (h) h->err = "host allocation failed";
    return SPB_ERR_ARG;
  } catch (const std::exception& e) {
    if (h) h->err = std::string("internal error: ") + e.what();
    return SPB_ERR_STATE;
  } catch (...) {
    if (h) h->err = "internal error (unknown exception)";
    return SPB_ERR_STATE;
  }
}

// Stage a host vector into a device buffer (or pass a device pointer through).
static const double* to_device(spb_handle h, const double* p, size_t len, int mem, DevBuf<double>& stage) {
  if (!p) return nullptr;
  if (mem == SPB_DEVICE) return p;
  SPB_CUDA(stage.alloc(len));
  SPB_CUDA(cudaMemcpyAsync(stage.p, p, len * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return stage.p;
}

extern "C" {

const char* spb_version(void) { return "saddlepoint_b200 0.1 (sm_100a)"; }

spb_status spb_create(int device, void* stream, spb_handle* out) {
  if (!out) return SPB_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return SPB_ERR_CUDA;  // no CPU fallback: without a GPU there is no handle
  }
  spb_context* h = new (std::nothrow) spb_context;
  if (!h) return SPB_ERR_ARG;
  spb_status s = guarded(h, [&]() {
    h->device = device;
    SPB_CUDA(cudaSetDevice(device));
    if (stream) {
      h->stream = (cudaStream_t)stream;
    } else {
      SPB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
      h->own_stream = true;
    }
    cudaDeviceProp prop;
    SPB_CUDA(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("SPB_PCG_MULTI")) h->pcg_persistent = !(e[0] == '1');
    if (!prop.cooperativeLaunch) h->pcg_persistent = false;
    // one GPU: the textbook three-phase kernel (pcg.cu) is the default -- measured 65.5 vs 69.3 us per iteration at C2;
    // "sr" selects the single-reduction kernel (pcg_peer.cu), which the multi-GPU halo mode always uses
    if (const char* e = getenv("SPB_PCG_VARIANT")) h->pcg_single_reduction = (e[0] == 's');
    if (const char* e = getenv("SPB_PCG_MINB")) h->pcg_minb = atoi(e);
    if (const char* e = getenv("SPB_PCG_RELAX")) h->pcg_relax_factor = atof(e);
    if (const char* e = getenv("SPB_PCG_PIN_MB")) h->pcg_pin_mb = atof(e);
    if (const char* e = getenv("SPB_PCG_COL16")) h->pcg_col16 = e[0] != '0';
    SPB_CUDA(cudaMallocHost((void**)&h->h_sc, sizeof(PcgScalars)));
    SPB_CUDA(cudaMallocHost((void**)&h->h_scalar, 8 * sizeof(double)));
    memset(h->h_sc, 0, sizeof(PcgScalars));
    return SPB_OK;
  });
  if (s != SPB_OK) {
    delete h;
    return s;
  }
  *out = h;
  return SPB_OK;
}

spb_status spb_destroy(spb_handle h) {
  if (!h) return SPB_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  chol_release(h);
  ic0_release(h);
  lm_release(h);
  batch_release(h);
  dist_release(h);
  for (auto& p : h->timing.pending) {
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  for (auto e : h->timing.pool) cudaEventDestroy(e);
  if (h->copy_stream) {
    for (auto& e : h->copy_ev)
      if (e) cudaEventDestroy(e);
    if (h->asm_done_ev) cudaEventDestroy(h->asm_done_ev);
    cudaStreamDestroy(h->copy_stream);
    h->copy_stream = nullptr;
  }
  if (h->asm_side_stream) {
    cudaEventDestroy(h->asm_fork_ev);
    cudaEventDestroy(h->asm_join_ev);
    cudaStreamDestroy(h->asm_side_stream);
    h->asm_side_stream = nullptr;
  }
  if (h->h_sc) cudaFreeHost(h->h_sc);
  if (h->h_scalar) cudaFreeHost(h->h_scalar);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return SPB_OK;
}

const char* spb_last_error(spb_handle h) { return h ? h->err.c_str() : "null handle"; }

spb_status spb_set_solver(spb_handle h, int solver, double pcg_rtol, int pcg_maxit) {
  if (!h || solver < 0 || solver > 2) return SPB_ERR_ARG;
  h->solver = solver;
  if (pcg_rtol > 0) h->pcg_rtol = pcg_rtol;
  if (pcg_maxit > 0) h->pcg_maxit = pcg_maxit;
  h->factor_valid = false;
  return SPB_OK;
}

static void validate_csc(int m, int n, const int32_t* colptr, const int32_t* rowidx) {
  if (m < 0 || n < 0 || !colptr || (colptr[n] > 0 && !rowidx)) throw ArgFail{"bad pattern arguments"};
  if (colptr[0] != 0) throw ArgFail{"colptr[0] must be 0"};
  for (int j = 0; j < n; ++j)
    if (colptr[j + 1] < colptr[j]) throw ArgFail{"colptr must be non-decreasing"};
}

spb_status spb_analyze(spb_handle h, int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx) {
  if (!h) return SPB_ERR_ARG;
  return guarded(h, [&]() {
    SPB_CUDA(cudaSetDevice(h->device));
    validate_csc(m, n, colptr, rowidx);
    h->analyzed = h->have_h = h->h_valid = h->factor_valid = false;
    const bool had_plain_h = h->trip_nnz < 0 && !h->partitioned && h->h_fingerprint == 0;
    h->trip_nnz = -1;
    h->pcg_last_iters = 0;
    h->keep.valid = false;
    h->m = m;
    h->n = n;
    h->nnzJ = colptr[n];
    h->partitioned = false;
    h->m_global = m;
    const bool timing = getenv("SPB_ANALYZE_TIMING") && atoi(getenv("SPB_ANALYZE_TIMING")) > 0;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (!timing) return;
      auto t1 = std::chrono::steady_clock::now();
      fprintf(stderr, "[spb analyze timing] %s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
      t0 = t1;
    };
    AssemblyPlan P;
    build_analysis(m, n, colptr, rowidx, h->H, h->S, P);
    lap("host symbolic (pattern, SpMV layout, assembly plan)");
    {
      // The ordering / supernode plan of the Cholesky path and the colouring of IC(0) depend on the structural pattern
      // of H only.  A Jacobian whose pattern changed but whose normal-matrix pattern did not (the iterative-rounding
      // driver appends one constraint row per outer iteration, IterativeRoundingTraits.h:76-79) keeps them: that phase
      // is ~6 s at the 500k-face size against 0.8 s per LM iteration.
      const uint64_t hh = pattern_fingerprint(n, n, h->H.colptr.data(), h->H.rowidx.data());
      if (!(had_plain_h && hh == h->h_pattern_hash)) {
        chol_release(h);
        ic0_release(h);
      }
      h->h_pattern_hash = hh;
    }
    {
      // block structure: every row of H holds its diagonal and at most the entry that couples it with the other unknown
      // of its aligned pair (2k, 2k+1) -- the extended Rosenbrock structure (benchmark/rosenbrock/main.cpp:29-38)
      bool b2 = n >= 2;
      for (int r = 0; r < n && b2; ++r) {
        const int q0 = h->H.colptr[r], q1 = h->H.colptr[(size_t)r + 1];
        if (q1 - q0 > 2) b2 = false;
        for (int q = q0; q < q1 && b2; ++q)
          if (h->H.rowidx[q] != r && h->H.rowidx[q] != (r ^ 1)) b2 = false;
      }
      h->h_block2 = b2;
    }
    upload_h_layout(h);
    lap("upload H layout");
    upload_analysis(h, colptr, rowidx, P);
    SPB_CUDA(cudaStreamSynchronize(h->stream));
    lap("pack + upload assembly plan");
    h->fingerprint = pattern_fingerprint(m, n, colptr, rowidx);
    lap("fingerprint");
    h->h_fingerprint = 0;
    h->analyzed = true;
    h->have_h = true;
    return SPB_OK;
  });
}

spb_status spb_pattern_matches(spb_handle h, int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx, int* same) {
  if (!h || !same || !colptr) return SPB_ERR_ARG;
  *same = (h->analyzed && h->m == m && h->n == n && h->nnzJ == colptr[n] &&
           h->fingerprint == pattern_fingerprint(m, n, colptr, rowidx))
              ? 1
              : 0;
  return SPB_OK;
}

spb_status spb_h_nnz(spb_handle h, int64_t* nnz) {
  if (!h || !nnz) return SPB_ERR_ARG;
  if (!h->have_h) return SPB_ERR_STATE;
  *nnz = h->H.nnz();
  return SPB_OK;
}

spb_status spb_h_pattern(spb_handle h, int32_t* colptr, int32_t* rowidx) {
  if (!h) return SPB_ERR_ARG;
  if (!h->have_h) return SPB_ERR_STATE;
  if (colptr) memcpy(colptr, h->H.colptr.data(), sizeof(int32_t) * ((size_t)h->n + 1));
  if (rowidx) memcpy(rowidx, h->H.rowidx.data(), sizeof(int32_t) * (size_t)h->H.nnz());
  return SPB_OK;
}

spb_status spb_h_values(spb_handle h, double* vals, int mem) {
  if (!h || !vals) return SPB_ERR_ARG;
  return guarded(h, [&]() {
    if (!h->h_valid) throw StateFail{"no numeric H: call spb_assemble or spb_set_matrix first"};
    SPB_CUDA(cudaSetDevice(h->device));
    size_t nnz = (size_t)h->H.nnz();
    if (mem == SPB_DEVICE) {
      gather_h_values(h, vals);
    } else {
      SPB_CUDA(h->d_vec_stage.alloc(nnz));
      gather_h_values(h, h->d_vec_stage.p);
      SPB_CUDA(cudaMemcpyAsync(vals, h->d_vec_stage.p, nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    SPB_CUDA(cudaStreamSynchronize(h->stream));
    return SPB_OK;
  });
}

spb_status spb_assemble(spb_handle h, const double* Jvals, const double* EVec, double lambda, int mem, double* foo) {
  if (!h || !Jvals) return SPB_ERR_ARG;
  return guarded(h, [&]() {
    if (!h->analyzed) throw StateFail{"spb_assemble before spb_analyze"};
    SPB_CUDA(cudaSetDevice(h->device));
    if (mem == SPB_HOST && run_assembly_streamed(h, Jvals, EVec, lambda, foo)) {
      SPB_CUDA(cudaStreamSynchronize(h->stream));  // host buffers may be reused by the caller
      return SPB_OK;
    }
    const double* dJ = to_device(h, Jvals, (size_t)h->nnzJ, mem, h->d_Jv_stage);
    const double* dF = to_device(h, EVec, (size_t)h->m, mem, h->d_f_stage);
    run_assembly(h, dJ, dF, lambda, foo);
    if (mem == SPB_HOST) SPB_CUDA(cudaStreamSynchronize(h->stream));  // host buffers may be reused by the caller
    return SPB_OK;
  });
}

static spb_status copy_out(spb_handle h, const double* d_src, double* dst, size_t len, int mem) {
  return guarded(h, [&]() {
    SPB_CUDA(cudaSetDevice(h->device));
    SPB_CUDA(cudaMemcpyAsync(dst, d_src, len * sizeof(double), mem == SPB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                             h->stream));
    SPB_CUDA(cudaStreamSynchronize(h->stream));
    return SPB_OK;
  });
}
spb_status spb_get_rhs(spb_handle h, double* rhs, int mem) {
  if (!h || !rhs) return SPB_ERR_ARG;
  if (!h->h_valid || !h->analyzed) return SPB_ERR_STATE;
  return copy_out(h, h->d_rhs.p, rhs, (size_t)h->n, mem);
}
spb_status spb_get_damping(spb_handle h, double* d, int mem) {
  if (!h || !d) return SPB_ERR_ARG;
  if (!h->h_valid || !h->analyzed) return SPB_ERR_STATE;
  return copy_out(h, h->d_dvec.p, d, (size_t)h->n, mem);
}

// Installs an explicit symmetric H pattern (CSC, both triangles, full diagonal) on the handle:
// SpMV layout, diagonal positions; invalidates everything derived from a previous pattern.
static void install_explicit_pattern(spb_handle h, int32_t n, const int32_t* colptr, const int32_t* rowidx, uint64_t fp) {
  h->analyzed = h->have_h = h->h_valid = h->factor_valid = false;
  h->h_pattern_hash = 0;
  h->keep.valid = false;
  h->trip_nnz = -1;
  chol_release(h);
  ic0_release(h);
  h->n = n;
  h->m = 0;
  h->nnzJ = 0;
  HPattern& H = h->H;
  H.n = n;
  H.colptr.assign(colptr, colptr + n + 1);
  H.rowidx.assign(rowidx, rowidx + colptr[n]);
  H.nup1.assign(n, 0);
  for (int j = 0; j < n; ++j) {
    int up = 0, prev = -1;
    bool diag = false;
    for (int q = colptr[j]; q < colptr[j + 1]; ++q) {
      int i = rowidx[q];
      if (i < 0 || i >= n || i <= prev) throw ArgFail{"matrix pattern must have ascending in-range row indices"};
      prev = i;
      if (i <= j) ++up;
      if (i == j) diag = true;
    }
    if (!diag) throw ArgFail{"matrix must have a structurally full diagonal"};
    H.nup1[j] = up;
  }
  // Structural symmetry is required, not assumed: Eigen's SimplicialLLT reads the lower triangle only
  // (EigenSolverWrapper.h:58), the SpMV and the PCG here read both, so an asymmetric pattern would give
  // triangle-dependent answers.
  for (int j = 0; j < n; ++j)
    for (int q = colptr[j]; q < colptr[j + 1]; ++q) {
      const int i = rowidx[q];
      if (i == j) continue;
      if (!std::binary_search(rowidx + colptr[i], rowidx + colptr[i + 1], j)) throw ArgFail{"matrix pattern is not structurally symmetric"};
    }
  build_sell(h->H, h->S);
  upload_h_layout(h);
  h->h_fingerprint = fp;
  h->have_h = true;
}

spb_status spb_set_matrix(spb_handle h, int32_t n, const int32_t* colptr, const int32_t* rowidx, const double* vals, int mem) {
  if (!h || !vals) return SPB_ERR_ARG;
  return guarded(h, [&]() {
    SPB_CUDA(cudaSetDevice(h->device));
    validate_csc(n, n, colptr, rowidx);
    uint64_t fp = pattern_fingerprint(n, n, colptr, rowidx);
    // new explicit pattern: must be structurally symmetric with a full diagonal
    if (!(h->have_h && !h->analyzed && h->trip_nnz < 0 && h->n == n && h->h_fingerprint == fp)) install_explicit_pattern(h, n, colptr, rowidx, fp);
    size_t nnz = (size_t)h->H.nnz();
    const double* dv = to_device(h, vals, nnz, mem, h->d_vec_stage);
    scatter_h_values(h, dv);
    SPB_CUDA(cudaStreamSynchronize(h->stream));
    return SPB_OK;
  });
}

// Legacy LinearSolver concept (GNSolver.h:131,169; SLSolver.h:123,159; EigenSolverWrapper.h:118-119):
// analyze(I, J, symmetric) takes the matrix as triplets -- with symmetric != 0 only entries on or above
// the diagonal, which are mirrored -- and factorize(S, symmetric) takes one value per triplet;
// duplicates are summed in insertion order.  The symbolic work (pattern, triplet -> entry map,
// SpMV layout, and on first factorize the ordering/supernodes) is done once per pattern.
spb_status spb_analyze_triplets(spb_handle h, int32_t n, int64_t nnz, const int32_t* rows, const int32_t* cols, int symmetric) {
  if (!h || n < 0 || nnz < 0 || (nnz > 0 && (!rows || !cols)) || nnz > 0x7fffffffLL) return SPB_ERR_ARG;
  return guarded(h, [&]() {
    SPB_CUDA(cudaSetDevice(h->device));
    // entry list: (column, row, source); a mirrored copy carries the same source
    struct Ent {
      int col, row, src;
    };
    std::vector<Ent> ent;
    ent.reserve((size_t)nnz * (symmetric ? 2 : 1));
    for (int64_t t = 0; t < nnz; ++t) {
      const int i = rows[t], j = cols[t];
      if (i < 0 || i >= n || j < 0 || j >= n) throw ArgFail{"triplet index out of range"};
      if (symmetric && i > j) throw ArgFail{"symmetric triplets must lie on or above the diagonal (row <= col)"};
      ent.push_back({j, i, (int)t});
      if (symmetric && i != j) ent.push_back({i, j, (int)t});
    }
    // stable: duplicates keep their insertion order
    std::stable_sort(ent.begin(), ent.end(), [](const Ent& a, const Ent& b) { return a.col != b.col ? a.col < b.col : a.row < b.row; });
    std::vector<int32_t> colptr((size_t)n + 1, 0), rowidx;
    std::vector<int64_t> tptr;
    std::vector<int> tsrc(ent.size());
    for (size_t e = 0; e < ent.size(); ++e) {
      if (e == 0 || ent[e].col != ent[e - 1].col || ent[e].row != ent[e - 1].row) {
        rowidx.push_back(ent[e].row);
        tptr.push_back((int64_t)e);
        colptr[(size_t)ent[e].col + 1]++;
      }
      tsrc[e] = ent[e].src;
    }
    tptr.push_back((int64_t)ent.size());
    for (int j = 0; j < n; ++j) colptr[(size_t)j + 1] += colptr[j];
    if (!symmetric) {  // the solvers need a structurally symmetric pattern
      for (int j = 0; j < n; ++j)
        for (int q = colptr[j]; q < colptr[(size_t)j + 1]; ++q) {
          const int i = rowidx[q];
          if (i == j) continue;
          if (!std::binary_search(rowidx.begin() + colptr[i], rowidx.begin() + colptr[(size_t)i + 1], j))
            throw ArgFail{"triplet pattern is not structurally symmetric"};
        }
    }
    install_explicit_pattern(h, n, colptr.data(), rowidx.data(), pattern_fingerprint(n, n, colptr.data(), rowidx.data()));
    SPB_CUDA(h->d_trip_ptr.upload(tptr, h->stream));
    SPB_CUDA(h->d_trip_src.upload(tsrc, h->stream));
    SPB_CUDA(cudaStreamSynchronize(h->stream));
    h->trip_nnz = nnz;
    return SPB_OK;
  });
}

spb_status spb_factorize_triplets(spb_handle h, const double* vals, int mem) {
  if (!h || !vals) return SPB_ERR_ARG;
  spb_status st = guarded(h, [&]() {
    if (h->trip_nnz < 0 || !h->have_h) throw StateFail{"spb_factorize_triplets before spb_analyze_triplets"};
    SPB_CUDA(cudaSetDevice(h->device));
    const double* dv = to_device(h, vals, (size_t)h->trip_nnz, mem, h->d_vec_stage2);
    SPB_CUDA(h->d_vec_stage.alloc((size_t)h->H.nnz()));
    sum_triplets(h, dv, h->d_vec_stage.p);
    scatter_h_values(h, h->d_vec_stage.p);
    if (mem == SPB_HOST) SPB_CUDA(cudaStreamSynchronize(h->stream));
    return SPB_OK;
  });
  if (st != SPB_OK) return st;
  return spb_factorize(h);
}

spb_status spb_factorize(spb_handle h) {
  if (!h) return SPB_ERR_ARG;
  return guarded(h, [&]() {
    if (!h->h_valid) throw StateFail{"spb_factorize without a numeric H"};
    SPB_CUDA(cudaSetDevice(h->device));
    spb_status s = SPB_OK;
    if (h->solver == SPB_SOLVER_CHOLESKY) s = chol_factorize(h);
    else if (h->solver == SPB_SOLVER_PCG_IC0) s = ic0_factorize(h);
    // PCG: the Jacobi diagonal is a by-product of assembly; "factorisation" is its inverse.
    else dev_invert_diag(h);
    if (s == SPB_OK) h->factor_valid = true;
    return s;
  });
}

spb_status spb_solve(spb_handle h, const double* rhs, double* x, int nrhs, int mem, int* iters) {
  if (!h || !x || nrhs < 1) return SPB_ERR_ARG;
  return guarded(h, [&]() {
    if (!h->factor_valid) throw StateFail{"spb_solve before spb_factorize"};
    h->solve_xx_valid = false;
    SPB_CUDA(cudaSetDevice(h->device));
    const size_t n = (size_t)h->n;
    int total_iters = 0;
    spb_status worst = SPB_OK;
    for (int c = 0; c < nrhs; ++c) {
      const double* db;
      if (!rhs) {
        if (!h->analyzed) throw StateFail{"no assembled gradient to solve for"};
        db = h->d_rhs.p;
      } else {
        db = to_device(h, rhs + (size_t)c * n, n, mem, h->d_vec_stage);
      }
      double* dx = x + (size_t)c * n;
      if (mem == SPB_HOST) {
        SPB_CUDA(h->d_x.alloc(n));
        dx = h->d_x.p;
      }
      int it = 0;
      if (h->halo_mode && h->solver != SPB_SOLVER_PCG_JACOBI) throw StateFail{"subdomain (halo) mode supports SPB_SOLVER_PCG_JACOBI only"};
      spb_status s = h->halo_mode ? ((h->peer && h->peer->ready) ? pcg_solve_halo_peer(h, db, dx, &it) : pcg_solve_halo(h, db, dx, &it))
                     : h->solver == SPB_SOLVER_CHOLESKY ? chol_solve(h, db, dx)
                     : h->solver == SPB_SOLVER_PCG_IC0 ? pcg_ic0_solve(h, db, dx, &it)
                     : (h->nranks > 1 ? pcg_solve_dist(h, db, dx, &it) : pcg_solve(h, db, dx, &it));
      total_iters += it;
      if (s != SPB_OK && worst == SPB_OK) worst = s;
      if (mem == SPB_HOST) {
        SPB_CUDA(cudaMemcpyAsync(x + (size_t)c * n, dx, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        SPB_CUDA(cudaStreamSynchronize(h->stream));
      }
    }
    if (iters) *iters = total_iters;
    SPB_CUDA(cudaStreamSynchronize(h->stream));
    return worst;
  });
}

spb_status spb_squared_norm(spb_handle h, const double* v, int64_t len, int mem, double* out) {
  if (!h || !out || (len > 0 && !v)) return SPB_ERR_ARG;
  return guarded(h, [&]() {
    SPB_CUDA(cudaSetDevice(h->device));
    const double* dv = to_device(h, v, (size_t)len, mem, h->d_vec_stage);
    dev_squared_norm(h, dv, len, out);
    return SPB_OK;
  });
}
spb_status spb_inf_norm(spb_handle h, const double* v, int64_t len, int mem, double* out) {
  if (!h || !out || (len > 0 && !v)) return SPB_ERR_ARG;
  return guarded(h, [&]() {
    SPB_CUDA(cudaSetDevice(h->device));
    const double* dv = to_device(h, v, (size_t)len, mem, h->d_vec_stage);
    dev_inf_norm(h, dv, len, out);
    return SPB_OK;
  });
}

spb_status spb_spmv(spb_handle h, const double* x, double* y, int mem) {
  if (!h || !x || !y) return SPB_ERR_ARG;
  return guarded(h, [&]() {
    if (!h->h_valid) throw StateFail{"spb_spmv without a numeric H"};
    SPB_CUDA(cudaSetDevice(h->device));
    const size_t n = (size_t)h->n;
    const double* dx = to_device(h, x, n, mem, h->d_vec_stage);
    double* dy = y;
    if (mem == SPB_HOST) {
      SPB_CUDA(h->d_vec_stage2.alloc(n));
      dy = h->d_vec_stage2.p;
    }
    dev_spmv(h, dx, dy);
    if (mem == SPB_HOST) SPB_CUDA(cudaMemcpyAsync(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    SPB_CUDA(cudaStreamSynchronize(h->stream));
    return SPB_OK;
  });
}

// ||H x - rhs|| / ||rhs|| for the assembled gradient rhs and the current H: the check a caller (bench.py, tests)
// runs on a step returned by spb_solve.  Subdomain (halo) mode: x carries valid ghost copies (as spb_solve
// returns it), the product is completed across ranks by one halo exchange and both norms are sums over the owned
// unknowns of all ranks, so every rank gets the same global number.
namespace {
__global__ void residual_diff_kernel(const double* __restrict__ y, const double* __restrict__ b, double* __restrict__ out, int64_t len) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) out[i] = y[i] - b[i];
}
}  // namespace
spb_status spb_residual_check(spb_handle h, const double* x, int mem, double* relres) {
  if (!h || !x || !relres) return SPB_ERR_ARG;
  return guarded(h, [&]() {
    if (!h->h_valid || !h->analyzed) throw StateFail{"spb_residual_check needs an assembled system"};
    SPB_CUDA(cudaSetDevice(h->device));
    const size_t n = (size_t)h->n;
    const double* dx = to_device(h, x, n, mem, h->d_vec_stage);
    SPB_CUDA(h->d_vec_stage2.alloc(n));
    double* dy = h->d_vec_stage2.p;
    dev_spmv(h, dx, dy);
    int lo = 0, hi = h->n;
    if (h->halo_mode) {
      halo_sum_both(h, dy);
      lo = h->own_lo;
      hi = h->own_hi;
    }
    const int64_t len = hi - lo;
    if (len > 0) {
      residual_diff_kernel<<<(int)std::min<int64_t>((len + 255) / 256, (int64_t)h->num_sms * 8), 256, 0, h->stream>>>(dy + lo, h->d_rhs.p + lo, dy + lo, len);
      h->launches++;
      SPB_CUDA(cudaGetLastError());
    }
    double rr = 0.0, bb = 0.0;
    if (h->halo_mode) {
      dev_squared_norm_global(h, dy + lo, len, &rr);
      dev_squared_norm_global(h, h->d_rhs.p + lo, len, &bb);
    } else {
      dev_squared_norm(h, dy + lo, len, &rr);
      dev_squared_norm(h, h->d_rhs.p + lo, len, &bb);
    }
    *relres = bb > 0.0 ? std::sqrt(rr / bb) : std::sqrt(rr);
    return SPB_OK;
  });
}

spb_status spb_factor_info(spb_handle h, int64_t* nnzL, double* flops, int32_t* nsuper, int32_t* nlevels) {
  if (!h) return SPB_ERR_ARG;
  int a = 0, b = 0;
  if (!chol_info(h, nnzL, flops, &a, &b)) return SPB_ERR_STATE;
  if (nsuper) *nsuper = a;
  if (nlevels) *nlevels = b;
  return SPB_OK;
}

int64_t spb_launch_count(spb_handle h) { return h ? h->launches : 0; }
int spb_assembly_form(spb_handle h) { return h ? h->last_asm_form : 0; }

spb_status spb_stage_ms(spb_handle h, int stage, double* ms, int64_t* launches) {
  if (!h || stage < 0 || stage >= ST_COUNT) return SPB_ERR_ARG;
  return guarded(h, [&]() {
    timing_flush(h);
    if (ms) *ms = h->timing.ms[stage];
    if (launches) *launches = h->timing.count[stage];
    return SPB_OK;
  });
}

spb_status spb_stage_reset(spb_handle h, int enable_timing) {
  if (!h) return SPB_ERR_ARG;
  return guarded(h, [&]() {
    timing_flush(h);
    for (int s = 0; s < ST_COUNT; ++s) {
      h->timing.ms[s] = 0;
      h->timing.count[s] = 0;
    }
    h->timing.enabled = enable_timing != 0;
    return SPB_OK;
  });
}

}  // extern "C"
