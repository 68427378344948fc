// spb_internal.h -- shared declarations of the library's translation units (not public).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <exception>
#include <string>
#include <thread>
#include <vector>

#include "../../include/saddlepoint_b200.h"

namespace spb {

// ---- device buffer --------------------------------------------------------------------
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    if (count <= n && p) return cudaSuccess;
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  cudaError_t upload(const std::vector<T>& h, cudaStream_t s) {
    cudaError_t e = alloc(h.size());
    if (e != cudaSuccess || h.empty()) return e;
    return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};

struct CudaFail {
  cudaError_t e;
  const char* where;
};
#define SPB_CUDA(expr)                                     \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) throw spb::CudaFail{_e, #expr}; \
  } while (0)

struct ArgFail {
  std::string msg;
};
struct NcclFail {
  int code;
  const char* where;
};
struct StateFail {
  std::string msg;
};

// ---- host symbolic products ---------------------------------------------------------
// Full symmetric structural pattern of dampJ^T*dampJ (LMSolver.h:136), CSC == CSR.
struct HPattern {
  int n = 0;
  std::vector<int> colptr, rowidx;
  std::vector<int> nup1;  // per column: number of entries with row <= column (diag rank + 1)
  int64_t nnz() const { return colptr.empty() ? 0 : colptr[n]; }
};

// SpMV storage of H: sliced ELL, slice height 32, column-major inside a slice; row r of the
// symmetric H is column r of the CSC pattern; entry k of row r sits at
// slice_off[r>>5] + k*32 + (r&31).  Padding entries have value 0 and are never read
// (rowlen guards them).
struct SellLayout {
  int n = 0, nslices = 0;
  std::vector<int64_t> slice_off;  // nslices+1, in entries
  std::vector<int> col;            // padded column indices
  std::vector<uint16_t> col16;     // column - row + 32768 where it fits (compressed index stream)
  std::vector<uint8_t> slice16;    // per slice: every delta of the slice fits 16 bits
  std::vector<int> rowlen;         // n
  int64_t padded = 0;
  int64_t pos(int row, int k) const { return slice_off[row >> 5] + (int64_t)k * 32 + (row & 31); }
};

// Assembly plan; record formats are documented in assembly.cu.
struct AssemblyPlan {
  int m = 0, n = 0;
  int64_t nnzJ = 0;
  int rec_bits = 16;  // 16: [first:1][po:7][qo:7] (J columns <= 128 entries); 32: [first:1][po:15][qo:15]
  std::vector<uint16_t> rec16;  // pair stream: slot-major, ascending residual row, chunks padded to 32
  std::vector<uint32_t> rec32;
  std::vector<uint16_t> grp_slot;  // per 32-record group: chunk-local slot of its first record
  // chunk c covers strictly-lower slots [chunk_slot[c], chunk_slot[c+1]) and padded records
  // [chunk_rec[c], chunk_rec[c]+chunk_npairs[c])
  std::vector<int> chunk_slot;
  std::vector<int64_t> chunk_rec;
  std::vector<int> chunk_npairs;
  // stage list of chunk c: J columns [chunk_stage_ptr[c], chunk_stage_ptr[c+1]) of stage_cols, with their
  // offsets in the shared-memory value buffer; chunk_staged[c]==0: list too big, gather from global
  std::vector<int> chunk_stage_ptr, stage_cols, stage_off;
  std::vector<uint16_t> stage_run;  // per stage entry: doubles of the bulk copy starting at this column's (aligned) window, 0 unless it heads a run of consecutive columns
  std::vector<uint8_t> chunk_staged;
  std::vector<uint16_t> slot_start;        // per slot: first record of the slot, relative to its chunk
  std::vector<uint16_t> slot_cnt;          // per slot: number of records (saturated; multi-tile slots are alone in their chunk)
  std::vector<int> proc_perm;              // processing order: position -> slot, sorted by descending count inside a chunk
  std::vector<uint16_t> slot_il, slot_jl;  // per slot: stage-list index of its row column i / its column j
  // per strictly-lower slot, column-major, ascending row inside a column
  std::vector<int> low_colptr;      // n+1
  std::vector<int> low_row;         // row index i of the slot
  std::vector<uint16_t> low_rank;   // rank of column j inside row i of H (position of H(i,j) in row i)
  std::vector<uint16_t> low_rank_up;  // rank of row i inside row j of H (position of the mirror H(j,i))
  int64_t npairs = 0;
  int nslots = 0;
  // column kernel chunks: chunk c covers columns [colchunk[c], colchunk[c+1])
  std::vector<int> colchunk;
};

// tunables shared by host planning and kernels
constexpr int kPairTile = 2048;   // products staged per tile
constexpr int kSlotCap = 512;     // strictly-lower slots per chunk
constexpr int kStageVals = 3072;  // J values staged in shared memory per chunk
constexpr int kStageCols = 192;   // J columns in a chunk's stage list
constexpr int kEntTile = 2048;    // J entries staged per tile in the column kernel
constexpr int kColCap = 512;      // columns per column-kernel chunk

// device-side packed forms of the plan (built by upload_analysis)
struct ChunkDesc {   // 32 bytes
  int slot0, nslots, npairs, stage0;
  uint16_t nstage, flags;  // flags: 1 staged in shared memory, 2 has columns > 32 entries, 4 TMA-safe windows
  int stage_bytes;         // sum of the 16-byte-aligned column windows (mbarrier expect_tx)
  int64_t rec0;
};
struct StageEntry {  // 16 bytes: one J column of a chunk's stage list (ascending columns)
  int src;             // Jcolptr[col]
  uint16_t off, len;   // offset in the shared value buffer, column length
  // slice_off[col>>5] + (col&31): storage position of H row `col`, entry 0 (low 48 bits); the high 16
  // bits hold the length in doubles of the bulk copy that starts at this column when it heads a run
  // of consecutive columns (0 otherwise)
  int64_t slice_base;
};
struct SlotMeta {    // 8 bytes
  uint16_t il, jl, rank, rank_up;
};

// ---- "stream" form of the pair assembly (pairs_v2.cpp packs it, assembly_v2.cu runs it) -----------------------
// Everything a chunk needs arrives in shared memory through the TMA unit: ONE bulk copy of the chunk's section
// (tables + slot metadata + pair records, contiguous in `blob`) and one bulk copy per run of consecutive J columns.
// Section layout (bytes, 16-byte aligned, length a multiple of 16, at most 65520; nsp = nslots rounded up to 32):
//   [0,16)   uint16 nslots, nstage, ngroups, off_meta, off_ginfo, off_cnt, off_rec, 0     (byte offsets of the tables)
//   uint2  stab[nstage]    per staged J column {storage position of H row `col`, entry 0 (slice_off[col>>5] + (col&31)),
//                          BYTE offset of the column in the staged value buffer}
//   uint32 meta[nsp]       [il:8][jl:8][rank:8][rank_up:8] per slot (lane t&31 of group t>>5)
//   uint32 ginfo[ngroups]  first record row of the group | rows << 16   (a row = 32 records, one per lane)
//   uint8  cnt[nsp]        products of the slot (0: padding lane)
//   uint16 rec[32*rows]    [po:7][qo:7] positions inside J column i / J column j
struct V2Section {  // pointers into one section
  int ns, nst, ng;
  const uint32_t *stab, *meta, *ginfo;  // stab: two words per column
  const uint16_t* rec;
  const uint8_t* cnt;
  explicit V2Section(const unsigned char* sec) {
    const uint16_t* hd = reinterpret_cast<const uint16_t*>(sec);
    ns = hd[0];
    nst = hd[1];
    ng = hd[2];
    stab = reinterpret_cast<const uint32_t*>(sec + 16);
    meta = reinterpret_cast<const uint32_t*>(sec + hd[3]);
    ginfo = reinterpret_cast<const uint32_t*>(sec + hd[4]);
    cnt = sec + hd[5];
    rec = reinterpret_cast<const uint16_t*>(sec + hd[6]);
  }
};
struct ChunkHdr2 {  // 32 bytes
  uint64_t blob_off;    // byte offset of the section
  uint32_t run0;        // first entry of the chunk's run list
  uint32_t val_bytes;   // bytes the runs bring in (mbarrier transaction count, together with the section)
  // A run whose even-padded window would end one element past the J value array is shortened by two elements and
  // the last real element is copied by a plain load/store: tail_src = its index in the value array + 1 (0 = none),
  // tail_dst its offset in the value buffer.
  uint32_t tail_src;
  uint16_t sec16;       // section length in 16-byte units
  uint16_t tail_dst;
  uint16_t nruns, nslots, nstage, ndiag;  // ndiag: number of 32-lane groups
};
struct Run2 {  // 8 bytes: one bulk copy of `len` doubles (even) from Jv + src (even) to the value buffer + dst (even)
  uint32_t src;
  uint16_t dst, len;
};
struct PairsV2Host {
  bool ok = false;
  const char* why = "";
  std::vector<ChunkHdr2> hdr;
  std::vector<Run2> runs;
  std::vector<unsigned char> blob;
  uint32_t max_val_bytes = 0, max_sec_bytes = 0;
};
void pack_pairs_v2(const int* colptr, const AssemblyPlan& P, const SellLayout& S, PairsV2Host& out);
// host execution of the packed form, in the kernel's own order (CPU tests of the packer)
void run_pairs_v2_host(const PairsV2Host& V, const double* Jv, double* hval);

// Host-side pieces of the last analysis kept for incremental re-analysis (analysis_update.cu): the J pattern and the
// parts of the assembly plan that depend on it.  Kept only for patterns below kPlanKeepMaxNnz entries.
struct PlanKeep {
  bool valid = false;
  std::vector<int> Jrow;                       // row indices of J (the column pointers are spb_context::Jcolptr_host)
  std::vector<int> stage_cols, chunk_stage_ptr;  // stage list of every chunk (ascending columns)
  std::vector<int> chunk_slot, chunk_npairs;
  std::vector<int64_t> chunk_rec;
  std::vector<uint8_t> chunk_staged;
  std::vector<int> low_colptr, low_row;        // strictly-lower slots, column-major (stream order)
  std::vector<uint16_t> slot_start, slot_cnt;  // per slot (stream order): first record relative to its chunk, record count
  std::vector<int> sorted_pos;                 // per chunk: its position in spb_context::d_chunk_desc_sorted / chunk_need
};
constexpr int64_t kPlanKeepMaxNnz = 120000000;

// Host threads for the symbolic phase (SPB_HOST_THREADS overrides; 1 = serial).
inline int host_threads() {
  static const int t = [] {
    const char* e = getenv("SPB_HOST_THREADS");
    int v = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(v, 32));
  }();
  return t;
}
// Runs fn(part) for part in [0,parts) on its own thread; rethrows the first exception.
template <class F>
inline void run_parts(int parts, F fn) {
  if (parts <= 1) {
    fn(0);
    return;
  }
  std::vector<std::thread> th;
  std::vector<std::exception_ptr> err((size_t)parts);
  for (int p = 0; p < parts; ++p)
    th.emplace_back([&, p] {
      try {
        fn(p);
      } catch (...) {
        err[p] = std::current_exception();
      }
    });
  for (auto& t : th) t.join();
  for (auto& e : err)
    if (e) std::rethrow_exception(e);
}

void build_h_pattern(int m, int n, const int* colptr, const int* rowidx, HPattern& H);
void build_sell(const HPattern& H, SellLayout& S);
void build_assembly_plan(int m, int n, const int* colptr, const int* rowidx, const HPattern& H, AssemblyPlan& P);
// the three above in one pass over a shared row view of J (multithreaded; SPB_HOST_THREADS)
void build_analysis(int m, int n, const int* colptr, const int* rowidx, HPattern& H, SellLayout& S, AssemblyPlan& P);
uint64_t pattern_fingerprint(int m, int n, const int* colptr, const int* rowidx);

}  // namespace spb

// ---- the handle -------------------------------------------------------------------------
namespace spb {

enum Stage { ST_ASSEMBLY = 0, ST_SPMV = 1, ST_PCGVEC = 2, ST_FACTOR = 3, ST_TRISOLVE = 4, ST_REDUCE = 5, ST_TRAITS = 6, ST_PCG = 7, ST_COUNT = 8 };

// Scalars of the PCG recurrences and of the LM reductions; lives in device memory with a
// pinned host mirror that is refreshed by explicit async copies.
struct PcgScalars {
  double rz, pq, rr, bb, alpha, beta, thresh;
  int iters, done, maxit, breakdown;
  unsigned int ticket_a, ticket_b;
  int pad0, pad1;
  double red[4];  // local partial sums handed to ncclAllReduce by the distributed PCG
  double xx;      // ||x||^2 of the solution, by-product of the persistent kernel (saves the LM loop a launch and a sync)
};

// ---- peer-memory exchange of the subdomain (halo) mode: dist.cu sets it up, pcg_peer.cu uses it --------------
// Every rank owns one WINDOW of device memory that its neighbours (halo slabs) and all ranks (dot-product
// slots) write directly over NVLink (CUDA IPC mappings); the layout is identical on every rank.
// One rank's contribution to a fused all-to-all sum of up to three doubles, in the flag-in-word form NCCL's LL
// protocol uses: every 8-byte word carries 32 bits of payload and the 32-bit round tag, so a word is complete the
// moment its tag matches (8-byte stores are atomic) and no fence or separate flag store is needed.
struct PeerSlot {
  unsigned long long w[8];  // w[2k], w[2k+1]: low / high half of value k; 64 bytes = one slot per (parity, source rank)
};
struct PeerLayout {  // byte offsets inside a window
  size_t flag_from_left = 0, flag_from_right = 0;  // (reserved)
  size_t abort_word = 0;                           // u32: a rank gave up waiting; everybody stops
  size_t slots = 0;                                // PeerSlot[2][nranks]  (round parity)
  size_t band_from_left = 0, band_from_right = 0;  // u64[2][band][2]: halo slabs, payload half + exchange tag per word (exchange parity)
  size_t bytes = 0;
};
struct PeerCtl {  // local (not peer-visible) counters that persist across solves
  unsigned long long band_seq;   // band exchanges completed so far
  unsigned long long round;      // all-to-all rounds completed so far
  unsigned long long reserved0, reserved1;
  int failed, pad;
};
struct PeerState {
  char* win = nullptr;         // own window (cudaMalloc; exported with cudaIpcGetMemHandle)
  std::vector<char*> peer;     // windows of all ranks as mapped here (peer[rank] == win)
  DevBuf<char*> d_peer;
  DevBuf<PeerCtl> d_ctl;
  PeerLayout L;
  int band = 0, nranks = 1, rank = 0;
  bool ready = false;   // every rank mapped every window: the fused kernel may be used
  bool failed = false;  // a solve timed out waiting for a peer: the exchange state is lost
};

struct Timing {
  bool enabled = false;
  double ms[ST_COUNT] = {0};
  int64_t count[ST_COUNT] = {0};
  struct Pending {
    cudaEvent_t a, b;
    int stage;
  };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
};

}  // namespace spb

struct spb_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  int num_sms = 148;

  int solver = SPB_SOLVER_PCG_JACOBI;
  double pcg_rtol = 1e-13;
  int pcg_maxit = 10000;

  // pattern state
  bool analyzed = false;      // J pattern analysed (assembly possible)
  bool have_h = false;        // H pattern + SpMV layout present (analyze or set_matrix)
  bool h_valid = false;       // numeric H present
  bool factor_valid = false;
  uint64_t fingerprint = 0;   // of the J pattern
  uint64_t h_fingerprint = 0; // of an explicitly set H pattern
  spb::DevBuf<long long> d_dbg;  // debug stopwatch buffer (SPB_ASM_TIMING)
  bool pairs_fast_path = false; // every assembly chunk is on the common path (pipelined kernel)
  int occ_pairs = 0;
  // "stream" form of the pair assembly (assembly_v2.cu): packed sections + run lists, ring geometry
  bool v2_ok = false;
  spb::DevBuf<unsigned char> d_v2_hdr;  // producer descriptors (header + first runs, 128 bytes per chunk)
  spb::DevBuf<spb::Run2> d_v2_runs;
  spb::DevBuf<unsigned char> d_v2_blob;
  uint32_t v2_val_bytes = 0, v2_sec_bytes = 0;
  int v2_stages = 0;
  size_t v2_smem = 0;
  int last_asm_form = 0;  // spb_assembly_form()
  bool h_block2 = false;  // H is block diagonal with blocks on the aligned pairs (2k, 2k+1) at most (extended Rosenbrock): the batched PCG preconditions with the exact inverse blocks
  cudaStream_t asm_side_stream = nullptr;  // the column kernel runs here, next to the stream-form pair kernel
  cudaEvent_t asm_fork_ev = nullptr, asm_join_ev = nullptr;
  // subdomain (halo) mode, dist.cu: this handle holds the LOCAL problem of one rank; local unknowns are
  // [left ghosts | owned | right ghosts] on a ring of ranks
  bool halo_mode = false;
  int halo = 0, own_lo = 0, own_hi = 0, left_rank = 0, right_rank = 0;
  spb::DevBuf<double> d_halo_tmp;  // 2*halo receive buffer of the reverse exchange
  cudaStream_t halo_stream = nullptr;  // the exchange of the PCG runs here, concurrently with the interior SpMV
  cudaEvent_t halo_ev_band = nullptr, halo_ev_done = nullptr;
  bool solve_xx_valid = false; // the last spb_solve left ||x||^2 in h_sc->xx
  // legacy triplet interface (spb_analyze_triplets): CSC entry q of H sums d_trip_src[d_trip_ptr[q] .. d_trip_ptr[q+1])
  int64_t trip_nnz = -1;      // number of input triplets; -1 = no triplet analysis on this handle
  spb::DevBuf<int64_t> d_trip_ptr;
  spb::DevBuf<int> d_trip_src;
  int m = 0, n = 0;
  int64_t nnzJ = 0;
  spb::HPattern H;            // host copy kept for export and for the Cholesky symbolic phase
  spb::SellLayout S;          // host: only slice_off/rowlen kept after upload
  int rec_bits = 16;
  int nchunks = 0, ncolchunks = 0, nslots = 0;
  // streamed assembly from pinned host memory (spb_assemble, SPB_HOST): J is uploaded in parts on a copy stream
  // while the parts already on the device are being assembled
  std::vector<int64_t> chunk_need;   // per assembly chunk (in d_chunk_desc_sorted order): J values [0, need) must be on the device
  spb::DevBuf<spb::ChunkDesc> d_chunk_desc_sorted;  // the chunk descriptors ordered by that need
  std::vector<int> colchunk_host;    // column boundaries of the column-kernel chunks
  std::vector<int> Jcolptr_host;
  spb::PlanKeep keep;                // for spb_analyze_update
  spb::DevBuf<int> d_keep_stage_cols, d_keep_stage_ptr, d_keep_sorted_pos, d_keep_flags;  // device copies, made by the first update
  uint64_t h_pattern_hash = 0;       // of the structural H pattern (the Cholesky / IC(0) symbolic state depends on it only)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_ev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t asm_done_ev = nullptr;
  bool asm_done_recorded = false;
  int64_t npairs = 0;

  // device: J pattern + assembly plan
  spb::DevBuf<int> d_Jcolptr, d_Jrow;
  spb::DevBuf<uint16_t> d_rec16;
  spb::DevBuf<uint32_t> d_rec32;
  spb::DevBuf<int> d_colchunk, d_nup1;
  spb::DevBuf<spb::ChunkDesc> d_chunk_desc;
  spb::DevBuf<spb::StageEntry> d_stage_ent;
  spb::DevBuf<spb::SlotMeta> d_slot_meta;
  spb::DevBuf<uint32_t> d_slot_start;  // (start << 16) | count, in processing order
  // device: H in SpMV layout
  spb::DevBuf<int> d_sell_col, d_rowlen;
  spb::DevBuf<uint16_t> d_sell_col16;
  spb::DevBuf<uint8_t> d_slice16;
  spb::DevBuf<int64_t> d_slice_off;
  spb::DevBuf<int64_t> d_pos_of_slot;  // lazily built: CSC slot -> storage position
  spb::DevBuf<double> d_hval, d_hdiag, d_dinv;
  spb::DevBuf<float> d_hval32;  // FP32 copy of the value stream (relaxed-precision products of the late PCG iterations)
  // device: vectors
  spb::DevBuf<double> d_rhs, d_dvec, d_Jv_stage, d_f_stage, d_vec_stage, d_vec_stage2;
  spb::DevBuf<double> d_r, d_p, d_q, d_x, d_partial, d_u, d_w;
  spb::DevBuf<spb::PcgScalars> d_sc;
  spb::DevBuf<unsigned long long> d_foo_bits;
  spb::PcgScalars* h_sc = nullptr;      // pinned
  double* h_scalar = nullptr;           // pinned scratch (8 doubles)

  // multi-GPU (dist.cu): NCCL communicator, row partition of the residuals
  void* comm = nullptr;        // ncclComm_t
  int rank = 0, nranks = 1;
  bool partitioned = false;    // spb_analyze_partitioned: this rank owns residual rows [row_begin,row_end)
  int m_global = 0, row_begin = 0, row_end = 0;
  std::vector<int> local_colptr, local_rowidx;
  spb::DevBuf<double> d_hval_part, d_rhs_part, d_dvec_part, d_pfull, d_xfull;
  int64_t collectives = 0;
  spb::PeerState* peer = nullptr;  // peer-memory windows of the halo mode (dist.cu), owned

  // batched independent problems (batch.cu): per-problem lambda of the column kernel; nullptr = the scalar argument
  const double* d_lambda_vec = nullptr;
  int lambda_n_per = 1;
  void* batch_ws = nullptr;  // spb::BatchWorkspace*, owned (batch.cu)
  int pcg_last_iters = 0;
  // relaxed-precision products: switch to the FP32 value stream once ||r||/||b|| <= pcg_rtol * pcg_relax_factor
  // (1e6: 1e-7 at the default rtol 1e-13); 0 = always FP64.  SPB_PCG_RELAX overrides.
  double pcg_relax_factor = 1e6;
  bool pcg_persistent = true;  // one cooperative kernel per solve; false: one launch per phase
  bool pcg_stream_ready = false;  // pcg_stream_kernel's shared-memory opt-in done
  bool pcg_single_reduction = false;  // the two-barrier Chronopoulos-Gear kernel (pcg_peer.cu); false: the three-barrier textbook kernel (pcg.cu)
  int occ_spmv = 0, occ_pcg = 0;
  bool pcg_col16 = true;       // persistent PCG reads 16-bit column deltas where a slice allows it
  double pcg_pin_mb = 0.0;     // MB of H the persistent PCG keeps in L2 (evict_last) across iterations
  int pcg_minb = 4;            // __launch_bounds__ min blocks/SM variant of the persistent kernel
  int64_t launches = 0;
  void* lm_ws = nullptr;  // spb::LmWorkspace*, owned (lm.cu)
  spb::Timing timing;
  void* chol = nullptr;  // spb::CholeskyState*, owned (cholesky.cu)
  void* ic0 = nullptr;   // spb::Ic0Device*, owned (pcg.cu)
  int occ_ic = 0;
};

namespace spb {
// timing helpers (context.cu)
void stage_begin(spb_context* h, int stage);
void stage_end(spb_context* h, int stage, int nlaunch);
void timing_flush(spb_context* h);

// assembly.cu
void upload_analysis(spb_context* h, const int* colptr, const int* rowidx, AssemblyPlan& P);
void upload_h_layout(spb_context* h);
void run_assembly(spb_context* h, const double* d_Jv, const double* d_f, double lambda, double* foo_host);
bool run_assembly_streamed(spb_context* h, const double* host_Jv, const double* host_f, double lambda, double* foo_host);  // false: not applicable
// assembly_v2.cu
void upload_pairs_v2(spb_context* h, const PairsV2Host& V);
bool run_pairs_v2(spb_context* h, const double* d_Jv, double* hv);  // false: not applicable, use the general kernels
void ensure_pos_of_slot(spb_context* h);
void gather_h_values(spb_context* h, double* d_out);   // CSC order
void scatter_h_values(spb_context* h, const double* d_in);
void sum_triplets(spb_context* h, const double* d_trip_vals, double* d_csc_vals);  // insertion-order sums per H entry  // CSC order -> storage (+ hdiag)

// problems.cu: initial point of a device-resident problem straight from the device (false: not one of ours)
bool builtin_x0_device(const spb_traits* T, double* d_dst, cudaStream_t st);
const double* builtin_constant_jacobian(const spb_traits* T);  // device CSC values of a constant Jacobian, or nullptr

// blas1.cu
void dev_squared_norm(spb_context* h, const double* d_v, int64_t len, double* host_out);
void dev_inf_norm(spb_context* h, const double* d_v, int64_t len, double* host_out);
void dev_add(spb_context* h, const double* a, const double* b, double* out, int64_t len);  // out=a+b
void dev_invert_diag(spb_context* h);  // d_dinv = 1/d_hdiag
void lm_release(spb_context* h);
void batch_release(spb_context* h);  // batch.cu

// pcg.cu
void dev_spmv(spb_context* h, const double* d_x, double* d_y);
spb_status pcg_solve(spb_context* h, const double* d_b, double* d_x, int* iters);

// dist.cu: in-stream collectives (no-ops / local copies when nranks == 1)
void dist_group_begin(spb_context* h);
void dist_group_end(spb_context* h);
void dist_allreduce_sum(spb_context* h, const double* send, double* recv, size_t count);
void dist_allgather(spb_context* h, const double* send, double* recv, size_t count_per_rank);
void dist_allreduce_max(spb_context* h, const double* send, double* recv, size_t count);
void halo_forward(spb_context* h, double* v);      // owners -> the neighbours' ghost copies
void halo_reverse_add(spb_context* h, double* v);  // ghost contributions -> owners (+=, left neighbour first)
void halo_sum_both(spb_context* h, double* v, cudaStream_t st = nullptr, bool add = true);  // one exchange: full sums on the owned edges AND on the ghosts
spb_status pcg_solve_halo(spb_context* h, const double* d_b, double* d_x, int* iters);
// pcg_peer.cu: the same solve as ONE persistent cooperative kernel per rank; halo slabs and dot products move
// through the peer windows (PeerState) -- no NCCL call and no kernel boundary inside the solve
spb_status pcg_solve_halo_peer(spb_context* h, const double* d_b, double* d_x, int* iters);
// pcg_peer.cu: the single-GPU solve with the same single-reduction kernel (no peers, no bands)
spb_status pcg_solve_sr(spb_context* h, const double* d_b, double* d_x, int* iters);
void peer_window_setup(spb_context* h);    // collective over the handle's communicator (dist.cu)
void peer_window_release(spb_context* h);
void dist_release(spb_context* h);
// pcg.cu: distributed Jacobi-PCG over a row partition of H (replicated H, partitioned vectors)
spb_status pcg_solve_dist(spb_context* h, const double* d_b, double* d_x, int* iters);
// sum of squares of a vector whose entries are split across ranks (allreduce of the local sums)
void dev_squared_norm_global(spb_context* h, const double* d_v, int64_t len, double* host_out);

// pcg.cu: multicolour IC(0) preconditioner + IC(0)-PCG
spb_status ic0_factorize(spb_context* h);
spb_status pcg_ic0_solve(spb_context* h, const double* d_b, double* d_x, int* iters);
void ic0_release(spb_context* h);

// cholesky.cu
spb_status chol_factorize(spb_context* h);
spb_status chol_solve(spb_context* h, const double* d_b, double* d_x);
void chol_release(spb_context* h);
bool chol_info(spb_context* h, int64_t* nnzL, double* flops, int* nsuper, int* nlevels);
}  // namespace spb
