// dist.cu -- multi-GPU plumbing for ONE large problem (config C5): NCCL communicator owned by
// the handle, row-partitioned symbolic analysis, and the in-stream collectives used by the
// assembly (sum of partial H / rhs / d) and by the distributed PCG (dot products, search
// direction).  One process per GPU; NCCL over NVLink 5 / NVSwitch.
//
// Why it shards: residual rows are independent, J^T J = sum_g J_g^T J_g and
// J^T f = sum_g J_g^T f_g, so every rank assembles the partial normal equations of its row
// block with the unchanged single-GPU kernels and one grouped ncclAllReduce combines them.
// The reference has no distributed path (single process, single thread).
//
// NCCL is bound at run time (dlopen) so that the library loads on machines without it and
// shares the copy PyTorch has already loaded when used from bench.py.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "spb_internal.h"

namespace spb {

struct Id128 {  // ncclUniqueId: 128 opaque bytes, passed by value
  char bytes[128];
};
namespace {
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8;  // ncclDataType_t::ncclDouble
constexpr int kNcclSum = 0;      // ncclRedOp_t::ncclSum
constexpr int kNcclMax = 2;      // ncclRedOp_t::ncclMax

bool load_nccl(std::string& err) {
  if (g_nccl.lib) return true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // reuse the copy already in the process (PyTorch's)
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
  if (!lib) {
    err = std::string("cannot load libnccl: ") + dlerror();
    return false;
  }
  auto sym = [&](const char* name) { return dlsym(lib, name); };
  g_nccl.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))sym("ncclAllGather");
  g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
  g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
  g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
  g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart ||
      !g_nccl.GroupEnd) {
    err = "libnccl is missing required symbols";
    return false;
  }
  g_nccl.lib = lib;
  return true;
}

#define SPB_NCCL(expr)                                 \
  do {                                                 \
    int _r = (expr);                                   \
    if (_r != 0) throw NcclFail{_r, #expr};            \
  } while (0)
}  // namespace

// ---- in-stream collectives used by assembly.cu / pcg.cu / lm.cu ---------------------------
void dist_group_begin(spb_context* h) {
  if (h->nranks > 1) SPB_NCCL(g_nccl.GroupStart());
}
void dist_group_end(spb_context* h) {
  if (h->nranks > 1) SPB_NCCL(g_nccl.GroupEnd());
}
void dist_allreduce_sum(spb_context* h, const double* send, double* recv, size_t count) {
  if (h->nranks <= 1) {
    if (send != recv) SPB_CUDA(cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return;
  }
  SPB_NCCL(g_nccl.AllReduce(send, recv, count, kNcclFloat64, kNcclSum, h->comm, h->stream));
  h->collectives++;
}
void dist_allgather(spb_context* h, const double* send, double* recv, size_t count_per_rank) {
  if (h->nranks <= 1) {
    if (send != recv) SPB_CUDA(cudaMemcpyAsync(recv, send, count_per_rank * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return;
  }
  SPB_NCCL(g_nccl.AllGather(send, recv, count_per_rank, kNcclFloat64, h->comm, h->stream));
  h->collectives++;
}

void dist_allreduce_max(spb_context* h, const double* send, double* recv, size_t count) {
  if (h->nranks <= 1) {
    if (send != recv) SPB_CUDA(cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return;
  }
  SPB_NCCL(g_nccl.AllReduce(send, recv, count, kNcclFloat64, kNcclMax, h->comm, h->stream));
  h->collectives++;
}

// ---- subdomain (halo) mode -------------------------------------------------------------------
// The scalable form of the one-large-problem case: a rank owns a contiguous range of unknowns and
// the residual rows based there, and analyses / assembles only its LOCAL J_g (owned + halo columns).
// The normal matrix is never summed across ranks: H x = sum_g R_g^T (J_g^T J_g) R_g x, so one product
// is  forward halo exchange of p  ->  local SpMV  ->  reverse exchange of the ghost rows of q, added
// on their owners (left neighbour's contribution first: fixed order, deterministic).  Per PCG
// iteration the traffic is two halo slabs each way plus two scalar allreduces, independent of n/ranks.
// Exchanges are grouped ncclSend/ncclRecv on the handle's stream over NVLink.
__global__ void halo_add_kernel(double* __restrict__ v, const double* __restrict__ from_left, const double* __restrict__ from_right,
                                int own_lo, int own_hi, int halo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= halo) return;
  v[own_lo + i] += from_left[i];          // the left neighbour's right-ghost rows are my first owned entries
  v[own_hi - halo + i] += from_right[i];  // the right neighbour's left-ghost rows are my last owned entries
}

void halo_forward(spb_context* h, double* v) {
  const size_t hl = (size_t)h->halo;
  if (h->nranks <= 1) {  // a single rank that is its own neighbour on the ring (test configuration)
    SPB_CUDA(cudaMemcpyAsync(v + h->own_hi, v + h->own_lo, hl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    SPB_CUDA(cudaMemcpyAsync(v, v + h->own_hi - hl, hl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return;
  }
  SPB_NCCL(g_nccl.GroupStart());
  SPB_NCCL(g_nccl.Send(v + h->own_lo, hl, kNcclFloat64, h->left_rank, h->comm, h->stream));          // my first owned -> left's right ghosts
  SPB_NCCL(g_nccl.Send(v + h->own_hi - hl, hl, kNcclFloat64, h->right_rank, h->comm, h->stream));    // my last owned  -> right's left ghosts
  SPB_NCCL(g_nccl.Recv(v + h->own_hi, hl, kNcclFloat64, h->right_rank, h->comm, h->stream));         // right ghosts <- right's first owned
  SPB_NCCL(g_nccl.Recv(v, hl, kNcclFloat64, h->left_rank, h->comm, h->stream));                      // left ghosts  <- left's last owned
  SPB_NCCL(g_nccl.GroupEnd());
  h->collectives++;
}

void halo_reverse_add(spb_context* h, double* v) {
  const size_t hl = (size_t)h->halo;
  SPB_CUDA(h->d_halo_tmp.alloc(2 * hl));
  double* from_right = h->d_halo_tmp.p;
  double* from_left = h->d_halo_tmp.p + hl;
  if (h->nranks <= 1) {  // self ring: my right-ghost rows are contributions to my first owned entries, and vice versa
    SPB_CUDA(cudaMemcpyAsync(from_left, v + h->own_hi, hl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    SPB_CUDA(cudaMemcpyAsync(from_right, v, hl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    halo_add_kernel<<<(h->halo + 255) / 256, 256, 0, h->stream>>>(v, from_left, from_right, h->own_lo, h->own_hi, h->halo);
    h->launches++;
    SPB_CUDA(cudaGetLastError());
    return;
  }
  SPB_NCCL(g_nccl.GroupStart());
  SPB_NCCL(g_nccl.Send(v, hl, kNcclFloat64, h->left_rank, h->comm, h->stream));                  // my left-ghost rows  -> left (its last owned)
  SPB_NCCL(g_nccl.Send(v + h->own_hi, hl, kNcclFloat64, h->right_rank, h->comm, h->stream));     // my right-ghost rows -> right (its first owned)
  SPB_NCCL(g_nccl.Recv(from_right, hl, kNcclFloat64, h->right_rank, h->comm, h->stream));
  SPB_NCCL(g_nccl.Recv(from_left, hl, kNcclFloat64, h->left_rank, h->comm, h->stream));
  SPB_NCCL(g_nccl.GroupEnd());
  halo_add_kernel<<<(h->halo + 255) / 256, 256, 0, h->stream>>>(v, from_left, from_right, h->own_lo, h->own_hi, h->halo);
  h->launches++;
  h->collectives++;
  SPB_CUDA(cudaGetLastError());
}

// Both directions in one exchange: afterwards v holds the FULL sum on the owned edges and on the
// ghosts (same bits on both holders: a + b is commutative).  A rank sends each neighbour its
// partials for the 2*halo entries they share -- [its ghost rows of that neighbour | its own edge rows
// next to that neighbour] -- and adds what it receives.
__global__ void halo_sum_both_kernel(double* __restrict__ v, const double* __restrict__ from_left, const double* __restrict__ from_right,
                                     int own_lo, int own_hi, int halo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * halo) return;
  // the shared band with the left neighbour is local [0, 2*halo) (ghosts, then my first owned); with the
  // right neighbour local [own_hi - halo, own_hi + halo).  Both ranks pack a band in ascending global order.
  v[i] += from_left[i];
  v[own_hi - halo + i] += from_right[i];
}

void halo_sum_both(spb_context* h, double* v, cudaStream_t st, bool add) {
  if (!st) st = h->stream;
  const size_t hl = (size_t)h->halo, band = 2 * hl;
  SPB_CUDA(h->d_halo_tmp.alloc(2 * band));
  double* from_right = h->d_halo_tmp.p;
  double* from_left = h->d_halo_tmp.p + band;
  if (h->nranks <= 1) {  // self ring: the left band [0,2h) and the right band [own_hi-h, own_hi+h) are the same global entries
    SPB_CUDA(cudaMemcpyAsync(from_left, v + h->own_hi - hl, band * sizeof(double), cudaMemcpyDeviceToDevice, st));
    SPB_CUDA(cudaMemcpyAsync(from_right, v, band * sizeof(double), cudaMemcpyDeviceToDevice, st));
  } else {
    SPB_NCCL(g_nccl.GroupStart());
    SPB_NCCL(g_nccl.Send(v, band, kNcclFloat64, h->left_rank, h->comm, st));                    // my view of the band shared with left
    SPB_NCCL(g_nccl.Send(v + h->own_hi - hl, band, kNcclFloat64, h->right_rank, h->comm, st));  // ... shared with right
    SPB_NCCL(g_nccl.Recv(from_right, band, kNcclFloat64, h->right_rank, h->comm, st));          // right's view of our band
    SPB_NCCL(g_nccl.Recv(from_left, band, kNcclFloat64, h->left_rank, h->comm, st));            // left's view of our band
    SPB_NCCL(g_nccl.GroupEnd());
    h->collectives++;
  }
  if (!add) return;  // the caller folds the additions into its next kernel (d_halo_tmp: [from_right | from_left])
  halo_sum_both_kernel<<<(int)((band + 255) / 256), 256, 0, st>>>(v, from_left, from_right, h->own_lo, h->own_hi, h->halo);
  h->launches++;
  SPB_CUDA(cudaGetLastError());
}

// ---- peer-memory windows of the halo mode (used by pcg_peer.cu) ---------------------------------------------
// One process per GPU: every rank allocates its window with cudaMalloc, exports it with cudaIpcGetMemHandle,
// the 64-byte handles are all-gathered over the handle's communicator, and every rank maps all windows
// (cudaIpcOpenMemHandle enables peer access over NVLink).  The fused kernel is used only if EVERY rank
// mapped every window (one max-allreduce of a failure flag); otherwise all ranks keep the NCCL path.
// A single rank that is its own neighbour (self ring; test configuration) maps nothing and uses its own window.
void peer_window_release(spb_context* h) {
  PeerState* S = h->peer;
  if (!S) return;
  for (size_t r = 0; r < S->peer.size(); ++r)
    if (S->peer[r] && S->peer[r] != S->win) cudaIpcCloseMemHandle(S->peer[r]);
  if (S->win) cudaFree(S->win);
  delete S;
  h->peer = nullptr;
}

void peer_window_setup(spb_context* h) {
  peer_window_release(h);
  const bool disabled = getenv("SPB_HALO_PEER") && atoi(getenv("SPB_HALO_PEER")) == 0;  // read at every setup: tests toggle it
  if (disabled || !h->halo_mode) return;
  const int R = std::max(1, h->nranks);
  PeerState* S = new PeerState;
  h->peer = S;
  S->nranks = R;
  S->rank = R > 1 ? h->rank : 0;
  S->band = 2 * h->halo;
  auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
  PeerLayout& L = S->L;
  size_t off = 0;
  L.flag_from_left = off;
  off += 256;
  L.flag_from_right = off;
  off += 256;
  L.abort_word = off;
  off += 256;
  L.slots = off;
  off = align(off + sizeof(PeerSlot) * 2 * (size_t)R);
  L.band_from_left = off;  // two 8-byte words (payload half + tag) per double, two parities
  off = align(off + sizeof(unsigned long long) * 2 * 2 * (size_t)S->band);
  L.band_from_right = off;
  off = align(off + sizeof(unsigned long long) * 2 * 2 * (size_t)S->band);
  L.bytes = off;
  cudaStream_t st = h->stream;
  SPB_CUDA(cudaMalloc((void**)&S->win, L.bytes));
  SPB_CUDA(cudaMemsetAsync(S->win, 0, L.bytes, st));
  SPB_CUDA(S->d_ctl.alloc(1));
  SPB_CUDA(cudaMemsetAsync(S->d_ctl.p, 0, sizeof(PeerCtl), st));
  S->peer.assign((size_t)R, nullptr);
  S->peer[(size_t)S->rank] = S->win;
  double fail = 0.0;
  if (R > 1) {
    // exchange the IPC handles (64 bytes each) through the communicator
    DevBuf<char> d_hnd;
    DevBuf<double> d_flag;
    SPB_CUDA(d_hnd.alloc((size_t)R * sizeof(cudaIpcMemHandle_t)));
    SPB_CUDA(d_flag.alloc(1));
    cudaIpcMemHandle_t mine;
    cudaError_t e = cudaIpcGetMemHandle(&mine, S->win);
    std::vector<cudaIpcMemHandle_t> all((size_t)R);
    if (e != cudaSuccess) {
      cudaGetLastError();
      fail = 1.0;
      memset(&mine, 0, sizeof(mine));
    }
    SPB_CUDA(cudaMemcpyAsync(d_hnd.p + (size_t)S->rank * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    SPB_NCCL(g_nccl.AllGather(d_hnd.p + (size_t)S->rank * sizeof(mine), d_hnd.p, sizeof(mine), 0 /* ncclInt8 */, h->comm, st));
    SPB_CUDA(cudaMemcpyAsync(all.data(), d_hnd.p, (size_t)R * sizeof(mine), cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaStreamSynchronize(st));  // also: every rank's window is zeroed before anybody can write into it
    if (fail == 0.0) {
      for (int r = 0; r < R; ++r) {
        if (r == S->rank) continue;
        void* ptr = nullptr;
        e = cudaIpcOpenMemHandle(&ptr, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
          cudaGetLastError();
          fail = 1.0;
          break;
        }
        S->peer[(size_t)r] = (char*)ptr;
      }
    }
    SPB_CUDA(cudaMemcpyAsync(d_flag.p, &fail, sizeof(double), cudaMemcpyHostToDevice, st));
    SPB_NCCL(g_nccl.AllReduce(d_flag.p, d_flag.p, 1, kNcclFloat64, kNcclMax, h->comm, st));
    SPB_CUDA(cudaMemcpyAsync(&fail, d_flag.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    h->collectives += 2;
  } else {
    SPB_CUDA(cudaStreamSynchronize(st));
  }
  if (fail != 0.0) {  // some rank cannot map some window: everybody stays on the NCCL exchange
    peer_window_release(h);
    return;
  }
  std::vector<char*> pv(S->peer.begin(), S->peer.end());
  SPB_CUDA(S->d_peer.upload(pv, st));
  SPB_CUDA(cudaStreamSynchronize(st));
  S->ready = true;
}

void dist_release(spb_context* h) {
  peer_window_release(h);
  if (h->halo_ev_band) cudaEventDestroy(h->halo_ev_band);
  if (h->halo_ev_done) cudaEventDestroy(h->halo_ev_done);
  if (h->halo_stream) cudaStreamDestroy(h->halo_stream);
  h->halo_ev_band = h->halo_ev_done = nullptr;
  h->halo_stream = nullptr;
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  h->comm = nullptr;
  h->nranks = 1;
  h->rank = 0;
}

}  // namespace spb

using namespace spb;

template <class F>
static spb_status guarded_dist(spb_handle h, F&& f) {
  try {
    cudaGetLastError();
    return f();
  } catch (const CudaFail& e) {
    if (h) h->err = std::string("CUDA error: ") + cudaGetErrorString(e.e) + " at " + e.where;
    return SPB_ERR_CUDA;
  } catch (const NcclFail& e) {
    if (h) h->err = std::string("NCCL error: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e.code) : "?") + " at " + e.where;
    return SPB_ERR_NCCL;
  } catch (const ArgFail& e) {
    if (h) h->err = e.msg;
    return SPB_ERR_ARG;
  } catch (const StateFail& e) {
    if (h) h->err = e.msg;
    return SPB_ERR_STATE;
  } catch (const std::exception& e) {
    if (h) h->err = std::string("internal error: ") + e.what();
    return SPB_ERR_STATE;
  } catch (...) {
    if (h) h->err = "internal error (unknown exception)";
    return SPB_ERR_STATE;
  }
}

extern "C" {

spb_status spb_comm_unique_id(void* id128) {
  if (!id128) return SPB_ERR_ARG;
  std::string err;
  if (!load_nccl(err)) return SPB_ERR_NCCL;
  return g_nccl.GetUniqueId(id128) == 0 ? SPB_OK : SPB_ERR_NCCL;
}

spb_status spb_comm_init(spb_handle h, const void* id128, int rank, int nranks) {
  if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return SPB_ERR_ARG;
  return guarded_dist(h, [&]() {
    std::string err;
    if (!load_nccl(err)) {
      h->err = err;
      return SPB_ERR_NCCL;
    }
    SPB_CUDA(cudaSetDevice(h->device));
    dist_release(h);
    Id128 id;
    memcpy(id.bytes, id128, 128);
    void* comm = nullptr;
    SPB_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    h->comm = comm;
    h->rank = rank;
    h->nranks = nranks;
    return SPB_OK;
  });
}

spb_status spb_dist_ring_halo(spb_handle h, int32_t own_begin, int32_t own_end, int32_t halo, int32_t left_rank, int32_t right_rank) {
  if (!h) return SPB_ERR_ARG;
  return guarded_dist(h, [&]() {
    if (halo == 0) {  // switch the mode off
      h->halo_mode = false;
      peer_window_release(h);
      return SPB_OK;
    }
    // without a communicator only the self ring (left == right == 0: one rank that is its own neighbour) is possible
    const bool self_ring = h->nranks <= 1 && left_rank == 0 && right_rank == 0;
    if (!self_ring && (!h->comm || h->nranks < 2)) throw StateFail{"spb_dist_ring_halo needs a communicator of at least two ranks (spb_comm_init)"};
    if (halo < 0 || own_begin != halo || own_end - own_begin < 2 * halo || left_rank < 0 || left_rank >= h->nranks || right_rank < 0 ||
        right_rank >= h->nranks || (!self_ring && (left_rank == h->rank || right_rank == h->rank)))
      throw ArgFail{"bad halo description: local unknowns are [halo ghosts | owned (>= 2*halo) | halo ghosts], neighbours are other ranks"};
    h->halo_mode = true;
    h->halo = halo;
    h->own_lo = own_begin;
    h->own_hi = own_end;
    h->left_rank = left_rank;
    h->right_rank = right_rank;
    SPB_CUDA(cudaSetDevice(h->device));
    peer_window_setup(h);  // collective: every rank of the communicator calls spb_dist_ring_halo
    return SPB_OK;
  });
}

spb_status spb_dist_peer_ready(spb_handle h, int* ready) {
  if (!h || !ready) return SPB_ERR_ARG;
  *ready = (h->halo_mode && h->peer && h->peer->ready && !h->peer->failed) ? 1 : 0;
  return SPB_OK;
}

spb_status spb_comm_destroy(spb_handle h) {
  if (!h) return SPB_ERR_ARG;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  dist_release(h);
  return SPB_OK;
}

// Same global Jacobian pattern on every rank; this rank owns residual rows [row_begin,row_end).
spb_status spb_analyze_partitioned(spb_handle h, int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx, int32_t row_begin,
                                   int32_t row_end) {
  if (!h || !colptr || row_begin < 0 || row_end < row_begin || row_end > m) return SPB_ERR_ARG;
  return guarded_dist(h, [&]() {
    SPB_CUDA(cudaSetDevice(h->device));
    h->analyzed = h->have_h = h->h_valid = h->factor_valid = false;
    h->trip_nnz = -1;
    h->pcg_last_iters = 0;
    h->h_pattern_hash = 0;
    h->keep.valid = false;
    chol_release(h);
    ic0_release(h);
    // H pattern from the GLOBAL Jacobian (identical on every rank, so the partial H arrays align)
    build_h_pattern(m, n, colptr, rowidx, h->H);
    build_sell(h->H, h->S);
    // local pattern: owned rows, renumbered from 0, CSC order preserved
    std::vector<int>& lp = h->local_colptr;
    std::vector<int>& li = h->local_rowidx;
    lp.assign((size_t)n + 1, 0);
    li.clear();
    for (int j = 0; j < n; ++j) {
      for (int q = colptr[j]; q < colptr[j + 1]; ++q) {
        int r = rowidx[q];
        if (r >= row_begin && r < row_end) li.push_back(r - row_begin);
      }
      lp[(size_t)j + 1] = (int)li.size();
    }
    const int ml = row_end - row_begin;
    h->m = ml;
    h->n = n;
    h->nnzJ = (int64_t)li.size();
    h->m_global = m;
    h->row_begin = row_begin;
    h->row_end = row_end;
    AssemblyPlan P;
    build_assembly_plan(ml, n, lp.data(), li.data(), h->H, P);
    upload_h_layout(h);
    upload_analysis(h, lp.data(), li.data(), P);
    h->keep = PlanKeep{};  // incremental re-analysis is a single-GPU, plain-handle feature
    // partial buffers (inactive slots stay zero forever: they are never written)
    SPB_CUDA(h->d_hval_part.alloc((size_t)h->S.padded));
    SPB_CUDA(cudaMemsetAsync(h->d_hval_part.p, 0, (size_t)h->S.padded * sizeof(double), h->stream));
    SPB_CUDA(h->d_rhs_part.alloc(n));
    SPB_CUDA(h->d_dvec_part.alloc(n));
    SPB_CUDA(cudaStreamSynchronize(h->stream));
    h->fingerprint = pattern_fingerprint(m, n, colptr, rowidx) ^ ((uint64_t)row_begin << 32) ^ (uint64_t)row_end;
    h->h_fingerprint = 0;
    h->partitioned = true;
    h->analyzed = true;
    h->have_h = true;
    return SPB_OK;
  });
}

spb_status spb_local_nnz(spb_handle h, int64_t* nnz) {
  if (!h || !nnz) return SPB_ERR_ARG;
  if (!h->analyzed) return SPB_ERR_STATE;
  *nnz = h->nnzJ;
  return SPB_OK;
}

spb_status spb_local_pattern(spb_handle h, int32_t* colptr, int32_t* rowidx) {
  if (!h) return SPB_ERR_ARG;
  if (!h->analyzed || !h->partitioned) return SPB_ERR_STATE;
  if (colptr) memcpy(colptr, h->local_colptr.data(), sizeof(int32_t) * h->local_colptr.size());
  if (rowidx) memcpy(rowidx, h->local_rowidx.data(), sizeof(int32_t) * h->local_rowidx.size());
  return SPB_OK;
}

int64_t spb_collective_count(spb_handle h) { return h ? h->collectives : 0; }

}  // extern "C"
