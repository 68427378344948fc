// assembly_v2.cu -- the "stream" form of the pair assembly: H(i,j) = sum_r J(r,i)*J(r,j) for the strictly-lower
// slots (and their mirrors), LMSolver.h:136, from the packed form of pairs_v2.cpp.
//
// One persistent CTA per SM, warp specialised.  The PRODUCER warp walks the CTA's chunks and feeds a ring of
// shared-memory stages through the TMA unit only: per chunk one bulk copy of the chunk's section (tables, 4-byte
// slot metadata, pair records in lane-interleaved rows) and one bulk copy per run of consecutive J columns, all
// completing on the stage's `full` mbarrier.  The CONSUMER warps wait on `full`, take 32 slots each (a group: slots of
// similar pair count, lanes chosen on the host so that their operand loads spread over the shared-memory banks; row d of
// the group's records is one contiguous 64-byte read), add the products in ascending residual row with separately rounded multiply and add
// (first product assigned) -- the same bits as the reference's sparse product and as the other assembly kernels --
// store H(i,j) and H(j,i) into the SpMV storage, and release the stage on its `empty` mbarrier.  There is no CTA-wide
// barrier: warps that finish early run ahead into the next stage, and the copies of the next stages are in flight
// while a stage is being summed.  The consumers execute ~60 instructions per slot (the register-pipelined kernel in
// assembly.cu: ~290), and a slot costs 4 index bytes + 2 per product instead of 12 + 2.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "spb_internal.h"
#include "tma.cuh"

namespace spb {

constexpr int kV2ConsumerWarps = 16;
constexpr int kV2Threads = (kV2ConsumerWarps + 1) * 32;
constexpr int kV2MaxStages = 8;
constexpr int kV2DescRing = 8;    // producer descriptors prefetched this many chunks ahead (cp.async into shared memory)
constexpr int kV2ExtraRuns = 4;   // further runs per lane the producer reads one chunk ahead (12 + 128 runs without a dependent load)
constexpr int kV2DescRuns = 12;   // runs carried inside a producer descriptor; further runs are read from the run list

// what the producer warp needs for one chunk, at a fixed stride so that the prefetch does not depend on another load
struct PDesc {  // 128 bytes
  ChunkHdr2 h;
  Run2 r[kV2DescRuns];
};
static_assert(sizeof(PDesc) == 128, "PDesc must be 128 bytes");

__device__ __forceinline__ void v2_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// shared-memory loads through 32-bit shared-window addresses (no generic-address arithmetic in the inner loops)
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint32_t v;
  asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kV2Threads, 1)
    assemble_pairs_stream_kernel(const PDesc* __restrict__ pdesc, int nchunks, const Run2* __restrict__ runs,
                                 const unsigned char* __restrict__ blob, const double* __restrict__ Jv, double* __restrict__ hval,
                                 int nstages, uint32_t val_bytes, uint32_t sec_bytes) {
  extern __shared__ __align__(128) unsigned char v2_smem[];
  // [ full[8] | empty[8] | producer descriptor ring | stage 0: values, section | stage 1 ... ]
  uint64_t* full = reinterpret_cast<uint64_t*>(v2_smem);
  uint64_t* empty = full + kV2MaxStages;
  PDesc* pring = reinterpret_cast<PDesc*>(v2_smem + 128);
  unsigned char* stages = v2_smem + 128 + sizeof(PDesc) * kV2DescRing;
  const uint32_t stage_bytes = val_bytes + sec_bytes;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nk = (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // chunks blockIdx.x + k*gridDim.x
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kV2ConsumerWarps);
    }
  }
  __syncthreads();
  if (nk <= 0) return;

  if (warp == kV2ConsumerWarps) {
    // ---- producer ------------------------------------------------------------------------------------------
    // the descriptors travel kV2DescRing chunks ahead of their use (cp.async, one commit group per chunk), so the
    // producer never waits for a global load: a chunk costs it one shared-memory read and a handful of TMA issues
    auto prefetch = [&](int kk) {
      if (kk < nk && lane < 8)
        v2_cp_async16(reinterpret_cast<unsigned char*>(pring + (kk % kV2DescRing)) + lane * 16,
                      reinterpret_cast<const unsigned char*>(pdesc + ((size_t)blockIdx.x + (size_t)kk * gridDim.x)) + lane * 16);
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int kk = 0; kk < kV2DescRing; ++kk) prefetch(kk);
    // Runs beyond the ones a descriptor carries (scattered stage lists: mesh problems have ~100 short runs per chunk) are
    // read from the run list ONE CHUNK AHEAD, kV2ExtraRuns per lane, into registers: the descriptor of chunk k+1 is
    // already in the ring when chunk k is issued, so these loads overlap the wait for a free stage.
    uint2 rx[kV2ExtraRuns];
    auto load_extra = [&](int kk) {  // requires the descriptor of chunk kk to have landed in the ring
      const PDesc* pn = pring + (kk % kV2DescRing);
      const uint32_t run0n = pn->h.run0;
      const int nrn = (int)pn->h.nruns;
#pragma unroll
      for (int q = 0; q < kV2ExtraRuns; ++q) {
        const int i = kV2DescRuns + lane + 32 * q;
        rx[q] = i < nrn ? __ldg(reinterpret_cast<const uint2*>(runs + run0n) + i) : make_uint2(0u, 0u);
      }
    };
    asm volatile("cp.async.wait_group %0;" ::"n"(kV2DescRing - 1) : "memory");
    __syncwarp();
    load_extra(0);
    int s = 0;
    uint32_t eph = 1;  // parity of the `empty` phase this stage's next reuse waits for (first pass: no wait)
    for (int k = 0; k < nk; ++k) {
      asm volatile("cp.async.wait_group %0;" ::"n"(kV2DescRing - 1) : "memory");
      __syncwarp();
      const PDesc* pd = pring + (k % kV2DescRing);
      const uint4 c0 = reinterpret_cast<const uint4*>(pd)[0], c1 = reinterpret_cast<const uint4*>(pd)[1];
      uint2 rc = make_uint2(0u, 0u);
      if (lane < kV2DescRuns) rc = reinterpret_cast<const uint2*>(pd->r)[lane];
      // ChunkHdr2 fields: c0 = {blob_off lo, blob_off hi, run0, val_bytes}; c1 = {tail_src, sec16 | tail_dst<<16, nruns | nslots<<16, nstage | ndiag<<16}
      const uint64_t blob_off = (uint64_t)c0.x | ((uint64_t)c0.y << 32);
      const uint32_t run0 = c0.z, vbytes = c0.w, tail_src = c1.x;
      const uint32_t sbytes = (c1.y & 0xffffu) * 16u, tail_dst = c1.y >> 16;
      const int nruns = (int)(c1.z & 0xffffu);
      unsigned char* st = stages + (size_t)s * stage_bytes;
      double* sval = reinterpret_cast<double*>(st);
      if (k >= nstages) mbar_wait(&empty[s], eph);
      if (lane == 0) {
        if (tail_src) sval[tail_dst] = Jv[tail_src - 1u];
        mbar_expect_tx(&full[s], sbytes + vbytes);
        tma_bulk_g2s(st + val_bytes, blob + blob_off, sbytes, &full[s]);
      }
      __syncwarp();
      if (lane < kV2DescRuns && lane < nruns && (rc.y >> 16) != 0u) tma_bulk_g2s(sval + (rc.y & 0xffffu), Jv + rc.x, (rc.y >> 16) * 8u, &full[s]);
#pragma unroll
      for (int q = 0; q < kV2ExtraRuns; ++q)
        if ((rx[q].y >> 16) != 0u) tma_bulk_g2s(sval + (rx[q].y & 0xffffu), Jv + rx[q].x, (rx[q].y >> 16) * 8u, &full[s]);
      for (int i = kV2DescRuns + 32 * kV2ExtraRuns + lane; i < nruns; i += 32) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(runs + run0) + i);
        if ((r.y >> 16) != 0u) tma_bulk_g2s(sval + (r.y & 0xffffu), Jv + r.x, (r.y >> 16) * 8u, &full[s]);
      }
      if (k + 1 < nk) {
        // descriptor of chunk k+1: committed kV2DescRing-1 groups ago at the latest
        asm volatile("cp.async.wait_group %0;" ::"n"(kV2DescRing - 2) : "memory");
        __syncwarp();
        load_extra(k + 1);
      }
      __syncwarp();
      prefetch(k + kV2DescRing);
      if (++s == nstages) {
        s = 0;
        eph ^= 1u;
      }
    }
    return;
  }

  // ---- consumers -------------------------------------------------------------------------------------------
  {
    const uint32_t full_u32 = smem_u32(full), empty_u32 = smem_u32(empty), stages_u32 = smem_u32(stages);
    int s = 0;
    uint32_t ph = 0;
    for (int k = 0; k < nk; ++k) {
      const uint32_t val_u32 = stages_u32 + (uint32_t)s * stage_bytes, sec_u32 = val_u32 + val_bytes;
      mbar_wait_u32(full_u32 + 8u * s, ph);
      const uint4 hd = lds_v4(sec_u32);  // nslots | nstage<<16, ngroups | off_meta<<16, off_ginfo | off_cnt<<16, off_rec
      const int ng = (int)(hd.y & 0xffffu);
      const uint32_t stab_u32 = sec_u32 + 16u, meta_u32 = sec_u32 + (hd.y >> 16), ginfo_u32 = sec_u32 + (hd.z & 0xffffu),
                     cnt_u32 = sec_u32 + (hd.z >> 16), rec_u32 = sec_u32 + (hd.w & 0xffffu);
      // the groups of 32 slots rotate over the warps from chunk to chunk (the first groups hold the longest slots)
      for (int g = (warp - k) & (kV2ConsumerWarps - 1); g < ng; g += kV2ConsumerWarps) {
        const int t = g * 32 + lane;
        const uint32_t gi = lds_u32(ginfo_u32 + 4u * g);
        const uint32_t m = lds_u32(meta_u32 + 4u * t);
        const int c = (int)lds_u8(cnt_u32 + t);  // 0: padding lane
        const int rows = (int)(gi >> 16);
        uint32_t ra = rec_u32 + (gi & 0xffffu) * 64u + 2u * lane;
        const uint32_t r = lds_u16(ra);  // padding lanes read record 0 = entry 0 of stage column 0: harmless
        const uint2 ti = lds_v2(stab_u32 + 8u * (m & 255u)), tj = lds_v2(stab_u32 + 8u * ((m >> 8) & 255u));
        const uint32_t vi = val_u32 + ti.y, vj = val_u32 + tj.y;
        const uint32_t pos_lo = ti.x + ((m >> 16) & 255u) * 32u;
        const uint32_t pos_up = tj.x + (m >> 24) * 32u;
        double acc = __dmul_rn(lds_f64(vi + ((r >> 4) & 0x3f8u)), lds_f64(vj + ((r << 3) & 0x3f8u)));
        int d = 1;
        ra += 64u;
        // four record rows per round: records and operands are all loaded before the first dependent add
        for (; d + 3 < rows; d += 4, ra += 256u) {
          const uint32_t r0 = lds_u16(ra), r1 = lds_u16(ra + 64u), r2 = lds_u16(ra + 128u), r3 = lds_u16(ra + 192u);
          const double a0 = lds_f64(vi + ((r0 >> 4) & 0x3f8u)), b0 = lds_f64(vj + ((r0 << 3) & 0x3f8u));
          const double a1 = lds_f64(vi + ((r1 >> 4) & 0x3f8u)), b1 = lds_f64(vj + ((r1 << 3) & 0x3f8u));
          const double a2 = lds_f64(vi + ((r2 >> 4) & 0x3f8u)), b2 = lds_f64(vj + ((r2 << 3) & 0x3f8u));
          const double a3 = lds_f64(vi + ((r3 >> 4) & 0x3f8u)), b3 = lds_f64(vj + ((r3 << 3) & 0x3f8u));
          if (d < c) acc = __dadd_rn(acc, __dmul_rn(a0, b0));
          if (d + 1 < c) acc = __dadd_rn(acc, __dmul_rn(a1, b1));
          if (d + 2 < c) acc = __dadd_rn(acc, __dmul_rn(a2, b2));
          if (d + 3 < c) acc = __dadd_rn(acc, __dmul_rn(a3, b3));
        }
        if (d + 1 < rows) {
          const uint32_t r0 = lds_u16(ra), r1 = lds_u16(ra + 64u);
          const double a0 = lds_f64(vi + ((r0 >> 4) & 0x3f8u)), b0 = lds_f64(vj + ((r0 << 3) & 0x3f8u));
          const double a1 = lds_f64(vi + ((r1 >> 4) & 0x3f8u)), b1 = lds_f64(vj + ((r1 << 3) & 0x3f8u));
          if (d < c) acc = __dadd_rn(acc, __dmul_rn(a0, b0));
          if (d + 1 < c) acc = __dadd_rn(acc, __dmul_rn(a1, b1));
          d += 2;
          ra += 128u;
        }
        if (d < rows) {
          const uint32_t r0 = lds_u16(ra);
          const double a0 = lds_f64(vi + ((r0 >> 4) & 0x3f8u)), b0 = lds_f64(vj + ((r0 << 3) & 0x3f8u));
          if (d < c) acc = __dadd_rn(acc, __dmul_rn(a0, b0));
        }
        if (c > 0) {
          hval[pos_lo] = acc;
          hval[pos_up] = acc;
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_u32 + 8u * s) : "memory");
      if (++s == nstages) {
        s = 0;
        ph ^= 1u;
      }
    }
  }
}

void upload_pairs_v2(spb_context* h, const PairsV2Host& V) {
  h->v2_ok = false;
  h->d_v2_hdr.release();
  h->d_v2_runs.release();
  h->d_v2_blob.release();
  static const bool off = getenv("SPB_ASM_V2") && atoi(getenv("SPB_ASM_V2")) == 0;
  if (!V.ok || off) return;
  cudaStream_t st = h->stream;
  {
    std::vector<PDesc> pd(V.hdr.size());
    for (size_t c = 0; c < V.hdr.size(); ++c) {
      pd[c].h = V.hdr[c];
      for (int r = 0; r < kV2DescRuns; ++r) pd[c].r[r] = r < (int)V.hdr[c].nruns ? V.runs[(size_t)V.hdr[c].run0 + r] : Run2{0u, 0, 0};
    }
    SPB_CUDA(h->d_v2_hdr.alloc(pd.size() * sizeof(PDesc)));
    SPB_CUDA(cudaMemcpyAsync(h->d_v2_hdr.p, pd.data(), pd.size() * sizeof(PDesc), cudaMemcpyHostToDevice, st));
    SPB_CUDA(cudaStreamSynchronize(st));
  }
  SPB_CUDA(h->d_v2_runs.upload(V.runs, st));
  SPB_CUDA(h->d_v2_blob.upload(V.blob, st));
  SPB_CUDA(cudaStreamSynchronize(st));
  h->v2_val_bytes = (V.max_val_bytes + 127u) & ~127u;
  h->v2_sec_bytes = (V.max_sec_bytes + 127u) & ~127u;
  int dev = 0, max_smem = 0;
  SPB_CUDA(cudaGetDevice(&dev));
  SPB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const uint32_t stage = h->v2_val_bytes + h->v2_sec_bytes;
  const int fixed = 128 + (int)sizeof(PDesc) * kV2DescRing;
  int ns = stage ? (int)((max_smem - fixed - 1024) / stage) : 0;
  ns = std::min(ns, kV2MaxStages);
  ns = std::min(ns, 4);  // deeper rings measured no faster; the rest of the SM's shared memory goes to the column kernel running alongside
  if (const char* e = getenv("SPB_ASM_V2_STAGES")) ns = std::max(2, std::min(std::min((int)((max_smem - fixed - 1024) / stage), kV2MaxStages), atoi(e)));
  if (ns < 2) return;
  h->v2_stages = ns;
  h->v2_smem = (size_t)fixed + (size_t)ns * stage;
  // the limit is a property of the function, not of the handle: handles with different ring sizes share it
  SPB_CUDA(cudaFuncSetAttribute(assemble_pairs_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  h->v2_ok = true;
}

// false: the stream form does not apply to this call (the caller falls back to the general kernels)
bool run_pairs_v2(spb_context* h, const double* d_Jv, double* hv) {
  if (!h->v2_ok || (((uintptr_t)d_Jv) & 15) != 0) return false;
  const int G = std::max(1, std::min(h->nchunks, h->num_sms));
  assemble_pairs_stream_kernel<<<G, kV2Threads, h->v2_smem, h->stream>>>(reinterpret_cast<const PDesc*>(h->d_v2_hdr.p), h->nchunks, h->d_v2_runs.p, h->d_v2_blob.p, d_Jv, hv,
                                                                        h->v2_stages, h->v2_val_bytes, h->v2_sec_bytes);
  return true;
}

}  // namespace spb
