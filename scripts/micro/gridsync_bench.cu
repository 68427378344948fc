// Microbenchmark: cost of a grid-wide barrier on B200 for different (blocks/SM, threads/block),
// cooperative_groups grid.sync() vs a hand-rolled sense-reversing barrier.
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
__global__ void k_cg(int iters, double* out) {
  cg::grid_group g = cg::this_grid();
  double a = threadIdx.x;
  for (int i = 0; i < iters; ++i) { a = a * 1.0000001 + 1.0; g.sync(); }
  if (a == 12345.678) out[0] = a;
}
__device__ unsigned int g_count = 0;
__device__ volatile unsigned int g_gen = 0;
__device__ __forceinline__ void my_barrier(unsigned int nblocks, unsigned int& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int target = gen + 1;
    if (atomicAdd(&g_count, 1u) == nblocks - 1) {
      g_count = 0;
      __threadfence();
      g_gen = target;
    } else {
      while (g_gen != target) {}
    }
    __threadfence();
  }
  gen += 1;
  __syncthreads();
}
__global__ void k_my(int iters, double* out) {
  unsigned int gen = g_gen;
  double a = threadIdx.x;
  for (int i = 0; i < iters; ++i) { a = a * 1.0000001 + 1.0; my_barrier(gridDim.x, gen); }
  if (a == 12345.678) out[0] = a;
}
int main() {
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  double* out; cudaMalloc(&out, 8);
  int iters = 2000;
  int cfg[][2] = {{1, 256}, {1, 1024}, {2, 512}, {4, 256}, {8, 256}, {2, 1024}};
  for (auto& c : cfg) {
    int blocks = p.multiProcessorCount * c[0], threads = c[1];
    for (int which = 0; which < 2; ++which) {
      void* args[] = {&iters, &out};
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      void* fn = which == 0 ? (void*)k_cg : (void*)k_my;
      cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(threads), args, 0, 0);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      cudaError_t err = cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(threads), args, 0, 0);
      cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("%s blocks/SM=%d threads=%d blocks=%d : %.3f us per barrier (%s)\n", which == 0 ? "cg::grid.sync" : "hand-rolled  ", c[0], threads, blocks,
             1e3 * ms / iters, cudaGetErrorString(err));
    }
  }
  return 0;
}
