// Microbenchmark: peak FP64 tensor-core throughput of one B200 (DMMA: mma.sync.aligned.m8n8k4 f64), the
// denominator of the supernode-update roofline (MEASURED_PEAKS.json has no FP64 number; tcgen05 has no FP64 path).
// Every warp keeps ACC independent accumulator tiles in registers and issues DMMAs back to back, operands in
// registers, no memory traffic.  One m8n8k4 DMMA = 8*8*4*2 = 512 flops per warp instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak scripts/micro/dmma_peak.cu && ./dmma_peak [json-out]
// Also measures plain DFMA throughput (the non-tensor FP64 pipe) for comparison.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ACC>
__global__ void __launch_bounds__(256) dmma_kernel(int iters, double* out) {
  double c0[ACC], c1[ACC];
#pragma unroll
  for (int k = 0; k < ACC; ++k) c0[k] = c1[k] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ACC; ++k) dmma(c0[k], c1[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < ACC; ++k) s += c0[k] + c1[k];
  if (s == 1.2345) out[0] = s;
}

template <int ACC>
__global__ void __launch_bounds__(256) dfma_kernel(int iters, double* out) {
  double c[ACC];
#pragma unroll
  for (int k = 0; k < ACC; ++k) c[k] = k;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ACC; ++k) c[k] = fma(c[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < ACC; ++k) s += c[k];
  if (s == 1.2345) out[0] = s;
}

template <class K>
static double time_ms(K launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();  // warm-up
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  double* out;
  cudaMalloc(&out, 8);
  const int sms = prop.multiProcessorCount, iters = 20000;
  double best_dmma = 0.0;
  int best_bps = 0, best_acc = 0;
  printf("%s: %d SMs\n", prop.name, sms);
  for (int bps = 1; bps <= 4; ++bps) {
    const int grid = sms * bps;
    const double warps = (double)grid * 8;
    double t4 = time_ms([&] { dmma_kernel<4><<<grid, 256>>>(iters, out); });
    double t8 = time_ms([&] { dmma_kernel<8><<<grid, 256>>>(iters, out); });
    double tf4 = warps * iters * 4 * 512.0 / (t4 * 1e-3) / 1e12, tf8 = warps * iters * 8 * 512.0 / (t8 * 1e-3) / 1e12;
    printf("DMMA m8n8k4  blocks/SM=%d  4 acc: %.2f TFLOP/s   8 acc: %.2f TFLOP/s\n", bps, tf4, tf8);
    if (tf4 > best_dmma) best_dmma = tf4, best_bps = bps, best_acc = 4;
    if (tf8 > best_dmma) best_dmma = tf8, best_bps = bps, best_acc = 8;
  }
  double best_dfma = 0.0;
  for (int bps = 2; bps <= 8; bps *= 2) {
    const int grid = sms * bps;
    double t = time_ms([&] { dfma_kernel<8><<<grid, 256>>>(iters, out); });
    double tf = (double)grid * 256 * iters * 8 * 2.0 / (t * 1e-3) / 1e12;
    printf("DFMA         blocks/SM=%d: %.2f TFLOP/s\n", bps, tf);
    if (tf > best_dfma) best_dfma = tf;
  }
  if (argc > 1) {
    FILE* f = fopen(argv[1], "w");
    fprintf(f, "{\"gpu\": \"%s\", \"sms\": %d, \"dmma_m8n8k4_tflops\": %.3f, \"dmma_best_blocks_per_sm\": %d, \"dmma_best_accumulators\": %d, \"dfma_tflops\": %.3f, "
               "\"how\": \"scripts/micro/dmma_peak.cu: register-resident mma.sync m8n8k4 f64 back to back, best of 5, CUDA events\"}\n",
            prop.name, sms, best_dmma, best_bps, best_acc, best_dfma);
    fclose(f);
  }
  printf("PEAK dmma=%.2f TFLOP/s dfma=%.2f TFLOP/s\n", best_dmma, best_dfma);
  return 0;
}
