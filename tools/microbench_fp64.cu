// tools/microbench_fp64.cu -- does the B200's FP64 pipe offer more multiplier throughput than IMAD.WIDE?
// Prints lane-ops/clk/SM for DFMA alone, DFMA next to IMAD.WIDE.U32 (different pipes?), DFMA next to
// 64-bit integer adds (the accumulate step of a 52-bit-limb split product), so that the design choice in
// DESIGN.md ("which pipe carries the 255-bit products") rests on a measurement.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int ITERS = 4096;

__global__ void k_dfma(double *out, double a, double b) {
  double acc[8];
  for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
  double x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = __fma_rz(x, y, acc[i]);
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 8 DFMA + 8 IMAD.WIDE per iteration
__global__ void k_dfma_imad(double *out, double a, double b, uint32_t u, uint32_t v) {
  double acc[8];
  unsigned long long iacc[8];
  for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i, iacc[i] = threadIdx.x + i;
  double x = a + threadIdx.x, y = b;
  uint32_t ux = u + threadIdx.x, uy = v;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      acc[i] = __fma_rz(x, y, acc[i]);
      iacc[i] = (unsigned long long)ux * uy + iacc[i];
    }
    uy += 1;
  }
  double s = 0;
  unsigned long long t = 0;
  for (int i = 0; i < 8; i++) s += acc[i], t ^= iacc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)t;
}
// split product: hi = fma_rz(a,b,C1); lo = fma_rz(a,b,C2-hi); integer-accumulate both bit patterns (2 x 64-bit adds)
__global__ void k_split(unsigned long long *out, double a, double b) {
  unsigned long long acch[4], accl[4];
  for (int i = 0; i < 4; i++) acch[i] = accl[i] = i;
  double x[4], y = b;
  for (int i = 0; i < 4; i++) x[i] = a + threadIdx.x + 3 * i;
  const double C1 = 20282409603651670423947251286016.0;          // 2^104
  const double C2 = 20282409603651674927546878656512.0;          // 2^104 + 2^52
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      double hi = __fma_rz(x[i], y, C1);
      double sub = C2 - hi;
      double lo = __fma_rz(x[i], y, sub);
      acch[i] += (unsigned long long)__double_as_longlong(hi);
      accl[i] += (unsigned long long)__double_as_longlong(lo);
    }
    y += 1.0;
  }
  unsigned long long t = 0;
  for (int i = 0; i < 4; i++) t ^= acch[i] ^ accl[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <typename K, typename... A>
static float timeit(K k, int grid, int block, A... args) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<<<grid, block>>>(args...);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<<<grid, block>>>(args...);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  double clk = prop.clockRate * 1e3;
  printf("device %s, %d SMs, max clock %.0f MHz\n", prop.name, sms, clk / 1e6);
  int grid = sms * 8, block = 256;
  void *buf;
  CK(cudaMalloc(&buf, (size_t)grid * block * 8));
  double threads = (double)grid * block;
  auto report = [&](const char *name, float ms, double ops_per_thread) {
    double ops = threads * ops_per_thread / (ms * 1e-3);
    printf("%-34s %8.3f ms  %10.3e ops/s  %7.2f lane-ops/clk/SM (at max clock)\n", name, ms, ops, ops / (sms * clk));
  };
  report("DFMA (rz)", timeit(k_dfma, grid, block, (double *)buf, 1.5, 2.5), 8.0 * ITERS);
  report("DFMA + IMAD.WIDE 1:1 (pairs)", timeit(k_dfma_imad, grid, block, (double *)buf, 1.5, 2.5, 3u, 5u), 8.0 * ITERS);
  report("split product (2 DFMA+DADD+2 IADD64)", timeit(k_split, grid, block, (unsigned long long *)buf, 1.5, 2.5), 4.0 * ITERS);
  return 0;
}
