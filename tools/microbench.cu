// tools/microbench.cu -- integer-pipe ceilings on the B200 for the field arithmetic in lcpc_b200/csrc/field.cuh.
// Prints ops/clk/SM for IMAD.WIDE.U32, IMAD, IADD3 and whole Montgomery multiplications per second,
// so NTT / SpMM kernels can be judged against the pipe that actually bounds them.
#include <cstdio>
#include <cuda_runtime.h>
#include "../lcpc_b200/csrc/field.cuh"
using namespace lcpc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void k_imad_wide(uint64_t *out, uint32_t a, uint32_t b) {
  uint64_t acc[8];
  for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
  uint32_t x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = (uint64_t)x * y + acc[i];  // IMAD.WIDE.U32
    y += 1;
  }
  uint64_t s = 0;
  for (int i = 0; i < 8; i++) s ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad(uint32_t *out, uint32_t a, uint32_t b) {
  uint32_t acc[8];
  for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
  uint32_t x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = acc[i] * x + y;
  }
  uint32_t s = 0;
  for (int i = 0; i < 8; i++) s ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd3(uint32_t *out, uint32_t a, uint32_t b) {
  uint32_t acc[8];
  for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
  uint32_t x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = (acc[i] ^ x) + y;  // LOP3 + IADD3, dependent
  }
  uint32_t s = 0;
  for (int i = 0; i < 8; i++) s ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// wide IMAD and ALU ops interleaved 1:2 -- do the two pipes overlap?
__global__ void k_mixed(uint64_t *out, uint32_t a, uint32_t b) {
  uint64_t acc[4];
  uint32_t alu[4];
  for (int i = 0; i < 4; i++) acc[i] = threadIdx.x + i, alu[i] = i;
  uint32_t x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      acc[i] = (uint64_t)(uint32_t)acc[i] * y + acc[i];
      alu[i] = (alu[i] ^ x) + y;
    }
  }
  uint64_t s = 0;
  for (int i = 0; i < 4; i++) s ^= acc[i] + alu[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int FID, int ILP>
__global__ void k_montmul(uint32_t *out, const uint32_t *in, int iters) {
  using F = Field<FID>;
  typename F::Elem x[ILP], w;
  for (int k = 0; k < ILP; k++)
    for (int i = 0; i < F::N; i++) x[k].v[i] = in[i] + threadIdx.x + k;
  for (int i = 0; i < F::N; i++) w.v[i] = in[F::N + i];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = F::mul(x[k], w);
  }
  uint32_t s = 0;
  for (int k = 0; k < ILP; k++)
    for (int i = 0; i < F::N; i++) s ^= x[k].v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int FID>
__global__ void k_butterfly(uint32_t *out, const uint32_t *in, int iters) {
  using F = Field<FID>;
  typename F::Elem a, b, w;
  for (int i = 0; i < F::N; i++) a.v[i] = in[i] + threadIdx.x, b.v[i] = in[i] ^ threadIdx.x, w.v[i] = in[F::N + i];
  a.v[F::N - 1] &= 0x3fffffff, b.v[F::N - 1] &= 0x3fffffff;
  for (int it = 0; it < iters; it++) {
    typename F::Elem s = F::add(a, b), d = F::sub(a, b);
    a = s, b = F::mul(d, w);
  }
  uint32_t s = 0;
  for (int i = 0; i < F::N; i++) s ^= a.v[i] ^ b.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("device %s, %d SMs, max clock %.0f MHz\n", prop.name, sms, clk_khz / 1e3);
  void *out;
  uint32_t *in;
  CK(cudaMalloc(&out, 64 << 20));
  CK(cudaMalloc(&in, 256));
  uint32_t hin[16] = {0x12345677, 0x23456789, 0x3456789a, 0x456789ab, 0x56789abc, 0x6789abcd, 0x789abcde, 0x0123456,
                      0x1f345677, 0x2f456789, 0x3f56789a, 0x4f6789ab, 0x5f789abc, 0x6f89abcd, 0x7f9abcde, 0x0f23456};
  CK(cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256;
  const double lanes = (double)blocks * threads;
  auto report = [&](const char *name, double ops_per_thread, float ms) {
    double ops = ops_per_thread * lanes;
    double per_clk_sm = ops / (ms * 1e-3) / (clk_khz * 1e3) / sms;
    printf("%-28s %8.3f ms  %10.3e ops/s  %7.2f lane-ops/clk/SM (at max clock)\n", name, ms, ops / (ms * 1e-3), per_clk_sm);
  };
  float ms;
#define RUN(name, ops, ...)                      \
  for (int rep = 0; rep < 2; rep++) {            \
    CK(cudaEventRecord(e0));                     \
    __VA_ARGS__;                                 \
    CK(cudaEventRecord(e1));                     \
    CK(cudaEventSynchronize(e1));                \
    CK(cudaGetLastError());                      \
    CK(cudaEventElapsedTime(&ms, e0, e1));       \
  }                                              \
  report(name, ops, ms);
  RUN("IMAD.WIDE.U32", 8.0 * ITERS, (k_imad_wide<<<blocks, threads>>>((uint64_t *)out, 3, 5)));
  RUN("IMAD (32-bit)", 8.0 * ITERS, (k_imad<<<blocks, threads>>>((uint32_t *)out, 3, 5)));
  RUN("ALU LOP3+IADD3 pairs", 16.0 * ITERS, (k_iadd3<<<blocks, threads>>>((uint32_t *)out, 3, 5)));
  RUN("mixed 4 wide + 8 ALU per it", 12.0 * ITERS, (k_mixed<<<blocks, threads>>>((uint64_t *)out, 3, 5)));
  const int MI = 2000;
  RUN("montmul Ft255 ILP1", (double)MI, (k_montmul<FT255, 1><<<blocks, threads>>>((uint32_t *)out, in, MI)));
  RUN("montmul Ft255 ILP2", 2.0 * MI, (k_montmul<FT255, 2><<<blocks, threads>>>((uint32_t *)out, in, MI)));
  RUN("montmul Ft127 ILP1", (double)MI, (k_montmul<FT127, 1><<<blocks, threads>>>((uint32_t *)out, in, MI)));
  RUN("montmul Ft127 ILP2", 2.0 * MI, (k_montmul<FT127, 2><<<blocks, threads>>>((uint32_t *)out, in, MI)));
  RUN("montmul Ft63 ILP2", 2.0 * MI, (k_montmul<FT63, 2><<<blocks, threads>>>((uint32_t *)out, in, MI)));
  RUN("butterfly Ft255", (double)MI, (k_butterfly<FT255><<<blocks, threads>>>((uint32_t *)out, in, MI)));
  RUN("butterfly Ft127", (double)MI, (k_butterfly<FT127><<<blocks, threads>>>((uint32_t *)out, in, MI)));
  // occupancy sweep for the Ft255 multiply: 128..1024 threads/SM
  for (int tpb : {64, 128, 256, 512}) {
    const int b2 = sms * 2;
    CK(cudaEventRecord(e0));
    k_montmul<FT255, 1><<<b2, tpb>>>((uint32_t *)out, in, MI);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("montmul Ft255, %4d threads/SM: %.3e mul/s\n", 2 * tpb, (double)b2 * tpb * MI / (ms * 1e-3));
  }
  return 0;
}
