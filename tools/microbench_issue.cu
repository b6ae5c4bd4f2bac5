// tools/microbench_issue.cu -- do ALU instructions (IADD3 / IADD3.X carry chains / LOP3) issue in the shadow
// of IMAD.WIDE.U32 on sm_100a, or do they serialise?  For R = 0..8 ALU ops per wide IMAD this prints the
// cycles per (wide IMAD + R ALU) group per SM sub-partition.  If the time stays flat up to R ~ 3 the FMA pipe
// hides ALU work (trading multiplies for adds, e.g. Karatsuba, pays); if it grows by ~1 cycle per ALU op it
// does not.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int ITERS = 2048;
constexpr int W = 8;  // independent wide-IMAD chains per thread

// KIND 0: plain adds (IADD3), 1: carry chains (add.cc / addc.cc), 2: LOP3 xor/and mix
template <int R, int KIND>
__global__ void k_mix(uint64_t *out, uint32_t a, uint32_t b) {
  uint64_t acc[W];
  uint32_t alu[W];
#pragma unroll
  for (int i = 0; i < W; i++) acc[i] = threadIdx.x + i, alu[i] = threadIdx.x * 3 + i;
  uint32_t x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < W; i++) {
      acc[i] = (uint64_t)x * (uint32_t)acc[i] + acc[i];  // IMAD.WIDE.U32, dependent on its own chain
      if (KIND == 0) {
#pragma unroll
        for (int r = 0; r < R; r++) alu[(i + r) % W] += alu[(i + r + 3) % W] + y;
      } else if (KIND == 1) {
        if (R > 0) {
          asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(alu[i]) : "r"(y));
#pragma unroll
          for (int r = 1; r < R; r++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(alu[(i + r) % W]) : "r"(x));
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; r++) alu[(i + r) % W] = (alu[(i + r) % W] ^ alu[(i + r + 3) % W]) & (y | alu[(i + r + 5) % W]);
      }
    }
    y += 1;
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < W; i++) s ^= acc[i] + alu[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int R, int KIND>
static int run(uint64_t *buf, int sms, double clk, int warps_per_smsp) {
  const int block = 128 * warps_per_smsp;  // 4 SMSPs x warps
  const int grid = sms;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k_mix<R, KIND><<<grid, block>>>(buf, 3, 5);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  k_mix<R, KIND><<<grid, block>>>(buf, 3, 5);
  cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1));
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double groups_per_smsp = (double)ITERS * W * warps_per_smsp;  // warp-level groups issued by one SMSP
  double cyc = ms * 1e-3 * clk / groups_per_smsp;
  printf("kind %d  R=%d ALU per wide IMAD, %d warps/SMSP: %6.2f cycles per group per SMSP\n", KIND, R, warps_per_smsp, cyc);
  return 0;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  double clk = prop.clockRate * 1e3;
  printf("device %s, %d SMs, max clock %.0f MHz (cycles assume max clock)\n", prop.name, sms, clk / 1e6);
  uint64_t *buf;
  CK(cudaMalloc(&buf, (size_t)sms * 1024 * 8));
  for (int w : {4, 8}) {
    run<0, 0>(buf, sms, clk, w);
    run<1, 0>(buf, sms, clk, w);
    run<2, 0>(buf, sms, clk, w);
    run<3, 0>(buf, sms, clk, w);
    run<4, 0>(buf, sms, clk, w);
    run<6, 0>(buf, sms, clk, w);
    run<8, 0>(buf, sms, clk, w);
    run<1, 1>(buf, sms, clk, w);
    run<2, 1>(buf, sms, clk, w);
    run<3, 1>(buf, sms, clk, w);
    run<4, 1>(buf, sms, clk, w);
    run<6, 1>(buf, sms, clk, w);
    run<2, 2>(buf, sms, clk, w);
    run<4, 2>(buf, sms, clk, w);
  }
  return 0;
}
