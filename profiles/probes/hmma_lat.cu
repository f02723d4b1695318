// Probe (not product): mma.sync m16n8k8 tf32 latency / throughput vs independent chains per warp and warps per SM
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
template <int CH>
__global__ void k(float* out, int iters) {
  float c[CH][4];
  for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 11, b1 = 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH> void run(int warps_per_sm, float* out) {
  const int iters = 4000;
  k<CH><<<148, warps_per_sm * 32>>>(out, 10); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<CH><<<148, warps_per_sm * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double cyc = ms * 1e-3 * 1.965e9;
  double per_smsp_mma = (double)iters * CH * warps_per_sm / 4.0;
  printf("chains/warp %d warps/SM %2d: %.1f cycles per HMMA per SMSP; per-chain step %.1f cycles\n", CH, warps_per_sm, cyc / per_smsp_mma, cyc / iters);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  run<1>(4, out); run<2>(4, out); run<4>(4, out); run<8>(4, out);
  run<1>(8, out); run<2>(8, out); run<4>(8, out);
  run<1>(16, out); run<2>(16, out); run<4>(16, out); run<4>(32, out);
  return 0;
}
