// Probe (not product): issue rate of the legacy mma.sync shapes on sm_100a -- TF32 m16n8k8 / m16n8k4,
// BF16 m16n8k16 / m16n8k8, FP16 m16n8k16 -- and of a 2:2 TF32/BF16 mix (the "1 TF32 + 2 half-rate
// BF16 correction" product considered for the conv layers).  16 warps per SM, 4 chains per warp.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define MMA_TF32K8(c) asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1))
#define MMA_TF32K4(c) asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0))
#define MMA_BF16K16(c) asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1))
#define MMA_BF16K8(c) asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0))
#define MMA_F16K16(c) asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1))
template <int KIND>
__global__ void k(float* out, int iters) {
  float c[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 11, b1 = 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (KIND == 0) MMA_TF32K8(c[i]);
      if (KIND == 1) MMA_TF32K4(c[i]);
      if (KIND == 2) MMA_BF16K16(c[i]);
      if (KIND == 3) MMA_BF16K8(c[i]);
      if (KIND == 4) MMA_F16K16(c[i]);
      if (KIND == 5) { if (i & 1) MMA_BF16K16(c[i]); else MMA_TF32K8(c[i]); }
    }
  }
  float s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND> void run(const char* name, int macs, float* out) {
  const int iters = 4000, warps = 16;
  k<KIND><<<148, warps * 32>>>(out, 10); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<KIND><<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double cyc = ms * 1e-3 * 1.965e9;
  double per_smsp = (double)iters * 4 * warps / 4.0;
  double tmacs = (double)iters * 4 * warps * 148 * macs / (ms * 1e-3) * 1e-12;
  printf("%-28s %.2f cycles per MMA per SMSP   %.1f TMAC/s = %.1f TFLOP/s\n", name, cyc / per_smsp, tmacs, 2 * tmacs);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  run<0>("tf32 m16n8k8", 16 * 8 * 8, out);
  run<1>("tf32 m16n8k4", 16 * 8 * 4, out);
  run<2>("bf16 m16n8k16", 16 * 8 * 16, out);
  run<3>("bf16 m16n8k8", 16 * 8 * 8, out);
  run<4>("f16 m16n8k16", 16 * 8 * 16, out);
  run<5>("mix tf32k8 : bf16k16 = 1:1", 16 * 8 * 12, out);
  return 0;
}
