// Probe (not product): tcgen05 kind::tf32 with NO-swizzle K-major descriptors whose start address is
// shifted in 16-byte steps (the "flat shift" implicit-GEMM conv idea), small N, and the legacy
// mma.sync tf32 rate.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nW_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra W_DONE;\nbra W_LOOP;\nW_DONE:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;  // swizzle none
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// A planes: [CQ][NPOS][4] floats; B: [CQ][N][4] floats.  D[m][n] = sum_c A[c/4][m + shift][c%4] * B[c/4][n][c%4]
template <int N>
__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* Bm, float* D, int CQ, int NPOS, int shift,
                                                    int reps, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* sA = reinterpret_cast<float*>(smem);
  float* sB = sA + CQ * NPOS * 4;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + CQ * N * 4);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < CQ * NPOS * 4; i += 128) sA[i] = A[i];
  for (int i = threadIdx.x; i < CQ * N * 4; i += 128) sB[i] = Bm[i];
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // generic-proxy smem writes -> visible to the async proxy (tensor core reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, N);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int ks = 0; ks < CQ / 2; ++ks) {
        uint64_t da = make_desc(smem_u32(sA) + (2 * ks) * NPOS * 16 + shift * 16, NPOS * 16, 128);
        uint64_t db = make_desc(smem_u32(sB) + (2 * ks) * N * 16, N * 16, 128);
        umma(tmem, da, db, idesc, (r == 0 && ks == 0) ? 0u : 1u);
      }
    }
    commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (cycles) *cycles = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = 0; c < N; c += 8) {
      uint32_t v[8];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * N + c + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(32));
  }
}

static float tf32r(float x) {  // truncate to tf32 (tensor core ignores the low 13 bits)
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

template <int N>
static void run_probe(int CQ, int shift, int reps) {
  const int NPOS = 160;
  std::vector<float> A(CQ * NPOS * 4), B(CQ * N * 4), D(128 * N), R(128 * N);
  srand(1);
  for (auto& v : A) v = tf32r((rand() % 2001 - 1000) / 500.f);
  for (auto& v : B) v = tf32r((rand() % 2001 - 1000) / 500.f);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int c = 0; c < CQ * 4; ++c) s += (double)A[((c / 4) * NPOS + m + shift) * 4 + c % 4] * B[((c / 4) * N + n) * 4 + c % 4];
      R[m * N + n] = (float)(s * reps);
    }
  float *dA, *dB, *dD;
  long long* dC;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMalloc(&dC, 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  size_t smem = (size_t)(CQ * NPOS * 4 + CQ * N * 4) * 4 + 64;
  probe_kernel<N><<<1, 128, smem>>>(dA, dB, dD, CQ, NPOS, shift, reps, dC);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < D.size(); ++i) {
    maxerr = fmax(maxerr, fabs((double)D[i] - R[i]));
    maxref = fmax(maxref, fabs((double)R[i]));
  }
  printf("probe N=%d CQ=%d shift=%d reps=%d: %s maxerr %.3g (maxref %.3g) cycles %lld (%.2f per MMA)\n", N, CQ, shift, reps,
         cudaGetErrorString(e), maxerr, maxref, cyc, (double)cyc / (reps * (CQ / 2)));
  if (maxerr > 1e-3 * maxref && reps == 1) {
    for (int m = 0; m < 4; ++m) {
      for (int n = 0; n < N && n < 8; ++n) printf(" %8.3f/%8.3f", D[m * N + n], R[m * N + n]);
      printf("\n");
    }
  }
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  cudaFree(dC);
}


template <int N>
__global__ void __launch_bounds__(128) rate_kernel(int nacc, int reps, int kper, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* sA = reinterpret_cast<float*>(smem);          // 16 planes x 160 pos x 16 B
  float* sB = sA + 16 * 160 * 4;                         // 16 planes x N x 16 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 16 * N * 4);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 16 * 160 * 4 + 16 * N * 4; i += 128) sA[i] = 0.001f * (i % 97);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, N);
    const uint64_t da0 = make_desc(smem_u32(sA), 160 * 16, 128);
    const uint64_t db0 = make_desc(smem_u32(sB), N * 16, 128);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int k = 0; k < kper; ++k) {
        for (int a = 0; a < nacc; ++a) {
          umma(tmem + a * N, da0 + (uint64_t)(a + k), db0 + (uint64_t)((k & 7) * 2 * N), idesc, (r | k) ? 1u : 0u);
        }
      }
    }
    commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    *cycles = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}
template <int N>
static void run_rate(int nacc, int reps, int kper) {
  long long* dC;
  cudaMalloc(&dC, 8);
  size_t smem = (size_t)(16 * 160 * 4 + 16 * N * 4) * 4 + 64;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rate_kernel<N><<<1, 128, smem>>>(nacc, reps, kper, dC);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
  printf("rate N=%d nacc=%d kper=%d reps=%d: %s  %.2f cycles per MMA\n", N, nacc, kper, reps, cudaGetErrorString(e),
         (double)cyc / ((double)reps * kper * nacc));
  cudaFree(dC);
}

// ---------------------------------------------------------------- legacy mma.sync tf32 rate
__global__ void __launch_bounds__(256) mma_sync_rate(float* out, int iters) {
  float c[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 11, b1 = 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  run_probe<16>(2, 0, 1);
  run_probe<16>(2, 3, 1);
  run_probe<16>(4, 19, 1);
  run_probe<8>(2, 5, 1);
  run_probe<24>(4, 7, 1);
  run_probe<32>(6, 7, 1);
  run_rate<16>(1, 1000, 1);
  run_rate<16>(1, 100, 27);
  run_rate<16>(4, 100, 27);
  run_rate<16>(8, 100, 27);
  run_rate<16>(16, 100, 27);
  run_rate<32>(1, 100, 27);
  run_rate<32>(4, 100, 27);
  run_rate<32>(8, 100, 27);
  run_rate<128>(1, 100, 27);
  run_rate<128>(2, 100, 27);
  run_rate<256>(1, 100, 27);
  run_rate<8>(8, 100, 27);
  {
    float* out;
    const int grid = 148 * 4, iters = 20000;
    cudaMalloc(&out, grid * 256 * 4);
    mma_sync_rate<<<grid, 256>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mma_sync_rate<<<grid, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double macs = (double)grid * 8 * iters * 4 * 16 * 8 * 8;
    printf("mma.sync m16n8k8 tf32: %.3f ms, %.1f TMAC/s = %.1f TFLOP/s (%.0f MAC/clk/SM at 1.965 GHz)\n", ms, macs / ms / 1e9,
           2 * macs / ms / 1e9, macs / (ms * 1e-3) / 148 / 1.965e9);
  }
  return 0;
}
