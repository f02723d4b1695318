// tcgen05 / TMEM / TMA GEMM for the two 8192x1024 dense layers (fc1, fc8), sm_100a.
//
//   D[M,N] = sum_t A_t[M,K] . B_t[N,K]^T      (all operands K-major fp32, kind::tf32)
//
// precision 1 ("tf32"):   one term, operands rounded to TF32         (rtol ~1e-3)
// precision 2 ("tf32x3"): three terms A_hi.B_hi + A_hi.B_lo + A_lo.B_hi with
//                         x = x_hi + x_lo split on the fly by a pre-pass -> fp32-level
//                         accuracy (rtol ~1e-6) on the tensor cores
//
// Every GEMM of the dense layers (forward NT, backward-data NN, backward-weight TN) is
// brought to this one K-major form by the pre-pass kernel below, which reads the fp32
// operand once and writes its TF32 hi/lo parts, optionally transposed and optionally
// ReLU-masked (the backward's [y>0]).
//
// Kernel anatomy (one 128x128 output tile per CTA, 256 threads):
//   warp 0 / lane 0 : TMA producer  -- cp.async.bulk.tensor.2d of 128x32-float boxes
//                     (128-byte swizzle) into a multi-stage smem ring, mbarrier complete_tx
//   warp 1 / lane 0 : MMA issuer    -- tcgen05.mma.cta_group::1.kind::tf32, M=128 N=128 K=8,
//                     fp32 accumulator in 128 TMEM columns; tcgen05.commit frees smem stages
//   warp 2          : TMEM allocator (tcgen05.alloc / dealloc)
//   warps 4..7      : epilogue      -- tcgen05.ld 32x32b, bias + activation, transpose through
//                     smem, coalesced global stores (or split-K partials)
#include <cuda.h>

#include "common.cuh"

namespace ava {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32;  // BK floats = one 128-byte swizzle atom
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;     // 16 KB per operand tile

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);       // start address
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                             // layout: SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=128
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format = F32
  d |= 2u << 7;                    // a_format = TF32
  d |= 2u << 10;                   // b_format = TF32
  d |= (uint32_t)(N >> 3) << 17;   // n_dim
  d |= (uint32_t)(M >> 4) << 24;   // m_dim
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

struct TcParams {
  float* D;            // output (direct mode)
  float* part;         // split-K partials [splits][M][N]
  const float* bias;   // per n
  int ldd, M, N, K;
  int act;
  int splits, kblocks_per_split;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread (lane = row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
        "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// The tensor core adds into its fp32 accumulator with truncation toward zero (measured on this
// kernel, profiles/probes/tc_gemm_accuracy.py: a K = 8192 chain shrinks the large outputs by
// 1.2e-5 -- ~8e-9 per accumulation -- where the fp32 FMA kernel is at 1e-9; at batch 1024 that
// systematic shrink moved 1334 ReLU units across zero and the whole-model gradients by 5e-4).
// The K loop is therefore cut into CHUNKS of TC_CHUNK k-blocks: each chunk is accumulated from
// zero in one of two 128-column TMEM accumulators (ping-pong), and while the tensor core works
// on the next chunk the epilogue warps drain the finished one and add it to running sums in
// registers with ordinary round-to-nearest FADDs.  A chain is then at most
// TC_CHUNK * 4 * TERMS accumulations long (48 in 3 terms: < 4e-7), and the drain is hidden
// behind the MMAs of the other accumulator.
constexpr int TC_CHUNK = 4;
constexpr int TC_THREADS = 384;   // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue

template <int TERMS, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
               const TcParams P) {
  constexpr int NT_A = (TERMS == 3) ? 2 : 1;              // operand tiles per stage (hi [, lo])
  constexpr int STAGE_BYTES = 2 * NT_A * TC_TILE_BYTES;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* tiles = tc_smem;                                // [STAGES][A_hi, A_lo, B_hi, B_lo]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tc_smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;                 // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;                      // [2] epilogue -> MMA (8 warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  // epilogue transpose buffers [8 warps][32][33]: the operand ring is free by then
  float* s_tr = reinterpret_cast<float*>(tiles);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;
  const int split = blockIdx.z;
  const int kb0 = split * P.kblocks_per_split;
  const int total_kb = P.K / TC_BK;
  const int nkb = min(P.kblocks_per_split, total_kb - kb0);
  const int nchunks = (nkb + TC_CHUNK - 1) / TC_CHUNK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------ TMA producer
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      mbar_expect_tx(&full_bar[s], STAGE_BYTES);
      uint8_t* st = tiles + s * STAGE_BYTES;
      const int kx = (kb0 + i) * TC_BK;
      tma_load_2d(st, &mapA0, kx, m0, &full_bar[s]);
      if (TERMS == 3) tma_load_2d(st + TC_TILE_BYTES, &mapA1, kx, m0, &full_bar[s]);
      tma_load_2d(st + NT_A * TC_TILE_BYTES, &mapB0, kx, n0, &full_bar[s]);
      if (TERMS == 3) tma_load_2d(st + (NT_A + 1) * TC_TILE_BYTES, &mapB1, kx, n0, &full_bar[s]);
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_tf32(TC_BM, TC_BN);
    int i = 0;
    for (int q = 0; q < nchunks; ++q) {
      const int b = q & 1;
      const uint32_t use = (uint32_t)(q >> 1);
      mbar_wait(&acc_empty[b], (use & 1) ^ 1);   // the epilogue has drained this accumulator's last chunk
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(b * TC_BN);
      const int iend = min(nkb, (q + 1) * TC_CHUNK);
      bool fresh = true;
      for (; i < iend; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(tiles + s * STAGE_BYTES);
        const uint64_t a_hi = make_smem_desc(sa);
        const uint64_t a_lo = make_smem_desc(sa + TC_TILE_BYTES);
        const uint64_t b_hi = make_smem_desc(sa + NT_A * TC_TILE_BYTES);
        const uint64_t b_lo = make_smem_desc(sa + (NT_A + 1) * TC_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          const uint64_t koff = (uint64_t)((k * 8 * 4) >> 4);  // 32 bytes per K=8 step inside the atom
          if (TERMS == 3) {
            // small terms first, the full-magnitude product last
            umma_tf32(tacc, a_hi + koff, b_lo + koff, idesc, fresh ? 0u : 1u);
            umma_tf32(tacc, a_lo + koff, b_hi + koff, idesc, 1u);
            umma_tf32(tacc, a_hi + koff, b_hi + koff, idesc, 1u);
          } else {
            umma_tf32(tacc, a_hi + koff, b_hi + koff, idesc, fresh ? 0u : 1u);
          }
          fresh = false;
        }
        umma_commit(&empty_bar[s]);  // smem stage reusable once these MMAs have read it
      }
      umma_commit(&acc_full[b]);     // this chunk's accumulator is complete
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue warps: drain chunks into running sums, then
    // registers -> smem transpose -> global
    const int wq = warp & 3;          // TMEM lane quarter: rows 32*wq .. 32*wq+31 of the tile
    const int ch = (warp - 4) >> 2;   // column half: columns 64*ch .. 64*ch+63
    float rs[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) rs[j] = 0.f;
#pragma unroll 1
    for (int q = 0; q < nchunks; ++q) {
      const int b = q & 1;
      mbar_wait(&acc_full[b], (uint32_t)((q >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(b * TC_BN + ch * 64);
      uint32_t v0[32], v1[32];
      tmem_ld32(taddr, v0);
      tmem_ld32(taddr + 32, v1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);   // the tensor core may overwrite this accumulator
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        rs[j] += __uint_as_float(v0[j]);
        rs[32 + j] += __uint_as_float(v1[j]);
      }
    }
    // all MMAs have completed (the last acc_full covers them): the operand ring is free
    float* tr = s_tr + (warp - 4) * 32 * 33;
    float* dst = (P.splits > 1) ? P.part + (size_t)split * P.M * P.N : P.D;
    const int ldd = (P.splits > 1) ? P.N : P.ldd;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      // lane = row (32*wq + lane), rs[32c + j] = column 64*ch + 32*c + j  ->  smem [row][col]
#pragma unroll
      for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = rs[32 * c + j];
      __syncwarp();
      // read back with lanes along columns: coalesced 128-byte row segments
      const int col = n0 + ch * 64 + c * 32 + lane;
      const float bv = (P.splits == 1 && P.bias) ? P.bias[col] : 0.f;
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        float x = tr[r * 33 + lane] + bv;
        if (P.splits == 1) {
          if (P.act == 1) x = fmaxf(x, 0.f);
          if (P.act == 2) x = expf(x);
        }
        dst[(size_t)(m0 + wq * 32 + r) * ldd + col] = x;
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

// ------------------------------------------------------------------ pre-pass
// src [R, C] (row stride ld) -> hi (and lo) in TF32, written either as [R, C] or transposed
// as [C, R]; optional mask (same layout as src): value = mask > 0 ? src : 0.
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

template <bool TRANSPOSE, bool WANT_LO>
__global__ void __launch_bounds__(256) tc_split_kernel(const float* __restrict__ src, const float* __restrict__ mask,
                                                       int ld, int R, int C, float* __restrict__ hi,
                                                       float* __restrict__ lo) {
  __shared__ float t_hi[32][33], t_lo[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    float x = 0.f;
    if (r < R && c < C) {
      x = src[(size_t)r * ld + c];
      if (mask != nullptr && !(mask[(size_t)r * ld + c] > 0.f)) x = 0.f;
    }
    const float h = to_tf32(x);
    const float l = WANT_LO ? to_tf32(x - h) : 0.f;
    if (TRANSPOSE) {
      t_hi[ty + 8 * i][tx] = h;
      if (WANT_LO) t_lo[ty + 8 * i][tx] = l;
    } else if (r < R && c < C) {
      hi[(size_t)r * C + c] = h;
      if (WANT_LO) lo[(size_t)r * C + c] = l;
    }
  }
  if (TRANSPOSE) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty + 8 * i, r = r0 + tx;  // output row = c, output column = r
      if (r < R && c < C) {
        hi[(size_t)c * R + r] = t_hi[tx][ty + 8 * i];
        if (WANT_LO) lo[(size_t)c * R + r] = t_lo[tx][ty + 8 * i];
      }
    }
  }
}

// ------------------------------------------------------------------ host side
// K-major fp32 matrix [rows, K] (contiguous rows) -> 128x32 boxes, 128-byte swizzle
static int make_map(CUtensorMap* map, const float* base, int rows, int K) {
  TensorMapEncodeFn fn = get_tensor_map_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return 1;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
  cuuint32_t box[2] = {TC_BK, TC_BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 1;
  }
  return 0;
}

int tc_gemm_supported(int M, int N, int K) {
  return M > 0 && N > 0 && K > 0 && M % TC_BM == 0 && N % TC_BN == 0 && K % TC_BK == 0;
}

static int tc_splits(int M, int N, int K) {
  long long tiles = (long long)(M / TC_BM) * (N / TC_BN);
  int kb = K / TC_BK;
  int s = 1;
  while (tiles * s < kNumSMs && s * 2 <= kb / 8 && s < 16) s *= 2;
  return s;
}

// workspace: [A_hi | A_lo | B_hi | B_lo | split-K partials]
long long tc_ws_bytes(int M, int N, int K) {
  long long a = (long long)M * K, b = (long long)N * K;
  long long parts = (long long)tc_splits(M, N, K) * M * N;
  return (2 * a + 2 * b + parts) * (long long)sizeof(float) + 4096;
}

void launch_splitk_reduce(const float* part, int splits, int M, int N, const float* bias, float* C, int ldc, int act,
                          cudaStream_t stream);

// D[M,N] = act( A'[M,K] . B'[N,K]^T + bias ), with A' = (a_trans ? a^T : a) (.) mask etc.
//   a: [M,K] with row stride lda, or if a_trans: stored [K,M] with row stride lda
//   b: [N,K] with row stride ldb, or if b_trans: stored [K,N] with row stride ldb
int tc_gemm(const float* a, int lda, bool a_trans, const float* a_mask, const float* b, int ldb, bool b_trans,
            const float* bias, float* d, int ldd, int M, int N, int K, int act, int precision, void* ws,
            long long ws_bytes, cudaStream_t stream) {
  if (!tc_gemm_supported(M, N, K)) {
    set_error("tc_gemm: M=%d N=%d K=%d not multiples of the 128x128x32 tile", M, N, K);
    return 1;
  }
  if (ws == nullptr || ws_bytes < tc_ws_bytes(M, N, K)) {
    set_error("tc_gemm: workspace too small (%lld < %lld)", ws_bytes, tc_ws_bytes(M, N, K));
    return 1;
  }
  const bool x3 = precision == 2;
  float* a_hi = reinterpret_cast<float*>(ws);
  float* a_lo = a_hi + (size_t)M * K;
  float* b_hi = a_lo + (size_t)M * K;
  float* b_lo = b_hi + (size_t)N * K;
  float* parts = b_lo + (size_t)N * K;
  // ---- pre-pass: TF32 hi/lo parts in K-major layout
  {
    // A: source is [M,K] (or [K,M] when transposed)
    const int R = a_trans ? K : M, C = a_trans ? M : K;
    dim3 grid((C + 31) / 32, (R + 31) / 32);
    if (a_trans) {
      if (x3) tc_split_kernel<true, true><<<grid, 256, 0, stream>>>(a, a_mask, lda, R, C, a_hi, a_lo);
      else tc_split_kernel<true, false><<<grid, 256, 0, stream>>>(a, a_mask, lda, R, C, a_hi, a_lo);
    } else {
      if (x3) tc_split_kernel<false, true><<<grid, 256, 0, stream>>>(a, a_mask, lda, R, C, a_hi, a_lo);
      else tc_split_kernel<false, false><<<grid, 256, 0, stream>>>(a, a_mask, lda, R, C, a_hi, a_lo);
    }
    if (check_launch("tc_split(A)")) return 1;
  }
  {
    const int R = b_trans ? K : N, C = b_trans ? N : K;
    dim3 grid((C + 31) / 32, (R + 31) / 32);
    if (b_trans) {
      if (x3) tc_split_kernel<true, true><<<grid, 256, 0, stream>>>(b, nullptr, ldb, R, C, b_hi, b_lo);
      else tc_split_kernel<true, false><<<grid, 256, 0, stream>>>(b, nullptr, ldb, R, C, b_hi, b_lo);
    } else {
      if (x3) tc_split_kernel<false, true><<<grid, 256, 0, stream>>>(b, nullptr, ldb, R, C, b_hi, b_lo);
      else tc_split_kernel<false, false><<<grid, 256, 0, stream>>>(b, nullptr, ldb, R, C, b_hi, b_lo);
    }
    if (check_launch("tc_split(B)")) return 1;
  }
  // ---- tensor maps
  CUtensorMap mA0, mA1, mB0, mB1;
  if (make_map(&mA0, a_hi, M, K) || make_map(&mA1, x3 ? a_lo : a_hi, M, K) || make_map(&mB0, b_hi, N, K) ||
      make_map(&mB1, x3 ? b_lo : b_hi, N, K))
    return 1;
  TcParams P;
  P.D = d;
  P.part = parts;
  P.bias = bias;
  P.ldd = ldd;
  P.M = M;
  P.N = N;
  P.K = K;
  P.act = act;
  P.splits = tc_splits(M, N, K);
  const int kb = K / TC_BK;
  P.kblocks_per_split = (kb + P.splits - 1) / P.splits;
  dim3 grid(N / TC_BN, M / TC_BM, P.splits);
  constexpr int EXTRA = 1024;  // barriers + tmem slot (the epilogue transpose reuses the operand ring)
  if (x3) {
    constexpr int ST = 3;
    const size_t smem = (size_t)ST * 4 * TC_TILE_BYTES + EXTRA;
    static bool cfg = false;
    if (!cfg) {
      cudaFuncSetAttribute(tc_gemm_kernel<3, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cfg = true;
    }
    tc_gemm_kernel<3, ST><<<grid, TC_THREADS, smem, stream>>>(mA0, mA1, mB0, mB1, P);
  } else {
    constexpr int ST = 6;
    const size_t smem = (size_t)ST * 2 * TC_TILE_BYTES + EXTRA;
    static bool cfg = false;
    if (!cfg) {
      cudaFuncSetAttribute(tc_gemm_kernel<1, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cfg = true;
    }
    tc_gemm_kernel<1, ST><<<grid, TC_THREADS, smem, stream>>>(mA0, mA1, mB0, mB1, P);
  }
  if (check_launch("tc_gemm")) return 1;
  if (P.splits > 1) launch_splitk_reduce(parts, P.splits, M, N, bias, d, ldd, act, stream);
  return 0;
}

}  // namespace ava
