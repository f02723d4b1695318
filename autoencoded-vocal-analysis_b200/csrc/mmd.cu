// Maximum mean discrepancy between sets of latent means (Gretton et al. 2012), the reference's
// downstream statistic on the VAE's output: ava/plotting/mmd_plots.py:255-312 (_estimate_mmd2,
// an O(n^2) Python double loop per pair of conditions) and :450-476 (estimate_median_sigma).
//
// ava_b200_mmd_block_sums: rows of x are grouped by condition (seg[i] = condition of row i);
//   S[a][b] = sum_{i in a, j in b} exp(A * ||x_i - x_j||^2)  for ALL pairs of conditions in one
//   pass over the N x N Gram matrix (fp64, direct differences as the reference computes them).
//   The whole MMD^2 matrix follows on the host from S and the group sizes.
// ava_b200_pair_kernel: out[k] = ||x[ia[k]] - x[ib[k]]||^2 (mode 0) or exp(A * that) (mode 1) for
//   explicit index pairs, summed in numpy's pairwise order (bit-exact for D = 32).
#include "common.cuh"

namespace ava {

constexpr int MMD_TILE = 64;   // 64 x 64 pairs per CTA, 4 x 4 per thread (256 threads)
constexpr int MMD_DC = 32;     // latent dimensions staged per pass (z_dim = 32: one pass)

__global__ void __launch_bounds__(256)
mmd_block_sums_kernel(const double* __restrict__ x, int N, int D, const int* __restrict__ seg, int n_seg,
                      double A, double* __restrict__ S) {
  __shared__ double sx[MMD_TILE][MMD_DC + 1], sy[MMD_TILE][MMD_DC + 1];
  __shared__ int segx[MMD_TILE], segy[MMD_TILE];
  __shared__ double red[8];
  const int i0 = blockIdx.y * MMD_TILE, j0 = blockIdx.x * MMD_TILE;
  const int tid = threadIdx.x;
  if (tid < MMD_TILE) {
    segx[tid] = (i0 + tid < N) ? seg[i0 + tid] : -1;
    segy[tid] = (j0 + tid < N) ? seg[j0 + tid] : -1;
  }
  const int ti = (tid >> 4) * 4, tj = (tid & 15) * 4;
  double d2[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) d2[a][b] = 0.0;
  for (int c0 = 0; c0 < D; c0 += MMD_DC) {
    const int dc = min(MMD_DC, D - c0);
    __syncthreads();
    for (int idx = tid; idx < MMD_TILE * dc; idx += 256) {
      const int r = idx / dc, c = idx - r * dc;
      sx[r][c] = (i0 + r < N) ? x[(size_t)(i0 + r) * D + c0 + c] : 0.0;
      sy[r][c] = (j0 + r < N) ? x[(size_t)(j0 + r) * D + c0 + c] : 0.0;
    }
    __syncthreads();
    for (int c = 0; c < dc; ++c) {
      double xa[4], yb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) xa[a] = sx[ti + a][c];
#pragma unroll
      for (int b = 0; b < 4; ++b) yb[b] = sy[tj + b][c];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const double d = xa[a] - yb[b];
          d2[a][b] = fma(d, d, d2[a][b]);
        }
    }
  }
  // a tile whose rows (and columns) all belong to one condition: one atomic per CTA
  const int last_i = min(MMD_TILE, N - i0) - 1, last_j = min(MMD_TILE, N - j0) - 1;
  const bool uniform = segx[0] == segx[last_i] && segy[0] == segy[last_j];
  if (uniform) {
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (segx[ti + a] >= 0 && segy[tj + b] >= 0) s += exp(A * d2[a][b]);
    s = warp_sum(s);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      atomicAdd(&S[(size_t)segx[0] * n_seg + segy[0]], t);
    }
  } else {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int sa = segx[ti + a], sb = segy[tj + b];
        if (sa >= 0 && sb >= 0) atomicAdd(&S[(size_t)sa * n_seg + sb], exp(A * d2[a][b]));
      }
  }
}

// numpy's pairwise summation of n < 128 terms: 8 interleaved accumulators, combined as
// ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)), then the remainder sequentially
__global__ void __launch_bounds__(256)
pair_kernel(const double* __restrict__ x, int D, const long long* __restrict__ ia, const long long* __restrict__ ib,
            long long n, double A, int mode, double* __restrict__ out) {
  const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
  if (k >= n) return;
  const double* a = x + (size_t)ia[k] * D;
  const double* b = x + (size_t)ib[k] * D;
  double res;
  if (D < 8) {
    res = 0.0;
    for (int c = 0; c < D; ++c) {
      const double d = a[c] - b[c];
      res = __dadd_rn(res, __dmul_rn(d, d));   // no FMA contraction: numpy multiplies, then adds
    }
  } else {
    double r[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double d = a[q] - b[q];
      r[q] = d * d;
    }
    int c = 8;
    for (; c + 8 <= D; c += 8) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const double d = a[c + q] - b[c + q];
        r[q] = __dadd_rn(r[q], __dmul_rn(d, d));
      }
    }
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; c < D; ++c) {
      const double d = a[c] - b[c];
      res = __dadd_rn(res, __dmul_rn(d, d));
    }
  }
  out[k] = mode ? exp(A * res) : res;
}

}  // namespace ava

using namespace ava;

extern "C" int ava_b200_mmd_block_sums(const double* x, int N, int D, const int* seg, int n_seg, double A, double* S,
                                       void* stream_) {
  AVA_REQUIRE(N >= 0 && D >= 1, "mmd_block_sums: N=%d D=%d", N, D);
  AVA_REQUIRE(n_seg >= 1, "mmd_block_sums: n_seg=%d", n_seg);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (cudaMemsetAsync(S, 0, sizeof(double) * (size_t)n_seg * n_seg, stream) != cudaSuccess) {
    set_error("mmd_block_sums: memset failed");
    return 1;
  }
  if (N == 0) return 0;
  const int nb = (N + MMD_TILE - 1) / MMD_TILE;
  mmd_block_sums_kernel<<<dim3(nb, nb), 256, 0, stream>>>(x, N, D, seg, n_seg, A, S);
  return check_launch("mmd_block_sums");
}

extern "C" int ava_b200_pair_kernel(const double* x, int D, const long long* ia, const long long* ib, long long n,
                                    double A, int mode, double* out, void* stream_) {
  AVA_REQUIRE(D >= 1 && n >= 0, "pair_kernel: D=%d n=%lld", D, n);
  if (n == 0) return 0;
  pair_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(x, D, ia, ib, n, A, mode, out);
  return check_launch("pair_kernel");
}
