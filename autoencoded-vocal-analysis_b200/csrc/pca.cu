// PCA of the latent means: replaces
//   sklearn.decomposition.PCA(n_components=2, copy=False, random_state=42).fit_transform(latent)
// as called by ava/data/data_container.py:538-551 (_make_latent_mean_pca_projection) on the
// [N, z_dim] float64 array get_latent returns.  For N > 500 rows scikit-learn (1.5+) takes its
// "covariance_eigh" route: mean, C = (X^T X - N m m^T)/(N-1), eigh(C), eigenvalues clipped at 0
// and sorted descending, component signs fixed by svd_flip(u_based_decision=False) (the entry
// of largest magnitude of every component is positive), projection (X - m) . V^T.
//
// Here: ONE pass over X builds the Gram matrix of (x - s) and the column sums, s being the mean
// of the first rows (a shift close to the mean, so the N m m^T correction does not cancel); the
// z x z symmetric eigenproblem is solved on the device by Jacobi rotations (round-robin order) in shared
// memory; a second pass projects.  All fp64; every reduction has a fixed order (bitwise
// reproducible).  Input rows may be fp64 or the fp32 latents exactly as the encoder wrote them.
//
// Bound: HBM.  Algorithmic bytes: fit N*D*sizeof(in); transform N*(D*sizeof(in) + K*8).
#include "common.cuh"

namespace ava {

constexpr int PCA_ROWS = 64;       // rows per tile
constexpr int PCA_THREADS = 256;
constexpr int PCA_MAX_CTAS = 2 * kNumSMs;
constexpr int PCA_SHIFT_ROWS = 1024;
constexpr int PCA_MAX_D = 64;

template <typename T>
__global__ void pca_shift_kernel(const T* __restrict__ x, long long N, int D, double* __restrict__ shift) {
  const int d = threadIdx.x;
  if (d >= D) return;
  const long long R = N < PCA_SHIFT_ROWS ? N : PCA_SHIFT_ROWS;
  double s = 0.0;
  for (long long r = 0; r < R; ++r) s += (double)x[r * D + d];
  shift[d] = s / (double)R;
}

// partial[cta][DP*DP + DP]: Gram block of the shifted rows this CTA owns, then their column sums.
// A group of (DP/4)^2 threads covers the DP x DP outputs in 4x4 register blocks; the 256/(that)
// groups of a CTA take alternate rows of a tile and are combined in a fixed order at the end.
template <typename T, int DP>
__global__ void __launch_bounds__(PCA_THREADS, 2)
pca_gram_kernel(const T* __restrict__ x, long long N, int D, const double* __restrict__ shift,
                double* __restrict__ partial) {
  constexpr int NB = DP / 4;
  constexpr int GS = NB * NB;                 // threads per group
  constexpr int NG = PCA_THREADS / GS;        // groups per CTA
  constexpr int XS_DOUBLES = PCA_ROWS * DP;
  constexpr int RED_DOUBLES = NG * DP * DP;
  constexpr int SM_DOUBLES = XS_DOUBLES > RED_DOUBLES ? XS_DOUBLES : RED_DOUBLES;
  __shared__ __align__(16) double smem[SM_DOUBLES];
  __shared__ double sshift[DP];
  const int tid = threadIdx.x;
  const int group = tid / GS, gt = tid % GS;
  const int bi = gt / NB, bj = gt % NB;
  if (tid < DP) sshift[tid] = tid < D ? shift[tid] : 0.0;
  double acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[p][q] = 0.0;
  double csum[4] = {0.0, 0.0, 0.0, 0.0};   // column sums, kept by the diagonal-block threads
  const long long n_tiles = (N + PCA_ROWS - 1) / PCA_ROWS;
  // the next tile's elements travel in registers while the current tile is multiplied
  constexpr int EPT = PCA_ROWS * DP / PCA_THREADS;   // elements per thread per tile
  double pre[EPT];
  auto fetch = [&](long long t) {
    const long long row0 = t * PCA_ROWS;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int idx = tid + e * PCA_THREADS;
      const int r = idx / DP, c = idx - r * DP;
      pre[e] = (c < D && row0 + r < N) ? (double)x[(row0 + r) * D + c] - sshift[c] : 0.0;
    }
  };
  __syncthreads();   // sshift
  if ((long long)blockIdx.x < n_tiles) fetch(blockIdx.x);
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < EPT; ++e) smem[tid + e * PCA_THREADS] = pre[e];
    __syncthreads();
    if (t + gridDim.x < n_tiles) fetch(t + gridDim.x);
    for (int r = group; r < PCA_ROWS; r += NG) {
      const double* xr = smem + r * DP;
      const double2 a01 = *reinterpret_cast<const double2*>(xr + 4 * bi);
      const double2 a23 = *reinterpret_cast<const double2*>(xr + 4 * bi + 2);
      const double2 b01 = *reinterpret_cast<const double2*>(xr + 4 * bj);
      const double2 b23 = *reinterpret_cast<const double2*>(xr + 4 * bj + 2);
      const double a[4] = {a01.x, a01.y, a23.x, a23.y};
      const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fma(a[p], b[q], acc[p][q]);
      if (bi == bj) {
#pragma unroll
        for (int p = 0; p < 4; ++p) csum[p] += a[p];
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) smem[group * DP * DP + (4 * bi + p) * DP + 4 * bj + q] = acc[p][q];
  __syncthreads();
  double* out = partial + (size_t)blockIdx.x * (DP * DP + DP);
  for (int idx = tid; idx < DP * DP; idx += PCA_THREADS) {
    double s = 0.0;
#pragma unroll
    for (int g = 0; g < NG; ++g) s += smem[g * DP * DP + idx];
    out[idx] = s;
  }
  // column sums: group g's diagonal thread bi holds columns 4*bi..4*bi+3
  __syncthreads();
  if (bi == bj) {
#pragma unroll
    for (int p = 0; p < 4; ++p) smem[group * DP + 4 * bi + p] = csum[p];
  }
  __syncthreads();
  if (tid < DP) {
    double s = 0.0;
#pragma unroll
    for (int g = 0; g < NG; ++g) s += smem[g * DP + tid];
    out[DP * DP + tid] = s;
  }
}

// Sum the per-CTA partials in CTA order (one thread per entry, many CTAs): sums[DP*DP + DP]
__global__ void __launch_bounds__(128)
pca_reduce_kernel(const double* __restrict__ partial, int n_parts, int stride, double* __restrict__ sums) {
  const int e = blockIdx.x * 128 + threadIdx.x;
  if (e >= stride) return;
  double s = 0.0;
  for (int k = 0; k < n_parts; ++k) s += partial[(size_t)k * stride + e];
  sums[e] = s;
}

// mean = shift + colsum/N;  cov = (G - colsum colsum^T / N) / (N - 1)   [D x D, row-major, unpadded]
__global__ void pca_finalize_kernel(const double* __restrict__ sums, int DP, int D, long long N,
                                    const double* __restrict__ shift, double* __restrict__ mean,
                                    double* __restrict__ cov) {
  const int tid = threadIdx.x;
  const double* cs = sums + DP * DP;
  if (tid < D) mean[tid] = shift[tid] + cs[tid] / (double)N;
  const double denom = N > 1 ? (double)(N - 1) : 1.0;
  for (int idx = tid; idx < D * D; idx += blockDim.x) {
    const int i = idx / D, j = idx - i * D;
    cov[idx] = (sums[i * DP + j] - cs[i] * cs[j] / (double)N) / denom;
  }
}

// Jacobi eigen-solver (parallel round-robin ordering) for one symmetric D x D matrix (D <= 64)
// in shared memory.
// evals: descending, clipped at 0; comps[k][:] = unit eigenvector k with its largest-magnitude
// entry made positive (first such entry on ties), as sklearn's svd_flip on V^T.
__global__ void __launch_bounds__(PCA_MAX_D)
pca_eigh_kernel(const double* __restrict__ cov, int D, double* __restrict__ evals, double* __restrict__ comps) {
  extern __shared__ double sm[];
  const int LD = D + 1;
  double* A = sm;
  double* V = sm + D * LD;
  __shared__ double red[PCA_MAX_D];
  __shared__ int order[PCA_MAX_D];
  __shared__ int converged;
  __shared__ int rot_p[PCA_MAX_D / 2], rot_q[PCA_MAX_D / 2];
  __shared__ double rot_c[PCA_MAX_D / 2], rot_s[PCA_MAX_D / 2];
  const int k = threadIdx.x;
  if (k < D)
    for (int j = 0; j < D; ++j) {
      // symmetrise: the two triangles of cov were summed in different orders
      A[k * LD + j] = 0.5 * (cov[k * D + j] + cov[j * D + k]);
      V[k * LD + j] = (j == k) ? 1.0 : 0.0;
    }
  __syncthreads();
  for (int sweep = 0; sweep < 40; ++sweep) {
    // off-diagonal and diagonal norms
    double off = 0.0, dg = 0.0;
    if (k < D) {
      for (int j = 0; j < D; ++j) {
        const double v = A[k * LD + j];
        if (j == k) dg = v * v; else off += v * v;
      }
    }
    red[k] = off;
    __syncthreads();
    if (k == 0) {
      double so = 0.0;
      for (int j = 0; j < D; ++j) so += red[j];
      red[0] = so;
    }
    __syncthreads();
    const double so = red[0];
    __syncthreads();
    red[k] = dg;
    __syncthreads();
    if (k == 0) {
      double sd = 0.0;
      for (int j = 0; j < D; ++j) sd += red[j];
      converged = (so <= 1e-30 * sd) || (so == 0.0);
    }
    __syncthreads();
    if (converged) break;
    // One sweep = n-1 rounds of a round-robin tournament over the columns: the D/2 pairs of a
    // round are disjoint, so their rotations are computed from the current matrix and applied
    // together -- columns (A <- A J, V <- V J; thread k owns row k), then rows (A <- J^T A;
    // thread k owns column k), then the rotated 2x2 blocks are set from the closed form.
    const int n = D + (D & 1);
    for (int r = 0; r < n - 1; ++r) {
      double t_rot = 0.0, app = 0.0, aqq = 0.0, apq = 0.0;
      int p = 0, q = 0;
      bool active = false;
      if (k < n / 2) {
        int u, v;
        if (k == 0) {
          u = r % (n - 1);
          v = n - 1;
        } else {
          u = (r + k) % (n - 1);
          v = (r - k + (n - 1)) % (n - 1);
        }
        p = u < v ? u : v;
        q = u < v ? v : u;
        double c = 1.0, sn = 0.0;
        if (q < D) {
          apq = A[p * LD + q];
          app = A[p * LD + p];
          aqq = A[q * LD + q];
          const double g100 = 100.0 * fabs(apq);
          const bool tiny = sweep > 3 && (fabs(app) + g100 == fabs(app)) && (fabs(aqq) + g100 == fabs(aqq));
          if (tiny) {
            active = true;            // zero the entry, no rotation (t = 0)
          } else if (apq != 0.0) {
            const double tau = (aqq - app) / (2.0 * apq);
            t_rot = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t_rot * t_rot);
            sn = t_rot * c;
            active = true;
          }
        }
        rot_p[k] = p;
        rot_q[k] = q < D ? q : -1;
        rot_c[k] = c;
        rot_s[k] = sn;
      }
      __syncthreads();
      if (k < D) {
        for (int i = 0; i < n / 2; ++i) {
          const int pi = rot_p[i], qi = rot_q[i];
          const double c = rot_c[i], sn = rot_s[i];
          if (qi < 0 || sn == 0.0) continue;
          const double a = A[k * LD + pi], b2 = A[k * LD + qi];
          A[k * LD + pi] = c * a - sn * b2;
          A[k * LD + qi] = sn * a + c * b2;
          const double va = V[k * LD + pi], vb = V[k * LD + qi];
          V[k * LD + pi] = c * va - sn * vb;
          V[k * LD + qi] = sn * va + c * vb;
        }
      }
      __syncthreads();
      if (k < D) {
        for (int i = 0; i < n / 2; ++i) {
          const int pi = rot_p[i], qi = rot_q[i];
          const double c = rot_c[i], sn = rot_s[i];
          if (qi < 0 || sn == 0.0) continue;
          const double a = A[pi * LD + k], b2 = A[qi * LD + k];
          A[pi * LD + k] = c * a - sn * b2;
          A[qi * LD + k] = sn * a + c * b2;
        }
      }
      __syncthreads();
      if (active) {
        A[p * LD + p] = app - t_rot * apq;
        A[q * LD + q] = aqq + t_rot * apq;
        A[p * LD + q] = 0.0;
        A[q * LD + p] = 0.0;
      }
      __syncthreads();
    }
  }
  // order eigenvalues (descending; stable in the index on ties)
  if (k == 0) {
    for (int j = 0; j < D; ++j) order[j] = j;
    for (int i = 0; i < D - 1; ++i) {
      int best = i;
      for (int j = i + 1; j < D; ++j)
        if (A[order[j] * LD + order[j]] > A[order[best] * LD + order[best]]) best = j;
      const int tmp = order[best];
      for (int j = best; j > i; --j) order[j] = order[j - 1];
      order[i] = tmp;
    }
  }
  __syncthreads();
  if (k < D) {
    const int col = order[k];
    const double lam = A[col * LD + col];
    evals[k] = lam > 0.0 ? lam : 0.0;
    double best = -1.0, sign = 1.0;
    for (int d = 0; d < D; ++d) {
      const double v = V[d * LD + col];
      if (fabs(v) > best) {
        best = fabs(v);
        sign = v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0);
      }
    }
    for (int d = 0; d < D; ++d) comps[k * D + d] = sign * V[d * LD + col];
  }
}

// out[n][k] = sum_d (x[n][d] - mean[d]) * comps[k][d], k < K.  A tile of 128 rows is staged in
// shared memory with coalesced loads (row stride D+1 doubles: conflict-free column walks); one
// thread per row, K accumulators in chunks of 8.
constexpr int PCA_TROWS = 128;
template <typename T>
__global__ void __launch_bounds__(PCA_TROWS)
pca_transform_kernel(const T* __restrict__ x, long long N, int D, const double* __restrict__ mean,
                     const double* __restrict__ comps, int K, double* __restrict__ out) {
  extern __shared__ double sm[];
  const int LD = D + 1;
  double* xs = sm;                      // [PCA_TROWS][LD]
  double* cs = sm + PCA_TROWS * LD;     // [K][D]
  __shared__ double smean[PCA_MAX_D];
  const int tid = threadIdx.x;
  if (tid < D) smean[tid] = mean[tid];
  for (int idx = tid; idx < K * D; idx += PCA_TROWS) cs[idx] = comps[idx];
  const long long n_tiles = (N + PCA_TROWS - 1) / PCA_TROWS;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long row0 = t * PCA_TROWS;
    __syncthreads();
    for (int idx = tid; idx < PCA_TROWS * D; idx += PCA_TROWS) {
      const int r = idx / D, c = idx - r * D;
      xs[r * LD + c] = (row0 + r < N) ? (double)x[(row0 + r) * D + c] - smean[c] : 0.0;
    }
    __syncthreads();
    const long long row = row0 + tid;
    if (row < N) {
      const double* xr = xs + tid * LD;
      for (int k0 = 0; k0 < K; k0 += 8) {
        double acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.0;
        for (int d = 0; d < D; ++d) {
          const double v = xr[d];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (k0 + q < K) acc[q] = fma(v, cs[(k0 + q) * D + d], acc[q]);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (k0 + q < K) out[row * K + k0 + q] = acc[q];
      }
    }
  }
}

static int padded_dim(int D) { return D <= 8 ? 8 : D <= 16 ? 16 : D <= 32 ? 32 : 64; }

static int gram_ctas(long long N) {
  const long long tiles = (N + PCA_ROWS - 1) / PCA_ROWS;
  return (int)(tiles < PCA_MAX_CTAS ? (tiles < 1 ? 1 : tiles) : PCA_MAX_CTAS);
}

template <typename T>
static void launch_gram(const T* x, long long N, int D, int DP, const double* shift, double* partial, int ctas,
                        cudaStream_t stream) {
  switch (DP) {
    case 8: pca_gram_kernel<T, 8><<<ctas, PCA_THREADS, 0, stream>>>(x, N, D, shift, partial); break;
    case 16: pca_gram_kernel<T, 16><<<ctas, PCA_THREADS, 0, stream>>>(x, N, D, shift, partial); break;
    case 32: pca_gram_kernel<T, 32><<<ctas, PCA_THREADS, 0, stream>>>(x, N, D, shift, partial); break;
    default: pca_gram_kernel<T, 64><<<ctas, PCA_THREADS, 0, stream>>>(x, N, D, shift, partial); break;
  }
}

}  // namespace ava

using namespace ava;

extern "C" long long ava_b200_pca_ws_bytes(int D) {
  if (D < 1 || D > PCA_MAX_D) return -1;
  const int DP = padded_dim(D);
  return (long long)sizeof(double) * ((long long)(PCA_MAX_CTAS + 1) * (DP * DP + DP) + PCA_MAX_D);
}

extern "C" int ava_b200_pca_fit(const void* x, int is_f32, long long N, int D, double* mean, double* cov,
                                double* evals, double* comps, void* ws, long long ws_bytes, void* stream_) {
  AVA_REQUIRE(D >= 1 && D <= PCA_MAX_D, "pca_fit: D=%d (1..%d supported)", D, PCA_MAX_D);
  AVA_REQUIRE(N >= 2, "pca_fit: needs at least 2 rows, got %lld", N);
  AVA_REQUIRE(ws != nullptr && ws_bytes >= ava_b200_pca_ws_bytes(D), "pca_fit: workspace of %lld bytes needed, got %lld",
              ava_b200_pca_ws_bytes(D), ws_bytes);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int DP = padded_dim(D);
  double* shift = (double*)ws;
  const int stride = DP * DP + DP;
  double* sums = shift + PCA_MAX_D;
  double* partial = sums + stride;
  const int ctas = gram_ctas(N);
  if (is_f32) {
    pca_shift_kernel<float><<<1, PCA_MAX_D, 0, stream>>>((const float*)x, N, D, shift);
    if (check_launch("pca_shift")) return 1;
    launch_gram<float>((const float*)x, N, D, DP, shift, partial, ctas, stream);
  } else {
    pca_shift_kernel<double><<<1, PCA_MAX_D, 0, stream>>>((const double*)x, N, D, shift);
    if (check_launch("pca_shift")) return 1;
    launch_gram<double>((const double*)x, N, D, DP, shift, partial, ctas, stream);
  }
  if (check_launch("pca_gram")) return 1;
  pca_reduce_kernel<<<(stride + 127) / 128, 128, 0, stream>>>(partial, ctas, stride, sums);
  if (check_launch("pca_reduce")) return 1;
  pca_finalize_kernel<<<1, 1024, 0, stream>>>(sums, DP, D, N, shift, mean, cov);
  if (check_launch("pca_finalize")) return 1;
  const size_t smem = sizeof(double) * 2 * (size_t)D * (D + 1);
  if (cudaFuncSetAttribute(pca_eigh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(double) * 2 * PCA_MAX_D * (PCA_MAX_D + 1))) != cudaSuccess) {
    set_error("pca_fit: cudaFuncSetAttribute failed");
    return 1;
  }
  pca_eigh_kernel<<<1, PCA_MAX_D, smem, stream>>>(cov, D, evals, comps);
  return check_launch("pca_eigh");
}

extern "C" int ava_b200_pca_transform(const void* x, int is_f32, long long N, int D, const double* mean,
                                      const double* comps, int K, double* out, void* stream_) {
  AVA_REQUIRE(D >= 1 && D <= PCA_MAX_D, "pca_transform: D=%d (1..%d supported)", D, PCA_MAX_D);
  AVA_REQUIRE(K >= 1 && K <= D, "pca_transform: K=%d with D=%d", K, D);
  AVA_REQUIRE(N >= 0, "pca_transform: N=%lld", N);
  if (N == 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t smem = sizeof(double) * ((size_t)PCA_TROWS * (D + 1) + (size_t)K * D);
  const int max_smem = (int)(sizeof(double) * (PCA_TROWS * (PCA_MAX_D + 1) + PCA_MAX_D * PCA_MAX_D));
  if (cudaFuncSetAttribute(pca_transform_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) !=
          cudaSuccess ||
      cudaFuncSetAttribute(pca_transform_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) !=
          cudaSuccess) {
    set_error("pca_transform: cudaFuncSetAttribute failed");
    return 1;
  }
  const long long tiles = (N + PCA_TROWS - 1) / PCA_TROWS;
  const int ctas = (int)(tiles < 4LL * kNumSMs ? tiles : 4LL * kNumSMs);
  if (is_f32)
    pca_transform_kernel<float><<<ctas, PCA_TROWS, smem, stream>>>((const float*)x, N, D, mean, comps, K, out);
  else
    pca_transform_kernel<double><<<ctas, PCA_TROWS, smem, stream>>>((const double*)x, N, D, mean, comps, K, out);
  return check_launch("pca_transform");
}
