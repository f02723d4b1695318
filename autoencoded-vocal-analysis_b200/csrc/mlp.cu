// The VAE's small dense layers as two row-wise fused kernels.
//
// Between the two 8192 x 1024 layers (fc1, fc8: tcgen05 GEMMs) the network is a chain of small
// per-sample layers (ava/models/vae.py:226-232 encoder tail, :298-311 reparameterised sample,
// :258-260 decoder head):
//     h1[1024] -fc2-> h2[256] -fc31|32|33-> h3[3x64] -fc41|42|43-> heads = (mu | u | log d)[3Z]
//     z = mu + u eps_W + sqrt(d) eps_D                                   (LowRankMultivariateNormal.rsample)
//     z[Z] -fc5-> t5[64] -fc6-> t6[256] -fc7-> t7[1024]
// 0.6 MFLOP per sample: as seven GEMM launches (+ split-K reductions) they cost ~100 us of launch
// latency and tiny-grid tails per direction whatever the batch size.  Every row is independent, so
// ONE kernel walks the whole chain for a tile of R rows with all intermediates in shared memory
// (the weights, 2.6 MB in all, stream from L2), and ONE kernel walks it backwards
// (dt7 -> ... -> gz -> analytic latent gradient -> ... -> dh1).  Arithmetic: fp32 FMA, fp32
// accumulation (the per-kernel parity bar of the dense layers, rtol 1e-4 vs float64, holds).
// The weight gradients (reductions over the batch) stay GEMMs: ava_b200_linear_bwd_weight.
#include "common.cuh"

namespace ava {

struct MlpParams {
  int B, Z, stages;
  // parameters (torch Linear layout [out][in]); w3/b3 = fc31|fc32|fc33 stacked, w4/b4 = fc41|fc42|fc43
  const float *w2, *b2, *w3, *b3, *w4, *b4, *w5, *b5, *w6, *b6, *w7, *b7;
  const float *eps_w, *eps_d;
  // forward: in h1 (stage 1) / z (stage 4 without stage 2); out everything else
  // backward: in dt7 + the saved activations; out dt6, dt5, gz, gheads, dh3, dh2, dh1
  float *h1, *h2, *h3, *heads, *z, *d, *t5, *t6, *t7;
  float *dt7, *dt6, *dt5, *gz, *gheads, *dh3, *dh2, *dh1;
  double* acc;
};

// y[r][n] = act(b[n] + sum_k x[r][xoff(n) + k] W[n][k]) for the tile's R rows: one warp per output
// neuron, lanes along k (coalesced 16-byte weight loads, each reused for the R rows).
// Grouped layers (fc4x): neuron n reads the x slice of its group, xoff = (n / ngrp) * xstride.
template <int R, bool RELU>
__device__ __forceinline__ void dense_rows(const float* __restrict__ W, const float* __restrict__ bias, int N, int K,
                                           const float* s_x, int ldx, int ngrp, int xstride, float* s_y, int ldy) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = warp; n < N; n += 8) {
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    const float* wr = W + (size_t)n * K;
    const float* xb = s_x + (n / ngrp) * xstride;
#pragma unroll 4
    for (int k = lane * 4; k < K; k += 128) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(xb + r * ldx + k);
        acc[r] = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, acc[r]))));
      }
    }
    float mine = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float t = warp_sum(acc[r]);
      if (lane == r) mine = t;
    }
    if (lane < R) {
      float v = mine + bias[n];
      if (RELU) v = fmaxf(v, 0.f);
      s_y[lane * ldy + n] = v;
    }
  }
}

// dx[r][k] = sum_n g[r][n] W[n][k] (g = the upstream gradient, already masked): threads along k
// (coalesced 16-byte weight loads), the n range split over 256 / (K/4) thread groups whose partial
// sums meet in shared memory.  Grouped layers (fc4x backward): output k of group kg = k / kgrp only
// sums the neurons of that group, W row n holds kgrp inputs.
template <int R>
__device__ __forceinline__ void dense_rows_bwd(const float* __restrict__ W, int N, int K, int kgrp, int ngrp,
                                               const float* s_g, int ldg, float* s_dx, int lddx, float* s_part) {
  const int KQ = K >> 2;
  int G = 256 / KQ;
  if (G < 1) G = 1;
  if (G > 32) G = 32;
  const int kq = threadIdx.x % KQ, ng = threadIdx.x / KQ;
  const bool active = ng < G;
  float acc[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
  if (active) {
    const int k = 4 * kq;
    const int kg = k / kgrp;                     // group of this output (0 for plain layers)
    const int n0 = kg * ngrp, n1 = (kgrp == K) ? N : n0 + ngrp;
    const int kin = k - kg * kgrp;               // column inside the group's weight rows
    const int wld = kgrp;
#pragma unroll 4
    for (int n = n0 + ng; n < n1; n += G) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * wld + kin));
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float g = s_g[r * ldg + n];
        acc[r][0] = fmaf(g, w.x, acc[r][0]);
        acc[r][1] = fmaf(g, w.y, acc[r][1]);
        acc[r][2] = fmaf(g, w.z, acc[r][2]);
        acc[r][3] = fmaf(g, w.w, acc[r][3]);
      }
    }
  }
  if (G == 1) {
    if (active) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        *reinterpret_cast<float4*>(s_dx + r * lddx + 4 * kq) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
    return;
  }
  // partial sums [G][R][K] -> fixed-order sum over G
  if (active) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      *reinterpret_cast<float4*>(s_part + ((size_t)ng * R + r) * K + 4 * kq) =
          make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < R * K; idx += 256) {
    const int r = idx / K, k = idx - r * K;
    float t = 0.f;
    for (int g = 0; g < G; ++g) t += s_part[((size_t)g * R + r) * K + k];
    s_dx[r * lddx + k] = t;
  }
}

// shared -> global, rows of the tile that exist (coalesced along the row)
template <int R>
__device__ __forceinline__ void store_rows(const float* s, int lds, float* g, int ldg, int N, int row0, int B) {
  if (g == nullptr) return;
  for (int idx = threadIdx.x; idx < R * N; idx += 256) {
    const int r = idx / N, n = idx - r * N;
    if (row0 + r < B) g[(size_t)(row0 + r) * ldg + n] = s[r * lds + n];
  }
}
template <int R>
__device__ __forceinline__ void load_rows(float* s, int lds, const float* g, int ldg, int N, int row0, int B,
                                          const float* mask = nullptr) {
  for (int idx = threadIdx.x; idx < R * N; idx += 256) {
    const int r = idx / N, n = idx - r * N;
    float v = 0.f;
    if (row0 + r < B) {
      v = g[(size_t)(row0 + r) * ldg + n];
      if (mask != nullptr && !(mask[(size_t)(row0 + r) * ldg + n] > 0.f)) v = 0.f;
    }
    s[r * lds + n] = v;
  }
}

constexpr int kMlpMaxZ = 64;

// shared-memory plan (floats), forward: x[R][1024] | h2[R][256] | h3[R][192] | heads[R][192] |
// z[R][64] | t5[R][64] | t6[R][256] ; t7 reuses x
template <int R>
struct MlpSmem {
  static constexpr int X = 0, H2 = X + R * 1024, H3 = H2 + R * 256, HD = H3 + R * 192, ZZ = HD + R * 192,
                       T5 = ZZ + R * 64, T6 = T5 + R * 64, END_FWD = T6 + R * 256;
  // backward: gA[R][1024] | gB[R][1024] (ping-pong gradients) | part[<= 32 KB]
  static constexpr int GA = 0, GB = GA + R * 1024, PART = GB + R * 1024, END_BWD = PART + 8192;
  static constexpr int FLOATS = END_FWD > END_BWD ? END_FWD : END_BWD;
};

template <int R>
__global__ void __launch_bounds__(256) mlp_fwd_kernel(const MlpParams P) {
  using S = MlpSmem<R>;
  extern __shared__ __align__(16) float sm[];
  __shared__ double s_z2[8], s_h[8];
  const int row0 = blockIdx.x * R;
  const int Z = P.Z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_x = sm + S::X;
  float* s_h2 = sm + S::H2;
  float* s_h3 = sm + S::H3;
  float* s_hd = sm + S::HD;
  float* s_z = sm + S::ZZ;
  float* s_t5 = sm + S::T5;
  float* s_t6 = sm + S::T6;
  if (P.stages & 1) {
    load_rows<R>(s_x, 1024, P.h1, 1024, 1024, row0, P.B);
    __syncthreads();
    dense_rows<R, true>(P.w2, P.b2, 256, 1024, s_x, 1024, 256, 0, s_h2, 256);
    __syncthreads();
    store_rows<R>(s_h2, 256, P.h2, 256, 256, row0, P.B);
    dense_rows<R, true>(P.w3, P.b3, 192, 256, s_h2, 256, 192, 0, s_h3, 192);
    __syncthreads();
    store_rows<R>(s_h3, 192, P.h3, 192, 192, row0, P.B);
    // three [Z, 64] heads on the three 64-wide slices of h3
    dense_rows<R, false>(P.w4, P.b4, 3 * Z, 64, s_h3, 192, Z, 64, s_hd, 192);
    __syncthreads();
    store_rows<R>(s_hd, 192, P.heads, 3 * Z, 3 * Z, row0, P.B);
  }
  if (P.stages & 2) {
    // reparameterised sample + the two latent terms of the loss, one warp per row
    // (LowRankMultivariateNormal.rsample / entropy, ava/models/vae.py:298-316; as latent_fwd_kernel)
    if (!(P.stages & 1)) {
      load_rows<R>(s_hd, 192, P.heads, 3 * Z, 3 * Z, row0, P.B);
      __syncthreads();
    }
    double z2 = 0.0, hb = 0.0;
    if (warp < R && row0 + warp < P.B) {
      const int b = row0 + warp;
      const float* row = s_hd + warp * 192;
      const float ew = P.eps_w[b];
      float sz2 = 0.f, s = 0.f, slogd = 0.f;
      for (int i = lane; i < Z; i += 32) {
        const float mu = row[i], u = row[Z + i], ld = row[2 * Z + i];
        const float d = expf(ld);
        const float zz = mu + u * ew + sqrtf(d) * P.eps_d[(size_t)b * Z + i];
        s_z[warp * 64 + i] = zz;
        P.z[(size_t)b * Z + i] = zz;
        if (P.d) P.d[(size_t)b * Z + i] = d;
        sz2 = fmaf(zz, zz, sz2);
        s += u * u / d;
        slogd += ld;
      }
      sz2 = warp_sum(sz2);
      s = warp_sum(s);
      slogd = warp_sum(slogd);
      z2 = sz2;
      hb = 0.5 * ((double)Z * (1.0 + 1.8378770664093453) + (double)log1pf(s) + (double)slogd);
    } else if (warp < R) {
      for (int i = lane; i < Z; i += 32) s_z[warp * 64 + i] = 0.f;
    }
    if (lane == 0) {
      s_z2[warp] = z2;
      s_h[warp] = hb;
    }
    __syncthreads();
    if (threadIdx.x == 0 && P.acc != nullptr) {
      double a = 0, h = 0;
      for (int w = 0; w < 8; ++w) {
        a += s_z2[w];
        h += s_h[w];
      }
      atomicAdd(&P.acc[0], a);
      atomicAdd(&P.acc[2], h);
    }
  } else if (P.stages & 4) {
    load_rows<R>(s_z, 64, P.z, Z, Z, row0, P.B);
    __syncthreads();
  }
  if (P.stages & 4) {
    dense_rows<R, true>(P.w5, P.b5, 64, Z, s_z, 64, 64, 0, s_t5, 64);
    __syncthreads();
    store_rows<R>(s_t5, 64, P.t5, 64, 64, row0, P.B);
    dense_rows<R, true>(P.w6, P.b6, 256, 64, s_t5, 64, 256, 0, s_t6, 256);
    __syncthreads();
    store_rows<R>(s_t6, 256, P.t6, 256, 256, row0, P.B);
    float* s_t7 = s_x;      // (the h1 tile is no longer needed)
    dense_rows<R, true>(P.w7, P.b7, 1024, 256, s_t6, 256, 1024, 0, s_t7, 1024);
    __syncthreads();
    store_rows<R>(s_t7, 1024, P.t7, 1024, 1024, row0, P.B);
  }
}

// Backward-data chain.  dt_l / dh_l are the gradients w.r.t. the layers' POST-activation outputs,
// stored unmasked exactly as the per-layer path stores them (the weight-gradient and bias kernels
// apply the ReLU masks themselves); the latent gradient is latent_bwd_kernel's arithmetic.
template <int R>
__global__ void __launch_bounds__(256) mlp_bwd_kernel(const MlpParams P) {
  using S = MlpSmem<R>;
  extern __shared__ __align__(16) float sm[];
  const int row0 = blockIdx.x * R;
  const int Z = P.Z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* gA = sm + S::GA;
  float* gB = sm + S::GB;
  float* part = sm + S::PART;
  // g7 = dt7 (.) [t7 > 0]
  load_rows<R>(gA, 1024, P.dt7, 1024, 1024, row0, P.B, P.t7);
  __syncthreads();
  dense_rows_bwd<R>(P.w7, 1024, 256, 256, 1024, gA, 1024, gB, 256, part);      // dt6 [256]
  __syncthreads();
  store_rows<R>(gB, 256, P.dt6, 256, 256, row0, P.B);
  for (int idx = threadIdx.x; idx < R * 256; idx += 256) {                     // (.) [t6 > 0]
    const int r = idx >> 8, n = idx & 255;
    if (row0 + r < P.B && !(P.t6[(size_t)(row0 + r) * 256 + n] > 0.f)) gB[r * 256 + n] = 0.f;
  }
  __syncthreads();
  dense_rows_bwd<R>(P.w6, 256, 64, 64, 256, gB, 256, gA, 64, part);            // dt5 [64]
  __syncthreads();
  store_rows<R>(gA, 64, P.dt5, 64, 64, row0, P.B);
  for (int idx = threadIdx.x; idx < R * 64; idx += 256) {
    const int r = idx >> 6, n = idx & 63;
    if (row0 + r < P.B && !(P.t5[(size_t)(row0 + r) * 64 + n] > 0.f)) gA[r * 64 + n] = 0.f;
  }
  __syncthreads();
  dense_rows_bwd<R>(P.w5, 64, Z, Z, 64, gA, 64, gB, 64, part);                 // gz [Z]
  __syncthreads();
  store_rows<R>(gB, 64, P.gz, Z, Z, row0, P.B);
  // latent: gheads from gz (decoder path) + prior + entropy terms, one warp per row
  if (warp < R) {
    float* grow = gA + warp * 192;
    if (row0 + warp < P.B) {
      const int b = row0 + warp;
      const float* row = P.heads + (size_t)b * 3 * Z;
      const float ew = P.eps_w[b];
      float s = 0.f;
      for (int i = lane; i < Z; i += 32) {
        const float u = row[Z + i];
        s += u * u * expf(-row[2 * Z + i]);
      }
      s = warp_sum(s);
      const float inv1s = 1.f / (1.f + s);
      for (int i = lane; i < Z; i += 32) {
        const float u = row[Z + i], ld = row[2 * Z + i];
        const float d = expf(ld);
        const float g = gB[warp * 64 + i] + P.z[(size_t)b * Z + i];
        const float uod = u / d;
        grow[i] = g;
        grow[Z + i] = g * ew - uod * inv1s;
        grow[2 * Z + i] = 0.5f * g * P.eps_d[(size_t)b * Z + i] * sqrtf(d) - 0.5f * (1.f - u * uod * inv1s);
      }
    } else {
      for (int i = lane; i < 3 * Z; i += 32) grow[i] = 0.f;
    }
  }
  __syncthreads();
  store_rows<R>(gA, 192, P.gheads, 3 * Z, 3 * Z, row0, P.B);
  // dh3[g*64 + k] = sum_j gheads[g*Z + j] W4[g][j][k]   (no activation on the heads)
  dense_rows_bwd<R>(P.w4, 3 * Z, 192, 64, Z, gA, 192, gB, 192, part);
  __syncthreads();
  store_rows<R>(gB, 192, P.dh3, 192, 192, row0, P.B);
  for (int idx = threadIdx.x; idx < R * 192; idx += 256) {
    const int r = idx / 192, n = idx - r * 192;
    if (row0 + r < P.B && !(P.h3[(size_t)(row0 + r) * 192 + n] > 0.f)) gB[r * 192 + n] = 0.f;
  }
  __syncthreads();
  dense_rows_bwd<R>(P.w3, 192, 256, 256, 192, gB, 192, gA, 256, part);         // dh2 [256]
  __syncthreads();
  store_rows<R>(gA, 256, P.dh2, 256, 256, row0, P.B);
  for (int idx = threadIdx.x; idx < R * 256; idx += 256) {
    const int r = idx >> 8, n = idx & 255;
    if (row0 + r < P.B && !(P.h2[(size_t)(row0 + r) * 256 + n] > 0.f)) gA[r * 256 + n] = 0.f;
  }
  __syncthreads();
  dense_rows_bwd<R>(P.w2, 256, 1024, 1024, 256, gA, 256, gB, 1024, part);      // dh1 [1024]
  __syncthreads();
  store_rows<R>(gB, 1024, P.dh1, 1024, 1024, row0, P.B);
}

template <int R>
static int launch_mlp(const MlpParams& P, bool backward, cudaStream_t stream) {
  const size_t smem = (size_t)MlpSmem<R>::FLOATS * sizeof(float);
  auto kf = mlp_fwd_kernel<R>;
  auto kb = mlp_bwd_kernel<R>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      set_error("mlp: cannot reserve %zu bytes of shared memory", smem);
      return 1;
    }
    configured = true;
  }
  const int grid = (P.B + R - 1) / R;
  if (backward)
    kb<<<grid, 256, smem, stream>>>(P);
  else
    kf<<<grid, 256, smem, stream>>>(P);
  return check_launch(backward ? "mlp_bwd" : "mlp_fwd");
}

static int run_mlp(const ava_b200_mlp_params* h, bool backward, void* stream_) {
  AVA_REQUIRE(h != nullptr, "mlp: null parameter block");
  AVA_REQUIRE(h->Z >= 4 && h->Z <= kMlpMaxZ && h->Z % 4 == 0, "mlp: z_dim %d not supported (multiple of 4, <= %d)", h->Z,
              kMlpMaxZ);
  if (h->B <= 0) return 0;
  MlpParams P = {};
  P.B = h->B;
  P.Z = h->Z;
  P.stages = h->stages;
  P.w2 = h->w2; P.b2 = h->b2; P.w3 = h->w3; P.b3 = h->b3; P.w4 = h->w4; P.b4 = h->b4;
  P.w5 = h->w5; P.b5 = h->b5; P.w6 = h->w6; P.b6 = h->b6; P.w7 = h->w7; P.b7 = h->b7;
  P.eps_w = h->eps_w; P.eps_d = h->eps_d;
  P.h1 = h->h1; P.h2 = h->h2; P.h3 = h->h3; P.heads = h->heads; P.z = h->z; P.d = h->d;
  P.t5 = h->t5; P.t6 = h->t6; P.t7 = h->t7;
  P.dt7 = h->dt7; P.dt6 = h->dt6; P.dt5 = h->dt5; P.gz = h->gz; P.gheads = h->gheads;
  P.dh3 = h->dh3; P.dh2 = h->dh2; P.dh1 = h->dh1;
  P.acc = h->acc;
  if (!backward) {
    AVA_REQUIRE(P.stages >= 1 && P.stages <= 7, "mlp_fwd: stages %d", P.stages);
    if (P.stages & 1) AVA_REQUIRE(P.h1 && P.h2 && P.h3 && P.heads && P.w2 && P.w3 && P.w4, "mlp_fwd: encoder tail pointers");
    if (P.stages & 2) AVA_REQUIRE(P.heads && P.z && P.eps_w && P.eps_d, "mlp_fwd: latent pointers");
    if (P.stages & 4) AVA_REQUIRE(P.z && P.t5 && P.t6 && P.t7 && P.w5 && P.w6 && P.w7, "mlp_fwd: decoder head pointers");
  } else {
    AVA_REQUIRE(P.dt7 && P.dt6 && P.dt5 && P.gz && P.gheads && P.dh3 && P.dh2 && P.dh1 && P.t7 && P.t6 && P.t5 && P.z &&
                    P.heads && P.h3 && P.h2 && P.eps_w && P.eps_d,
                "mlp_bwd: missing pointer");
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  // rows per CTA: enough CTAs to cover the SMs, as many rows per weight pass as that allows
  if (P.B >= 8 * 120) return launch_mlp<8>(P, backward, stream);
  if (P.B >= 4 * 120) return launch_mlp<4>(P, backward, stream);
  if (P.B >= 2 * 120) return launch_mlp<2>(P, backward, stream);
  return launch_mlp<1>(P, backward, stream);
}

}  // namespace ava

extern "C" int ava_b200_mlp_fwd(const ava_b200_mlp_params* h_params, void* stream) {
  return ava::run_mlp(h_params, false, stream);
}
extern "C" int ava_b200_mlp_bwd(const ava_b200_mlp_params* h_params, void* stream) {
  return ava::run_mlp(h_params, true, stream);
}
