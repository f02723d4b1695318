// Shared helpers for the ava_b200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ava_b200.h"

namespace ava {

constexpr float kBnEps = 1e-5f;
constexpr int kStatsStride = AVA_STATS_STRIDE;  // doubles per BN layer
constexpr int kNumSMs = 148;

void set_error(const char* fmt, ...);
// cuTensorMapEncodeTiled, fetched through the runtime (no libcuda link dependency)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn get_tensor_map_encode_fn();
void count_launch(int n = 1);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  count_launch();
  return 0;
}

#define AVA_REQUIRE(cond, ...)   \
  do {                           \
    if (!(cond)) {               \
      ava::set_error(__VA_ARGS__); \
      return 1;                  \
    }                            \
  } while (0)

// Programmatic dependent launch.  A kernel launched through launch_pdl() may be scheduled while the
// kernel before it in the stream is still draining; it must call pdl_wait() before it reads or
// writes anything an earlier kernel of the stream produced (parameters are the exception the
// prologues use: within a step only the optimizer kernels write them, and those never trigger
// their dependents early).  pdl_launch_dependents() -- always AFTER pdl_wait(), so that an early
// started kernel only ever overlaps its immediate predecessor -- lets the next kernel's CTAs
// take the SM resources this kernel's CTAs free as they finish, instead of waiting for the whole
// grid to retire and a launch gap on top.  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Measured (profiles/r02_summary.md): inside the captured CUDA graph of the step the attribute
// buys nothing (7.11 vs 7.06 ms at batch 1024, 1.337 vs 1.326 ms at batch 64 -- graph launches
// already leave no gap worth hiding), so it is OFF unless AVA_B200_PDL=1.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The Adam update, shared by adam_kernel (adam.cu) and the fused data-parallel kernel (dp.cu) with
// every rounding spelled out, so that the two produce the same bits from the same summed gradient
// whatever the compiler contracts around them: a 1-ulp difference in the first steps is amplified
// by ReLU kinks and Adam's scale invariance into lr-sized differences of ~1 % of the parameters
// within a few steps (seen when the two kernels were written independently).
struct AdamCoef {
  float b1, b2, omb1, omb2, eps, step_size, inv_sqrt_bc2;
};
// Scalars exactly as torch.optim.Adam forms them (Python doubles, rounded to fp32 once); t = the
// step number of THIS update (torch increments `step` before using it)
__device__ __forceinline__ AdamCoef adam_coef(double lr_d, double b1_d, double b2_d, double eps_d, double t) {
  const double bc1 = 1.0 - pow(b1_d, t);
  const double bc2 = 1.0 - pow(b2_d, t);
  AdamCoef k;
  k.step_size = (float)(lr_d / bc1);
  k.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  k.b1 = (float)b1_d;
  k.b2 = (float)b2_d;
  k.eps = (float)eps_d;
  k.omb1 = (float)(1.0 - b1_d);
  k.omb2 = (float)(1.0 - b2_d);
  return k;
}
__device__ __forceinline__ void adam_update(const AdamCoef& k, float gg, float& p, float& m, float& v) {
  m = __fmaf_rn(k.b1, m, __fmul_rn(k.omb1, gg));
  v = __fmaf_rn(k.b2, v, __fmul_rn(__fmul_rn(k.omb2, gg), gg));
  const float denom = __fmaf_rn(__fsqrt_rn(v), k.inv_sqrt_bc2, k.eps);
  p = __fmaf_rn(-k.step_size, __fdiv_rn(m, denom), p);
}

// BatchNorm coefficients for one channel from accumulated sums (train) or running
// buffers (eval):  bn(x) = x*scale + shift.
struct BnCoef {
  float mean, invstd, scale, shift;
};
__device__ __forceinline__ BnCoef bn_coef(const double* stats, int c, double count, const float* gamma,
                                          const float* beta, const float* rmean, const float* rvar, bool train) {
  double mean, var;
  if (train) {
    mean = stats[c] / count;
    var = stats[32 + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
  } else {
    mean = (double)rmean[c];
    var = (double)rvar[c];
  }
  double invstd = rsqrt(var + (double)kBnEps);
  BnCoef k;
  k.mean = (float)mean;
  k.invstd = (float)invstd;
  double g = gamma ? (double)gamma[c] : 1.0;
  double b = beta ? (double)beta[c] : 0.0;
  k.scale = (float)(g * invstd);
  k.shift = (float)(b - mean * g * invstd);
  return k;
}

// Backward of the BatchNorm that consumes y (train mode), SURVEY.md appendix A:
//   dx = gamma*invstd*(dy - dbeta/N - xhat*dgamma/N)
//      = p*(dy - c1) + q*(y - mean),   p = gamma*invstd, c1 = dbeta/N,
//                                      q = -gamma*invstd^3 * S/N, S = sum dy*(y-mean)
// with dstats[c] = sum dy (= dbeta) and dstats[32+c] = S (= dgamma/invstd).
// Evaluated in fp64: the two terms cancel heavily (BN backward projects out the mean and
// the xhat direction), and a coefficient rounded to fp32 would add a *coherent* error to
// every element of the channel, which bias / BN-affine gradients then sum up.
struct DzCoef {
  double p, q, mean, c1;
};
__device__ __forceinline__ DzCoef dz_coef(const float* gamma, const double* stats, const double* dstats, int c,
                                          double count) {
  DzCoef k;
  if (gamma == nullptr) {
    k.p = 1.0;
    k.q = 0.0;
    k.mean = 0.0;
    k.c1 = 0.0;
    return k;
  }
  double mean = stats[c] / count;
  double var = stats[32 + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  double invstd = rsqrt(var + (double)kBnEps);
  double g = (double)gamma[c];
  k.p = g * invstd;
  k.q = -g * invstd * invstd * invstd * dstats[32 + c] / count;
  k.mean = mean;
  k.c1 = dstats[c] / count;
  return k;
}
__device__ __forceinline__ float dz_apply(const DzCoef& k, float g, float y) {
  return (float)(k.p * ((double)g - k.c1) + k.q * ((double)y - k.mean));
}

}  // namespace ava
