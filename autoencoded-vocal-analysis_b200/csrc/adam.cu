// Fused Adam over the flat parameter buffer (torch.optim.Adam defaults;
// ava/models/vae.py:119,353).  One elementwise pass: reads p,g,m,v; writes p,m,v.
#include "common.cuh"

namespace ava {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n4, long long n, const float* step_count, double lr_d, double b1_d, double b2_d,
            double eps_d, float gscale, const double* __restrict__ hyper) {
  if (hyper != nullptr) {   // hyper-parameters read from device memory (CUDA-graph replay safe)
    lr_d = hyper[0];
    b1_d = hyper[1];
    b2_d = hyper[2];
    eps_d = hyper[3];
  }
  const AdamCoef k = adam_coef(lr_d, b1_d, b2_d, eps_d, (double)step_count[0] + 1.0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_update(k, __fmul_rn(gv.x, gscale), pv.x, mv.x, vv.x);
    adam_update(k, __fmul_rn(gv.y, gscale), pv.y, mv.y, vv.y);
    adam_update(k, __fmul_rn(gv.z, gscale), pv.z, mv.z, vv.z);
    adam_update(k, __fmul_rn(gv.w, gscale), pv.w, mv.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // scalar tail
  if (blockIdx.x == 0) {
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      float pp = p[i], mm = m[i], vv = v[i];
      adam_update(k, __fmul_rn(g[i], gscale), pp, mm, vv);
      p[i] = pp;
      m[i] = mm;
      v[i] = vv;
    }
  }
}

__global__ void adam_bump_step_kernel(float* step_count) { step_count[0] += 1.f; }

}  // namespace ava

static int adam_launch(float* p, const float* g, float* m, float* v, long long n, float* step_count, double lr,
                       double beta1, double beta2, double eps, float grad_scale, const double* hyper,
                       void* stream_) {
  using namespace ava;
  if (n <= 0) return 0;
  AVA_REQUIRE(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                  ((uintptr_t)v % 16 == 0),
              "adam_step: buffers must be 16-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  long long n4 = n / 4;
  long long want = (n4 + 255) / 256;
  int grid = (int)(want < 1 ? 1 : (want > 16 * kNumSMs ? 16 * kNumSMs : want));
  adam_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, n4, n, step_count, lr, beta1, beta2, eps, grad_scale, hyper);
  if (check_launch("adam")) return 1;
  adam_bump_step_kernel<<<1, 1, 0, stream>>>(step_count);
  return check_launch("adam_bump_step");
}

extern "C" int ava_b200_adam_step(float* p, const float* g, float* m, float* v, long long n, float* step_count,
                                  double lr, double beta1, double beta2, double eps, float grad_scale,
                                  void* stream_) {
  return adam_launch(p, g, m, v, n, step_count, lr, beta1, beta2, eps, grad_scale, nullptr, stream_);
}

extern "C" int ava_b200_adam_step_dev(float* p, const float* g, float* m, float* v, long long n,
                                      float* step_count, const double* hyper, float grad_scale, void* stream_) {
  AVA_REQUIRE(hyper != nullptr, "adam_step_dev: hyper-parameter buffer required");
  return adam_launch(p, g, m, v, n, step_count, 0.0, 0.0, 0.0, 0.0, grad_scale, hyper, stream_);
}
