// tcgen05/TMEM GEMM path for the two 8192x1024 dense layers (placeholder: reports
// "unsupported" until the tensor-core kernel lands; precision=0 is always available).
#include "common.cuh"

namespace ava {

int tc_gemm_supported(int M, int N, int K) {
  (void)M; (void)N; (void)K;
  return 0;
}

int tc_linear_fwd(const float*, int, const float*, const float*, float*, int, int, int, int, int, void*,
                  long long, cudaStream_t) {
  set_error("tensor-core linear path not built");
  return 1;
}

}  // namespace ava
