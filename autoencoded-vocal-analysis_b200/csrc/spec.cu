// Batched GPU `get_spec`: STFT -> log magnitude -> bilinear resampling onto the
// 128x128 (freq x time) grid -> normalise/clip.  One CTA per window, fp64 throughout
// (B200 runs fp64 at half the fp32 rate; the path is bound by the 64 KiB/window output
// write, not by arithmetic), so parity with the reference's float64 scipy path is ~1e-12.
//
// Reference: ava/preprocessing/utils.py:59-104 and scipy.signal.stft (_spectral_helper):
// zero boundary extension by nperseg/2, zero padding to a hop multiple, no detrend,
// window multiply, one-sided FFT, scale 1/sum(window).
//
// Frames are streamed: after frame k's log-spectrum is in the 2-deep ring buffer, every
// target time whose lower frame is k-1 is interpolated and written, so shared memory
// holds two spectra, not the whole spectrogram.  Frames no target touches are skipped.
#include "common.cuh"

namespace ava {

struct SpecParams {
  const void* audio;
  int is_f32;
  const long long* seg_start;
  const int* seg_len;
  int nperseg, hop, log2n, remove_dc;
  const double* window;
  double scale;
  const int* t_idx;
  const double* t_frac;
  int n_t;
  const int* f_idx;
  const double* f_frac;
  int n_f;
  int max_frames;
  double spec_min, inv_range;
  float* out;
  double* out64;
};

__device__ __forceinline__ double load_sample(const SpecParams& P, long long i) {
  if (P.is_f32) return (double)reinterpret_cast<const float*>(P.audio)[i];
  return (double)reinterpret_cast<const short*>(P.audio)[i];
}

__global__ void __launch_bounds__(256) get_spec_kernel(const SpecParams P) {
  extern __shared__ __align__(16) double sm[];
  const int N = P.nperseg, NB = N / 2 + 1;
  double2* s_x = reinterpret_cast<double2*>(sm);        // [N] FFT buffer
  double2* s_tw = s_x + N;                               // [N/2] twiddles
  double* s_log = reinterpret_cast<double*>(s_tw + N / 2);  // [2][NB] ring of log spectra
  double* s_red = s_log + 2 * NB;                        // [8]
  int* s_list = reinterpret_cast<int*>(s_red + 8);       // [n_t]
  int* s_cnt = s_list + P.n_t;                           // [1]
  unsigned char* s_need = reinterpret_cast<unsigned char*>(s_cnt + 1);  // [max_frames+1]

  const int tid = threadIdx.x;
  const int w = blockIdx.x;
  const int len = P.seg_len[w];
  const size_t out_base = (size_t)w * P.n_f * P.n_t;
  const int n_out = P.n_f * P.n_t;

  // every output starts at 0 (fill value / too-short segment / out-of-range targets)
  for (int i = tid; i < n_out; i += 256) {
    if (P.out) P.out[out_base + i] = 0.f;
    if (P.out64) P.out64[out_base + i] = 0.0;
  }
  if (len <= 0) return;
  const long long start = P.seg_start[w];
  const int half = N / 2;
  const int L = len + 2 * half;
  const int nadd = ((P.hop - (L - N) % P.hop) % P.hop) % N;
  const int K = (L + nadd - N) / P.hop + 1;

  // twiddles exp(-2 pi i j / N)
  for (int j = tid; j < N / 2; j += 256) {
    double s, c;
    sincospi(-2.0 * (double)j / (double)N, &s, &c);
    s_tw[j] = make_double2(c, s);
  }
  // which frames are referenced by a target time
  for (int k = tid; k <= K && k <= P.max_frames; k += 256) s_need[k] = 0;
  __syncthreads();
  const int* tix = P.t_idx + (size_t)w * P.n_t;
  const double* tfr = P.t_frac + (size_t)w * P.n_t;
  for (int j = tid; j < P.n_t; j += 256) {
    int i = tix[j];
    if (i >= 0 && i + 1 < K) {
      s_need[i] = 1;
      s_need[i + 1] = 1;
    }
  }
  // mean (np.mean: exact for int16 in fp64)
  double mean = 0.0;
  if (P.remove_dc) {
    double acc = 0.0;
    for (int i = tid; i < len; i += 256) acc += load_sample(P, start + i);
    acc = warp_sum(acc);
    if ((tid & 31) == 0) s_red[tid >> 5] = acc;
    __syncthreads();
    for (int i = 0; i < 8; ++i) mean += s_red[i];
    mean /= (double)len;
  }
  __syncthreads();

  for (int k = 0; k < K; ++k) {
    if (!s_need[k]) continue;  // uniform across the CTA
    // ---- windowed frame, bit-reversed order
    for (int j = tid; j < N; j += 256) {
      int idx = k * P.hop + j - half;
      double v = 0.0;
      if (idx >= 0 && idx < len) v = (load_sample(P, start + idx) - mean) * P.window[j];
      int r = (int)(__brev((unsigned)j) >> (32 - P.log2n));
      s_x[r] = make_double2(v, 0.0);
    }
    __syncthreads();
    // ---- radix-2 DIT
    for (int s = 1; s <= P.log2n; ++s) {
      const int mh = 1 << (s - 1);
      const int tstep = N >> s;
      for (int b = tid; b < N / 2; b += 256) {
        int grp = b >> (s - 1), pos = b & (mh - 1);
        int i0 = (grp << s) + pos, i1 = i0 + mh;
        double2 wv = s_tw[pos * tstep];
        double2 a = s_x[i0], c = s_x[i1];
        double tr = wv.x * c.x - wv.y * c.y, ti = wv.x * c.y + wv.y * c.x;
        s_x[i0] = make_double2(a.x + tr, a.y + ti);
        s_x[i1] = make_double2(a.x - tr, a.y - ti);
      }
      __syncthreads();
    }
    // ---- log magnitude (ava/preprocessing/utils.py:79)
    double* lg = s_log + (k & 1) * NB;
    for (int j = tid; j < NB; j += 256) {
      double2 z = s_x[j];
      lg[j] = log(hypot(z.x * P.scale, z.y * P.scale) + 1e-12);
    }
    if (tid == 0) *s_cnt = 0;
    __syncthreads();
    // ---- targets whose lower frame is k-1 can now be written
    if (k >= 1) {
      for (int j = tid; j < P.n_t; j += 256)
        if (tix[j] == k - 1) s_list[atomicAdd(s_cnt, 1)] = j;
      __syncthreads();
      const int cnt = *s_cnt;
      const double* lo = s_log + ((k - 1) & 1) * NB;
      const double* hi = lg;
      for (int i = tid; i < cnt * P.n_f; i += 256) {
        int f = i / cnt, t = s_list[i - f * cnt];
        int fi = P.f_idx[f];
        if (fi < 0) continue;
        double wy = P.f_frac[f], wx = tfr[t];
        double s00 = lo[fi], s01 = hi[fi], s10 = lo[fi + 1], s11 = hi[fi + 1];
        double v = (1.0 - wy) * ((1.0 - wx) * s00 + wx * s01) + wy * ((1.0 - wx) * s10 + wx * s11);
        v = (v - P.spec_min) * P.inv_range;
        v = fmin(fmax(v, 0.0), 1.0);
        if (P.out) P.out[out_base + (size_t)f * P.n_t + t] = (float)v;
        if (P.out64) P.out64[out_base + (size_t)f * P.n_t + t] = v;
      }
    }
    __syncthreads();
  }
}

}  // namespace ava

// Target-time tables of a batch of fixed-duration windows, on the device.  Bit-for-bit the
// float64 arithmetic of the host path (preprocessing/utils.py::bracket), which follows
// ava/models/window_vae_dataset.py:231-235 (target_times = np.linspace(onset, offset, n_t)) and
// the scipy interp2d bracketing / fill rule behind ava/preprocessing/utils.py:80-81,99:
//   T_j   = fl(fl(j * step) + tstart), step = fl(fl(tstop - tstart) / (n_t-1)), T_{n_t-1} = tstop
//   g_k   = fl(base[k] + grid0)                     (frame times, the reference's t += max(0,t1))
//   i     = the interval [g_i, g_{i+1}] containing T (i <= K-2), w = (T-g_i)/(g_{i+1}-g_i)
//   T < g_0 or T > g_{K-1}: i = -1, w = 0 (fill value)
// No FMA contraction anywhere (numpy multiplies, then adds).
namespace ava {
__global__ void __launch_bounds__(128)
time_tables_kernel(const double* __restrict__ grid0, const int* __restrict__ K, const double* __restrict__ base,
                   const double* __restrict__ tstart, const double* __restrict__ tstop, int n, int n_t,
                   int* __restrict__ t_idx, double* __restrict__ t_frac) {
  const int w = blockIdx.x;
  if (w >= n) return;
  const double g0 = grid0[w];
  const int Kw = K[w];
  const long long km2 = Kw - 2;
  const double start = tstart[w], stop = tstop[w];
  const double step = __ddiv_rn(__dsub_rn(stop, start), (double)(n_t - 1));
  const double dt = __dsub_rn(base[1], base[0]);
  const double first = __dadd_rn(base[0], g0);
  const double last = __dadd_rn(base[Kw - 1], g0);
  for (int j = threadIdx.x; j < n_t; j += blockDim.x) {
    const double T = (j == n_t - 1 && n_t > 1) ? stop : __dadd_rn(__dmul_rn((double)j, step), start);
    double f = floor(__ddiv_rn(__dsub_rn(T, first), dt));
    f = fmin(fmax(f, 0.0), (double)km2);
    long long i = (long long)f;
    for (int it = 0; it < 2; ++it) {
      if (__dadd_rn(base[i], g0) > T) --i;
      i = i < 0 ? 0 : (i > km2 ? km2 : i);
      if (__dadd_rn(base[i + 1], g0) <= T && i < km2) ++i;
    }
    const double lo = __dadd_rn(base[i], g0), hi = __dadd_rn(base[i + 1], g0);
    const double wgt = __ddiv_rn(__dsub_rn(T, lo), __dsub_rn(hi, lo));
    const bool bad = (T < first) || (T > last);
    t_idx[(size_t)w * n_t + j] = bad ? -1 : (int)i;
    t_frac[(size_t)w * n_t + j] = bad ? 0.0 : wgt;
  }
}
}  // namespace ava

extern "C" int ava_b200_window_time_tables(const double* grid0, const int* K, const double* base, int kmax,
                                           const double* tstart, const double* tstop, int n, int n_t, int* t_idx,
                                           double* t_frac, void* stream) {
  using namespace ava;
  AVA_REQUIRE(kmax >= 3 && n_t >= 2, "window_time_tables: kmax=%d n_t=%d", kmax, n_t);
  if (n <= 0) return 0;
  time_tables_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(grid0, K, base, tstart, tstop, n, n_t, t_idx, t_frac);
  return check_launch("window_time_tables");
}

extern "C" int ava_b200_get_spec_batch(const void* audio, int is_f32, const long long* seg_start,
                                       const int* seg_len, int n, int nperseg, int noverlap, int remove_dc,
                                       const double* window, double scale, const int* t_idx,
                                       const double* t_frac, int n_t, const int* f_idx, const double* f_frac,
                                       int n_f, int max_frames, double spec_min, double spec_max, float* out,
                                       double* out64, void* stream) {
  using namespace ava;
  AVA_REQUIRE(nperseg >= 16 && nperseg <= 4096 && (nperseg & (nperseg - 1)) == 0,
              "get_spec: nperseg=%d must be a power of two in [16,4096]", nperseg);
  AVA_REQUIRE(noverlap >= 0 && noverlap < nperseg, "get_spec: bad noverlap %d", noverlap);
  AVA_REQUIRE(spec_max != spec_min, "get_spec: spec_max_val == spec_min_val");
  if (n <= 0) return 0;
  SpecParams P;
  P.audio = audio;
  P.is_f32 = is_f32;
  P.seg_start = seg_start;
  P.seg_len = seg_len;
  P.nperseg = nperseg;
  P.hop = nperseg - noverlap;
  P.log2n = 0;
  while ((1 << P.log2n) < nperseg) ++P.log2n;
  P.remove_dc = remove_dc;
  P.window = window;
  P.scale = scale;
  P.t_idx = t_idx;
  P.t_frac = t_frac;
  P.n_t = n_t;
  P.f_idx = f_idx;
  P.f_frac = f_frac;
  P.n_f = n_f;
  P.max_frames = max_frames;
  P.spec_min = spec_min;
  P.inv_range = 1.0 / (spec_max - spec_min);
  P.out = out;
  P.out64 = out64;
  size_t smem = (size_t)nperseg * 16 + (size_t)(nperseg / 2) * 16 + (size_t)2 * (nperseg / 2 + 1) * 8 + 64 +
                (size_t)n_t * 4 + 4 + (size_t)max_frames + 16;
  smem = (smem + 15) / 16 * 16;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(get_spec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess) {
      cudaGetLastError();
      set_error("get_spec: %zu bytes of shared memory unavailable", smem);
      return 1;
    }
    configured = smem;
  }
  get_spec_kernel<<<n, 256, smem, (cudaStream_t)stream>>>(P);
  return check_launch("get_spec");
}
