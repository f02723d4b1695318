/*
 * ava_b200 -- C ABI of the B200-native AVA hot path (VAE train/infer + get_spec).
 *
 * The reference (pearsonlab/autoencoded-vocal-analysis) is pure Python and has no
 * FFI of its own; the drop-in boundary for a maintainer is the Python module
 * surface (ava.models.vae.VAE, ava.models.vae_dataset, ava.models.window_vae_dataset,
 * p['get_spec']).  This header is the native boundary underneath that surface:
 * every entry point replaces one group of library calls the reference makes from
 * Python, cited per function as <reference file>:<lines>.
 *
 * Conventions
 *   - plain C types only; all pointers are DEVICE pointers on the current CUDA
 *     device unless the name starts with h_ (host).  The library never allocates
 *     persistent memory, never frees caller memory and never copies host<->device.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and
 *     is CUDA-graph capturable; no implicit synchronisation.
 *   - every call returns 0 on success, nonzero on error; ava_b200_last_error()
 *     gives the message (thread-local).  No C++ exceptions cross the boundary.
 *   - tensors are contiguous fp32 NCHW / row-major exactly as the reference's
 *     torch tensors are; statistics accumulators are fp64.
 *
 * Layer ids (hard-coded network, ava/models/vae.py:128-168):
 *   0..6   = bn1+conv1 .. bn7+conv7      (encoder, ava/models/vae.py:217-223)
 *   7..13  = bn8+convt1 .. bn14+convt7   (decoder, ava/models/vae.py:263-269)
 */
#ifndef AVA_B200_H
#define AVA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVA_B200_ABI_VERSION 3
#define AVA_NUM_BN_LAYERS 14
#define AVA_STATS_STRIDE 64 /* doubles per layer in a stats block: [0..31]=sum, [32..63]=sum of squares */

const char* ava_b200_last_error(void);
int ava_b200_abi_version(void);
/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
long long ava_b200_launch_count(void);

/* ------------------------------------------------------------------ BatchNorm
 * Per-channel sums of x[B,C,HW]: stats[c] += sum x, stats[32+c] += sum x^2.
 * Replaces the statistics half of torch.nn.BatchNorm2d for bn1 (input spectrogram)
 * and bn8 (fc8 output viewed as [B,32,16,16]); ava/models/vae.py:217,262-263.
 * All other BN layers get their statistics from the producing layer's epilogue. */
int ava_b200_channel_stats(const float* x, int B, int C, int HW, double* stats, void* stream);

/* Running-buffer update for all 14 BN layers in one launch (momentum 0.1, unbiased
 * variance, num_batches_tracked += 1): torch.nn.BatchNorm2d train-mode side effect,
 * ava/models/vae.py:135-141,162-168.  stats: [14][64] doubles as above; counts[l] =
 * B*H*W of layer l's input; running: flat fp32 buffer, layer l's running_mean at
 * rm_off[l] and running_var at rv_off[l] (element offsets); nbt: int64[14]. */
int ava_b200_bn_update_running(const double* stats, const int* h_channels, const long long* h_counts,
                               float* running, const int* h_rm_off, const int* h_rv_off,
                               long long* nbt, float momentum, void* stream);

/* ------------------------------------------------- fused BN -> conv -> ReLU layers
 * Forward of layer `layer` (0..13): y = act(conv(bn(x)) + b) with
 *   bn: train!=0 -> batch statistics from stats_in (sums over x, N=B*H*W), else
 *       running_mean/running_var; eps 1e-5; zero padding applied AFTER bn.
 *   conv: Conv2d 3x3 pad 1 stride 1|2 (layers 0-6), ConvTranspose2d 3x3 pad 1
 *       stride 1 | stride 2 output_padding 1 (layers 7-13).
 *   act: ReLU except layer 13.
 * If stats_out != NULL, sums of y (the next BN's batch statistics) are accumulated
 * into it in the epilogue.  Replaces bn_k + conv_k + F.relu, ava/models/vae.py:217-223,
 * 263-269. */
int ava_b200_bnconv_fwd(int layer, int B, const float* x, float* y, const float* w, const float* b,
                        const float* gamma, const float* beta, const double* stats_in,
                        const float* running_mean, const float* running_var, int train,
                        double* stats_out, void* stream);

/* Arithmetic of the conv inner products (process-wide, like the reference side's
 * torch.backends.cudnn.allow_tf32 switch that governs the same layers on a GPU):
 *   0  fp32 FMA (default)
 *   1  TF32 tensor cores (mma.sync m16n8k8), operands rounded to TF32: rtol ~2e-3
 *   2  error-compensated 3xTF32 on the tensor cores (x = x_hi + x_lo; a_lo*b_hi + a_hi*b_lo +
 *      a_hi*b_hi, fp32 accumulate): fp32-level accuracy (rtol 1e-4 parity holds)
 *   3  as 2, but where a kernel has the variant (the weight-gradient kernels) the two correction
 *      terms a_lo*b + a*b_lo run as BF16 mma.m16n8k16 -- two k-chunks per instruction, 4 instead
 *      of 6 tensor-core instructions per pair of chunks; per-product error <= 2^-18, unbiased:
 *      rtol 1e-4 parity holds (tests/test_gpu_kernels.py, tests/test_gpu_model.py mode tf32x3b)
 *   4  as 3, and the forward / backward-data kernels form both correction terms of a tap in ONE
 *      BF16 mma.m16n8k16 (k16 = {a_lo | a} x {b | b_lo}): 2 instead of 3 tensor-core instructions
 *      per tap; per-product error <= 2^-19 (3xTF32: 2^-22): rtol 1e-4 parity holds (same tests,
 *      mode tf32x3c) with less margin -- batch-1024 gradients within 7.2e-5 of float64 (mode 3:
 *      4.2e-5), twice as many ReLU units on the other side of zero
 *   5  as 4 for the weight-gradient and backward-data kernels only; the forward kernels (whose
 *      rounding decides ReLU states and feeds every later layer) keep three TF32 terms
 *   6  as 5, plus the forward kernels of the DECODER (layers 7..13): their roundings reach only the
 *      reconstruction, not the latent path -- batch-1024 gradients 4.0e-5 (as mode 5), 260 ReLU units
 *      on the other side of zero (mode 5: 191, mode 4: 403, the float32 reference itself: 129)
 * Layers whose channel counts are not multiples of 8 always use fp32 FMA. */
int ava_b200_set_conv_precision(int mode);
int ava_b200_get_conv_precision(void);

/* Backward of layer `layer`.  `dz` is the gradient w.r.t. this layer's PRE-activation conv output
 * (for layer 13 -- no ReLU, nothing after it -- the loss gradient itself).  Order per layer:
 *
 *   1. ava_b200_dz_border_sums(dz)  -> tsums: nine per-channel sums of dz (total / first and last
 *      row and column / corners; for the stride-2 conv-transpose layers 8, 10, 12 the four
 *      row-parity x column-parity classes and the last row / column), from which the sum of dz over
 *      the pixels whose tap-k partner lies inside the image follows for each of the 9 taps.
 *   2. ava_b200_bnconv_bwd_weight   -> dw, db (OVERWRITTEN) and this layer's two BatchNorm-backward
 *      reductions dstats = (sum g, sum g*(x-mean)), ACCUMULATED (zero them first).  g is the gradient
 *      w.r.t. the BatchNorm output; it is never materialised: both reductions are linear in the
 *      centred raw product Rc = sum dz*(x-mean)_pad the kernel accumulates anyway,
 *         dw = gamma*invstd*Rc + beta*T_k,  sum g = <w, T_k>,  sum g*(x-mean) = <w, Rc>.
 *   3. ava_b200_bnconv_bwd_data     -> dz_prev = [x > 0] * BN-backward(g) with g formed tile by tile
 *      on chip: the gradient w.r.t. the PREVIOUS layer's pre-activation output (same shape as x;
 *      relu_mask = 0 skips the [x > 0] factor).  Not needed for layer 0.
 *
 * Replaces autograd's cudnn_convolution_backward (input + weight), native_batch_norm_backward and
 * threshold_backward for ava/models/vae.py:217-223,263-269 (called from vae.py:352).
 * mode: 1 for the stride-2 conv-transpose layers (8, 10, 12), else 0.  tsums: [9][32] doubles,
 * ACCUMULATED (zero them first). */
#define AVA_TSUM_STRIDE 288 /* doubles per layer in a tsums block: [slot 0..8][channel 0..31] */
int ava_b200_dz_border_sums(const float* dz, int B, int C, int H, int W, int mode, double* tsums, void* stream);

/* x/stats_in: this layer's input and its batch statistics; w/gamma/beta: this layer's parameters.
 * ws: scratch of ava_b200_bnconv_bwd_weight_ws(layer,B) bytes. */
int ava_b200_bnconv_bwd_weight(int layer, int B, const float* dz, const float* x, const float* w,
                               const float* gamma, const float* beta, const double* stats_in,
                               const double* tsums, float* dw, float* db, double* dstats, void* ws,
                               void* stream);
long long ava_b200_bnconv_bwd_weight_ws(int layer, int B);

/* gamma == NULL: no BatchNorm backward (dz_prev = [x > 0] * g).
 * tsums_prev (optional): the nine border sums of dz_prev -- layer-1's tsums block -- are accumulated
 * in the epilogue, which saves step 1 (a full read of dz_prev) for layer-1. */
int ava_b200_bnconv_bwd_data(int layer, int B, const float* dz, const float* w, const float* x,
                             const float* gamma, const double* stats_in, const double* dstats, int relu_mask,
                             float* dz_prev, double* tsums_prev, void* stream);

/* BN backward finalisation: dgamma[c] = invstd*dstats[32+c], dbeta[c] = dstats[c]
 * for all 14 layers (dstats[32+c] holds sum g*(x-mean)). */
int ava_b200_bn_param_grads(const double* stats, const double* dstats, const int* h_channels,
                            const long long* h_counts, float* grads, const int* h_dgamma_off,
                            const int* h_dbeta_off, void* stream);

/* out = [a>0] * (p*(g - c1) + q*(a - mean)), per channel c = (i / HW) % C: the backward of
 * the BatchNorm that consumes `a` (train mode; fp64 coefficient math) followed by the ReLU
 * backward of the layer that produced `a`, as one elementwise pass; out may alias g.
 * gamma == NULL: no BatchNorm, only the ReLU mask.  relu == 0: no mask.  The conv layers apply
 * this inside ava_b200_bnconv_bwd_data; the stand-alone pass serves the fc1 -> conv7 seam (ReLU
 * mask of conv7's output on the gradient arriving from fc1).  tsums != NULL: the nine border sums
 * of `out` (rows of width W; mode 0 of ava_b200_dz_border_sums) are accumulated on the way. */
int ava_b200_bn_relu_bwd_apply(const float* g, const float* a, const float* gamma, const double* stats,
                               const double* dstats, int B, int C, int HW, int relu, float* out,
                               double* tsums, int W, void* stream);

/* ------------------------------------------------------------- dense (Linear) layers
 * Y[M,N] = act(X[M,K] . W[N,K]^T + b): torch.nn.Linear (+F.relu / torch.exp),
 * ava/models/vae.py:225-232,258-261.  act: 0 none, 1 relu, 2 exp.
 * `groups` > 1 runs a strided batch: group g uses X + g*x_gs, W + g*w_gs, b + g*b_gs,
 * Y + g*y_gs (the three posterior heads).  precision: 0 = fp32 SIMT,
 * 1 = tcgen05 TF32 tensor cores (rtol 3e-3), 2 = tcgen05 3xTF32 error-compensated
 * (fp32-level, rtol 2e-5).  The tensor-core paths need groups == 1 and M,N multiples of
 * 128, K of 32 (for all three of fwd / bwd_data / bwd_weight); other shapes run on the SIMT
 * kernel regardless of `precision`. */
int ava_b200_linear_fwd(const float* x, int ldx, const float* w, const float* b, float* y, int ldy, int M, int N,
                        int K, int act, int groups, long long x_gs, long long w_gs, long long b_gs,
                        long long y_gs, int precision, void* ws, long long ws_bytes, void* stream);
/* dX[M,K] = (dY (.) mask) . W ; mask = [Ymask>0] if Ymask != NULL (ReLU backward). */
int ava_b200_linear_bwd_data(const float* dy, int lddy, const float* ymask, const float* w, float* dx, int lddx,
                             int M, int N, int K, int groups, long long dy_gs, long long w_gs, long long dx_gs,
                             int accumulate, int precision, void* ws, long long ws_bytes, void* stream);
/* dW[N,K] = (dY (.) mask)^T . X ; db[N] = column sums of dY (.) mask.  Overwrites. */
int ava_b200_linear_bwd_weight(const float* dy, int lddy, const float* ymask, const float* x, int ldx, float* dw,
                               float* db, int M, int N, int K, int groups, long long dy_gs, long long x_gs,
                               long long dw_gs, long long db_gs, int precision, void* ws, long long ws_bytes,
                               void* stream);

/* The weight gradients of up to AVA_MAX_WGRAD_JOBS small Linear layers in ONE launch (fp32 FMA, no
 * split-K, no workspace): job j is ava_b200_linear_bwd_weight's dW for (dy, ymask, x) of shape
 * M x N / M x K with `groups` strided groups (ymask may be NULL).  h_jobs is a HOST array. */
#define AVA_MAX_WGRAD_JOBS 8
typedef struct {
  const float* dy;
  const float* ymask;
  const float* x;
  float* dw;
  int lddy, ldx, M, N, K, groups;
  long long dy_gs, x_gs, dw_gs;
} ava_b200_wgrad_job;
int ava_b200_linear_bwd_weight_multi(const ava_b200_wgrad_job* h_jobs, int njobs, void* stream);

/* Bias gradients of up to AVA_MAX_BIAS_JOBS Linear layers in two launches: job j computes
 * db[n] = sum_m (mask > 0 ? dy : 0)[m, n] over dy [M, N] (row stride ld; mask may be NULL) --
 * what ava_b200_linear_bwd_weight does per layer when given db; the step passes db = NULL there
 * and collects the layers of a backward segment into one call.  h_jobs is a HOST array.
 * Workspace: ava_b200_bias_grads_ws_bytes.  Replaces the bias half of autograd's Linear backward
 * (ava/models/vae.py:226-232,258-260). */
#define AVA_MAX_BIAS_JOBS 8
typedef struct {
  const float* dy;
  const float* mask;
  float* db;
  int ld, M, N;
} ava_b200_bias_job;
long long ava_b200_bias_grads_ws_bytes(const ava_b200_bias_job* h_jobs, int njobs);
int ava_b200_bias_grads(const ava_b200_bias_job* h_jobs, int njobs, void* ws, long long ws_bytes, void* stream);
long long ava_b200_linear_ws_bytes(int M, int N, int K);

/* ----------------------------------------------------------------------- ELBO
 * Latent part (torch.distributions.LowRankMultivariateNormal(mu,u,d).rsample() /
 * .entropy() + the prior term; ava/models/vae.py:312-316,323):
 *   d = exp(logd); z = mu + u*eps_w + sqrt(d)*eps_d;
 *   acc[0] += sum z^2 ; acc[2] += sum_b H_b,  H_b = 1/2(Z(1+ln 2pi) + ln(1+sum u^2/d) + sum ln d)
 * heads: [B, 3*Z] row-major = (mu | u | logd) as produced by the grouped fc4x layer. */
int ava_b200_latent_fwd(const float* heads, const float* eps_w, const float* eps_d, int B, int Z, float* z,
                        float* d_out, double* acc, void* stream);
/* Gradient of the loss w.r.t. the heads (mu | u | logd) given gz = dL/dz from the
 * decoder (prior term z added here): analytic, SURVEY.md 8(a) "Analytic gradients". */
int ava_b200_latent_bwd(const float* heads, const float* eps_w, const float* eps_d, const float* z,
                        const float* gz, int B, int Z, float* g_heads, void* stream);
/* Reconstruction term (ava/models/vae.py:319-320): acc[1] += sum (x - x_rec)^2 and,
 * if g != NULL, g = precision * (x_rec - x) = dL/dx_rec.  tsums != NULL (with g): also the nine
 * border sums of g viewed as [n/(H*W)] single-channel H x W planes (mode 0 of
 * ava_b200_dz_border_sums, channel 0), accumulated while g is written -- g is the decoder's last
 * layer's dz, whose weight-gradient finalisation needs them. */
int ava_b200_recon(const float* x, const float* x_rec, long long n, float precision, float* g, double* acc,
                   double* tsums, int H, int W, void* stream);
/* loss = 1/2(acc0 + Z ln 2pi) + 1/2 XDIM ln(2pi/prec) + 1/2 prec acc1 - acc2, written as
 * fp32 to loss[0] and accumulated (fp64) into loss_sum[0] if non-NULL (device-side epoch
 * accumulator replacing loss.item() per step, ava/models/vae.py:351). */
int ava_b200_elbo_finalize(const double* acc, int Z, int xdim, float precision, float* loss, double* loss_sum,
                           void* stream);

/* ------------------------------------------------------- fused small dense layers
 * The chain between the two 8192x1024 layers as ONE row-wise kernel per direction (csrc/mlp.cu):
 *   forward  (ava/models/vae.py:226-232, 298-316, 258-260): stages bit 0: h1 -fc2-> h2 -fc31|32|33->
 *            h3 -fc41|42|43-> heads = (mu | u | log d); bit 1: z = mu + u eps_w + sqrt(d) eps_d,
 *            acc[0] += sum z^2, acc[2] += entropy (as ava_b200_latent_fwd); bit 2: z -fc5-> t5 -fc6->
 *            t6 -fc7-> t7.  Every intermediate is also written out (saved for the backward pass).
 *   backward (autograd of the same): dt7 -> dt6 -> dt5 -> gz -> gheads (analytic latent gradient, as
 *            ava_b200_latent_bwd) -> dh3 -> dh2 -> dh1; gradients w.r.t. post-activation outputs,
 *            stored unmasked (ava_b200_linear_bwd_weight / ava_b200_bias_grads apply the ReLU masks).
 * w3/b3 and w4/b4 are the stacked fc31|fc32|fc33 and fc41|fc42|fc43 parameters.  z_dim: a multiple
 * of 4, <= 64.  h_params is a HOST struct of DEVICE pointers.  fp32 FMA arithmetic. */
typedef struct {
  int B, Z, stages;
  const float *w2, *b2, *w3, *b3, *w4, *b4, *w5, *b5, *w6, *b6, *w7, *b7;
  const float *eps_w, *eps_d;
  float *h1, *h2, *h3, *heads, *z, *d, *t5, *t6, *t7;
  float *dt7, *dt6, *dt5, *gz, *gheads, *dh3, *dh2, *dh1;
  double* acc;
} ava_b200_mlp_params;
int ava_b200_mlp_fwd(const ava_b200_mlp_params* h_params, void* stream);
int ava_b200_mlp_bwd(const ava_b200_mlp_params* h_params, void* stream);

/* ----------------------------------------------------------------------- Adam
 * torch.optim.Adam (defaults: no weight decay, no amsgrad), one launch over the flat
 * parameter buffer; ava/models/vae.py:119,353.  step_count: device fp32 scalar holding
 * the step number BEFORE this update (incremented by the kernel; mirrors torch's
 * per-tensor `step` state).  lr/betas/eps are doubles, as torch holds them. grad_scale multiplies g on load (1.0 normally). */
int ava_b200_adam_step(float* p, const float* g, float* m, float* v, long long n, float* step_count, double lr,
                       double beta1, double beta2, double eps, float grad_scale, void* stream);

/* Same update with the hyper-parameters read from DEVICE memory: hyper = {lr, beta1, beta2, eps}
 * (doubles), so that a captured CUDA graph of the step follows optimizer.param_groups (learning-
 * rate schedules, the lr a checkpoint restores: ava/models/vae.py:470) without re-capture. */
int ava_b200_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float* step_count,
                           const double* hyper, float grad_scale, void* stream);

/* Data-parallel optimizer step fused with its collectives over NVLink / NVSwitch peer memory
 * (csrc/dp.cu): the data-parallel form of optimizer.step(), ava/models/vae.py:353 -- replaces
 * "NCCL all-reduce(SUM) of the flat gradient, then Adam on every rank".  Rank r sums the
 * world gradient copies of ITS 1/world slice (peer loads, or one multimem.ld_reduce through the
 * switch when multicast pointers are given), applies Adam to that slice only (moments are sharded)
 * and stores the new parameters into every rank's parameter buffer; two flag barriers order it
 * against the ranks' backward / next forward passes.
 *   h_peers: HOST struct of DEVICE pointers: every rank's flat gradient buffer, flat parameter
 *            buffer and flag block (2*AVA_DP_MAX_WORLD uint32, zero-initialised once), all in
 *            symmetric (peer-mapped) memory; index = rank; grad_mc/param_mc = multicast views or NULL
 *   m, v:    this rank's LOCAL moment buffers (only its slice is read and written)
 *   local:   LOCAL device uint32[4], zero-initialised once: [0] step sequence number,
 *            [1] CTA counter, [2] set to 1 if a peer did not arrive within 20 s (the kernel
 *            then carries on instead of hanging; the caller checks it)
 * Every rank must make the same sequence of calls.  n % 4 == 0; world <= AVA_DP_MAX_WORLD. */
#define AVA_DP_MAX_WORLD 8
typedef struct {
  const float* grad[AVA_DP_MAX_WORLD];
  float* param[AVA_DP_MAX_WORLD];
  unsigned int* flags[AVA_DP_MAX_WORLD];
  const float* grad_mc;
  float* param_mc;
} ava_b200_dp_peers;
int ava_b200_adam_step_dp(const ava_b200_dp_peers* h_peers, int rank, int world, float* m, float* v,
                          long long n, float* step_count, const double* hyper, float grad_scale,
                          unsigned int* local, void* stream);

/* ------------------------------------------------------------------- get_spec
 * Batched spectrogram front end: ava/preprocessing/utils.py:18-110 (get_spec) with
 * scipy.signal.stft semantics (periodic Hann, zero boundary extension, zero padding to a
 * hop multiple, scale 1/sum(win)).  One CTA per window; computed in fp64.
 *   audio:      concatenated audio of all files on device: int16 (is_f32 = 0), fp32 (1) or fp64 (2)
 *   seg_start:  [n] first sample of each segment in `audio`
 *   seg_len:    [n] number of samples (0 => output zeros, the reference's
 *               "too short" branch, ava/preprocessing/utils.py:69-71)
 *   window/scale: the analysis window [nperseg] (fp64) and 1/sum(window), computed by
 *               the host with the same scipy call the reference makes
 *   t_idx/t_frac [n,n_t]: lower frame index and weight of each target time;
 *               t_idx < 0 marks an out-of-range target (fill value -> 0 after clip)
 *   f_idx/f_frac [n_f]: same for the target frequencies (shared by all windows)
 *   max_frames: upper bound on the number of STFT frames of any segment
 *   out: [n, n_f, n_t] fp32 (numpy_to_tensor's float32, ava/models/utils.py:444-446)
 *   out64: optional [n, n_f, n_t] fp64 (the array get_spec itself returns) */
int ava_b200_get_spec_batch(const void* audio, int is_f32, const long long* seg_start, const int* seg_len, int n,
                            int nperseg, int noverlap, int remove_dc, const double* window, double scale,
                            const int* t_idx, const double* t_frac, int n_t, const int* f_idx,
                            const double* f_frac, int n_f, int max_frames, double spec_min, double spec_max,
                            float* out, double* out64, void* stream);

/* within_syll_normalize (ava/preprocessing/utils.py:106-109), per spectrogram of m = n_f*n_t values:
 *   spec -= np.quantile(spec, q); spec[spec < 0] = 0; spec /= spec.max() + 1e-12
 * in float64 on spec64 [n, m] in place (the quantile by an exact radix select + numpy's linear
 * interpolation); spec32 (optional) receives the float32 copy. */
int ava_b200_quantile_normalize(double* spec64, float* spec32, int n, int m, double q, void* stream);

/* Target-time tables of a batch of fixed-duration windows computed on the device (the
 * shotgun path, ava/models/window_vae_dataset.py:231-235: target_times = linspace(onset,
 * offset, n_t); bracketing and fill rule of scipy interp2d, ava/preprocessing/utils.py:80-81,99).
 * Same float64 operations as the host tables, bit for bit.
 *   grid0 [n]: max(0, t1) of each window; K [n]: STFT frames of each window (>= 3);
 *   base [kmax]: frame times of a segment starting at 0; tstart/tstop [n]: first / last target
 *   t_idx/t_frac [n,n_t]: outputs in the layout ava_b200_get_spec_batch consumes. */
int ava_b200_window_time_tables(const double* grid0, const int* K, const double* base, int kmax,
                                const double* tstart, const double* tstop, int n, int n_t, int* t_idx,
                                double* t_frac, void* stream);

/* ------------------------------------------------- MMD^2 between sets of latent means
 * Downstream statistic on get_latent's output, ava/plotting/mmd_plots.py:255-312, 450-476.
 * x: [N,D] float64 latent means; seg[i] in [0,n_seg) = condition of row i.
 * S[a*n_seg+b] = sum_{i in a, j in b} exp(A*||x_i-x_j||^2), A = -0.5/sigma^2 (S is zeroed by
 * the call).  MMD^2(a,b) = (S_aa-n_a)/(n_a(n_a-1)) + (S_bb-n_b)/(n_b(n_b-1)) - 2 S_ab/(n_a n_b). */
int ava_b200_mmd_block_sums(const double* x, int N, int D, const int* seg, int n_seg, double A, double* S,
                            void* stream);
/* out[k] = ||x[ia[k]]-x[ib[k]]||^2 (mode 0) or exp(A * that) (mode 1); numpy summation order. */
int ava_b200_pair_kernel(const double* x, int D, const long long* ia, const long long* ib, long long n,
                         double A, int mode, double* out, void* stream);

/* ------------------------------------------------- PCA of the latent means
 * sklearn.decomposition.PCA(n_components=K).fit_transform on get_latent's output,
 * ava/data/data_container.py:538-551 (_make_latent_mean_pca_projection).
 * x: [N,D] rows, float64 (is_f32 = 0) or the encoder's float32 latents (is_f32 = 1), D <= 64.
 * fit: mean[D]; cov[D*D] = sample covariance (divisor N-1); evals[D] descending, clipped at 0;
 * comps[D*D]: row k = unit eigenvector k, its largest-magnitude entry positive (scikit-learn's
 * svd_flip with u_based_decision=False).  ws: ava_b200_pca_ws_bytes(D) bytes of scratch.
 * transform: out[N,K] = (x - mean) . comps[:K]^T. */
long long ava_b200_pca_ws_bytes(int D);
int ava_b200_pca_fit(const void* x, int is_f32, long long N, int D, double* mean, double* cov, double* evals,
                     double* comps, void* ws, long long ws_bytes, void* stream);
int ava_b200_pca_transform(const void* x, int is_f32, long long N, int D, const double* mean, const double* comps,
                           int K, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVA_B200_H */
