/* starcop_b200 C ABI -- the drop-in boundary of the B200-native STARCOP hot path.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all memory
 *     (PyTorch allocates it); the library never allocates persistent device memory;
 *   - every entry point launches on the `stream` it is given (a cudaStream_t passed as void*),
 *     never synchronises the device, never throws; it returns SC_OK (0) or a negative sc_status;
 *   - activation tensors are NHWC with an explicit channel stride `ld` (elements), so a channel
 *     slice of a wider buffer (the decoder's concat buffers) is addressed without a copy;
 *   - `dtype` is SC_F32 or SC_BF16 and names the STORAGE type of activations; all arithmetic
 *     accumulates in fp32 (statistics and loss sums in fp64);
 *   - there is no CPU fallback anywhere behind this header.
 *
 * Each entry point cites the reference call it replaces (paths relative to the STARCOP tree).
 *
 * Implementation selectors (environment, read per call; the tests use them to check two independent kernels of
 * one operator against each other -- they never select a CPU path, there is none):
 *   STARCOP_MAG1C_STREAMING  sc_mag1c_filter: always the streaming kernel (no group-resident fast path)
 *   STARCOP_MAG1C_NO_TC      sc_mag1c_filter: the fp64-FMA group-resident kernel instead of the tensor-core one
 *   STARCOP_RATIO_NOCLUSTER / STARCOP_RATIO_CLUSTER   sc_ratio_product: force the single-CTA / the cluster select
 *   STARCOP_RATIO_CLUSTER8 / STARCOP_RATIO_CLUSTER12   sc_ratio_product: the eight-CTA (one band at a time) cluster kernel /
 *                            twelve-CTA clusters instead of the automatic choice between sixteen and twelve
 *   STARCOP_BN_NOFLAT        sc_bn_bwd_reduce: always the register-streaming kernel (no cp.async.bulk ring)
 *   STARCOP_NO_WGRAD_HALO    sc_tc_conv_wgrad: always the per-tap kernel (no halo-patch kernel for the thin layers)
 */
#ifndef STARCOP_B200_H
#define STARCOP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SC_OK = 0,
  SC_ERR_BAD_ARG = -1,      /* shape / alignment / dtype contract violated */
  SC_ERR_CUDA = -2,         /* a CUDA runtime call failed (see sc_last_cuda_error) */
  SC_ERR_UNSUPPORTED = -3,  /* configuration not implemented by this kernel */
  SC_ERR_NO_DEVICE = -4     /* no sm_100 device / driver entry point unavailable */
} sc_status;

enum { SC_F32 = 0, SC_BF16 = 1 };
enum { SC_ACT_NONE = 0, SC_ACT_RELU = 1, SC_ACT_RELU6 = 2 };

int sc_abi_version(void);
/* last cudaError_t seen by the library on this thread, as text (host pointer, static storage) */
const char* sc_last_cuda_error(void);

/* ---- A1: DataNormalizer.normalize_x (starcop/data/normalizer_module.py:134-135) -------------
 * x: (B,C,H,W) f32 NCHW raw products.  off/fac/lo/hi: C doubles each (device).  f64_path is a bit
 * mask of which reference parameter arrays are float64 (bit0 offsets, bit1 factors): the reference
 * builds them with np.array(python numbers), so one non-integer entry promotes the whole
 * subtraction / division to float64 before the final .float().
 * out_nhwc: (B,H,W,ld_out) `dtype`, channels [C,ld_out) zero filled; may be NULL.
 * out_nchw: (B,C,H,W) f32 normalised copy ("input_norm", model_module.py:196); may be NULL. */
int sc_normalize_pack(const float* x, const double* off, const double* fac, const double* lo,
                      const double* hi, int f64_path, int B, int C, int H, int W,
                      void* out_nhwc, int ld_out, int dtype, float* out_nchw, void* stream);

/* ---- A2: smp.Unet(mobilenet_v2) building blocks (model_module.py:98, 238-251) ---------------
 * Dense convolution, NHWC, weights packed [KH*KW][Cin][Cout] f32 (sc_pack_weights).
 * y[n,ho,wo,co] (+)= sum x[n, ho*stride-pad+kh, wo*stride-pad+kw, ci] * w[kh,kw,ci,co] (+ bias)
 * accumulate != 0 adds onto the existing y (gradient accumulation for residual / skip fan-out). */
int sc_conv_fprop(const void* x, int ldx, const float* w_packed, const float* bias, void* y, int ldy,
                  int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                  int dtype, int accumulate, void* stream);
/* dW[co,ci,kh,kw] (OIHW f32, torch layout) += sum_p x[p*stride-pad+k][ci] * dy[p][co]; caller zeroes dW.
 * Deterministic: the pixel splits write private partial gradients into `workspace`
 * (sc_conv_wgrad_workspace_bytes bytes, may be NULL when that is 0) and are summed in split order. */
int64_t sc_conv_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
int sc_conv_wgrad(const void* x, int ldx, const void* dy, int lddy, float* dw_oihw, float* workspace,
                  int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                  int dtype, void* stream);
/* OIHW f32 -> packed [KH*KW][Cin][Cout] f32; flip_transpose != 0 builds the data-gradient filter
 * (taps mirrored, Cin/Cout swapped) so that dgrad of a stride-1 conv is sc_conv_fprop on dy. */
int sc_pack_weights(const float* w_oihw, float* w_packed, int Cout, int Cin, int KH, int KW,
                    int flip_transpose, void* stream);

/* depthwise 3x3, pad 1, stride 1|2; weights (C,1,3,3) f32 torch layout.
 * The *_bnact variants read x through the producer's BatchNorm+activation on the fly:
 * xin = act(x*scale[c]+shift[c]) (scale==NULL -> identity), so the x6 expanded tensor of an
 * inverted-residual block is never written normalised.  stats != NULL: fprop also emits the
 * BatchNorm partial-sum rows of its stored outputs (see sc_bn_stats; *stats_rows_host = row count). */
int sc_dwconv_fprop(const void* x, int ldx, const float* scale, const float* shift, int act,
                    const float* w, void* y, int ldy, double* stats, int* stats_rows_host, int N, int H, int W,
                    int C, int stride, int dtype, void* stream);
int sc_dwconv_dgrad(const void* dy, int lddy, const float* w, void* dx, int lddx,
                    int N, int H, int W, int C, int stride, int dtype, void* stream);
/* dw (C,1,3,3) += ...; workspace: sc_dwconv_wgrad_workspace_bytes(C) bytes of per-block partial rows */
int64_t sc_dwconv_wgrad_workspace_bytes(int C);
int sc_dwconv_wgrad(const void* x, int ldx, const float* scale, const float* shift, int act,
                    const void* dy, int lddy, float* dw, float* workspace, int N, int H, int W, int C,
                    int stride, int dtype, void* stream);

/* training-mode BatchNorm2d (eps 1e-5, momentum 0.1 -- torchvision / smp defaults), split as
 * statistics -> finalize -> apply so that the normalise+activation runs in the consumer.
 * Statistics are produced as ROWS OF PARTIAL SUMS (no global atomics, deterministic): a producer
 * writes nrows <= SC_BN_MAX_PARTIALS rows of [sum(0..C) | sum of squares(C..2C)] fp64 into a buffer
 * of sc_bn_partials_bytes(C) bytes and reports nrows through *nrows_host; the consumer sums them. */
#define SC_BN_MAX_PARTIALS 296
int64_t sc_bn_partials_bytes(int C);
int sc_bn_stats(const void* y, int ldy, double* partials, int* nrows_host, int64_t P, int C, int dtype,
                void* stream);
/* training != 0: mean/var from the partial rows over P pixels, running stats updated (unbiased var);
 * training == 0: uses running_mean/var.  Emits scale = gamma*invstd, shift = beta - mean*scale,
 * and saves mean / invstd for the backward. */
int sc_bn_finalize(const double* partials, int nrows, int64_t P, int C, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, float momentum, float eps, int training,
                   float* scale, float* shift, float* save_mean, float* save_invstd, void* stream);
/* z = act(y*scale+shift) (+ residual); written at channel stride ldz; upsample2 != 0 writes each
 * pixel to its 2x2 nearest-neighbour block of a (N,2H,2W,ldz) buffer (F.interpolate(scale 2,
 * "nearest") + torch.cat of the decoder fused into the producer's store). */
int sc_bn_act(const void* y, int ldy, const float* scale, const float* shift, int act,
              const void* residual, int ldr, void* z, int ldz, int N, int H, int W, int C,
              int upsample2, int dtype, void* stream);
/* BN backward, phase 1: with g = dz * act'(y*scale+shift) (dz optionally 2x2 sum-pooled from a
 * (N,2H,2W,lddz) buffer when pooled != 0), partial rows of [sum g | sum g*xhat] (fp64, as above). */
int sc_bn_bwd_reduce(const void* dz, int lddz, int pooled, const void* y, int ldy,
                     const float* scale, const float* shift, const float* mean, const float* invstd,
                     int act, double* partials, int* nrows_host, int N, int H, int W, int C, int dtype,
                     void* stream);
/* phase 2: dy = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)); dgamma += sum g*xhat, dbeta += sum g.
 * Row [nrows] of `partials` receives the totals. */
int sc_bn_bwd_apply(const void* dz, int lddz, int pooled, const void* y, int ldy,
                    const float* scale, const float* shift, const float* mean, const float* invstd,
                    const float* gamma, int act, double* partials, int nrows, void* dy, int lddy,
                    float* dgamma, float* dbeta, int N, int H, int W, int C, int dtype, void* stream);
/* out (+)= a  [2x2 sum-pooled when pooled]; plain gradient routing for skip / residual fan-out */
int sc_add_into(const void* a, int lda, int pooled, void* out, int ldo, int accumulate,
                int N, int H, int W, int C, int dtype, void* stream);

/* segmentation head Conv2d(16->1, k3, pad1, bias) (smp SegmentationHead) -- bandwidth bound, no MMA */
int sc_head_fprop(const void* x, int ldx, const float* w, const float* bias, float* logits,
                  int N, int H, int W, int C, int dtype, void* stream);
/* data gradient of the head (the weight / bias gradient is sc_head_wgrad_tiled) */
int sc_head_bwd(const void* x, int ldx, const float* w, const float* dlogits, void* dx, int lddx,
                int N, int H, int W, int C, int dtype, void* stream);

/* the head's weight / bias gradient through the depthwise tile machinery (TMA halo-tile ring, per-CTA partial
 * rows: deterministic, no global atomics); dw / dbias are ACCUMULATED.  C in {8,16,24,32,48,64}. */
int64_t sc_head_wgrad_workspace_bytes(int C);
int sc_head_wgrad_tiled(const void* x, int ldx, const float* dlogits, float* dw, float* dbias, float* workspace,
                        int N, int H, int W, int C, int dtype, void* stream);

/* ---- A3/A4/A6: weighted BCE + decisions + confusion counts in ONE pass ----------------------
 * (model_module.py:76-79 train, :115-135 val, :191-212 batch_with_preds; torchmetrics
 * ConfusionMatrix.update).  logits/y/w: n = B*HW f32 (w may be NULL = no weight_loss).
 * loss_sum: NULL, or sc_bce_loss_words(B, HW) doubles: [0] += sum l*w (fp64), [1] a ticket word that must be
 * zero on entry and is zero again on exit, [2...] per-block partial sums (scratch).  The total is formed in
 * block order by the last block to finish: bit-identical from run to run (no floating-point atomics).
 * Optional outputs (NULL to skip):
 *   grad        d mean(l*w)/d logits * grad_scale            (ATen's backward form)
 *   cm          int64[4] += [[TN,FP],[FN,TP]] for pred = logits >= 0      (val_step, :124)
 *   pred_count  int64[B] += per-tile sum of that pred                     (pred_classification)
 *   cm_sig / pred_count_sig: the same for pred = sigmoid(logits) > .5     (batch_with_preds :204)
 *   prediction, loss_px, loss_px_w (f32), pred_binary, differences (int64) per pixel. */
int64_t sc_bce_loss_words(int B, int64_t HW);
int sc_bce_fused(const float* logits, const float* y, const float* w, float pos_weight,
                 int B, int64_t HW, float grad_scale, double* loss_sum, float* grad,
                 int64_t* cm, int64_t* pred_count, int64_t* cm_sig, int64_t* pred_count_sig,
                 float* prediction, float* loss_px, float* loss_px_w, int64_t* pred_binary,
                 int64_t* differences, void* stream);

/* ---- A16: torch.optim.Adam(lr) on one flat fp32 arena (model_module.py:172-174) -------------
 * step_host is the 1-based step count; grad_scale multiplies g first (DDP mean). */
int sc_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                 float beta2, float eps, int step_host, float grad_scale, void* stream);
/* same update with the step counter (incremented here) and the learning rate in device memory, so a
 * captured CUDA graph of the train step stays valid across steps and lr-scheduler changes */
int sc_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev,
                     float beta1, float beta2, float eps, int* step_dev, float grad_scale, void* stream);

/* ---- A8/A10: mag1c matched filter (starcop/models/mag1c.py:176-348) -------------------------
 * One pixel group per CTA (a detector column of a tile, process_aviris.py:211-212, or any pixel set
 * func_by_groups builds, mag1c.py:161-172).  x: radiance, band s of pixel `pix` at
 * x[pix*pixel_stride + s] for the S window bands (the caller offsets x to the first window band);
 * element type float (AVIRIS, process_aviris.py:199) or double (EMIT, mag1c_emit.py:46) per `fp64`.
 * pix_idx: (G, pmax) int32 pixel indices of each group, counts: G valid lengths (NULL = all pmax).
 * Groups with <= skip_le pixels are skipped and their outputs keep the caller's pre-fill (-9999): func_by_groups
 * passes 10 (mag1c.py:166), mag1c_emit processes every group that has a valid pixel (mag1c_emit.py:66-68).  tmpl: S doubles.  mf_out / albedo_out: per-pixel outputs indexed by `pix` (scatter).
 * num_iter = 0 is the plain matched filter `rmf`; alpha = diagonal loading (mag1c.py:246).
 * status: optional device int, incremented per group whose covariance was not positive definite. */
int64_t sc_mag1c_smem_bytes(int S, int pmax, int elem_bytes);   /* must be <= 220 KB */
int sc_mag1c_filter(const void* x, int64_t pixel_stride, const int32_t* pix_idx, const int32_t* counts,
                    int pmax, const double* tmpl, void* mf_out, void* albedo_out, int G, int S,
                    int num_iter, double alpha, int fp64, int skip_le, int* status, void* stream);
/* diagnostics: clock64 stamps of group 0 at the phase boundaries of the last tensor-core mag1c launch
 * (start, loaded, centred, MMAs done, covariance assembled, inverse done, rmf end, iteration 0 end, end) */
int sc_debug_mag1c_clocks(long long* out16);
/* x / d for 0 <= x < 2^31, d >= 1, evaluated with the multiply-high constants the persistent tile loops use (host only; tests) */
unsigned int sc_debug_fastdiv(unsigned int x, unsigned int d);

/* ---- A11/A13: band ratio product (starcop/data/feature_extration.py:32-56) ------------------
 * per tile: exact 5/95 percentiles (np.percentile linear) of each band by radix select, inlier
 * sums, c = sum_bg/sum_sig, R = (c*sig-bg)/(bg+1e-6), R = zero_value where both < 1e-6. */
int64_t sc_ratio_workspace_bytes(int T, int64_t HW);
int sc_ratio_product(const float* bg, const float* sig, float* out, int T, int64_t HW,
                     float percentile, float zero_value, void* workspace, void* stream);
int sc_weight_mag1c(const float* mag1c, float* out, int64_t n, void* stream);

/* ---- A12: MLR band reconstruction (starcop/data/feature_extration.py:58-124, sklearn
 * LinearRegression with intercept): bands (T,K,HW) f32, target (T,HW) -> recon (T,HW) = X.coef + b,
 * normal equations of the centred data in fp64.  K <= 9.  The ratio itself is then
 * sc_ratio_product(bg = target, sig = recon, zero_value = -0.5) + sc_zero_override(target). */
int64_t sc_mlr_workspace_bytes(int T);
int sc_mlr_reconstruct(const float* bands, const float* target, float* recon, int T, int K, int64_t HW,
                       void* workspace, void* stream);
/* out[i] = value where ref[i] == 0 (np.where(band_target_signal == 0.0, zero_value_out, R), :111) */
int sc_zero_override(const float* ref, float* out, int64_t n, float value, void* stream);

/* ---- A17: EMIT input rescale (starcop/emit_tools/emit_dataset.py:62-101): magic (H,W), rgb (3,H,W)
 * -> out (4, H//32*32, W//32*32) = [clip(mf/240,0,2)*1750, clip(rgb/20,0,2)*60], nan_to_num */
int sc_emit_rescale(const float* magic, const float* rgb, float* out, int H, int W, void* stream);

/* ---- A14: binary opening with the 3x3 cross (starcop/baselines.py:25-27, 54-58) ------------- */
int sc_threshold_opening(const float* pred, float threshold, int64_t* out, uint8_t* scratch,
                         int B, int H, int W, void* stream);

/* ---- (f)3: sensor simulation by spectral response functions (starcop/data/aviris.py:262-338, product at :324-326)
 * over a BIP cube: out[k][p] = sum_c weights[k][c] * cube[p][c] (only where weights != 0), `fill` where any band
 * with a non-zero weight equals `fill` (the reference's missing_values rule).  cube: n_pixels x C f32 (any
 * 4-byte aligned address inside a device allocation), weights: K x C f32 device table (K <= 16; build it with the host recipe of aviris.py:275-316), out:
 * (n_pixels / tile_pixels) x K x tile_pixels planar (the reference's (K,H,W) for each of the stacked scenes).  band_ranges_host: 2K ints on the HOST, [c0, c1) = the bands
 * output band k touches (an SRF is a contiguous window), or NULL = all bands.  One pass: n_pixels * (C + K) * 4
 * algorithmic bytes. */
int sc_srf_aggregate(const float* cube_bip, int64_t n_pixels, int64_t tile_pixels, int C, const float* weights,
                     const int32_t* band_ranges_host, int K, float fill, float* out_planar, void* stream);

/* ---- configs[2] glue (process_aviris.py:183-219 -> dataset -> model_module.py:98): matched-filter output + the three
 * cube bands of the RGB products -> the network's normalised NHWC input in one pass (what sc_normalize_pack makes of
 * the NCHW batch [clip(mf,0,1e4), R, G, B]; integer normaliser factors only, i.e. the HyperSTARCOP products), and
 * optionally the raw NCHW batch (B,4,H,W) and weight_mag1c = clip(mf/400, .1, 1) (feature_extration.py:32-35). */
int sc_chain_pack(const float* mf, const float* cube_bip, int C, int band_r, int band_g, int band_b, const double* off,
                  const double* fac, const double* lo, const double* hi, int B, int64_t HW, void* out_nhwc, int ld,
                  int dtype, float* out_nchw, float* weight_loss, void* stream);

/* ---- (f)1: training augmentation on the device (starcop/data/datamodule.py:128-134, kornia RandomRotation(p=.5,
 * degrees=90) + RandomHorizontalFlip + RandomVerticalFlip applied jointly to input / label / loss weight): the
 * composite is one affine resampling per sample.  in/out: (B,C,H,W) f32, distinct buffers; mats: B x 6 floats mapping
 * OUTPUT pixel (x,y) to the source position (m0 x + m1 y + m2, m3 x + m4 y + m5), pixel centres at integers;
 * bilinear (ATen grid_sampler weights) or nearest (nearbyint), zeros outside the image. */
int sc_affine_warp(const float* in, float* out, const float* mats_dst_to_src, int B, int C, int H, int W,
                   int nearest, void* stream);

/* ---- A7: threshold sweep of run_validation (starcop/validation.py:37-42, 118-125) in one pass.  thresholds: K <= 64
 * floats, ASCENDING (device).  hist: int64 [2][K+1], accumulated: hist[t][i] = pixels with y.long() == t whose
 * prediction is > exactly i thresholds; the confusion matrix of threshold k follows by prefix sums (exact). */
int sc_threshold_sweep(const float* pred, const float* y, const float* thresholds_ascending, int K, int64_t n,
                       int64_t* hist, void* stream);

/* ---- tcgen05 tensor-core path (bf16 storage, fp32 accumulate in TMEM) -----------------------
 * Implicit-GEMM convolution: A = activation tile fetched by TMA (shifted box per filter tap, zero
 * fill = padding), B = packed bf16 weights [Cout][KH*KW][Cin], accumulators in TMEM.
 * Supported: stride 1, KH=KW in {1,3}, Cin % 8 == 0, Cout % 8 == 0, W % 16 == 0, H % 8 == 0
 * (anything else returns SC_ERR_UNSUPPORTED and the caller uses the fp32-FMA kernels).
 * Packed weights: bf16 [cout_pad][tap][cin_pad] with cin_pad = Cin rounded up to the channel chunk
 * (64 if Cin%64==0, 32 if Cin%32==0, else 16) -- sc_tc_cin_pad(Cin).
 * stride 2 (the encoder stem) samples the input through the tensor map's element strides.
 * Optional epilogue: per-channel sum / sum-of-squares of the stored outputs as BatchNorm partial
 * rows (`stats` + *stats_rows_host, see sc_bn_stats); accumulate != 0 adds onto the existing y. */
int sc_tc_cin_pad(int Cin);
int sc_tc_supported(void);
int sc_tc_pack_weights(const float* w_oihw, void* w_bf16, int Cout, int Cin, int KH, int KW,
                       int flip_transpose, int cin_pad, int cout_pad, void* stream);
/* the same re-pack for MANY layers in one launch (the weights change once per optimiser step: the engine
 * re-packs every tensor-core layer's fprop and dgrad filters at the start of a step).  descs_dev: device array
 * of n <= 128 jobs; offset = running sum of the jobs' output element counts (rows * kk * cols), total = their sum. */
typedef struct {
  const float* w;          /* OIHW f32 */
  void* out;               /* bf16 [rows][kk][cols] */
  int64_t offset;
  int32_t cout, cin, kk, flip_transpose, cin_pad, cout_pad;
} sc_tc_pack_desc;
int sc_tc_pack_weights_batch(const sc_tc_pack_desc* descs_dev, int n, int64_t total, void* stream);
int sc_tc_conv_fprop(const void* x, int ldx, const void* w_bf16, void* y, int ldy, double* stats,
                     int* stats_rows_host, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                     int stride, int accumulate, void* stream);
/* dW (OIHW f32) += X^T dY.  The pixel range is split across CTAs; every split stores its partial tile into
 * `partials` (sc_tc_conv_wgrad_workspace_bytes bytes; may be NULL when that is 0) and a second kernel sums the
 * splits in order: no floating-point atomics, bit-identical from run to run. */
int64_t sc_tc_conv_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride);
int sc_tc_conv_wgrad(const void* x, int ldx, const void* dy, int lddy, float* dw_oihw, float* partials,
                     int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, void* stream);

/* Thin-layer weight gradient (3x3, stride 1, pad 1; Cin >= 16, Cout in {16, 32}, H % 8 == 0, W % 16 == 0): the halo
 * patch of a 16x8 pixel tile is fetched once per 32-channel group and the horizontal taps are stacked along the MMA's
 * M dimension (one pixel of descriptor stride per tap), ~4x fewer TMA rows than one shifted box per tap.  One partial
 * gradient per CTA in `partials` (sc_tc_wgrad_halo_workspace_bytes), summed in CTA order: deterministic.
 * sc_tc_conv_wgrad dispatches here by itself when sc_tc_wgrad_halo_supported(...) (its workspace query covers it);
 * STARCOP_NO_WGRAD_HALO=1 disables the route. */
int sc_tc_wgrad_halo_supported(int N, int H, int W, int Cin, int Cout);
int64_t sc_tc_wgrad_halo_workspace_bytes(int N, int H, int W, int Cin, int Cout);
int sc_tc_wgrad_halo(const void* x, int ldx, const void* dy, int lddy, float* dw_oihw, float* partials,
                     int N, int H, int W, int Cin, int Cout, void* stream);

/* Thin-layer variant for 3x3 / stride 1 / pad 1 (decoder blocks 2-4, fprop and dgrad): the 18x10 input halo
 * patch of a 16x8 pixel tile is staged ONCE (9x less L2->SM traffic than one shifted box per tap) and all nine
 * taps' weights stay resident in shared memory.  Packed weights: bf16 [Cout][9][sc_tc_halo_cin_pad(Cin)]
 * (sc_tc_pack_weights with cin_pad = that value).  Any H, W; Cin % 8 == 0; Cout % 16 == 0, Cout <= 128. */
int sc_tc_halo_cin_pad(int Cin);
/* Stem Conv2d(C <= 4 -> Cout, 3x3, stride 2, pad 1) -- the first layer of torchvision MobileNetV2 inside
 * smp.Unet(mobilenet_v2), starcop/models/model_module.py:238-251 -- as a space-to-depth convolution:
 * sc_stem_s2d regroups the bf16 NHWC input (N,H,W,ld) into (N,H/2,W/2,16), sc_stem_s2d_pack_weights writes the
 * equivalent stride-1 3x3 weights over 16 channels in the tensor-core packing [Cout][9][16], sc_stem_s2d_unpack_grad
 * adds the gradient of those weights (Cout,16,3,3 fp32) back into the OIHW gradient (Cout,C,3,3). */
int sc_stem_s2d(const void* x_nhwc, int ldx, int C, void* xs, int N, int H, int W, void* stream);
int sc_stem_s2d_pack_weights(const float* w_oihw, void* w_bf16, int Cout, int C, void* stream);
int sc_stem_s2d_unpack_grad(const float* g_s2d, float* dw_oihw, int Cout, int C, void* stream);
int sc_tc_halo_supported(int Cin, int Cout);
int sc_tc_conv3x3_halo(const void* x, int ldx, const void* w_bf16, void* y, int ldy, double* stats,
                       int* stats_rows_host, int N, int H, int W, int Cin, int Cout, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STARCOP_B200_H */
