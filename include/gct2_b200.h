/* gct2_b200.h -- C ABI of the B200-native training-step kernels for relgukxilef/GAN-Class-Transfer2.
 *
 * The reference (train.py, TensorFlow/Keras) has no FFI of its own: every op below replaces the TensorFlow
 * library op that one line of train.py expands to.  Each entry point names that line.  A maintainer binds
 * these with ctypes; the in-repo binding is gan_class_transfer2_b200/_lib.py.
 *
 * Conventions
 *   - all tensor pointers are DEVICE pointers owned by the caller; nothing is allocated or freed here;
 *   - activations are NHWC bf16 (uint16_t storage), addressed as base + pixel*ld + channel, where `ld` is the
 *     pixel stride in elements: producers write straight into channel slices of the U-Net's concat buffers
 *     (train.py:113-119 tf.concat is never materialised as a copy);
 *   - kernels ("w") are in the Keras variable layouts: Conv2D [4,4,Cin,Cout], Conv2DTranspose [4,4,Cout,Cin];
 *     the tensor-core ops take the bf16 shadow copy maintained by gct2_adam_keras / gct2_cast_bf16;
 *   - gradients, Adam state and master weights are fp32;
 *   - `stream` is a cudaStream_t (CUstream) passed as void*; every call only enqueues work on it (no host
 *     synchronisation), so a whole step can be captured into a CUDA graph;
 *   - return value 0 = success; otherwise gct2_last_error() describes the failure (thread-local string);
 *   - channel counts of tensor-core ops must be multiples of 64; spatial extents powers of two >= 4.
 */
#ifndef GCT2_B200_H
#define GCT2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCT2_ABI_VERSION 2

/* `flags` of the four tensor-core fprop / dgrad entry points.
 * GCT2_WEIGHTS_STABLE: the kernel tensor `w` is not being written by any launch that may still be running when this one
 * starts (the engine's optimiser runs on another stream and is joined by an event), so its first tiles may be fetched
 * before the programmatic dependency on the previous launch of `stream` resolves.  Without the flag every load waits. */
#define GCT2_WEIGHTS_STABLE 1

/* Objective switches of train.py:29-32 as understood by gct2_dense_mse / gct2_sample_update (`target_mode`). */
#define GCT2_TARGET_X 0        /* predict_x = True (the reference's default)                          train.py:243-244 */
#define GCT2_TARGET_EPSILON 1  /* predict_x = False: the network predicts the noise                  train.py:245-246 */
#define GCT2_TARGET_SCALED 2   /* | predict_scaled_epsilon                                            train.py:247-248 */
#define GCT2_TARGET_WEIGHTED 4 /* | prediction_weighting                                              train.py:250-252 */
#define GCT2_TARGET_ODE 8      /* ordinary_differential_equation (takes precedence)                   train.py:238-242 */

int gct2_abi_version(void);
const char* gct2_last_error(void);
/* Selects the device, resolves the driver entry points, raises the kernels' shared-memory limits.
 * Fails (non-zero) when the device is not sm_100. */
int gct2_init(int device);
int gct2_num_sms(void);
/* Number of kernels (and memset nodes of split-K paths) this library has enqueued so far in this process. */
long long gct2_launch_count(void);
/* Test hook (not part of the drop-in surface): key 0/1 override the MN-major UMMA descriptor LBO/SBO bytes,
 * key 2 = verbose plan logging, key 3 = force N tile, key 4 = force split-K, key 7 = record per-CTA phase timestamps
 * (gct2_debug_timeline; only in libraries built with -DGCT2_TIMELINE -- the production main loops carry no stamps),
 * key 8 != 0 = launch without programmatic dependent launch, keys 9 / 10 = CTA budget of the wgrad / dgrad
 * launches (0 = all SMs), key 11 = whole-step launch trace (gct2_debug_trace), key 12 != 0 = finish split-K with a
 * separate kernel instead of inside the launch, key 13 = grid cap of the Adam kernel (0 = 8 blocks per SM), key 15 = grid
 * cap of the down0 weight-gradient kernel (0 = 2 blocks per SM), key 16 = which point of the TMA producer's start-up
 * timeline stamp [7] records, key 19 = CTA pairs (cta_group::2): 0 heuristic, 1 wherever legal, 2 never, key 20 =
 * split-K rendezvous watchdog in polls of ~40 ns (0 = none; default 2^28), key 21 != 0 = never fetch weights before the
 * programmatic dependency resolves, key 22 = gct2_set_sm_budget, key 23 = gct2_set_adam_sms, key 24 = blocks per SM of
 * the Dense+MSE kernel, key 25 = split-K inside a thread-block cluster (partials through distributed shared memory): 0 =
 * when the cost model picks it, 1 = never (the default), 2 = whenever legal, key 26 != 0 = never use the L2 rendezvous form of the
 * in-launch split-K finish (data-parallel steps: it needs all CTAs of a launch resident at once). */
void gct2_debug_set(int key, int value);
/* Test hook: after gct2_debug_set(7, 1) every tensor-core conv launch records, per CTA, 8 %globaltimer stamps (ns):
 * [0] entry, [1] prologue done, [2] first operands landed, [3] MMAs of the first tile issued, [4] first accumulator
 * complete, [5] first epilogue done, [6] CTA done, [7] see key 16.  Synchronises the device and copies the stamps of the most
 * recent launch (up to max_ctas CTAs) to `host`; returns the number of CTAs. */
int gct2_debug_timeline(unsigned long long* host, int max_ctas);
/* Test hook: the plan of the most recent tensor-core conv launch: out8 = {BN, split-K factor, CTA pairs (0/1), split-K
 * finish (0 = none / separate kernel, 1 = in-launch L2 rendezvous, 2 = thread-block cluster through distributed shared
 * memory), grid, ring slots, ring rounds per work item, weights fetched early (0/1)}. */
void gct2_debug_last_plan(int* out8);
/* CTAs (= SMs) a tensor-core conv launch may occupy; 0 = all.  Data-parallel callers leave room for the NCCL kernels
 * that run beside backward, so that a conv launch never queues a second wave behind them. */
void gct2_set_sm_budget(int sms);
/* SMs of the following gct2_adam_apply launches: 0 = the whole chip (many small blocks); n > 0 = n CTAs of 1024 threads,
 * each alone on its SM (it requests most of the SM's shared memory), the form that runs BESIDE the tensor-core launches
 * of backward on disjoint SMs (measured: ~98 GB/s of optimiser traffic per SM up to 48 SMs, 5.8 TB/s from 64).  Both
 * forms compute bit-identical results. */
void gct2_set_adam_sms(int sms);
/* train.py:34,43-45 -- the reference's precision switch (`mixed_precision`, Keras policy 'mixed_float16').  fp16 != 0:
 * every 16-bit tensor of the following launches (activations, their gradients, the weights' shadow copy; the pointers
 * typed uint16_t below) is IEEE fp16 and the tensor cores run kind::f16 on fp16 operands; 0 (default): bf16.  Master
 * weights, gradients of the variables, optimiser state, loss and accumulation stay fp32 in both.  Host-side state, read
 * when a launch is enqueued. */
void gct2_set_policy(int fp16);
int gct2_get_policy(void);
/* train.py:82-83 -- tf.keras.mixed_precision.LossScaleOptimizer (dynamic).  State `ls` = device float[4]: {scale, good
 * steps, all-gradients-finite flag, 1/scale}; initialise to {2^15, 0, 1, 2^-15}.  gct2_dense_mse(loss_scale = ls)
 * multiplies the gradient of the loss by the scale; gct2_loss_scale_check clears the flag when any of g[0..n) is inf or
 * NaN; gct2_adam_apply(loss_scale_state = ls) skips the whole update (and the iteration count) when the flag is clear
 * and otherwise divides the gradients by the scale; gct2_loss_scale_update halves the scale after a skipped step,
 * doubles it after `growth_steps` consecutive good ones (Keras default 2000) and re-arms the flag. */
int gct2_loss_scale_check(const float* g, long long n, float* ls, void* stream);
int gct2_loss_scale_update(float* ls, int growth_steps, void* stream);
/* Test hook: after gct2_debug_set(11, 1) the first and last block of EVERY launch of this library append
 * {kernel id, blockIdx | gridDim << 32, entry ns, exit ns}; this call synchronises, copies up to max_records records
 * (4 x u64 each) to `host`, clears the buffer and returns the count.  Kernel ids: 1 noise, 2 step_begin, 3/4 down0
 * fprop/wgrad, 5 dense+mse, 6 bias grads, 7 adam_prepare, 8 adam, 9 cast, 10 sample update, 20 split-K finish, 21 wgrad reduce,
 * 100 + 10*mode + BN/64 tensor-core conv (mode 0 strided, 1 phase, 2 wgrad). */
int gct2_debug_trace(unsigned long long* host, int max_records);

/* train.py:224-234 + :85-93 -- Trainer.call noising with alpha_dash:
 *   noised = x*sqrt(abar(t)) + eps*sqrt(1-abar(t)), abar(t) = (1 - t/(steps+1))^2 * 0.25.
 * x, eps, noised: fp32 [B, elems_per_image]; t_int: int32 [B]. RNG (t_int, eps) is an input. */
int gct2_noise_images(const float* x, const float* eps, const int32_t* t_int, float* noised, int B,
                      int elems_per_image, int steps, void* stream);

/* train.py:158-169 DownShuffle on the 3-channel image (down0): relu(conv2d(x, w[4,4,3,Cout], s=2, SAME) + b).
 * x fp32 [B,H,W,3]; w, bias fp32; y bf16 [B,H/2,W/2,Cout] with pixel stride ldy. CUDA-core direct conv. */
int gct2_conv4s2_c3_fprop(const float* x, const float* w, const float* bias, uint16_t* y, int ldy, int B, int H,
                          int W, int Cout, void* stream);
/* Backward of the above w.r.t. kernel and bias (Keras train_step, implicit at train.py:516):
 * dz bf16 [B,H/2,W/2,Cout] (already ReLU-masked); dw fp32 [4,4,3,Cout]; db fp32 [Cout] or NULL (bias gradient taken
 * elsewhere).  accumulate == 0: dw, db are overwritten; != 0: added into (the caller zeroed them). */
int gct2_conv4s2_c3_wgrad(const float* x, const uint16_t* dz, int lddz, float* dw, float* db, int B, int H, int W,
                          int Cout, int accumulate, void* stream);

/* train.py:158-169 DownShuffle forward (Cin % 64 == 0): y = relu(conv2d(x, w[4,4,Cin,Cout], s=2, SAME) + b).
 * x bf16 [B,H,W,Cin] stride ldx; y bf16 [B,H/2,W/2,Cout] stride ldy. tcgen05 implicit GEMM (strided form).
 * ws: fp32 split-K scratch (contents irrelevant on entry and exit).  Split-K is used only when `splits` partial
 * outputs (splits * B*(H/2)*(W/2)*Cout floats) fit in ws_bytes; partials are summed in a fixed order, so results are
 * bit-reproducible.  ws may be NULL (no split-K).  The same holds for every ws argument below.
 * Split-K is finished inside the launch: preferably by a thread-block cluster per tile (its CTAs hold the K slices and
 * exchange partials through distributed shared memory; co-scheduled by the hardware, safe beside anything), otherwise,
 * when every work item has its own resident CTA, by a rendezvous at counters of the launch's own over fp32 slabs in
 * `ws` -- that form needs all CTAs of the launch resident at once: do not run two such launches concurrently on
 * different streams of one device unless gct2_set_sm_budget leaves room for both (gct2_debug_set(26, 1) switches it
 * off); wgrad calls never wait and may overlap anything.  flags: GCT2_WEIGHTS_STABLE or 0. */
int gct2_conv4s2_fprop(const uint16_t* x, int ldx, const uint16_t* w, const float* bias, uint16_t* y, int ldy,
                       int B, int H, int W, int Cin, int Cout, float* ws, size_t ws_bytes, int flags, void* stream);
/* Backward-data of DownShuffle: dx[b,iy,ix,ci] (+)= sum dy[b,oy,ox,co]*w[ky,kx,ci,co], then ReLU-masked by the
 * producer's saved output: dx = (acc + (add_old ? dx : 0)) * (act > 0).  dy bf16 [B,H/2,W/2,Cout]; dx, act bf16
 * [B,H,W,Cin].  tcgen05 implicit GEMM (phase form).  ws >= B*H*W*Cin floats. */
int gct2_conv4s2_dgrad(const uint16_t* dy, int lddy, const uint16_t* w, uint16_t* dx, int lddx,
                       const uint16_t* act, int ldact, int add_old, int B, int H, int W, int Cin, int Cout,
                       float* ws, size_t ws_bytes, int flags, void* stream);
/* Backward-filter of DownShuffle: dw[ky,kx,ci,co] = sum x[b,2oy-1+ky,2ox-1+kx,ci]*dy[b,oy,ox,co]; fp32, overwritten.
 * ws: split-K scratch for splits * 16*Cin*Cout floats (must not be shared with a concurrently running call). */
int gct2_conv4s2_wgrad(const uint16_t* x, int ldx, const uint16_t* dy, int lddy, float* dw, int B, int H, int W,
                       int Cin, int Cout, float* ws, size_t ws_bytes, void* stream);

/* train.py:145-156 UpShuffle forward: y = relu(conv2d_transpose(x, w[4,4,Cout,Cin], s=2, SAME) + b).
 * x bf16 [B,H,W,Cin] stride ldx; y bf16 [B,2H,2W,Cout] stride ldy (phase form). ws >= B*2H*2W*Cout floats. */
int gct2_convT4s2_fprop(const uint16_t* x, int ldx, const uint16_t* w, const float* bias, uint16_t* y, int ldy,
                        int B, int H, int W, int Cin, int Cout, float* ws, size_t ws_bytes, int flags, void* stream);
/* Backward-data of UpShuffle: dx[b,iy,ix,ci] = sum dy[b,2iy-1+ky,2ix-1+kx,co]*w[ky,kx,co,ci]; channels
 * [0,mask_channels) are ReLU-masked by act (the saved activation co-located with dx), the rest stored raw
 * (they are the skip-path gradient, consumed by gct2_conv4s2_dgrad(add_old=1)).  Strided form.
 * dy bf16 [B,2H,2W,Cout]; dx, act bf16 [B,H,W,Cin]. ws >= B*H*W*Cin floats. */
int gct2_convT4s2_dgrad(const uint16_t* dy, int lddy, const uint16_t* w, uint16_t* dx, int lddx,
                        const uint16_t* act, int ldact, int mask_channels, int B, int H, int W, int Cin, int Cout,
                        float* ws, size_t ws_bytes, int flags, void* stream);
/* Backward-filter of UpShuffle: dw[ky,kx,co,ci] = sum dy[b,2iy-1+ky,2ix-1+kx,co]*x[b,iy,ix,ci]; fp32, overwritten. */
int gct2_convT4s2_wgrad(const uint16_t* x, int ldx, const uint16_t* dy, int lddy, float* dw, int B, int H, int W,
                        int Cin, int Cout, float* ws, size_t ws_bytes, void* stream);

/* train.py:131-139 Block's Conv2D(filters, 3, 1, 'same', relu) -- the dormant block_depth > 0 branch --
 * on the same tcgen05 implicit-GEMM template as the 4x4 / stride-2 family: a filter tap is a whole-tile shift of one
 * 4-D TMA box (zero fill at the border = SAME padding).  ks is the kernel side: 3, or 1 for a per-pixel projection
 * (the Dense(input_channels) of train.py:106-112 is the 1x1 case).  Cin % 64 == 0, Cout % 64 == 0.
 *   fprop: y = relu(conv2d(x, w[ks,ks,Cin,Cout], s=1, SAME) + b); x, y 16-bit [B,H,W,*] with pixel strides ldx / ldy.
 *   dgrad: dx = mask(conv2d_backprop_input(dy, w) (+ dx when add_old)): columns [0, mask_channels) are zeroed where the
 *          saved activation `act` (co-located with dx, stride ldact) is not positive, the rest is stored raw (the skip
 *          slice of a concat buffer, completed later by the DownShuffle dgrad that also reads it).
 *   wgrad: dw[ks,ks,Cin,Cout] fp32 = conv2d_backprop_filter(x, dy); overwritten.  One side's channel count must be a
 *          multiple of 128.
 * ws / flags as for gct2_conv4s2_fprop. */
int gct2_conv3s1_fprop(const uint16_t* x, int ldx, const uint16_t* w, const float* bias, uint16_t* y, int ldy, int B,
                       int H, int W, int Cin, int Cout, int ks, float* ws, size_t ws_bytes, int flags, void* stream);
/* train.py:106-112 Residual with residual = True: y = res + Dense(Cout, use_bias=False)(x), i.e. the ks = 1 case of the
 * stride-1 map without bias and activation and with the layer's input `res` (16-bit [B,H,W,Cout], stride ldres) added in
 * the epilogue.  Backward: gct2_conv3s1_dgrad / _wgrad with ks = 1 (the identity path's gradient is the incoming one). */
int gct2_conv3s1_fprop_add(const uint16_t* x, int ldx, const uint16_t* w, const uint16_t* res, int ldres, uint16_t* y,
                           int ldy, int B, int H, int W, int Cin, int Cout, int ks, int flags, void* stream);
int gct2_conv3s1_dgrad(const uint16_t* dy, int lddy, const uint16_t* w, uint16_t* dx, int lddx, const uint16_t* act,
                       int ldact, int mask_channels, int add_old, int B, int H, int W, int Cin, int Cout, int ks,
                       float* ws, size_t ws_bytes, int flags, void* stream);
int gct2_conv3s1_wgrad(const uint16_t* x, int ldx, const uint16_t* dy, int lddy, float* dw, int B, int H, int W,
                       int Cin, int Cout, int ks, float* ws, size_t ws_bytes, void* stream);
/* The first convolution of the outermost Block (train.py:192 with block_depth > 0) reads the 3-channel image:
 * y = relu(conv2d(x, w[3,3,3,Cout], s=1, SAME) + b), x fp32 [B,H,W,3], y 16-bit stride ldy; and its weight gradient
 * dw fp32 [3,3,3,Cout] from dz (16-bit, already ReLU-masked).  CUDA-core direct convolutions (K = 27).  Cout % 8 == 0.
 * accumulate == 0: dw is overwritten; != 0: added into (the caller zeroed it). */
int gct2_conv3s1_c3_fprop(const float* x, const float* w, const float* bias, uint16_t* y, int ldy, int B, int H,
                          int W, int Cout, void* stream);
int gct2_conv3s1_c3_wgrad(const float* x, const uint16_t* dz, int lddz, float* dw, int B, int H, int W, int Cout,
                          int accumulate, void* stream);

/* BiasAddGrad of every conv layer: db[c] = sum over rows of dz[row*ld + c]; dz bf16, db fp32 (overwritten). */
int gct2_bias_grad(const uint16_t* dz, int ld, long long rows, int C, float* db, void* stream);
/* The same for n <= 16 tensors in ONE launch (all conv layers of a step).  dz, ld, rows, C, db are HOST arrays of
 * length n holding device pointers / sizes.  accumulate as above. */
int gct2_bias_grad_multi(int n, const uint16_t* const* dz, const int* ld, const long long* rows, const int* C,
                         float* const* db, int accumulate, void* stream);

/* train.py:198-202 Dense(3) on concat([up0_out(64), noised(3)]) fused with train.py:262-272 MSE and their
 * backward.  u0 bf16 [pixels,64] stride ldu; noised, x fp32 [pixels,3]; wd fp32 [67,3]; bd fp32 [3].
 * pred (nullable) fp32 [pixels,3]; loss: one fp32, overwritten with sum((pred-x)^2)*inv_n; inv_n = 1/(global
 * element count) so data-parallel ranks produce partial means.  When backward != 0 also writes
 * du0 = (u0>0) * (dpred . wd^T) (bf16, stride lddu), dwd fp32 [Cu+3,3], dbd fp32 [3], dpred = 2(pred-x)*inv_n.
 * Cu is 64 or 128.  accumulate == 0: loss, dwd, dbd are overwritten; != 0: added into (the caller zeroed them).
 * noised == NULL: the layer reads the Cu 16-bit channels only (wd, dwd fp32 [Cu,3]) -- Dense(3) behind a Block
 * (block_depth > 0) or without the concat skip (concat = False), where no image channels reach it.
 * loss_scale: NULL, or the device state of gct2_loss_scale_* (dpred is multiplied by loss_scale[0]; the loss is not).
 * target_mode selects what the loss compares (train.py:238-252): 0 = predict_x (target = x; eps, t_int unused, may be
 * NULL); otherwise the network predicts the noise -- GCT2_TARGET_EPSILON, optionally | GCT2_TARGET_SCALED (target
 * eps*sqrt(1-abar(t))) | GCT2_TARGET_WEIGHTED (target and prediction both times sqrt(1-abar(t))) -- or
 * GCT2_TARGET_ODE (target = the noised image of step t-1).  eps fp32 [pixels,3] and t_int int32 [images] are the draws
 * of the noising step; pixels_per_image maps a pixel to its image. */
int gct2_dense_mse(const uint16_t* u0, int ldu, const float* noised, const float* x, const float* wd,
                   const float* bd, float* pred, float* loss, uint16_t* du0, int lddu, float* dwd, float* dbd,
                   long long pixels, int Cu, float inv_n, int backward, int accumulate, const float* loss_scale,
                   const float* eps, const int32_t* t_int, long long pixels_per_image, int target_mode, int steps,
                   void* stream);

/* train.py:106-112 with residual = True at the image level (block_depth = 0): the outermost Residual returns
 * r = noised + up0 . wp (Dense(3, use_bias=False), wp fp32 [U,3]) and Dense(3) (wd fp32 [3,3], train.py:198-202) follows, so
 * pred = up0 . (wp wd) + noised . wd + bd: gct2_dense_mse with the effective kernel weff = [wp wd ; wd] (fp32 [U+3,3]) that
 * gct2_res0_compose writes.  gct2_res0_decompose maps gct2_dense_mse's dwd output (dweff, fp32 [U+3,3]) back:
 * dwp = dweff[:U] . wd^T, dwd = wp^T . dweff[:U] + dweff[U:] (both overwritten). */
int gct2_res0_compose(const float* wp, const float* wd, float* weff, int U, void* stream);
int gct2_res0_decompose(const float* dweff, const float* wp, const float* wd, float* dwp, float* dwd, int U, void* stream);

/* train.py:50-65,75 -- tf.keras.optimizers.Adam(WarmUp(base_lr, warmup_steps)), Keras formula (epsilon added to
 * the un-bias-corrected sqrt(v)).  All n parameters live in flat fp32 buffers; w_bf16 receives the shadow copy.
 * iterations: device int64 (0-based step count, incremented here); hyper: device float[2] scratch
 * (alpha, lr of this step).  g is multiplied by grad_scale first (1 for summed data-parallel gradients). */
int gct2_adam_keras(float* w, float* m, float* v, const float* g, uint16_t* w_bf16, long long n,
                    long long* iterations, float* hyper, float base_lr, int warmup_steps, float beta1, float beta2,
                    float eps, float grad_scale, void* stream);
/* The same optimiser in two parts, so that the update of a contiguous range of variables can start as soon as that
 * range's gradients are complete (overlapping the rest of backward): gct2_adam_prepare once per step (computes
 * alpha/lr of this step into hyper[0..1], increments *iterations), then gct2_adam_apply per range
 * (iterations_inc: NULL, or a counter to increment by one -- used together with gct2_step_begin; loss_scale_state: NULL,
 * or the device state of gct2_loss_scale_*). */
int gct2_adam_prepare(long long* iterations, float* hyper, float base_lr, int warmup_steps, float beta1, float beta2,
                      void* stream);
int gct2_adam_apply(float* w, float* m, float* v, const float* g, uint16_t* w_bf16, long long n, const float* hyper,
                    float beta1, float beta2, float eps, float grad_scale, long long* iterations_inc,
                    const float* loss_scale_state, void* stream);
/* The same update with the gradient given as bf16 (data parallel: the gradient reduce-scatter runs on a bf16 copy of
 * the bucket -- half the NVLink bytes -- and the optimiser of this rank's slice reads the summed bf16 values). */
int gct2_adam_apply_g16(float* w, float* m, float* v, const uint16_t* g_bf16, uint16_t* w_bf16, long long n,
                        const float* hyper, float beta1, float beta2, float eps, float grad_scale,
                        long long* iterations_inc, void* stream);
/* Data parallel, the fused form: gradient exchange + Keras-Adam + weight broadcast of one rank's slice of a
 * gradient bucket in ONE kernel over NVLink peer memory -- no collective library on the path.  g_bf16_ptrs / w16_ptrs are
 * HOST arrays of `world` device pointers: the BASE of every rank's bf16 gradient buffer and of every rank's 16-bit weight
 * shadow, all mapped into this process (CUDA IPC / a symmetric-memory allocator; index = rank, this rank included).  The slice
 * is elements [elem_offset, elem_offset + n) of those flat buffers; w, m, v point at THIS rank's fp32 masters of the
 * slice.  For every element the kernel sums the bf16 gradients of all ranks in fp32 (rank order), applies the update,
 * and writes the new 16-bit weight into every rank's shadow (write_all != 0) or only into w16_ptrs[0] (write_all == 0:
 * a range every rank updates redundantly).  g_multicast / w16_multicast: NULL, or the NVLS multicast addresses of the two
 * buffers -- then the sum is one multimem.ld_reduce (added inside the NVSwitch) and the broadcast one multimem.st.
 * The caller orders it across ranks: all ranks' gradients of the slice complete before, all writes landed after. */
int gct2_adam_apply_p2p(float* w, float* m, float* v, const uint16_t* const* g_bf16_ptrs, uint16_t* const* w16_ptrs,
                        const uint16_t* g_multicast, uint16_t* w16_multicast, int world, long long elem_offset, long long n,
                        const float* hyper, float beta1, float beta2, float eps, float grad_scale, int write_all,
                        void* stream);
/* Data parallel: the small replicated region (down0's kernel, biases, Dense) and the scalar loss summed over all ranks by
 * peer loads -- out_a[i] = sum_r src_ptrs[r][i] for i < n_a, out_b[i - n_a] likewise for the next n_b elements, ranks added in
 * index order (the same bits on every rank).  src_ptrs: HOST array of `world` device pointers to every rank's fp32 staging
 * buffer (symmetric memory, this rank included).  The caller orders it across ranks like gct2_adam_apply_p2p. */
int gct2_sum_peers_f32(const float* const* src_ptrs, int world, float* out_a, long long n_a, float* out_b, long long n_b,
                       void* stream);
/* Everything a step needs before its first convolution, in one launch (train.py:224-234 plus optimiser bookkeeping):
 * draws t_int ~ U{1..steps} per image and eps ~ N(0,1) per element on the device (Philox4x32-10 keyed by `seed`, offset by
 * *iterations so every step differs; the reference draws with TF's unseeded generators), writes
 * noised = x*sqrt(abar(t)) + eps*sqrt(1-abar(t)); optionally stores eps / t_int (eps_out, t_out may be NULL); zeroes
 * gsmall[0..nsmall) and *loss; writes this step's Adam alpha/lr to hyper[0..1] WITHOUT incrementing *iterations (pass
 * iterations_inc to the step's last gct2_adam_apply instead). */
int gct2_step_begin(const float* x, float* noised, float* eps_out, int32_t* t_out, int B, int elems_per_image, int steps,
                    unsigned long long seed, const long long* iterations, float* hyper, float base_lr, int warmup_steps,
                    float beta1, float beta2, float* gsmall, long long nsmall, float* loss, void* stream);
/* The same with the batch as it leaves the reference's decode_file (train.py:285-293), i.e. one step EARLIER than x:
 * img uint8 [B,H,W,3]; x = img/128 - 1, image b mirrored left-right when flip != NULL and flip[b] != 0
 * (tf.image.random_flip_left_right; the draw is the caller's, like the crop).  x_out fp32 [B,H,W,3] receives the decoded
 * image (the loss target); everything else as gct2_step_begin.  A quarter of the host-to-device bytes per step. */
int gct2_step_begin_u8(const uint8_t* img, const uint8_t* flip, float* x_out, int width, float* noised, float* eps_out,
                       int32_t* t_out, int B, int elems_per_image, int steps, unsigned long long seed,
                       const long long* iterations, float* hyper, float base_lr, int warmup_steps, float beta1,
                       float beta2, float* gsmall, long long nsmall, float* loss, void* stream);

/* train.py:365-398 and :441-468 -- log_sample's diffusion loops (predict_x branch), the arithmetic between two
 * Denoiser calls in one launch.  After the call at step t:  x_theta = pred;  eps_theta = (fake - sqrt(abar(t)) x_theta)
 * / sqrt(1 - abar(t));  then, if 1 <= t_next <= steps, the next call's input  fake = sqrt(abar(t_next)) x_theta +
 * sqrt(1 - abar(t_next)) eps_theta  (in place).  pred == NULL: only the mix, from the given x_theta / eps_theta (the
 * loop's first iteration).  All tensors fp32 with n elements.  target_mode as for gct2_dense_mse: with GCT2_TARGET_EPSILON
 * (| _SCALED) the prediction is the (scaled) noise and x_theta = (fake - scaled noise) / sqrt(abar(t)) (train.py:400-413);
 * with GCT2_TARGET_ODE x_theta follows train.py:382-391 and eps_theta is left unchanged. */
int gct2_sample_update(const float* pred, float* fake, float* x_theta, float* eps_theta, int t, int t_next, int steps,
                       long long n, int target_mode, void* stream);
/* train.py:418-432 -- the latent edits between log_sample's two loops, one launch: from epsilon_theta fp32 [size,size,3]
 * and dictionary fp32 [size,size,entries,3] writes out fp32 [4,size,size,3] = {epsilon_theta, pixelated (avg_pool2d 4 +
 * nearest UpSampling2D 4), shifted (tf.roll by 1 along both axes), quantised (nearest dictionary entry per pixel)}. */
int gct2_latent_edits(const float* eps_theta, const float* dictionary, float* out, int size, int entries, void* stream);
/* train.py:357-361 -- log_sample's 'example loss': out[0] = sqrt(mean((a - b)^2)) over n fp32 elements. */
int gct2_rmse(const float* a, const float* b, long long n, float* out, void* stream);
/* fp32 -> bf16 (round to nearest even); builds the first shadow copy of the weights. */
int gct2_cast_bf16(const float* src, uint16_t* dst, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCT2_B200_H */
