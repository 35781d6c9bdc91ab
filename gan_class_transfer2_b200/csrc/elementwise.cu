// HBM-bound kernels of the training step: noising, the 3-channel first conv (down0) and its weight
// gradient, the fused Dense(3)+MSE forward/backward, per-channel bias gradients, and Keras-Adam.
// Reference semantics: train.py:85-93 (alpha_dash), :224-234 (noising), :158-169 (DownShuffle on the
// 3-channel image), :198-202 (Dense(3)), :262-272 (MSE), :50-65,75 (WarmUp + Adam, Keras formula).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "conv_host.cuh"
#include "elementwise.cuh"
#include "ptx.cuh"

namespace gct2 {

void trace_set_elementwise(unsigned long long* buf) { cudaMemcpyToSymbol(g_trace_buf, &buf, sizeof(buf)); }

static int g_ew_sms = 148;
static int g_c3w_blocks = 0;   // debug key 15: grid cap of down0's weight-gradient kernel (0 = 2 blocks per SM)
static int g_adam_blocks = 0;  // debug key 13: grid cap of the Adam kernel (0 = 8 blocks per SM)
static int g_adam_sms = 0;     // gct2_set_adam_sms / debug key 23: > 0 = run Keras-Adam on that many SMs, one 1024-thread
                               // CTA each (SM-exclusive through its shared-memory request), leaving the rest to the convs
static int g_f16 = 0;          // gct2_set_policy: 16-bit storage format of activations / gradients / weight shadow (0 bf16, 1 fp16)
void elementwise_set_f16(int f16) { g_f16 = f16 ? 1 : 0; }
static int g_dense_bps = 2;    // debug key 24: blocks per SM of the fused Dense+MSE kernel
void elementwise_set_debug(int key, int value) {
  if (key == 24) g_dense_bps = value > 0 ? value : 2;
  if (key == 13) g_adam_blocks = value;
  if (key == 15) g_c3w_blocks = value;
  if (key == 23) g_adam_sms = value > 0 ? value : 0;
}
void elementwise_set_sms(int n) { g_ew_sms = n; }
void elementwise_set_adam_sms(int n) { g_adam_sms = n > 0 ? n : 0; }

#define GCT2_CHECK_LAUNCH(name)                                       \
  do {                                                                \
    cudaError_t e__ = cudaGetLastError();                             \
    if (e__ != cudaSuccess) {                                         \
      set_error("%s launch: %s", name, cudaGetErrorString(e__));      \
      return 1;                                                       \
    }                                                                 \
    count_launch();                                                   \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------ noising (a1)
// noised = x*sqrt(abar(t)) + eps*sqrt(1-abar(t)),  abar(t) = (1 - t/(steps+1))^2 * 0.25   (train.py:85-93,231-234)
__global__ void noise_kernel(const float4* __restrict__ x, const float4* __restrict__ eps,
                             const int* __restrict__ t_int, float4* __restrict__ out, int B, int vecPerImage,
                             int steps) {
  TraceScope trace(1);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  const long long total = (long long)B * vecPerImage;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / vecPerImage);
    float t = (float)__ldg(t_int + b);
    t = t / (float)(steps + 1);
    const float om = 1.f - t;
    const float abar = om * om * 0.25f;
    const float sa = sqrtf(abar), sb = sqrtf(1.f - abar);
    const float4 xv = __ldg(x + i), ev = __ldg(eps + i);
    out[i] = make_float4(xv.x * sa + ev.x * sb, xv.y * sa + ev.y * sb, xv.z * sa + ev.z * sb, xv.w * sa + ev.w * sb);
  }
  trace.end();
}

int noise_images(const float* x, const float* eps, const int* t_int, float* noised, int B, int elemsPerImage,
                 int steps, cudaStream_t st) {
  if (elemsPerImage % 4) {
    set_error("noise_images: elements per image must be a multiple of 4");
    return 1;
  }
  const int vec = elemsPerImage / 4;
  const long long total = (long long)B * vec;
  int blocks = (int)((total + 255) / 256);
  if (blocks > g_ew_sms * 8) blocks = g_ew_sms * 8;
  if (blocks < 1) blocks = 1;
  launch_k(noise_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(eps), t_int,
                                      reinterpret_cast<float4*>(noised), B, vec, steps);
  GCT2_CHECK_LAUNCH("noise_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ sampling-loop update (f1)
// train.py:365-398 / :441-468 (predict_x branch), everything between two Denoiser calls of log_sample in one launch:
//   x_theta = prediction ; eps_theta = (fake - sqrt(abar_t) x_theta) / sqrt(1 - abar_t)            (after the call at t)
//   fake'   = sqrt(abar_t') x_theta + sqrt(1 - abar_t') eps_theta                                  (input of the call at t')
// pred == nullptr: only the second line, from the given x_theta / eps_theta (the loop's first mix).  t_next outside
// [1, steps]: only the first line (the loop's last update).  fake is read (at t) and overwritten (for t_next) in place.
__device__ __forceinline__ float abar_of(int t, int steps) {
  const float om = 1.f - (float)t / (float)(steps + 1);
  return om * om * 0.25f;
}
__device__ __forceinline__ float abar_of_f(float t, int steps) {
  const float om = 1.f - t / (float)(steps + 1);
  return om * om * 0.25f;
}
__global__ void sample_update_kernel(const float4* __restrict__ pred, float4* __restrict__ fake, float4* __restrict__ x_theta,
                                     float4* __restrict__ eps_theta, int t, int t_next, int steps, long long nvec,
                                     int mode) {
  TraceScope trace(10);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  const float a = abar_of(t, steps), sa = sqrtf(a), sb = sqrtf(1.f - a);
  const bool mix = t_next >= 1 && t_next <= steps;
  const float an = abar_of(mix ? t_next : t, steps), san = sqrtf(an), sbn = sqrtf(1.f - an);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float4 xt, et;
    if (pred != nullptr) {
      const float4 pr = __ldg(pred + i);
      const float4 f = fake[i];
      if (mode & 8) {
        // ordinary_differential_equation (train.py:382-391, 452-461): x_theta from the predicted previous step;
        // epsilon_theta is left untouched by the reference
        const float a1 = abar_of(t - 1, steps), sa1 = sqrtf(a1), sb1 = sqrtf(1.f - a1);
        const float den = sa1 * sb - sa * sb1;
        xt = make_float4((pr.x * sb - f.x * sb1) / den, (pr.y * sb - f.y * sb1) / den, (pr.z * sb - f.z * sb1) / den,
                         (pr.w * sb - f.w * sb1) / den);
        et = eps_theta[i];
      } else if (mode == 0) {  // predict_x (train.py:394-398, 464-468)
        xt = pr;
        et = make_float4((f.x - sa * xt.x) / sb, (f.y - sa * xt.y) / sb, (f.z - sa * xt.z) / sb, (f.w - sa * xt.w) / sb);
      } else {  // the network predicts (scaled) epsilon (train.py:400-413, 470-479)
        float4 sc4;
        if (mode & 2) {
          et = make_float4(pr.x / sb, pr.y / sb, pr.z / sb, pr.w / sb);
          sc4 = pr;
        } else {
          et = pr;
          sc4 = make_float4(pr.x * sb, pr.y * sb, pr.z * sb, pr.w * sb);
        }
        xt = make_float4((f.x - sc4.x) / sa, (f.y - sc4.y) / sa, (f.z - sc4.z) / sa, (f.w - sc4.w) / sa);
      }
      x_theta[i] = xt;
      eps_theta[i] = et;
    } else {
      xt = x_theta[i];
      et = eps_theta[i];
    }
    if (mix)
      fake[i] = make_float4(san * xt.x + sbn * et.x, san * xt.y + sbn * et.y, san * xt.z + sbn * et.z,
                            san * xt.w + sbn * et.w);
  }
  trace.end();
}

int sample_update(const float* pred, float* fake, float* x_theta, float* eps_theta, int t, int t_next, int steps,
                  long long n, int mode, cudaStream_t st) {
  if (n % 4 || steps < 1 || (pred != nullptr && (t < 1 || t > steps))) {
    set_error("sample_update: n must be a multiple of 4 and 1 <= t <= steps (n=%lld t=%d steps=%d)", n, t, steps);
    return 1;
  }
  const long long nvec = n / 4;
  int blocks = (int)((nvec + 255) / 256);
  if (blocks > g_ew_sms * 8) blocks = g_ew_sms * 8;
  if (blocks < 1) blocks = 1;
  launch_k(sample_update_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const float4*>(pred),
           reinterpret_cast<float4*>(fake), reinterpret_cast<float4*>(x_theta), reinterpret_cast<float4*>(eps_theta), t,
           t_next, steps, nvec, mode);
  GCT2_CHECK_LAUNCH("sample_update_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ fused step prologue
// One launch for everything the step needs before the first convolution (train.py:224-234 + optimiser bookkeeping):
//   t_int ~ U{1..steps} per image, eps ~ N(0,1) per element (Philox4x32-10 keyed by `seed`, offset by the optimiser
//   iteration so every step draws fresh numbers; Box-Muller), noised = x*sqrt(abar) + eps*sqrt(1-abar);
//   zeroes the atomically-accumulated gradient region and the loss; computes this step's Adam alpha/lr.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
__device__ __forceinline__ float u32_to_unit_open(uint32_t v) {  // (0, 1]
  return ((float)(v >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

// U8: the batch arrives as the decoded JPEG bytes (train.py:285-293 decode_file): x = u8 / 128 - 1, optionally mirrored
// left-right per image (tf.image.random_flip_left_right, the draw stays with the caller); the fp32 image is written
// for the loss and never read back by this kernel.
template <bool U8>
__global__ void __launch_bounds__(256) step_begin_kernel(const float4* __restrict__ x, const uint8_t* __restrict__ x_u8,
                                                         const uint8_t* __restrict__ flip, float4* __restrict__ x_out,
                                                         int W, float4* __restrict__ noised,
                                                         float4* __restrict__ eps_out, int* __restrict__ t_out,
                                                         int B, int vecPerImage, int steps, unsigned long long seed,
                                                         const long long* __restrict__ iterations,
                                                         float* __restrict__ hyper, float base, int warmup, float b1,
                                                         float b2, float4* __restrict__ gsmall, long long nsmallVec,
                                                         float* __restrict__ loss) {
  TraceScope trace(2);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  const long long step = *iterations;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  if (tid == 0) {
    float lr = base;
    if (step < warmup) lr = base * (float)(step + 1) / (float)(warmup + 1);
    const float t = (float)(step + 1);
    hyper[0] = lr * sqrtf(1.f - powf(b2, t)) / (1.f - powf(b1, t));
    hyper[1] = lr;
    *loss = 0.f;
  }
  for (long long i = tid; i < nsmallVec; i += nthreads) gsmall[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long total = (long long)B * vecPerImage;
  for (long long i = tid; i < total; i += nthreads) {
    const int b = (int)(i / vecPerImage);
    // stream 1: one draw per image; stream 0: four normals per float4
    const uint4 ti = philox4x32_10(make_uint4((uint32_t)b, 0u, (uint32_t)step, 0x80000000u | (uint32_t)(step >> 32)), key);
    const int tint = 1 + (int)(ti.x % (uint32_t)steps);
    if (t_out != nullptr && i % vecPerImage == 0) t_out[b] = tint;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)step, (uint32_t)(step >> 32) & 0x7fffffffu),
                                  key);
    const float ra = sqrtf(-2.f * __logf(u32_to_unit_open(r.x))), rb = sqrtf(-2.f * __logf(u32_to_unit_open(r.z)));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u32_to_unit_open(r.y), &s0, &c0);
    __sincosf(6.283185307179586f * u32_to_unit_open(r.w), &s1, &c1);
    const float4 ev = make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
    float t = (float)tint / (float)(steps + 1);
    const float om = 1.f - t;
    const float abar = om * om * 0.25f;
    const float sa = sqrtf(abar), sb = sqrtf(1.f - abar);
    float4 xv;
    if (U8) {
      const long long e0 = (i - (long long)b * vecPerImage) * 4;  // first of my four scalars inside image b (HWC, C = 3)
      const uint8_t* img = x_u8 + (long long)b * vecPerImage * 4;
      const bool mirror = flip != nullptr && flip[b] != 0;
      float f[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        long long e = e0 + j;
        if (mirror) {
          const long long pix = e / 3;
          const int c = (int)(e - pix * 3);
          const long long row = pix / W;
          const int col = (int)(pix - row * W);
          e = (row * W + (W - 1 - col)) * 3 + c;
        }
        f[j] = (float)__ldg(img + e) * (1.0f / 128.0f) - 1.0f;
      }
      xv = make_float4(f[0], f[1], f[2], f[3]);
      x_out[i] = xv;
    } else {
      xv = __ldg(x + i);
    }
    if (eps_out != nullptr) eps_out[i] = ev;
    noised[i] = make_float4(xv.x * sa + ev.x * sb, xv.y * sa + ev.y * sb, xv.z * sa + ev.z * sb, xv.w * sa + ev.w * sb);
  }
  trace.end();
}

int step_begin(const float* x, const uint8_t* x_u8, const uint8_t* flip, float* x_out, int W, float* noised,
               float* eps_out, int* t_out, int B, int elemsPerImage, int steps,
               unsigned long long seed, const long long* iterations, float* hyper, float base_lr, int warmup_steps,
               float beta1, float beta2, float* gsmall, long long nsmall, float* loss, cudaStream_t st) {
  if (elemsPerImage % 4 || nsmall % 4) {
    set_error("step_begin: elements per image and the small gradient region must be multiples of 4");
    return 1;
  }
  const int vec = elemsPerImage / 4;
  const long long total = (long long)B * vec;
  int blocks = (int)((total + 255) / 256);
  if (blocks > g_ew_sms * 8) blocks = g_ew_sms * 8;
  if (blocks < 1) blocks = 1;
  if (x_u8 != nullptr) {
    if (x_out == nullptr || W < 1 || elemsPerImage % (3 * W)) {
      set_error("step_begin: the uint8 input needs x_out and a width that divides the image (W=%d)", W);
      return 1;
    }
    launch_k(step_begin_kernel<true>, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const float4*>(x), x_u8, flip,
             reinterpret_cast<float4*>(x_out), W, reinterpret_cast<float4*>(noised), reinterpret_cast<float4*>(eps_out),
             t_out, B, vec, steps, seed, iterations, hyper, base_lr, warmup_steps, beta1, beta2,
             reinterpret_cast<float4*>(gsmall), nsmall / 4, loss);
    GCT2_CHECK_LAUNCH("step_begin_kernel<u8>");
    return 0;
  }
  launch_k(step_begin_kernel<false>, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const float4*>(x), x_u8, flip,
           reinterpret_cast<float4*>(x_out), W, reinterpret_cast<float4*>(noised),
                                           reinterpret_cast<float4*>(eps_out), t_out, B, vec, steps, seed, iterations,
                                           hyper, base_lr, warmup_steps, beta1, beta2,
                                           reinterpret_cast<float4*>(gsmall), nsmall / 4, loss);
  GCT2_CHECK_LAUNCH("step_begin_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ down0 (Cin = 3)
// Direct conv on CUDA cores: K = 48 is too thin for a tensor-core tile and the layer is bound by its 128-channel
// output write.  The 48 taps of one output value are consumed as 24 packed pairs: one 8-byte shared-memory load brings
// two neighbouring patch values, one fma.rn.f32x2 (FFMA2, two fp32 FMAs per instruction on sm_100) multiplies them with
// the matching pair of weights into an (even-tap, odd-tap) accumulator pair that is summed at the end.  A thread owns
// TWO output channels, so every patch load feeds four FMAs: the loop is bound by the FMA pipe, not by shared memory.
constexpr int C3_T = 8;                 // output tile width (and height of the wgrad tile)
constexpr int C3_P = 2 * C3_T + 2;      // input patch width (18)
constexpr int C3_ROW = C3_P * 3 + 2;    // padded patch row (56 floats, 16-byte aligned rows)

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {  // c + a * b, both lanes
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// rows [oy0, oy0 + TY) x columns [ox0, ox0 + 8) of the output need input rows 2*oy0-1 .. 2*oy0+2*TY, columns 2*ox0-1 ..
template <int TY>
__device__ __forceinline__ void c3_load_patch(float (*patch)[C3_ROW], const float* __restrict__ x, int b, int oy0,
                                               int ox0, int H, int W) {
  constexpr int PY = 2 * TY + 2;
  for (int i = threadIdx.x; i < PY * C3_P * 3; i += blockDim.x) {
    const int c = i % 3, xx = (i / 3) % C3_P, yy = i / (3 * C3_P);
    const int iy = 2 * oy0 - 1 + yy, ix = 2 * ox0 - 1 + xx;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + (((long long)b * H + iy) * W + ix) * 3 + c);
    patch[yy][xx * 3 + c] = v;
  }
}

// fprop: block = 128 threads = 64 channel pairs x 2 row groups; tile = 8 x 4 output pixels (two rows per group).
constexpr int C3F_TY = 4;
__global__ void __launch_bounds__(128) conv_c3_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias,
                                                            __nv_bfloat16* __restrict__ y, int ldy, int B, int H,
                                                            int W, int Cout, int f16) {
  TraceScope trace(3);
  __shared__ __align__(16) float patch[2 * C3F_TY + 2][C3_ROW];
  const int Ho = H / 2, Wo = W / 2;
  const int tilesX = Wo / C3_T, tilesY = Ho / C3F_TY;
  const int tile = blockIdx.x;
  const int b = tile / (tilesX * tilesY);
  const int oy0 = ((tile / tilesX) % tilesY) * C3F_TY, ox0 = (tile % tilesX) * C3_T;
  const int cp = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int co = blockIdx.y * 128 + 2 * cp;
  // the kernel and the bias are not produced by the previous launch: fetch them before the dependency resolves
  float2 w2[2][24];
#pragma unroll
  for (int j = 0; j < 24; ++j) {  // HWIO: ((ky*4+kx)*3+c)*Cout + co ; pair j = taps (2j, 2j+1)
    const float2 lo = __ldg(reinterpret_cast<const float2*>(w + (2 * j) * Cout + co));
    const float2 hi = __ldg(reinterpret_cast<const float2*>(w + (2 * j + 1) * Cout + co));
    w2[0][j] = make_float2(lo.x, hi.x);
    w2[1][j] = make_float2(lo.y, hi.y);
  }
  const float2 bv = __ldg(reinterpret_cast<const float2*>(bias + co));
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  c3_load_patch<C3F_TY>(patch, x, b, oy0, ox0, H, W);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < C3F_TY / 2; ++r) {
    const int py = grp * (C3F_TY / 2) + r;
#pragma unroll 2
    for (int px = 0; px < C3_T; ++px) {
      float2 a0 = make_float2(bv.x, 0.f), a1 = make_float2(bv.y, 0.f);
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        // the 12 floats of this tap row start 24*px bytes into a 16-byte aligned row: six 8-byte broadcast loads
        const float2* row = reinterpret_cast<const float2*>(&patch[2 * py + ky][2 * px * 3]);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const float2 xv = row[j];
          a0 = ffma2(xv, w2[0][ky * 6 + j], a0);
          a1 = ffma2(xv, w2[1][ky * 6 + j], a1);
        }
      }
      const long long pix = ((long long)b * Ho + oy0 + py) * Wo + ox0 + px;
      *reinterpret_cast<uint32_t*>(y + pix * ldy + co) = pack_h2(fmaxf(a0.x + a0.y, 0.f), fmaxf(a1.x + a1.y, 0.f), f16);
    }
  }
  trace.end();
}

int conv4s2_c3_fprop(const float* x, const float* w, const float* bias, __nv_bfloat16* y, int ldy, int B, int H,
                     int W, int Cout, cudaStream_t st) {
  if ((H / 2) % C3_T || (W / 2) % C3_T || Cout % 128 || ldy % 2) {
    set_error("conv4s2_c3_fprop: unsupported shape H=%d W=%d Cout=%d ldy=%d", H, W, Cout, ldy);
    return 1;
  }
  dim3 grid(B * (H / 2 / C3F_TY) * (W / 2 / C3_T), Cout / 128);
  launch_k(conv_c3_fprop_kernel, dim3(grid), dim3(128), 0, st, x, w, bias, y, ldy, B, H, W, Cout, g_f16);
  GCT2_CHECK_LAUNCH("conv_c3_fprop_kernel");
  return 0;
}

// dW[ky,kx,c,co] = sum_pix x[pix@tap, c] * dz[pix, co];  db[co] = sum_pix dz[pix, co]
// 256 threads = 64 channel pairs x 4 row groups of an 8 x 8 tile; the thread's 2 x 48 sums live in 48 packed register
// pairs (FFMA2 with the gradient broadcast to both lanes).  The four row groups are combined with shared-memory atomics
// (conflict-free: consecutive threads, consecutive addresses) and the block adds each value to dw / db once.
__global__ void __launch_bounds__(256, 2) conv_c3_wgrad_kernel(const float* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ dz, int lddz,
                                                            float* __restrict__ dw, float* __restrict__ db, int B,
                                                            int H, int W, int Cout, int numTiles, int f16) {
  TraceScope trace(4);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  __shared__ __align__(16) float patch[C3_P][C3_ROW];
  __shared__ float comb[49][128];
  const int Ho = H / 2, Wo = W / 2;
  const int tilesX = Wo / C3_T, tilesY = Ho / C3_T;
  const int cp = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int co = blockIdx.y * 128 + 2 * cp;
  for (int i = threadIdx.x; i < 49 * 128; i += blockDim.x) (&comb[0][0])[i] = 0.f;
  float2 acc[2][24];
#pragma unroll
  for (int j = 0; j < 24; ++j) acc[0][j] = acc[1][j] = make_float2(0.f, 0.f);
  float2 accb = make_float2(0.f, 0.f);
  for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
    const int b = tile / (tilesX * tilesY);
    const int oy0 = ((tile / tilesX) % tilesY) * C3_T, ox0 = (tile % tilesX) * C3_T;
    __syncthreads();
    c3_load_patch<C3_T>(patch, x, b, oy0, ox0, H, W);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < C3_T / 4; ++r) {
      const int py = grp * (C3_T / 4) + r;
      uint32_t g[C3_T];
#pragma unroll
      for (int px = 0; px < C3_T; ++px)  // the row's 8 gradient pairs are requested together
        g[px] = __ldg(reinterpret_cast<const uint32_t*>(dz + (((long long)b * Ho + oy0 + py) * Wo + ox0 + px) * lddz + co));
#pragma unroll 2
      for (int px = 0; px < C3_T; ++px) {
        const float g0 = h_lo(g[px], f16), g1 = h_hi(g[px], f16);
        accb.x += g0;
        accb.y += g1;
        const float2 g00 = make_float2(g0, g0), g11 = make_float2(g1, g1);
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          const float2* row = reinterpret_cast<const float2*>(&patch[2 * py + ky][2 * px * 3]);
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const float2 xv = row[j];
            acc[0][ky * 6 + j] = ffma2(xv, g00, acc[0][ky * 6 + j]);
            acc[1][ky * 6 + j] = ffma2(xv, g11, acc[1][ky * 6 + j]);
          }
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 24; ++j) {
    atomicAdd(&comb[2 * j][2 * cp], acc[0][j].x);
    atomicAdd(&comb[2 * j][2 * cp + 1], acc[1][j].x);
    atomicAdd(&comb[2 * j + 1][2 * cp], acc[0][j].y);
    atomicAdd(&comb[2 * j + 1][2 * cp + 1], acc[1][j].y);
  }
  atomicAdd(&comb[48][2 * cp], accb.x);
  atomicAdd(&comb[48][2 * cp + 1], accb.y);
  __syncthreads();
  const int cbase = blockIdx.y * 128;
  for (int i = threadIdx.x; i < 48 * 128; i += blockDim.x) atomicAdd(dw + (i >> 7) * Cout + cbase + (i & 127), (&comb[0][0])[i]);
  if (db != nullptr && threadIdx.x < 128) atomicAdd(db + cbase + threadIdx.x, comb[48][threadIdx.x]);
  trace.end();
}

int conv4s2_c3_wgrad(const float* x, const __nv_bfloat16* dz, int lddz, float* dw, float* db, int B, int H, int W,
                     int Cout, int zero, cudaStream_t st) {
  if ((H / 2) % C3_T || (W / 2) % C3_T || Cout % 128 || lddz % 2) {
    set_error("conv4s2_c3_wgrad: unsupported shape H=%d W=%d Cout=%d lddz=%d", H, W, Cout, lddz);
    return 1;
  }
  if (zero) {
    cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)48 * Cout * sizeof(float), st);
    if (e == cudaSuccess && db != nullptr) e = cudaMemsetAsync(db, 0, (size_t)Cout * sizeof(float), st);
    if (e != cudaSuccess) {
      set_error("conv4s2_c3_wgrad memset: %s", cudaGetErrorString(e));
      return 1;
    }
  }
  const int numTiles = B * (H / 2 / C3_T) * (W / 2 / C3_T);
  // One tile per block up to two blocks per SM; beyond that a block walks several tiles, because the global atomics
  // per block (49 x 128) are the cost that does not shrink with the tile count (debug key 15 varies the cap).
  int gx = g_c3w_blocks > 0 ? g_c3w_blocks : 2 * g_ew_sms;
  if (gx > numTiles) gx = numTiles;
  dim3 grid(gx, Cout / 128);
  launch_k(conv_c3_wgrad_kernel, dim3(grid), dim3(256), 0, st, x, dz, lddz, dw, db, B, H, W, Cout, numTiles, g_f16);
  GCT2_CHECK_LAUNCH("conv_c3_wgrad_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ 3x3 / stride-1 conv on the image
// train.py:131-139 with block_depth > 0: the first Conv2D of the outermost Block (train.py:192) reads the 3-channel
// image (K = 27): CUDA-core direct convolution.  A thread owns one pixel x 8 output channels; the kernel (27 x Cout fp32)
// and the bias sit in shared memory, the 27 inputs of a pixel are broadcast loads shared by the threads of that pixel.
__global__ void __launch_bounds__(256) conv3s1_c3_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias,
                                                               __nv_bfloat16* __restrict__ y, int ldy, int B, int H,
                                                               int W, int Cout, int f16) {
  TraceScope trace(30);
  extern __shared__ __align__(16) float c3s1_smem[];
  float* ws = c3s1_smem;             // [27][Cout]
  float* bs = c3s1_smem + 27 * Cout;  // [Cout]
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) ws[i] = __ldg(w + i);
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) bs[i] = __ldg(bias + i);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  __syncthreads();
  const int groups = Cout / 8, ppb = blockDim.x / groups;
  const int grp = threadIdx.x % groups, pl = threadIdx.x / groups;
  const long long pixels = (long long)B * H * W;
  if (pl < ppb) {
    for (long long p = (long long)blockIdx.x * ppb + pl; p < pixels; p += (long long)gridDim.x * ppb) {
      const int xx = (int)(p % W), yy = (int)((p / W) % H);
      const long long b = p / ((long long)W * H);
      float acc[8];
      const float4 b0 = *reinterpret_cast<const float4*>(bs + grp * 8), b1 = *reinterpret_cast<const float4*>(bs + grp * 8 + 4);
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = yy + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = xx + kx - 1;
          const bool in = iy >= 0 && iy < H && ix >= 0 && ix < W;
          const float* xp = x + ((b * H + iy) * W + ix) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float v = in ? __ldg(xp + c) : 0.f;
            const float* wr = ws + ((ky * 3 + kx) * 3 + c) * Cout + grp * 8;
            const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
            acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]);
            acc[3] = fmaf(v, w0.w, acc[3]); acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
            acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
          }
        }
      }
      uint4 o;
      o.x = pack_h2(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), f16);
      o.y = pack_h2(fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f), f16);
      o.z = pack_h2(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f), f16);
      o.w = pack_h2(fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f), f16);
      *reinterpret_cast<uint4*>(y + p * ldy + grp * 8) = o;
    }
  }
  trace.end();
}

int conv3s1_c3_fprop(const float* x, const float* w, const float* bias, __nv_bfloat16* y, int ldy, int B, int H,
                     int W, int Cout, cudaStream_t st) {
  if (Cout % 8 || Cout < 8 || Cout > 1024 || ldy % 8 || B < 1 || H < 1 || W < 1) {
    set_error("conv3s1_c3_fprop: unsupported shape B=%d H=%d W=%d Cout=%d ldy=%d", B, H, W, Cout, ldy);
    return 1;
  }
  const size_t smem = (size_t)28 * Cout * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(conv3s1_c3_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("conv3s1_c3_fprop: %s", cudaGetErrorString(e));
      return 1;
    }
  }
  const int groups = Cout / 8, ppb = 256 / groups;
  if (ppb < 1) {
    set_error("conv3s1_c3_fprop: Cout=%d needs more than one block per pixel", Cout);
    return 1;
  }
  const long long pixels = (long long)B * H * W;
  long long blocks = (pixels + ppb - 1) / ppb;
  if (blocks > (long long)g_ew_sms * 8) blocks = (long long)g_ew_sms * 8;
  launch_k(conv3s1_c3_fprop_kernel, dim3((int)blocks), dim3(256), smem, st, x, w, bias, y, ldy, B, H, W, Cout, g_f16);
  GCT2_CHECK_LAUNCH("conv3s1_c3_fprop_kernel");
  return 0;
}

// wgrad: dw[ky,kx,c,co] += sum_pix x[pix + (ky-1, kx-1), c] * dz[pix, co].  A thread owns one output channel and a
// stride of the block's pixels with its 27 sums in registers (the pixel's 27 inputs are the same address for all
// threads of a warp: broadcast loads); lanes are combined in shared memory, one global atomic per value per block.
__global__ void __launch_bounds__(256) conv3s1_c3_wgrad_kernel(const float* __restrict__ x,
                                                               const __nv_bfloat16* __restrict__ dz, int lddz,
                                                               float* __restrict__ dw, int B, int H, int W, int Cout,
                                                               int chBlock, long long pixPerBlock, int f16) {
  TraceScope trace(31);
  extern __shared__ __align__(16) float c3s1w_smem[];  // [27][chBlock]
  for (int i = threadIdx.x; i < 27 * chBlock; i += blockDim.x) c3s1w_smem[i] = 0.f;
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  __syncthreads();
  const int lanes = blockDim.x / chBlock;
  const int cl = threadIdx.x % chBlock, lane = threadIdx.x / chBlock;
  const int co = blockIdx.y * chBlock + cl;
  const long long pixels = (long long)B * H * W;
  const long long p0 = (long long)blockIdx.x * pixPerBlock;
  const long long p1 = p0 + pixPerBlock < pixels ? p0 + pixPerBlock : pixels;
  float acc[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) acc[k] = 0.f;
  if (lane < lanes) {
    const unsigned short* dzs = reinterpret_cast<const unsigned short*>(dz);
    for (long long p = p0 + lane; p < p1; p += lanes) {
      const unsigned short raw = __ldg(dzs + p * lddz + co);
      const float g = h_lo((uint32_t)raw, f16);
      const int xx = (int)(p % W), yy = (int)((p / W) % H);
      const long long b = p / ((long long)W * H);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = yy + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = xx + kx - 1;
          const bool in = iy >= 0 && iy < H && ix >= 0 && ix < W;
          const float* xp = x + ((b * H + iy) * W + ix) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float v = in ? __ldg(xp + c) : 0.f;
            acc[(ky * 3 + kx) * 3 + c] = fmaf(v, g, acc[(ky * 3 + kx) * 3 + c]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 27; ++k) atomicAdd(&c3s1w_smem[k * chBlock + cl], acc[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * chBlock; i += blockDim.x)
    atomicAdd(dw + (long long)(i / chBlock) * Cout + blockIdx.y * chBlock + (i % chBlock), c3s1w_smem[i]);
  trace.end();
}

int conv3s1_c3_wgrad(const float* x, const __nv_bfloat16* dz, int lddz, float* dw, int B, int H, int W, int Cout,
                     int zero, cudaStream_t st) {
  int chBlock = Cout < 256 ? Cout : 256;
  if (Cout < 8 || 256 % chBlock || Cout % chBlock || B < 1 || H < 1 || W < 1) {
    set_error("conv3s1_c3_wgrad: Cout must be a power of two >= 8 or a multiple of 256 (got %d)", Cout);
    return 1;
  }
  if (zero) {
    cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)27 * Cout * sizeof(float), st);
    if (e != cudaSuccess) {
      set_error("conv3s1_c3_wgrad memset: %s", cudaGetErrorString(e));
      return 1;
    }
  }
  const long long pixels = (long long)B * H * W;
  long long gx = 2LL * g_ew_sms;
  if (gx > (pixels + 63) / 64) gx = (pixels + 63) / 64;
  if (gx < 1) gx = 1;
  const long long ppb = (pixels + gx - 1) / gx;
  gx = (pixels + ppb - 1) / ppb;
  launch_k(conv3s1_c3_wgrad_kernel, dim3((int)gx, Cout / chBlock), dim3(256), (size_t)27 * chBlock * sizeof(float), st, x, dz,
           lddz, dw, B, H, W, Cout, chBlock, ppb, g_f16);
  GCT2_CHECK_LAUNCH("conv3s1_c3_wgrad_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ Dense(3) + MSE (a6,a7)
// LPP lanes share one pixel (lane `sub` owns 8 of the Cu = 8*LPP up0 channels: one 16-byte load / store), so a warp
// walks 32/LPP pixels per iteration; the 3 image channels are handled by every lane of the group (broadcast loads).
//   pred = [u0 | noised] . Wd + bd ;  loss = sum((pred - x)^2) * invN ;  dpred = 2 (pred - x) * invN
//   du0 = relu'(u0) * (dpred . Wd^T)  (bf16) ;  dWd, dbd by warp-shuffle + shared-memory reduction, one atomic per
//   block per value (the caller zeroes loss / dwd / dbd).
template <int LPP>
__global__ void __launch_bounds__(256) dense_mse_kernel(const __nv_bfloat16* __restrict__ u0, int ldu,
                                                        const float* __restrict__ noised,
                                                        const float* __restrict__ x, const float* __restrict__ wd,
                                                        const float* __restrict__ bd, float* __restrict__ pred,
                                                        float* __restrict__ loss, __nv_bfloat16* __restrict__ du0,
                                                        int lddu, float* __restrict__ dwd, float* __restrict__ dbd,
                                                        long long pixels, float invN, int backward, int f16,
                                                        const float* __restrict__ loss_scale,
                                                        const float* __restrict__ eps, const int* __restrict__ t_int,
                                                        long long ppi, int mode, int steps) {
  TraceScope trace(5);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  constexpr int PPW = 32 / LPP;        // pixels per warp iteration
  constexpr int CU = 8 * LPP;
  constexpr int NRED = CU * 3 + 9 + 3 + 1;  // dWd(u0 part) | dWd(image part) | dbd | loss
  __shared__ float red[8][NRED];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPP, pg = lane / LPP;
  const long long warpGlobal = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const long long numWarps = (long long)gridDim.x * (blockDim.x >> 5);
  float w[8][3], wn[9], bv[3];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int j = 0; j < 3; ++j) w[c][j] = __ldg(wd + (sub * 8 + c) * 3 + j);
  const bool img = noised != nullptr;  // false: Dense(3) on the CU 16-bit channels only (wd is [CU,3])
#pragma unroll
  for (int k = 0; k < 9; ++k) wn[k] = img ? __ldg(wd + CU * 3 + k) : 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) bv[j] = __ldg(bd + j);
  const float gradScale = loss_scale != nullptr ? invN * __ldg(loss_scale) : invN;
  float g[8][3], gn[9], gb[3] = {0.f, 0.f, 0.f}, lossAcc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) g[c][0] = g[c][1] = g[c][2] = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) gn[k] = 0.f;

  for (long long p0 = warpGlobal * PPW; p0 < pixels; p0 += numWarps * PPW) {
    const long long p = p0 + pg;
    const bool live = p < pixels;
    float a[8];
    float nz[3] = {0.f, 0.f, 0.f}, xv[3] = {0.f, 0.f, 0.f};
    // train.py:238-252: target = ca*x + cb*eps, compared with sc*prediction (predict_x, the default: ca = sc = 1, cb = 0)
    float ca = 1.f, cb = 0.f, sc = 1.f;
    if (mode != 0 && live) {
      const float t = (float)__ldg(t_int + (int)(p / ppi));
      const float ab = abar_of_f(t, steps), sb = sqrtf(1.f - ab);
      if (mode & 8) {  // ordinary_differential_equation: the noised image of step t - 1
        const float a1 = abar_of_f(t - 1.f, steps);
        ca = sqrtf(a1);
        cb = sqrtf(1.f - a1);
      } else {
        ca = 0.f;
        cb = (mode & 2) ? sb : 1.f;        // predict_scaled_epsilon
        if (mode & 4) {                    // prediction_weighting
          cb *= sb;
          sc = sb;
        }
      }
    }
    if (live) {
      const uint4 uv = __ldg(reinterpret_cast<const uint4*>(u0 + p * ldu + sub * 8));
      a[0] = h_lo(uv.x, f16); a[1] = h_hi(uv.x, f16); a[2] = h_lo(uv.y, f16); a[3] = h_hi(uv.y, f16);
      a[4] = h_lo(uv.z, f16); a[5] = h_hi(uv.z, f16); a[6] = h_lo(uv.w, f16); a[7] = h_hi(uv.w, f16);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (img) nz[c] = __ldg(noised + p * 3 + c);
        xv[c] = ca * __ldg(x + p * 3 + c);
        if (cb != 0.f) xv[c] = fmaf(cb, __ldg(eps + p * 3 + c), xv[c]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) a[c] = 0.f;
    }
    float s[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int j = 0; j < 3; ++j) s[j] = fmaf(a[c], w[c][j], s[j]);
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1)
#pragma unroll
      for (int j = 0; j < 3; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
    float d[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float pj = s[j] + nz[0] * wn[j] + nz[1] * wn[3 + j] + nz[2] * wn[6 + j] + bv[j];
      if (pred != nullptr && live && sub == j) pred[p * 3 + j] = pj;
      const float diff = live ? sc * pj - xv[j] : 0.f;
      if (sub == 0) lossAcc = fmaf(diff, diff, lossAcc);
      d[j] = 2.f * diff * sc * gradScale;  // loss scaling (mixed precision): the backward pass carries scale * gradient
    }
    if (backward && live) {
      float r[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        r[c] = a[c] > 0.f ? d[0] * w[c][0] + d[1] * w[c][1] + d[2] * w[c][2] : 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) g[c][j] = fmaf(a[c], d[j], g[c][j]);
      }
      uint4 o;
      o.x = pack_h2(r[0], r[1], f16); o.y = pack_h2(r[2], r[3], f16);
      o.z = pack_h2(r[4], r[5], f16); o.w = pack_h2(r[6], r[7], f16);
      *reinterpret_cast<uint4*>(du0 + p * lddu + sub * 8) = o;
      if (sub == 0) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          gb[j] += d[j];
#pragma unroll
          for (int c = 0; c < 3; ++c) gn[c * 3 + j] = fmaf(nz[c], d[j], gn[c * 3 + j]);
        }
      }
    }
  }
  // reduce over the pixel groups of the warp (lanes with equal `sub`), then over the block's warps
#pragma unroll
  for (int o = LPP; o < 32; o <<= 1) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int j = 0; j < 3; ++j) g[c][j] += __shfl_xor_sync(0xffffffffu, g[c][j], o);
#pragma unroll
    for (int k = 0; k < 9; ++k) gn[k] += __shfl_xor_sync(0xffffffffu, gn[k], o);
#pragma unroll
    for (int j = 0; j < 3; ++j) gb[j] += __shfl_xor_sync(0xffffffffu, gb[j], o);
    lossAcc += __shfl_xor_sync(0xffffffffu, lossAcc, o);
  }
  if (pg == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int j = 0; j < 3; ++j) red[warp][(sub * 8 + c) * 3 + j] = g[c][j];
    if (sub == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) red[warp][CU * 3 + k] = gn[k];
#pragma unroll
      for (int j = 0; j < 3; ++j) red[warp][CU * 3 + 9 + j] = gb[j];
      red[warp][CU * 3 + 12] = lossAcc;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NRED; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) v += red[wv][i];
    if (i == CU * 3 + 12)
      atomicAdd(loss, v * invN);
    else if (backward) {
      if (i < CU * 3)
        atomicAdd(dwd + i, v);
      else if (i < CU * 3 + 9) {
        if (img) atomicAdd(dwd + i, v);
      } else
        atomicAdd(dbd + (i - CU * 3 - 9), v);
    }
  }
  trace.end();
}

int dense_mse(const __nv_bfloat16* u0, int ldu, const float* noised, const float* x, const float* wd,
              const float* bd, float* pred, float* loss, __nv_bfloat16* du0, int lddu, float* dwd, float* dbd,
              long long pixels, int Cu, float invN, int backward, int zero, const float* loss_scale, const float* eps,
              const int* t_int, long long pixels_per_image, int target_mode, int steps, cudaStream_t st) {
  if (target_mode != 0 && (t_int == nullptr || pixels_per_image < 1 || steps < 1 || ((target_mode & 7) && eps == nullptr) ||
                           ((target_mode & 8) && eps == nullptr))) {
    set_error("dense_mse: target mode %d needs t_int, eps, pixels_per_image and steps", target_mode);
    return 1;
  }
  if (Cu != 64 && Cu != 128) {
    set_error("dense_mse: the fused kernel expects 64 or 128 up0 channels (+3 image channels), got %d", Cu);
    return 1;
  }
  if ((ldu % 8) || (backward && (lddu % 8))) {
    set_error("dense_mse: pixel strides must be multiples of 8 elements");
    return 1;
  }
  if (zero) {
    cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), st);
    if (backward && e == cudaSuccess) e = cudaMemsetAsync(dwd, 0, (size_t)(Cu + (noised ? 3 : 0)) * 3 * sizeof(float), st);
    if (backward && e == cudaSuccess) e = cudaMemsetAsync(dbd, 0, 3 * sizeof(float), st);
    if (e != cudaSuccess) {
      set_error("dense_mse memset: %s", cudaGetErrorString(e));
      return 1;
    }
  }
  long long want = (pixels * (Cu / 8) + 255) / 256;  // one 16-byte vector per thread per iteration
  int blocks = (int)(want < (long long)g_ew_sms * g_dense_bps ? want : (long long)g_ew_sms * g_dense_bps);
  if (blocks < 1) blocks = 1;
  if (Cu == 64)
    launch_k(dense_mse_kernel<8>, dim3(blocks), dim3(256), 0, st, u0, ldu, noised, x, wd, bd, pred, loss, du0, lddu, dwd, dbd, pixels, invN,
                                                backward, g_f16, loss_scale, eps, t_int, pixels_per_image, target_mode, steps);
  else
    launch_k(dense_mse_kernel<16>, dim3(blocks), dim3(256), 0, st, u0, ldu, noised, x, wd, bd, pred, loss, du0, lddu, dwd, dbd, pixels,
                                                 invN, backward, g_f16, loss_scale, eps, t_int, pixels_per_image, target_mode, steps);
  GCT2_CHECK_LAUNCH("dense_mse_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ residual = True at the image level
// train.py:106-112 with residual = True and block_depth = 0: the outermost Residual returns
//   r = noised + up0 . Wp          (Dense(3, use_bias=False) on up0's U channels, train.py:107-111)
// and Dense(3) follows (train.py:198-202): pred = r . Wd + bd = up0 . (Wp Wd) + noised . Wd + bd -- the fused
// Dense(3)+MSE kernel with the effective kernel Weff = [Wp Wd ; Wd] ([U+3, 3]).  Its outputs map back linearly:
//   dWp = dWeff[:U] . Wd^T ,  dWd = Wp^T . dWeff[:U] + dWeff[U:] ;  du0 = relu'(up0) (dpred . Weff[:U]^T) is already right.
// Both maps are U x 3 x 3 multiply-adds: one small block each.
__global__ void res0_compose_kernel(const float* __restrict__ wp, const float* __restrict__ wd, float* __restrict__ weff, int U) {
  TraceScope trace(32);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (U + 3) * 3; i += gridDim.x * blockDim.x) {
    const int r = i / 3, j = i - 3 * r;
    float v;
    if (r < U)
      v = wp[r * 3] * wd[j] + wp[r * 3 + 1] * wd[3 + j] + wp[r * 3 + 2] * wd[6 + j];
    else
      v = wd[(r - U) * 3 + j];
    weff[i] = v;
  }
  trace.end();
}
__global__ void res0_decompose_kernel(const float* __restrict__ dweff, const float* __restrict__ wp,
                                      const float* __restrict__ wd, float* __restrict__ dwp, float* __restrict__ dwd, int U) {
  TraceScope trace(33);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  for (int i = threadIdx.x; i < U * 3; i += blockDim.x) {
    const int r = i / 3, k = i - 3 * r;
    dwp[i] = dweff[r * 3] * wd[k * 3] + dweff[r * 3 + 1] * wd[k * 3 + 1] + dweff[r * 3 + 2] * wd[k * 3 + 2];
  }
  if (threadIdx.x < 9) {
    const int k = threadIdx.x / 3, j = threadIdx.x - 3 * k;
    float v = dweff[(U + k) * 3 + j];
    for (int r = 0; r < U; ++r) v = fmaf(wp[r * 3 + k], dweff[r * 3 + j], v);
    dwd[threadIdx.x] = v;
  }
  trace.end();
}
int res0_compose(const float* wp, const float* wd, float* weff, int U, cudaStream_t st) {
  if (U < 1) {
    set_error("res0_compose: U must be positive");
    return 1;
  }
  launch_k(res0_compose_kernel, dim3(1), dim3(256), 0, st, wp, wd, weff, U);
  GCT2_CHECK_LAUNCH("res0_compose_kernel");
  return 0;
}
int res0_decompose(const float* dweff, const float* wp, const float* wd, float* dwp, float* dwd, int U, cudaStream_t st) {
  if (U < 1) {
    set_error("res0_decompose: U must be positive");
    return 1;
  }
  launch_k(res0_decompose_kernel, dim3(1), dim3(256), 0, st, dweff, wp, wd, dwp, dwd, U);
  GCT2_CHECK_LAUNCH("res0_decompose_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ bias gradients
// db[c] = sum_rows dz[row, c] for up to BG_MAX_SEG tensors in ONE launch (every conv layer's BiasAddGrad).  A block
// owns a row range of one segment; thread = channel pair, blockDim.x/pairs row groups; shared-memory reduce, one
// atomic per block per channel (the caller zeroes the outputs).
constexpr int BG_MAX_SEG = 16;
struct BiasGradSegs {
  const __nv_bfloat16* dz[BG_MAX_SEG];
  float* db[BG_MAX_SEG];
  long long rows[BG_MAX_SEG];
  int ld[BG_MAX_SEG], C[BG_MAX_SEG], firstBlock[BG_MAX_SEG + 1], rowsPerBlock[BG_MAX_SEG];
  int n, f16;
};

__global__ void __launch_bounds__(256) bias_grad_kernel(const __grid_constant__ BiasGradSegs sg) {
  TraceScope trace(6);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  extern __shared__ float red[];  // [rowsPerIter][C]
  int s = 0;
  while (s + 1 < sg.n && (int)blockIdx.x >= sg.firstBlock[s + 1]) ++s;
  const __nv_bfloat16* __restrict__ dz = sg.dz[s];
  const int C = sg.C[s], ld = sg.ld[s], f16 = sg.f16;
  const int vecs = C / 8;                  // threads per row, 16 bytes each
  const int rpi = blockDim.x / vecs;       // rows per iteration
  const int v = threadIdx.x % vecs, rg = threadIdx.x / vecs;
  const long long r0 = (long long)(blockIdx.x - sg.firstBlock[s]) * sg.rowsPerBlock[s];
  long long r1 = r0 + sg.rowsPerBlock[s];
  if (r1 > sg.rows[s]) r1 = sg.rows[s];
  if (rg < rpi) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    long long r = r0 + rg;
    for (; r + 3LL * rpi < r1; r += 4LL * rpi) {  // four independent 16-byte loads in flight
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(dz + (r + (long long)u * rpi) * ld) + v);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[0] += h_lo(q[u].x, f16); acc[1] += h_hi(q[u].x, f16); acc[2] += h_lo(q[u].y, f16); acc[3] += h_hi(q[u].y, f16);
        acc[4] += h_lo(q[u].z, f16); acc[5] += h_hi(q[u].z, f16); acc[6] += h_lo(q[u].w, f16); acc[7] += h_hi(q[u].w, f16);
      }
    }
    for (; r < r1; r += rpi) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(dz + r * ld) + v);
      acc[0] += h_lo(q.x, f16); acc[1] += h_hi(q.x, f16); acc[2] += h_lo(q.y, f16); acc[3] += h_hi(q.y, f16);
      acc[4] += h_lo(q.z, f16); acc[5] += h_hi(q.z, f16); acc[6] += h_lo(q.w, f16); acc[7] += h_hi(q.w, f16);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rg * C + v * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int g = 0; g < rpi; ++g) t += red[g * C + c];
    atomicAdd(sg.db[s] + c, t);
  }
  trace.end();
}

int bias_grad_multi(int n, const __nv_bfloat16* const* dz, const int* ld, const long long* rows, const int* C,
                    float* const* db, int zero, cudaStream_t st) {
  if (n < 1 || n > BG_MAX_SEG) {
    set_error("bias_grad: between 1 and %d tensors per call, got %d", BG_MAX_SEG, n);
    return 1;
  }
  BiasGradSegs sg;
  sg.n = n;
  sg.f16 = g_f16;
  const int threads = 256;
  int block = 0;
  double totalBytes = 0;
  for (int i = 0; i < n; ++i) {
    if (C[i] % 8 || C[i] > 2048 || C[i] < 8 || rows[i] < 1 || ld[i] % 8) {
      set_error("bias_grad: unsupported tensor %d (C=%d ld=%d rows=%lld): channels and pixel stride must be multiples of 8",
                i, C[i], ld[i], rows[i]);
      return 1;
    }
    totalBytes += (double)rows[i] * C[i];
  }
  const double budget = (double)g_ew_sms * 8;  // blocks over all segments, shared in proportion to bytes
  size_t smem = 0;
  for (int i = 0; i < n; ++i) {
    const int rpi = threads / (C[i] / 8);
    long long nb = (long long)(budget * ((double)rows[i] * C[i]) / totalBytes + 0.999);
    const long long cap = (rows[i] + 4LL * rpi - 1) / (4LL * rpi);  // at least one unrolled iteration per block
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    const int rpb = (int)((rows[i] + nb - 1) / nb);
    nb = (rows[i] + rpb - 1) / rpb;
    sg.dz[i] = dz[i]; sg.db[i] = db[i]; sg.rows[i] = rows[i]; sg.ld[i] = ld[i]; sg.C[i] = C[i];
    sg.firstBlock[i] = block; sg.rowsPerBlock[i] = rpb;
    block += (int)nb;
    const size_t need = (size_t)rpi * C[i] * sizeof(float);
    if (need > smem) smem = need;
    if (zero) {
      cudaError_t e = cudaMemsetAsync(db[i], 0, (size_t)C[i] * sizeof(float), st);
      if (e != cudaSuccess) {
        set_error("bias_grad memset: %s", cudaGetErrorString(e));
        return 1;
      }
    }
  }
  sg.firstBlock[n] = block;
  launch_k(bias_grad_kernel, dim3(block), dim3(threads), smem, st, sg);
  GCT2_CHECK_LAUNCH("bias_grad_kernel");
  return 0;
}

int bias_grad(const __nv_bfloat16* dz, int ld, long long rows, int C, float* db, cudaStream_t st) {
  return bias_grad_multi(1, &dz, &ld, &rows, &C, &db, 1, st);
}

// ------------------------------------------------------------------------------------ Keras Adam (a9)
// state[0] = iteration count (int64, 0-based), written back incremented; alpha goes to hyper[0].
//   lr(step) = base*(step+1)/(warmup+1) while step < warmup else base            (train.py:57-65)
//   alpha = lr*sqrt(1-b2^t)/(1-b1^t), t = step+1 ; m += (g-m)(1-b1) ; v += (g*g-v)(1-b2) ; w -= alpha*m/(sqrt(v)+eps)
__global__ void adam_prepare_kernel(long long* __restrict__ iterations, float* __restrict__ hyper, float base,
                                    int warmup, float b1, float b2) {
  TraceScope trace(7);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  const long long step = *iterations;
  float lr = base;
  if (step < warmup) lr = base * (float)(step + 1) / (float)(warmup + 1);
  const float t = (float)(step + 1);
  const float b1p = powf(b1, t), b2p = powf(b2, t);
  hyper[0] = lr * sqrtf(1.f - b2p) / (1.f - b1p);
  hyper[1] = lr;
  *iterations = step + 1;
  trace.end();
}

// Memory-level parallelism is what this kernel lives on: it shares SMs with the conv CTAs (which leave room for ~14 K
// registers, i.e. one small block), so each thread keeps ADAM_U float4 of every array in flight (16 x 16-byte loads
// issued before the first use) and a block is only 128 threads.  Element order and arithmetic are unchanged.
#ifndef GCT2_ADAM_U
#define GCT2_ADAM_U 1
#endif
#ifndef GCT2_ADAM_THREADS
#define GCT2_ADAM_THREADS 256
#endif
#ifndef GCT2_ADAM_MINB
#define GCT2_ADAM_MINB (640 / GCT2_ADAM_THREADS)  // A/B hook: resident blocks per SM the register budget is sized for
#endif
constexpr int ADAM_U = GCT2_ADAM_U;
constexpr int ADAM_THREADS = GCT2_ADAM_THREADS;
__device__ __forceinline__ void adam_update(float4& wv, float4& mv, float4& vv, float4 gv, float gscale, float c1,
                                            float c2, float alpha, float eps) {
  gv.x *= gscale; gv.y *= gscale; gv.z *= gscale; gv.w *= gscale;
  adam_elem4(wv, mv, vv, gv, c1, c2, alpha, eps);
}
__device__ __forceinline__ float4 round_f16x4(float4 v) {
  const uint32_t a = pack_f16x2(v.x, v.y), b = pack_f16x2(v.z, v.w);
  return make_float4(f16_lo(a), f16_hi(a), f16_lo(b), f16_hi(b));
}
// G16: the gradient arrives as bf16 (data parallel: the reduce-scatter ran on a bf16 copy, half the NVLink bytes)
template <bool G16>
__device__ __forceinline__ float4 adam_load_grad(const void* __restrict__ g, long long i) {
  if (G16) {
    const uint2 q = __ldcs(reinterpret_cast<const uint2*>(g) + i);
    return make_float4(bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y));
  }
  return __ldcs(reinterpret_cast<const float4*>(g) + i);  // the gradient is dead after this read
}
template <bool G16>
__global__ void __launch_bounds__(ADAM_THREADS, GCT2_ADAM_MINB) adam_kernel(float4* __restrict__ w, float4* __restrict__ m,
                                                            float4* __restrict__ v, const void* __restrict__ g,
                                                            uint2* __restrict__ wb, long long nvec,
                                                            const float* __restrict__ hyper, float b1, float b2,
                                                            float eps, float gscale,
                                                            long long* __restrict__ iterations_inc, int f16,
                                                            const float* __restrict__ ls) {
  TraceScope trace(8);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  const float alpha = __ldg(hyper);
  // dynamic loss scaling (train.py:82-83, Keras LossScaleOptimizer): ls = {scale, good steps, all gradients finite, 1/scale}.
  // A step whose gradients overflowed is skipped as a whole -- no update, no iteration count -- and the gradients carry
  // the scale, removed here.
  if (ls != nullptr) {
    if (__ldg(ls + 2) == 0.f) {
      trace.end();
      return;
    }
    gscale *= __ldg(ls + 3);
  }
  // mixed precision: the reference's weight gradients are fp16 tensors (cast to fp32 before the unscaling): same rounding
  const bool roundG = ls != nullptr && f16 != 0;
  if (iterations_inc != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *iterations_inc += 1;
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  const long long T = (long long)gridDim.x * blockDim.x;
  long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  // full groups: ADAM_U strided vectors per thread, all loads in flight before the first use
  for (; i0 + (ADAM_U - 1) * T < nvec; i0 += T * ADAM_U) {
    float4 gv[ADAM_U], mv[ADAM_U], vv[ADAM_U], wv[ADAM_U];
#pragma unroll
    for (int u = 0; u < ADAM_U; ++u) {
      gv[u] = adam_load_grad<G16>(g, i0 + u * T);
      mv[u] = m[i0 + u * T];
      vv[u] = v[i0 + u * T];
      wv[u] = w[i0 + u * T];
    }
#pragma unroll
    for (int u = 0; u < ADAM_U; ++u) {
      if (roundG) gv[u] = round_f16x4(gv[u]);
      adam_update(wv[u], mv[u], vv[u], gv[u], gscale, c1, c2, alpha, eps);
      m[i0 + u * T] = mv[u];
      v[i0 + u * T] = vv[u];
      w[i0 + u * T] = wv[u];
      uint2 o;
      o.x = pack_h2(wv[u].x, wv[u].y, f16);
      o.y = pack_h2(wv[u].z, wv[u].w, f16);
      wb[i0 + u * T] = o;
    }
  }
  for (; i0 < nvec; i0 += T) {  // ragged tail
    float4 gv = adam_load_grad<G16>(g, i0), mv = m[i0], vv = v[i0], wv = w[i0];
    if (roundG) gv = round_f16x4(gv);
    adam_update(wv, mv, vv, gv, gscale, c1, c2, alpha, eps);
    m[i0] = mv;
    v[i0] = vv;
    w[i0] = wv;
    uint2 o;
    o.x = pack_h2(wv.x, wv.y, f16);
    o.y = pack_h2(wv.z, wv.w, f16);
    wb[i0] = o;
  }
  trace.end();
}

// SM-partitioned variant: few CTAs, each alone on its SM (the dynamic shared-memory request keeps every other block --
// ours or a conv CTA -- off that SM), 1024 threads x 2 float4 of each of the four arrays in flight per thread = 128 KB
// of loads per SM.  HBM-bound work does not need all 148 SMs to draw most of the bandwidth, and the tensor-core convs of
// the same step do not need HBM: running the two side by side on disjoint SM sets hides the optimiser (engine.py).
// Same element order inside a vector and the same adam_elem arithmetic as adam_kernel: bit-identical results.
constexpr int ADAM_WIDE_THREADS = 1024;
constexpr int ADAM_WIDE_U = 2;
constexpr int ADAM_WIDE_SMEM = 120 * 1024;
__global__ void __launch_bounds__(ADAM_WIDE_THREADS, 1) adam_wide_kernel(float4* __restrict__ w, float4* __restrict__ m,
                                                                         float4* __restrict__ v, const float4* __restrict__ g,
                                                                         uint2* __restrict__ wb, long long nvec,
                                                                         const float* __restrict__ hyper, float b1, float b2,
                                                                         float eps, float gscale,
                                                                         long long* __restrict__ iterations_inc) {
  TraceScope trace(8);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  const float alpha = __ldg(hyper);
  if (iterations_inc != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *iterations_inc += 1;
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  const long long T = (long long)gridDim.x * blockDim.x;
  long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (; i0 + (ADAM_WIDE_U - 1) * T < nvec; i0 += T * ADAM_WIDE_U) {
    float4 gv[ADAM_WIDE_U], mv[ADAM_WIDE_U], vv[ADAM_WIDE_U], wv[ADAM_WIDE_U];
#pragma unroll
    for (int u = 0; u < ADAM_WIDE_U; ++u) {
      gv[u] = __ldcs(g + i0 + u * T);
      mv[u] = __ldcs(m + i0 + u * T);
      vv[u] = __ldcs(v + i0 + u * T);
      wv[u] = __ldcs(w + i0 + u * T);
    }
#pragma unroll
    for (int u = 0; u < ADAM_WIDE_U; ++u) {
      adam_update(wv[u], mv[u], vv[u], gv[u], gscale, c1, c2, alpha, eps);
      __stcs(m + i0 + u * T, mv[u]);
      __stcs(v + i0 + u * T, vv[u]);
      __stcs(w + i0 + u * T, wv[u]);
      uint2 o;
      o.x = pack_bf16x2(wv[u].x, wv[u].y);
      o.y = pack_bf16x2(wv[u].z, wv[u].w);
      wb[i0 + u * T] = o;
    }
  }
  for (; i0 < nvec; i0 += T) {
    float4 gv = __ldcs(g + i0), mv = m[i0], vv = v[i0], wv = w[i0];
    adam_update(wv, mv, vv, gv, gscale, c1, c2, alpha, eps);
    m[i0] = mv;
    v[i0] = vv;
    w[i0] = wv;
    uint2 o;
    o.x = pack_bf16x2(wv.x, wv.y);
    o.y = pack_bf16x2(wv.z, wv.w);
    wb[i0] = o;
  }
  trace.end();
}

// ------------------------------------------------------------------------------------ data parallel: gradient exchange +
// Keras-Adam + weight broadcast in ONE kernel over NVLink peer memory (no NCCL kernel, no SM reserved for a collective)
// Every rank owns 1/N of each gradient bucket.  For its slice the kernel
//   1. reads the bf16 gradients of ALL ranks -- P2P loads from the peers' buffers (torch symmetric memory maps them into
//      this process) or, with NVLS multicast, ONE multimem.ld_reduce per vector that the NVSwitch sums on the way --
//      and accumulates them in fp32 in rank order;
//   2. applies Keras-Adam to its fp32 masters (same adam_elem arithmetic as adam_kernel);
//   3. writes the new bf16 weights straight into the shadow buffer of EVERY rank (P2P stores / one multimem.st).
// That is reduce-scatter + sharded optimiser + all-gather without a collective library: the NVLink transfers are the
// kernel's own loads and stores and overlap its arithmetic element by element.  The caller brackets it with two
// cross-rank barriers (engine.py): all gradients of the bucket complete before, all weight writes landed after.
struct P2PPtrs {
  const uint16_t* g[8];   // per rank: base of its bf16 gradient buffer (symmetric: same layout everywhere)
  uint16_t* w16[8];       // per rank: base of its bf16 / fp16 weight shadow
  const uint16_t* g_mc;   // multicast address of the gradient buffer (nullptr: per-peer loads)
  uint16_t* w16_mc;       // multicast address of the shadow buffer (nullptr: per-peer stores)
  int world;
};
__device__ __forceinline__ uint4 ld_sys_u4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_u4(void* p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// NVLS: the switch adds the bf16 pairs of all ranks (fp32 accumulation inside the switch) and returns the sums
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_u4(void* mc, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}
__global__ void __launch_bounds__(256) adam_p2p_kernel(float4* __restrict__ w, float4* __restrict__ m, float4* __restrict__ v,
                                                       const __grid_constant__ P2PPtrs pp, long long elemOff, long long n8,
                                                       const float* __restrict__ hyper, float b1, float b2, float eps,
                                                       float gscale, int f16, int writeAll) {
  TraceScope trace(15);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  const float alpha = __ldg(hyper);
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const long long e = elemOff + 8 * i;  // first of my eight elements in the flat buffers
    float gs[8];
    if (pp.g_mc != nullptr) {
      const uint4 q = multimem_ld_reduce_bf16x8(pp.g_mc + e);
      gs[0] = bf16_lo(q.x); gs[1] = bf16_hi(q.x); gs[2] = bf16_lo(q.y); gs[3] = bf16_hi(q.y);
      gs[4] = bf16_lo(q.z); gs[5] = bf16_hi(q.z); gs[6] = bf16_lo(q.w); gs[7] = bf16_hi(q.w);
    } else {
      uint4 q[8];
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r < pp.world) q[r] = ld_sys_u4(pp.g[r] + e);  // all peers' vectors in flight together
#pragma unroll
      for (int j = 0; j < 8; ++j) gs[j] = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r < pp.world) {  // rank order: the sum does not depend on which rank computes it
          gs[0] += bf16_lo(q[r].x); gs[1] += bf16_hi(q[r].x); gs[2] += bf16_lo(q[r].y); gs[3] += bf16_hi(q[r].y);
          gs[4] += bf16_lo(q[r].z); gs[5] += bf16_hi(q[r].z); gs[6] += bf16_lo(q[r].w); gs[7] += bf16_hi(q[r].w);
        }
    }
    float4 w0 = w[2 * i], w1 = w[2 * i + 1], m0 = m[2 * i], m1 = m[2 * i + 1], v0 = v[2 * i], v1 = v[2 * i + 1];
    adam_update(w0, m0, v0, make_float4(gs[0], gs[1], gs[2], gs[3]), gscale, c1, c2, alpha, eps);
    adam_update(w1, m1, v1, make_float4(gs[4], gs[5], gs[6], gs[7]), gscale, c1, c2, alpha, eps);
    w[2 * i] = w0; w[2 * i + 1] = w1;
    m[2 * i] = m0; m[2 * i + 1] = m1;
    v[2 * i] = v0; v[2 * i + 1] = v1;
    uint4 o;
    o.x = pack_h2(w0.x, w0.y, f16); o.y = pack_h2(w0.z, w0.w, f16);
    o.z = pack_h2(w1.x, w1.y, f16); o.w = pack_h2(w1.z, w1.w, f16);
    if (!writeAll) {
      // replicated range (every rank updates the same values itself): only my own shadow
      *reinterpret_cast<uint4*>(pp.w16[0] + e) = o;
    } else if (pp.w16_mc != nullptr) {
      multimem_st_u4(pp.w16_mc + e, o);
    } else {
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r < pp.world) st_sys_u4(pp.w16[r] + e, o);
    }
  }
  trace.end();
}

int adam_apply_p2p(float* w, float* m, float* v, const uint16_t* const* g_ptrs, uint16_t* const* w16_ptrs,
                   const uint16_t* g_mc, uint16_t* w16_mc, int world, long long elem_offset, long long n,
                   const float* hyper, float beta1, float beta2, float eps, float grad_scale, int write_all,
                   cudaStream_t st) {
  if (world < 1 || world > 8 || n % 8 || elem_offset % 8 ||
      (reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) % 16) {
    set_error("adam_apply_p2p: 1..8 ranks, ranges of whole 8-element vectors on 16-byte boundaries (world=%d n=%lld off=%lld)",
              world, n, elem_offset);
    return 1;
  }
  if (n == 0) return 0;
  P2PPtrs pp;
  for (int r = 0; r < 8; ++r) {
    pp.g[r] = r < world ? g_ptrs[r] : nullptr;
    pp.w16[r] = r < world ? w16_ptrs[r] : nullptr;
  }
  pp.g_mc = g_mc;
  pp.w16_mc = w16_mc;
  pp.world = world;
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = g_adam_blocks > 0 ? g_adam_blocks : (long long)g_ew_sms * 4;
  if (blocks > cap) blocks = cap;
  launch_k(adam_p2p_kernel, dim3((int)blocks), dim3(256), 0, st, reinterpret_cast<float4*>(w), reinterpret_cast<float4*>(m),
           reinterpret_cast<float4*>(v), pp, elem_offset, n8, hyper, beta1, beta2, eps, grad_scale, g_f16, write_all);
  GCT2_CHECK_LAUNCH("adam_p2p_kernel");
  return 0;
}

// The replicated head region (down0's kernel, biases, Dense: a few thousand fp32 gradients) and the scalar loss, summed
// over all ranks without a collective library: every rank reads every rank's copy (in symmetric memory) and adds them in
// rank order -- the same sum bit for bit on every rank, so the redundant Keras-Adam updates stay identical.  Replaces two
// small NCCL all-reduces (~15-30 us of launch latency each) at the very end of the step, where nothing can hide them.
struct PeerF32 {
  const float* src[8];
  int world;
};
__global__ void __launch_bounds__(256) sum_peers_f32_kernel(const __grid_constant__ PeerF32 pp, float* __restrict__ out_a,
                                                            long long n_a, float* __restrict__ out_b, long long n_b) {
  TraceScope trace(16);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_a + n_b; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < pp.world) asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v[r]) : "l"(pp.src[r] + i) : "memory");
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < pp.world) sum += v[r];
    if (i < n_a)
      out_a[i] = sum;
    else
      out_b[i - n_a] = sum;
  }
  trace.end();
}
int sum_peers_f32(const float* const* src_ptrs, int world, float* out_a, long long n_a, float* out_b, long long n_b,
                  cudaStream_t st) {
  if (world < 1 || world > 8 || n_a < 0 || n_b < 0 || n_a + n_b == 0) {
    set_error("sum_peers_f32: 1..8 ranks and a non-empty range (world=%d n=%lld+%lld)", world, n_a, n_b);
    return 1;
  }
  PeerF32 pp;
  for (int r = 0; r < 8; ++r) pp.src[r] = r < world ? src_ptrs[r] : nullptr;
  pp.world = world;
  long long blocks = (n_a + n_b + 255) / 256;
  if (blocks > g_ew_sms * 2) blocks = g_ew_sms * 2;
  launch_k(sum_peers_f32_kernel, dim3((int)blocks), dim3(256), 0, st, pp, out_a, n_a, out_b, n_b);
  GCT2_CHECK_LAUNCH("sum_peers_f32_kernel");
  return 0;
}

int adam_prepare(long long* iterations, float* hyper, float base_lr, int warmup_steps, float beta1, float beta2,
                 cudaStream_t st) {
  launch_k(adam_prepare_kernel, dim3(1), dim3(1), 0, st, iterations, hyper, base_lr, warmup_steps, beta1, beta2);
  GCT2_CHECK_LAUNCH("adam_prepare_kernel");
  return 0;
}

int adam_apply(float* w, float* m, float* v, const void* g, int g_is_bf16, __nv_bfloat16* w_bf16, long long n,
               const float* hyper, float beta1, float beta2, float eps, float grad_scale, long long* iterations_inc,
               const float* ls, cudaStream_t st) {
  if (n % 4 || (reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) % 16 ||
      reinterpret_cast<uintptr_t>(g) % (g_is_bf16 ? 8 : 16) || reinterpret_cast<uintptr_t>(w_bf16) % 8) {
    set_error("adam: ranges must start on 16-byte boundaries and hold a multiple of 4 elements (n=%lld)", n);
    return 1;
  }
  const long long nvec = n / 4;
  if (nvec == 0) return 0;
  if (!g_is_bf16 && !g_f16 && ls == nullptr && g_adam_sms > 0 && nvec >= (long long)g_adam_sms * ADAM_WIDE_THREADS * ADAM_WIDE_U) {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(adam_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ADAM_WIDE_SMEM);
      attr = true;
    }
    launch_k(adam_wide_kernel, dim3(g_adam_sms), dim3(ADAM_WIDE_THREADS), ADAM_WIDE_SMEM, st, reinterpret_cast<float4*>(w),
             reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), reinterpret_cast<const float4*>(g),
             reinterpret_cast<uint2*>(w_bf16), nvec, hyper, beta1, beta2, eps, grad_scale, iterations_inc);
    GCT2_CHECK_LAUNCH("adam_wide_kernel");
    return 0;
  }
  long long blocks = (nvec + ADAM_THREADS * ADAM_U - 1) / (ADAM_THREADS * ADAM_U);
  const long long cap = g_adam_blocks > 0 ? g_adam_blocks : (long long)g_ew_sms * 8;
  if (blocks > cap) blocks = cap;
  if (g_is_bf16)
    launch_k(adam_kernel<true>, dim3((int)blocks), dim3(ADAM_THREADS), 0, st, reinterpret_cast<float4*>(w),
             reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), g, reinterpret_cast<uint2*>(w_bf16), nvec, hyper,
             beta1, beta2, eps, grad_scale, iterations_inc, g_f16, ls);
  else
    launch_k(adam_kernel<false>, dim3((int)blocks), dim3(ADAM_THREADS), 0, st, reinterpret_cast<float4*>(w),
             reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), g, reinterpret_cast<uint2*>(w_bf16), nvec, hyper,
             beta1, beta2, eps, grad_scale, iterations_inc, g_f16, ls);
  GCT2_CHECK_LAUNCH("adam_kernel");
  return 0;
}

int adam_keras(float* w, float* m, float* v, const float* g, __nv_bfloat16* w_bf16, long long n,
               long long* iterations, float* hyper, float base_lr, int warmup_steps, float beta1, float beta2,
               float eps, float grad_scale, cudaStream_t st) {
  if (adam_prepare(iterations, hyper, base_lr, warmup_steps, beta1, beta2, st)) return 1;
  return adam_apply(w, m, v, g, 0, w_bf16, n, hyper, beta1, beta2, eps, grad_scale, nullptr, nullptr, st);
}

// ------------------------------------------------------------------------------------ fp32 -> bf16 shadow
__global__ void cast_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, long long nvec, int f16) {
  TraceScope trace(9);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 s = __ldg(src + i);
    uint2 o;
    o.x = pack_h2(s.x, s.y, f16);
    o.y = pack_h2(s.z, s.w, f16);
    dst[i] = o;
  }
  trace.end();
}

int cast_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t st) {
  if (n % 4) {
    set_error("cast_bf16: element count must be a multiple of 4, got %lld", n);
    return 1;
  }
  const long long nvec = n / 4;
  long long blocks = (nvec + 255) / 256;
  if (blocks > g_ew_sms * 8) blocks = g_ew_sms * 8;
  if (blocks < 1) blocks = 1;
  launch_k(cast_bf16_kernel, dim3((int)blocks), dim3(256), 0, st, reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst),
                                                nvec, g_f16);
  GCT2_CHECK_LAUNCH("cast_bf16_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ dynamic loss scaling (f3)
// train.py:82-83 wraps the optimiser in tf.keras.mixed_precision.LossScaleOptimizer (dynamic): the loss is multiplied by
// `scale` before backward, the gradients divided by it before the update; a step with a non-finite gradient is skipped
// and halves the scale, `growth` consecutive good steps double it.  State on the device: ls = {scale, good steps,
// finite flag, 1/scale}.
// `limit`: the largest finite value of the gradients' storage format in the reference -- under Keras' mixed_float16 policy
// the weight gradients leave the fp16 convolutions as fp16 tensors, so a scaled gradient beyond 65504 IS an overflow there
// (this implementation accumulates them in fp32 and applies the same limit to skip the same steps).
__global__ void __launch_bounds__(256) loss_scale_check_kernel(const float4* __restrict__ g, long long nvec, float* __restrict__ ls,
                                                               float limit) {
  TraceScope trace(11);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  bool bad = false;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(g + i);
    // |x| <= limit is false for NaN, inf and everything the 16-bit format cannot hold
    bad |= !(fabsf(v.x) <= limit && fabsf(v.y) <= limit && fabsf(v.z) <= limit && fabsf(v.w) <= limit);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) ls[2] = 0.f;
  trace.end();
}
__global__ void loss_scale_update_kernel(float* __restrict__ ls, int growth) {
  TraceScope trace(12);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  float scale = ls[0], good = ls[1];
  if (ls[2] != 0.f) {
    good += 1.f;
    if (good >= (float)growth) {
      if (scale * 2.f < 3.0e38f) scale *= 2.f;
      good = 0.f;
    }
  } else {
    scale = fmaxf(scale * 0.5f, 1.f);
    good = 0.f;
  }
  ls[0] = scale;
  ls[1] = good;
  ls[2] = 1.f;  // armed for the next step
  ls[3] = 1.f / scale;
  trace.end();
}
int loss_scale_check(const float* g, long long n, float* ls, cudaStream_t st) {
  if (n % 4 || reinterpret_cast<uintptr_t>(g) % 16) {
    set_error("loss_scale_check: the gradient range must be 16-byte aligned and a multiple of 4 elements");
    return 1;
  }
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > g_ew_sms * 8) blocks = g_ew_sms * 8;
  if (blocks < 1) blocks = 1;
  launch_k(loss_scale_check_kernel, dim3((int)blocks), dim3(256), 0, st, reinterpret_cast<const float4*>(g), n / 4, ls,
           g_f16 ? 65504.0f : 3.3895314e38f /* bf16 max */);
  GCT2_CHECK_LAUNCH("loss_scale_check_kernel");
  return 0;
}
int loss_scale_update(float* ls, int growth_steps, cudaStream_t st) {
  launch_k(loss_scale_update_kernel, dim3(1), dim3(1), 0, st, ls, growth_steps);
  GCT2_CHECK_LAUNCH("loss_scale_update_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ log_sample's latent edits (f1)
// train.py:418-432: from the inverted latent epsilon_theta [S,S,3] the four latents that are decoded again --
//   out[0] = epsilon_theta, out[1] = pixelated (4x4 average, nearest up-sampling), out[2] = shifted by one pixel along
//   both axes (tf.roll, wrapping), out[3] = quantised: per pixel the nearest (squared distance, first minimum) of the K
//   entries of dictionary [S,S,K,3].  One thread per pixel.
__global__ void __launch_bounds__(256) latent_edits_kernel(const float* __restrict__ e, const float* __restrict__ dict,
                                                           float* __restrict__ out, int S, int K) {
  TraceScope trace(13);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  const long long n = (long long)S * S;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(p / S), x = (int)(p - (long long)y * S);
    const float v0 = __ldg(e + p * 3), v1 = __ldg(e + p * 3 + 1), v2 = __ldg(e + p * 3 + 2);
    float* o = out + p * 3;
    o[0] = v0; o[1] = v1; o[2] = v2;
    // pixelated: mean of the 4x4 block, summed in row-major order (the order of a pooling window)
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    const int by = y & ~3, bx = x & ~3;
    for (int dy = 0; dy < 4; ++dy)
      for (int dx = 0; dx < 4; ++dx) {
        const float* q = e + ((long long)(by + dy) * S + bx + dx) * 3;
        s0 += __ldg(q); s1 += __ldg(q + 1); s2 += __ldg(q + 2);
      }
    o += n * 3;
    o[0] = s0 * 0.0625f; o[1] = s1 * 0.0625f; o[2] = s2 * 0.0625f;
    // shifted: out[y][x] = e[y-1][x-1] with wrap-around
    const float* q = e + ((long long)((y + S - 1) % S) * S + (x + S - 1) % S) * 3;
    o += n * 3;
    o[0] = __ldg(q); o[1] = __ldg(q + 1); o[2] = __ldg(q + 2);
    // quantised
    const float* d = dict + p * K * 3;
    int best = 0;
    float bestErr = 3.4e38f;
    for (int k = 0; k < K; ++k) {
      const float a0 = v0 - __ldg(d + k * 3), a1 = v1 - __ldg(d + k * 3 + 1), a2 = v2 - __ldg(d + k * 3 + 2);
      const float err = a0 * a0 + a1 * a1 + a2 * a2;
      if (err < bestErr) {
        bestErr = err;
        best = k;
      }
    }
    o += n * 3;
    o[0] = __ldg(d + best * 3); o[1] = __ldg(d + best * 3 + 1); o[2] = __ldg(d + best * 3 + 2);
  }
  trace.end();
}
int latent_edits(const float* eps_theta, const float* dictionary, float* out, int S, int K, cudaStream_t st) {
  if (S < 4 || S % 4 || K < 1) {
    set_error("latent_edits: the image side must be a multiple of 4 and the dictionary non-empty (S=%d K=%d)", S, K);
    return 1;
  }
  int blocks = (int)(((long long)S * S + 255) / 256);
  if (blocks > g_ew_sms * 8) blocks = g_ew_sms * 8;
  launch_k(latent_edits_kernel, dim3(blocks), dim3(256), 0, st, eps_theta, dictionary, out, S, K);
  GCT2_CHECK_LAUNCH("latent_edits_kernel");
  return 0;
}

// train.py:357-361 'example loss': out[0] = sqrt(mean((a - b)^2)) over n elements; one block (n is one image).
__global__ void __launch_bounds__(1024) rmse_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                                    float* __restrict__ out) {
  TraceScope trace(14);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = __ldg(a + i) - __ldg(b + i);
    acc = fmaf(d, d, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = sqrtf(v / (float)n);
  }
  trace.end();
}
int rmse(const float* a, const float* b, long long n, float* out, cudaStream_t st) {
  if (n < 1) {
    set_error("rmse: empty input");
    return 1;
  }
  launch_k(rmse_kernel, dim3(1), dim3(1024), 0, st, a, b, n, out);
  GCT2_CHECK_LAUNCH("rmse_kernel");
  return 0;
}

}  // namespace gct2
