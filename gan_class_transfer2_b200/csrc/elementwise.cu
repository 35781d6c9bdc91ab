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

static int g_ew_sms = 148;
void elementwise_set_sms(int n) { g_ew_sms = n; }

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
  noise_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(eps), t_int,
                                      reinterpret_cast<float4*>(noised), B, vec, steps);
  GCT2_CHECK_LAUNCH("noise_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ down0 (Cin = 3)
// Direct conv on CUDA cores: K = 48 is too thin for a tensor-core tile and the layer is bound by its 128-channel
// output write.  Block = 8x8 output pixels, thread = output channel; the 18x18x3 input patch sits in smem and the
// thread's 48 weights in registers.
constexpr int C3_T = 8;                 // output tile edge
constexpr int C3_P = 2 * C3_T + 2;      // input patch edge (18)
constexpr int C3_ROW = C3_P * 3 + 2;    // padded patch row (56 floats, 16-byte aligned rows)

__device__ __forceinline__ void c3_load_patch(float (*patch)[C3_ROW], const float* __restrict__ x, int b, int oy0,
                                               int ox0, int H, int W) {
  for (int i = threadIdx.x; i < C3_P * C3_P * 3; i += blockDim.x) {
    const int c = i % 3, xx = (i / 3) % C3_P, yy = i / (3 * C3_P);
    const int iy = 2 * oy0 - 1 + yy, ix = 2 * ox0 - 1 + xx;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + (((long long)b * H + iy) * W + ix) * 3 + c);
    patch[yy][xx * 3 + c] = v;
  }
}

__global__ void __launch_bounds__(128) conv_c3_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias,
                                                            __nv_bfloat16* __restrict__ y, int ldy, int B, int H,
                                                            int W, int Cout) {
  __shared__ __align__(16) float patch[C3_P][C3_ROW];
  const int Ho = H / 2, Wo = W / 2;
  const int tilesX = Wo / C3_T, tilesY = Ho / C3_T;
  const int tile = blockIdx.x;
  const int b = tile / (tilesX * tilesY);
  const int oy0 = ((tile / tilesX) % tilesY) * C3_T, ox0 = (tile % tilesX) * C3_T;
  const int co = blockIdx.y * blockDim.x + threadIdx.x;
  c3_load_patch(patch, x, b, oy0, ox0, H, W);
  float wr[48];
#pragma unroll
  for (int k = 0; k < 48; ++k) wr[k] = __ldg(w + k * Cout + co);  // HWIO: ((ky*4+kx)*3+c)*Cout + co
  const float bv = __ldg(bias + co);
  __syncthreads();
  for (int py = 0; py < C3_T; ++py) {
#pragma unroll
    for (int px = 0; px < C3_T; ++px) {
      float acc = bv;
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const float* row = &patch[2 * py + ky][2 * px * 3];
#pragma unroll
        for (int j = 0; j < 12; ++j) acc = fmaf(row[j], wr[ky * 12 + j], acc);
      }
      const long long pix = ((long long)b * Ho + oy0 + py) * Wo + ox0 + px;
      y[pix * ldy + co] = __float2bfloat16(fmaxf(acc, 0.f));
    }
  }
}

int conv4s2_c3_fprop(const float* x, const float* w, const float* bias, __nv_bfloat16* y, int ldy, int B, int H,
                     int W, int Cout, cudaStream_t st) {
  if ((H / 2) % C3_T || (W / 2) % C3_T || Cout % 128) {
    set_error("conv4s2_c3_fprop: unsupported shape H=%d W=%d Cout=%d", H, W, Cout);
    return 1;
  }
  dim3 grid(B * (H / 2 / C3_T) * (W / 2 / C3_T), Cout / 128);
  conv_c3_fprop_kernel<<<grid, 128, 0, st>>>(x, w, bias, y, ldy, B, H, W, Cout);
  GCT2_CHECK_LAUNCH("conv_c3_fprop_kernel");
  return 0;
}

// dW[ky,kx,c,co] = sum_pix x[pix@tap, c] * dz[pix, co];  db[co] = sum_pix dz[pix, co]
__global__ void __launch_bounds__(128) conv_c3_wgrad_kernel(const float* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ dz, int lddz,
                                                            float* __restrict__ dw, float* __restrict__ db, int B,
                                                            int H, int W, int Cout, int numTiles) {
  __shared__ __align__(16) float patch[C3_P][C3_ROW];
  const int Ho = H / 2, Wo = W / 2;
  const int tilesX = Wo / C3_T, tilesY = Ho / C3_T;
  const int co = blockIdx.y * blockDim.x + threadIdx.x;
  float acc[48];
#pragma unroll
  for (int k = 0; k < 48; ++k) acc[k] = 0.f;
  float accb = 0.f;
  for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
    const int b = tile / (tilesX * tilesY);
    const int oy0 = ((tile / tilesX) % tilesY) * C3_T, ox0 = (tile % tilesX) * C3_T;
    __syncthreads();
    c3_load_patch(patch, x, b, oy0, ox0, H, W);
    __syncthreads();
    for (int py = 0; py < C3_T; ++py) {
#pragma unroll
      for (int px = 0; px < C3_T; ++px) {
        const long long pix = ((long long)b * Ho + oy0 + py) * Wo + ox0 + px;
        const float g = __bfloat162float(dz[pix * lddz + co]);
        accb += g;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          const float* row = &patch[2 * py + ky][2 * px * 3];
#pragma unroll
          for (int j = 0; j < 12; ++j) acc[ky * 12 + j] = fmaf(row[j], g, acc[ky * 12 + j]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 48; ++k) atomicAdd(dw + k * Cout + co, acc[k]);
  atomicAdd(db + co, accb);
}

int conv4s2_c3_wgrad(const float* x, const __nv_bfloat16* dz, int lddz, float* dw, float* db, int B, int H, int W,
                     int Cout, cudaStream_t st) {
  if ((H / 2) % C3_T || (W / 2) % C3_T || Cout % 128) {
    set_error("conv4s2_c3_wgrad: unsupported shape H=%d W=%d Cout=%d", H, W, Cout);
    return 1;
  }
  cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)48 * Cout * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(db, 0, (size_t)Cout * sizeof(float), st);
  if (e != cudaSuccess) {
    set_error("conv4s2_c3_wgrad memset: %s", cudaGetErrorString(e));
    return 1;
  }
  const int numTiles = B * (H / 2 / C3_T) * (W / 2 / C3_T);
  int gx = numTiles < 2 * g_ew_sms ? numTiles : 2 * g_ew_sms;
  dim3 grid(gx, Cout / 128);
  conv_c3_wgrad_kernel<<<grid, 128, 0, st>>>(x, dz, lddz, dw, db, B, H, W, Cout, numTiles);
  GCT2_CHECK_LAUNCH("conv_c3_wgrad_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ Dense(3) + MSE (a6,a7)
// One warp walks pixels; lane l owns channels 2l, 2l+1 of the 64-channel up0 output, the 3 image channels are
// handled redundantly by every lane.  Emits pred (optional), the loss partial, du0 = relu'(u0) * (dpred . Wd^T)
// as bf16, and the Dense weight/bias gradients (warp-shuffle reductions, one atomic per warp per value).
//   pred = [u0 | noised] . Wd + bd ;  loss = mean((x - pred)^2) ;  dpred = 2 (pred - x) / Ntot
__global__ void __launch_bounds__(256) dense_mse_kernel(const __nv_bfloat16* __restrict__ u0, int ldu,
                                                        const float* __restrict__ noised,
                                                        const float* __restrict__ x, const float* __restrict__ wd,
                                                        const float* __restrict__ bd, float* __restrict__ pred,
                                                        float* __restrict__ loss, __nv_bfloat16* __restrict__ du0,
                                                        int lddu, float* __restrict__ dwd, float* __restrict__ dbd,
                                                        long long pixels, float invN, int backward) {
  const int lane = threadIdx.x & 31;
  const int warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int numWarps = (gridDim.x * blockDim.x) >> 5;
  float w0[3], w1[3], wn[9], bv[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    w0[j] = __ldg(wd + (2 * lane) * 3 + j);
    w1[j] = __ldg(wd + (2 * lane + 1) * 3 + j);
    bv[j] = __ldg(bd + j);
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) wn[k] = __ldg(wd + 64 * 3 + k);
  float g0[3] = {0.f, 0.f, 0.f}, g1[3] = {0.f, 0.f, 0.f}, gn[9], gb[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 9; ++k) gn[k] = 0.f;
  float lossAcc = 0.f;
  for (long long p = warpGlobal; p < pixels; p += numWarps) {
    const uint32_t uv = __ldg(reinterpret_cast<const uint32_t*>(u0 + p * ldu) + lane);
    const float a0 = bf16_lo(uv), a1 = bf16_hi(uv);
    float s[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) s[j] = warp_sum(a0 * w0[j] + a1 * w1[j]);
    float nz[3], xv[3], d[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      nz[c] = __ldg(noised + p * 3 + c);
      xv[c] = __ldg(x + p * 3 + c);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float pj = s[j] + nz[0] * wn[j] + nz[1] * wn[3 + j] + nz[2] * wn[6 + j] + bv[j];
      if (pred != nullptr && lane == j) pred[p * 3 + j] = pj;
      const float diff = pj - xv[j];
      lossAcc += diff * diff;
      d[j] = 2.f * diff * invN;
    }
    if (backward) {
      const float r0 = a0 > 0.f ? d[0] * w0[0] + d[1] * w0[1] + d[2] * w0[2] : 0.f;
      const float r1 = a1 > 0.f ? d[0] * w1[0] + d[1] * w1[1] + d[2] * w1[2] : 0.f;
      reinterpret_cast<uint32_t*>(du0 + p * lddu)[lane] = pack_bf16x2(r0, r1);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        g0[j] = fmaf(a0, d[j], g0[j]);
        g1[j] = fmaf(a1, d[j], g1[j]);
        gb[j] += d[j];
#pragma unroll
        for (int c = 0; c < 3; ++c) gn[c * 3 + j] = fmaf(nz[c], d[j], gn[c * 3 + j]);
      }
    }
  }
  // every lane accumulated the same lossAcc / gn / gb (redundant work), so lane 0 publishes them.
  if (lane == 0) atomicAdd(loss, lossAcc * invN);
  if (backward) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      atomicAdd(dwd + (2 * lane) * 3 + j, g0[j]);
      atomicAdd(dwd + (2 * lane + 1) * 3 + j, g1[j]);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) atomicAdd(dwd + 64 * 3 + k, gn[k]);
#pragma unroll
      for (int j = 0; j < 3; ++j) atomicAdd(dbd + j, gb[j]);
    }
  }
}

int dense_mse(const __nv_bfloat16* u0, int ldu, const float* noised, const float* x, const float* wd,
              const float* bd, float* pred, float* loss, __nv_bfloat16* du0, int lddu, float* dwd, float* dbd,
              long long pixels, int Cu, float invN, int backward, cudaStream_t st) {
  if (Cu != 64) {
    set_error("dense_mse: the fused kernel expects 64 up0 channels (+3 image channels), got %d", Cu);
    return 1;
  }
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), st);
  if (backward && e == cudaSuccess) e = cudaMemsetAsync(dwd, 0, 67 * 3 * sizeof(float), st);
  if (backward && e == cudaSuccess) e = cudaMemsetAsync(dbd, 0, 3 * sizeof(float), st);
  if (e != cudaSuccess) {
    set_error("dense_mse memset: %s", cudaGetErrorString(e));
    return 1;
  }
  const int blocks = g_ew_sms * 4;
  dense_mse_kernel<<<blocks, 256, 0, st>>>(u0, ldu, noised, x, wd, bd, pred, loss, du0, lddu, dwd, dbd, pixels, invN,
                                          backward);
  GCT2_CHECK_LAUNCH("dense_mse_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ bias gradient
// db[c] = sum_rows dz[row, c]; thread = channel pair, blockDim.y row groups, one atomic per block per channel.
__global__ void bias_grad_kernel(const __nv_bfloat16* __restrict__ dz, int ld, long long rows, int C,
                                 float* __restrict__ db, int rowsPerBlock) {
  extern __shared__ float red[];  // [groups][C]
  const int pairs = C / 2;
  const int groups = blockDim.x / pairs;
  const int pr = threadIdx.x % pairs, grp = threadIdx.x / pairs;
  const long long r0 = (long long)blockIdx.x * rowsPerBlock;
  long long r1 = r0 + rowsPerBlock;
  if (r1 > rows) r1 = rows;
  float s0 = 0.f, s1 = 0.f;
  if (grp < groups) {
    for (long long r = r0 + grp; r < r1; r += groups) {
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(dz + r * ld) + pr);
      s0 += bf16_lo(v);
      s1 += bf16_hi(v);
    }
    red[grp * C + 2 * pr] = s0;
    red[grp * C + 2 * pr + 1] = s1;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int g = 0; g < groups; ++g) s += red[g * C + c];
    atomicAdd(db + c, s);
  }
}

int bias_grad(const __nv_bfloat16* dz, int ld, long long rows, int C, float* db, cudaStream_t st) {
  if (C % 2 || C > 1024 || C < 2) {
    set_error("bias_grad: unsupported channel count %d", C);
    return 1;
  }
  cudaError_t e = cudaMemsetAsync(db, 0, (size_t)C * sizeof(float), st);
  if (e != cudaSuccess) {
    set_error("bias_grad memset: %s", cudaGetErrorString(e));
    return 1;
  }
  const int pairs = C / 2;
  int threads = pairs >= 256 ? pairs : (256 / pairs) * pairs;
  const int groups = threads / pairs;
  long long blocks = (rows + 63) / 64;
  if (blocks > g_ew_sms * 4) blocks = g_ew_sms * 4;
  if (blocks < 1) blocks = 1;
  const int rowsPerBlock = (int)((rows + blocks - 1) / blocks);
  blocks = (rows + rowsPerBlock - 1) / rowsPerBlock;
  bias_grad_kernel<<<(int)blocks, threads, (size_t)groups * C * sizeof(float), st>>>(dz, ld, rows, C, db,
                                                                                     rowsPerBlock);
  GCT2_CHECK_LAUNCH("bias_grad_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ Keras Adam (a9)
// state[0] = iteration count (int64, 0-based), written back incremented; alpha goes to hyper[0].
//   lr(step) = base*(step+1)/(warmup+1) while step < warmup else base            (train.py:57-65)
//   alpha = lr*sqrt(1-b2^t)/(1-b1^t), t = step+1 ; m += (g-m)(1-b1) ; v += (g*g-v)(1-b2) ; w -= alpha*m/(sqrt(v)+eps)
__global__ void adam_prepare_kernel(long long* __restrict__ iterations, float* __restrict__ hyper, float base,
                                    int warmup, float b1, float b2) {
  const long long step = *iterations;
  float lr = base;
  if (step < warmup) lr = base * (float)(step + 1) / (float)(warmup + 1);
  const float t = (float)(step + 1);
  const float b1p = powf(b1, t), b2p = powf(b2, t);
  hyper[0] = lr * sqrtf(1.f - b2p) / (1.f - b1p);
  hyper[1] = lr;
  *iterations = step + 1;
}

__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ w, float4* __restrict__ m,
                                                   float4* __restrict__ v, const float4* __restrict__ g,
                                                   uint2* __restrict__ wb, long long nvec,
                                                   const float* __restrict__ hyper, float b1, float b2, float eps,
                                                   float gscale) {
  const float alpha = __ldg(hyper);
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float4 gv = __ldg(g + i);
    gv.x *= gscale; gv.y *= gscale; gv.z *= gscale; gv.w *= gscale;
    float4 mv = m[i], vv = v[i], wv = w[i];
    mv.x += (gv.x - mv.x) * c1; mv.y += (gv.y - mv.y) * c1; mv.z += (gv.z - mv.z) * c1; mv.w += (gv.w - mv.w) * c1;
    vv.x += (gv.x * gv.x - vv.x) * c2; vv.y += (gv.y * gv.y - vv.y) * c2;
    vv.z += (gv.z * gv.z - vv.z) * c2; vv.w += (gv.w * gv.w - vv.w) * c2;
    wv.x -= alpha * mv.x / (sqrtf(vv.x) + eps); wv.y -= alpha * mv.y / (sqrtf(vv.y) + eps);
    wv.z -= alpha * mv.z / (sqrtf(vv.z) + eps); wv.w -= alpha * mv.w / (sqrtf(vv.w) + eps);
    m[i] = mv;
    v[i] = vv;
    w[i] = wv;
    uint2 o;
    o.x = pack_bf16x2(wv.x, wv.y);
    o.y = pack_bf16x2(wv.z, wv.w);
    wb[i] = o;
  }
}

int adam_keras(float* w, float* m, float* v, const float* g, __nv_bfloat16* w_bf16, long long n,
               long long* iterations, float* hyper, float base_lr, int warmup_steps, float beta1, float beta2,
               float eps, float grad_scale, cudaStream_t st) {
  if (n % 4) {
    set_error("adam_keras: parameter count must be a multiple of 4 (pad the flat buffer), got %lld", n);
    return 1;
  }
  adam_prepare_kernel<<<1, 1, 0, st>>>(iterations, hyper, base_lr, warmup_steps, beta1, beta2);
  GCT2_CHECK_LAUNCH("adam_prepare_kernel");
  const long long nvec = n / 4;
  long long blocks = (nvec + 255) / 256;
  if (blocks > g_ew_sms * 8) blocks = g_ew_sms * 8;
  adam_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<float4*>(w), reinterpret_cast<float4*>(m),
                                           reinterpret_cast<float4*>(v), reinterpret_cast<const float4*>(g),
                                           reinterpret_cast<uint2*>(w_bf16), nvec, hyper, beta1, beta2, eps,
                                           grad_scale);
  GCT2_CHECK_LAUNCH("adam_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------ fp32 -> bf16 shadow
__global__ void cast_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, long long nvec) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 s = __ldg(src + i);
    uint2 o;
    o.x = pack_bf16x2(s.x, s.y);
    o.y = pack_bf16x2(s.z, s.w);
    dst[i] = o;
  }
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
  cast_bf16_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst),
                                                nvec);
  GCT2_CHECK_LAUNCH("cast_bf16_kernel");
  return 0;
}

}  // namespace gct2
