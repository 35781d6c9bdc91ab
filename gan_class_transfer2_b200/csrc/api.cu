// extern "C" surface declared in include/gct2_b200.h: argument checking + translation to the internal launchers.
#include "../../include/gct2_b200.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_host.cuh"
#include "conv_umma.cuh"
#include "elementwise.cuh"

using namespace gct2;

namespace {
int g_force_bn = 0, g_force_splits = 0, g_sms = 0, g_policy_f16 = 0;
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline const __nv_bfloat16* CB(const uint16_t* p) { return reinterpret_cast<const __nv_bfloat16*>(p); }
inline __nv_bfloat16* MB(uint16_t* p) { return reinterpret_cast<__nv_bfloat16*>(p); }

bool pow2_ge(int v, int lo) { return v >= lo && (v & (v - 1)) == 0; }

int check_conv(const char* name, int B, int Hlo, int Wlo, int Ca, int Cb) {
  if (B < 1 || !pow2_ge(Hlo, 4) || !pow2_ge(Wlo, 4)) {
    set_error("%s: need B >= 1 and power-of-two lo-res extent >= 4 (got B=%d, %dx%d)", name, B, Hlo, Wlo);
    return 1;
  }
  if (Ca % 64 || Cb % 64 || Ca < 64 || Cb < 64) {
    set_error("%s: channel counts must be positive multiples of 64 (got %d, %d)", name, Ca, Cb);
    return 1;
  }
  return 0;
}
ConvArgs blank(int mode, int B, int Hlo, int Wlo) {
  ConvArgs a{};
  a.mode = mode;
  a.B = B;
  a.Hlo = Hlo;
  a.Wlo = Wlo;
  a.forceBN = g_force_bn;
  a.forceSplits = g_force_splits;
  a.f16 = g_policy_f16;
  return a;
}
}  // namespace

extern "C" {

int gct2_abi_version(void) { return GCT2_ABI_VERSION; }
const char* gct2_last_error(void) { return last_error(); }

int gct2_init(int device) {
  if (conv_init(device)) return 1;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    set_error("cudaGetDeviceProperties failed");
    return 1;
  }
  g_sms = prop.multiProcessorCount;
  elementwise_set_sms(g_sms);
  return 0;
}
int gct2_num_sms(void) { return g_sms; }
long long gct2_launch_count(void) { return launch_count(); }

int gct2_debug_trace(unsigned long long* host, int max_records) { return debug_read_trace(host, max_records); }
int gct2_debug_timeline(unsigned long long* host, int max_ctas) { return debug_read_timeline(host, max_ctas); }
void gct2_debug_last_plan(int* out8) { debug_last_plan(out8); }
void gct2_set_sm_budget(int sms) { conv_set_sm_budget(sms); }
void gct2_set_adam_sms(int sms) { elementwise_set_adam_sms(sms); }
void gct2_set_policy(int fp16) {
  g_policy_f16 = fp16 ? 1 : 0;
  elementwise_set_f16(g_policy_f16);
}
int gct2_get_policy(void) { return g_policy_f16; }
int gct2_loss_scale_check(const float* g, long long n, float* ls, void* stream) { return loss_scale_check(g, n, ls, S(stream)); }
int gct2_loss_scale_update(float* ls, int growth_steps, void* stream) { return loss_scale_update(ls, growth_steps, S(stream)); }

void gct2_debug_set(int key, int value) {
  if (key == 3)
    g_force_bn = value;
  else if (key == 4)
    g_force_splits = value;
  else if (key == 13 || key == 15 || key == 23 || key == 24)
    elementwise_set_debug(key, value);
  else
    conv_set_debug(key, value);
}

int gct2_noise_images(const float* x, const float* eps, const int32_t* t_int, float* noised, int B,
                      int elems_per_image, int steps, void* stream) {
  return noise_images(x, eps, t_int, noised, B, elems_per_image, steps, S(stream));
}

int gct2_conv4s2_c3_fprop(const float* x, const float* w, const float* bias, uint16_t* y, int ldy, int B, int H,
                          int W, int Cout, void* stream) {
  return conv4s2_c3_fprop(x, w, bias, MB(y), ldy, B, H, W, Cout, S(stream));
}
int gct2_conv4s2_c3_wgrad(const float* x, const uint16_t* dz, int lddz, float* dw, float* db, int B, int H, int W,
                          int Cout, int accumulate, void* stream) {
  return conv4s2_c3_wgrad(x, CB(dz), lddz, dw, db, B, H, W, Cout, accumulate ? 0 : 1, S(stream));
}

int gct2_conv4s2_fprop(const uint16_t* x, int ldx, const uint16_t* w, const float* bias, uint16_t* y, int ldy,
                       int B, int H, int W, int Cin, int Cout, float* ws, size_t ws_bytes, int flags, void* stream) {
  if (check_conv("gct2_conv4s2_fprop", B, H / 2, W / 2, Cin, Cout)) return 1;
  ConvArgs a = blank(MODE_S, B, H / 2, W / 2);
  a.hi = CB(x); a.ldHi = ldx; a.Chi = Cin;
  a.w = CB(w); a.R = Cin; a.Cc = Cout;
  a.epi = EPI_BIAS_RELU; a.out = MB(y); a.ldo = ldy; a.bias = bias;
  a.ws = ws; a.wsBytes = ws_bytes; a.flags = (flags & GCT2_WEIGHTS_STABLE) ? CONV_WEIGHTS_STABLE : 0;
  return conv_launch(a, S(stream));
}

int gct2_conv4s2_dgrad(const uint16_t* dy, int lddy, const uint16_t* w, uint16_t* dx, int lddx,
                       const uint16_t* act, int ldact, int add_old, int B, int H, int W, int Cin, int Cout,
                       float* ws, size_t ws_bytes, int flags, void* stream) {
  if (check_conv("gct2_conv4s2_dgrad", B, H / 2, W / 2, Cin, Cout)) return 1;
  ConvArgs a = blank(MODE_P, B, H / 2, W / 2);
  a.lo = CB(dy); a.ldLo = lddy; a.Clo = Cout;
  a.w = CB(w); a.R = Cin; a.Cc = Cout;
  a.epi = EPI_DGRAD; a.out = MB(dx); a.ldo = lddx; a.act = CB(act); a.ldact = ldact; a.maskN = Cin;
  a.addOld = add_old ? 1 : 0;
  a.ws = ws; a.wsBytes = ws_bytes; a.flags = (flags & GCT2_WEIGHTS_STABLE) ? CONV_WEIGHTS_STABLE : 0;
  return conv_launch(a, S(stream));
}

int gct2_conv4s2_wgrad(const uint16_t* x, int ldx, const uint16_t* dy, int lddy, float* dw, int B, int H, int W,
                       int Cin, int Cout, float* ws, size_t ws_bytes, void* stream) {
  if (check_conv("gct2_conv4s2_wgrad", B, H / 2, W / 2, Cin, Cout)) return 1;
  ConvArgs a = blank(MODE_W, B, H / 2, W / 2);
  a.hi = CB(x); a.ldHi = ldx; a.Chi = Cin;
  a.lo = CB(dy); a.ldLo = lddy; a.Clo = Cout;
  a.dw = dw;
  a.ws = ws; a.wsBytes = ws_bytes;
  return conv_launch(a, S(stream));
}

int gct2_convT4s2_fprop(const uint16_t* x, int ldx, const uint16_t* w, const float* bias, uint16_t* y, int ldy,
                        int B, int H, int W, int Cin, int Cout, float* ws, size_t ws_bytes, int flags, void* stream) {
  if (check_conv("gct2_convT4s2_fprop", B, H, W, Cin, Cout)) return 1;
  ConvArgs a = blank(MODE_P, B, H, W);
  a.lo = CB(x); a.ldLo = ldx; a.Clo = Cin;
  a.w = CB(w); a.R = Cout; a.Cc = Cin;
  a.epi = EPI_BIAS_RELU; a.out = MB(y); a.ldo = ldy; a.bias = bias;
  a.ws = ws; a.wsBytes = ws_bytes; a.flags = (flags & GCT2_WEIGHTS_STABLE) ? CONV_WEIGHTS_STABLE : 0;
  return conv_launch(a, S(stream));
}

int gct2_convT4s2_dgrad(const uint16_t* dy, int lddy, const uint16_t* w, uint16_t* dx, int lddx,
                        const uint16_t* act, int ldact, int mask_channels, int B, int H, int W, int Cin, int Cout,
                        float* ws, size_t ws_bytes, int flags, void* stream) {
  if (check_conv("gct2_convT4s2_dgrad", B, H, W, Cin, Cout)) return 1;
  if (mask_channels % 32 || mask_channels < 0 || mask_channels > Cin) {
    set_error("gct2_convT4s2_dgrad: mask_channels must be a multiple of 32 in [0, Cin] (got %d)", mask_channels);
    return 1;
  }
  ConvArgs a = blank(MODE_S, B, H, W);
  a.hi = CB(dy); a.ldHi = lddy; a.Chi = Cout;
  a.w = CB(w); a.R = Cout; a.Cc = Cin;
  a.epi = EPI_DGRAD; a.out = MB(dx); a.ldo = lddx; a.act = CB(act); a.ldact = ldact; a.maskN = mask_channels;
  a.addOld = 0;
  a.ws = ws; a.wsBytes = ws_bytes; a.flags = (flags & GCT2_WEIGHTS_STABLE) ? CONV_WEIGHTS_STABLE : 0;
  return conv_launch(a, S(stream));
}

int gct2_convT4s2_wgrad(const uint16_t* x, int ldx, const uint16_t* dy, int lddy, float* dw, int B, int H, int W,
                        int Cin, int Cout, float* ws, size_t ws_bytes, void* stream) {
  if (check_conv("gct2_convT4s2_wgrad", B, H, W, Cin, Cout)) return 1;
  ConvArgs a = blank(MODE_W, B, H, W);
  a.hi = CB(dy); a.ldHi = lddy; a.Chi = Cout;
  a.lo = CB(x); a.ldLo = ldx; a.Clo = Cin;
  a.dw = dw;
  a.ws = ws; a.wsBytes = ws_bytes;
  return conv_launch(a, S(stream));
}

int gct2_conv3s1_fprop(const uint16_t* x, int ldx, const uint16_t* w, const float* bias, uint16_t* y, int ldy, int B,
                       int H, int W, int Cin, int Cout, int ks, float* ws, size_t ws_bytes, int flags, void* stream) {
  if (check_conv("gct2_conv3s1_fprop", B, H, W, Cin, Cout)) return 1;
  ConvArgs a = blank(MODE_CF, B, H, W);
  a.ks = ks;
  a.lo = CB(x); a.ldLo = ldx; a.Clo = Cin;
  a.w = CB(w); a.R = Cin; a.Cc = Cout;
  a.epi = EPI_BIAS_RELU; a.out = MB(y); a.ldo = ldy; a.bias = bias;
  a.ws = ws; a.wsBytes = ws_bytes; a.flags = (flags & GCT2_WEIGHTS_STABLE) ? CONV_WEIGHTS_STABLE : 0;
  return conv_launch(a, S(stream));
}

int gct2_conv3s1_fprop_add(const uint16_t* x, int ldx, const uint16_t* w, const uint16_t* res, int ldres, uint16_t* y,
                           int ldy, int B, int H, int W, int Cin, int Cout, int ks, int flags, void* stream) {
  if (check_conv("gct2_conv3s1_fprop_add", B, H, W, Cin, Cout)) return 1;
  if (res == nullptr) {
    set_error("gct2_conv3s1_fprop_add: res must not be NULL");
    return 1;
  }
  ConvArgs a = blank(MODE_CF, B, H, W);
  a.ks = ks;
  a.lo = CB(x); a.ldLo = ldx; a.Clo = Cin;
  a.w = CB(w); a.R = Cin; a.Cc = Cout;
  // the data-gradient epilogue with nothing masked and the add operand taken from `res`: y = acc + res
  a.epi = EPI_DGRAD; a.out = MB(y); a.ldo = ldy; a.act = CB(res); a.ldact = ldres; a.maskN = 0; a.addOld = 1;
  a.addSrc = CB(res); a.ldAdd = ldres;
  a.forceSplits = 1;
  a.flags = (flags & GCT2_WEIGHTS_STABLE) ? CONV_WEIGHTS_STABLE : 0;
  return conv_launch(a, S(stream));
}

int gct2_conv3s1_dgrad(const uint16_t* dy, int lddy, const uint16_t* w, uint16_t* dx, int lddx, const uint16_t* act,
                       int ldact, int mask_channels, int add_old, int B, int H, int W, int Cin, int Cout, int ks,
                       float* ws, size_t ws_bytes, int flags, void* stream) {
  if (check_conv("gct2_conv3s1_dgrad", B, H, W, Cin, Cout)) return 1;
  if (mask_channels % 32 || mask_channels < 0 || mask_channels > Cin) {
    set_error("gct2_conv3s1_dgrad: mask_channels must be a multiple of 32 in [0, Cin] (got %d)", mask_channels);
    return 1;
  }
  ConvArgs a = blank(MODE_CD, B, H, W);
  a.ks = ks;
  a.lo = CB(dy); a.ldLo = lddy; a.Clo = Cout;
  a.w = CB(w); a.R = Cin; a.Cc = Cout;
  a.epi = EPI_DGRAD; a.out = MB(dx); a.ldo = lddx; a.act = CB(act); a.ldact = ldact; a.maskN = mask_channels;
  a.addOld = add_old ? 1 : 0;
  a.ws = ws; a.wsBytes = ws_bytes; a.flags = (flags & GCT2_WEIGHTS_STABLE) ? CONV_WEIGHTS_STABLE : 0;
  return conv_launch(a, S(stream));
}

int gct2_conv3s1_wgrad(const uint16_t* x, int ldx, const uint16_t* dy, int lddy, float* dw, int B, int H, int W,
                       int Cin, int Cout, int ks, float* ws, size_t ws_bytes, void* stream) {
  if (check_conv("gct2_conv3s1_wgrad", B, H, W, Cin, Cout)) return 1;
  ConvArgs a = blank(MODE_CW, B, H, W);
  a.ks = ks;
  a.hi = CB(x); a.ldHi = ldx; a.Chi = Cin;
  a.lo = CB(dy); a.ldLo = lddy; a.Clo = Cout;
  a.dw = dw;
  a.ws = ws; a.wsBytes = ws_bytes;
  return conv_launch(a, S(stream));
}

int gct2_conv3s1_c3_fprop(const float* x, const float* w, const float* bias, uint16_t* y, int ldy, int B, int H,
                          int W, int Cout, void* stream) {
  return conv3s1_c3_fprop(x, w, bias, MB(y), ldy, B, H, W, Cout, S(stream));
}
int gct2_conv3s1_c3_wgrad(const float* x, const uint16_t* dz, int lddz, float* dw, int B, int H, int W, int Cout,
                          int accumulate, void* stream) {
  return conv3s1_c3_wgrad(x, CB(dz), lddz, dw, B, H, W, Cout, accumulate ? 0 : 1, S(stream));
}

int gct2_bias_grad(const uint16_t* dz, int ld, long long rows, int C, float* db, void* stream) {
  return bias_grad(CB(dz), ld, rows, C, db, S(stream));
}
int gct2_bias_grad_multi(int n, const uint16_t* const* dz, const int* ld, const long long* rows, const int* C,
                         float* const* db, int accumulate, void* stream) {
  return bias_grad_multi(n, reinterpret_cast<const __nv_bfloat16* const*>(dz), ld, rows, C, db, accumulate ? 0 : 1,
                         S(stream));
}

int gct2_dense_mse(const uint16_t* u0, int ldu, const float* noised, const float* x, const float* wd,
                   const float* bd, float* pred, float* loss, uint16_t* du0, int lddu, float* dwd, float* dbd,
                   long long pixels, int Cu, float inv_n, int backward, int accumulate, const float* loss_scale,
                   const float* eps, const int32_t* t_int, long long pixels_per_image, int target_mode, int steps,
                   void* stream) {
  return dense_mse(CB(u0), ldu, noised, x, wd, bd, pred, loss, MB(du0), lddu, dwd, dbd, pixels, Cu, inv_n, backward,
                   accumulate ? 0 : 1, loss_scale, eps, t_int, pixels_per_image, target_mode, steps, S(stream));
}

int gct2_res0_compose(const float* wp, const float* wd, float* weff, int U, void* stream) {
  return res0_compose(wp, wd, weff, U, S(stream));
}
int gct2_res0_decompose(const float* dweff, const float* wp, const float* wd, float* dwp, float* dwd, int U, void* stream) {
  return res0_decompose(dweff, wp, wd, dwp, dwd, U, S(stream));
}

int gct2_adam_keras(float* w, float* m, float* v, const float* g, uint16_t* w_bf16, long long n,
                    long long* iterations, float* hyper, float base_lr, int warmup_steps, float beta1, float beta2,
                    float eps, float grad_scale, void* stream) {
  return adam_keras(w, m, v, g, MB(w_bf16), n, iterations, hyper, base_lr, warmup_steps, beta1, beta2, eps,
                    grad_scale, S(stream));
}

int gct2_adam_prepare(long long* iterations, float* hyper, float base_lr, int warmup_steps, float beta1, float beta2,
                      void* stream) {
  return adam_prepare(iterations, hyper, base_lr, warmup_steps, beta1, beta2, S(stream));
}
int gct2_adam_apply(float* w, float* m, float* v, const float* g, uint16_t* w_bf16, long long n, const float* hyper,
                    float beta1, float beta2, float eps, float grad_scale, long long* iterations_inc,
                    const float* loss_scale_state, void* stream) {
  return adam_apply(w, m, v, g, 0, MB(w_bf16), n, hyper, beta1, beta2, eps, grad_scale, iterations_inc, loss_scale_state,
                    S(stream));
}
int gct2_adam_apply_g16(float* w, float* m, float* v, const uint16_t* g_bf16, uint16_t* w_bf16, long long n,
                        const float* hyper, float beta1, float beta2, float eps, float grad_scale,
                        long long* iterations_inc, void* stream) {
  return adam_apply(w, m, v, g_bf16, 1, MB(w_bf16), n, hyper, beta1, beta2, eps, grad_scale, iterations_inc, nullptr,
                    S(stream));
}
int gct2_adam_apply_p2p(float* w, float* m, float* v, const uint16_t* const* g_bf16_ptrs, uint16_t* const* w16_ptrs,
                        const uint16_t* g_multicast, uint16_t* w16_multicast, int world, long long elem_offset, long long n,
                        const float* hyper, float beta1, float beta2, float eps, float grad_scale, int write_all,
                        void* stream) {
  return adam_apply_p2p(w, m, v, g_bf16_ptrs, w16_ptrs, g_multicast, w16_multicast, world, elem_offset, n, hyper, beta1,
                        beta2, eps, grad_scale, write_all, S(stream));
}
int gct2_sum_peers_f32(const float* const* src_ptrs, int world, float* out_a, long long n_a, float* out_b, long long n_b,
                       void* stream) {
  return sum_peers_f32(src_ptrs, world, out_a, n_a, out_b, n_b, S(stream));
}
int gct2_step_begin(const float* x, float* noised, float* eps_out, int32_t* t_out, int B, int elems_per_image, int steps,
                    unsigned long long seed, const long long* iterations, float* hyper, float base_lr, int warmup_steps,
                    float beta1, float beta2, float* gsmall, long long nsmall, float* loss, void* stream) {
  return step_begin(x, nullptr, nullptr, nullptr, 0, noised, eps_out, t_out, B, elems_per_image, steps, seed, iterations,
                    hyper, base_lr, warmup_steps, beta1, beta2, gsmall, nsmall, loss, S(stream));
}
int gct2_step_begin_u8(const uint8_t* img, const uint8_t* flip, float* x_out, int width, float* noised, float* eps_out,
                       int32_t* t_out, int B, int elems_per_image, int steps, unsigned long long seed,
                       const long long* iterations, float* hyper, float base_lr, int warmup_steps, float beta1,
                       float beta2, float* gsmall, long long nsmall, float* loss, void* stream) {
  if (img == nullptr) {
    set_error("gct2_step_begin_u8: img is null");
    return 1;
  }
  return step_begin(nullptr, img, flip, x_out, width, noised, eps_out, t_out, B, elems_per_image, steps, seed, iterations,
                    hyper, base_lr, warmup_steps, beta1, beta2, gsmall, nsmall, loss, S(stream));
}

int gct2_sample_update(const float* pred, float* fake, float* x_theta, float* eps_theta, int t, int t_next, int steps,
                       long long n, int target_mode, void* stream) {
  return sample_update(pred, fake, x_theta, eps_theta, t, t_next, steps, n, target_mode, S(stream));
}
int gct2_latent_edits(const float* eps_theta, const float* dictionary, float* out, int size, int entries, void* stream) {
  return latent_edits(eps_theta, dictionary, out, size, entries, S(stream));
}
int gct2_rmse(const float* a, const float* b, long long n, float* out, void* stream) { return rmse(a, b, n, out, S(stream)); }

int gct2_cast_bf16(const float* src, uint16_t* dst, long long n, void* stream) {
  return cast_bf16(src, MB(dst), n, S(stream));
}

}  // extern "C"
