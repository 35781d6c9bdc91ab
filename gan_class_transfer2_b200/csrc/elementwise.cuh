// Internal declarations of the HBM-bound kernels' launchers (definitions in elementwise.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace gct2 {

void elementwise_set_sms(int n);
void elementwise_set_debug(int key, int value);
void elementwise_set_adam_sms(int n);  // see gct2_set_adam_sms
int noise_images(const float* x, const float* eps, const int* t_int, float* noised, int B, int elemsPerImage,
                 int steps, cudaStream_t st);
int conv4s2_c3_fprop(const float* x, const float* w, const float* bias, __nv_bfloat16* y, int ldy, int B, int H,
                     int W, int Cout, cudaStream_t st);
int conv4s2_c3_wgrad(const float* x, const __nv_bfloat16* dz, int lddz, float* dw, float* db, int B, int H, int W,
                     int Cout, int zero, cudaStream_t st);
int conv3s1_c3_fprop(const float* x, const float* w, const float* bias, __nv_bfloat16* y, int ldy, int B, int H,
                     int W, int Cout, cudaStream_t st);
int conv3s1_c3_wgrad(const float* x, const __nv_bfloat16* dz, int lddz, float* dw, int B, int H, int W, int Cout,
                     int zero, cudaStream_t st);
int res0_compose(const float* wp, const float* wd, float* weff, int U, cudaStream_t st);
int res0_decompose(const float* dweff, const float* wp, const float* wd, float* dwp, float* dwd, int U, cudaStream_t st);
int dense_mse(const __nv_bfloat16* u0, int ldu, const float* noised, const float* x, const float* wd,
              const float* bd, float* pred, float* loss, __nv_bfloat16* du0, int lddu, float* dwd, float* dbd,
              long long pixels, int Cu, float invN, int backward, int zero, const float* loss_scale, const float* eps,
              const int* t_int, long long pixels_per_image, int target_mode, int steps, cudaStream_t st);
int latent_edits(const float* eps_theta, const float* dictionary, float* out, int S, int K, cudaStream_t st);
int rmse(const float* a, const float* b, long long n, float* out, cudaStream_t st);
int bias_grad_multi(int n, const __nv_bfloat16* const* dz, const int* ld, const long long* rows, const int* C,
                    float* const* db, int zero, cudaStream_t st);
int bias_grad(const __nv_bfloat16* dz, int ld, long long rows, int C, float* db, cudaStream_t st);
int adam_keras(float* w, float* m, float* v, const float* g, __nv_bfloat16* w_bf16, long long n,
               long long* iterations, float* hyper, float base_lr, int warmup_steps, float beta1, float beta2,
               float eps, float grad_scale, cudaStream_t st);
int adam_prepare(long long* iterations, float* hyper, float base_lr, int warmup_steps, float beta1, float beta2,
                 cudaStream_t st);
int adam_apply(float* w, float* m, float* v, const void* g, int g_is_bf16, __nv_bfloat16* w_bf16, long long n,
               const float* hyper, float beta1, float beta2, float eps, float grad_scale, long long* iterations_inc,
               const float* ls, cudaStream_t st);
int adam_apply_p2p(float* w, float* m, float* v, const uint16_t* const* g_ptrs, uint16_t* const* w16_ptrs,
                   const uint16_t* g_mc, uint16_t* w16_mc, int world, long long elem_offset, long long n,
                   const float* hyper, float beta1, float beta2, float eps, float grad_scale, int write_all,
                   cudaStream_t st);
int sum_peers_f32(const float* const* src_ptrs, int world, float* out_a, long long n_a, float* out_b, long long n_b,
                  cudaStream_t st);
int loss_scale_check(const float* g, long long n, float* ls, cudaStream_t st);
int loss_scale_update(float* ls, int growth_steps, cudaStream_t st);
void elementwise_set_f16(int f16);
int sample_update(const float* pred, float* fake, float* x_theta, float* eps_theta, int t, int t_next, int steps,
                  long long n, int mode, cudaStream_t st);
int step_begin(const float* x, const uint8_t* x_u8, const uint8_t* flip, float* x_out, int W, float* noised,
               float* eps_out, int* t_out, int B, int elemsPerImage, int steps,
               unsigned long long seed, const long long* iterations, float* hyper, float base_lr, int warmup_steps,
               float beta1, float beta2, float* gsmall, long long nsmall, float* loss, cudaStream_t st);
int cast_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t st);

}  // namespace gct2
