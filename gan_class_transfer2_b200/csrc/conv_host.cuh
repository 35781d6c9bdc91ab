// Host-side description of one conv-family contraction and its launcher (see conv_umma.cuh).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace gct2 {

struct ConvArgs {
  int mode;                      // MODE_S / MODE_P / MODE_W / MODE_CF / MODE_CD / MODE_CW
  int B, Hlo, Wlo;               // lo-res spatial extent; the hi-res side is (2*Hlo, 2*Wlo)
  const __nv_bfloat16* hi;       // hi-res operand (S: A gather; W: G), base already offset to its first channel
  int ldHi, Chi;                 // pixel stride (elements) and number of channels used
  const __nv_bfloat16* lo;       // lo-res operand (P: A; W: P)
  int ldLo, Clo;
  const __nv_bfloat16* w;        // S/P: bf16 kernel, flat [16][R][Cc] in the Keras layout (HWIO or HWOI)
  int R, Cc;
  // epilogue (S/P)
  int epi;                       // EPI_BIAS_RELU or EPI_DGRAD
  __nv_bfloat16* out;
  int ldo;
  const float* bias;
  const __nv_bfloat16* act;
  int ldact, maskN, addOld;
  const __nv_bfloat16* addSrc;   // stride-1 modes: tensor added instead of out's old contents (needs addOld, no split-K)
  int ldAdd;
  float* ws;                     // fp32 split-K workspace (contents irrelevant on entry and on exit)
  size_t wsBytes;
  int flags;                     // CONV_WEIGHTS_STABLE: `w` is not being written by any launch that may still be running
                                 // when this one starts, so its boxes may be fetched before griddepcontrol.wait
  // W
  float* dw;                     // fp32 [16][Chi][Clo]
  int ks;                        // stride-1 modes (MODE_CF / MODE_CD / MODE_CW): kernel side, 3 or 1; the single-resolution
                                 // activation operand goes in `lo` (CF: x, CD: dy, CW: dy) and CW's gathered x in `hi`
  int f16;                       // 16-bit storage format of every bf16-typed pointer above: 0 = bf16, 1 = fp16
  // tuning overrides (0 = heuristic)
  int forceBN, forceSplits;
};
enum : int { CONV_WEIGHTS_STABLE = 1 };

extern int g_use_pdl;  // 1 = launch with programmatic stream serialisation (debug key 8 toggles)

// <<<>>> replacement that adds the programmatic-dependent-launch attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int conv_init(int device);                       // once per process/device
int conv_launch(const ConvArgs& a, cudaStream_t stream);
int debug_read_timeline(unsigned long long* host, int max_ctas);  // test hook, see gct2_debug_timeline
int debug_read_trace(unsigned long long* host, int max_records);   // test hook, see gct2_debug_trace
void conv_set_debug(int key, int value);         // test hooks, see gct2_debug_set in include/gct2_b200.h
void conv_set_sm_budget(int n);                  // CTAs a conv launch may occupy (0 = all SMs)
void debug_last_plan(int* out8);                 // test hook, see gct2_debug_last_plan
const char* last_error();
void count_launch(int n = 1);                   // kernels/memsets enqueued by this library (gct2_launch_count)
long long launch_count();
void set_error(const char* fmt, ...);

}  // namespace gct2
