// Launcher for the tcgen05 conv family: builds TMA tensor maps, picks the tile shape / split-K factor,
// and launches conv_umma_kernel (+ the split-K finishing pass).  See conv_umma.cuh for the data flow.
#include "conv_umma.cuh"
#include "conv_host.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace gct2 {

// ------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
const char* last_error() { return g_err; }
int debug_read_timeline(unsigned long long* host, int max_ctas);
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int g_use_pdl = 1;
static int g_pair = 0;                 // debug key 19: 0 = CTA pairs (cta_group::2) when a launch has more tiles than SMs,
                                       // 1 = wherever legal, 2 = never
static int g_fuse_finish = 1;          // debug key 12 != 0 disables the in-kernel split-K finish
static int g_cap_w = 0, g_cap_sp = 0;  // debug keys 9 / 10: CTA budget of wgrad / dgrad launches (0 = all SMs)
static int g_sm_budget = 0;            // key 22 / gct2_set_sm_budget: CTAs any conv launch may occupy (0 = all SMs)
static int g_b_early = 1;              // debug key 21 != 0 disables the weight fetch before griddepcontrol.wait
static int g_tap_fuse = 1;             // debug key 27 != 0 disables the four-tap weight-gradient items of 64-channel layers
static int g_no_l2_finish = 0;         // key 26 != 0: never finish split-K with the L2 rendezvous (it needs every CTA of the
                                       // launch resident at once, which nothing guarantees beside NCCL kernels); the
                                       // cluster form and the finishing kernel remain
static int g_csplit = 1;               // key 25: split-K inside a cluster (DSMEM): 0 = when the cost model picks it, 1 = never
                                       // (partials through L2; the single-GPU default: measured 2-5 % slower per step at
                                       // batch 1 -- its epilogue is 1-2 us shorter but the cluster launch costs 0.5 us
                                       // and the plans it enables use fewer CTAs), 2 = whenever legal.  Data-parallel
                                       // steps use 0: the cluster form is safe beside NCCL kernels, the L2 form is not
static unsigned g_spin_limit = 1u << 28;  // debug key 20: rendezvous watchdog in polls of ~40 ns (default ~10 s; 0 = none)
static int g_stamp_pos = 0;            // debug key 16 (-DGCT2_TIMELINE builds): see ConvParams::stampPos
static long long g_launches = 0;
void count_launch(int n) { g_launches += n; }
long long launch_count() { return g_launches; }
static int g_last_plan[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // BN, splits, pair, fused, grid, stages, rounds, early
void debug_last_plan(int* out8) {
  for (int i = 0; i < 8; ++i) out8[i] = g_last_plan[i];
}

// ------------------------------------------------------------------------------------ globals
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_mn_lbo = 8192, g_mn_sbo = 1024, g_verbose = 0;
static unsigned long long* g_dbg = nullptr;  // test hook (key 7): phase timestamps of the most recent conv launch
static int g_dbg_ctas = 0;
constexpr int DBG_MAX_CTAS = 512;

// Per-device state: the library may be initialised on several devices of one process (one engine per device); every
// launch looks its device up with cudaGetDevice().
constexpr int MAX_DEVICES = 16;
constexpr int CNT_RING_INTS = 1 << 20;  // rendezvous counters of the fused split-K finish: a ring every fused launch cuts
                                        // its own [numTiles][2] region from, so launches in flight on different streams
                                        // (a second engine, sampling beside training) never share counters
struct DeviceState {
  bool inited = false;
  int num_sms = 148;
  int* cnt = nullptr;
  size_t cnt_next = 0;
  int max_pairs[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // co-resident 2-CTA clusters [mode][BN index] (0 = unsupported)
  int max_clusters[2][3][4] = {};  // co-resident clusters of 2 / 4 / 8 plain CTAs [mode S|P][BN index][log2 size]
};
static DeviceState g_dev[MAX_DEVICES];
static DeviceState* cur_dev() {
  int d = -1;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES || !g_dev[d].inited) return nullptr;
  return &g_dev[d];
}

int debug_read_timeline(unsigned long long* host, int max_ctas) {
  if (g_dbg == nullptr) return 0;
  const int n = g_dbg_ctas < max_ctas ? g_dbg_ctas : max_ctas;
  if (cudaDeviceSynchronize() != cudaSuccess ||
      cudaMemcpy(host, g_dbg, (size_t)n * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("debug_read_timeline: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  return n;
}

static unsigned long long* g_trace_host_ptr = nullptr;
void trace_set_elementwise(unsigned long long* buf);  // elementwise.cu's copy of the pointer
int debug_read_trace(unsigned long long* host, int max_records) {
  if (g_trace_host_ptr == nullptr) return 0;
  unsigned long long n = 0;
  if (cudaDeviceSynchronize() != cudaSuccess ||
      cudaMemcpy(&n, g_trace_host_ptr, sizeof(n), cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("debug_read_trace: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  if (n > TRACE_MAX_RECORDS) n = TRACE_MAX_RECORDS;
  if ((int)n > max_records) n = max_records;
  if (cudaMemcpy(host, g_trace_host_ptr + 1, n * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemset(g_trace_host_ptr, 0, sizeof(unsigned long long)) != cudaSuccess) {
    set_error("debug_read_trace: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  return (int)n;
}

void conv_set_sm_budget(int n) { g_sm_budget = n > 0 ? n : 0; }

void conv_set_debug(int key, int value) {
  if (key == 11) {
    if (value && g_trace_host_ptr == nullptr) {
      const size_t bytes = (1 + TRACE_MAX_RECORDS * 4) * sizeof(unsigned long long);
      if (cudaMalloc(&g_trace_host_ptr, bytes) != cudaSuccess || cudaMemset(g_trace_host_ptr, 0, bytes) != cudaSuccess) {
        set_error("step trace buffer: %s", cudaGetErrorString(cudaGetLastError()));
        g_trace_host_ptr = nullptr;
        return;
      }
    }
    unsigned long long* dev = value ? g_trace_host_ptr : nullptr;
    if (cudaMemcpyToSymbol(g_trace_buf, &dev, sizeof(dev)) != cudaSuccess)
      set_error("step trace symbol: %s", cudaGetErrorString(cudaGetLastError()));
    trace_set_elementwise(dev);
  }
  if (key == 0) g_mn_lbo = value;
  if (key == 1) g_mn_sbo = value;
  if (key == 2) g_verbose = value;
  if (key == 8) g_use_pdl = value ? 0 : 1;  // key 8 != 0 disables programmatic dependent launch
  if (key == 9) g_cap_w = value;
  if (key == 10) g_cap_sp = value;
  if (key == 12) g_fuse_finish = value ? 0 : 1;
  if (key == 16) g_stamp_pos = value;
  if (key == 19) g_pair = value;
  if (key == 20) g_spin_limit = (unsigned)value;
  if (key == 21) g_b_early = value ? 0 : 1;
  if (key == 27) g_tap_fuse = value ? 0 : 1;
  if (key == 22) conv_set_sm_budget(value);
  if (key == 25) g_csplit = value;
  if (key == 26) g_no_l2_finish = value;
  if (key == 7) {
    if (value && g_dbg == nullptr &&
        cudaMalloc(&g_dbg, (size_t)DBG_MAX_CTAS * 8 * sizeof(unsigned long long)) != cudaSuccess) {
      set_error("timeline buffer: %s", cudaGetErrorString(cudaGetLastError()));
      g_dbg = nullptr;
    }
    if (!value && g_dbg != nullptr) {
      cudaFree(g_dbg);
      g_dbg = nullptr;
    }
  }
}

template <int MODE, int BN, int PAIR = 0, int CS = 0>
static int set_attr() {
  cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<MODE, BN, PAIR, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_umma_kernel<%d,%d,%d>): %s", MODE, BN, PAIR, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

static void query_all_pairs(DeviceState& ds);

int conv_init(int device) {
  if (device < 0 || device >= MAX_DEVICES) {
    set_error("gct2_init: device %d out of range", device);
    return 1;
  }
  if (g_dev[device].inited) return 0;
  int prev = -1;
  cudaGetDevice(&prev);
  struct Restore {
    int prev;
    ~Restore() {
      if (prev >= 0) cudaSetDevice(prev);  // never leave the caller on another device
    }
  } restore{prev};
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    return 1;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return 1;
  }
  if (prop.major != 10) {
    set_error("gct2 requires an sm_100a device (B200); found sm_%d%d", prop.major, prop.minor);
    return 1;
  }
  DeviceState& ds = g_dev[device];
  ds.num_sms = prop.multiProcessorCount;
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
      return 1;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  int rc = 0;  // function attributes are per device: set them on every device that is initialised
  rc |= set_attr<MODE_S, 64>() | set_attr<MODE_S, 128>() | set_attr<MODE_S, 256>();
  rc |= set_attr<MODE_P, 64>() | set_attr<MODE_P, 128>() | set_attr<MODE_P, 256>();
  rc |= set_attr<MODE_W, 64>() | set_attr<MODE_W, 128>() | set_attr<MODE_W, 256>();
  rc |= set_attr<MODE_S, 128, 1>() | set_attr<MODE_S, 256, 1>() | set_attr<MODE_P, 128, 1>() | set_attr<MODE_P, 256, 1>();
  rc |= set_attr<MODE_W, 128, 1>() | set_attr<MODE_W, 256, 1>();
  rc |= set_attr<MODE_S, 64, 0, 1>() | set_attr<MODE_S, 128, 0, 1>() | set_attr<MODE_S, 256, 0, 1>();
  rc |= set_attr<MODE_P, 64, 0, 1>() | set_attr<MODE_P, 128, 0, 1>() | set_attr<MODE_P, 256, 0, 1>();
  rc |= set_attr<MODE_CF, 64>() | set_attr<MODE_CF, 128>() | set_attr<MODE_CF, 256>();
  rc |= set_attr<MODE_CD, 64>() | set_attr<MODE_CD, 128>() | set_attr<MODE_CD, 256>();
  rc |= set_attr<MODE_CW, 64>() | set_attr<MODE_CW, 128>() | set_attr<MODE_CW, 256>();
  if (rc) return 1;
  query_all_pairs(ds);
  if (cudaMalloc(&ds.cnt, (size_t)CNT_RING_INTS * sizeof(int)) != cudaSuccess ||
      cudaMemset(ds.cnt, 0, (size_t)CNT_RING_INTS * sizeof(int)) != cudaSuccess) {
    set_error("could not allocate the split-K rendezvous counters");
    return 1;
  }
  ds.cnt_next = 0;
  ds.inited = true;
  return 0;
}

// ------------------------------------------------------------------------------------ tensor maps
static thread_local int t_map_f16 = 0;  // storage format of the launch whose tensor maps are being encoded
static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(m, t_map_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
              rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
    return 1;
  }
  return 0;
}

// hi-res NHWC tensor viewed as (px*ld + c, W/2, py, H/2, B): one box = one filter tap of a pixel tile.
static int map_hi5(CUtensorMap* m, const __nv_bfloat16* p, int ld, int C, int B, int Hhi, int Whi, int Wt, int Ht,
                   int Nb) {
  cuuint64_t dims[5] = {(cuuint64_t)(ld + C), (cuuint64_t)(Whi / 2), 2, (cuuint64_t)(Hhi / 2), (cuuint64_t)B};
  cuuint64_t st[4] = {(cuuint64_t)2 * ld * 2, (cuuint64_t)Whi * ld * 2, (cuuint64_t)2 * Whi * ld * 2,
                      (cuuint64_t)Hhi * Whi * ld * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)Wt, 1, (cuuint32_t)Ht, (cuuint32_t)Nb};
  return encode(m, p, 5, dims, st, box);
}
// lo-res NHWC tensor (C, W, H, B).
static int map_lo4(CUtensorMap* m, const __nv_bfloat16* p, int ld, int C, int B, int H, int W, int Wt, int Ht,
                   int Nb) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t st[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)Wt, (cuuint32_t)Ht, (cuuint32_t)Nb};
  return encode(m, p, 4, dims, st, box);
}
// kernel [taps][R][Cc] viewed as (Cc, R, taps).
static int map_w3(CUtensorMap* m, const __nv_bfloat16* p, int R, int Cc, int boxRows, int taps = 16) {
  cuuint64_t dims[3] = {(cuuint64_t)Cc, (cuuint64_t)R, (cuuint64_t)taps};
  cuuint64_t st[2] = {(cuuint64_t)Cc * 2, (cuuint64_t)R * Cc * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)boxRows, 1};
  return encode(m, p, 3, dims, st, box);
}

// ------------------------------------------------------------------------------------ split-K finish
// ws fp32 [splits][pixels][N] -> bf16 out with the real epilogue.  Slabs are summed in split order (deterministic).
__global__ void splitk_finish_kernel(const float* __restrict__ ws, long long splitStride, int splits, int N,
                                     long long pixels, int epi, __nv_bfloat16* __restrict__ out, int ldo,
                                     const float* __restrict__ bias, const __nv_bfloat16* __restrict__ act, int ldact,
                                     int maskN, int addOld, int f16) {
  TraceScope trace(20);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  const int vecPerRow = N / 4;
  const long long total = pixels * vecPerRow;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / vecPerRow;
    const int n = (int)(i % vecPerRow) * 4;
    const float4* wp = reinterpret_cast<const float4*>(ws + pix * N + n);
    float4 v = __ldg(wp);
    for (int sidx = 1; sidx < splits; ++sidx) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(ws + sidx * splitStride + pix * N + n));
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    __nv_bfloat16* o = out + pix * ldo + n;
    if (epi == EPI_BIAS_RELU) {
      const float4 b = *reinterpret_cast<const float4*>(bias + n);
      v.x = fmaxf(v.x + b.x, 0.f);
      v.y = fmaxf(v.y + b.y, 0.f);
      v.z = fmaxf(v.z + b.z, 0.f);
      v.w = fmaxf(v.w + b.w, 0.f);
    } else {
      if (addOld) {
        const uint2 old = *reinterpret_cast<const uint2*>(o);
        v.x += h_lo(old.x, f16);
        v.y += h_hi(old.x, f16);
        v.z += h_lo(old.y, f16);
        v.w += h_hi(old.y, f16);
      }
      if (n < maskN) {
        const uint2 a = *reinterpret_cast<const uint2*>(act + pix * ldact + n);
        v.x = h_pos_lo(a.x) ? v.x : 0.f;
        v.y = h_pos_hi(a.x) ? v.y : 0.f;
        v.z = h_pos_lo(a.y) ? v.z : 0.f;
        v.w = h_pos_hi(a.y) ? v.w : 0.f;
      }
    }
    uint2 r;
    r.x = pack_h2(v.x, v.y, f16);
    r.y = pack_h2(v.z, v.w, f16);
    *reinterpret_cast<uint2*>(o) = r;
  }
  trace.end();
}

// dw[i] = sum over splits of slab[s][i] (weight-gradient split-K), fixed order.
__global__ void wgrad_reduce_kernel(const float4* __restrict__ ws, long long splitStrideVec, int splits,
                                    float4* __restrict__ dw, long long nvec) {
  TraceScope trace(21);
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(ws + i);
    for (int sidx = 1; sidx < splits; ++sidx) {
      const float4 u = __ldg(ws + sidx * splitStrideVec + i);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    dw[i] = v;
  }
  trace.end();
}

// ------------------------------------------------------------------------------------ heuristics
// Ring: every shape fills the same 192 KB of pipeline (conv_umma.cuh: kStages / kStageBytes) -- 4 x 48 KB at BN = 64 and
// 3 x 64 KB at BN = 128 (two k-chunks per slot), 4 x 48 KB at BN = 256; a CTA of a pair stages only half of the B tile:
// 4 x 48 KB at BN = 128, 6 x 32 KB at BN = 256.  (Two- to four-slot rings that would let two CTAs share an SM were
// measured in round 1: -16 % at batch 1, -25 % at batch 32.)
// pipeline stages + barriers (256 B) + 1 KB alignment slack
static size_t smem_for(int BN, bool pair = false) {
  (void)BN;
  (void)pair;
  return (size_t)192 * 1024 + 1024 + 256;
}
static int stages_for(int BN, bool pair) {
  if (pair) return BN == 256 ? kStages<256, 1>() : kStages<128, 1>();
  return BN == 256 ? kStages<256, 0>() : (BN == 128 ? kStages<128, 0>() : kStages<64, 0>());
}
static int kps_for(int BN) { return BN == 256 ? 1 : 2; }

struct Choice {
  int BN, splits;
  int pair = 0;
  int csplit = 0;  // the splits are the CTAs of one cluster (partials through distributed shared memory)
};

static int bn_index(int BN) { return BN == 64 ? 0 : (BN == 128 ? 1 : 2); }

template <int MODE, int BN>
static void query_pairs(DeviceState& ds) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * 32);
  cfg.blockDim = dim3(kConvThreads<BN>());
  cfg.dynamicSmemBytes = smem_for(BN, true);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, conv_umma_kernel<MODE, BN, 1>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  ds.max_pairs[MODE][bn_index(BN)] = n;
}

template <int MODE, int BN>
static void query_clusters(DeviceState& ds) {
  for (int lg = 1; lg <= 3; ++lg) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((1 << lg) * 16);
    cfg.blockDim = dim3(kConvThreads<BN>());
    cfg.dynamicSmemBytes = smem_for(BN, false);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1 << lg;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, conv_umma_kernel<MODE, BN, 0, 1>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    ds.max_clusters[MODE][bn_index(BN)][lg] = n;
  }
}

static void query_all_pairs(DeviceState& ds) {
  query_clusters<MODE_S, 64>(ds); query_clusters<MODE_S, 128>(ds); query_clusters<MODE_S, 256>(ds);
  query_clusters<MODE_P, 64>(ds); query_clusters<MODE_P, 128>(ds); query_clusters<MODE_P, 256>(ds);
  query_pairs<MODE_S, 128>(ds); query_pairs<MODE_S, 256>(ds);
  query_pairs<MODE_P, 128>(ds); query_pairs<MODE_P, 256>(ds);
  query_pairs<MODE_W, 128>(ds); query_pairs<MODE_W, 256>(ds);
}

// CTAs a launch may occupy: all SMs, or the caller's budget (gct2_set_sm_budget: SMs left to a concurrently running
// NCCL kernel or optimiser launch), further capped by the per-chain test hooks.
static int cta_budget(const DeviceState& ds, int mode, bool dgradEpi) {
  int n = ds.num_sms;
  if (g_sm_budget > 0 && g_sm_budget < n) n = g_sm_budget;
  const int cap = mode == MODE_W ? g_cap_w : (dgradEpi ? g_cap_sp : 0);
  if (cap > 0 && cap < n) n = cap;
  return n;
}

// Cost model (SM cycles at ~1.97 GHz) for one launch, fitted to per-CTA phase timelines measured on B200
// (tools/timeline.py; profiles/): fixed start-up + ring rounds + epilogue, where
//   * a ring round is issue-bound (MMA issuer: wait, 4 UMMAs per k-chunk, commit; three TMA producers keep up) until all
//     SMs together exceed what L2 delivers -- wide tiles ingest fewer bytes per FLOP and win at large batch;
//   * split-K buys parallelism for the price of a partial tile's round trip through L2 and a rendezvous (fused) or an
//     extra launch (finishing kernel) -- with cheap k-steps it pays only when a layer has very few tiles;
//   * epilogues scale with the tile's real rows: deep layers at batch 1 fill 16-64 of a tile's 128 rows.
static Choice choose(const DeviceState& ds, int mode, int mTiles, int phases, int N, int kTotal, long long outElems,
                     bool dgradEpi, int forceBN, int forceSplits, size_t slabBytes, size_t wsBytes, int wTaps = 16,
                     bool allowCluster = true) {
  Choice best{0, 1};
  double bestCost = 1e30;
  const int bns[3] = {256, 128, 64};
  const bool isW = mode == MODE_W;
  const int maxCtas = cta_budget(ds, mode, dgradEpi);
  for (int bi = 0; bi < 3; ++bi) {
    const int BN = bns[bi];
    if (N % BN) continue;
    if (forceBN && BN != forceBN) continue;
    const int nTiles = N / BN;
    for (int splits = 1; splits <= 64; splits *= 2) {
      if (kTotal % splits) break;
      if (forceSplits && splits != forceSplits) continue;
      // partial slabs must fit the caller's workspace (tile-major slabs of the fused finish hold whole 128-row tiles)
      const size_t tileSlab = isW ? 0 : (size_t)mTiles * phases * 128 * N * sizeof(float);
      if (splits > 1 && (slabBytes > tileSlab ? slabBytes : tileSlab) * splits > wsBytes) break;
      const int kIters = kTotal / splits;
      const long long items = (long long)mTiles * phases * (isW ? wTaps : 1) * nTiles * splits;
      const long long active = items < maxCtas ? items : maxCtas;
      const long long waves = (items + active - 1) / active;
      // one 64-wide k-chunk: issue-bound at small grids.  Selection constants: 290 cycles with two chunks per ring slot
      // (BN <= 128), 408 + 0.96 BN with one (BN = 256), never below the tensor pipe's ~2.1 BN, and the chip-wide L2 -> SM
      // rate (6000 B/clk) when every SM pulls at once.  Round 2 also measured the steady state per CTA directly
      // (tools/timeline.py: 285 / 375 / 665 cycles at BN = 64 / 128 / 256 -- every shape moves ~170 bytes per clock through
      // shared memory, TMA writes plus MMA reads), but plans chosen with those figures were slower in the step (0.582 vs
      // 0.564 ms at batch 1): the split-K epilogue terms below are calibrated against the constants kept here.
      double tk = kps_for(BN) == 2 ? 290.0 : 408.0 + 0.96 * BN;
      if (tk < 2.1 * BN) tk = 2.1 * BN;
      const double l2 = (16384.0 + BN * 128.0) * (double)active / 6000.0;
      if (l2 > tk) tk = l2;
      const double main = kIters * tk;
      // rows of a 128-row tile that hold real pixels (deep layers at batch 1 have 16 .. 64)
      double validRows = isW ? 128.0 : (double)outElems / ((double)N * mTiles * phases);
      if (validRows > 128.0) validRows = 128.0;
      double epi;
      int useCluster = 0;
      if (splits > 1 && !isW) {
        // partial tile to its slab and back through L2 (~14.5 B/clk per CTA each way) + the rendezvous
        epi = 5000.0 + 2.0 * validRows * BN * 4.0 / 14.5;
        // finished by a separate kernel: a launch + every slab read once more, at the chip-wide rate
        if (items > maxCtas || !g_fuse_finish || g_no_l2_finish) epi += 6000.0 + (double)outElems * 4.0 * (splits + 0.5) / 3000.0;
        // the splits as one cluster: partial tile to the CTA's own shared memory, each CTA reads 128/splits rows of
        // every peer through DSMEM; no global traffic, no residency requirement beyond the cluster itself
        int lg = 0;
        while ((1 << lg) < splits) ++lg;
        const int clusters = (allowCluster && splits <= 8 && g_csplit != 1 && g_fuse_finish) ? ds.max_clusters[mode][bn_index(BN)][lg] : 0;
        if (clusters > 0) {
          const double epiC = 2500.0 + 20.0 * BN;
          const bool fits = g_sm_budget == 0 || items <= maxCtas;
          if (fits && (epiC < epi || g_csplit == 2)) {
            epi = epiC;
            useCluster = 1;
          }
        }
      } else if (isW) {
        epi = 35.0 * BN;                                // fp32 tile straight to HBM
        if (splits > 1) epi += 5000.0 + (double)outElems * 4.0 * (splits + 1.0) / 3000.0;  // + reduction kernel
      } else {
        epi = dgradEpi ? 600.0 + 30.0 * BN : 1000.0 + 15.0 * BN;
        epi *= 0.25 + 0.75 * validRows / 128.0;
      }
      long long wavesEff = waves;
      if (useCluster) {
        int lg = 0;
        while ((1 << lg) < splits) ++lg;
        long long cap = (long long)ds.max_clusters[mode][bn_index(BN)][lg] * splits;
        if (cap > maxCtas) cap = maxCtas - maxCtas % splits;
        if (cap < splits) cap = splits;
        wavesEff = (items + cap - 1) / cap;
        // a cluster launch always has one CTA per work item: under an SM budget (room left for NCCL / the optimiser) it is
        // only used when it fits the budget in one wave
      }
      // fixed per launch: prologue 0.7 us + first loads 1.45 us + teardown 0.25 us
      const double cost = 4800.0 + (useCluster ? 600.0 : 0.0) + wavesEff * main + epi * (useCluster ? wavesEff : 1) +
                          (useCluster ? 0.0 : (waves - 1) * (epi > main ? epi - main : 0.0));
      if (cost < bestCost) {
        bestCost = cost;
        best = Choice{BN, splits};
        best.csplit = useCluster;
      }
    }
  }
  return best;
}

template <int MODE, int BN, int PAIR = 0, int CS = 0>
static cudaError_t launch_one(int grid, size_t smem, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b,
                              const ConvParams& p) {
  const int clusterSize = PAIR ? 2 : (CS ? p.splits : 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kConvThreads<BN>());
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  cfg.attrs = at;
  cfg.numAttrs = 0;
  if (clusterSize > 1) {
    at[cfg.numAttrs].id = cudaLaunchAttributeClusterDimension;
    at[cfg.numAttrs].val.clusterDim.x = clusterSize;
    at[cfg.numAttrs].val.clusterDim.y = 1;
    at[cfg.numAttrs].val.clusterDim.z = 1;
    ++cfg.numAttrs;
  }
  if (g_use_pdl) {
    at[cfg.numAttrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[cfg.numAttrs].val.programmaticStreamSerializationAllowed = 1;
    ++cfg.numAttrs;
  }
  return cudaLaunchKernelEx(&cfg, conv_umma_kernel<MODE, BN, PAIR, CS>, a, b, p);
}

template <int MODE>
static cudaError_t launch_bn(int BN, int grid, size_t smem, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b,
                             const ConvParams& p) {
  if (p.cm == 2) {
    if constexpr (mode_is_s1(MODE)) {
      return cudaErrorInvalidValue;  // the stride-1 maps have no cta_group::2 instantiations
    } else {
      if (BN < 128) return cudaErrorInvalidValue;
      return BN == 128 ? launch_one<MODE, 128, 1>(grid, smem, st, a, b, p) : launch_one<MODE, 256, 1>(grid, smem, st, a, b, p);
    }
  }
  if (p.csplit) {
    if constexpr (mode_is_w(MODE) || mode_is_s1(MODE)) {
      return cudaErrorInvalidValue;
    } else {
      switch (BN) {
        case 64: return launch_one<MODE, 64, 0, 1>(grid, smem, st, a, b, p);
        case 128: return launch_one<MODE, 128, 0, 1>(grid, smem, st, a, b, p);
        default: return launch_one<MODE, 256, 0, 1>(grid, smem, st, a, b, p);
      }
    }
  }
  switch (BN) {
    case 64: return launch_one<MODE, 64>(grid, smem, st, a, b, p);
    case 128: return launch_one<MODE, 128>(grid, smem, st, a, b, p);
    default: return launch_one<MODE, 256>(grid, smem, st, a, b, p);
  }
}

static void pixel_tile(int rows, int H, int W, int* Wt, int* Ht, int* Nb) {
  int wt = W < 16 ? W : 16;
  if (rows == 64 && wt > 8) wt = 8;
  int ht = rows / wt;
  if (ht > H) ht = H;
  *Wt = wt;
  *Ht = ht;
  *Nb = rows / (wt * ht);
}

int conv_launch(const ConvArgs& a, cudaStream_t stream) {
  DeviceState* dsp = cur_dev();
  if (dsp == nullptr) {
    set_error("gct2_init was not called for the current device");
    return 1;
  }
  DeviceState& ds = *dsp;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap mapA, mapB;
  p.B = a.B;
  p.Hlo = a.Hlo;
  p.Wlo = a.Wlo;
  p.mnLbo = g_mn_lbo;
  p.mnSbo = g_mn_sbo;
  p.f16 = a.f16 ? 1 : 0;
  t_map_f16 = p.f16;
  const bool s1 = mode_is_s1(a.mode);
  if (s1 && a.ks != 1 && a.ks != 3) {
    set_error("stride-1 conv: kernel side must be 1 or 3 (got %d)", a.ks);
    return 1;
  }
  const int s1Taps = s1 ? a.ks * a.ks : 0;
  p.ks = a.ks;
  const int rows = mode_is_w(a.mode) ? 64 : 128;
  pixel_tile(rows, a.Hlo, a.Wlo, &p.Wt, &p.Ht, &p.Nb);
  if (a.Wlo % p.Wt || a.Hlo % p.Ht || p.Wt * p.Ht * p.Nb != rows) {
    set_error("unsupported spatial extent %dx%d (tile %dx%dx%d)", a.Hlo, a.Wlo, p.Nb, p.Ht, p.Wt);
    return 1;
  }
  p.tilesX = a.Wlo / p.Wt;
  p.tilesY = a.Hlo / p.Ht;
  const int tilesB = (a.B + p.Nb - 1) / p.Nb;
  const int pixTiles = p.tilesX * p.tilesY * tilesB;
  const int maxCtas = cta_budget(ds, a.mode, a.epi == EPI_DGRAD);
  int BN = 0;
  p.cm = 1;

  if (!mode_is_w(a.mode)) {
    // weight-side roles: S and CF contract over the kernel's rows (HWIO: R = Cin) and produce its columns; P and CD the
    // other way round
    const bool rowsAreK = a.mode == MODE_S || a.mode == MODE_CF;
    const int Ck = a.mode == MODE_S ? a.Chi : a.Clo;
    const int N = rowsAreK ? a.Cc : a.R;
    const int wk = rowsAreK ? a.R : a.Cc;
    if (Ck % 64 || N % 64 || wk != Ck) {
      set_error("conv: channel counts must be multiples of 64 and match the kernel (Ck=%d N=%d kernel %dx%d)", Ck, N,
                a.R, a.Cc);
      return 1;
    }
    const int taps = s1 ? s1Taps : (a.mode == MODE_S ? 16 : 4);
    const int phases = a.mode == MODE_P ? 4 : 1;
    p.kcPer = Ck / 64;
    const int kTotal = taps * p.kcPer;
    const size_t slab = (size_t)a.B * phases * a.Hlo * a.Wlo * N * sizeof(float);
    // the cost model knows two shapes of main loop: MN-major weights (S) and K-major weights (P)
    const int modelMode = rowsAreK ? MODE_S : MODE_P;
    Choice c = choose(ds, modelMode, pixTiles, phases, N, kTotal, (long long)(slab / sizeof(float)), a.epi == EPI_DGRAD,
                      a.forceBN, a.forceSplits, slab, a.ws ? a.wsBytes : 0, 16, !s1);
    if (c.BN == 0) {
      set_error("conv: no tile shape for N=%d", N);
      return 1;
    }
    // CTA pairs: consecutive M tiles of one (phase, N tile, split) share their B tile through cta_group::2
    // (measured: +4 % step throughput at 8 images/GPU, +6 % at 32, -1 % at batch 1 where a CTA owns one tile and
    // the cluster launch costs more than the shared B tile saves -- hence only for launches of more than one wave)
    const long long itemsPlain = (long long)pixTiles * (N / c.BN) * phases * c.splits;
    if (!s1 && (g_pair == 1 || (g_pair == 0 && itemsPlain > maxCtas)) && c.BN >= 128 && pixTiles % 2 == 0 &&
        ds.max_pairs[a.mode][bn_index(c.BN)] > 0) {
      c.pair = 1;
      c.csplit = 0;  // a cluster is either a cta_group::2 pair or the K slices of one tile
    }
    BN = c.BN;
    p.mTiles = pixTiles;
    p.nTiles = N / BN;
    p.splits = c.splits;
    p.kIters = kTotal / c.splits;
    p.numItems = pixTiles * p.nTiles * phases * c.splits;
    p.cm = c.pair ? 2 : 1;
    p.N = N;
    p.Hout = a.mode == MODE_P ? 2 * a.Hlo : a.Hlo;
    p.Wout = a.mode == MODE_P ? 2 * a.Wlo : a.Wlo;
    p.out = a.out;
    p.ldo = a.ldo;
    p.bias = a.bias;
    p.act = a.act;
    p.ldact = a.ldact;
    p.maskN = a.maskN;
    p.addOld = a.addOld;
    p.addSrc = s1 ? a.addSrc : nullptr;
    p.ldAdd = a.ldAdd;
    if (p.addSrc != nullptr && (c.splits != 1 || a.epi != EPI_DGRAD || !a.addOld)) {
      set_error("conv: the separate add operand needs the add epilogue without split-K");
      return 1;
    }
    p.epi = a.epi;
    p.bEarly = (g_b_early && (a.flags & CONV_WEIGHTS_STABLE)) ? 1 : 0;
    if (c.splits > 1) {
      p.epi = EPI_WS_SLAB;
      p.ws = a.ws;
      p.wsSplitStride = (long long)a.B * p.Hout * p.Wout * N;
      p.numTiles = phases * p.nTiles * pixTiles;
      // finish inside the launch when every item has its own resident CTA (so the splits of a tile can wait for each
      // other); otherwise a finishing kernel sums the slabs
      int resident = maxCtas;  // CTAs of this launch that can be on the chip at once
      if (c.pair && ds.max_pairs[a.mode][bn_index(BN)] * 2 < resident) resident = ds.max_pairs[a.mode][bn_index(BN)] * 2;
      p.csplit = c.csplit;
      p.fused = (!c.csplit && g_fuse_finish && !g_no_l2_finish && p.numItems <= resident &&
                 (size_t)p.numTiles * 2 <= (size_t)CNT_RING_INTS / 4) ? 1 : 0;
      p.realEpi = a.epi;
      if (p.fused) {
        // this launch's own counter region (zero at rest: the last split through the rendezvous re-arms it)
        const size_t need = (size_t)p.numTiles * 2;
        if (ds.cnt_next + need > (size_t)CNT_RING_INTS) ds.cnt_next = 0;
        p.cnt = ds.cnt + ds.cnt_next;
        ds.cnt_next += need;
        p.spinLimit = g_spin_limit;
      }
    }
    if (a.mode == MODE_S) {
      p.ldG = a.ldHi;
      if (map_hi5(&mapA, a.hi, a.ldHi, a.Chi, a.B, 2 * a.Hlo, 2 * a.Wlo, p.Wt, p.Ht, p.Nb)) return 1;
      if (map_w3(&mapB, a.w, a.R, a.Cc, 64)) return 1;
    } else if (s1) {
      if (map_lo4(&mapA, a.lo, a.ldLo, a.Clo, a.B, a.Hlo, a.Wlo, p.Wt, p.Ht, p.Nb)) return 1;
      if (map_w3(&mapB, a.w, a.R, a.Cc, a.mode == MODE_CF ? 64 : BN, taps)) return 1;
    } else {
      if (map_lo4(&mapA, a.lo, a.ldLo, a.Clo, a.B, a.Hlo, a.Wlo, p.Wt, p.Ht, p.Nb)) return 1;
      if (map_w3(&mapB, a.w, a.R, a.Cc, c.pair ? BN / 2 : BN)) return 1;
    }
  } else {
    // wgrad: dw[tap][Chi][Clo]; the M side must be a multiple of 128
    p.gIsA = (a.Chi % 128 == 0) ? 1 : 0;
    const int Mch = p.gIsA ? a.Chi : a.Clo;
    const int Nch = p.gIsA ? a.Clo : a.Chi;
    if (Mch % 128 || Nch % 64) {
      set_error("wgrad: unsupported channel counts (%d, %d)", a.Chi, a.Clo);
      return 1;
    }
    const int chunks = pixTiles;
    const int wTaps = s1 ? s1Taps : 16;
    const size_t slab = (size_t)wTaps * a.Chi * a.Clo * sizeof(float);
    // A 64-channel gathered N side (up0: 64 output channels) would run on 128 x 64 tiles, the shape that spends most
    // shared-memory traffic per FLOP (ncu, 32 images: 36-43 % tensor-pipe active).  The un-shifted operand is the same for
    // every filter tap, so one work item takes FOUR taps: a 256-wide tile whose 64-column blocks are the gathered operand
    // at taps 4 ph .. 4 ph + 3 -- one A load per four B loads, the BN = 256 main loop.
    const bool tapFuse = g_tap_fuse && !s1 && !p.gIsA && Nch == 64 && (a.forceBN == 0 || a.forceBN == 256);
    Choice c = choose(ds, MODE_W, Mch / 128, 1, tapFuse ? 256 : Nch, chunks, (long long)(slab / sizeof(float)), false,
                      tapFuse ? 256 : a.forceBN, a.forceSplits, slab, a.ws ? a.wsBytes : 0, tapFuse ? 4 : wTaps);
    if (c.BN == 0) {
      set_error("wgrad: no tile shape for N=%d", Nch);
      return 1;
    }
    const long long itemsPlainW = (long long)wTaps * (Mch / 128) * (Nch / c.BN) * c.splits;
    if (!s1 && !tapFuse && (g_pair == 1 || (g_pair == 0 && itemsPlainW > maxCtas)) && c.BN >= 128 && (Mch / 128) % 2 == 0 &&
        ds.max_pairs[MODE_W][bn_index(c.BN)] > 0)
      c.pair = 1;
    BN = c.BN;
    p.mTiles = Mch / 128;
    p.nTiles = tapFuse ? 1 : Nch / BN;
    p.tapFuse = tapFuse ? 4 : 0;
    p.splits = c.splits;
    p.kIters = chunks / c.splits;
    p.numItems = (tapFuse ? wTaps / 4 : wTaps) * p.mTiles * p.nTiles * c.splits;
    p.cm = c.pair ? 2 : 1;
    p.N = Nch;
    p.ldG = a.ldHi;
    p.epi = EPI_WGRAD;
    p.dw = a.dw;
    p.tapStride = (long long)a.Chi * a.Clo;
    p.rowStride = p.gIsA ? a.Clo : 1;
    p.colStride = p.gIsA ? 1 : a.Clo;
    p.atomic = c.splits > 1;
    CUtensorMap mg, mp;
    if (s1) {  // the gathered operand is the layer's input at the same extent as dy
      if (map_lo4(&mg, a.hi, a.ldHi, a.Chi, a.B, a.Hlo, a.Wlo, p.Wt, p.Ht, p.Nb)) return 1;
    } else if (map_hi5(&mg, a.hi, a.ldHi, a.Chi, a.B, 2 * a.Hlo, 2 * a.Wlo, p.Wt, p.Ht, p.Nb)) {
      return 1;
    }
    if (map_lo4(&mp, a.lo, a.ldLo, a.Clo, a.B, a.Hlo, a.Wlo, p.Wt, p.Ht, p.Nb)) return 1;
    mapA = p.gIsA ? mg : mp;
    mapB = p.gIsA ? mp : mg;
    if (p.atomic) {
      p.ws = a.ws;
      p.wsSplitStride = (long long)wTaps * a.Chi * a.Clo;
    }
  }

  const int kps = kps_for(BN);
  p.rounds = (p.kIters + kps - 1) / kps;
  p.stampPos = g_stamp_pos;
  {
    auto lg2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
    if ((p.Wt & (p.Wt - 1)) || (p.Ht & (p.Ht - 1)) || (p.tilesX & (p.tilesX - 1)) || (p.tilesY & (p.tilesY - 1))) {
      set_error("conv: pixel-tile geometry must be power-of-two (tile %dx%d, tiles %dx%d)", p.Ht, p.Wt, p.tilesY, p.tilesX);
      return 1;
    }
    p.lgWt = lg2(p.Wt); p.lgHt = lg2(p.Ht); p.lgTilesX = lg2(p.tilesX); p.lgTilesY = lg2(p.tilesY);
    p.fdMTilesC = make_fastdiv((uint32_t)(p.mTiles / p.cm));
    p.fdNTiles = make_fastdiv((uint32_t)p.nTiles);
    p.fdSplits = make_fastdiv((uint32_t)p.splits);
    p.fdKcPer = make_fastdiv((uint32_t)(p.kcPer > 0 ? p.kcPer : 1));
  }
  const size_t smem = smem_for(BN, p.cm == 2);
  p.numClusterItems = p.csplit ? p.numItems / p.splits : p.numItems / p.cm;
  int ctas = maxCtas;  // one CTA per SM (194 KB of shared memory, 61 K registers)
  if (p.cm == 2) {
    const int mp = ds.max_pairs[a.mode][bn_index(BN)] * 2;
    if (mp < ctas) ctas = mp;
  }
  int grid = p.numItems < ctas ? p.numItems : ctas;
  grid -= grid % p.cm;
  // cluster split-K: exactly one tile per cluster (its partial tile lives in the ring's shared memory, so a CTA must
  // not start loading another item); clusters beyond the chip's capacity simply run in waves
  if (p.csplit) grid = p.numItems;
  g_last_plan[0] = BN; g_last_plan[1] = p.splits; g_last_plan[2] = p.cm == 2; g_last_plan[3] = p.csplit ? 2 : p.fused;
  g_last_plan[4] = grid; g_last_plan[5] = stages_for(BN, p.cm == 2); g_last_plan[6] = p.rounds; g_last_plan[7] = p.bEarly;
  if (g_verbose)
    fprintf(stderr,
            "gct2 conv mode %d B %d lo %dx%d tile %dx%dx%d BN %d splits %d pair %d kIters %d rounds %d items %d grid %d "
            "stages %d finish %s early %d\n",
            a.mode, a.B, a.Hlo, a.Wlo, p.Nb, p.Ht, p.Wt, BN, p.splits, p.cm == 2, p.kIters, p.rounds, p.numItems, grid,
            stages_for(BN, p.cm == 2), p.csplit ? "cluster" : (p.fused ? "l2" : (p.splits > 1 ? "kernel" : "-")), p.bEarly);
#ifdef GCT2_TIMELINE
  if (g_dbg != nullptr && grid <= DBG_MAX_CTAS) {
    cudaMemsetAsync(g_dbg, 0, (size_t)grid * 8 * sizeof(unsigned long long), stream);
    p.dbg = g_dbg;
    g_dbg_ctas = grid;
  }
#endif
  cudaError_t e;
  if (a.mode == MODE_S)
    e = launch_bn<MODE_S>(BN, grid, smem, stream, mapA, mapB, p);
  else if (a.mode == MODE_P)
    e = launch_bn<MODE_P>(BN, grid, smem, stream, mapA, mapB, p);
  else if (a.mode == MODE_W)
    e = launch_bn<MODE_W>(BN, grid, smem, stream, mapA, mapB, p);
  else if (a.mode == MODE_CF)
    e = launch_bn<MODE_CF>(BN, grid, smem, stream, mapA, mapB, p);
  else if (a.mode == MODE_CD)
    e = launch_bn<MODE_CD>(BN, grid, smem, stream, mapA, mapB, p);
  else
    e = launch_bn<MODE_CW>(BN, grid, smem, stream, mapA, mapB, p);
  if (e != cudaSuccess) {
    set_error("conv_umma_kernel launch: %s", cudaGetErrorString(e));
    return 1;
  }
  count_launch();
  if (!mode_is_w(a.mode) && p.splits > 1 && !p.fused && !p.csplit) {
    const long long pixels = (long long)a.B * p.Hout * p.Wout;
    const long long total = pixels * (p.N / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > ds.num_sms * 8) blocks = ds.num_sms * 8;
    e = launch_k(splitk_finish_kernel, dim3(blocks), dim3(256), 0, stream, a.ws, p.wsSplitStride, p.splits, p.N, pixels,
                 a.epi, a.out, a.ldo, a.bias, a.act, a.ldact, a.maskN, a.addOld, a.f16);
    if (e != cudaSuccess) {
      set_error("splitk_finish_kernel launch: %s", cudaGetErrorString(e));
      return 1;
    }
    count_launch();
  }
  if (mode_is_w(a.mode) && p.splits > 1) {
    const long long nvec = (long long)(s1 ? s1Taps : 16) * a.Chi * a.Clo / 4;
    int blocks = (int)((nvec + 255) / 256);
    if (blocks > ds.num_sms * 8) blocks = ds.num_sms * 8;
    e = launch_k(wgrad_reduce_kernel, dim3(blocks), dim3(256), 0, stream, reinterpret_cast<const float4*>(a.ws),
                 p.wsSplitStride / 4, p.splits, reinterpret_cast<float4*>(a.dw), nvec);
    if (e != cudaSuccess) {
      set_error("wgrad_reduce_kernel launch: %s", cudaGetErrorString(e));
      return 1;
    }
    count_launch();
  }
  return 0;
}

}  // namespace gct2
