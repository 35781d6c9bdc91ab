// Microbenchmark (diagnostic, not product code): how fast can ONE SM -- and all SMs together -- pull operand boxes
// through TMA into shared memory, as a function of the box shape / tensor-map rank / ring depth / CTA count?
// The conv main loop is bound by operand ingest (tools/timeline.py), so this is its roofline.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_ingest_probe tma_ingest_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (++spins > (1u << 26)) { printf("probe: mbarrier watchdog\n"); __trap(); }
  }
}
__device__ __forceinline__ void tma2(void* d, const CUtensorMap* m, uint64_t* b, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(d)), "l"((uint64_t)m), "r"(smem_u32(b)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma3(void* d, const CUtensorMap* m, uint64_t* b, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(d)), "l"((uint64_t)m), "r"(smem_u32(b)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma4(void* d, const CUtensorMap* m, uint64_t* b, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(d)), "l"((uint64_t)m), "r"(smem_u32(b)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma5(void* d, const CUtensorMap* m, uint64_t* b, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(d)), "l"((uint64_t)m), "r"(smem_u32(b)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// variant: 0 = 2-D rows x 64 box of a [rows][C] matrix (plain GEMM operand), 1 = 4-D lo-res NHWC box (P mode A),
// 2 = 5-D parity-split hi-res box (S mode A), 3 = 3-D weight boxes 64x64 (count = boxes per stage)
struct Args { int variant, stages, iters, boxes, stageBytes; };

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap map, Args a, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + a.stages * a.stageBytes);
  uint64_t* empty = full + a.stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map) : "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = 0, t1 = 0;
  if (warp == 0 && lane == 0) {
    uint32_t st = 0, ph = 0;
    const int bid = blockIdx.x;
    for (int it = 0; it < a.iters; ++it) {
      mbar_wait(&empty[st], ph ^ 1);
      mbar_expect(&full[st], a.stageBytes);
      uint8_t* dst = smem + st * a.stageBytes;
      const int k = (it * 7 + bid * 3);
      if (a.variant == 0) {
        for (int j = 0; j < a.boxes; ++j) tma2(dst + j * 16384, &map, &full[st], (k % 8) * 64, ((k + j * 5) % 32) * 128);
      } else if (a.variant == 1) {
        for (int j = 0; j < a.boxes; ++j) tma4(dst + j * 16384, &map, &full[st], (k % 8) * 64, ((k + j) % 4) * 16 - (k & 1), ((k / 4) % 8) * 8 - ((k >> 1) & 1), 0);
      } else if (a.variant == 2) {
        for (int j = 0; j < a.boxes; ++j) tma5(dst + j * 16384, &map, &full[st], (k & 1) * 512 + ((k >> 1) % 8) * 64, ((k + j) % 2) * 16 - ((k >> 2) & 1), (k >> 3) & 1, ((k / 16) % 4) * 8 - ((k >> 4) & 1), 0);
      } else {
        for (int j = 0; j < a.boxes; ++j) tma3(dst + j * 8192, &map, &full[st], ((k + j) % 8) * 64, ((k / 8) % 8) * 64, k % 16);
      }
      if (++st == (uint32_t)a.stages) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    uint32_t st = 0, ph = 0;
    for (int it = 0; it < a.iters; ++it) {
      mbar_wait(&full[st], ph);
      if (it == a.stages) t0 = clock64();   // steady state only
      mbar_arrive(&empty[st]);
      if (++st == (uint32_t)a.stages) { st = 0; ph ^= 1; }
    }
    t1 = clock64();
    out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  __syncthreads();
}

// The real main loop's traffic: one 5-D A box (16 KB) + nb 64x64 weight boxes (8 KB each) per stage, issued by
// `producers` warps taking the k-iterations round-robin (warp p handles it = p, p + producers, ...).
__global__ void __launch_bounds__(256, 1) probe_ab(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                   int stages, int iters, int nb, int producers, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const int stageBytes = 16384 + nb * 8192;
  uint64_t* full = (uint64_t*)(smem + stages * stageBytes);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapB) : "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < producers && lane == 0) {
    const int bid = blockIdx.x;
    for (int it = warp; it < iters; it += producers) {
      const uint32_t st = it % stages, ph = (it / stages) & 1;
      mbar_wait(&empty[st], ph ^ 1);
      mbar_expect(&full[st], stageBytes);
      uint8_t* dst = smem + st * stageBytes;
      const int k = (it * 7 + bid * 3);
      tma5(dst, &mapA, &full[st], (k & 1) * 512 + ((k >> 1) % 8) * 64, (k % 2) * 16 - ((k >> 2) & 1), (k >> 3) & 1, ((k / 16) % 4) * 8 - ((k >> 4) & 1), 0);
      for (int j = 0; j < nb; ++j) tma3(dst + 16384 + j * 8192, &mapB, &full[st], ((k + j) % 8) * 64, ((k / 8) % 8) * 64, k % 16);
    }
  } else if (warp == 7 && lane == 0) {
    uint32_t st = 0, ph = 0;
    long long t0 = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[st], ph);
      if (it == 2 * stages) t0 = clock64();
      mbar_arrive(&empty[st]);
      if (++st == (uint32_t)stages) { st = 0; ph ^= 1; }
    }
    out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
  __syncthreads();
}

// Latency of ONE TMA box load (issue -> mbarrier complete), first use of the descriptor in the launch vs later uses.
__global__ void __launch_bounds__(128, 1) probe_lat(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                    int prefetchDesc, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 65536);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (prefetchDesc) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapB) : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (prefetchDesc) { long long t = clock64(); while (clock64() - t < 4000) {} }   // give the prefetch 2 us
    uint32_t ph = 0;
    for (int rep = 0; rep < 6; ++rep) {
      const long long t0 = clock64();
      mbar_expect(bar, rep < 3 ? 16384 : 8192);
      if (rep < 3) tma5(smem, &mapA, bar, (rep & 1) * 512, rep * 16, 0, 8 * rep, 0);
      else tma3(smem + 16384, &mapB, bar, 64 * rep, 64 * rep, rep);
      mbar_wait(bar, ph);
      ph ^= 1;
      out[blockIdx.x * 8 + rep] = (unsigned long long)(clock64() - t0);
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  CK(cudaSetDevice(0));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fn;
  // activations: hi-res 64x64x512 bf16 (4 MB) = lo-res view 64x64x512; matrix view [4096][512]; weights [16][512][512] (8 MB)
  __nv_bfloat16 *act, *wgt; unsigned long long* out;
  CK(cudaMalloc(&act, 64 * 64 * 512 * 2)); CK(cudaMemset(act, 0, 64 * 64 * 512 * 2));
  CK(cudaMalloc(&wgt, 16 * 512 * 512 * 2)); CK(cudaMemset(wgt, 0, 16 * 512 * 512 * 2));
  CK(cudaMalloc(&out, 1024 * 8));
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUtensorMap m2, m4, m5, m3;
  { cuuint64_t d[2] = {512, 4096}, s[1] = {1024}; cuuint32_t b[2] = {64, 128};
    if (enc(&m2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, act, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc2 failed\n"); return 1; } }
  { cuuint64_t d[4] = {512, 64, 64, 1}, s[3] = {1024, 64 * 1024, 64 * 64 * 1024}; cuuint32_t b[4] = {64, 16, 8, 1};
    if (enc(&m4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, act, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc4 failed\n"); return 1; } }
  { // hi-res 64x64x512 viewed as (px*ld + c, W/2, py, H/2, B)
    cuuint64_t d[5] = {1024, 32, 2, 32, 1}, s[4] = {2 * 512 * 2, 64 * 512 * 2, 2 * 64 * 512 * 2, 64 * 64 * 512 * 2}; cuuint32_t b[5] = {64, 16, 1, 8, 1};
    if (enc(&m5, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, act, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc5 failed\n"); return 1; } }
  { cuuint64_t d[3] = {512, 512, 16}, s[2] = {1024, 512 * 1024}; cuuint32_t b[3] = {64, 64, 1};
    if (enc(&m3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wgt, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc3 failed\n"); return 1; } }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("variant,boxes,stage_KB,stages,ctas,cycles_per_stage,B_per_clk_per_SM,GBs_per_SM_at_1965MHz,aggregate_TBs\n");
  const char* names[4] = {"2d_gemm_128x64", "4d_phase_A", "5d_strided_A", "3d_weight_64x64"};
  struct Cfg { int variant, boxes; } cfgs[] = {{0, 1}, {0, 2}, {1, 1}, {2, 1}, {3, 1}, {3, 2}, {3, 4}};
  for (auto c : cfgs) {
    const int stageBytes = c.variant == 3 ? c.boxes * 8192 : c.boxes * 16384;
    for (int stages : {2, 4, 8}) {
      if ((size_t)stages * stageBytes > 200 * 1024) continue;
      for (int ctas : {1, 32, 128, 148}) {
        Args a{c.variant, stages, 512, c.boxes, stageBytes};
        const CUtensorMap& m = c.variant == 0 ? m2 : c.variant == 1 ? m4 : c.variant == 2 ? m5 : m3;
        const size_t smem = 200 * 1024 + 2048;  // one CTA per SM regardless of the ring size
        for (int rep = 0; rep < 2; ++rep) probe<<<ctas, 128, smem>>>(m, a, out);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(ctas);
        CK(cudaMemcpy(h.data(), out, ctas * 8, cudaMemcpyDeviceToHost));
        double mean = 0; for (auto v : h) mean += (double)v; mean /= ctas;
        const double perStage = mean / (a.iters - stages);
        const double bpc = stageBytes / perStage;
        printf("%s,%d,%d,%d,%d,%.1f,%.1f,%.1f,%.2f\n", names[c.variant], c.boxes, stageBytes / 1024, stages, ctas, perStage, bpc, bpc * 1.965, bpc * 1.965 * ctas / 1000.0);
      }
    }
  }
  CK(cudaFuncSetAttribute(probe_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  printf("\nlatency_cycles,prefetch_desc,A5d_first,A5d_2nd,A5d_3rd,B3d_first,B3d_2nd,B3d_3rd\n");
  for (int pf : {0, 1}) {
    for (int rep = 0; rep < 2; ++rep) {
      probe_lat<<<1, 128, 90 * 1024>>>(m5, m3, pf, out);
      CK(cudaDeviceSynchronize());
      unsigned long long h[8];
      CK(cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost));
      printf("lat,%d,%llu,%llu,%llu,%llu,%llu,%llu\n", pf, h[0], h[1], h[2], h[3], h[4], h[5]);
    }
  }
  CK(cudaFuncSetAttribute(probe_ab, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  printf("\nkernel_like,nb,stage_KB,stages,producers,ctas,cycles_per_stage,B_per_clk_per_SM,aggregate_TBs\n");
  for (int nb : {1, 2, 4}) {
    const int stageBytes = 16384 + nb * 8192;
    const int stages = 192 * 1024 / stageBytes;
    for (int producers : {1, 3}) {
      for (int ctas : {32, 148}) {
        const int iters = 1200;
        for (int rep = 0; rep < 2; ++rep) probe_ab<<<ctas, 256, 200 * 1024 + 2048>>>(m5, m3, stages, iters, nb, producers, out);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(ctas);
        CK(cudaMemcpy(h.data(), out, ctas * 8, cudaMemcpyDeviceToHost));
        double mean = 0; for (auto v : h) mean += (double)v; mean /= ctas;
        const double perStage = mean / (iters - 2 * stages);
        printf("A5d+B3d,%d,%d,%d,%d,%d,%.1f,%.1f,%.2f\n", nb, stageBytes / 1024, stages, producers, ctas, perStage, stageBytes / perStage, stageBytes / perStage * 1.965 * ctas / 1000.0);
      }
    }
  }
  return 0;
}
