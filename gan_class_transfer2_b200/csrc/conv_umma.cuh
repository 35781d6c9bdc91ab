// Implicit-GEMM 4x4 / stride-2 convolution family on tcgen05 + TMEM, operands staged by TMA.
//
// One persistent, warp-specialised kernel template serves all 24 tensor-core contractions of the
// reference's U-Net step (reference: train.py:145-169 DownShuffle/UpShuffle; the backward passes are
// implicit in Keras' train_step, SURVEY.md 8(a) rows a2,a3,a8).  Three index maps ("modes"):
//
//   MODE_S  strided form     out_lo[pix, n] = sum_{tap(16), k} G_hi[pix @ tap, k] * Wt[tap][k][n]
//           DownShuffle fprop (G = x, Wt = HWIO kernel) and UpShuffle dgrad (G = dy, Wt = HWOI kernel).
//           A: K-major im2col tile (5-D TMA box on the parity-split hi-res tensor), B: MN-major weights.
//   MODE_P  phase form       out_hi[2m+py, 2n+px, n] = sum_{tap(2x2), k} X_lo[(m,n)+d(tap,phase), k] * Wt[tap][n][k]
//           UpShuffle fprop (Wt = HWOI kernel) and DownShuffle dgrad (Wt = HWIO kernel).
//           A: K-major shifted tile (4-D TMA box, zero fill at the border), B: K-major weights.
//   MODE_W  weight gradient  dW[tap][g][p] = sum_pix G_hi[pix @ tap, g] * P_lo[pix, p]
//           both operands MN-major (pixels are the contraction dim and are the slow smem axis).
//
// Stride-1 index maps (reference: train.py:131-139, Block's ks x ks / stride-1 Conv2D with ks = 3; the dormant
// block_depth > 0 branch).  Input and output have the same extent; a tap is a whole-tile shift:
//   MODE_CF fprop            out[pix, n] = sum_{tap(ks*ks), k} X[pix + (tap - ks/2), k] * Wt[tap][k][n]     (B MN-major, HWIO)
//   MODE_CD dgrad            dx[pix, n]  = sum_{tap, k} dY[pix - (tap - ks/2), k] * Wt[tap][n][k]           (B K-major,  HWIO)
//   MODE_CW weight gradient  dW[tap][g][p] = sum_pix X[pix + (tap - ks/2), g] * dY[pix, p]
//           A of CF / CD and the X operand of CW: 4-D TMA box shifted by the tap, zero fill at the border (SAME padding).
//
// Warp roles: warps 0, 2, 3 = TMA producers (ring rounds round-robin; warp 2 also allocates TMEM), warp 1 = MMA
// issuer, warps 4.. = epilogue (8 warps; TMEM -> registers -> global).  Accumulators are
// double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Ring: a slot holds KPS k-chunks of 64 (KPS = 2 for BN <= 128, 1 for BN = 256), i.e. one full/empty barrier round trip
// feeds 8 (or 4) K = 16 MMAs.  The MMA-issuing thread's own instruction stream bounds the main loop (round 1: ~410
// cycles per 4-MMA round against a 128-cycle tensor floor at BN = 64), so narrow tiles pay the barrier round once per
// 128 of K instead of once per 64.
//
// PAIR = 1 instances run as clusters of two CTAs on tcgen05 cta_group::2: each CTA keeps its own 128-row tile, stages
// half of the B tile, and rank 0 issues one 256-row MMA per K = 16 slice for both (see ptx.cuh, "CTA pairs").
//
// Weights before the dependency: the S/P launches fetch the B (weight) boxes of their first ring pass BEFORE
// griddepcontrol.wait when the caller vouches that the kernel tensor is not being written by anything still running
// (ConvParams::bEarly, GCT2_WEIGHTS_STABLE in the C ABI); the A boxes (activations of the previous launch) follow after it.
//
// Per-CTA phase stamps (tools/timeline.py) are compiled in only with -DGCT2_TIMELINE: the production main loops carry
// no test-hook instructions.
#pragma once
#include <cuda_bf16.h>
#include <type_traits>
#include "ptx.cuh"

namespace gct2 {

enum : int { MODE_S = 0, MODE_P = 1, MODE_W = 2, MODE_CF = 3, MODE_CD = 4, MODE_CW = 5 };
__host__ __device__ constexpr bool mode_is_w(int m) { return m == MODE_W || m == MODE_CW; }
__host__ __device__ constexpr bool mode_is_s1(int m) { return m >= MODE_CF; }
enum : int { EPI_BIAS_RELU = 0, EPI_DGRAD = 1, EPI_WS_SLAB = 2, EPI_WGRAD = 3 };

__device__ __forceinline__ void epi_bar_sync(int nthreads) {  // named barrier 1: epilogue warps only
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// n / d for 0 <= n < 2^31 with a multiply-high and a shift (host-side magic numbers): an integer division by a run-time
// value costs a single thread 150-300 cycles, and the producer's start-up alone had ten of them on the critical path
// of every launch (1.2 us of the 2.4 us between griddepcontrol.wait and the first MMA, tools/timeline.py).
struct FastDiv {
  uint32_t d, mul, shr;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{d, 0u, 0u};
  if (d > 1) {
    uint32_t lg = 0;
    while ((1u << lg) < d) ++lg;
    const uint32_t pw = 31 + lg;
    f.mul = (uint32_t)(((1ull << pw) + d - 1) / d);
    f.shr = pw - 32;
  }
  return f;
}
__device__ __forceinline__ int fd_div(const FastDiv& f, int n) {
  return f.d == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr);
}
__device__ __forceinline__ int fd_mod(const FastDiv& f, int n) { return n - fd_div(f, n) * (int)f.d; }

struct ConvParams {
  int B, Hlo, Wlo;             // lo-res spatial extent (hi-res = 2x)
  int Wt, Ht, Nb;              // pixel-tile geometry; rows = Nb*Ht*Wt (128 for S/P, 64 for W)
  int tilesX, tilesY;          // pixel tiles along W and H (batch tiles = mTiles / (tilesX*tilesY))
  int mTiles;                  // S/P: pixel tiles; W: (M-side channels)/128
  int nTiles;                  // N / BN
  int splits;                  // split-K factor
  int kIters;                  // 64-wide k-chunks per work item
  int rounds;                  // ring rounds per work item = ceil(kIters / KPS)
  int kcPer;                   // S/P: 64-channel chunks per tap
  int numItems;                // total work items
  int ldG;                     // S / W(G operand): pixel stride (elements) of the gathered hi-res tensor
  int gIsA;                    // W: 1 when the gathered operand supplies the M side
  int mnLbo, mnSbo;            // MN-major descriptor offsets (bytes)
  int cm;                      // CTAs per cluster: 1, or 2 for a cta_group::2 pair (two consecutive M tiles of one
                               // (phase, N tile, split) share their B tile)
  int numClusterItems;         // numItems / cm
  int csplit;                  // split-K inside a thread-block cluster: the `splits` CTAs of a cluster own the K slices of ONE
                               // tile, keep their partial accumulators in their own shared memory and each finishes
                               // 128/splits rows by reading its peers' partials through distributed shared memory
  int fused;                   // split-K finished inside this launch (tile-major slabs + arrive/depart counters)
  int realEpi;                 // fused: the epilogue to apply after the slabs are summed (EPI_BIAS_RELU / EPI_DGRAD)
  int numTiles;                // fused: phases * nTiles * mTiles
  int* cnt;                    // fused: [numTiles][2] arrive / depart counters of THIS launch, zero between launches
  unsigned spinLimit;          // fused: rendezvous watchdog (polls of ~40 ns; 0 = wait for ever)
  int bEarly;                  // S/P: fetch the first ring pass of weight boxes before griddepcontrol.wait
  int f16;                     // 16-bit storage format of operands and outputs: 0 = bf16, 1 = fp16 (gct2_set_policy)
  int ks;                      // stride-1 modes: kernel side (3, or 1 for a per-pixel projection); taps = ks * ks
  int tapFuse;                 // MODE_W with a 64-channel gathered N side (up0): 4 = one work item computes FOUR filter taps --
                               // the un-shifted A tile is loaded once, the four 64-column blocks of a 256-wide B tile are the
                               // gathered operand at four different taps, and the epilogue scatters the blocks to their taps
  // fast division by the launch constants used in index decoding (all set by conv_launch)
  FastDiv fdMTilesC, fdNTiles, fdSplits, fdKcPer;
  int lgWt, lgHt, lgTilesX, lgTilesY;  // pixel-tile geometry is power-of-two by construction
  int stampPos;                // -DGCT2_TIMELINE only: which point of the producer's start-up stamp 7 records
  unsigned long long* dbg;     // -DGCT2_TIMELINE only: per-CTA phase timestamps (8 x u64 per CTA, %globaltimer ns)
  // epilogue
  int epi;
  int N;                       // total output columns (S/P) ; W: N-side channels
  int Hout, Wout;              // output spatial extent (S: lo-res, P: hi-res)
  __nv_bfloat16* out;          // bf16 output (pixel stride ldo)
  int ldo;
  const float* bias;
  const __nv_bfloat16* act;    // EPI_DGRAD: saved post-ReLU activation co-located with out (pixel stride ldact)
  int ldact;
  int maskN;                   // EPI_DGRAD: columns [0,maskN) are ReLU-masked
  int addOld;                  // EPI_DGRAD: add the value already stored in out (skip-path gradient)
  const __nv_bfloat16* addSrc; // stride-1 modes, EPI_DGRAD without split-K: add THIS tensor (pixel stride ldAdd) instead of
  int ldAdd;                   // out's old contents -- the `input + Dense(module(input))` of train.py:110-111
  float* ws;                   // EPI_WS_SLAB: fp32 [splits][outPixels][N]; EPI_WGRAD split-K: fp32 [splits][16*Chi*Clo]
  long long wsSplitStride;     // elements between consecutive split slabs
  float* dw;                   // EPI_WGRAD: fp32 [16][...]
  long long tapStride;
  int rowStride, colStride;    // element strides of the (M-row, N-col) accumulator tile inside one tap
  int atomic;                  // EPI_WGRAD: split-K -> each split stores its partial into its own slab of ws
};

struct WorkItem {
  int mt, nt, ph, split;       // ph: phase (P) or tap (W)
};

// item = index of a cluster item; rm = this CTA's rank inside its cluster.  Pairs: a cluster item is two consecutive M
// tiles and rm picks one; cluster split-K: a cluster item is one tile and rm is the K slice.
template <int MODE>
__device__ __forceinline__ WorkItem decode_item(const ConvParams& p, int item, int rm) {
  WorkItem w;
  if (p.csplit) {
    w.mt = fd_mod(p.fdMTilesC, item);
    const int r = fd_div(p.fdMTilesC, item);
    w.nt = fd_mod(p.fdNTiles, r);
    w.ph = fd_div(p.fdNTiles, r);
    w.split = rm;
    return w;
  }
  w.mt = fd_mod(p.fdMTilesC, item) * p.cm + rm;
  int r = fd_div(p.fdMTilesC, item);
  if (mode_is_w(MODE)) {
    w.nt = fd_mod(p.fdNTiles, r);
    r = fd_div(p.fdNTiles, r);
    w.split = fd_mod(p.fdSplits, r);
    w.ph = fd_div(p.fdSplits, r);
  } else {
    w.split = fd_mod(p.fdSplits, r);
    r = fd_div(p.fdSplits, r);
    w.nt = fd_mod(p.fdNTiles, r);
    w.ph = fd_div(p.fdNTiles, r);
  }
  return w;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#ifdef GCT2_TIMELINE
#define GCT2_STAMP(slot)                                                              \
  do {                                                                                \
    if (p.dbg != nullptr) p.dbg[(size_t)blockIdx.x * 8 + (slot)] = globaltimer_ns();  \
  } while (0)
#else
#define GCT2_STAMP(slot) do { } while (0)
#endif

// Epilogue warps per tile width.  Round 1 gave the BN = 128 / 256 instantiations 16 of them (640 threads, so 96 registers
// per thread and spills in every such kernel); measured in round 2 (profiles/r2_epilogue_warps_ab.jsonl): 8 warps with 128
// registers and no spills are faster at every batch size (0.5435 -> 0.5382 ms at batch 1, +1.9 % at 8 images, +0.8 % at 32)
// -- the epilogue of tile i overlaps the main loop of tile i + 1 anyway, and at batch 1 the split-K wait dominates it.
#ifndef GCT2_EPI_WARPS_WIDE
#define GCT2_EPI_WARPS_WIDE 8  // A/B hook (make variant)
#endif
template <int BN>
__host__ __device__ constexpr int kEpilogueWarps() {
  return BN == 64 ? 8 : GCT2_EPI_WARPS_WIDE;
}
template <int BN>
__host__ __device__ constexpr int kConvThreads() {
  return 128 + 32 * kEpilogueWarps<BN>();
}
// Ring geometry (shared by the kernel and the host-side shared-memory size).
template <int BN>
__host__ __device__ constexpr int kChunksPerStage() {
  return BN == 256 ? 1 : 2;
}
template <int BN, int PAIR>
__host__ __device__ constexpr int kSubBytes() {  // one k-chunk: 128 x 64 bf16 of A + (half of) BN x 64 bf16 of B
  return 128 * 128 + (PAIR ? BN * 64 : BN * 128);
}
template <int BN, int PAIR>
__host__ __device__ constexpr int kStageBytes() {
  return kChunksPerStage<BN>() * kSubBytes<BN, PAIR>();
}
template <int BN, int PAIR>
__host__ __device__ constexpr int kStages() {  // every shape fills the same 192 KB of pipeline
  return (192 * 1024) / kStageBytes<BN, PAIR>();
}

// One k-chunk (64 of K) of one work item on its way: the A box(es) and/or the B box(es) of ring position (sa, sb).
template <int MODE, int BN, int PAIR>
__device__ __forceinline__ void issue_chunk(const ConvParams& p, const CUtensorMap* mapA, const CUtensorMap* mapB,
                                            uint64_t* bar, uint8_t* sa, uint8_t* sb, const WorkItem& w, int x0, int y0,
                                            int b0, int kit, int rm, bool doA, bool doB) {
  constexpr int BLK = 64 * 128;  // one 64x64 bf16 block = one TMA box of an MN-major operand
  constexpr bool pair = PAIR != 0;
  if (MODE == MODE_S) {
    const int tap = fd_div(p.fdKcPer, kit), kc = kit - tap * p.kcPer;
    const int ky = tap >> 2, kx = tap & 3;
    const int py = (ky + 1) & 1, px = (kx + 1) & 1;
    const int hy = ((ky + 1) >> 1) - 1, hx = ((kx + 1) >> 1) - 1;
    if (doA) {
      if (pair)
        tma_load_5d_pair(sa, mapA, bar, px * p.ldG + kc * 64, x0 + hx, py, y0 + hy, b0);
      else
        tma_load_5d(sa, mapA, bar, px * p.ldG + kc * 64, x0 + hx, py, y0 + hy, b0);
    }
    if (doB) {
      if (pair) {
#pragma unroll
        for (int j = 0; j < BN / 128; ++j)  // my half of the tile's columns
          tma_load_3d_pair(sb + j * BLK, mapB, bar, w.nt * BN + rm * (BN / 2) + j * 64, kc * 64, tap);
      } else {
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_3d(sb + j * BLK, mapB, bar, w.nt * BN + j * 64, kc * 64, tap);
      }
    }
  } else if (MODE == MODE_P) {
    const int t4 = fd_div(p.fdKcPer, kit), kc = kit - t4 * p.kcPer;
    const int ty = t4 >> 1, tx = t4 & 1;
    const int py = w.ph >> 1, px = w.ph & 1;
    const int ky = (1 - py) + 2 * ty, kx = (1 - px) + 2 * tx;
    if (doA) {
      if (pair)
        tma_load_4d_pair(sa, mapA, bar, kc * 64, x0 + (px - tx), y0 + (py - ty), b0);
      else
        tma_load_4d(sa, mapA, bar, kc * 64, x0 + (px - tx), y0 + (py - ty), b0);
    }
    if (doB) {
      if (pair)  // my half of the tile's rows (the host built this map with BN/2-row boxes)
        tma_load_3d_pair(sb, mapB, bar, kc * 64, w.nt * BN + rm * (BN / 2), ky * 4 + kx);
      else
        tma_load_3d(sb, mapB, bar, kc * 64, w.nt * BN, ky * 4 + kx);
    }
  } else if (MODE == MODE_CF || MODE == MODE_CD) {
    const int tap = fd_div(p.fdKcPer, kit), kc = kit - tap * p.kcPer;
    const int ky = p.ks == 3 ? (tap * 11) >> 5 : 0, kx = tap - ky * p.ks, off = p.ks >> 1;
    const int sx = MODE == MODE_CF ? kx - off : off - kx, sy = MODE == MODE_CF ? ky - off : off - ky;
    if (doA) tma_load_4d(sa, mapA, bar, kc * 64, x0 + sx, y0 + sy, b0);
    if (doB) {
      if (MODE == MODE_CF) {
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_3d(sb + j * BLK, mapB, bar, w.nt * BN + j * 64, kc * 64, tap);
      } else {
        tma_load_3d(sb, mapB, bar, kc * 64, w.nt * BN, tap);
      }
    }
  } else if (MODE == MODE_CW) {
    const int cx = (kit & (p.tilesX - 1)) << p.lgWt;
    const int cy = ((kit >> p.lgTilesX) & (p.tilesY - 1)) << p.lgHt;
    const int cb = (kit >> (p.lgTilesX + p.lgTilesY)) * p.Nb;
    const int ky = p.ks == 3 ? (w.ph * 11) >> 5 : 0, kx = w.ph - ky * p.ks, off = p.ks >> 1;
    if (p.gIsA) {
#pragma unroll
      for (int j = 0; j < 2; ++j) tma_load_4d(sa + j * BLK, mapA, bar, w.mt * 128 + j * 64, cx + kx - off, cy + ky - off, cb);
#pragma unroll
      for (int j = 0; j < BN / 64; ++j) tma_load_4d(sb + j * BLK, mapB, bar, w.nt * BN + j * 64, cx, cy, cb);
    } else {
#pragma unroll
      for (int j = 0; j < 2; ++j) tma_load_4d(sa + j * BLK, mapA, bar, w.mt * 128 + j * 64, cx, cy, cb);
#pragma unroll
      for (int j = 0; j < BN / 64; ++j)
        tma_load_4d(sb + j * BLK, mapB, bar, w.nt * BN + j * 64, cx + kx - off, cy + ky - off, cb);
    }
  } else {
    // pixel chunk -> (batch tile, y tile, x tile); both operands are activations: always loaded together
    const int cx = (kit & (p.tilesX - 1)) << p.lgWt;
    const int cy = ((kit >> p.lgTilesX) & (p.tilesY - 1)) << p.lgHt;
    const int cb = (kit >> (p.lgTilesX + p.lgTilesY)) * p.Nb;
    const int ky = w.ph >> 2, kx = w.ph & 3;
    const int py = (ky + 1) & 1, px = (kx + 1) & 1;
    const int hy = ((ky + 1) >> 1) - 1, hx = ((kx + 1) >> 1) - 1;
    if (pair) {
      // my own 128 M-side channels, my half of the N-side channels; both land on the leader's barrier
      const int nb0 = w.nt * BN + rm * (BN / 2);
      if (p.gIsA) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_load_5d_pair(sa + j * BLK, mapA, bar, px * p.ldG + w.mt * 128 + j * 64, cx + hx, py, cy + hy, cb);
#pragma unroll
        for (int j = 0; j < BN / 128; ++j) tma_load_4d_pair(sb + j * BLK, mapB, bar, nb0 + j * 64, cx, cy, cb);
      } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) tma_load_4d_pair(sa + j * BLK, mapA, bar, w.mt * 128 + j * 64, cx, cy, cb);
#pragma unroll
        for (int j = 0; j < BN / 128; ++j)
          tma_load_5d_pair(sb + j * BLK, mapB, bar, px * p.ldG + nb0 + j * 64, cx + hx, py, cy + hy, cb);
      }
    } else if (p.gIsA) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
        tma_load_5d(sa + j * BLK, mapA, bar, px * p.ldG + w.mt * 128 + j * 64, cx + hx, py, cy + hy, cb);
#pragma unroll
      for (int j = 0; j < BN / 64; ++j) tma_load_4d(sb + j * BLK, mapB, bar, w.nt * BN + j * 64, cx, cy, cb);
    } else {
#pragma unroll
      for (int j = 0; j < 2; ++j) tma_load_4d(sa + j * BLK, mapA, bar, w.mt * 128 + j * 64, cx, cy, cb);
      if (BN == 256 && p.tapFuse) {
        // four taps share the A tile: block j of the B tile is the gathered operand's 64 channels at tap 4 * ph + j
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) {
          const int kx2 = j, ky2 = w.ph;  // tap = 4 * ph + j  ->  (ky, kx) = (ph, j)
          const int py2 = (ky2 + 1) & 1, px2 = (kx2 + 1) & 1;
          const int hy2 = ((ky2 + 1) >> 1) - 1, hx2 = ((kx2 + 1) >> 1) - 1;
          tma_load_5d(sb + j * BLK, mapB, bar, px2 * p.ldG, cx + hx2, py2, cy + hy2, cb);
        }
      } else {
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_5d(sb + j * BLK, mapB, bar, px * p.ldG + w.nt * BN + j * 64, cx + hx, py, cy + hy, cb);
      }
    }
  }
}

// Split-K finished in place (called by the epilogue warps once their partial tile is stored): split s sums rows
// [s*R, (s+1)*R) of the tile over all partials in split order (bit-reproducible), applies the real epilogue and writes
// the 16-bit output -- coalesced, no extra launch.
//   cluster form (csplit): the splits are the CTAs of one cluster (co-scheduled by the hardware, so they may wait for
//     each other whatever else runs on the GPU); partials are read from the peers' shared memory;
//   L2 form (fused): every split of the tile has its own CTA and the host launched no more CTAs than can be resident
//     at once; partials go through tile-major fp32 slabs in global memory and a counter rendezvous.
// (inlined: as a separate function it cost 11 us per step at batch 1 -- call overhead on the critical path of every
// split-K launch)
#define GCT2_FINISH_ATTR __forceinline__
#ifdef GCT2_NO_F16
#define GCT2_F16_OF(p) 0
#else
#define GCT2_F16_OF(p) ((p).f16)
#endif
template <int MODE, int BN, int CS>
__device__ GCT2_FINISH_ATTR void splitk_finish_in_place(const ConvParams& p, const WorkItem& w, int tileId,
                                                    uint8_t* smem, uint64_t* red_full, int warp, int lane) {
  constexpr bool csplit = CS != 0;
  constexpr int NE = kEpilogueWarps<BN>();
  const int f16 = GCT2_F16_OF(p);
      // ---- split-K finished in place: split s sums rows [s*R, (s+1)*R) of the tile over all partials in split
      // order (bit-reproducible), applies the real epilogue and writes the bf16 output -- coalesced, no extra launch.
      //   cluster form (csplit): the splits are the CTAs of one cluster (co-scheduled by the hardware, so they may
      //     wait for each other whatever else runs on the GPU); partials are read from the peers' shared memory;
      //   L2 form (fused): every split of the tile has its own CTA and the host launched no more CTAs than can be
      //     resident at once; partials go through tile-major fp32 slabs in global memory and a counter rendezvous.
      constexpr int NT = NE * 32;
      const int et = (warp - 4) * 32 + lane;
      const uint32_t part0 = smem_u32(smem);
      if (csplit) {
        __syncwarp();
        if (lane == 0) {
          const uint32_t bar = smem_u32(red_full);
          for (int d = 0; d < p.splits; ++d) mbar_arrive_cluster_release(map_to_rank(bar, (uint32_t)d));
        }
        mbar_wait_cluster_acquire(red_full, 0);
      } else {
        __threadfence();
        epi_bar_sync(NT);
        if (et == 0) {
          atomicAdd(p.cnt + 2 * tileId, 1);
          uint32_t spins = 0;
          while (ld_acquire_gpu(p.cnt + 2 * tileId) < p.splits) {
            __nanosleep(40);
            if (p.spinLimit != 0u && ++spins > p.spinLimit) {
              printf("gct2: split-K rendezvous watchdog block %d tile %d\n", (int)blockIdx.x, tileId);
              __trap();
            }
          }
        }
        epi_bar_sync(NT);
      }
      const int R = 128 / p.splits, vecPerRow = BN / 4;
      const float* slab0 = p.ws + ((long long)tileId * 128) * BN;
      const long long splitStride = (long long)p.numTiles * 128 * BN;
      const int items = R * vecPerRow;
      if (!csplit) {
        // L2 form, pass 1: ALL partials of this CTA's rows are requested at once, global -> shared memory without
        // passing through registers (cp.async into the operand ring, idle now: the CTA's only work item has retired its
        // MMAs).  The sum below then costs one L2 round trip whatever the split factor; loading the partials one after
        // the other was most of a 6-9 us epilogue at 16 / 32 splits (the 4x4 and 8x8 layers at batch 1).  Layout
        // [split][item]: a warp's 16-byte reads are consecutive.  Every thread reads back only what it requested itself.
        for (int idx = et; idx < items; idx += NT) {
          const int rr = w.split * R + idx / vecPerRow, c4 = (idx % vecPerRow) * 4;
          const int bl = rr >> (p.lgWt + p.lgHt);
          if ((w.mt >> (p.lgTilesX + p.lgTilesY)) * p.Nb + bl >= p.B) continue;
          const float* sp = slab0 + (long long)rr * BN + c4;
          const uint32_t dst = part0 + (uint32_t)idx * 16u;
          for (int sidx = 0; sidx < p.splits; ++sidx)
            cp_async_cg16(dst + (uint32_t)(sidx * items) * 16u, sp + (long long)sidx * splitStride);
        }
        cp_async_wait_all();
      }
      for (int idx = et; idx < items; idx += NT) {
        const int rr = w.split * R + idx / vecPerRow, c4 = (idx % vecPerRow) * 4;
        const int xl = rr & (p.Wt - 1), yl = (rr >> p.lgWt) & (p.Ht - 1), bl = rr >> (p.lgWt + p.lgHt);
        const int b = (w.mt >> (p.lgTilesX + p.lgTilesY)) * p.Nb + bl;
        if (b >= p.B) continue;
        int oy = (((w.mt >> p.lgTilesX) & (p.tilesY - 1)) << p.lgHt) + yl, ox = ((w.mt & (p.tilesX - 1)) << p.lgWt) + xl;
        if (MODE == MODE_P) {
          oy = 2 * oy + (w.ph >> 1);
          ox = 2 * ox + (w.ph & 1);
        }
        const long long opix = ((long long)b * p.Hout + oy) * p.Wout + ox;
        float4 v;
        if (csplit) {
          // four remote loads in flight per batch: the sum is bound by one distributed-shared-memory round trip per
          // batch, not per split; summation order stays split 0, 1, 2, ... (bit-reproducible)
          v = make_float4(0.f, 0.f, 0.f, 0.f);
          const uint32_t off = part0 + (uint32_t)(rr * BN * 4) + (uint32_t)((((c4 >> 2) ^ (rr & 7))) << 4);
          for (int s0 = 0; s0 < p.splits; s0 += 4) {
            float4 u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (s0 + j < p.splits) u[j] = ld_dsmem_f4(map_to_rank(off, (uint32_t)(s0 + j)));
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (s0 + j < p.splits) {
                v.x += u[j].x; v.y += u[j].y; v.z += u[j].z; v.w += u[j].w;
              }
          }
        } else {
          // pass 2: summation order split 0, 1, 2, ... -- bit-identical to the finishing kernel
          const uint32_t src = part0 + (uint32_t)idx * 16u;
          v = ld_smem_f4(src);
          for (int sidx = 1; sidx < p.splits; ++sidx) {
            const float4 u = ld_smem_f4(src + (uint32_t)(sidx * items) * 16u);
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
          }
        }
        const int nn = w.nt * BN + c4;
        __nv_bfloat16* o = p.out + opix * p.ldo + nn;
        if (p.realEpi == EPI_BIAS_RELU) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + nn));
          v.x = fmaxf(v.x + bb.x, 0.f); v.y = fmaxf(v.y + bb.y, 0.f);
          v.z = fmaxf(v.z + bb.z, 0.f); v.w = fmaxf(v.w + bb.w, 0.f);
        } else {
          if (p.addOld) {
            const uint2 old = *reinterpret_cast<const uint2*>(o);
            v.x += h_lo(old.x, f16); v.y += h_hi(old.x, f16); v.z += h_lo(old.y, f16); v.w += h_hi(old.y, f16);
          }
          if (nn < p.maskN) {
            const uint2 am = __ldg(reinterpret_cast<const uint2*>(p.act + opix * p.ldact + nn));
            v.x = h_pos_lo(am.x) ? v.x : 0.f; v.y = h_pos_hi(am.x) ? v.y : 0.f;
            v.z = h_pos_lo(am.y) ? v.z : 0.f; v.w = h_pos_hi(am.y) ? v.w : 0.f;
          }
        }
        uint2 res;
        res.x = pack_h2(v.x, v.y, f16);
        res.y = pack_h2(v.z, v.w, f16);
        *reinterpret_cast<uint2*>(o) = res;
      }
      if (!csplit) {
        epi_bar_sync(NT);
        if (et == 0) {
          const int old = atomicAdd(p.cnt + 2 * tileId + 1, 1);
          if (old == p.splits - 1) {  // every split has passed the rendezvous: re-arm the counters
            p.cnt[2 * tileId] = 0;
            p.cnt[2 * tileId + 1] = 0;
          }
        }
      }
}

#ifndef GCT2_CONV_EXTRA_BOUND
#define GCT2_CONV_EXTRA_BOUND 128  // launch bound 512 for 384-thread CTAs: at most 128 registers per thread
#endif
// PAIR = 1: the cta_group::2 variant (a separate instantiation: a kernel that contains cta_group::2 instructions can
// only be launched as a cluster of an even number of CTAs).
// CS = 1: the variant whose split-K is finished inside a thread-block cluster (DSMEM).  A separate instantiation so that
// the default kernels do not carry its code: with it compiled in, every launch of the step was slower (0.582 vs 0.564 ms
// per step at batch 1 with the path unused -- the kernels grow past what the instruction cache holds at start-up).
template <int MODE, int BN, int PAIR = 0, int CS = 0>
__global__ void __launch_bounds__(kConvThreads<BN>() + GCT2_CONV_EXTRA_BOUND, 1) conv_umma_kernel(const __grid_constant__ CUtensorMap mapA,
                                                        const __grid_constant__ CUtensorMap mapB,
                                                        const __grid_constant__ ConvParams p) {
  constexpr int A_BYTES = 128 * 128;  // 128 rows x 64 bf16 (S/P) or 2 blocks of 64 pixels x 64 channels (W)
  constexpr bool pair = PAIR != 0;
  constexpr int KPS = kChunksPerStage<BN>();
  constexpr int SUB_BYTES = kSubBytes<BN, PAIR>();
  constexpr int STAGE_BYTES = kStageBytes<BN, PAIR>();
  constexpr int S = kStages<BN, PAIR>();
  constexpr uint32_t TMEM_COLS = 2 * BN;
  static_assert(S >= 3, "three producers need at least three ring slots");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * STAGE_BYTES);
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint64_t* tempty = tfull + 2;
  uint64_t* red_full = tempty + 2;  // cluster split-K: every peer's partial tile is in its shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(red_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  TraceScope trace(100 + MODE * 10 + BN / 64);
#ifdef GCT2_TIMELINE
  const unsigned long long t_entry = (p.dbg != nullptr && threadIdx.x == 0) ? globaltimer_ns() : 0ull;
#endif

  constexpr bool csplit = CS != 0 && !pair && !mode_is_w(MODE);
  const int rm = (pair || csplit) ? (int)cluster_ctarank() : 0;
  const int clusterId = pair ? (int)(blockIdx.x >> 1) : (csplit ? fd_div(p.fdSplits, (int)blockIdx.x) : (int)blockIdx.x);
  const int numClusters = pair ? (int)(gridDim.x >> 1) : (csplit ? fd_div(p.fdSplits, (int)gridDim.x) : (int)gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);  // pair: the leader's single commit is multicast to both CTAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      // pair: the leader's accumulator buffer is free once the epilogues of BOTH CTAs have drained theirs
      mbar_init(&tempty[i], (pair ? 2 : 1) * kEpilogueWarps<BN>());
    }
    if (csplit) mbar_init(red_full, (uint32_t)(p.splits * kEpilogueWarps<BN>()));  // one arrival per epilogue warp of every peer
    fence_mbar_init();
  }
  if (warp == 2) {
    if (pair) {
      tmem_alloc_pair(tmem_slot, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (pair || csplit) cluster_sync_all();  // the peers' barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // producer identity (warps 0, 2, 3 -> producers 0, 1, 2); elect.sync (not `lane == 0`) picks the thread: ptxas then
  // knows exactly one thread runs the loop and feeds the uniform-datapath instructions (UTMALDG / UTCHMMA / UTCBAR)
  // directly instead of wrapping each one in a vote-and-branch loop over "possibly several" active threads.
  constexpr uint32_t NP = 3;
  const bool producerWarp = warp == 0 || warp == 2 || warp == 3;
  const uint32_t pj = warp == 0 ? 0u : (uint32_t)warp - 1u;
  const bool producer = producerWarp && elect_one();
  // the full-barrier arrival of one ring round: pair -> one arrival on the LEADER's barrier announces the bytes of both
  // CTAs (the loads of both complete there)
  auto expect_round = [&](uint32_t stage, int nch) {
    if (!pair)
      mbar_arrive_expect_tx(&full[stage], (uint32_t)(nch * SUB_BYTES));
    else if (rm == 0)
      mbar_arrive_expect_tx(&full[stage], (uint32_t)(2 * nch * SUB_BYTES));
  };
  const int lastCh = p.kIters - (p.rounds - 1) * KPS;  // chunks of an item's last round (1 or KPS)

  // ---- weights of the first ring pass, before the dependency resolves (they are not produced by the previous launch)
  const bool early = !mode_is_w(MODE) && p.bEarly != 0;
  if (early && producer && clusterId < p.numClusterItems) {
    const WorkItem w = decode_item<MODE>(p, clusterId, rm);
    const int firstPass = p.rounds < S ? p.rounds : S;
    for (int rd = (int)pj; rd < firstPass; rd += (int)NP) {
      const int nch = rd == p.rounds - 1 ? lastCh : KPS;
      expect_round((uint32_t)rd, nch);
      uint8_t* st = smem + rd * STAGE_BYTES;
#pragma unroll
      for (int c = 0; c < KPS; ++c)
        if (c < nch)
          issue_chunk<MODE, BN, PAIR>(p, &mapA, &mapB, &full[rd], st + c * SUB_BYTES, st + c * SUB_BYTES + A_BYTES, w, 0,
                                      0, 0, w.split * p.kIters + rd * KPS + c, rm, false, true);
    }
  }

  // Programmatic dependent launch: everything above ran while the previous kernel of the stream was still draining;
  // from here on activations in global memory are touched, so wait for that kernel to complete -- and let the next
  // kernel start its own prologue now.
  pdl_launch_dependents();
  pdl_wait();
  trace.ready();
#ifdef GCT2_TIMELINE
  if (threadIdx.x == 0 && p.dbg != nullptr) p.dbg[(size_t)blockIdx.x * 8 + 0] = t_entry;  // kernel entry
  if (threadIdx.x == 0) GCT2_STAMP(1);  // prologue done (and the previous kernel complete)
#endif

  if (producerWarp) {
    // ------------------------------------------------------------------ TMA producers (three warps)
    // One thread needs ~600-900 cycles to get one stage on its way (empty-wait, expect_tx and 2-5 TMA instructions
    // issue back to back at ~150 cycles each: tools/probes/tma_ingest_probe.cu), more than the tensor pipe needs to
    // consume it, and a deeper ring does not help because the limit is issue, not latency.  Three producer threads in
    // three warps take the ring rounds round-robin (producer j owns global round g = j mod 3): measured 2.6x the
    // single-producer rate in isolation.  S >= 3 keeps a producer's parity wait from aliasing (its previous round
    // g - 3 could be issued only after the consumer released g - 3 - S, and releases happen in order).
    if (producer) {
#ifdef GCT2_TIMELINE
      if (warp == 0 && p.stampPos == 0) GCT2_STAMP(7);
#endif
      uint32_t gbase = 0;
      for (int item = clusterId; item < p.numClusterItems; item += numClusters, gbase += (uint32_t)p.rounds) {
        const WorkItem w = decode_item<MODE>(p, item, rm);
        int x0 = 0, y0 = 0, b0 = 0;
        if (!mode_is_w(MODE)) {
          x0 = (w.mt & (p.tilesX - 1)) << p.lgWt;
          y0 = ((w.mt >> p.lgTilesX) & (p.tilesY - 1)) << p.lgHt;
          b0 = (w.mt >> (p.lgTilesX + p.lgTilesY)) * p.Nb;
        }
        const uint32_t gm = gbase % NP;
        for (int rd = (int)((pj + NP - gm) % NP); rd < p.rounds; rd += (int)NP) {
          const uint32_t g = gbase + (uint32_t)rd;
          const uint32_t ring = g / (uint32_t)S, stage = g - ring * (uint32_t)S, phase = ring & 1u;
          const int nch = rd == p.rounds - 1 ? lastCh : KPS;
          // the weights of the first item's first ring pass are already on their way (and its slots were never used)
          const bool wEarly = early && item == clusterId && rd < S;
          if (!wEarly) {
            mbar_wait(&empty[stage], phase ^ 1);
            expect_round(stage, nch);
          }
          uint8_t* st = smem + stage * STAGE_BYTES;
#pragma unroll
          for (int c = 0; c < KPS; ++c)
            if (c < nch)
              issue_chunk<MODE, BN, PAIR>(p, &mapA, &mapB, &full[stage], st + c * SUB_BYTES,
                                          st + c * SUB_BYTES + A_BYTES, w, x0, y0, b0,
                                          w.split * p.kIters + rd * KPS + c, rm, true, !wEarly);
#ifdef GCT2_TIMELINE
          if (g == 0 && p.stampPos == 4) GCT2_STAMP(7);
#endif
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (pair: the leader CTA only)
    if ((!pair || rm == 0) && elect_one()) {
      constexpr int A_MN = mode_is_w(MODE) ? 1 : 0;
      constexpr int B_MN = (MODE == MODE_P || MODE == MODE_CD) ? 0 : 1;
      // cta_group::2: 256 rows over the pair; operand format by the launch's storage policy
      const uint32_t idesc = GCT2_F16_OF(p) ? make_idesc_bf16(pair ? 256 : 128, BN, A_MN, B_MN, 0)
                                   : make_idesc_bf16(pair ? 256 : 128, BN, A_MN, B_MN, 1);
      const uint32_t a_lbo = A_MN ? (uint32_t)p.mnLbo : 16u, a_sbo = A_MN ? (uint32_t)p.mnSbo : 1024u;
      const uint32_t b_lbo = B_MN ? (uint32_t)p.mnLbo : 16u, b_sbo = B_MN ? (uint32_t)p.mnSbo : 1024u;
      constexpr uint32_t a_kstep = A_MN ? 2048u : 32u;  // bytes per UMMA_K = 16 along K
      constexpr uint32_t b_kstep = B_MN ? 2048u : 32u;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      // Everything that can be carried across iterations is: the stage's descriptor pair and barrier addresses
      // advance by constants and wrap with the ring; the K = 16 slices add immediates to the 14-bit address field.
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t da_first = make_smem_desc(smem_base, a_lbo, a_sbo);
      const uint64_t db_first = make_smem_desc(smem_base + A_BYTES, b_lbo, b_sbo);
      constexpr uint64_t dstep = (uint64_t)(STAGE_BYTES >> 4);
      constexpr uint64_t dsub = (uint64_t)(SUB_BYTES >> 4);
      const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
      constexpr uint64_t ka = a_kstep >> 4, kb = b_kstep >> 4;
      auto mma4 = [&](uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t accumulate) {
        if (pair) {
          umma_bf16_pair(d_tmem, da, db, idesc, accumulate);
          umma_bf16_pair(d_tmem, da + ka, db + kb, idesc, 1u);
          umma_bf16_pair(d_tmem, da + 2 * ka, db + 2 * kb, idesc, 1u);
          umma_bf16_pair(d_tmem, da + 3 * ka, db + 3 * kb, idesc, 1u);
        } else {
          umma_bf16(d_tmem, da, db, idesc, accumulate);
          umma_bf16(d_tmem, da + ka, db + kb, idesc, 1u);
          umma_bf16(d_tmem, da + 2 * ka, db + 2 * kb, idesc, 1u);
          umma_bf16(d_tmem, da + 3 * ka, db + 3 * kb, idesc, 1u);
        }
      };
      uint64_t da = da_first, db = db_first;
      uint32_t full_a = full0, empty_a = empty0;
      for (int item = clusterId; item < p.numClusterItems; item += numClusters) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
#ifdef GCT2_TIMELINE
        if (item == clusterId && p.dbg != nullptr) {  // when did the first operands land?
          mbar_wait_addr(full_a, phase);
          GCT2_STAMP(2);
        }
#endif
        for (int rd = 0; rd < p.rounds; ++rd) {
          // one ring round: wait for the operands, 4 * nch MMAs, release the slot.  (No tcgen05.fence here: the operands
          // were written by the async proxy (TMA) and are read by the async proxy (tcgen05.mma); the mbarrier's
          // completion orders the two.)
          mbar_wait_addr(full_a, phase);
          mma4(d_tmem, da, db, rd != 0 ? 1u : 0u);
          if (KPS == 2 && (rd + 1 < p.rounds || lastCh == 2)) mma4(d_tmem, da + dsub, db + dsub, 1u);
          // frees the smem slot once these MMAs have read it (pair: in both CTAs)
          if (pair)
            umma_commit_pair_addr(empty_a, 0x3);
          else
            umma_commit_addr(empty_a);
          da += dstep;
          db += dstep;
          full_a += 8;
          empty_a += 8;
          if (++stage == (uint32_t)S) {
            stage = 0;
            phase ^= 1;
            da = da_first;
            db = db_first;
            full_a = full0;
            empty_a = empty0;
          }
        }
        if (pair)
          umma_commit_pair(&tfull[acc], 0x3);  // both CTAs' accumulator halves complete -> both epilogues
        else
          umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
#ifdef GCT2_TIMELINE
        if (item == clusterId) GCT2_STAMP(3);  // all MMAs of the first item issued
#endif
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 4 .. 4+NE-1)
    // At batch 1 a CTA owns a single tile, so nothing overlaps its epilogue: it is spread over NE warps.  Warp w may
    // only read TMEM lanes 32*(w%4)..+31 (hardware rule), so warps w, w+4, w+8, ... share a lane quarter and split the
    // tile's columns.  Lane l of a warp owns accumulator row (q*32 + l): 16-byte vector accesses per row.
    {
      constexpr int NE = kEpilogueWarps<BN>();
      constexpr int COLS = BN / (NE / 4);  // columns per warp (32 or 64)
      const int q = warp & 3;              // TMEM lane quarter this warp may access
      const int cgrp = (warp - 4) >> 2;    // which slice of the tile's columns
      const int r = q * 32 + lane;
      const int f16 = GCT2_F16_OF(p);
      uint32_t acc = 0, acc_phase = 0;
      for (int item = clusterId; item < p.numClusterItems; item += numClusters) {
        const WorkItem w = decode_item<MODE>(p, item, rm);
        const int n0 = w.nt * BN + cgrp * COLS;
        const int tileId = (w.ph * p.nTiles + w.nt) * p.mTiles + w.mt;
        // row -> output location
        bool valid = true;
        long long pix = 0;
        if (!mode_is_w(MODE)) {
          const int xl = r & (p.Wt - 1), yl = (r >> p.lgWt) & (p.Ht - 1), bl = r >> (p.lgWt + p.lgHt);
          const int x = ((w.mt & (p.tilesX - 1)) << p.lgWt) + xl;
          const int y = (((w.mt >> p.lgTilesX) & (p.tilesY - 1)) << p.lgHt) + yl;
          const int b = (w.mt >> (p.lgTilesX + p.lgTilesY)) * p.Nb + bl;
          valid = b < p.B;
          int oy = y, ox = x;
          if (MODE == MODE_P) {
            oy = 2 * y + (w.ph >> 1);
            ox = 2 * x + (w.ph & 1);
          }
          pix = ((long long)b * p.Hout + oy) * p.Wout + ox;
        }
        mbar_wait(&tfull[acc], acc_phase);
#ifdef GCT2_TIMELINE
        if (item == clusterId && warp == 4 && lane == 0) GCT2_STAMP(4);  // first accumulator complete
#endif
        tc_fence_after();
        const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN + cgrp * COLS;
#pragma unroll 1
        for (int c0 = 0; c0 < COLS; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + c0, v);
          const int n = n0 + c0;
          if (p.epi == EPI_BIAS_RELU) {
            tmem_ld_wait();
            if (valid) {
              const float4* bp = reinterpret_cast<const float4*>(p.bias + n);
              uint4* dst = reinterpret_cast<uint4*>(p.out + pix * p.ldo + n);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const float4 b0 = __ldg(bp + 2 * g), b1 = __ldg(bp + 2 * g + 1);
                uint4 o;
                o.x = pack_h2(fmaxf(__uint_as_float(v[8 * g + 0]) + b0.x, 0.f),
                              fmaxf(__uint_as_float(v[8 * g + 1]) + b0.y, 0.f), f16);
                o.y = pack_h2(fmaxf(__uint_as_float(v[8 * g + 2]) + b0.z, 0.f),
                              fmaxf(__uint_as_float(v[8 * g + 3]) + b0.w, 0.f), f16);
                o.z = pack_h2(fmaxf(__uint_as_float(v[8 * g + 4]) + b1.x, 0.f),
                              fmaxf(__uint_as_float(v[8 * g + 5]) + b1.y, 0.f), f16);
                o.w = pack_h2(fmaxf(__uint_as_float(v[8 * g + 6]) + b1.z, 0.f),
                              fmaxf(__uint_as_float(v[8 * g + 7]) + b1.w, 0.f), f16);
                dst[g] = o;
              }
            }
          } else if (p.epi == EPI_DGRAD) {
            // the saved activation and the skip gradient are requested while the TMEM load is in flight
            uint4* dst = reinterpret_cast<uint4*>(p.out + pix * p.ldo + n);
            const uint4* ap = reinterpret_cast<const uint4*>(p.act + pix * p.ldact + n);
            const uint4* addp = dst;
            if constexpr (mode_is_s1(MODE)) {
              if (p.addSrc != nullptr) addp = reinterpret_cast<const uint4*>(p.addSrc + pix * p.ldAdd + n);
            }
            const bool masked = valid && n < p.maskN, add = valid && p.addOld;
            uint4 av[4], ov[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              av[g] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);  // positive in both formats = keep
              ov[g] = make_uint4(0u, 0u, 0u, 0u);
              if (masked) av[g] = __ldg(ap + g);
              if (add) ov[g] = addp[g];
            }
            tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[8 * g + j]);
                const uint4 o = ov[g], a = av[g];
                f[0] += h_lo(o.x, f16); f[1] += h_hi(o.x, f16); f[2] += h_lo(o.y, f16); f[3] += h_hi(o.y, f16);
                f[4] += h_lo(o.z, f16); f[5] += h_hi(o.z, f16); f[6] += h_lo(o.w, f16); f[7] += h_hi(o.w, f16);
                f[0] = h_pos_lo(a.x) ? f[0] : 0.f; f[1] = h_pos_hi(a.x) ? f[1] : 0.f;
                f[2] = h_pos_lo(a.y) ? f[2] : 0.f; f[3] = h_pos_hi(a.y) ? f[3] : 0.f;
                f[4] = h_pos_lo(a.z) ? f[4] : 0.f; f[5] = h_pos_hi(a.z) ? f[5] : 0.f;
                f[6] = h_pos_lo(a.w) ? f[6] : 0.f; f[7] = h_pos_hi(a.w) ? f[7] : 0.f;
                uint4 res;
                res.x = pack_h2(f[0], f[1], f16);
                res.y = pack_h2(f[2], f[3], f16);
                res.z = pack_h2(f[4], f[5], f16);
                res.w = pack_h2(f[6], f[7], f16);
                dst[g] = res;
              }
            }
          } else if (p.epi == EPI_WS_SLAB) {
            // split-K partial: plain stores into this split's slab (no atomics: the finishing pass sums the slabs in
            // a fixed order, so the step is bit-reproducible)
            tmem_ld_wait();
            if (csplit) {
              // cluster split-K: the partial tile goes to THIS CTA's shared memory (the ring is idle: every MMA of the
              // CTA's only work item has completed), rows of BN floats with the 16-byte chunks XOR-swizzled by the row
              // so that the 32 lanes of a warp (32 rows, same chunk) spread over all banks
              float4* rowp = reinterpret_cast<float4*>(smem) + (size_t)r * (BN / 4);
              const int ch0 = (cgrp * COLS + c0) >> 2;
#pragma unroll
              for (int g = 0; g < 8; ++g)
                rowp[(ch0 + g) ^ (r & 7)] = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                                                         __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
            } else if (valid) {
              // fused finish: tile-major slab [split][tile][row][col] (compact, read back coalesced below);
              // finishing kernel: pixel-major slab [split][pixel][N]
              float4* dst = reinterpret_cast<float4*>(
                  p.fused ? p.ws + (((long long)w.split * p.numTiles + tileId) * 128 + r) * BN + (n - w.nt * BN)
                          : p.ws + (long long)w.split * p.wsSplitStride + pix * p.N + n);
#pragma unroll
              for (int g = 0; g < 8; ++g)
                dst[g] = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                                     __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
            }
          } else {  // EPI_WGRAD: row = M-side channel, columns = N-side channels
            tmem_ld_wait();
            // (tap-fused items: the tile's 64-column block n / 64 belongs to tap 4 * ph + n / 64)
            const int tapOf = (BN == 256 && p.tapFuse) ? w.ph * 4 + (n >> 6) : w.ph;
            const int colOf = (BN == 256 && p.tapFuse) ? (n & 63) : n;
            float* base = (p.atomic ? p.ws + (long long)w.split * p.wsSplitStride : p.dw) +
                          (long long)tapOf * p.tapStride + (long long)(w.mt * 128 + r) * p.rowStride +
                          (long long)colOf * p.colStride;
            if (p.colStride == 1) {
              float4* d4 = reinterpret_cast<float4*>(base);
#pragma unroll
              for (int g = 0; g < 8; ++g)
                d4[g] = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                                    __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) base[(long long)j * p.colStride] = __uint_as_float(v[j]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if constexpr (!mode_is_w(MODE)) {  // (weight gradients never finish in place: keep the code out of their kernels)
          if (p.fused || csplit) splitk_finish_in_place<MODE, BN, csplit ? 1 : 0>(p, w, tileId, smem, red_full, warp, lane);
        }
#ifdef GCT2_TIMELINE
        if (item == clusterId && warp == 4 && lane == 0) GCT2_STAMP(5);  // first epilogue (of warp 4) done
#endif
        if (lane == 0) {
          if (pair && rm != 0)
            mbar_arrive_remote(&tempty[acc], 0);  // the leader's MMA thread waits for both CTAs' epilogues
          else
            mbar_arrive(&tempty[acc]);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  // no CTA leaves while a peer may still signal its barriers or read its shared memory
  if (pair || csplit) cluster_sync_all();
#ifdef GCT2_TIMELINE
  if (threadIdx.x == 0) GCT2_STAMP(6);  // all work of this CTA done
#endif
  trace.end();
  if (warp == 2) {
    tc_fence_after();
    if (pair)
      tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else
      tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace gct2
