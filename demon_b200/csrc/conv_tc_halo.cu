// tcgen05 implicit-GEMM convolution, "halo" variant: every input pixel is fetched ONCE per output tile.
//
// The output tile is 16 rows x 8 columns and, per 32-channel chunk, ONE TMA box per stride-parity plane brings the tile
// plus its halo into shared memory ((16+nqy-1) x (8+nqx-1) pixels, 128 bytes per pixel, 128-byte swizzle).  Every filter
// tap is then just a different START ADDRESS inside the same shared-memory image:
//
//     GEMM row m = (y, x) of the tile  ->  halo pixel (y + qy - qy_min, x + qx - qx_min)
//
// A STEP of the main loop is one (32-channel chunk, input shift) pair.  For a plain convolution a shift is a filter tap;
// for a transposed convolution k4 s2 (four sub-pixel classes of 2x2 taps each) the 16 class-taps use only NINE distinct
// input shifts, so the A operand of a shift is staged once and multiplied with the weight blocks of every class that uses
// it (4 classes at shift (0,0), 2 on the edges, 1 in the corners): 9 steps per chunk instead of 16.
//
// Warp roles (persistent over tiles, mbarrier rings between them, every wait bounded):
//     warp 0        A producer: the halo boxes (TMA)
//     warp 1        MMA thread: TS-mode tcgen05.mma kind::tf32 (A from TMEM, B = weights from shared memory)
//     warps 2..     stagers, G groups of four warps taking turns: shifted pixel rows of the halo image -> registers ->
//                   A_hi (raw fp32) and A_lo = A - trunc_tf32(A) in a TMEM ring slot (tcgen05.st)
//     next 4 warps  epilogue: tcgen05.ld -> transpose through shared memory -> bias, leaky ReLU -> coalesced NHWC stores
//     last warp     W producer: the pre-swizzled weight blocks of the step (cp.async.bulk)
//
// ONE operand ring of kRing = 4 slots couples them: slot s = (TMEM columns of A_hi | A_lo, the weight blocks in shared
// memory, barrier full[s], barrier free[s]).  full[s] collects the four stager warps of the step plus the weight bytes
// (expect_tx), so the MMA thread polls ONE barrier per step; free[s] is ONE tcgen05.commit per step that releases the TMEM
// columns to the stagers and the weight slot to the W producer at the same time.
//
// Narrow single-n-tile layers keep their whole packed weight set (<= 96 KB) RESIDENT in shared memory instead of streaming
// one block per step through the ring: the small bulk copies of one SM execute one after the other at ~440 cycles each, which
// capped those layers at ~440 cycles per step (profiles/r02_ring_latency.md).  Layers with few tiles split their K loop over
// several work items (split-K, halo_splitk_reduce_kernel).
//
// The MMA thread is the pacemaker of the CTA (profiles/r01_wait_counters.txt: it never waits, it IS the critical path),
// and what it pays for is its own dependent instruction stream: every value that travels from a vector register to the
// uniform datapath (R2UR, VOTEU) in front of a UTCHMMA / UTCBAR costs tens of cycles.  The loop is therefore written so
// that the ring slot is a COMPILE-TIME constant (switch on step & 3 into four copies of the step body): TMEM operand
// addresses, weight descriptors and barrier addresses are loop-invariant base + immediate, and the only per-step values
// are the accumulator base (changes per tile) and the class mask of the step.
//
// 3xTF32 in two instructions per K8 slice when N <= 64 ("stacked", MODE 1): the weight block holds [W_hi ; W_lo] as 2N
// consecutive rows, D[:, 0:2N] (+)= A_hi * [W_hi ; W_lo] is ONE UMMA of N' = 2N and D[:, N:2N] += A_lo * W_hi the second;
// the epilogue adds the halves.  MODE 2 = three instructions (64 < N <= 128, or four classes of N = 64), MODE 0 = plain
// single-pass TF32.  Variants: PER_TAP (images that are not made of whole 16 x 8 tiles: one 128-pixel TMA box per step
// instead of a halo) and CIN8 (8-channel inputs: four taps x 8 channels per K = 32 step).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "conv_tc.cuh"
#include "conv_tc_ptx.cuh"

namespace demon {

namespace {

constexpr int kRing = 4;                               // operand ring slots (TMEM A slot of 64 columns + weight slot)
#ifndef DEMON_MAX_STAGER_GROUPS
#define DEMON_MAX_STAGER_GROUPS 2
#endif
constexpr int kMaxGroups = DEMON_MAX_STAGER_GROUPS;    // stager groups of four warps (3: 19 warps, 96 registers per thread)
constexpr int kMaxThreads = 32 * (2 + 4 * kMaxGroups + 4 + 1);   // 15 warps with two stager groups
constexpr int kMaxAStages = 4;
constexpr int kTileH = 16, kTileW = 8;
constexpr int kMaxPlanes = 4;
constexpr int kEpiRowBytes = 144;                  // 32 fp32 + 16 B pad: the epilogue's transpose buffer, conflict free both ways
constexpr int kEpiStageBytes = 32 * kEpiRowBytes;  // per epilogue warp

struct HaloPlane {
  int c_off;      // coordinate offset in dim 0 (rx * in_pitch)
  int ry;         // coordinate in dim 2
  int qx_min, qy_min;
  int cols, rows; // halo box
  int smem_off;   // byte offset of the plane inside the A region (1024-aligned)
  int bytes;      // rows * cols * 128
};

struct HaloParams {
  int tiles_x, tiles_y, B, n_tiles, total_tiles;
  uint32_t mul_n_tiles, mul_tiles_x, mul_tiles_y;   // fast_div multipliers
  int k_chunks, nplanes, nsteps, nclass;
  // split-K: a work ITEM is (tile, z), z < ksplit: kc_split = k_chunks / ksplit chunks of the tile's K loop, accumulated into
  // partial buffer z (out + z * part_stride, bias-free, no activation); halo_splitk_reduce_kernel sums the buffers.
  // ksplit == 1: an item is a tile and the epilogue writes the layer's output directly.
  int ksplit, kc_split, total_items;
  uint32_t mul_ksplit;
  long long part_stride;
  int ngroups;                                      // stager groups (2 or 3)
  // per step (one input shift of one 32-channel chunk)
  int st_plane[kMaxTaps], st_aoff[kMaxTaps];        // halo mode: plane and byte offset of the shift's first pixel
  int st_cmask[kMaxTaps], st_fmask[kMaxTaps];       // classes multiplied in this step / classes that start their sum here (chunk 0)
  int st_woff[kMaxTaps], st_wbytes[kMaxTaps];       // weight blocks of the step inside a chunk's weight image
  // per-tap mode (layers whose images are not made of whole 16x8 tiles): the A stage is ONE shift's 128-pixel tile
  // (tb images x th rows x tw columns, one TMA box per step, box coordinates carry the shift)
  int per_tap, tw, th, tb, tiles_b, aempty_count;
  int tap_c[kMaxTaps], tap_qx[kMaxTaps], tap_ry[kMaxTaps], tap_qy[kMaxTaps];
  // 8-channel mode (Cin == 8, e.g. the image pair): a pixel is 32 bytes in shared memory (no swizzle), the halo is loaded
  // once per tile, and one K = 32 step gathers FOUR taps x 8 channels into the TMEM A operand (`nsteps` then counts these
  // groups, g_plane / g_aoff describe the real taps, -1 = missing tap -> zero columns)
  int cin8, ntaps_real;
  int g_plane[kMaxTaps], g_aoff[kMaxTaps];
  HaloPlane planes[kMaxPlanes];
  int a_region_bytes;   // one halo stage (all planes)
  int sa;               // A (halo, shared memory) stages
  int w_chunk_bytes;    // weight bytes of one (n tile, chunk): sum of st_wbytes
  int w_stage_bytes;    // weight ring slot = the largest step
  int w_region_bytes;   // shared memory of the weights: kRing slots, or the whole layer when resident
  int w_resident;       // 1: ALL weight blocks of the (single) n tile live in shared memory for the whole kernel, loaded once
  int cls_bytes;        // one class block: [W_hi ; W_lo] (3xTF32) or W alone
  int n_tile, nsplit, tmem_cols;
  int nbuf;             // accumulator buffers (2 = epilogue overlaps the next tile; 1 when TMEM is short)
  int mode;             // 0 single pass, 1 stacked 3xTF32, 2 three-instruction 3xTF32
  int acc_w;            // accumulator columns per class: 2 * n_tile in stacked mode ([big | small] terms), n_tile otherwise
  const unsigned char* w;
  float* out;
  int out_pitch, Ho, Wo, Hfull, Wfull, osy, osx, Cout;
  int cls_ooy[4], cls_oox[4];
  const float* bias;
  int leaky;
  int* err;
  long long* timing;   // optional [gridDim.x][16] wait-cycle counters (debug), nullptr in production
};

struct HaloMaps {
  CUtensorMap m[kMaxPlanes];
};

__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Bounded wait, fully inline (a call in the MMA thread's loop would force every loop-invariant operand of the UTCHMMAs back
// through a vector register and an R2UR per step).  Fast path: one try_wait.  Slow path: spin with a cycle budget; the
// cycles spent there go to `acc` (the wait counters of the debug timing cost nothing on the fast path).  On timeout the
// error flag is raised and the kernel runs to completion with garbage; the host turns the flag into DEMON_E_STATE.
__device__ __forceinline__ void wait_t(uint32_t bar, uint32_t parity, int* err, long long& acc, bool timed) {
  // (a try_wait on a pending phase may suspend the thread in hardware and still return true: with `timed` the whole
  // call is bracketed, otherwise only the spin path is)
  const long long t_in = timed ? clock64() : 0;
  if (mbar_try(bar, parity)) {
    if (timed) acc += clock64() - t_in;
    return;
  }
  const long long t0 = clock64();
  for (;;) {
    if (mbar_try(bar, parity)) break;
    if (clock64() - t0 > kTimeoutCycles || *reinterpret_cast<volatile int*>(err) != 0) {
      atomicExch(err, 1);
      break;
    }
  }
  acc += clock64() - (timed ? t_in : t0);
}

// x / d for small d by one multiply (mul = 2^32 / d + 1, exact while x * d < 2^32; d == 1 -> mul = 0): the tile decode
// runs once per tile in four roles, and integer division costs ~25 instructions on a thread that has few to spare
__device__ __forceinline__ int fast_div(int x, int d, uint32_t mul) { return mul ? (int)__umulhi((uint32_t)x, mul) : x; }

template <bool PER_TAP>
__device__ __forceinline__ void halo_decode_tile(const HaloParams& p, int tile, int& nt, int& n, int& y0, int& x0) {
  int m = fast_div(tile, p.n_tiles, p.mul_n_tiles);
  nt = tile - m * p.n_tiles;
  const int m1 = fast_div(m, p.tiles_x, p.mul_tiles_x);
  const int xb = m - m1 * p.tiles_x;
  n = fast_div(m1, p.tiles_y, p.mul_tiles_y);
  const int yb = m1 - n * p.tiles_y;
  if (PER_TAP) { n *= p.tb; y0 = yb * p.th; x0 = xb * p.tw; }
  else { y0 = yb * kTileH; x0 = xb * kTileW; }
}

// ---- MMA thread ---------------------------------------------------------------------------------------------------------
// loop invariants of the MMA thread (all of them end up as uniform registers or immediates)
struct MmaCtx {
  uint32_t full0, free0;      // barrier rings
  uint32_t t_ring;            // TMEM address of ring slot 0 (A_hi at +0, A_lo at +32)
  uint32_t w_ring;            // shared-memory address of weight slot 0
  uint32_t w_stage_bytes, cls_bytes;
  uint32_t w_resident;        // weights resident: the block of step l of a tile is at w_ring + l * w_stage_bytes
  uint32_t n_tile, acc_w;
  uint32_t idesc, idesc2;
  int* err;
#ifdef DEMON_TC_TIMING_FULL
  mutable long long* evlog;   // CTA 0: clock64 per event of the first 64 global steps (rows of 64 behind the wait counters)
  mutable int ev_step;
#endif
};

// all MMAs of ONE class block of a step: K = 32 as four K8 slices
template <int MODE>
__device__ __forceinline__ void issue_block(const MmaCtx& c, uint32_t d, uint32_t a_hi, uint32_t wb, bool fresh) {
  const uint64_t w_hi = umma_desc_sw128_sbo(wb, 1024);
  const uint32_t a_lo = a_hi + 32;
  if (MODE == 1) {
    // D[:, 0:2N] (+)= A_hi * [W_hi ; W_lo]  (one UMMA of N' = 2N: big term | first small term);  D[:, N:2N] += A_lo * W_hi
    if (fresh) umma_tf32_ts(d, a_hi, w_hi, c.idesc2, 0u); else umma_tf32_ts(d, a_hi, w_hi, c.idesc2, 1u);
    umma_tf32_ts(d + c.n_tile, a_lo, w_hi, c.idesc, 1u);
#pragma unroll
    for (int j = 1; j < 4; ++j) {   // + 8 TMEM columns / + 32 bytes inside the swizzled weight row per K8 slice
      umma_tf32_ts(d, a_hi + 8 * j, w_hi + (uint64_t)(2 * j), c.idesc2, 1u);
      umma_tf32_ts(d + c.n_tile, a_lo + 8 * j, w_hi + (uint64_t)(2 * j), c.idesc, 1u);
    }
  } else if (MODE == 2) {
    const uint64_t w_lo = umma_desc_sw128_sbo(wb + c.n_tile * 128u, 1024);
    if (fresh) umma_tf32_ts(d, a_hi, w_lo, c.idesc, 0u); else umma_tf32_ts(d, a_hi, w_lo, c.idesc, 1u);
    umma_tf32_ts(d, a_lo, w_hi, c.idesc, 1u);
    umma_tf32_ts(d, a_hi, w_hi, c.idesc, 1u);
#pragma unroll
    for (int j = 1; j < 4; ++j) {
      umma_tf32_ts(d, a_hi + 8 * j, w_lo + (uint64_t)(2 * j), c.idesc, 1u);
      umma_tf32_ts(d, a_lo + 8 * j, w_hi + (uint64_t)(2 * j), c.idesc, 1u);
      umma_tf32_ts(d, a_hi + 8 * j, w_hi + (uint64_t)(2 * j), c.idesc, 1u);
    }
  } else {
    if (fresh) umma_tf32_ts(d, a_hi, w_hi, c.idesc, 0u); else umma_tf32_ts(d, a_hi, w_hi, c.idesc, 1u);
#pragma unroll
    for (int j = 1; j < 4; ++j) umma_tf32_ts(d, a_hi + 8 * j, w_hi + (uint64_t)(2 * j), c.idesc, 1u);
  }
}

// One step with the ring slot as a compile-time constant (the main loop is unrolled by kRing, so the slot is the
// position in the unrolled body and every operand address is a loop-invariant base plus an immediate).
//   - wait for full[SLOT] unless the early poll issued during the previous step already saw the phase complete
//   - poll full[SLOT + 1] (mbarrier.test_wait: never suspends the thread) BEFORE issuing, so that its latency overlaps the
//     UTCHMMA issue
//   - the step's MMAs: single-class layers (MULTI == false) one block, fresh only at the first step of a tile;
//     transposed convolutions walk the class mask of the step
//   - ONE commit: frees the TMEM columns (stagers) and the weight slot (W producer)
template <int MODE, int SLOT, bool MULTI>
__device__ __forceinline__ bool mma_step(const MmaCtx& c, uint32_t d_base, uint32_t par, uint32_t npar, bool rdy, bool fresh, uint32_t cm,
                                         uint32_t fm, long long& w_full, int l) {
#ifdef DEMON_TC_TIMING_FULL
  // diagnostic build: every wait bracketed by clock reads, and an event log of CTA 0's first 64 steps (tools/bench_conv.py prints it)
  if (c.evlog && c.ev_step < 64) { c.evlog[512 + c.ev_step] = clock64(); c.evlog[576 + c.ev_step] = rdy ? 1 : 0; }   // step entered, early probe result
  if (!rdy) wait_t(c.full0 + 8 * SLOT, par, c.err, w_full, true);
  if (c.evlog && c.ev_step < 64) c.evlog[c.ev_step] = clock64();         // full[] of the step observed
  tc_fence_after();
  if (c.evlog && c.ev_step < 64) c.evlog[384 + c.ev_step] = clock64();   // after tcgen05.fence::after_thread_sync
#else
  if (!rdy) wait_t(c.full0 + 8 * SLOT, par, c.err, w_full, false);   // (the MMA thread's waits are never bracketed: no clock reads on its fast path)
  tc_fence_after();
#endif
  const bool next_rdy = mbar_test(c.full0 + 8 * ((SLOT + 1) & (kRing - 1)), npar);
  const uint32_t a_hi = c.t_ring + (uint32_t)(SLOT * 64);
  uint32_t wb = c.w_ring + (c.w_resident ? (uint32_t)l : (uint32_t)SLOT) * c.w_stage_bytes;
  if (!MULTI) {
    issue_block<MODE>(c, d_base, a_hi, wb, fresh);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (cm & (1u << k)) {
        issue_block<MODE>(c, d_base + (uint32_t)k * c.acc_w, a_hi, wb, (fm & (1u << k)) != 0);
        wb += c.cls_bytes;
      }
  }
#ifdef DEMON_TC_TIMING_FULL
  if (c.evlog && c.ev_step < 64) c.evlog[448 + c.ev_step] = clock64();   // the step's MMAs issued
#endif
  umma_commit(c.free0 + 8 * SLOT);
#ifdef DEMON_TC_TIMING_FULL
  if (c.evlog && c.ev_step < 64) c.evlog[64 + c.ev_step] = clock64();    // the step's commit issued
  ++c.ev_step;
#endif
  return next_rdy;
}

// The MMA thread's whole main loop.  State per step is kept minimal: l (step inside the tile), the early-poll flag, the
// ring parity; everything else is loop invariant.  The tile boundary is handled inline after the step that ends a tile.
template <int MODE, bool MULTI>
__device__ __forceinline__ void mma_main_loop(const HaloParams& p, const MmaCtx& c, uint32_t tmem_base, int acc_cols, uint32_t cfull0,
                                              uint32_t cempty0, long long& w_cempty, long long& w_full) {
  const int steps_per_tile = p.kc_split * p.nsteps;   // steps of one work item
  const int nsteps = p.nsteps;
  const int nbuf = p.nbuf;
  uint64_t cmasks = 0, fmasks = 0;   // class / fresh masks of the steps of a chunk, four bits per step (MULTI only)
  if (MULTI)
    for (int t = 0; t < nsteps; ++t) { cmasks |= (uint64_t)p.st_cmask[t] << (4 * t); fmasks |= (uint64_t)p.st_fmask[t] << (4 * t); }
  int tile = blockIdx.x;                               // (work item index)
  if (tile >= p.total_items) return;
  int it = 0, l = 0, t = 0;
  uint32_t par = 0;
  wait_t(cempty0, 1u, c.err, w_cempty, false);          // accumulator buffer 0, first use
  uint32_t d_base = tmem_base;
  uint32_t cfull_bar = cfull0;
  bool rdy = mbar_test(c.full0, 0);

#define DEMON_MMA_STEP(SLOT)                                                                                              \
  {                                                                                                                       \
    uint32_t cm = 1u, fm = 0u;                                                                                            \
    if (MULTI) {                                                                                                          \
      cm = (uint32_t)(cmasks >> (4 * t)) & 15u;                                                                           \
      fm = (l < nsteps) ? ((uint32_t)(fmasks >> (4 * t)) & 15u) : 0u;                                                     \
      if (++t == nsteps) t = 0;                                                                                           \
    }                                                                                                                     \
    rdy = mma_step<MODE, SLOT, MULTI>(c, d_base, par, (SLOT == kRing - 1) ? (par ^ 1u) : par, rdy, l == 0, cm, fm, w_full, l); \
    if (++l == steps_per_tile) {                                                                                          \
      umma_commit(cfull_bar);   /* the tile's accumulators are complete once everything issued so far has retired */      \
      tile += gridDim.x;                                                                                                  \
      if (tile >= p.total_items) return;                                                                                  \
      ++it;                                                                                                               \
      l = 0;                                                                                                              \
      const int a = (nbuf == 2) ? (it & 1) : 0;                                                                           \
      const uint32_t cphase = (nbuf == 2) ? ((it >> 1) & 1) : (it & 1);                                                   \
      wait_t(cempty0 + 8 * a, cphase ^ 1u, c.err, w_cempty, false);                                                       \
      d_base = tmem_base + (uint32_t)(a * acc_cols);                                                                      \
      cfull_bar = cfull0 + 8 * a;                                                                                         \
    }                                                                                                                     \
  }
  for (;;) {
    DEMON_MMA_STEP(0)
    DEMON_MMA_STEP(1)
    DEMON_MMA_STEP(2)
    DEMON_MMA_STEP(3)
    par ^= 1u;
  }
#undef DEMON_MMA_STEP
}

template <bool PER_TAP, bool CIN8, int MODE>
__global__ void __launch_bounds__(kMaxThreads, 1) conv_tc_halo_kernel(const __grid_constant__ HaloMaps maps, const HaloParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // bars: A_full[4], A_empty[4], full[4], free[4], accum_full[2], accum_empty[2]
  __shared__ __align__(8) uint64_t bars[2 * kMaxAStages + 2 * kRing + 4 + 1];   // (+ W_resident)
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_stager_warps = 4 * p.ngroups;
  const int epi_warp0 = 2 + n_stager_warps;
  const int w_warp = epi_warp0 + 4;
  const uint32_t afull0 = smem_u32(&bars[0]), aempty0 = smem_u32(&bars[kMaxAStages]);
  const uint32_t full0 = smem_u32(&bars[2 * kMaxAStages]), free0 = smem_u32(&bars[2 * kMaxAStages + kRing]);
  const uint32_t cfull0 = smem_u32(&bars[2 * kMaxAStages + 2 * kRing]);
  const uint32_t cempty0 = smem_u32(&bars[2 * kMaxAStages + 2 * kRing + 2]);
  const uint32_t wres_bar = smem_u32(&bars[2 * kMaxAStages + 2 * kRing + 4]);
  const int a_stage_bytes = p.a_region_bytes;
  unsigned char* w_ring = smem + (size_t)p.sa * a_stage_bytes;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nplanes; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[i]) : "memory");
    for (int s = 0; s < kMaxAStages; ++s) {
      mbar_init(afull0 + 8 * s, 1);
      mbar_init(aempty0 + 8 * s, (uint32_t)p.aempty_count);
    }
    for (int s = 0; s < kRing; ++s) {
      mbar_init(full0 + 8 * s, p.w_resident ? 4 : 4 + 1);   // the four warps of the stager group that owns the step (+ the W producer's expect_tx)
      mbar_init(free0 + 8 * s, 1);       // the MMA thread's commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(cfull0 + 8 * s, 1);
      mbar_init(cempty0 + 8 * s, 4);
    }
    mbar_init(wres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int acc_cols = p.nclass * p.acc_w;           // columns of one accumulator buffer
  const uint32_t t_ring = tmem_base + (uint32_t)(p.nbuf * acc_cols);   // A-operand ring: kRing slots of 64 columns (hi | lo)
  const bool timed = p.timing != nullptr;
  // Programmatic dependent launch: from here on the next kernel of the stream may become resident (on SMs this grid has
  // left).  This kernel's own prologue above ran while its predecessor was still draining; the roles that read what the
  // predecessor wrote (A producer) or write global memory (epilogue) wait for it below, the W producer does not (weights
  // are constants) and fills the weight ring in the meantime.
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== A producer: one halo box per plane and 32-channel chunk ======================================================
    if (elect_one_sync()) {
      int sa = 0;
      uint32_t pa = 0;
      uint32_t a_bytes = 0;
      long long w_aempty = 0;
      const long long t_begin = clock64();
      for (int i = 0; i < p.nplanes; ++i) a_bytes += (uint32_t)p.planes[i].bytes;
      pdl_wait();   // the input is the previous kernel's output
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const int tile = fast_div(item, p.ksplit, p.mul_ksplit);
        const int kc0 = (item - tile * p.ksplit) * p.kc_split;
        int nt, n, y0, x0;
        halo_decode_tile<PER_TAP>(p, tile, nt, n, y0, x0);
        for (int kc = kc0; kc < kc0 + p.kc_split; ++kc) {
          if (PER_TAP) {
            for (int t = 0; t < p.nsteps; ++t) {
              wait_t(aempty0 + 8 * sa, pa ^ 1, p.err, w_aempty, timed);
              mbar_expect_tx(afull0 + 8 * sa, 128u * 128u);
              tma_load_5d(smem_u32(smem + (size_t)sa * a_stage_bytes), &maps.m[0], afull0 + 8 * sa, p.tap_c[t] + kc * 32, x0 + p.tap_qx[t],
                          p.tap_ry[t], y0 + p.tap_qy[t], n);
              if (++sa == p.sa) { sa = 0; pa ^= 1; }
            }
            continue;
          }
          wait_t(aempty0 + 8 * sa, pa ^ 1, p.err, w_aempty, timed);
          const uint32_t abase = smem_u32(smem + (size_t)sa * a_stage_bytes);
          mbar_expect_tx(afull0 + 8 * sa, a_bytes);
          for (int i = 0; i < p.nplanes; ++i) {
            const HaloPlane& pl = p.planes[i];
            tma_load_5d(abase + pl.smem_off, &maps.m[i], afull0 + 8 * sa, pl.c_off + kc * 32, x0 + pl.qx_min, pl.ry, y0 + pl.qy_min, n);
          }
          if (++sa == p.sa) { sa = 0; pa ^= 1; }
        }
      }
      if (timed) { p.timing[blockIdx.x * 16 + 0] = w_aempty; p.timing[blockIdx.x * 16 + 8] = clock64() - t_begin; }
    }
  } else if (warp == w_warp) {
    // ===== W producer: the weight blocks of every step into the step's ring slot ==========================================
    // The copy's bytes complete on the step's `full` barrier (the one the stager group arrives on).  Both producers of a
    // slot wait for free[slot] of the previous round before they arrive for the next one, so all five arrivals and the
    // bytes of a step belong to the same barrier phase.
    const bool w_elected = elect_one_sync();
    if (w_elected && p.w_resident) {
      // Resident weights (narrow single-n-tile layers: the whole layer's packed weights are 24 .. 96 KB): a few large copies
      // at the start instead of one 4 .. 8 KB cp.async.bulk per step.  The small copies of one SM execute one after the
      // other at ~440 cycles each (event log in profiles/r02_ring_latency.md: a copy lands ~2900 cycles after it is issued
      // with six ahead of it), and THAT is the ~440-cycle step of the narrow layers, whatever N and whatever the ring depth.
      const uint32_t total = (uint32_t)(p.k_chunks * p.w_chunk_bytes);
      mbar_expect_tx(wres_bar, total);
      for (int kc = 0; kc < p.k_chunks; ++kc)
        bulk_load(smem_u32(w_ring + (size_t)kc * p.w_chunk_bytes), p.w + (size_t)kc * p.w_chunk_bytes, (uint32_t)p.w_chunk_bytes, wres_bar);
    } else if (w_elected) {
      int slot = 0;
      uint32_t use = 0;
      long long w_free = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const int tile = fast_div(item, p.ksplit, p.mul_ksplit);
        const int kc0 = (item - tile * p.ksplit) * p.kc_split;
        const int nt = tile - fast_div(tile, p.n_tiles, p.mul_n_tiles) * p.n_tiles;
        const unsigned char* wsrc = p.w + ((size_t)nt * p.k_chunks + kc0) * p.w_chunk_bytes;
        for (int kc = 0; kc < p.kc_split; ++kc, wsrc += p.w_chunk_bytes) {
          for (int t = 0; t < p.nsteps; ++t) {
            wait_t(free0 + 8 * slot, use ^ 1, p.err, w_free, timed);
            const uint32_t bytes = (uint32_t)p.st_wbytes[t];
            mbar_expect_tx(full0 + 8 * slot, bytes);
            bulk_load(smem_u32(w_ring + (size_t)slot * p.w_stage_bytes), wsrc + p.st_woff[t], bytes, full0 + 8 * slot);
            if (++slot == kRing) { slot = 0; use ^= 1; }
          }
        }
      }
      if (timed) p.timing[blockIdx.x * 16 + 1] = w_free;
    }
  } else if (warp == 1) {
    // ===== MMA issuer: A operand from TMEM (staged by the stager warps), B operand (weights) from shared memory =========
    if (elect_one_sync()) {
      MmaCtx c;
      c.full0 = full0; c.free0 = free0; c.t_ring = t_ring; c.w_ring = smem_u32(w_ring);
      c.w_stage_bytes = (uint32_t)p.w_stage_bytes; c.cls_bytes = (uint32_t)p.cls_bytes;
      c.n_tile = (uint32_t)p.n_tile; c.acc_w = (uint32_t)p.acc_w;
      c.idesc = umma_idesc_tf32(p.n_tile); c.idesc2 = umma_idesc_tf32(2 * p.n_tile);
      c.err = p.err;
#ifdef DEMON_TC_TIMING_FULL
      c.evlog = (timed && blockIdx.x == 0) ? p.timing + 160 * 16 : nullptr;
      c.ev_step = 0;
#endif
      c.w_resident = (uint32_t)p.w_resident;
      long long w_cempty = 0, w_full = 0;
      const long long t_begin = clock64();
      if (p.w_resident) wait_t(wres_bar, 0u, p.err, w_full, false);
      if (p.nclass == 1) mma_main_loop<MODE, false>(p, c, tmem_base, acc_cols, cfull0, cempty0, w_cempty, w_full);
      else mma_main_loop<MODE, true>(p, c, tmem_base, acc_cols, cfull0, cempty0, w_cempty, w_full);
      if (timed) {
        long long* tm = p.timing + blockIdx.x * 16;
        tm[2] = w_cempty; tm[3] = w_full; tm[9] = clock64() - t_begin;
      }
    }
  } else if (warp < epi_warp0) {
    // ===== stagers: halo image (shared memory) -> A operand of one shift in TMEM ========================================
    // Thread (quadrant q, lane l) owns GEMM row m = 32q + l = tile pixel (m / 8, m % 8).  For every step it reads its
    // (shifted) 128-byte pixel out of the swizzled halo image -- conflict free: the 8 lanes of a quarter warp hit 8
    // different 16-byte columns -- and writes the raw fp32 values (A_hi: the tensor core ignores the low 13 mantissa bits)
    // and A_lo = A - trunc_tf32(A) to its TMEM lane.  The groups of four warps take the steps in turns.
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int g8 = m >> 3, r = m & 7;
    const int ngroups = p.ngroups;
    long long w_safull = 0, w_free = 0;
#ifdef DEMON_TC_TIMING_FULL
    long long d_load = 0, d_store = 0, d_arrive = 0;
#endif
    const long long t_begin = clock64();
    // The group visits only its OWN steps: global step gs = grp, grp + G, grp + 2G, ...; t_own is that step's index
    // inside the current chunk.  Ring slot and parity follow from gs; in per-tap mode so does the A stage (one per step).
    uint32_t gs = (uint32_t)grp;
    int t_own = grp;
    int sa = PER_TAP ? (grp % p.sa) : 0;          // halo mode: stage of the current chunk; per-tap: stage of step gs
    uint32_t pa = PER_TAP ? (uint32_t)((grp / p.sa) & 1) : 0u;
    const int nsteps = p.nsteps;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      for (int kc = 0; kc < p.kc_split; ++kc) {
        uint32_t abase = 0;
        if (!PER_TAP) {
          wait_t(afull0 + 8 * sa, pa, p.err, w_safull, timed);
          __syncwarp();
          abase = smem_u32(smem + (size_t)sa * a_stage_bytes);
        }
        for (; t_own < nsteps; t_own += ngroups, gs += (uint32_t)ngroups) {
          const int t = t_own;
          const int slot_cur = (int)(gs & (kRing - 1));
          const uint32_t use_cur = (gs >> 2) & 1u;
          const int sa_cur = sa;
          uint32_t row;
#ifdef DEMON_TC_TIMING_FULL
          const long long c0 = clock64();
#endif
          if (PER_TAP) {
            wait_t(afull0 + 8 * sa_cur, pa, p.err, w_safull, timed);
            __syncwarp();
            row = smem_u32(smem + (size_t)sa_cur * a_stage_bytes) + (uint32_t)(g8 * 1024 + r * 128);
            sa += ngroups;                                  // the A stage of this group's next step
            while (sa >= p.sa) { sa -= p.sa; pa ^= 1u; }
          } else if (!CIN8) {
            const HaloPlane& pl = p.planes[p.st_plane[t]];
            row = abase + (uint32_t)(pl.smem_off + p.st_aoff[t] + g8 * pl.cols * 128 + r * 128);
          } else {
            row = 0;
          }
          uint32_t hi[32], lo[32];   // the shared-memory loads are in flight while the slot's barrier is polled
          if (CIN8) {
            // four taps x 8 channels -> the 32 columns of this K step; a pixel is 32 contiguous bytes
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int rt = 4 * t + j;
              const int pi = (rt < p.ntaps_real) ? p.g_plane[rt] : -1;
              float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
              if (pi >= 0) {
                const HaloPlane& pl = p.planes[pi];
                const uint32_t px = abase + (uint32_t)(pl.smem_off + p.g_aoff[rt] + (g8 * pl.cols + r) * 32);
                v0 = lds128(px); v1 = lds128(px + 16);
              }
              hi[8 * j + 0] = __float_as_uint(v0.x); hi[8 * j + 1] = __float_as_uint(v0.y);
              hi[8 * j + 2] = __float_as_uint(v0.z); hi[8 * j + 3] = __float_as_uint(v0.w);
              hi[8 * j + 4] = __float_as_uint(v1.x); hi[8 * j + 5] = __float_as_uint(v1.y);
              hi[8 * j + 6] = __float_as_uint(v1.z); hi[8 * j + 7] = __float_as_uint(v1.w);
            }
          } else {
            const uint32_t phase = (row >> 7) & 7u;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 v = lds128(row + ((c ^ phase) << 4));
              hi[4 * c + 0] = __float_as_uint(v.x); hi[4 * c + 1] = __float_as_uint(v.y);
              hi[4 * c + 2] = __float_as_uint(v.z); hi[4 * c + 3] = __float_as_uint(v.w);
            }
          }
#ifdef DEMON_TC_TIMING_FULL
          const long long c1 = clock64();
#endif
          wait_t(free0 + 8 * slot_cur, use_cur ^ 1, p.err, w_free, timed);
          __syncwarp();
          tc_fence_after();
#ifdef DEMON_TC_TIMING_FULL
          const long long c2 = clock64();
          if (timed && blockIdx.x == 0 && q == 0 && lane == 0 && gs < 64) p.timing[160 * 16 + 128 + gs] = c2;   // free[] observed
#endif
          const uint32_t taddr = t_ring + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot_cur * 64);
          tmem_st_x32(taddr, hi);
          if (MODE != 0) {
#pragma unroll
            for (int c = 0; c < 32; c += 2) tf32_lo2(hi[c], hi[c + 1], lo[c], lo[c + 1]);
            tmem_st_x32(taddr + 32, lo);
          }
          tmem_st_wait();
#ifdef DEMON_TC_TIMING_FULL
          const long long c3 = clock64();
#endif
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(full0 + 8 * slot_cur);
            if (PER_TAP) mbar_arrive(aempty0 + 8 * sa_cur);   // the shift's tile has been consumed: stage back to the producer
          }
#ifdef DEMON_TC_TIMING_FULL
          d_load += c1 - c0; d_store += c3 - c2; d_arrive += clock64() - c3;
          if (timed && blockIdx.x == 0 && q == 0 && lane == 0 && gs < 64) {
            p.timing[160 * 16 + 192 + gs] = clock64();   // arrived on full[]
            p.timing[160 * 16 + 256 + gs] = c0;          // started on the step (address math, shared loads)
            p.timing[160 * 16 + 320 + gs] = c3;          // tcgen05.wait::st returned
          }
#endif
        }
        t_own -= nsteps;   // index of the next own step inside the NEXT chunk
        if (!PER_TAP) {
          __syncwarp();
          if (lane == 0) mbar_arrive(aempty0 + 8 * sa);   // this warp is done reading the halo stage
          if (++sa == p.sa) { sa = 0; pa ^= 1; }
        }
      }
    }
    if (timed && lane == 0 && q == 0) {
      if (grp == 0) { p.timing[blockIdx.x * 16 + 6] = w_safull; p.timing[blockIdx.x * 16 + 4] = w_free; p.timing[blockIdx.x * 16 + 10] = clock64() - t_begin; }
      else if (grp == 1) p.timing[blockIdx.x * 16 + 12] = w_free;
#ifdef DEMON_TC_TIMING_FULL
      if (grp == 0) { p.timing[blockIdx.x * 16 + 13] = d_load; p.timing[blockIdx.x * 16 + 14] = d_store; p.timing[blockIdx.x * 16 + 15] = d_arrive; }
#endif
    }
  } else if (warp < w_warp) {
    // ===== epilogue ======================================================================================================
    // After the TMEM load a thread holds one output pixel (GEMM row) x 32 channels, and consecutive pixels are a whole
    // pixel pitch apart in global memory: storing rows directly costs 32 memory wavefronts per store instruction.  Each
    // warp therefore transposes its 32 x 32 block through shared memory (row pitch 144 B: conflict free both ways) and
    // stores with lane = (row % 4, 16-byte chunk), i.e. four pixels x 128 contiguous bytes per instruction; the bias is
    // then one float4 per lane and chunk.
    const int q = warp & 3;
    const int sub = lane >> 3, chunk = lane & 7;
    const uint32_t stg = smem_u32(w_ring + (size_t)p.w_region_bytes) + (uint32_t)(q * kEpiStageBytes);
    int rowoff[8];          // element offset of row 4 i + sub of this warp's block inside the output tile
    uint32_t rowpos[8];     // its (yl, xl, nl) for the bounds test
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = q * 32 + 4 * i + sub;
      const int xl = PER_TAP ? (m % p.tw) : (m & 7);
      const int yl = PER_TAP ? ((m / p.tw) % p.th) : (m >> 3);
      const int nl = PER_TAP ? (m / (p.tw * p.th)) : 0;
      rowoff[i] = ((nl * p.Hfull + yl * p.osy) * p.Wfull + xl * p.osx) * p.out_pitch;
      rowpos[i] = (uint32_t)yl | ((uint32_t)xl << 10) | ((uint32_t)nl << 20);
    }
    const float slope = p.leaky ? 0.1f : 1.0f;
    long long w_cfull = 0;
    const long long t_begin = clock64();
    int it = 0;
    pdl_wait();   // the output buffer may still be read (or written) by the previous kernel
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      const int tile = fast_div(item, p.ksplit, p.mul_ksplit);
      const int z = item - tile * p.ksplit;
      int nt, n, y0, x0;
      halo_decode_tile<PER_TAP>(p, tile, nt, n, y0, x0);
      const int a = (p.nbuf == 2) ? (it & 1) : 0;
      wait_t(cfull0 + 8 * a, (p.nbuf == 2) ? ((it >> 1) & 1) : (it & 1), p.err, w_cfull, timed);
      __syncwarp();
      tc_fence_after();
      uint32_t rowmask = 0xFFu;
      const bool whole = PER_TAP ? (y0 + p.th <= p.Ho && x0 + p.tw <= p.Wo && n + p.tb <= p.B) : (y0 + kTileH <= p.Ho && x0 + kTileW <= p.Wo);
      if (!whole) {
        rowmask = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int yl = (int)(rowpos[i] & 1023u), xl = (int)((rowpos[i] >> 10) & 1023u), nl = (int)(rowpos[i] >> 20);
          if (y0 + yl < p.Ho && x0 + xl < p.Wo && n + nl < p.B) rowmask |= 1u << i;
        }
      }
      float* tile_out = p.out + (size_t)z * p.part_stride + ((size_t)(n * p.Hfull + y0 * p.osy) * p.Wfull + x0 * p.osx) * p.out_pitch;
      const int cbase = nt * p.n_tile;
      for (int cls = 0; cls < p.nclass; ++cls) {
        float* cls_out = tile_out + (size_t)(p.cls_ooy[cls] * p.Wfull + p.cls_oox[cls]) * p.out_pitch;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * acc_cols + cls * p.acc_w);
        for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
          uint32_t v[32];
          const int ncol = (p.n_tile - c0) >= 32 ? 32 : 16;
          if (ncol == 32) tmem_ld_x32(t_row + c0, v); else tmem_ld_x16(t_row + c0, v);
          if (MODE == 1) {   // big term + small terms
            uint32_t u[32];
            if (ncol == 32) tmem_ld_x32(t_row + p.n_tile + c0, u); else tmem_ld_x16(t_row + p.n_tile + c0, u);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
          }
          tmem_ld_wait();
          if (cls == p.nclass - 1 && c0 + 32 >= p.n_tile) {   // last TMEM read of the tile: hand the accumulators back now
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(cempty0 + 8 * a);
          }
          __syncwarp();   // the previous block has been read back
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (4 * c < ncol)
              sts128(stg + (uint32_t)(lane * kEpiRowBytes + c * 16),
                     make_float4(__uint_as_float(v[4 * c + 0]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                                 __uint_as_float(v[4 * c + 3])));
          __syncwarp();
          const int col = cbase + c0 + 4 * chunk;
          if (4 * chunk < ncol && col < p.Cout) {
            const float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);   // (null: split-K partial sums)
            float* colp = cls_out + col;
            const uint32_t sbase = stg + (uint32_t)(sub * kEpiRowBytes + chunk * 16);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (!((rowmask >> i) & 1u)) continue;
              const float4 x = lds128(sbase + (uint32_t)(4 * i * kEpiRowBytes));
              float4 o;
              o.x = x.x + b.x; o.y = x.y + b.y; o.z = x.z + b.z; o.w = x.w + b.w;
              // max(slope * x, x) with slope 0.1 (leaky ReLU, helpers.py:36-38) or 1 (identity, exact)
              o.x = fmaxf(slope * o.x, o.x); o.y = fmaxf(slope * o.y, o.y);
              o.z = fmaxf(slope * o.z, o.z); o.w = fmaxf(slope * o.w, o.w);
              *reinterpret_cast<float4*>(colp + rowoff[i]) = o;
            }
          }
        }
      }
    }
    if (timed && q == 0 && lane == 0) { p.timing[blockIdx.x * 16 + 7] = w_cfull; p.timing[blockIdx.x * 16 + 11] = clock64() - t_begin; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// Split-K second pass: out[pix][c] = act(bias[c] + sum_z part[z][pix][c]); the partial buffers have the geometry of the
// layer's full output image ([B][Hfull][Wfull][cpitch] floats each), so sub-pixel classes are already interleaved.
__global__ void __launch_bounds__(256) halo_splitk_reduce_kernel(const float* __restrict__ part, long long part_stride, int ksplit,
                                                                 float* __restrict__ out, int out_pitch, int cpitch, int Cout, long long npix,
                                                                 const float* __restrict__ bias, int leaky) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = cpitch >> 2;
  const long long total = npix * c4n;
  const float slope = leaky ? 0.1f : 1.0f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long pix = i / c4n;
    const int c = (int)(i - pix * c4n) * 4;
    if (c >= Cout) continue;
    float4 acc = __ldg(reinterpret_cast<const float4*>(bias + c));
    for (int z = 0; z < ksplit; ++z) {   // fixed order: deterministic
      const float4 v = *reinterpret_cast<const float4*>(part + (size_t)z * part_stride + (size_t)pix * cpitch + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x = fmaxf(slope * acc.x, acc.x); acc.y = fmaxf(slope * acc.y, acc.y);
    acc.z = fmaxf(slope * acc.z, acc.z); acc.w = fmaxf(slope * acc.w, acc.w);
    *reinterpret_cast<float4*>(out + (size_t)pix * out_pitch + c) = acc;
  }
}

int floor_div_h(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------------------
struct HaloPlan {
  HaloParams prm;
  HaloMaps maps;
  int smem_bytes;
  // weight source of every (step, class): tap index inside the class, -1 = class not in the step
  int st_tap[kMaxTaps][4];
};

static int pow2_ceil_h(int v) { int r = 1; while (r < v) r <<= 1; return r; }
static int popcount4(int m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1); }

static int stager_groups() {
  static const int g = []() {
    const char* e = getenv("DEMON_STAGER_GROUPS");
    int v = e ? atoi(e) : 2;
    return (v < 1 || v > kMaxGroups) ? 2 : v;
  }();
  return g;
}

static bool w_resident_enabled() {   // DEMON_W_RESIDENT=0: weights always through the 4-slot ring (A/B measurements)
  static const bool v = []() { const char* e = getenv("DEMON_W_RESIDENT"); return !(e && e[0] == '0'); }();
  return v;
}

static bool splitk_enabled() {   // DEMON_TC_SPLITK=0: never split the K loop of a tensor-core layer (A/B measurements)
  static const bool v = []() { const char* e = getenv("DEMON_TC_SPLITK"); return !(e && e[0] == '0'); }();
  return v;
}

static bool wave_model_enabled() {   // DEMON_WAVE_MODEL=0: N tile by width alone (A/B measurements)
  static const bool v = []() { const char* e = getenv("DEMON_WAVE_MODEL"); return !(e && e[0] == '0'); }();
  return v;
}

// Steps of one chunk: the distinct input shifts of all classes, each with the mask of the classes that use it.
// max_cls: at most this many classes per step (a wide step needs a wide weight slot; a shift is repeated if necessary).
struct ShiftStep { int ry, rx, qy, qx, cmask, tap[4]; };

static bool halo_build(const ConvProblem* probs, int nclass, int nsplit, HaloPlan& plan, bool encode) {
  const ConvProblem& p = probs[0];
  HaloParams& prm = plan.prm;
  const int budget = 224 * 1024 - 4 * kEpiStageBytes;   // dynamic shared memory minus the epilogue's transpose buffers
  // candidates, widest first: classes per step (a wide step needs a wide weight slot), then the N tile cap
  struct Cand { int max_cls, n_cap; };
  std::vector<Cand> cands;
  for (int mc = nclass; mc >= 1; mc = (mc > 2 ? 2 : mc - 1)) cands.push_back({mc, 256});
  cands.push_back({1, 64});
  cands.push_back({1, 32});
  bool found = false;
  for (const Cand& cand : cands) {
    const int max_cls = cand.max_cls;
    memset(&prm, 0, sizeof(prm));
    memset(plan.st_tap, -1, sizeof(plan.st_tap));
    prm.nclass = nclass;
    prm.nsplit = nsplit;
    prm.ngroups = stager_groups();
    prm.cin8 = (p.Cin == 8) ? 1 : 0;
    const int px_bytes = prm.cin8 ? 32 : 128;   // bytes of one pixel of a halo plane in shared memory
    prm.per_tap = ((p.Ho % kTileH) != 0 || (p.Wo % kTileW) != 0) ? 1 : 0;
    if (prm.cin8 && (prm.per_tap || nclass != 1)) return false;
    // ---- steps ---------------------------------------------------------------------------------------------------
    std::vector<ShiftStep> steps;
    for (int c = 0; c < nclass; ++c)
      for (int i = 0; i < probs[c].ntaps; ++i) {
        const int qy = floor_div_h(probs[c].dy[i], p.sy), ry = probs[c].dy[i] - qy * p.sy;
        const int qx = floor_div_h(probs[c].dx[i], p.sx), rx = probs[c].dx[i] - qx * p.sx;
        int si = -1;
        for (size_t k = 0; k < steps.size(); ++k)
          if (steps[k].ry == ry && steps[k].rx == rx && steps[k].qy == qy && steps[k].qx == qx && !((steps[k].cmask >> c) & 1) &&
              popcount4(steps[k].cmask) < max_cls) { si = (int)k; break; }
        if (si < 0) { steps.push_back({ry, rx, qy, qx, 0, {-1, -1, -1, -1}}); si = (int)steps.size() - 1; }
        steps[si].cmask |= 1 << c;
        steps[si].tap[c] = i;
      }
    const int nreal = (int)steps.size();
    if (nreal > kMaxTaps) return false;
    int m_tiles = 0;
    if (prm.per_tap) {
      // one 128-pixel tile per step, tb images x th rows x tw columns
      for (int t = 0; t < nreal; ++t) {
        prm.tap_qy[t] = steps[t].qy; prm.tap_ry[t] = steps[t].ry;
        prm.tap_qx[t] = steps[t].qx; prm.tap_c[t] = steps[t].rx * p.in_pitch;
      }
      prm.nplanes = 1;
      long best_tiles = -1;
      const int tw = std::min(128, pow2_ceil_h(p.Wo));
      for (int th = 1; th * tw <= 128; th <<= 1) {
        const int tb = 128 / (tw * th);
        const long tiles = (long)ceil_div(p.Wo, tw) * ceil_div(p.Ho, th) * ceil_div(p.B, tb);
        if (best_tiles < 0 || tiles <= best_tiles) {
          best_tiles = tiles;
          prm.tw = tw; prm.th = th; prm.tb = tb;
          prm.tiles_x = ceil_div(p.Wo, tw); prm.tiles_y = ceil_div(p.Ho, th); prm.tiles_b = ceil_div(p.B, tb);
        }
      }
      prm.planes[0].cols = prm.tw; prm.planes[0].rows = prm.th; prm.planes[0].bytes = 128 * 128;
      prm.a_region_bytes = 128 * 128;
      prm.aempty_count = 4;     // the four stager warps of the group that consumed the shift release its shared-memory stage
      m_tiles = prm.tiles_x * prm.tiles_y * prm.tiles_b;
    } else {
      prm.aempty_count = 4 * prm.ngroups;
      // planes: shifts grouped by stride parity
      struct PInfo { int ry, rx, qy_min, qy_max, qx_min, qx_max; };
      std::vector<PInfo> pinfo;
      std::vector<int> step_plane(nreal);
      for (int t = 0; t < nreal; ++t) {
        int pi = -1;
        for (size_t k = 0; k < pinfo.size(); ++k)
          if (pinfo[k].ry == steps[t].ry && pinfo[k].rx == steps[t].rx) pi = (int)k;
        if (pi < 0) { pinfo.push_back({steps[t].ry, steps[t].rx, steps[t].qy, steps[t].qy, steps[t].qx, steps[t].qx}); pi = (int)pinfo.size() - 1; }
        PInfo& pl = pinfo[pi];
        pl.qy_min = std::min(pl.qy_min, steps[t].qy); pl.qy_max = std::max(pl.qy_max, steps[t].qy);
        pl.qx_min = std::min(pl.qx_min, steps[t].qx); pl.qx_max = std::max(pl.qx_max, steps[t].qx);
        step_plane[t] = pi;
      }
      if ((int)pinfo.size() > kMaxPlanes) return false;
      prm.nplanes = (int)pinfo.size();
      int off = 0;
      for (int i = 0; i < prm.nplanes; ++i) {
        HaloPlane& pl = prm.planes[i];
        pl.c_off = pinfo[i].rx * p.in_pitch; pl.ry = pinfo[i].ry;
        pl.qx_min = pinfo[i].qx_min; pl.qy_min = pinfo[i].qy_min;
        pl.cols = kTileW + pinfo[i].qx_max - pinfo[i].qx_min;
        pl.rows = kTileH + pinfo[i].qy_max - pinfo[i].qy_min;
        if (pl.cols > 256 || pl.rows > 256) return false;
        pl.bytes = pl.rows * pl.cols * px_bytes;
        pl.smem_off = off;
        off += (pl.bytes + 1023) / 1024 * 1024;
      }
      prm.a_region_bytes = off;
      for (int t = 0; t < nreal; ++t) {
        const HaloPlane& pl = prm.planes[step_plane[t]];
        prm.st_plane[t] = step_plane[t];
        prm.st_aoff[t] = ((steps[t].qy - pl.qy_min) * pl.cols + (steps[t].qx - pl.qx_min)) * px_bytes;
      }
      prm.tiles_x = ceil_div(p.Wo, kTileW); prm.tiles_y = ceil_div(p.Ho, kTileH); prm.tiles_b = p.B;
      m_tiles = prm.tiles_x * prm.tiles_y * p.B;
    }
    // class / fresh masks, weight source of every (step, class)
    int seen = 0;
    for (int t = 0; t < nreal; ++t) {
      prm.st_cmask[t] = steps[t].cmask;
      prm.st_fmask[t] = steps[t].cmask & ~seen;
      seen |= steps[t].cmask;
      for (int c = 0; c < 4; ++c) plan.st_tap[t][c] = steps[t].tap[c];
    }
    prm.nsteps = nreal;
    if (prm.cin8) {   // regroup: one K step = four consecutive taps (nclass == 1: every real step is one tap of class 0)
      prm.ntaps_real = nreal;
      for (int t = 0; t < nreal; ++t) { prm.g_plane[t] = prm.st_plane[t]; prm.g_aoff[t] = prm.st_aoff[t]; }
      prm.nsteps = (nreal + 3) / 4;
      for (int t = 0; t < prm.nsteps; ++t) { prm.st_plane[t] = 0; prm.st_aoff[t] = 0; prm.st_cmask[t] = 1; prm.st_fmask[t] = (t == 0) ? 1 : 0; }
    }
    // ---- TMEM budget (512 columns): nbuf accumulator buffers x nclass x acc_w  +  the A-operand ring, kRing slots of 64
    // columns (A_hi | A_lo of one shift).  acc_w = 2N in stacked 3xTF32 mode (big | small terms side by side), N otherwise.
    const int ring_cols = kRing * 64;
    const int cout16 = (p.Cout + 15) / 16 * 16;
    int n_tile = std::min(std::min(cout16, nsplit == 1 ? 256 : 128), cand.n_cap);
    while (n_tile > 16 && nclass * n_tile + ring_cols > 512) n_tile = (n_tile > 32) ? (n_tile / 2 + 15) / 16 * 16 : n_tile - 16;
    // narrow the N tile (down to 64) while the layer would leave SMs idle
    while (n_tile >= 128 && (n_tile % 32) == 0 && (long)m_tiles * ceil_div(p.Cout, n_tile) < 148) n_tile /= 2;
    // wave quantisation: a persistent grid of 148 CTAs needs ceil(tiles / 148) rounds.  A step costs ~800 cycles at N = 128
    // (three-instruction 3xTF32) and ~470 at N = 64 (stacked): e.g. 192 tiles of N = 128 take 2 rounds x 800, the same layer
    // as 384 tiles of N = 64 takes 3 x 470.  Pick the cheaper (single-class 3xTF32 layers; measured per-step costs).
    if (wave_model_enabled() && nsplit == 3 && nclass == 1 && n_tile == 128 && p.Cout % 64 == 0) {
      const long t128 = (long)m_tiles * ceil_div(p.Cout, 128), t64 = (long)m_tiles * ceil_div(p.Cout, 64);
      const long c128 = ((t128 + 147) / 148) * 800, c64 = ((t64 + 147) / 148) * 470;
      if (c64 * 100 < c128 * 95) n_tile = 64;
    }
    if (n_tile < 16 || nclass * n_tile + ring_cols > 512) return false;
    prm.mode = (nsplit == 1) ? 0 : 2;
    if (nsplit == 3 && n_tile <= 64 && nclass * 2 * n_tile + ring_cols <= 512) prm.mode = 1;
    // Three-instruction mode instead of the stacked one where the stacked accumulators would cost the double buffering (four
    // classes of N = 32, netRefine/refine0: 0.626 -> 0.585 ms, the epilogue of tile i overlaps tile i + 1 again) and where the
    // epilogue is the bound (8-channel layers: half the TMEM reads and no add per tile, conv1y 0.124 -> 0.119 ms).
    // DEMON_TC_MODE2=0 switches both off (A/B), bit 0 = multi-class, bit 1 = 8-channel.
    {
      static const int mode2 = []() { const char* e = getenv("DEMON_TC_MODE2"); return e ? atoi(e) : 3; }();
      if (prm.mode == 1 && (((mode2 & 1) && nclass == 4 && 2 * nclass * 2 * n_tile + ring_cols > 512 && 2 * nclass * n_tile + ring_cols <= 512) ||
                            ((mode2 & 2) && prm.cin8)))
        prm.mode = 2;
    }
    prm.n_tile = n_tile;
    prm.acc_w = (prm.mode == 1 ? 2 : 1) * n_tile;
    prm.nbuf = (2 * nclass * prm.acc_w + ring_cols <= 512) ? 2 : 1;
    prm.n_tiles = ceil_div(p.Cout, n_tile);
    int cols = 32;
    while (cols < prm.nbuf * nclass * prm.acc_w + ring_cols) cols <<= 1;
    prm.tmem_cols = cols;
    prm.k_chunks = prm.cin8 ? 1 : p.Cin / 32;
    // ---- weights: one class block = [W_hi ; W_lo] (3xTF32) or W alone, 1024-byte aligned; a step holds the blocks of its
    // classes in ascending class order; the ring slot is as large as the widest step
    prm.cls_bytes = (nsplit == 3) ? n_tile * 256 : (n_tile * 128 + 1023) / 1024 * 1024;
    int woff = 0, wmax = 0;
    for (int t = 0; t < prm.nsteps; ++t) {
      prm.st_woff[t] = woff;
      prm.st_wbytes[t] = popcount4(prm.st_cmask[t]) * prm.cls_bytes;
      woff += prm.st_wbytes[t];
      wmax = std::max(wmax, prm.st_wbytes[t]);
    }
    prm.w_chunk_bytes = woff;
    prm.w_stage_bytes = wmax;
    // shared memory: the weights (a ring of kRing slots, or the whole layer: see the W producer), then up to 4 halo stages
    // (at least 2).  Resident: one n tile, one class, every step the same size, <= 96 KB in all and at least 2 A stages left.
    const int w_total = prm.k_chunks * prm.w_chunk_bytes;
    prm.w_resident = 0;
    prm.w_region_bytes = kRing * wmax;
    if (w_resident_enabled() && nclass == 1 && prm.n_tiles == 1 && w_total <= 96 * 1024 && wmax * prm.nsteps == prm.w_chunk_bytes &&
        budget - w_total >= 2 * prm.a_region_bytes) {
      prm.w_resident = 1;
      prm.w_region_bytes = w_total;
    }
    const int rest = budget - prm.w_region_bytes;
    if (rest < 2 * prm.a_region_bytes) continue;   // next candidate: narrower steps / narrower N tile
    prm.sa = std::min(kMaxAStages, rest / prm.a_region_bytes);
    plan.smem_bytes = prm.sa * prm.a_region_bytes + prm.w_region_bytes + 4 * kEpiStageBytes + 1024;
    found = true;
    break;
  }
  if (!found || prm.sa < 2) return false;
  prm.B = p.B;
  const int m_tiles = prm.tiles_x * prm.tiles_y * prm.tiles_b;
  prm.total_tiles = m_tiles * prm.n_tiles;
  // ---- split-K: layers with few tiles and a long K loop (the 6x8 / 12x16 levels; everything at batch 1) leave most SMs
  // idle and run their K loop serially.  Cost model in cycles: rounds of 148 items x (steps of an item x ~600 + ~3000 of
  // prologue / epilogue), + ~12000 for the second pass + the partial sums' trip through L2; a split has to win 10 %.
  prm.ksplit = 1;
  if (splitk_enabled() && !prm.cin8 && !prm.w_resident && prm.k_chunks >= 4) {
    // the partial sums travel to the second pass through L2 (~4000 B per cycle for write + read back): measured on
    // netFlow2/refine3/upconv at batch 64 (25 MB of output), three slices win 8 %, not the 24 % the rounds alone promise
    const long out_bytes = (long)p.B * p.Hfull * p.Wfull * ((p.Cout + 3) / 4 * 4) * 4;
    auto cost = [&](int ks) {
      const long items = (long)prm.total_tiles * ks;
      return ((items + 147) / 148) * ((long)(prm.k_chunks / ks) * prm.nsteps * 600 + 3000) + (ks > 1 ? 12000 + ks * out_bytes * 2 / 4000 : 0);
    };
    long best = cost(1);
    for (int ks = 2; ks <= 8; ++ks) {
      if (prm.k_chunks % ks != 0 || prm.k_chunks / ks < 2) continue;
      const long c = cost(ks);
      if (c * 10 < best * 9 && c < cost(prm.ksplit)) prm.ksplit = ks;
    }
  }
  prm.kc_split = prm.k_chunks / prm.ksplit;
  prm.total_items = prm.total_tiles * prm.ksplit;
  prm.part_stride = 0;
  auto fd = [](int d) { return d <= 1 ? 0u : (uint32_t)((1ull << 32) / (uint64_t)d + 1ull); };
  prm.mul_n_tiles = fd(prm.n_tiles); prm.mul_tiles_x = fd(prm.tiles_x); prm.mul_tiles_y = fd(prm.tiles_y);
  prm.mul_ksplit = fd(prm.ksplit);
  if ((uint64_t)prm.total_items * (uint64_t)prm.ksplit >= (1ull << 32)) return false;
  if ((uint64_t)prm.total_tiles * (uint64_t)std::max(prm.n_tiles, std::max(prm.tiles_x, prm.tiles_y)) >= (1ull << 32)) return false;
  prm.out = p.out; prm.out_pitch = p.out_pitch; prm.Ho = p.Ho; prm.Wo = p.Wo; prm.Hfull = p.Hfull; prm.Wfull = p.Wfull;
  prm.osy = p.osy; prm.osx = p.osx; prm.Cout = p.Cout; prm.bias = p.bias; prm.leaky = p.leaky;
  for (int c = 0; c < nclass; ++c) { prm.cls_ooy[c] = probs[c].ooy; prm.cls_oox[c] = probs[c].oox; }
  if (!encode) return true;
  // TMA descriptors: 5-D view {sx*C, W/sx, sy, H/sy, B} of the input slice, one box shape per plane
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return false;
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  const cuuint64_t cp = (cuuint64_t)p.in_pitch;
  cuuint64_t gdim[5] = {(cuuint64_t)(p.sx - 1) * cp + (cuuint64_t)p.Cin, (cuuint64_t)(p.Wi / p.sx), (cuuint64_t)p.sy,
                        (cuuint64_t)(p.Hi / p.sy), (cuuint64_t)p.B};
  cuuint64_t gstr[4] = {(cuuint64_t)p.sx * cp * 4, (cuuint64_t)p.Wi * cp * 4, (cuuint64_t)p.sy * p.Wi * cp * 4,
                        (cuuint64_t)p.Hi * p.Wi * cp * 4};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < prm.nplanes; ++i) {
    cuuint32_t box[5] = {(cuuint32_t)(prm.cin8 ? 8 : 32), (cuuint32_t)prm.planes[i].cols, 1, (cuuint32_t)prm.planes[i].rows,
                         (cuuint32_t)(prm.per_tap ? prm.tb : 1)};
    CUresult r = enc(&plan.maps.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(p.in), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, prm.cin8 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
  }
  for (int i = prm.nplanes; i < kMaxPlanes; ++i) plan.maps.m[i] = plan.maps.m[0];
  return true;
}

bool tc_halo_supported(const ConvProblem* probs, int nclass) {
  if (nclass < 1 || nclass > 4) return false;
  for (int c = 0; c < nclass; ++c) {
    ConvProblem q = probs[c];
    if (q.Cin == 8 && q.in_pitch == 8) q.Cin = 32;   // 8-channel mode: everything but the channel rule must hold
    if (!tc_layer_supported(q)) return false;
    if (probs[c].in != probs[0].in || probs[c].Cout != probs[0].Cout || probs[c].sy != probs[0].sy || probs[c].sx != probs[0].sx) return false;
  }
  HaloPlan plan;
  return halo_build(probs, nclass, 3, plan, false);
}

// debug: the plan halo_build chooses for a layer, as text (no device needed)
int tc_halo_describe(const ConvProblem* probs, int nclass, int nsplit, char* buf, int buflen) {
  HaloPlan plan;
  if (!halo_build(probs, nclass, nsplit, plan, false)) return snprintf(buf, buflen, "halo: unsupported");
  const HaloParams& q = plan.prm;
  int n = snprintf(buf, buflen, "halo %s mode %d n_tile %d x%d nbuf %d tmem %d steps %d x %d chunks sa %d a_stage %d w_slot %d smem %d tiles %d groups %d ksplit %d wres %d |",
                   q.per_tap ? "per-tap" : (q.cin8 ? "cin8" : "halo"), q.mode, q.n_tile, q.n_tiles, q.nbuf, q.tmem_cols, q.nsteps, q.k_chunks, q.sa,
                   q.a_region_bytes, q.w_stage_bytes, plan.smem_bytes, q.total_tiles, q.ngroups, q.ksplit, q.w_resident);
  for (int t = 0; t < q.nsteps && n < buflen - 32; ++t)
    n += snprintf(buf + n, buflen - n, " [c%x f%x w%d+%d]", q.st_cmask[t], q.st_fmask[t], q.st_woff[t], q.st_wbytes[t]);
  return n;
}

static float tf32_round_h(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return x;
  u += 0x00000FFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

int tc_halo_prepare(TcLayer& t, const ConvProblem* probs, const float* const* w_hosts, int nclass, int precision) {
  const ConvProblem& p = probs[0];
  const int nsplit = (precision == DEMON_PREC_TF32) ? 1 : 3;
  HaloPlan* plan = new HaloPlan();
  if (!halo_build(probs, nclass, nsplit, *plan, true)) {
    delete plan;
    return fail(DEMON_E_CUDA, "tc_halo_prepare: could not build the plan (tensor map encode failed?)");
  }
  const HaloParams& prm = plan->prm;
  // weights: [n tile][chunk][step][class of the step] blocks of [W_hi | W_lo], n_tile rows x 32 fp32, K-major, pre-swizzled
  const size_t total = (size_t)prm.n_tiles * prm.k_chunks * prm.w_chunk_bytes;
  std::vector<unsigned char> packed(total, 0);
  for (int nt = 0; nt < prm.n_tiles; ++nt)
    for (int kc = 0; kc < prm.k_chunks; ++kc)
      for (int tt = 0; tt < prm.nsteps; ++tt) {
        int idx = 0;
        for (int cls = 0; cls < 4; ++cls) {
          if (!((prm.st_cmask[tt] >> cls) & 1)) continue;
          unsigned char* blk = packed.data() + ((size_t)nt * prm.k_chunks + kc) * prm.w_chunk_bytes + prm.st_woff[tt] + (size_t)idx * prm.cls_bytes;
          ++idx;
          const int tap = prm.cin8 ? 0 : plan->st_tap[tt][cls];
          for (int r = 0; r < prm.n_tile; ++r) {
            const int co = nt * prm.n_tile + r;
            for (int k = 0; k < 32; ++k) {
              float w = 0.f;
              if (prm.cin8) {   // K index = (tap within the group of four, channel)
                const int rt = 4 * tt + k / 8;
                if (co < p.Cout && rt < prm.ntaps_real) w = w_hosts[0][((size_t)plan->st_tap[rt][0] * 8 + (k & 7)) * p.Cout_pad + co];
              } else if (co < p.Cout) {
                w = w_hosts[cls][((size_t)tap * p.Cin + kc * 32 + k) * p.Cout_pad + co];
              }
              const float hi = (nsplit == 3) ? tf32_round_h(w) : w;
              const float lo = w - hi;
              const size_t off = (size_t)r * 128 + (size_t)(((k >> 2) ^ (r & 7)) << 4) + (size_t)(k & 3) * 4;
              memcpy(blk + off, &hi, 4);
              if (nsplit == 3) memcpy(blk + (size_t)prm.n_tile * 128 + off, &lo, 4);
            }
          }
        }
      }
  void* dw = nullptr;
  cudaError_t e = cudaMalloc(&dw, total);
  if (e == cudaSuccess) e = cudaMemcpy(dw, packed.data(), total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { delete plan; return fail(DEMON_E_CUDA, "tc_halo_prepare: %s", cudaGetErrorString(e)); }
  plan->prm.w = static_cast<const unsigned char*>(dw);
  t.w_packed = dw;
  t.halo_plan = plan;
  t.per_tap = prm.per_tap;
  t.nclass = nclass;
  t.n_tile = prm.n_tile; t.n_tiles = prm.n_tiles; t.k_chunks = prm.k_chunks; t.nsplit = nsplit;
  t.th = kTileH; t.tw = kTileW; t.tb = 1; t.stages = prm.sa; t.smem_bytes = plan->smem_bytes;
  t.ksplit = prm.ksplit;
  t.splitk_bytes = (prm.ksplit > 1) ? (size_t)prm.ksplit * p.B * p.Hfull * p.Wfull * ((p.Cout + 3) / 4 * 4) * sizeof(float) : 0;
  return DEMON_OK;
}

void tc_halo_free(TcLayer& t) {
  if (t.halo_plan) delete static_cast<HaloPlan*>(t.halo_plan);
  t.halo_plan = nullptr;
}

static long long* g_timing_dev = nullptr;
void tc_halo_enable_timing(bool on) {
  if (on && !g_timing_dev) { cudaMalloc(&g_timing_dev, 256 * 16 * sizeof(long long)); }
  if (g_timing_dev) cudaMemset(g_timing_dev, 0, 256 * 16 * sizeof(long long));
  if (!on && g_timing_dev) { cudaFree(g_timing_dev); g_timing_dev = nullptr; }
}
int tc_halo_read_timing(long long* host, int nblocks) {
  if (!g_timing_dev) return -1;
  cudaMemcpy(host, g_timing_dev, (size_t)nblocks * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
  return 0;
}

template <bool PER_TAP, bool CIN8, int MODE>
static int launch_variant(const HaloPlan* plan, const HaloParams& prm, int grid, int threads, cudaStream_t stream) {
  TcDeviceState& ds = tc_device_state();
  const int slot = (PER_TAP ? 1 : 0) + (CIN8 ? 2 : 0) + 3 * MODE;
  if (!(ds.halo_attr_set & (1u << slot))) {
    DEMON_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_halo_kernel<PER_TAP, CIN8, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    ds.halo_attr_set |= 1u << slot;
  }
  cudaError_t le = launch_pdl(conv_tc_halo_kernel<PER_TAP, CIN8, MODE>, dim3(grid), dim3(threads), (size_t)plan->smem_bytes, stream, plan->maps, prm);
  if (le != cudaSuccess) return fail(DEMON_E_CUDA, "conv_tc_halo launch: %s", cudaGetErrorString(le));
  DEMON_LAUNCH_CHECK();
  return DEMON_OK;
}

int conv_tc_halo_launch(const TcLayer& t, const ConvProblem* probs, cudaStream_t stream) {
  HaloPlan* plan = static_cast<HaloPlan*>(t.halo_plan);
  HaloParams prm = plan->prm;
  prm.out = probs[0].out;          // the output slice may be re-pointed between calls (caller-owned result buffers)
  prm.out_pitch = probs[0].out_pitch;
  TcDeviceState& ds = tc_device_state();
  if (!ds.err_dev) return fail(DEMON_E_CUDA, "tcgen05 path: no error flag on device %d", ds.device);
  prm.err = ds.err_dev;
  prm.timing = g_timing_dev;
  const int cpitch = (prm.Cout + 3) / 4 * 4;
  const long long npix = (long long)prm.B * prm.Hfull * prm.Wfull;
  float* const final_out = prm.out;
  const int final_pitch = prm.out_pitch;
  if (prm.ksplit > 1) {   // partial sums into the caller's scratch, bias and activation in the second pass
    if (!probs[0].partial) return fail(DEMON_E_STATE, "conv_tc_halo: split-K layer launched without a scratch buffer");
    prm.out = probs[0].partial;
    prm.out_pitch = cpitch;
    prm.part_stride = npix * cpitch;
    prm.bias = nullptr;
    prm.leaky = 0;
  }
  const int grid = std::min(prm.total_items, ds.sms);
  const int threads = 32 * (2 + 4 * prm.ngroups + 4 + 1);
  int rc;
  if (prm.per_tap) {
    if (prm.mode == 0) rc = launch_variant<true, false, 0>(plan, prm, grid, threads, stream);
    else if (prm.mode == 1) rc = launch_variant<true, false, 1>(plan, prm, grid, threads, stream);
    else rc = launch_variant<true, false, 2>(plan, prm, grid, threads, stream);
  } else if (prm.cin8) {
    if (prm.mode == 0) rc = launch_variant<false, true, 0>(plan, prm, grid, threads, stream);
    else if (prm.mode == 1) rc = launch_variant<false, true, 1>(plan, prm, grid, threads, stream);
    else rc = launch_variant<false, true, 2>(plan, prm, grid, threads, stream);
  } else {
    if (prm.mode == 0) rc = launch_variant<false, false, 0>(plan, prm, grid, threads, stream);
    else if (prm.mode == 1) rc = launch_variant<false, false, 1>(plan, prm, grid, threads, stream);
    else rc = launch_variant<false, false, 2>(plan, prm, grid, threads, stream);
  }
  if (rc != DEMON_OK || prm.ksplit == 1) return rc;
  const long long total = npix * (cpitch / 4);
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)ds.sms * 8);
  cudaError_t le = launch_pdl(halo_splitk_reduce_kernel, dim3(blocks), dim3(256), 0, stream, (const float*)probs[0].partial, prm.part_stride,
                              prm.ksplit, final_out, final_pitch, cpitch, prm.Cout, npix, plan->prm.bias, plan->prm.leaky);
  if (le != cudaSuccess) return fail(DEMON_E_CUDA, "halo_splitk_reduce launch: %s", cudaGetErrorString(le));
  DEMON_LAUNCH_CHECK();
  return DEMON_OK;
}

}  // namespace demon
