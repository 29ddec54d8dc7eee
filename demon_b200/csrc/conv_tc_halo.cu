// tcgen05 implicit-GEMM convolution, "halo" variant: every input pixel is fetched ONCE per output tile.
//
// conv_tc.cu fetches one shifted 128-pixel A tile per filter tap, i.e. it reads the input kh*kw times (16 times for the
// four sub-pixel classes of a transposed convolution) and splits it into hi/lo as often; at the resolutions of the
// first and last layers that makes the kernel L2-bandwidth bound.  Here the output tile is 16 rows x 8 columns and,
// per 32-channel chunk, ONE TMA box per stride-parity plane brings the tile plus its halo into shared memory
// ((16+nqy-1) x (8+nqx-1) pixels, 128 bytes per pixel, 128-byte swizzle).  Every tap is then just a different START
// ADDRESS inside the same shared-memory image:
//
//     GEMM row m = (y, x) of the tile  ->  halo pixel (y + qy - qy_min, x + qx - qx_min)
//
// Warp roles (15 warps, persistent over tiles, mbarrier rings between them, every wait bounded):
//     warp 0        A producer: the halo boxes (TMA)
//     warp 14       W producer: one pre-swizzled [W_hi ; W_lo] block per (chunk, tap) step (cp.async.bulk); its bytes complete
//                   on the step's barrier
//     warps 2..9    stagers, two groups of four alternating steps: shifted pixel rows of the halo image -> registers ->
//                   A_hi (raw fp32) and A_lo = A - trunc_tf32(A) in a TMEM ring slot (tcgen05.st)
//     warp 1        MMA thread: TS-mode tcgen05.mma kind::tf32 (A from TMEM, B = weights from shared memory), 3xTF32 as two
//                   stacked instructions for N <= 64, one poll and two commits per step
//     warps 10..13  epilogue: tcgen05.ld -> transpose through shared memory -> bias, leaky ReLU -> coalesced NHWC stores
//
// (The first version of this kernel fed the tap-shifted halo rows to the tensor core directly from shared memory -- a
// descriptor start shifted by whole 128-byte rows and a stride that is not a multiple of 1024 read exactly the expected
// rows, because the 128-byte swizzle is applied on ABSOLUTE shared-memory address bits, tools/umma_shift_probe.cu --
// but a shared-memory A operand costs ~43 cycles per instruction on top of N/2, tools/umma_rate_probe.cu.)
//
// All sub-pixel classes of a transposed convolution accumulate side by side in TMEM (nclass x N columns, double
// buffered when they fit), so the input halo is shared by all 16 taps.  Variants: PER_TAP (images that are not made of
// whole 16 x 8 tiles: one 128-pixel box per step instead of a halo) and CIN8 (8-channel inputs: four taps per K = 32 step).
#include <cuda.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "conv_tc.cuh"
#include "conv_tc_ptx.cuh"

namespace demon {

namespace {

constexpr int kStagerWarps = 8;
constexpr int kWProducerWarp = 2 + kStagerWarps + 4;
constexpr int kThreads = 32 * (kWProducerWarp + 1);   // 15 warps: A producer, MMA, 8 stagers, 4 epilogue, W producer
constexpr int kMaxAStages = 4;
constexpr int kMaxTStages = 8;
constexpr int kTileH = 16, kTileW = 8;
constexpr int kMaxPlanes = 4;
constexpr int kMaxWStages = 8;
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

struct HaloTap {
  int plane;
  int a_off;      // byte offset of the tap's first pixel inside its plane
  int cls;
  int first;      // 1: first tap of its class (accumulator starts from zero at chunk 0)
};

struct HaloParams {
  int tiles_x, tiles_y, B, n_tiles, total_tiles;
  uint32_t mul_n_tiles, mul_tiles_x, mul_tiles_y;   // fast_div multipliers
  int k_chunks, nplanes, ntaps, nclass;
  // per-tap mode (layers whose images are not made of whole 16x8 tiles): the A stage is ONE tap's 128-pixel tile
  // (tb images x th rows x tw columns, one TMA box per (chunk, tap) step, box coordinates carry the tap like in conv_tc.cu)
  int per_tap, tw, th, tb, tiles_b, aempty_count;
  int tap_c[kMaxTaps], tap_qx[kMaxTaps], tap_ry[kMaxTaps], tap_qy[kMaxTaps];
  // 8-channel mode (Cin == 8, e.g. the image pair): a pixel is 32 bytes in shared memory (no swizzle), the halo is loaded
  // once per tile, and one K = 32 step gathers FOUR taps x 8 channels into the TMEM A operand (`ntaps` then counts these
  // groups, g_plane / g_aoff describe the real taps, -1 = missing tap -> zero columns)
  int cin8, ntaps_real;
  int g_plane[kMaxTaps], g_aoff[kMaxTaps];
  HaloPlane planes[kMaxPlanes];
  HaloTap taps[kMaxTaps];
  int a_region_bytes;   // hi image of all planes (lo image follows at the same offsets)
  int sa;               // A (halo, shared memory) stages
  int st;               // A-operand (TMEM) ring slots of 64 columns
  int w_stage_bytes, sw;
  int n_tile, nsplit, tmem_cols;
  int nbuf;             // accumulator buffers (2 = epilogue overlaps the next tile; 1 when TMEM is short)
  int stacked;          // 1: 3xTF32 as two instructions on [W_hi ; W_lo] (N <= 64), 0: three instructions
  int acc_w;            // accumulator columns per class: 2 * n_tile in 3xTF32 mode ([big | small] terms), n_tile otherwise
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

__device__ __forceinline__ void wait_t(uint32_t bar, uint32_t parity, int* err, long long& acc, bool timed) {
  if (!timed) { mbar_wait(bar, parity, err); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity, err);
  acc += clock64() - t0;
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

template <bool PER_TAP, bool CIN8>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_halo_kernel(const __grid_constant__ HaloMaps maps, const HaloParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // bars: A_full[4], A_empty[4], T_full[8], T_empty[8], W_full[8], W_empty[8], accum_full[2], accum_empty[2]
  __shared__ __align__(8) uint64_t bars[2 * kMaxAStages + 2 * kMaxTStages + 2 * kMaxWStages + 4];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t afull0 = smem_u32(&bars[0]), aempty0 = smem_u32(&bars[kMaxAStages]);
  const uint32_t tfull0 = smem_u32(&bars[2 * kMaxAStages]), tempty0 = smem_u32(&bars[2 * kMaxAStages + kMaxTStages]);
  const uint32_t wfull0 = smem_u32(&bars[2 * kMaxAStages + 2 * kMaxTStages]);
  const uint32_t wempty0 = smem_u32(&bars[2 * kMaxAStages + 2 * kMaxTStages + kMaxWStages]);
  const uint32_t cfull0 = smem_u32(&bars[2 * kMaxAStages + 2 * kMaxTStages + 2 * kMaxWStages]);
  const uint32_t cempty0 = smem_u32(&bars[2 * kMaxAStages + 2 * kMaxTStages + 2 * kMaxWStages + 2]);
  const int a_stage_bytes = p.a_region_bytes;
  unsigned char* w_ring = smem + (size_t)p.sa * a_stage_bytes;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nplanes; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[i]) : "memory");
    for (int s = 0; s < kMaxAStages; ++s) {
      mbar_init(afull0 + 8 * s, 1);
      mbar_init(aempty0 + 8 * s, (uint32_t)p.aempty_count);
    }
    for (int s = 0; s < kMaxTStages; ++s) {
      mbar_init(tfull0 + 8 * s, 4 + 1);   // the four warps of the owning stager group + the W producer's expect_tx
      mbar_init(tempty0 + 8 * s, 1);
    }
    for (int s = 0; s < kMaxWStages; ++s) {
      mbar_init(wfull0 + 8 * s, 1);
      mbar_init(wempty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(cfull0 + 8 * s, 1);
      mbar_init(cempty0 + 8 * s, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int acc_cols = p.nclass * p.acc_w;           // columns of one accumulator buffer
  const uint32_t t_ring = tmem_base + (uint32_t)(p.nbuf * acc_cols);   // A-operand ring: st slots of 64 columns (hi | lo)
  const bool timed = p.timing != nullptr;
  const int steps_per_tile = p.k_chunks * p.ntaps;

  if (warp == 0) {
    // ===== A producer: one halo box per plane and 32-channel chunk ======================================================
    if (elect_one_sync()) {
      int sa = 0;
      uint32_t pa = 0;
      uint32_t a_bytes = 0;
      long long w_aempty = 0;
      const long long t_begin = clock64();
      for (int i = 0; i < p.nplanes; ++i) a_bytes += (uint32_t)p.planes[i].bytes;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int nt, n, y0, x0;
        halo_decode_tile<PER_TAP>(p, tile, nt, n, y0, x0);
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          if (PER_TAP) {
            for (int t = 0; t < p.ntaps; ++t) {
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
  } else if (warp == kWProducerWarp) {
    // ===== W producer: one weight block per (chunk, tap), its own ring ====================================================
    // The copy's bytes complete on the STEP's barrier (the one the stager group arrives on), so the MMA thread polls one
    // barrier per step instead of two (a completed poll costs it ~150 cycles, tools/umma_contention_probe.cu).  The weight
    // ring is at most as deep as the A-operand ring (sw <= st): once W_empty of step s - sw has been committed, step
    // s - st has long been consumed, i.e. the step barrier is in the phase this expect_tx belongs to.
    if (elect_one_sync()) {
      int sw = 0, ts = 0;
      uint32_t pw = 0;
      long long w_wempty = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile - fast_div(tile, p.n_tiles, p.mul_n_tiles) * p.n_tiles;
        const unsigned char* wsrc = p.w + (size_t)nt * steps_per_tile * p.w_stage_bytes;
        for (int kt = 0; kt < steps_per_tile; ++kt) {
          wait_t(wempty0 + 8 * sw, pw ^ 1, p.err, w_wempty, timed);
          mbar_expect_tx(tfull0 + 8 * ts, (uint32_t)p.w_stage_bytes);
          bulk_load(smem_u32(w_ring + (size_t)sw * p.w_stage_bytes), wsrc + (size_t)kt * p.w_stage_bytes, (uint32_t)p.w_stage_bytes,
                    tfull0 + 8 * ts);
          if (++sw == p.sw) { sw = 0; pw ^= 1; }
          if (++ts == p.st) ts = 0;
        }
      }
      if (timed) p.timing[blockIdx.x * 16 + 1] = w_wempty;
    }
  } else if (warp == 1) {
    // ===== MMA issuer: A operand from TMEM (staged by the stager warps), B operand (weights) from shared memory =========
    // This thread is the pacemaker of the whole CTA, so its per-step overhead is kept minimal: the barrier polls of step
    // s+1 are issued BEFORE the MMAs of step s and only looked at afterwards (an mbarrier try_wait takes ~90 cycles even
    // when the phase is already complete), and the per-tap class / first-of-class flags live in two register bitmasks.
    if (elect_one_sync()) {
      int st = 0, sw = 0;
      uint32_t pt = 0;
      // 3xTF32 as TWO instructions per K8 step when N <= 64 ("stacked"): the weight block holds [W_hi ; W_lo] as 2N
      // consecutive rows, so
      //   D[:, 0:2N]  (+)= A_hi * [W_hi ; W_lo]      (one UMMA of N' = 2N: big term | first small term)
      //   D[:, N:2N]   += A_lo * W_hi                 (second small term)
      // and the epilogue adds the two halves: fewer, wider instructions (the tensor core has a fixed cost per
      // instruction, tools/umma_rate_probe.cu), and the small terms accumulate apart from the big one.
      const uint32_t idesc = umma_idesc_tf32(p.n_tile), idesc2 = umma_idesc_tf32(2 * p.n_tile);
      uint32_t cls_bits = 0, first_bits = 0;
      for (int t = 0; t < p.ntaps; ++t) { cls_bits |= (uint32_t)p.taps[t].cls << (2 * t); first_bits |= (uint32_t)p.taps[t].first << t; }
      const uint32_t w_lo_off = (uint32_t)p.n_tile * 128u;
      const uint32_t w_ring_addr = smem_u32(w_ring);
      long long w_cempty = 0, w_tfull = 0, w_poll = 0, w_issue = 0, w_commit = 0;
      const long long t_begin = clock64();
      int it = 0;
      bool rdy_t = mbar_try(tfull0, 0);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int a = (p.nbuf == 2) ? (it & 1) : 0;
        const uint32_t cphase = (p.nbuf == 2) ? ((it >> 1) & 1) : (it & 1);
        wait_t(cempty0 + 8 * a, cphase ^ 1, p.err, w_cempty, timed);
        const uint32_t d_base = tmem_base + (uint32_t)(a * acc_cols);
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          for (int t = 0; t < p.ntaps; ++t) {
            const long long tq0 = timed ? clock64() : 0;
            if (!rdy_t) wait_t(tfull0 + 8 * st, pt, p.err, w_tfull, timed);
            tc_fence_after();
            const uint32_t a_hi = t_ring + (uint32_t)(st * 64), a_lo = a_hi + 32;
            const uint32_t wbase = w_ring_addr + (uint32_t)(sw * p.w_stage_bytes);
            const int st_cur = st, sw_cur = sw;
            if (++st == p.st) { st = 0; pt ^= 1; }
            if (++sw == p.sw) sw = 0;
            rdy_t = mbar_try(tfull0 + 8 * st, pt);   // poll the NEXT step's barrier
            const long long tq1 = timed ? clock64() : 0;
            const uint64_t w_hi = umma_desc_sw128_sbo(wbase, 1024);
            const uint32_t d_tmem = d_base + ((cls_bits >> (2 * t)) & 3u) * (uint32_t)p.acc_w;
            const bool fresh = (kc == 0) && ((first_bits >> t) & 1u);
            if (p.stacked) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {   // 4 x K8: +8 TMEM columns / +32 bytes inside the swizzled weight row
                umma_tf32_ts(d_tmem, a_hi + 8 * j, w_hi + (uint64_t)(2 * j), idesc2, !(fresh && j == 0));
                umma_tf32_ts(d_tmem + (uint32_t)p.n_tile, a_lo + 8 * j, w_hi + (uint64_t)(2 * j), idesc, 1);
              }
            } else if (p.nsplit == 3) {
              const uint64_t w_lo = umma_desc_sw128_sbo(wbase + w_lo_off, 1024);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                umma_tf32_ts(d_tmem, a_hi + 8 * j, w_lo + (uint64_t)(2 * j), idesc, !(fresh && j == 0));
                umma_tf32_ts(d_tmem, a_lo + 8 * j, w_hi + (uint64_t)(2 * j), idesc, 1);
                umma_tf32_ts(d_tmem, a_hi + 8 * j, w_hi + (uint64_t)(2 * j), idesc, 1);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_tf32_ts(d_tmem, a_hi + 8 * j, w_hi + (uint64_t)(2 * j), idesc, !(fresh && j == 0));
            }
            const long long tq2 = timed ? clock64() : 0;
            umma_commit(tempty0 + 8 * st_cur);
            umma_commit(wempty0 + 8 * sw_cur);
            if (kc == p.k_chunks - 1 && t == p.ntaps - 1) umma_commit(cfull0 + 8 * a);
            if (timed) { const long long tq3 = clock64(); w_poll += tq1 - tq0; w_issue += tq2 - tq1; w_commit += tq3 - tq2; }
          }
        }
      }
      if (timed) {
        long long* tm = p.timing + blockIdx.x * 16;
        tm[2] = w_cempty; tm[3] = w_tfull; tm[5] = 0; tm[13] = w_poll; tm[14] = w_issue; tm[15] = w_commit; tm[9] = clock64() - t_begin;
      }
    }
  } else if (warp < 2 + kStagerWarps) {
    // ===== stagers: halo image (shared memory) -> A operand of one tap in TMEM =========================================
    // Thread (quadrant q, lane l) owns GEMM row m = 32q + l = tile pixel (m / 8, m % 8).  For every (chunk, tap) step it
    // reads its (shifted) 128-byte pixel out of the swizzled halo image -- conflict free: the 8 lanes of a quarter warp
    // hit 8 different 16-byte columns -- and writes the raw fp32 values (A_hi: the tensor core ignores the low 13
    // mantissa bits) and A_lo = A - trunc_tf32(A) to its TMEM lane.  Two groups of four warps alternate steps.
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int g = m >> 3, r = m & 7;
    int sa = 0;
    uint32_t pa = 0;
    long long w_safull = 0, w_tempty = 0;
    const long long t_begin = clock64();
    int step = 0;    // global (chunk, tap) step counter of this CTA
    int slot = 0;    // = step % st, kept incrementally (no division in the loop)
    uint32_t use = 0;   // parity of step / st
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        uint32_t abase = 0;
        if (!PER_TAP) {
          wait_t(afull0 + 8 * sa, pa, p.err, w_safull, timed);
          __syncwarp();
          abase = smem_u32(smem + (size_t)sa * a_stage_bytes);
        }
        for (int t = 0; t < p.ntaps; ++t, ++step) {
          const int slot_cur = slot;
          const uint32_t use_cur = use;
          if (++slot == p.st) { slot = 0; use ^= 1; }
          const int sa_cur = sa;
          const uint32_t pa_cur = pa;
          if (PER_TAP) { if (++sa == p.sa) { sa = 0; pa ^= 1; } }   // one A stage per step
          if ((step & 1) != grp) continue;
          uint32_t row;
          if (PER_TAP) {
            wait_t(afull0 + 8 * sa_cur, pa_cur, p.err, w_safull, timed);
            __syncwarp();
            row = smem_u32(smem + (size_t)sa_cur * a_stage_bytes) + (uint32_t)(g * 1024 + r * 128);
          } else if (!CIN8) {
            const HaloTap& tp = p.taps[t];
            const HaloPlane& pl = p.planes[tp.plane];
            row = abase + (uint32_t)(pl.smem_off + tp.a_off + g * pl.cols * 128 + r * 128);
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
                const uint32_t px = abase + (uint32_t)(pl.smem_off + p.g_aoff[rt] + (g * pl.cols + r) * 32);
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
          wait_t(tempty0 + 8 * slot_cur, use_cur ^ 1, p.err, w_tempty, timed);
          __syncwarp();
          tc_fence_after();
          const uint32_t taddr = t_ring + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot_cur * 64);
          tmem_st_x32(taddr, hi);
          if (p.nsplit == 3) {
#pragma unroll
            for (int c = 0; c < 32; ++c) lo[c] = __float_as_uint(tf32_lo(__uint_as_float(hi[c])));
            tmem_st_x32(taddr + 32, lo);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(tfull0 + 8 * slot_cur);
            if (PER_TAP) mbar_arrive(aempty0 + 8 * sa_cur);   // the tap's tile has been consumed: stage back to the producer
          }
        }
        if (!PER_TAP) {
          __syncwarp();
          if (lane == 0) mbar_arrive(aempty0 + 8 * sa);   // this warp is done reading the halo stage
          if (++sa == p.sa) { sa = 0; pa ^= 1; }
        }
      }
    }
    if (timed && (threadIdx.x == 64 || threadIdx.x == 64 + 128)) {
      if (grp == 0) { p.timing[blockIdx.x * 16 + 6] = w_safull; p.timing[blockIdx.x * 16 + 4] = w_tempty; p.timing[blockIdx.x * 16 + 10] = clock64() - t_begin; }
      else p.timing[blockIdx.x * 16 + 12] = w_tempty;
    }
  } else if (warp < 2 + kStagerWarps + 4) {
    // ===== epilogue ======================================================================================================
    // After the TMEM load a thread holds one output pixel (GEMM row) x 32 channels, and consecutive pixels are a whole
    // pixel pitch apart in global memory: storing rows directly costs 32 memory wavefronts per store instruction, and
    // the wait counters showed the epilogue (not the MMA thread) bounding the layers with few steps per tile (conv1y,
    // conv2y: ~3400 / ~9000 cycles per tile).  Each warp therefore transposes its 32 x 32 block through shared memory
    // (row pitch 144 B: conflict free both ways) and stores with lane = (row % 4, 16-byte chunk), i.e. four pixels x 128
    // contiguous bytes per instruction; the bias is then one float4 per lane and chunk.
    const int q = warp & 3;
    const int sub = lane >> 3, chunk = lane & 7;
    const uint32_t stg = smem_u32(w_ring + (size_t)p.sw * p.w_stage_bytes) + (uint32_t)(q * kEpiStageBytes);
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
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
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
      float* tile_out = p.out + ((size_t)(n * p.Hfull + y0 * p.osy) * p.Wfull + x0 * p.osx) * p.out_pitch;
      const int cbase = nt * p.n_tile;
      for (int cls = 0; cls < p.nclass; ++cls) {
        float* cls_out = tile_out + (size_t)(p.cls_ooy[cls] * p.Wfull + p.cls_oox[cls]) * p.out_pitch;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * acc_cols + cls * p.acc_w);
        for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
          uint32_t v[32];
          const int ncol = (p.n_tile - c0) >= 32 ? 32 : 16;
          if (ncol == 32) tmem_ld_x32(t_row + c0, v); else tmem_ld_x16(t_row + c0, v);
          if (p.stacked) {   // big term + small terms
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
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
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

int floor_div_h(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------------------
struct HaloPlan {
  HaloParams prm;
  HaloMaps maps;
  int smem_bytes;
};

static int pow2_ceil_h(int v) { int r = 1; while (r < v) r <<= 1; return r; }

static bool halo_build(const ConvProblem* probs, int nclass, int n_tile_req, int nsplit, HaloPlan& plan, bool encode) {
  const ConvProblem& p = probs[0];
  HaloParams& prm = plan.prm;
  memset(&prm, 0, sizeof(prm));
  prm.nclass = nclass;
  prm.nsplit = nsplit;
  prm.cin8 = (p.Cin == 8) ? 1 : 0;
  const int px_bytes = prm.cin8 ? 32 : 128;   // bytes of one pixel of a halo plane in shared memory
  prm.per_tap = ((p.Ho % kTileH) != 0 || (p.Wo % kTileW) != 0) ? 1 : 0;
  if (prm.cin8 && (prm.per_tap || nclass != 1)) return false;
  int m_tiles = 0;
  if (prm.per_tap) {
    // one 128-pixel tile per (chunk, tap) step, tb images x th rows x tw columns (same tiling rule as conv_tc.cu)
    int ntaps = 0;
    for (int c = 0; c < nclass; ++c)
      for (int i = 0; i < probs[c].ntaps; ++i, ++ntaps) {
        if (ntaps >= kMaxTaps) return false;
        const int qy = floor_div_h(probs[c].dy[i], p.sy), qx = floor_div_h(probs[c].dx[i], p.sx);
        prm.tap_qy[ntaps] = qy; prm.tap_ry[ntaps] = probs[c].dy[i] - qy * p.sy;
        prm.tap_qx[ntaps] = qx; prm.tap_c[ntaps] = (probs[c].dx[i] - qx * p.sx) * p.in_pitch;
        prm.taps[ntaps].plane = 0; prm.taps[ntaps].a_off = 0; prm.taps[ntaps].cls = c; prm.taps[ntaps].first = (i == 0) ? 1 : 0;
      }
    prm.ntaps = ntaps;
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
    prm.aempty_count = 4;     // the four stager warps of the group that consumed the tap release its shared-memory stage
    m_tiles = prm.tiles_x * prm.tiles_y * prm.tiles_b;
  } else {
  prm.aempty_count = kStagerWarps;
  // planes: taps grouped by stride parity
  struct PInfo { int ry, rx, qy_min, qy_max, qx_min, qx_max; };
  std::vector<PInfo> pinfo;
  struct TInfo { int plane, qy, qx, cls; };
  std::vector<TInfo> tinfo;
  for (int c = 0; c < nclass; ++c)
    for (int i = 0; i < probs[c].ntaps; ++i) {
      const int qy = floor_div_h(probs[c].dy[i], p.sy), ry = probs[c].dy[i] - qy * p.sy;
      const int qx = floor_div_h(probs[c].dx[i], p.sx), rx = probs[c].dx[i] - qx * p.sx;
      int pi = -1;
      for (size_t k = 0; k < pinfo.size(); ++k)
        if (pinfo[k].ry == ry && pinfo[k].rx == rx) pi = (int)k;
      if (pi < 0) { pinfo.push_back({ry, rx, qy, qy, qx, qx}); pi = (int)pinfo.size() - 1; }
      PInfo& pl = pinfo[pi];
      pl.qy_min = std::min(pl.qy_min, qy); pl.qy_max = std::max(pl.qy_max, qy);
      pl.qx_min = std::min(pl.qx_min, qx); pl.qx_max = std::max(pl.qx_max, qx);
      tinfo.push_back({pi, qy, qx, c});
    }
  if ((int)pinfo.size() > kMaxPlanes || (int)tinfo.size() > kMaxTaps) return false;
  prm.nplanes = (int)pinfo.size();
  prm.ntaps = (int)tinfo.size();
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
  int last_cls = -1;
  for (int t = 0; t < prm.ntaps; ++t) {
    HaloTap& tp = prm.taps[t];
    const HaloPlane& pl = prm.planes[tinfo[t].plane];
    tp.plane = tinfo[t].plane;
    tp.a_off = ((tinfo[t].qy - pl.qy_min) * pl.cols + (tinfo[t].qx - pl.qx_min)) * px_bytes;
    tp.cls = tinfo[t].cls;
    tp.first = (tp.cls != last_cls) ? 1 : 0;
    last_cls = tp.cls;
  }
  if (prm.cin8) {   // regroup: one K step = four consecutive taps
    prm.ntaps_real = prm.ntaps;
    for (int t = 0; t < prm.ntaps_real; ++t) { prm.g_plane[t] = prm.taps[t].plane; prm.g_aoff[t] = prm.taps[t].a_off; }
    prm.ntaps = (prm.ntaps_real + 3) / 4;
    for (int t = 0; t < prm.ntaps; ++t) { prm.taps[t].plane = 0; prm.taps[t].a_off = 0; prm.taps[t].cls = 0; prm.taps[t].first = (t == 0) ? 1 : 0; }
  }
  prm.tiles_x = ceil_div(p.Wo, kTileW); prm.tiles_y = ceil_div(p.Ho, kTileH); prm.tiles_b = p.B;
  m_tiles = prm.tiles_x * prm.tiles_y * p.B;
  }
  // TMEM budget (512 columns): nbuf accumulator buffers x nclass x acc_w  +  the A-operand ring, `st` slots of 64
  // columns (A_hi | A_lo of one tap), at least 4 slots so that the stagers run ahead of the tensor core.
  // acc_w = 2N in stacked 3xTF32 mode (big | small terms side by side, see the MMA role), N otherwise.
  const int cout16 = (p.Cout + 15) / 16 * 16;
  auto fits = [&](int n, int wm, int& nbuf_out) {
    for (int nb = 2; nb >= 1; --nb)
      if (nb * nclass * wm * n + 4 * 64 <= 512) { nbuf_out = nb; return true; }
    return false;
  };
  int n_tile = std::min(cout16, 256), nbuf = 2;
  while (n_tile >= 16 && !fits(n_tile, 1, nbuf)) n_tile = (n_tile > 32) ? (n_tile / 2 + 15) / 16 * 16 : n_tile - 16;
  if (n_tile_req > 0) n_tile = std::min(n_tile, n_tile_req);
  // narrow the N tile (down to 64) while the layer would leave SMs idle
  while (n_tile >= 128 && (n_tile % 32) == 0 && (long)m_tiles * ceil_div(p.Cout, n_tile) < 148) n_tile /= 2;
  if (n_tile < 16 || !fits(n_tile, 1, nbuf)) return false;
  prm.stacked = 0;
  int nbuf2 = 0;
  if (nsplit == 3 && n_tile <= 64 && fits(n_tile, 2, nbuf2)) { prm.stacked = 1; nbuf = nbuf2; }
  prm.nbuf = nbuf;
  prm.n_tile = n_tile;
  prm.acc_w = (prm.stacked ? 2 : 1) * n_tile;
  prm.n_tiles = ceil_div(p.Cout, n_tile);
  prm.st = std::min(kMaxTStages, (512 - prm.nbuf * nclass * prm.acc_w) / 64);
  int cols = 32;
  while (cols < prm.nbuf * nclass * prm.acc_w + prm.st * 64) cols <<= 1;
  prm.tmem_cols = cols;
  prm.k_chunks = prm.cin8 ? 1 : p.Cin / 32;
  // weight ring slot = one (chunk, tap) block: [W_hi | W_lo] (3xTF32) or W_hi alone, 1024-byte aligned
  const int slot = (nsplit == 3) ? n_tile * 256 : (n_tile * 128 + 1023) / 1024 * 1024;
  prm.w_stage_bytes = slot;
  // shared memory: up to 4 halo stages (at least 2), the rest for the weight ring
  const int budget = 224 * 1024 - 4 * kEpiStageBytes;   // minus the epilogue's transpose buffers
  prm.sa = kMaxAStages;
  while (prm.sa > 2 && budget - prm.sa * prm.a_region_bytes < 4 * slot) --prm.sa;
  const int rest = budget - prm.sa * prm.a_region_bytes;
  if (rest < 2 * slot) return false;
  prm.sw = std::min(std::min(kMaxWStages, prm.st), rest / slot);   // sw <= st, see the W producer
  plan.smem_bytes = prm.sa * prm.a_region_bytes + prm.sw * slot + 4 * kEpiStageBytes + 1024;
  prm.B = p.B;
  prm.total_tiles = m_tiles * prm.n_tiles;
  auto fd = [](int d) { return d <= 1 ? 0u : (uint32_t)((1ull << 32) / (uint64_t)d + 1ull); };
  prm.mul_n_tiles = fd(prm.n_tiles); prm.mul_tiles_x = fd(prm.tiles_x); prm.mul_tiles_y = fd(prm.tiles_y);
  if ((uint64_t)prm.total_tiles * (uint64_t)std::max(prm.n_tiles, std::max(prm.tiles_x, prm.tiles_y)) >= (1ull << 32)) return false;
  prm.out = p.out; prm.out_pitch = p.out_pitch; prm.Ho = p.Ho; prm.Wo = p.Wo; prm.Hfull = p.Hfull; prm.Wfull = p.Wfull;
  prm.osy = p.osy; prm.osx = p.osx; prm.Cout = p.Cout; prm.bias = p.bias; prm.leaky = p.leaky;
  for (int c = 0; c < nclass; ++c) { prm.cls_ooy[c] = probs[c].ooy; prm.cls_oox[c] = probs[c].oox; }
  if (!encode) return true;
  // TMA descriptors: same 5-D view as conv_tc.cu, one box shape per plane
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
  const ConvProblem& p = probs[0];
  for (int c = 0; c < nclass; ++c) {
    ConvProblem q = probs[c];
    if (q.Cin == 8 && q.in_pitch == 8) q.Cin = 32;   // 8-channel mode: everything but the channel rule must hold
    if (!tc_layer_supported(q)) return false;
  }
  // per-tap TS mode pays off for plain convolutions at low resolution; the 4-class transposed convolutions there are
  // TMEM-limited (single accumulator buffer, split N) and stay on the shared-memory-operand kernel (conv_tc.cu)
  if (((p.Ho % kTileH) != 0 || (p.Wo % kTileW) != 0) && nclass != 1) return false;
  HaloPlan plan;
  return halo_build(probs, nclass, 0, 3, plan, false);
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
  if (!halo_build(probs, nclass, 0, nsplit, *plan, true)) {
    delete plan;
    return fail(DEMON_E_CUDA, "tc_halo_prepare: could not build the plan (tensor map encode failed?)");
  }
  const HaloParams& prm = plan->prm;
  // weights: [n_tile][chunk][tap (class major)] blocks of [W_hi | W_lo], n_tile rows x 32 fp32, K-major, pre-swizzled
  const int slot = prm.w_stage_bytes;
  const size_t total = (size_t)prm.n_tiles * prm.k_chunks * prm.ntaps * slot;
  std::vector<unsigned char> packed(total, 0);
  std::vector<int> tap_in_class(prm.ntaps);
  {
    int cnt[4] = {0, 0, 0, 0};
    for (int tt = 0; tt < prm.ntaps; ++tt) tap_in_class[tt] = cnt[prm.taps[tt].cls]++;
  }
  for (int nt = 0; nt < prm.n_tiles; ++nt)
    for (int kc = 0; kc < prm.k_chunks; ++kc)
      for (int tt = 0; tt < prm.ntaps; ++tt) {
        unsigned char* blk = packed.data() + ((size_t)(nt * prm.k_chunks + kc) * prm.ntaps + tt) * slot;
        const int cls = prm.taps[tt].cls, tap = tap_in_class[tt];
        for (int r = 0; r < prm.n_tile; ++r) {
          const int co = nt * prm.n_tile + r;
          for (int k = 0; k < 32; ++k) {
            float w = 0.f;
            if (prm.cin8) {   // K index = (tap within the group of four, channel)
              const int rt = 4 * tt + k / 8;
              if (co < p.Cout && rt < prm.ntaps_real) w = w_hosts[0][((size_t)rt * 8 + (k & 7)) * p.Cout_pad + co];
            } else
            if (co < p.Cout) w = w_hosts[cls][((size_t)tap * p.Cin + kc * 32 + k) * p.Cout_pad + co];
            const float hi = (nsplit == 3) ? tf32_round_h(w) : w;
            const float lo = w - hi;
            const size_t off = (size_t)r * 128 + (size_t)(((k >> 2) ^ (r & 7)) << 4) + (size_t)(k & 3) * 4;
            memcpy(blk + off, &hi, 4);
            if (nsplit == 3) memcpy(blk + (size_t)prm.n_tile * 128 + off, &lo, 4);
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
  return DEMON_OK;
}

void tc_halo_free(TcLayer& t) {
  if (t.halo_plan) delete static_cast<HaloPlan*>(t.halo_plan);
  t.halo_plan = nullptr;
}

extern int* tc_error_flag();

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

int conv_tc_halo_launch(const TcLayer& t, const ConvProblem* probs, cudaStream_t stream) {
  HaloPlan* plan = static_cast<HaloPlan*>(t.halo_plan);
  HaloParams prm = plan->prm;
  prm.out = probs[0].out;          // the output slice may be re-pointed between calls (caller-owned result buffers)
  prm.out_pitch = probs[0].out_pitch;
  prm.err = tc_error_flag();
  prm.timing = g_timing_dev;
  static bool attr_set = false;
  if (!attr_set) {
    DEMON_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_halo_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    DEMON_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_halo_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    DEMON_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_halo_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int grid = std::min(prm.total_tiles, sms);
  if (prm.per_tap) conv_tc_halo_kernel<true, false><<<grid, kThreads, plan->smem_bytes, stream>>>(plan->maps, prm);
  else if (prm.cin8) conv_tc_halo_kernel<false, true><<<grid, kThreads, plan->smem_bytes, stream>>>(plan->maps, prm);
  else conv_tc_halo_kernel<false, false><<<grid, kThreads, plan->smem_bytes, stream>>>(plan->maps, prm);
  DEMON_LAUNCH_CHECK();
  return DEMON_OK;
}

}  // namespace demon
