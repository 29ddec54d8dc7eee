// Hardware probe (sm_100a): does tcgen05.st / tcgen05.ld traffic of other warps slow down a stream of TS-mode
// tcgen05.mma kind::tf32 (A operand in TMEM)?  One CTA: warp 1 issues "stacked 3xTF32" steps (4 x [N'=2N, N] MMAs,
// one commit per step), warps 2..9 write 2 x 32 columns per iteration with tcgen05.st (what the stager warps of
// conv_tc_halo_kernel do), warps 10..13 read 32 columns with tcgen05.ld (the epilogue).  Prints cycles per step / per
// iteration for every combination.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.b32 %0, 1, 0, px;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
      "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

struct Cfg { int mma_steps, n, st_warps, ld_warps, st_iters, ld_iters, st_per_iter, commits, polls; };
struct Out { long long mma_cycles, st_cycles, ld_cycles; long long st_done, ld_done; };

__global__ void __launch_bounds__(448, 1) probe(const Cfg* cfgs, Out* outs, int ncfg, float* sink) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t cbar[4];
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (256 * 128) / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  uint32_t phase = 0;
  for (int v = 0; v < ncfg; ++v) {
    const Cfg c = cfgs[v];
    if (tid == 0) stop = 0;
    __syncthreads();
    if (warp == 1) {
      if (c.mma_steps > 0 && elect_one()) {
        const int n = c.n;
        const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * n) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t bd = make_desc(smem_u32(smem));
        const long long t0 = clock64();
        for (int s = 0; s < c.mma_steps; ++s) {
          const uint32_t a_hi = tmem + 256 + (uint32_t)((s & 3) * 64), a_lo = a_hi + 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem),
                         "r"(a_hi + 8 * j), "l"(bd + (uint64_t)(2 * j)), "r"(idesc2), "r"(1u) : "memory");
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem + (uint32_t)n),
                         "r"(a_lo + 8 * j), "l"(bd + (uint64_t)(2 * j)), "r"(idesc1), "r"(1u) : "memory");
          }
          for (int k = 0; k < c.commits; ++k)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&cbar[k])) : "memory");
          for (int k = 0; k < c.polls; ++k) {   // a try_wait on an mbarrier whose awaited phase is already complete
            uint32_t ok;
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&cbar[3])), "r"(1u) : "memory");
            if (!ok) sink[0] = 1.f;
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        outs[v].mma_cycles = clock64() - t0;
        stop = 1;
      }
      if (c.mma_steps > 0) phase ^= 1;
    } else if (warp >= 2 && warp < 10) {
      if (warp - 2 < c.st_warps) {
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (uint32_t)(((warp - 2) >> 2) * 128);
        uint32_t vals[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) vals[i] = __float_as_uint(1.0f + lane + i);
        const long long t0 = clock64();
        long long it = 0;
        for (;; ++it) {
          if (c.mma_steps > 0 ? stop : (it >= c.st_iters)) break;
          st32(taddr, vals);
          if (c.st_per_iter > 1) st32(taddr + 32, vals);
          if (c.st_per_iter > 2) { st32(taddr + 64, vals); st32(taddr + 96, vals); }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (lane == 0 && warp == 2) { outs[v].st_cycles = clock64() - t0; outs[v].st_done = it; }
      }
    } else if (warp >= 10 && warp < 14) {
      if (warp - 10 < c.ld_warps) {
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t vals[32];
        float acc = 0.f;
        const long long t0 = clock64();
        long long it = 0;
        for (;; ++it) {
          if (c.mma_steps > 0 ? stop : (it >= c.ld_iters)) break;
          ld32(taddr + (uint32_t)((it & 1) * 32), vals);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          acc += __uint_as_float(vals[it & 31]);
        }
        if (lane == 0 && warp == 10) { outs[v].ld_cycles = clock64() - t0; outs[v].ld_done = it; }
        if (acc == 123.456f) sink[tid] = acc;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  std::vector<Cfg> c;
  const int steps = 2000;
  for (int n : {16, 32, 64}) c.push_back({steps, n, 0, 0, 0, 0, 2, 0, 0});   // MMA alone
  for (int n : {16, 32})
    for (int cm : {1, 2, 3}) c.push_back({steps, n, 0, 0, 0, 0, 2, cm, 0});    // + commits per step
  for (int n : {16, 32})
    for (int pl : {1, 2}) c.push_back({steps, n, 0, 0, 0, 0, 2, 0, pl});       // + completed try_wait polls per step
  c.push_back({steps, 16, 0, 0, 0, 0, 2, 2, 2});
  c.push_back({steps, 32, 0, 0, 0, 0, 2, 2, 2});
  c.push_back({steps, 32, 8, 0, 0, 0, 2, 2, 2});
  for (int w : {4, 8}) c.push_back({0, 32, w, 0, 4000, 0, 2});               // st alone (2 x 32 columns per iteration)
  for (int w : {4, 8}) c.push_back({0, 32, w, 0, 4000, 0, 4});               // st alone (4 x 32 columns per iteration)
  c.push_back({0, 32, 4, 0, 4000, 0, 1, 0, 0});                                     // st alone (1 x 32 columns)
  c.push_back({0, 32, 0, 4, 0, 4000, 2, 0, 0});                                     // ld alone
  for (int n : {16, 32, 64})
    for (int w : {4, 8}) c.push_back({steps, n, w, 0, 0, 0, 2});              // MMA + st
  for (int n : {16, 32}) c.push_back({steps, n, 0, 4, 0, 0, 2});              // MMA + ld
  for (int n : {16, 32}) c.push_back({steps, n, 8, 4, 0, 0, 2});              // MMA + st + ld
  Cfg* dc; Out* dout; float* sink;
  cudaMalloc(&dc, c.size() * sizeof(Cfg));
  cudaMalloc(&dout, c.size() * sizeof(Out));
  cudaMalloc(&sink, 4096);
  cudaMemset(dout, 0, c.size() * sizeof(Out));
  cudaMemcpy(dc, c.data(), c.size() * sizeof(Cfg), cudaMemcpyHostToDevice);
  const int smem = 256 * 128 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 448, smem>>>(dc, dout, (int)c.size(), sink);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<Out> h(c.size());
  cudaMemcpy(h.data(), dout, c.size() * sizeof(Out), cudaMemcpyDeviceToHost);
  for (size_t i = 0; i < h.size(); ++i) {
    printf("N=%2d mma=%d commits=%d polls=%d st_warps=%d (x%d) ld_warps=%d :", c[i].n, c[i].mma_steps > 0, c[i].commits, c[i].polls, c[i].st_warps, c[i].st_per_iter, c[i].ld_warps);
    if (c[i].mma_steps) printf("  %7.1f cycles per stacked step (8 MMAs)", (double)h[i].mma_cycles / c[i].mma_steps);
    if (c[i].st_warps && h[i].st_done) printf("  st: %7.1f cycles per iteration per warp", (double)h[i].st_cycles / h[i].st_done);
    if (c[i].ld_warps && h[i].ld_done) printf("  ld: %7.1f cycles per x32 load per warp", (double)h[i].ld_cycles / h[i].ld_done);
    printf("\n");
  }
  return 0;
}
