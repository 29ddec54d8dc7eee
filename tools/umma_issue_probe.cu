// Hardware probe (sm_100a): what bounds a stream of narrow TS-mode tcgen05.mma kind::tf32 "stacked 3xTF32" steps
// (4 x [N' = 2N, N] MMAs per step, as conv_tc_halo_kernel issues them) when every step also pays one tcgen05.commit and one
// mbarrier poll?  Variants: (a) commits / polls only every `every`-th step, (b) two accumulator sets used in turns by ONE
// issuing thread (is it the dependent accumulation?), (c) TWO issuing threads (warps 1 and 14) with own accumulators and own
// barriers (is it the issuing thread's instruction stream?).  Prints cycles per step over all steps issued.
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

struct Cfg { int mma_steps, n, st_warps, ld_warps, st_iters, ld_iters, st_per_iter, commits, polls, every, nacc, nthreads, fence, a_base; };
struct Out { long long mma_cycles, st_cycles, ld_cycles; long long st_done, ld_done; long long mma2_cycles; };

__global__ void __launch_bounds__(480, 1) probe(const Cfg* cfgs, Out* outs, int ncfg, float* sink) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar, bar2;
  __shared__ __align__(8) uint64_t cbar2[4];
  __shared__ __align__(8) uint64_t cbar[4];
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (256 * 128) / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[i])));
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar2[i])));
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
    if (warp == 1 || warp == 14) {
      const int me = (warp == 1) ? 0 : 1;
      if (c.mma_steps > 0 && me < c.nthreads && elect_one()) {
        const int n = c.n;
        const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * n) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t bd = make_desc(smem_u32(smem));
        uint64_t* my_cbar = me ? cbar2 : cbar;
        uint64_t* my_bar = me ? &bar2 : &bar;
        const uint32_t acc0 = tmem + (uint32_t)(me * 128);   // accumulators of this thread: columns [me*128, me*128 + 128)
        const long long t0 = clock64();
        for (int s = 0; s < c.mma_steps; ++s) {
          if (c.fence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = tmem + (uint32_t)(c.a_base ? c.a_base : 256) + (uint32_t)((s & 3) * 64), a_lo = a_hi + 32;
          const uint32_t d = acc0 + (uint32_t)((c.nacc > 1 ? (s % c.nacc) : 0) * 2 * n);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(d),
                         "r"(a_hi + 8 * j), "l"(bd + (uint64_t)(2 * j)), "r"(idesc2), "r"(1u) : "memory");
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(d + (uint32_t)n),
                         "r"(a_lo + 8 * j), "l"(bd + (uint64_t)(2 * j)), "r"(idesc1), "r"(1u) : "memory");
          }
          if ((s % c.every) == c.every - 1) {
            for (int k = 0; k < c.commits; ++k)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&my_cbar[k])) : "memory");
            for (int k = 0; k < c.polls; ++k) {   // a try_wait on an mbarrier whose awaited phase is already complete
              uint32_t ok;
              asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&my_cbar[3])), "r"(1u) : "memory");
              if (!ok) sink[0] = 1.f;
            }
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(my_bar)) : "memory");
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(my_bar)), "r"(phase) : "memory");
        if (me) outs[v].mma2_cycles = clock64() - t0; else outs[v].mma_cycles = clock64() - t0;
        if (!me) stop = 1;
      }
      if (c.mma_steps > 0 && me < c.nthreads) phase ^= 1;
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
  //                      steps n  stw ldw sti ldi stpi commits polls every nacc nthreads
  for (int n : {16, 32}) {
    c.push_back({steps, n, 0, 0, 0, 0, 2, 0, 0, 1, 1, 1});   // MMAs alone, one accumulator set (dependent accumulation)
    c.push_back({steps, n, 0, 0, 0, 0, 2, 0, 0, 1, 2, 1});   // two accumulator sets in turns
    c.push_back({steps, n, 0, 0, 0, 0, 2, 1, 1, 1, 1, 1});   // + commit + poll every step (what the kernel does)
    c.push_back({steps, n, 0, 0, 0, 0, 2, 1, 1, 2, 1, 1});   // ... every second step
    c.push_back({steps, n, 0, 0, 0, 0, 2, 1, 1, 4, 1, 1});   // ... every fourth step
    c.push_back({steps, n, 0, 0, 0, 0, 2, 1, 1, 1, 2, 1});   // every step, two accumulator sets
    c.push_back({steps, n, 0, 0, 0, 0, 2, 0, 0, 1, 1, 2});   // two issuing threads, MMAs alone
    c.push_back({steps, n, 0, 0, 0, 0, 2, 1, 1, 1, 1, 2});   // two issuing threads, commit + poll every step
    c.push_back({steps, n, 0, 0, 0, 0, 2, 1, 1, 2, 1, 2});   // two issuing threads, every second step
    c.push_back({steps, n, 8, 4, 0, 0, 2, 1, 1, 1, 1, 1});   // one thread, with st + ld traffic
    c.push_back({steps, n, 8, 4, 0, 0, 2, 1, 1, 1, 1, 2});   // two threads, with st + ld traffic
    c.push_back({steps, n, 0, 0, 0, 0, 2, 1, 1, 1, 1, 1, 1});   // one thread, commit + poll + tcgen05.fence::after_thread_sync every step
    c.push_back({steps, n, 0, 0, 0, 0, 2, 1, 1, 1, 1, 2, 1});   // two threads, the same
    c.push_back({steps, n, 8, 4, 0, 0, 2, 1, 1, 1, 1, 1, 1});   // one thread, fence, with st + ld traffic
    c.push_back({steps, n, 0, 0, 0, 0, 2, 0, 0, 1, 1, 2, 0, 64});    // two threads, MMAs alone, A ring right behind the accumulators (columns 64..319)
    c.push_back({steps, n, 0, 0, 0, 0, 2, 1, 1, 1, 1, 2, 0, 64});    // ... with commit + poll
    c.push_back({steps, n, 0, 0, 0, 0, 2, 1, 1, 1, 1, 1, 0, 64});    // one thread, commit + poll, A ring at 64
  }
  Cfg* dc; Out* dout; float* sink;
  cudaMalloc(&dc, c.size() * sizeof(Cfg));
  cudaMalloc(&dout, c.size() * sizeof(Out));
  cudaMalloc(&sink, 4096);
  cudaMemset(dout, 0, c.size() * sizeof(Out));
  cudaMemcpy(dc, c.data(), c.size() * sizeof(Cfg), cudaMemcpyHostToDevice);
  const int smem = 256 * 128 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 480, smem>>>(dc, dout, (int)c.size(), sink);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<Out> h(c.size());
  cudaMemcpy(h.data(), dout, c.size() * sizeof(Out), cudaMemcpyDeviceToHost);
  for (size_t i = 0; i < h.size(); ++i) {
    const long long cyc = h[i].mma_cycles > h[i].mma2_cycles ? h[i].mma_cycles : h[i].mma2_cycles;
    printf("N=%2d threads=%d accumulator sets=%d commit+poll=%d every %d step(s) fence=%d A ring at column %d st_warps=%d ld_warps=%d : %7.1f cycles per step (all %d steps of %d thread(s))\n",
           c[i].n, c[i].nthreads, c[i].nacc, c[i].commits, c[i].every, c[i].fence, c[i].a_base ? c[i].a_base : 256, c[i].st_warps, c[i].ld_warps, (double)cyc / (c[i].mma_steps * c[i].nthreads),
           c[i].mma_steps * c[i].nthreads, c[i].nthreads);
  }
  return 0;
}
