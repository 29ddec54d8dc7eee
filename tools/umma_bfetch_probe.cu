// Hardware probe (sm_100a): what does a TS-mode tcgen05.mma kind::tf32 (M = 128, K = 8) pay for its B operand in shared
// memory?  Stacked 3xTF32 steps (4 x [N' = 2N, N] MMAs) whose B descriptor (a) never changes, (b) cycles over four ring
// slots (a new weight block every step, as in conv_tc_halo_kernel), with 128-byte-swizzled K-major rows (a K8 slice is
// 32 bytes of every 128-byte row) or (c) compact 32-byte-swizzled rows (a K8 slice is contiguous).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.b32 %0, 1, 0, px;\n}" : "=r"(pred));
  return pred != 0;
}
struct Cfg { int n, slots, layout, steps; long long cycles; };

template <int SLOTS, int LAYOUT>
__device__ void run(Cfg& c, uint32_t tmem, unsigned char* smem, uint32_t bar, uint32_t& phase) {
  const int n = c.n;
  const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * n) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t sbo = (LAYOUT == 2) ? 1024u : 256u;
  const uint64_t kstep = (LAYOUT == 2) ? 2 : (uint64_t)((2 * n * 32) >> 4);   // next K8 slice: +32 B in the row / next compact block
  const uint64_t bd0 = make_desc(smem_u32(smem), sbo, LAYOUT);
  const long long t0 = clock64();
#pragma unroll 1
  for (int s = 0; s < c.steps; ++s) {
    const uint64_t bd = bd0 + (uint64_t)((s & (SLOTS - 1)) * (16384 >> 4));
    const uint32_t a_hi = tmem + 256 + (uint32_t)((s & 3) * 64), a_lo = a_hi + 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem),
                   "r"(a_hi + 8 * j), "l"(bd + kstep * j), "r"(idesc2), "r"(1u) : "memory");
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem + (uint32_t)n),
                   "r"(a_lo + 8 * j), "l"(bd + kstep * j), "r"(idesc1), "r"(1u) : "memory");
    }
  }
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
  phase ^= 1;
  c.cycles = clock64() - t0;
}

__global__ void __launch_bounds__(128, 1) probe(Cfg* cfgs, int ncfg) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
  if (warp == 1 && elect_one()) {
    uint32_t phase = 0;
    for (int v = 0; v < ncfg; ++v) {
      Cfg c = cfgs[v];
      if (c.slots == 1 && c.layout == 2) run<1, 2>(c, tmem, smem, smem_u32(&bar), phase);
      else if (c.slots == 4 && c.layout == 2) run<4, 2>(c, tmem, smem, smem_u32(&bar), phase);
      else if (c.slots == 1 && c.layout == 6) run<1, 6>(c, tmem, smem, smem_u32(&bar), phase);
      else run<4, 6>(c, tmem, smem, smem_u32(&bar), phase);
      cfgs[v].cycles = c.cycles;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  std::vector<Cfg> c;
  for (int layout : {2, 6})
    for (int slots : {1, 4})
      for (int n : {16, 32, 64}) c.push_back({n, slots, layout, 2000, 0});
  Cfg* d;
  cudaMalloc(&d, c.size() * sizeof(Cfg));
  cudaMemcpy(d, c.data(), c.size() * sizeof(Cfg), cudaMemcpyHostToDevice);
  const int smem = 65536 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(d, (int)c.size());
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(c.data(), d, c.size() * sizeof(Cfg), cudaMemcpyDeviceToHost);
  for (auto& x : c)
    printf("N=%2d  B %s, %s : %7.1f cycles per stacked step (8 MMAs, tensor math alone: %d)\n", x.n, x.slots == 1 ? "fixed      " : "4 ring slots",
           x.layout == 2 ? "128B swizzle" : "32B swizzle ", (double)x.cycles / x.steps, 6 * x.n);
  return 0;
}
