// Hardware probe (sm_100a): issue/execute rate of back-to-back tcgen05.mma kind::tf32 (M = 128, K = 8) as a function
// of N, with the A operand in shared memory (SS) or in TMEM (TS).  One CTA, one issuing thread; prints cycles per MMA.
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

struct Res { long long cycles; int n; int ts; int iters; int commit_every; int rot; int acc; };

__global__ void __launch_bounds__(128, 1) probe(Res* res, int nres) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 128 + 256 * 128) / 4; i += 128) ((float*)smem)[i] = 1.0f;
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
  if (warp == 1) {
    if (elect_one()) {
      uint32_t phase = 0;
      for (int v = 0; v < nres; ++v) {
        const int n = res[v].n, ts = res[v].ts, iters = res[v].iters, ce = res[v].commit_every, rot = res[v].rot; const uint32_t acc = (uint32_t)res[v].acc;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t ad = make_desc(smem_u32(smem)), bd = make_desc(smem_u32(smem + 128 * 128));
        const uint32_t a_t = tmem + 448, d_t0 = tmem;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
          const uint64_t adv = (uint64_t)(2 * (i & 3));
          const uint32_t d_t = d_t0 + (uint32_t)((i & (rot - 1)) * n);
          if (ts)
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(d_t),
                         "r"(a_t + 8 * (i & 3)), "l"(bd + adv), "r"(idesc), "r"(acc) : "memory");
          else
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_t),
                         "l"(ad + adv), "l"(bd + adv), "r"(idesc), "r"(acc) : "memory");
        }
        const long long t_issue = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        phase ^= 1;
        const long long t1 = clock64();
        res[v].cycles = t1 - t0;
        res[v].iters = (int)(t_issue - t0);   // reuse: cycles spent issuing
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  std::vector<Res> r;
  const int iters = 2000;
  for (int ts = 0; ts < 2; ++ts)
    for (int n : {16, 64, 256}) r.push_back({0, n, ts, iters, 0, 1, 1});
  for (int ts = 0; ts < 2; ++ts)
    for (int rot : {2, 4})
      for (int n : {16, 64}) r.push_back({0, n, ts, iters, 0, rot, 1});      // rotate over independent accumulators
  for (int ts = 0; ts < 2; ++ts)
    for (int n : {16, 64}) r.push_back({0, n, ts, iters, 0, 1, 0});          // no accumulate (overwrite D)
  Res* d;
  cudaMalloc(&d, r.size() * sizeof(Res));
  cudaMemcpy(d, r.data(), r.size() * sizeof(Res), cudaMemcpyHostToDevice);
  const int smem = 128 * 128 + 256 * 128 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(d, (int)r.size());
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<Res> h(r.size());
  cudaMemcpy(h.data(), d, r.size() * sizeof(Res), cudaMemcpyDeviceToHost);
  for (size_t i = 0; i < h.size(); ++i)
    printf("%s N=%3d accumulators=%d accumulate=%d : %7.1f cycles/MMA total, %6.1f cycles/MMA issuing   (math at 2048 MAC/clk: %5.1f)\n",
           h[i].ts ? "TS" : "SS", h[i].n, r[i].rot, r[i].acc, (double)h[i].cycles / iters, (double)h[i].iters / iters, 128.0 * h[i].n * 8 / 2048);
  return 0;
}
