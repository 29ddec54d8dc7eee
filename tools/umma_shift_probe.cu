// Hardware probe (sm_100a): how does tcgen05.mma address a K-major SWIZZLE_128B operand whose start address is NOT
// 1024-byte aligned (shifted by whole 128-byte rows), with and without the descriptor's base_offset field, and with a
// stride-byte-offset that is not a multiple of 1024?  Answers whether one halo tile in shared memory can serve all the
// taps of a convolution.  Prints, for each variant, which smem row every MMA row read and whether the K order was intact.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

struct Variant { int shift_rows; int base_offset; int sbo; };

__global__ void __launch_bounds__(128, 1) probe(const Variant* vars, int nvar, float* out) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  float* A = (float*)smem;                 // 512 rows x 32 floats, swizzled by ABSOLUTE row index
  float* Bm = (float*)(smem + 512 * 128);  // 32 rows (n) x 32 floats (k): identity
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 512 * 32; i += 128) {
    const int r = i / 32, k = i % 32;
    const float v = (k < 16) ? (float)r : (float)k;
    const uint32_t off = r * 128 + (((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4;
    *(float*)(smem + off) = v;
  }
  for (int i = tid; i < 32 * 32; i += 128) {
    const int n = i / 32, k = i % 32;
    const uint32_t off = n * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4;
    *(float*)((unsigned char*)Bm + off) = (n == k) ? 1.f : 0.f;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  uint32_t phase = 0;
  for (int v = 0; v < nvar; ++v) {
    if (tid == 0) {
      const uint32_t a_addr = smem_u32(A) + vars[v].shift_rows * 128;
      for (int j = 0; j < 4; ++j) {
        const uint64_t ad = make_desc(a_addr, vars[v].sbo, vars[v].base_offset) + (uint64_t)(2 * j);
        const uint64_t bd = make_desc(smem_u32(Bm), 1024, 0) + (uint64_t)(2 * j);
        const uint32_t acc = j != 0;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
                     "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                     : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
    }
    phase ^= 1;
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; ++c) out[((size_t)v * 128 + tid) * 32 + c] = __uint_as_float(r[c]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  std::vector<Variant> vars = {{0, 0, 1024}, {8, 0, 1024}, {1, 0, 1024}, {1, 1, 1024}, {3, 0, 1024}, {3, 3, 1024}, {5, 5, 1024},
                               {0, 0, 1280}, {1, 0, 1280}, {1, 1, 1280}, {0, 0, 2048}, {2, 0, 2048}, {2, 2, 2048}, {16, 0, 2048}};
  Variant* dv;
  float* dout;
  cudaMalloc(&dv, vars.size() * sizeof(Variant));
  cudaMemcpy(dv, vars.data(), vars.size() * sizeof(Variant), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, vars.size() * 128 * 32 * sizeof(float));
  const int smem = 512 * 128 + 32 * 128 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(dv, (int)vars.size(), dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> h(vars.size() * 128 * 32);
  cudaMemcpy(h.data(), dout, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
  for (size_t v = 0; v < vars.size(); ++v) {
    const int groups_pitch = vars[v].sbo / 128;
    int rows_ok = 0, korder_ok = 0;
    for (int m = 0; m < 128; ++m) {
      const float* d = &h[(v * 128 + m) * 32];
      const int expect_row = vars[v].shift_rows + (m / 8) * groups_pitch + (m % 8);
      bool rok = true, kok = true;
      for (int c = 0; c < 16; ++c) rok = rok && d[c] == (float)expect_row;
      for (int c = 16; c < 32; ++c) kok = kok && d[c] == (float)c;
      rows_ok += rok; korder_ok += kok;
    }
    printf("shift %2d base_offset %d sbo %4d : rows as expected %3d/128, k order intact %3d/128 | m=0..9 row ids:", vars[v].shift_rows,
           vars[v].base_offset, vars[v].sbo, rows_ok, korder_ok);
    for (int m = 0; m < 10; ++m) printf(" %g", h[(v * 128 + m) * 32]);
    printf(" | m=1 cols:");
    for (int c = 0; c < 32; c += 4) printf(" %g", h[(v * 128 + 1) * 32 + c]);
    printf("\n");
  }
  return 0;
}
