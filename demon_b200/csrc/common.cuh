// Shared helpers of the demon_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/demon_b200.h"

namespace demon {

// ---- error plumbing: C ABI returns a code, message kept thread-local -------------------------
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launch_count;

#define DEMON_CHECK_CUDA(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::demon::fail(DEMON_E_CUDA, "%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, \
                           cudaGetErrorString(_e));                                         \
  } while (0)

// Every kernel launch goes through this: counts launches (gpu_launches in bench.py) and turns a
// launch failure into an error code, like _CHECK_CUDA_ERROR did in the reference
// (lmbspecialops/src/cuda_helper.h:25-35) but without throwing across the C ABI.
#define DEMON_LAUNCH_CHECK()                                                                    \
  do {                                                                                          \
    ::demon::g_launch_count.fetch_add(1, std::memory_order_relaxed);                            \
    cudaError_t _e = cudaPeekAtLastError();                                                     \
    if (_e != cudaSuccess) {                                                                    \
      cudaGetLastError();                                                                       \
      return ::demon::fail(DEMON_E_CUDA, "%s:%d: kernel launch failed: %s", __FILE__, __LINE__, \
                           cudaGetErrorString(_e));                                             \
    }                                                                                           \
  } while (0)

#define DEMON_REQUIRE(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return ::demon::fail(DEMON_E_INVALID, __VA_ARGS__); \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- IEEE ops without FMA contraction ---------------------------------------------------------
// The geometry ops restate x86 code compiled without FMA; using the _rn intrinsics keeps nvcc from
// contracting a*b+c so the float results match the CPU oracle operation for operation.
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double fmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double fadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double fsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double fdiv(double a, double b) { return __ddiv_rn(a, b); }

}  // namespace demon
