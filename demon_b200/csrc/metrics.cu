// Evaluation metrics of the DeMoN path on the device: the masked per-sample sums behind
// depthmotionnet.evaluation.metrics.compute_errors / evaluate_depth / compute_flow_epe
// (python/depthmotionnet/evaluation/metrics.py:25-38,62-237,240-372,377-387), so that a dataset-level accuracy run never
// copies a depth map to the host: one pass over prediction and ground truth produces every sum the eleven distances and
// the least-squares scale factor need.  HBM-bound streaming reductions: 8 bytes in per pixel, nothing out but
// [n][16] doubles.
//
// Element-wise arithmetic is float32 with IEEE operations like numpy's (reciprocal, division, subtraction; log / log10 are
// CUDA's logf / log10f, within 1-2 ulp of numpy's), accumulation is double in a FIXED order (per-thread strided partial
// sums -> warp shuffle tree -> per-CTA slots -> one warp folds the slots in index order), so results are deterministic
// run to run; the reference accumulates pairwise in float32, which is where the documented 1e-5 tolerance comes from.
#include "common.cuh"
#include <cmath>

namespace demon {
namespace {

constexpr int kSums = 16;
constexpr int kMetricThreads = 256;
constexpr int kMaxSlots = 64;   // CTAs per sample

__device__ __forceinline__ bool valid_pair(float a, float b) { return isfinite(a) && isfinite(b) && a > 0.f && b > 0.f; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// partial[n][slot][kSums]
__global__ void __launch_bounds__(kMetricThreads) depth_sums_kernel(const float* __restrict__ pred, const float* __restrict__ gt, long hw,
                                                                   bool inverse_pred, bool inverse_gt, const float* __restrict__ gt_div,
                                                                   const float* __restrict__ pred_scale, double* __restrict__ partial) {
  const int n = blockIdx.y, slot = blockIdx.x, nslots = gridDim.x;
  const float* p = pred + (long)n * hw;
  const float* g = gt + (long)n * hw;
  const float gdiv = gt_div ? __ldg(gt_div + n) : 1.0f;
  const float pscale = pred_scale ? __ldg(pred_scale + n) : 1.0f;
  const float l125 = logf(1.25f), l156 = logf(1.5625f), l195 = logf(1.953125f);
  double acc[kSums];
#pragma unroll
  for (int k = 0; k < kSums; ++k) acc[k] = 0.0;
  for (long i = (long)slot * kMetricThreads + threadIdx.x; i < hw; i += (long)nslots * kMetricThreads) {
    const float pi = __ldg(p + i), gi = __ldg(g + i);
    if (!valid_pair(pi, gi)) continue;                       // compute_valid_depth_mask on the inputs (metrics.py:337)
    float dp = inverse_pred ? fdiv(1.0f, pi) : pi;           // metrics.py:339-342
    float dg = inverse_gt ? fdiv(1.0f, gi) : gi;
    if (gt_div) dg = fdiv(dg, gdiv);                         // metrics.py:349-355
    // least-squares scale factor sums on the UNSCALED prediction (metrics.py:283-318)
    const float pp = fmul(dp, dp), pg = fmul(dp, dg);
    if (isfinite(pg) && pg > 0.f) { acc[12] += (double)pp; acc[13] += (double)pg; }
    const float ip = fdiv(1.0f, dp), ig = fdiv(1.0f, dg);
    const float ipp = fmul(ip, ip), ipg = fmul(ip, ig);
    if (isfinite(ipg) && ipg > 0.f) { acc[14] += (double)ipp; acc[15] += (double)ipg; }
    if (pred_scale) dp = fmul(dp, pscale);                   // metrics.py:362
    if (!valid_pair(dp, dg)) continue;                       // compute_errors masks again (metrics.py:252)
    const float d = fsub(dp, dg);
    const float ld = fsub(logf(dp), logf(dg));
    acc[0] += 1.0;
    acc[1] += (double)fabsf(d);
    acc[2] += (double)fabsf(fsub(fdiv(1.0f, dp), fdiv(1.0f, dg)));
    acc[3] += (double)ld;
    acc[4] += (double)fmul(ld, ld);
    acc[5] += (double)fdiv(fabsf(d), dg);
    acc[6] += (double)fdiv(fmul(d, d), dg);
    acc[7] += (double)fabsf(fsub(log10f(dp), log10f(dg)));
    acc[8] += (double)fmul(d, d);
    const float ald = fabsf(ld);
    acc[9] += (ald < l125) ? 1.0 : 0.0;
    acc[10] += (ald < l156) ? 1.0 : 0.0;
    acc[11] += (ald < l195) ? 1.0 : 0.0;
  }
  __shared__ double red[kMetricThreads / 32][kSums];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < kSums; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kSums) {
    double v = 0.0;
    for (int w = 0; w < kMetricThreads / 32; ++w) v += red[w][threadIdx.x];
    partial[((long)n * nslots + slot) * kSums + threadIdx.x] = v;
  }
}

__global__ void fold_kernel(const double* __restrict__ partial, int nslots, int width, double* __restrict__ sums) {
  const int n = blockIdx.x, k = threadIdx.x;
  if (k >= width) return;
  double v = 0.0;
  for (int s = 0; s < nslots; ++s) v += partial[((long)n * nslots + s) * width + k];
  sums[(long)n * width + k] = v;
}

// scale[n] from the folded sums: mode 0 'abs', 1 'log', 2 'inv' (metrics.py:283-318)
__global__ void scale_kernel(const double* __restrict__ sums, int n, int mode, float* __restrict__ scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* s = sums + (long)i * kSums;
  double r = 1.0;
  if (mode == 0) { if (s[12] > 0.0) r = s[13] / s[12]; }
  else if (mode == 1) { if (s[0] > 0.0) r = exp(-s[3] / s[0]); }
  else { if (s[14] > 0.0) r = 1.0 / (s[15] / s[14]); }
  scale[i] = (float)r;
}

__global__ void __launch_bounds__(kMetricThreads) epe_sums_kernel(const float* __restrict__ f1, const float* __restrict__ f2, long hw,
                                                                 double* __restrict__ partial) {
  const int n = blockIdx.y, slot = blockIdx.x, nslots = gridDim.x;
  const float* a = f1 + (long)n * 2 * hw;
  const float* b = f2 + (long)n * 2 * hw;
  double sum = 0.0, cnt = 0.0;
  for (long i = (long)slot * kMetricThreads + threadIdx.x; i < hw; i += (long)nslots * kMetricThreads) {
    const float dx = fsub(__ldg(a + i), __ldg(b + i)), dy = fsub(__ldg(a + hw + i), __ldg(b + hw + i));
    const float epe = sqrtf(fadd(fmul(dx, dx), fmul(dy, dy)));     // metrics.py:379-380
    if (isfinite(epe) && epe > 0.f) { sum += (double)epe; cnt += 1.0; }   // compute_valid_depth_mask(epe), metrics.py:382
  }
  __shared__ double red[kMetricThreads / 32][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  sum = warp_sum(sum); cnt = warp_sum(cnt);
  if (lane == 0) { red[warp][0] = sum; red[warp][1] = cnt; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double v = 0.0;
    for (int w = 0; w < kMetricThreads / 32; ++w) v += red[w][threadIdx.x];
    partial[((long)n * nslots + slot) * 2 + threadIdx.x] = v;
  }
}

int slots_for(int64_t hw) {
  int64_t s = (hw + 4 * kMetricThreads - 1) / (4 * kMetricThreads);
  return (int)(s < 1 ? 1 : (s > kMaxSlots ? kMaxSlots : s));
}

}  // namespace
}  // namespace demon

using namespace demon;

extern "C" {

int64_t demon_metric_workspace_bytes(int n, int64_t hw) {
  if (n <= 0 || hw <= 0) return 0;
  return (int64_t)n * slots_for(hw) * kSums * (int64_t)sizeof(double);
}

int demon_depth_error_sums_f32(const float* pred, const float* gt, int n, int64_t hw, int inverse_pred, int inverse_gt, const float* gt_div,
                               const float* pred_scale, double* sums, void* workspace, void* stream) {
  DEMON_REQUIRE(n >= 0 && hw >= 0 && n <= 65535, "depth_error_sums: bad size");
  if (n == 0) return DEMON_OK;
  DEMON_REQUIRE(sums && workspace && (hw == 0 || (pred && gt)), "depth_error_sums: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int nslots = slots_for(hw);
  depth_sums_kernel<<<dim3(nslots, n), kMetricThreads, 0, s>>>(pred, gt, (long)hw, inverse_pred != 0, inverse_gt != 0, gt_div, pred_scale,
                                                               static_cast<double*>(workspace));
  DEMON_LAUNCH_CHECK();
  fold_kernel<<<n, 32, 0, s>>>(static_cast<const double*>(workspace), nslots, kSums, sums);
  DEMON_LAUNCH_CHECK();
  return DEMON_OK;
}

int demon_depth_scale_factor(const double* sums, int n, int mode, float* scale, void* stream) {
  DEMON_REQUIRE(n >= 0 && mode >= 0 && mode <= 2, "depth_scale_factor: bad argument");
  if (n == 0) return DEMON_OK;
  DEMON_REQUIRE(sums && scale, "depth_scale_factor: null pointer");
  scale_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(sums, n, mode, scale);
  DEMON_LAUNCH_CHECK();
  return DEMON_OK;
}

int demon_flow_epe_sums_f32(const float* flow1, const float* flow2, int n, int64_t hw, double* sums, void* workspace, void* stream) {
  DEMON_REQUIRE(n >= 0 && hw >= 0 && n <= 65535, "flow_epe_sums: bad size");
  if (n == 0) return DEMON_OK;
  DEMON_REQUIRE(sums && workspace && (hw == 0 || (flow1 && flow2)), "flow_epe_sums: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int nslots = slots_for(hw);
  epe_sums_kernel<<<dim3(nslots, n), kMetricThreads, 0, s>>>(flow1, flow2, (long)hw, static_cast<double*>(workspace));
  DEMON_LAUNCH_CHECK();
  fold_kernel<<<n, 32, 0, s>>>(static_cast<const double*>(workspace), nslots, 2, sums);
  DEMON_LAUNCH_CHECK();
  return DEMON_OK;
}

}  // extern "C"
