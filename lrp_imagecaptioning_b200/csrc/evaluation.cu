// Explanation-evaluation reductions on the device: the consumers of the pixel maps in the
// reference are HBM-bound reductions over [words, hw, hw, 3] tensors; running them next to the maps avoids shipping
// 602 KB per word to the host.
//   heat map     exaimin_word.py:96-103, :127-128 (mean over channels after postprocess(), project = x / max|x|)
//                evaluate_bbox.py:79-84          (negated, relu, mean over channels, project with the (x+1)/2 shift)
//   pooling      exaimin_word.py:64-77           (16x16 max / average pooling of the channel mean, then project)
//   correctness  evaluate_bbox.py:191-208        (relevance mass inside a bounding box / total mass, per threshold)
#include "../../include/lrpcap.h"
#include "common.cuh"

namespace lrpcap {
namespace {

constexpr int kEvalThreads = 512;
constexpr int kMaxThresholds = 16;

__device__ __forceinline__ float channel_value(const float* m, int p, int mode) {
  // postprocess(..., 'BGRtoRGB') reverses the channel axis before the mean: the float32 sum runs c2, c1, c0
  const float a = m[3 * p + 2], b = m[3 * p + 1], c = m[3 * p];
  if (mode == 0) return ((a + b) + c) / 3.0f;
  const float s = mode == 1 ? -1.f : 1.f;
  return ((fmaxf(s * a, 0.f) + fmaxf(s * b, 0.f)) + fmaxf(s * c, 0.f)) / 3.0f;
}

__device__ __forceinline__ float block_max(float v, float* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int st = kEvalThreads / 2; st > 0; st >>= 1) {
    if (threadIdx.x < st) sh[threadIdx.x] = fmaxf(sh[threadIdx.x], sh[threadIdx.x + st]);
    __syncthreads();
  }
  const float r = sh[0];
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_sum(double v, double* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int st = kEvalThreads / 2; st > 0; st >>= 1) {
    if (threadIdx.x < st) sh[threadIdx.x] += sh[threadIdx.x + st];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// One block per word. window == 1: heat map at full resolution; window > 1: window x window max (type 0) or average
// (type 1) pooling of the channel value first (the reference pools before it projects).
__global__ void __launch_bounds__(kEvalThreads)
heatmap_kernel(const float* __restrict__ maps, int hw, int mode, int shift_negative, int window, int pool_type,
               float* __restrict__ out, float* __restrict__ means) {
  __shared__ float shf[kEvalThreads];
  __shared__ double shd[kEvalThreads];
  const int ow = hw / window;
  const float* m = maps + (size_t)blockIdx.x * hw * hw * 3;
  float* o = out + (size_t)blockIdx.x * ow * ow;
  float amax = 0.f, vmin = 0.f;
  for (int q = threadIdx.x; q < ow * ow; q += kEvalThreads) {
    float v;
    if (window == 1) {
      v = channel_value(m, q, mode);
    } else {
      const int oy = q / ow, ox = q - oy * ow;
      float mx = -INFINITY, sum = 0.f;
      for (int dy = 0; dy < window; ++dy)
        for (int dx = 0; dx < window; ++dx) {
          const float c = channel_value(m, (oy * window + dy) * hw + ox * window + dx, mode);
          mx = fmaxf(mx, c);
          sum += c;
        }
      v = pool_type == 0 ? mx : sum / (float)(window * window);
    }
    o[q] = v;
    amax = fmaxf(amax, fabsf(v));
    vmin = fminf(vmin, v);
  }
  amax = block_max(amax, shf);
  const float neg = block_max(-vmin, shf);   // > 0 iff some value is negative
  double sum = 0.0;
  for (int q = threadIdx.x; q < ow * ow; q += kEvalThreads) {
    float v = 0.f;
    if (amax != 0.f) {
      v = o[q] / amax;
      if (shift_negative && neg > 0.f) v = (v + 1.f) / 2.f;
    }
    o[q] = v;
    sum += v;
  }
  sum = block_sum(sum, shd);
  if (means && threadIdx.x == 0) means[blockIdx.x] = (float)(sum / (double)(ow * ow));
}

struct Thresholds {
  float t[kMaxThresholds];
};

// One block per box: ratio[box, j] = sum_{p in box, hm > t_j} hm / sum_{hm > t_j} hm, capped at 1, 0 when empty.
__global__ void __launch_bounds__(kEvalThreads)
bbox_kernel(const float* __restrict__ hm, int hw, const int* __restrict__ boxes, Thresholds th, int n_th,
            float* __restrict__ ratio) {
  __shared__ double shd[kEvalThreads];
  const int* b = boxes + 5 * blockIdx.x;
  const float* m = hm + (size_t)b[0] * hw * hw;
  const int x0 = b[1], y0 = b[2], x1 = b[3], y1 = b[4];
  double tot[kMaxThresholds], cor[kMaxThresholds];
#pragma unroll
  for (int j = 0; j < kMaxThresholds; ++j) tot[j] = cor[j] = 0.0;
  for (int p = threadIdx.x; p < hw * hw; p += kEvalThreads) {
    const float v = m[p];
    const int y = p / hw, x = p - y * hw;
    const bool in = (y >= y0 && y < y1 && x >= x0 && x < x1);
#pragma unroll
    for (int j = 0; j < kMaxThresholds; ++j)
      if (j < n_th && v > th.t[j]) {
        tot[j] += v;
        if (in) cor[j] += v;
      }
  }
#pragma unroll
  for (int j = 0; j < kMaxThresholds; ++j) {
    if (j >= n_th) break;   // uniform
    const double t = block_sum(tot[j], shd), c = block_sum(cor[j], shd);
    if (threadIdx.x == 0) {
      double r = t == 0.0 ? 0.0 : c / t;
      if (r > 1.0) r = 1.0;
      ratio[(size_t)blockIdx.x * n_th + j] = (float)r;
    }
  }
}

}  // namespace
}  // namespace lrpcap

using namespace lrpcap;

extern "C" int lrpcap_heatmaps(const float* d_maps, int n_words, int hw, int mode, int shift_negative, int window,
                               int pool_type, float* d_out, float* h_means, void* stream) {
  LRPCAP_REQUIRE(d_maps && d_out && n_words > 0 && hw > 0, kErrInvalidArg, "heatmaps: bad argument");
  LRPCAP_REQUIRE(mode >= 0 && mode <= 2, kErrInvalidArg, "heatmaps: mode must be 0 (mean), 1 (negative part) or 2 (positive part)");
  LRPCAP_REQUIRE(window >= 1 && hw % window == 0 && (pool_type == 0 || pool_type == 1), kErrShape,
                 "heatmaps: window %d must divide the map size %d; pool_type 0 (max) or 1 (average)", window, hw);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* d_means = nullptr;
  if (h_means) LRPCAP_CUDA(cudaMalloc(&d_means, (size_t)n_words * sizeof(float)));
  heatmap_kernel<<<n_words, kEvalThreads, 0, s>>>(d_maps, hw, mode, shift_negative, window, pool_type, d_out, d_means);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && h_means) {
    e = cudaMemcpyAsync(h_means, d_means, (size_t)n_words * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  }
  if (d_means) cudaFree(d_means);
  LRPCAP_REQUIRE(e == cudaSuccess, kErrCuda, "heatmaps: %s", cudaGetErrorString(e));
  return kOk;
}

extern "C" int lrpcap_bbox_correctness(const float* d_heatmaps, int n_maps, int hw, const int* h_boxes, int n_boxes,
                                       const float* h_thresholds, int n_thresholds, float* h_ratio, void* stream) {
  LRPCAP_REQUIRE(d_heatmaps && h_boxes && h_thresholds && h_ratio && n_maps > 0 && n_boxes > 0 && hw > 0, kErrInvalidArg,
                 "bbox_correctness: bad argument");
  LRPCAP_REQUIRE(n_thresholds >= 1 && n_thresholds <= kMaxThresholds, kErrInvalidArg,
                 "bbox_correctness: 1..%d thresholds", kMaxThresholds);
  for (int i = 0; i < n_boxes; ++i)
    LRPCAP_REQUIRE(h_boxes[5 * i] >= 0 && h_boxes[5 * i] < n_maps, kErrInvalidArg,
                   "bbox_correctness: box %d refers to map %d outside [0,%d)", i, h_boxes[5 * i], n_maps);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  Thresholds th;
  for (int j = 0; j < kMaxThresholds; ++j) th.t[j] = j < n_thresholds ? h_thresholds[j] : 0.f;
  int* d_boxes = nullptr;
  float* d_ratio = nullptr;
  cudaError_t e = cudaMalloc(&d_boxes, (size_t)n_boxes * 5 * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&d_ratio, (size_t)n_boxes * n_thresholds * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_boxes, h_boxes, (size_t)n_boxes * 5 * sizeof(int), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) {
    bbox_kernel<<<n_boxes, kEvalThreads, 0, s>>>(d_heatmaps, hw, d_boxes, th, n_thresholds, d_ratio);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(h_ratio, d_ratio, (size_t)n_boxes * n_thresholds * sizeof(float), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (d_boxes) cudaFree(d_boxes);
  if (d_ratio) cudaFree(d_ratio);
  LRPCAP_REQUIRE(e == cudaSuccess, kErrCuda, "bbox_correctness: %s", cudaGetErrorString(e));
  return kOk;
}
