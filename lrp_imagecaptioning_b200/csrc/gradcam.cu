// Grad-CAM heat-map and Guided-Grad-CAM scaling (reference: /root/reference/models/explainers.py:930-949, 1634-1653).
//   weights = mean_{xy} grads ; cam = sum_k weights_k F_k ; pyramid_expand(cam, upscale, sigma) ; relu ; / (max|cam| + 1e-6)
// skimage.transform.pyramid_expand (un-vendored, version unpinned -> parity unpinned) is restated as
// order-1 resize with skimage 'reflect' boundary (= scipy.ndimage 'mirror') followed by scipy.ndimage.gaussian_filter
// (mode 'reflect', truncate 4).  All HBM-bound: one block per word (and row), shared-memory staging, separable blur.
#include "../../include/lrpcap.h"
#include "common.cuh"
#include <cmath>
#include <vector>

namespace lrpcap {
namespace {

__global__ void __launch_bounds__(256)
cam_kernel(const float* __restrict__ F, const int* __restrict__ img_index, const float* __restrict__ grads,
           float* __restrict__ cam, int L, int D) {
  extern __shared__ float wts[];
  const int w = blockIdx.x;
  const float* g = grads + (size_t)w * L * D;
  for (int d = threadIdx.x; d < D; d += 256) {
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += g[(size_t)l * D + d];
    wts[d] = s / L;
  }
  __syncthreads();
  const float* f = F + (size_t)img_index[w] * L * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int l = warp; l < L; l += 8) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(wts[d], f[(size_t)l * D + d], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) cam[(size_t)w * L + l] = s;
  }
}

__device__ __forceinline__ int mirror(int i, int n) {   // d c b | a b c d | c b a
  if (n == 1) return 0;
  const int p = 2 * (n - 1);
  i = ((i % p) + p) % p;
  return i < n ? i : p - i;
}
__device__ __forceinline__ int reflect(int i, int n) {  // d c b a | a b c d | d c b a
  const int p = 2 * n;
  i = ((i % p) + p) % p;
  return i < n ? i : p - 1 - i;
}

// one block per (word, output row): bilinear resize of that row, then horizontal Gaussian
__global__ void __launch_bounds__(256)
resize_blurx_kernel(const float* __restrict__ cam, float* __restrict__ tmp, const float* __restrict__ gw, int fh, int hw,
                    int up, int radius) {
  extern __shared__ float row[];
  const int w = blockIdx.y, y = blockIdx.x;
  const float* c = cam + (size_t)w * fh * fh;
  const float iy = (y + 0.5f) / up - 0.5f;
  const int y0 = (int)floorf(iy);
  const float ty = iy - y0;
  const int ya = mirror(y0, fh), yb = mirror(y0 + 1, fh);
  for (int x = threadIdx.x; x < hw; x += 256) {
    const float ix = (x + 0.5f) / up - 0.5f;
    const int x0 = (int)floorf(ix);
    const float tx = ix - x0;
    const int xa = mirror(x0, fh), xb = mirror(x0 + 1, fh);
    const float top = c[ya * fh + xa] * (1.f - tx) + c[ya * fh + xb] * tx;
    const float bot = c[yb * fh + xa] * (1.f - tx) + c[yb * fh + xb] * tx;
    row[x] = top * (1.f - ty) + bot * ty;
  }
  __syncthreads();
  for (int x = threadIdx.x; x < hw; x += 256) {
    float s = 0.f;
    for (int k = -radius; k <= radius; ++k) s = fmaf(gw[k + radius], row[reflect(x + k, hw)], s);
    tmp[((size_t)w * hw + y) * hw + x] = s;
  }
}

// one block per (word, 32-column strip): vertical Gaussian + relu
__global__ void __launch_bounds__(256)
blury_kernel(const float* __restrict__ tmp, float* __restrict__ out, const float* __restrict__ gw, int hw, int radius) {
  extern __shared__ float col[];   // [hw][32]
  const int w = blockIdx.y, x0 = blockIdx.x * 32;
  for (int i = threadIdx.x; i < hw * 32; i += 256) {
    const int y = i / 32, x = x0 + (i & 31);
    col[i] = x < hw ? tmp[((size_t)w * hw + y) * hw + x] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hw * 32; i += 256) {
    const int y = i / 32, xl = i & 31;
    if (x0 + xl >= hw) continue;
    float s = 0.f;
    for (int k = -radius; k <= radius; ++k) s = fmaf(gw[k + radius], col[reflect(y + k, hw) * 32 + xl], s);
    out[((size_t)w * hw + y) * hw + x0 + xl] = fmaxf(s, 0.f);
  }
}

__global__ void __launch_bounds__(256) normalize_kernel(float* __restrict__ out, int n) {
  float* o = out + (size_t)blockIdx.x * n;
  __shared__ float red[256];
  float m = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) m = fmaxf(m, fabsf(o[i]));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  const float inv = 1.f / (red[0] + 1e-6f);
  for (int i = threadIdx.x; i < n; i += 256) o[i] *= inv;
}

__global__ void scale_maps_kernel(float* __restrict__ maps, const float* __restrict__ cam, size_t pixels) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels) return;
  const float c = cam[i];
  maps[3 * i] *= c;
  maps[3 * i + 1] *= c;
  maps[3 * i + 2] *= c;
}

}  // namespace
}  // namespace lrpcap

using namespace lrpcap;

extern "C" int lrpcap_gradcam(const float* d_features, const int* h_img_index, const float* d_grads, int n_words, int fh,
                              int D, int upscale, float sigma, float* d_cam, void* stream) {
  LRPCAP_REQUIRE(d_features && h_img_index && d_grads && d_cam, kErrInvalidArg, "gradcam: null argument");
  LRPCAP_REQUIRE(n_words > 0 && fh > 0 && D > 0 && upscale > 0 && sigma > 0.f, kErrShape, "gradcam: bad shape");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int L = fh * fh, hw = fh * upscale;
  const int radius = (int)(4.0f * sigma + 0.5f);
  LRPCAP_REQUIRE((size_t)hw * 32 * sizeof(float) <= 48 * 1024 && (size_t)D * sizeof(float) <= 48 * 1024, kErrShape,
                 "gradcam: map too large for the staging buffers");
  std::vector<float> gw(2 * radius + 1);
  double sum = 0.0;
  for (int k = -radius; k <= radius; ++k) sum += (gw[k + radius] = (float)std::exp(-0.5 * k * k / ((double)sigma * sigma)));
  for (float& v : gw) v = (float)(v / sum);
  int* d_idx = nullptr;
  float *d_gw = nullptr, *d_small = nullptr, *d_tmp = nullptr;
  int st = kOk;
  auto run = [&]() -> int {
    LRPCAP_CUDA(cudaMalloc(&d_idx, (size_t)n_words * sizeof(int)));
    LRPCAP_CUDA(cudaMalloc(&d_gw, gw.size() * sizeof(float)));
    LRPCAP_CUDA(cudaMalloc(&d_small, (size_t)n_words * L * sizeof(float)));
    LRPCAP_CUDA(cudaMalloc(&d_tmp, (size_t)n_words * hw * hw * sizeof(float)));
    LRPCAP_CUDA(cudaMemcpyAsync(d_idx, h_img_index, (size_t)n_words * sizeof(int), cudaMemcpyHostToDevice, s));
    LRPCAP_CUDA(cudaMemcpyAsync(d_gw, gw.data(), gw.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    cam_kernel<<<n_words, 256, D * sizeof(float), s>>>(d_features, d_idx, d_grads, d_small, L, D);
    resize_blurx_kernel<<<dim3(hw, n_words), 256, hw * sizeof(float), s>>>(d_small, d_tmp, d_gw, fh, hw, upscale, radius);
    blury_kernel<<<dim3((hw + 31) / 32, n_words), 256, (size_t)hw * 32 * sizeof(float), s>>>(d_tmp, d_cam, d_gw, hw, radius);
    normalize_kernel<<<n_words, 256, 0, s>>>(d_cam, hw * hw);
    LRPCAP_CUDA(cudaGetLastError());
    LRPCAP_CUDA(cudaStreamSynchronize(s));
    return kOk;
  };
  st = run();
  cudaFree(d_idx); cudaFree(d_gw); cudaFree(d_small); cudaFree(d_tmp);
  return st;
}

extern "C" int lrpcap_scale_maps(float* d_maps, const float* d_cam, int n_words, int hw, void* stream) {
  LRPCAP_REQUIRE(d_maps && d_cam && n_words > 0 && hw > 0, kErrInvalidArg, "scale_maps: bad argument");
  const size_t pixels = (size_t)n_words * hw * hw;
  scale_maps_kernel<<<(unsigned)((pixels + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_maps, d_cam, pixels);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}
