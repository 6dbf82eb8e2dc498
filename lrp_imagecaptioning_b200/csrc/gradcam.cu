// Grad-CAM heat-map and Guided-Grad-CAM scaling (reference: /root/reference/models/explainers.py:930-949, 1634-1653).
//   weights = mean_{xy} grads ; cam = sum_k weights_k F_k ; pyramid_expand(cam, upscale, sigma) ; relu ; / (max|cam| + 1e-6)
// skimage.transform.pyramid_expand (un-vendored, version unpinned -> parity unpinned) is restated as
// order-1 resize with skimage 'reflect' boundary (= scipy.ndimage 'mirror') followed by scipy.ndimage.gaussian_filter
// (mode 'reflect', truncate 4).  Both steps are linear and separable, so they are folded into one [hw, fh] matrix.
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

// pyramid_expand is linear and separable: out = A cam A^T with A [hw, fh] = (Gaussian, 'reflect') o (order-1 resize,
// 'mirror'), built once per call on the host in double precision (build_expand_matrix below).  One block per word:
// T = A cam (hw x fh, shared memory), out = relu(T A^T), block-wide max, scale by 1 / (max + 1e-6).
__global__ void __launch_bounds__(256)
expand_kernel(const float* __restrict__ cam, const float* __restrict__ A, float* __restrict__ out, int fh, int hw) {
  extern __shared__ float sm[];
  float* As = sm;                    // [hw][fh]
  float* Ts = sm + hw * fh;          // [hw][fh]
  float* cs = Ts + hw * fh;          // [fh][fh]
  __shared__ float red[256];
  const int w = blockIdx.x;
  for (int i = threadIdx.x; i < hw * fh; i += 256) As[i] = A[i];
  for (int i = threadIdx.x; i < fh * fh; i += 256) cs[i] = cam[(size_t)w * fh * fh + i];
  __syncthreads();
  for (int i = threadIdx.x; i < hw * fh; i += 256) {
    const int y = i / fh, j = i - y * fh;
    float s = 0.f;
    for (int k = 0; k < fh; ++k) s = fmaf(As[y * fh + k], cs[k * fh + j], s);
    Ts[i] = s;
  }
  __syncthreads();
  float* o = out + (size_t)w * hw * hw;
  float m = 0.f;
  for (int i = threadIdx.x; i < hw * hw; i += 256) {
    const int y = i / hw, x = i - y * hw;
    float s = 0.f;
    for (int j = 0; j < fh; ++j) s = fmaf(Ts[y * fh + j], As[x * fh + j], s);
    s = fmaxf(s, 0.f);
    o[i] = s;
    m = fmaxf(m, s);
  }
  red[threadIdx.x] = m;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + st]);
    __syncthreads();
  }
  const float inv = 1.f / (red[0] + 1e-6f);
  for (int i = threadIdx.x; i < hw * hw; i += 256) o[i] *= inv;   // each thread rescales what it wrote
}

inline int mirror_h(int i, int n) {   // d c b | a b c d | c b a   (skimage 'reflect' = ndimage 'mirror')
  if (n == 1) return 0;
  const int p = 2 * (n - 1);
  i = ((i % p) + p) % p;
  return i < n ? i : p - i;
}
inline int reflect_h(int i, int n) {  // d c b a | a b c d | d c b a   (ndimage 'reflect')
  const int p = 2 * n;
  i = ((i % p) + p) % p;
  return i < n ? i : p - 1 - i;
}

// A[x, j]: weight of source sample j in output sample x of  gaussian_filter1d(resize1d(.))
std::vector<float> build_expand_matrix(int fh, int up, float sigma) {
  const int hw = fh * up, radius = (int)(4.0f * sigma + 0.5f);
  std::vector<double> gw(2 * radius + 1);
  double sum = 0.0;
  for (int k = -radius; k <= radius; ++k) sum += (gw[k + radius] = std::exp(-0.5 * k * k / ((double)sigma * sigma)));
  for (double& v : gw) v /= sum;
  std::vector<double> R((size_t)hw * fh, 0.0);
  for (int x = 0; x < hw; ++x) {
    const double ix = (x + 0.5) / up - 0.5;
    const int x0 = (int)std::floor(ix);
    const double t = ix - x0;
    R[(size_t)x * fh + mirror_h(x0, fh)] += 1.0 - t;
    R[(size_t)x * fh + mirror_h(x0 + 1, fh)] += t;
  }
  std::vector<float> A((size_t)hw * fh);
  for (int x = 0; x < hw; ++x)
    for (int j = 0; j < fh; ++j) {
      double s = 0.0;
      for (int k = -radius; k <= radius; ++k) s += gw[k + radius] * R[(size_t)reflect_h(x + k, hw) * fh + j];
      A[(size_t)x * fh + j] = (float)s;
    }
  return A;
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
  const size_t smem = ((size_t)2 * hw * fh + (size_t)fh * fh) * sizeof(float);
  LRPCAP_REQUIRE(smem <= 200 * 1024 && (size_t)D * sizeof(float) <= 48 * 1024, kErrShape,
                 "gradcam: map too large for the staging buffers");
  const std::vector<float> A = build_expand_matrix(fh, upscale, sigma);
  int* d_idx = nullptr;
  float *d_A = nullptr, *d_small = nullptr;
  int st = kOk;
  auto run = [&]() -> int {
    LRPCAP_CUDA(cudaMalloc(&d_idx, (size_t)n_words * sizeof(int)));
    LRPCAP_CUDA(cudaMalloc(&d_A, A.size() * sizeof(float)));
    LRPCAP_CUDA(cudaMalloc(&d_small, (size_t)n_words * L * sizeof(float)));
    LRPCAP_CUDA(cudaMemcpyAsync(d_idx, h_img_index, (size_t)n_words * sizeof(int), cudaMemcpyHostToDevice, s));
    LRPCAP_CUDA(cudaMemcpyAsync(d_A, A.data(), A.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    cam_kernel<<<n_words, 256, D * sizeof(float), s>>>(d_features, d_idx, d_grads, d_small, L, D);
    static int smem_state[kMaxDevices] = {};
    LRPCAP_CUDA(ensure_dynamic_smem(expand_kernel, (int)smem, smem_state));
    expand_kernel<<<n_words, 256, smem, s>>>(d_small, d_A, d_cam, fh, hw);
    LRPCAP_CUDA(cudaGetLastError());
    LRPCAP_CUDA(cudaStreamSynchronize(s));
    return kOk;
  };
  st = run();
  cudaFree(d_idx); cudaFree(d_A); cudaFree(d_small);
  return st;
}

extern "C" int lrpcap_scale_maps(float* d_maps, const float* d_cam, int n_words, int hw, void* stream) {
  LRPCAP_REQUIRE(d_maps && d_cam && n_words > 0 && hw > 0, kErrInvalidArg, "scale_maps: bad argument");
  const size_t pixels = (size_t)n_words * hw * hw;
  scale_maps_kernel<<<(unsigned)((pixels + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_maps, d_cam, pixels);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}
