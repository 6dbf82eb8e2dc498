// LRP-inference word weights (reference: models/model.py:1641-1691, :2013-2062).
// Per explained word: hp = channel-mean of the pixel relevance map (the BGR->RGB flip of postprocess() does not change a
// mean over channels), hp /= max|hp| (`project`), score = mean(hp) | mean(relu(hp)).  One block per word, HBM-bound.
#include "../../include/lrpcap.h"
#include "common.cuh"

namespace lrpcap {
namespace {

__global__ void __launch_bounds__(512)
lrp_score_kernel(const float* __restrict__ maps, int pixels, int mode, float* __restrict__ scores) {
  const float* m = maps + (size_t)blockIdx.x * pixels * 3;
  double sum = 0.0, psum = 0.0;
  float amax = 0.f;
  for (int p = threadIdx.x; p < pixels; p += 512) {
    const float hp = (m[3 * p] + m[3 * p + 1] + m[3 * p + 2]) / 3.0f;
    sum += hp;
    psum += fmaxf(hp, 0.f);
    amax = fmaxf(amax, fabsf(hp));
  }
  __shared__ double s1[512], s2[512];
  __shared__ float s3[512];
  s1[threadIdx.x] = sum; s2[threadIdx.x] = psum; s3[threadIdx.x] = amax;
  __syncthreads();
  for (int st = 256; st > 0; st >>= 1) {
    if (threadIdx.x < st) {
      s1[threadIdx.x] += s1[threadIdx.x + st];
      s2[threadIdx.x] += s2[threadIdx.x + st];
      s3[threadIdx.x] = fmaxf(s3[threadIdx.x], s3[threadIdx.x + st]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double a = s3[0];
    scores[blockIdx.x] = a == 0.0 ? 0.f : (float)(((mode == 0) ? s1[0] : s2[0]) / a / pixels);
  }
}

}  // namespace
}  // namespace lrpcap

using namespace lrpcap;

extern "C" int lrpcap_lrp_inference_scores(const float* d_maps, int n_words, int hw, int mode, float* h_scores, void* stream) {
  LRPCAP_REQUIRE(d_maps && h_scores && n_words > 0 && hw > 0, kErrInvalidArg, "lrp_inference_scores: bad argument");
  // reference: NotImplementedError("the lrp inference mode is not available") (model.py:1685-1686)
  LRPCAP_REQUIRE(mode == 0 || mode == 1, kErrUnsupported, "lrp_inference_scores: mode must be 0 ('mean') or 1 ('pos_mean')");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* d = nullptr;
  LRPCAP_CUDA(cudaMalloc(&d, (size_t)n_words * sizeof(float)));
  lrp_score_kernel<<<n_words, 512, 0, s>>>(d_maps, hw * hw, mode, d);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_scores, d, (size_t)n_words * sizeof(float), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d);
  LRPCAP_REQUIRE(e == cudaSuccess, kErrCuda, "lrp_inference_scores: %s", cudaGetErrorString(e));
  return kOk;
}
