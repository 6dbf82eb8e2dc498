// Shared helpers for the lrpcap CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>

namespace lrpcap {

// ---- status codes returned across the C ABI (include/lrpcap.h) ----
enum : int {
  kOk = 0,
  kErrInvalidArg = -1,
  kErrShape = -2,
  kErrCuda = -3,
  kErrUnsupported = -4,
  kErrState = -5,
};

void set_last_error(const char* fmt, ...);
const char* get_last_error();

#define LRPCAP_CUDA(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::lrpcap::set_last_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,          \
                               cudaGetErrorString(_e));                                    \
      return ::lrpcap::kErrCuda;                                                           \
    }                                                                                      \
  } while (0)

#define LRPCAP_TRY(expr)                  \
  do {                                    \
    int _s = (expr);                      \
    if (_s != ::lrpcap::kOk) return _s;   \
  } while (0)

#define LRPCAP_REQUIRE(cond, code, ...)          \
  do {                                           \
    if (!(cond)) {                               \
      ::lrpcap::set_last_error(__VA_ARGS__);     \
      return (code);                             \
    }                                            \
  } while (0)

// ---- split-bf16 storage: a float tensor of n elements is kept as two bf16 planes,
//      hi[n] followed by lo[n], with v ~= float(hi) + float(lo) (16 mantissa bits).
//      Same 4 bytes/element as fp32; both planes are tcgen05 kind::f16 operands. ----
struct SplitPtr {
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
};
__host__ __device__ inline SplitPtr split_ptr(void* base, size_t n_elems) {
  SplitPtr p;
  p.hi = reinterpret_cast<__nv_bfloat16*>(base);
  p.lo = p.hi + n_elems;
  return p;
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ float join_bf16(__nv_bfloat16 hi, __nv_bfloat16 lo) {
  return __bfloat162float(hi) + __bfloat162float(lo);
}

// pack two floats' hi parts / lo parts into 32-bit words (element 0 in the low half)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  __nv_bfloat16 ah, al, bh, bl;
  split_bf16(a, ah, al);
  split_bf16(b, bh, bl);
  hi2 = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
  lo2 = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
}

// ---- split-fp16 storage (forward activations / weights, optional): v ~= float(hi) + float(lo) with two IEEE half planes,
//      22 mantissa bits inside the half range (|v| < 65504; lo underflows gradually below |v| ~ 0.1).  Plane code 4. ----
constexpr int kPlanesF16x2 = 4;
constexpr int kPlanesH1x2 = 5;    // two-product backward: A one fp16 plane, B two fp16 planes
constexpr int kPlanesH1F8 = 6;    // fp16 + fp8 backward: A = fp16 plane + E4M3 [top bits | residual] plane, B = fp16 hi plane +
                                  // E4M3 [low part | high part] plane: one kind::f16 and one double-length kind::f8f6f4 product
__device__ __forceinline__ void split2h(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  const __half ah = __float2half_rn(a), bh = __float2half_rn(b);
  const __half al = __float2half_rn(a - __half2float(ah)), bl = __float2half_rn(b - __half2float(bh));
  hi2 = (uint32_t)__half_as_ushort(ah) | ((uint32_t)__half_as_ushort(bh) << 16);
  lo2 = (uint32_t)__half_as_ushort(al) | ((uint32_t)__half_as_ushort(bl) << 16);
}
__device__ __forceinline__ float f16lo_to_float(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); }
__device__ __forceinline__ float f16hi_to_float(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }

__device__ __forceinline__ float bf16lo_to_float(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_to_float(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// ---- rule arithmetic shared by every kernel ----
// epsilon rule: z + sgn+(z)*eps, sgn+(0) = +1   (relevance_rule.py:131)
__device__ __forceinline__ float stab_eps(float z, float eps) { return z + (z >= 0.f ? eps : -eps); }
// SafeDivide denominator: z + [z==0]*1e-7      (layers.py:456-458)
__device__ __forceinline__ float safe_den(float z) { return z + (z == 0.f ? 1e-7f : 0.f); }

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Per-device launch bookkeeping (a process may drive several GPUs, several host threads may launch concurrently).
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline int device_sm_count() {
  static int sms[kMaxDevices] = {};
  const int dev = current_device();
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 1;
    sms[dev] = n;   // benign race: every writer stores the same value
  }
  return sms[dev];
}
// Raises a kernel's dynamic shared-memory limit once per device (and again if a larger size is asked for).
// `state` is a function-local static of the caller: one slot per device.
template <class Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, int (&state)[kMaxDevices]) {
  const int dev = current_device();
  if (state[dev] >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) state[dev] = bytes;
  return e;
}

}  // namespace lrpcap
