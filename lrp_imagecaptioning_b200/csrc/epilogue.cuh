// Fused rule arithmetic applied to a run of NV consecutive output channels of one pixel.
// Shared by the tcgen05 kernel (NV = 16, split-bf16 storage) and the fp32 SIMT kernel (NV = 4, fp32 storage),
// so both precisions implement exactly the same rules:
//   EpsilonRule / ZRule          innvestigate/analyzer/relevance_based/relevance_rule.py:74-144
//   AlphaBetaRule (z+ forward)   relevance_rule.py:216-322
//   max-pool routing             relevance_analyzer.py:459-480 (fused as a 2x2 up-sampling store with a
//                                pre-masked per-image multiplier)
//   Gradient / GuidedBackprop    analyzer/gradient_based.py:101-172, 228-265
#pragma once
#include "tc_conv.cuh"

namespace lrpcap {

struct EpiDev {
  const float* bias;
  void* out;          // activation (forward) or message (backward) tensor, storage-typed
  size_t out_elems;
  float* out_f32;
  float* G;
  float* Mseed;
  int gmode;
  int rule_bias;      // 0: the rule's z excludes the bias (IgnoreBias variants); the true forward always adds it
  float eps;
  const void* x_act;  // FWD_ZACT: true activation, storage-typed
  size_t x_elems;
  const int* img_index;
  const float* Gin;
  const float* Gin2;
  const unsigned* Gidx;
  int up;
  int relu_acc;
  float acc_scale;    // forward: z = acc * acc_scale (+ bias); 2^-k when the weights were pre-scaled by 2^k (half planes)
  int* overflow;      // forward, half planes: set to 1 when an activation leaves the half range
  int g_up;           // layout of G written by the forward epilogues (see g_offset)
  int out_planar_f32;    // backward: message written as fp32 [items][NO][H][W] (the last message)
  // two-product backward (StoreH1 messages): one fp16 plane holding 2^kt[item] * s, kt chosen per item and layer so that
  // the plane's maximum sits near 2^kMsgTargetExp (fp16 keeps 11 bits over 2^-14 .. 2^16 only)
  const unsigned* mx_in;   // [items] float bits of max |stored value| of the incoming message
  unsigned* mx_out;        // [items] running maximum of the message being written (atomicMax on the float bits)
  const int* kt_in;        // [items] log2 of the incoming message's scale
  int* kt_out;             // [items] log2 of the outgoing message's scale (every tile of an item writes the same value)
  int target_exp;          // the predicted maximum of the outgoing plane is 2^target_exp
  size_t g_elems;          // LRPCAP_DEBUG_BOUNDS: floats behind G / Gin (one multiplier tensor), 0 = unknown
  size_t aux_elems;        // LRPCAP_DEBUG_BOUNDS: floats behind out_f32 / Mseed, 0 = unknown
};

// Debug build (python -m lrp_imagecaptioning_b200.build --debug -> liblrpcap_dbg.so, run by tools/run_bounds_check.py):
// every epilogue address is checked against the logical size of the tensor it belongs to; a violation prints the site
// and traps (compute-sanitizer is closed on this GPU pool).
#ifdef LRPCAP_DEBUG_BOUNDS
#define LRPCAP_BOUNDS(what, off, n, limit)                                                                              \
  do {                                                                                                                  \
    if ((limit) != 0 && (size_t)(off) + (size_t)(n) > (size_t)(limit)) {                                                \
      printf("lrpcap bounds: %s offset %llu + %d > %llu (block %d thread %d)\n", what, (unsigned long long)(off), (int)(n),  \
             (unsigned long long)(limit), (int)blockIdx.x, (int)threadIdx.x);                                           \
      __trap();                                                                                                         \
    }                                                                                                                   \
  } while (0)
#else
#define LRPCAP_BOUNDS(what, off, n, limit) ((void)0)
#endif

constexpr int kMsgTargetExp = 4;   // default: predicted maximum 2^4: 2^12 of head-room against growth, 2^-28 of the maximum resolved
// log2 of the factor that brings a plane whose maximum has float bits `mbits` to the target maximum 2^target
__device__ __forceinline__ int msg_rescale_exp(unsigned mbits, int target) {
  const int ex = (int)((mbits >> 23) & 0xffu);
  int d = ex ? target - (ex - 127) : 0;
  return d < -100 ? -100 : (d > 100 ? 100 : d);
}
__device__ __forceinline__ float pow2i(int d) { return __int_as_float((127 + d) << 23); }   // d in [-126, 127]

// Per-image multipliers G are only ever touched by epilogues (never by TMA), so they are stored in the order the
// backward epilogue reads them: 16-channel runs of consecutive accumulator pixels are contiguous,
//   [img][C/16][up*up sub-pixel][H/up][W/up][16],   up = 2 when a 2x2 max-pool follows the layer.
// A warp of the backward epilogue (one accumulator pixel per lane, 16 channels per access) then reads 64 B per lane
// from consecutive addresses instead of one 64 B piece per 4*C-byte pixel row: 4x fewer L1 wavefronts.
__host__ __device__ __forceinline__ size_t g_offset(int img, int y, int x, int n, int H, int W, int C, int up) {
  const int Hc = H / up, Wc = W / up;
  const int sub = (y % up) * up + (x % up);
  return (((((size_t)img * (C >> 4) + (n >> 4)) * (up * up) + sub) * Hc + y / up) * Wc + x / up) * 16 + (n & 15);
}

// ---- 256-bit global accesses (sm_100: LDG.256 / STG.256). The epilogue's accesses are one pixel per lane, i.e. one
// L1 wavefront per lane whatever the width, so doubling the width halves the LSU wavefronts per tile.
__device__ __forceinline__ void stg256(void* p, const uint32_t (&u)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]),
               "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
               : "memory");
}
__device__ __forceinline__ void ldg256_nc(const void* p, uint32_t (&u)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "l"(p));
}

// ---- storage policies ----
struct StoreSplit {
  static constexpr bool kScaled = false;
  template <int NV>
  static __device__ __forceinline__ void store(void* base, size_t elems, size_t off, const float (&v)[NV]) {
    static_assert(NV % 4 == 0, "split storage moves 4 or 8 elements per vector");
    __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(base);
    __nv_bfloat16* lo = hi + elems;
    if constexpr (NV % 16 == 0) {
#pragma unroll
      for (int j = 0; j < NV / 16; ++j) {
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split2(v[16 * j + 2 * i], v[16 * j + 2 * i + 1], h[i], l[i]);
        stg256(hi + off + 16 * j, h);
        stg256(lo + off + 16 * j, l);
      }
    } else if constexpr (NV % 8 == 0) {
#pragma unroll
      for (int j = 0; j < NV / 8; ++j) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split2(v[8 * j + 2 * i], v[8 * j + 2 * i + 1], h[i], l[i]);
        reinterpret_cast<uint4*>(hi + off)[j] = make_uint4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<uint4*>(lo + off)[j] = make_uint4(l[0], l[1], l[2], l[3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV / 4; ++j) {
        uint32_t h[2], l[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) split2(v[4 * j + 2 * i], v[4 * j + 2 * i + 1], h[i], l[i]);
        reinterpret_cast<uint2*>(hi + off)[j] = make_uint2(h[0], h[1]);
        reinterpret_cast<uint2*>(lo + off)[j] = make_uint2(l[0], l[1]);
      }
    }
  }
  template <int NV>
  static __device__ __forceinline__ void load(const void* base, size_t elems, size_t off, float (&v)[NV]) {
    static_assert(NV % 4 == 0, "split storage moves 4 or 8 elements per vector");
    const __nv_bfloat16* hi = reinterpret_cast<const __nv_bfloat16*>(base);
    const __nv_bfloat16* lo = hi + elems;
    if constexpr (NV % 8 == 0) {
#pragma unroll
      for (int j = 0; j < NV / 8; ++j) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi + off) + j);
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(lo + off) + j);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[8 * j + 2 * i] = bf16lo_to_float(aw[i]) + bf16lo_to_float(bw[i]);
          v[8 * j + 2 * i + 1] = bf16hi_to_float(aw[i]) + bf16hi_to_float(bw[i]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV / 4; ++j) {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(hi + off) + j);
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(lo + off) + j);
        v[4 * j + 0] = bf16lo_to_float(a.x) + bf16lo_to_float(b.x);
        v[4 * j + 1] = bf16hi_to_float(a.x) + bf16hi_to_float(b.x);
        v[4 * j + 2] = bf16lo_to_float(a.y) + bf16lo_to_float(b.y);
        v[4 * j + 3] = bf16hi_to_float(a.y) + bf16hi_to_float(b.y);
      }
    }
  }
};

// two IEEE half planes (common.cuh: split2h); same interface and access widths as StoreSplit
struct StoreSplitH {
  static constexpr bool kScaled = false;
  template <int NV>
  static __device__ __forceinline__ void store(void* base, size_t elems, size_t off, const float (&v)[NV]) {
    static_assert(NV % 4 == 0, "split storage moves 4, 8 or 16 elements per vector");
    __half* hi = reinterpret_cast<__half*>(base);
    __half* lo = hi + elems;
    if constexpr (NV % 16 == 0) {
#pragma unroll
      for (int j = 0; j < NV / 16; ++j) {
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split2h(v[16 * j + 2 * i], v[16 * j + 2 * i + 1], h[i], l[i]);
        stg256(hi + off + 16 * j, h);
        stg256(lo + off + 16 * j, l);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV / 4; ++j) {
        uint32_t h[2], l[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) split2h(v[4 * j + 2 * i], v[4 * j + 2 * i + 1], h[i], l[i]);
        reinterpret_cast<uint2*>(hi + off)[j] = make_uint2(h[0], h[1]);
        reinterpret_cast<uint2*>(lo + off)[j] = make_uint2(l[0], l[1]);
      }
    }
  }
  template <int NV>
  static __device__ __forceinline__ void load(const void* base, size_t elems, size_t off, float (&v)[NV]) {
    static_assert(NV % 4 == 0, "split storage moves 4 elements per vector");
    const __half* hi = reinterpret_cast<const __half*>(base);
    const __half* lo = hi + elems;
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
      const uint2 a = __ldg(reinterpret_cast<const uint2*>(hi + off) + j);
      const uint2 b = __ldg(reinterpret_cast<const uint2*>(lo + off) + j);
      v[4 * j + 0] = f16lo_to_float(a.x) + f16lo_to_float(b.x);
      v[4 * j + 1] = f16hi_to_float(a.x) + f16hi_to_float(b.x);
      v[4 * j + 2] = f16lo_to_float(a.y) + f16lo_to_float(b.y);
      v[4 * j + 3] = f16hi_to_float(a.y) + f16hi_to_float(b.y);
    }
  }
};

// one IEEE half plane, saturating (the scaled message of the two-product backward); `elems` unused
struct StoreH1 {
  static constexpr bool kScaled = true;
  template <int NV>
  static __device__ __forceinline__ void store(void* base, size_t, size_t off, const float (&v)[NV]) {
    static_assert(NV % 8 == 0, "single-plane storage moves 8 or 16 elements per vector");
    __half* p = reinterpret_cast<__half*>(base) + off;
    uint32_t u[NV / 2];
#pragma unroll
    for (int i = 0; i < NV / 2; ++i) {
      const float a = fminf(fmaxf(v[2 * i], -65504.f), 65504.f), b = fminf(fmaxf(v[2 * i + 1], -65504.f), 65504.f);
      const __half2 h = __floats2half2_rn(a, b);
      u[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    if constexpr (NV == 16) {
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = u[i];
      stg256(p, w);
    } else {
#pragma unroll
      for (int j = 0; j < NV / 8; ++j) reinterpret_cast<uint4*>(p)[j] = make_uint4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
    }
  }
  template <int NV>
  static __device__ __forceinline__ void load(const void* base, size_t, size_t off, float (&v)[NV]) {
    const __half* p = reinterpret_cast<const __half*>(base) + off;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __half2float(p[i]);
  }
};

// fp16 + fp8 message (kPlanesH1F8): the scaled value v is stored as a16 = half(v) in the fp16 plane, and in the byte plane
// (at base + 2 * elems bytes, 128 bytes per 64-channel block) as [e4m3(a16) (64 B) | e4m3((v - a16) * 2^kResShift) (64 B)]:
// the top four bits of the message (pairs with the weights' low part) and its rounding residual (pairs with the weights'
// high part).  a16 * w_hi runs as one kind::f16 product, [a8 | r8] * [w_lo8 ; w_8] as one double-length kind::f8f6f4
// product at twice the rate: two product-equivalents for ~15 bits (tools/sim_h1f8.py: 1e-5 per layer vs 2e-4 of the plain
// two-product mode and 4e-6 of the three-product mode).
constexpr int kResShift = 13;
struct StoreH1F8 {
  static constexpr bool kScaled = true;
  template <int NV>
  static __device__ __forceinline__ void store(void* base, size_t elems, size_t off, const float (&v)[NV]) {
    static_assert(NV == 16 || NV == 8, "fp16 + fp8 storage moves 8 or 16 elements per vector");
    __half* p16 = reinterpret_cast<__half*>(base) + off;
    uint8_t* p8 = reinterpret_cast<uint8_t*>(base) + elems * 2 + (off >> 6) * 128 + (off & 63);
    uint32_t u[NV / 2];
    uint16_t a8[NV / 2], r8[NV / 2];
#pragma unroll
    for (int i = 0; i < NV / 2; ++i) {
      const float a = fminf(fmaxf(v[2 * i], -65504.f), 65504.f), b = fminf(fmaxf(v[2 * i + 1], -65504.f), 65504.f);
      const __half2 h = __floats2half2_rn(a, b);
      u[i] = *reinterpret_cast<const uint32_t*>(&h);
      const float2 hf = __half22float2(h);
      a8[i] = __nv_cvt_float2_to_fp8x2(hf, __NV_SATFINITE, __NV_E4M3);
      r8[i] = __nv_cvt_float2_to_fp8x2(make_float2((a - hf.x) * (float)(1 << kResShift), (b - hf.y) * (float)(1 << kResShift)),
                                       __NV_SATFINITE, __NV_E4M3);
    }
    if constexpr (NV == 16) {
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = u[i];
      stg256(p16, w);
      *reinterpret_cast<uint4*>(p8) = make_uint4(a8[0] | ((uint32_t)a8[1] << 16), a8[2] | ((uint32_t)a8[3] << 16),
                                                 a8[4] | ((uint32_t)a8[5] << 16), a8[6] | ((uint32_t)a8[7] << 16));
      *reinterpret_cast<uint4*>(p8 + 64) = make_uint4(r8[0] | ((uint32_t)r8[1] << 16), r8[2] | ((uint32_t)r8[3] << 16),
                                                      r8[4] | ((uint32_t)r8[5] << 16), r8[6] | ((uint32_t)r8[7] << 16));
    } else {
      *reinterpret_cast<uint4*>(p16) = make_uint4(u[0], u[1], u[2], u[3]);
      *reinterpret_cast<uint2*>(p8) = make_uint2(a8[0] | ((uint32_t)a8[1] << 16), a8[2] | ((uint32_t)a8[3] << 16));
      *reinterpret_cast<uint2*>(p8 + 64) = make_uint2(r8[0] | ((uint32_t)r8[1] << 16), r8[2] | ((uint32_t)r8[3] << 16));
    }
  }
  template <int NV>
  static __device__ __forceinline__ void load(const void* base, size_t, size_t off, float (&v)[NV]) {
    const __half* p = reinterpret_cast<const __half*>(base) + off;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __half2float(p[i]);
  }
};

struct StoreSplit3 {
  static constexpr bool kScaled = false;
  template <int NV>
  static __device__ __forceinline__ void store(void* base, size_t elems, size_t off, const float (&v)[NV]) {
    static_assert(NV % 4 == 0, "split storage moves 4 elements per vector");
    __nv_bfloat16* p0 = reinterpret_cast<__nv_bfloat16*>(base);
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
      uint32_t w[3][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float a = v[4 * j + 2 * i], b = v[4 * j + 2 * i + 1];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
          w[p][i] = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
          a -= __bfloat162float(ah);
          b -= __bfloat162float(bh);
        }
      }
#pragma unroll
      for (int p = 0; p < 3; ++p) reinterpret_cast<uint2*>(p0 + (size_t)p * elems + off)[j] = make_uint2(w[p][0], w[p][1]);
    }
  }
  template <int NV>
  static __device__ __forceinline__ void load(const void* base, size_t elems, size_t off, float (&v)[NV]) {
    static_assert(NV % 4 == 0, "split storage moves 4 elements per vector");
    const __nv_bfloat16* p0 = reinterpret_cast<const __nv_bfloat16*>(base);
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
      const uint2 a = __ldg(reinterpret_cast<const uint2*>(p0 + off) + j);
      const uint2 b = __ldg(reinterpret_cast<const uint2*>(p0 + elems + off) + j);
      const uint2 c = __ldg(reinterpret_cast<const uint2*>(p0 + 2 * elems + off) + j);
      v[4 * j + 0] = bf16lo_to_float(a.x) + bf16lo_to_float(b.x) + bf16lo_to_float(c.x);
      v[4 * j + 1] = bf16hi_to_float(a.x) + bf16hi_to_float(b.x) + bf16hi_to_float(c.x);
      v[4 * j + 2] = bf16lo_to_float(a.y) + bf16lo_to_float(b.y) + bf16lo_to_float(c.y);
      v[4 * j + 3] = bf16hi_to_float(a.y) + bf16hi_to_float(b.y) + bf16hi_to_float(c.y);
    }
  }
};

struct StoreF32 {
  static constexpr bool kScaled = false;
  template <int NV>
  static __device__ __forceinline__ void store(void* base, size_t, size_t off, const float (&v)[NV]) {
    static_assert(NV % 4 == 0, "fp32 storage moves 4 elements per 16 B");
    if constexpr (NV % 8 == 0) {
#pragma unroll
      for (int j = 0; j < NV / 8; ++j) {
        uint32_t u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = __float_as_uint(v[8 * j + i]);
        stg256(reinterpret_cast<float*>(base) + off + 8 * j, u);
      }
    } else {
      float4* q = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
#pragma unroll
      for (int i = 0; i < NV / 4; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
  }
  template <int NV>
  static __device__ __forceinline__ void load(const void* base, size_t, size_t off, float (&v)[NV]) {
    if constexpr (NV % 8 == 0) {
#pragma unroll
      for (int j = 0; j < NV / 8; ++j) {
        uint32_t u[8];
        ldg256_nc(reinterpret_cast<const float*>(base) + off + 8 * j, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 * j + i] = __uint_as_float(u[i]);
      }
    } else {
      const float4* q = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
#pragma unroll
      for (int i = 0; i < NV / 4; ++i) {
        const float4 t = __ldg(q + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    }
  }
};

template <int NV>
__device__ __forceinline__ void load_f32(const float* p, float (&v)[NV]) { StoreF32::load<NV>(p, 0, 0, v); }
template <int NV>
__device__ __forceinline__ void store_f32(float* p, const float (&v)[NV]) { StoreF32::store<NV>(p, 0, 0, v); }

// Backward epilogue for one run of NV channels: s_prev = acc * G'[img] at the UP x UP pixels the (pooled) pixel routes to.
// All multiplier loads are issued before the first use so their latencies overlap (8 epilogue warps give little
// thread-level parallelism to hide them otherwise).
template <int NV, class ST>
__device__ __forceinline__ void epi_store_msg(const EpiDev& e, size_t item_pixels, int item, size_t pix, int NO, int n,
                                              const float (&o)[NV]) {
  if (!e.out_planar_f32) {
    LRPCAP_BOUNDS("message", ((size_t)item * item_pixels + pix) * NO + n, NV, e.out_elems);
    ST::template store<NV>(e.out, e.out_elems, ((size_t)item * item_pixels + pix) * NO + n, o);
  } else {
    LRPCAP_BOUNDS("planar message", ((size_t)item * NO + n + NV - 1) * item_pixels + pix, 1, e.out_elems);
    // last message: fp32, fully channel-planar [item][NO][pixels] -- what the 64 -> 3 transposed conv (fp32 FMA, reads a
    // 3x3 window per channel through TMA) consumes; adjacent lanes are adjacent pixels, so the stores coalesce
    float* dst = reinterpret_cast<float*>(e.out) + ((size_t)item * NO + n) * item_pixels + pix;
#pragma unroll
    for (int i = 0; i < NV; ++i) dst[(size_t)i * item_pixels] = o[i];
  }
}

// Multiplier of sub-pixel `sub` from the compact form: the window's only non-zero sits at the arg-max position.
template <int UP>
__device__ __forceinline__ float g_select(float g, unsigned idx, int k, int sub) {
  if (UP == 1) return g;
  return ((idx >> (2 * k)) & 3u) == (unsigned)sub ? g : 0.f;
}

template <int UP, int NV, class ST>
__device__ __forceinline__ void epi_bwd(const EpiDev& e, int H, int W, int Nout, int item, int y, int x, int n,
                                        const float (&v)[NV]) {
  const int img = __ldg(e.img_index + item);
  const int WW = W * UP;
  const size_t item_pixels = (size_t)H * UP * WW;
  const int NO = e.Gin2 ? 2 * Nout : Nout;
  // (H, W) is the accumulator grid (= the pooled grid when UP == 2): one 16-channel run per accumulator pixel
  const size_t goff = (((size_t)img * (Nout >> 4) + (n >> 4)) * H + y) * W + x;
  const unsigned idx = UP == 2 ? __ldg(e.Gidx + goff) : 0u;
  for (int pass = 0; pass < (e.Gin2 ? 2 : 1); ++pass) {   // pass 1: inhibitor branch (beta != 0) -> channels [Nout, 2 Nout)
    float gg[NV];
    load_f32<NV>((pass ? e.Gin2 : e.Gin) + goff * 16 + (n & 15), gg);
#pragma unroll
    for (int sub = 0; sub < UP * UP; ++sub) {
      const size_t pix = (size_t)(y * UP + sub / UP) * WW + (x * UP + sub % UP);
      float o[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) o[i] = v[i] * g_select<UP>(gg[i], idx, (n & 15) + i, sub);
      epi_store_msg<NV, ST>(e, item_pixels, item, pix, NO, pass * Nout + n, o);
    }
  }
}

// Backward epilogue of one accumulator row (pixel) over NCH 16-channel chunks, software-pipelined: the multiplier run
// (and, for up-sampling layers, the arg-max word) of chunk c+1 is loaded while chunk c is multiplied and stored, and the
// image index is read once.  Written for the tcgen05 kernels, whose 8 epilogue warps have too little thread-level
// parallelism to hide a load -> use latency per chunk (ncu: the epilogue, not the tensor pipe, paced the 64-channel layers).
// load_acc(c, v) must fetch chunk c of the accumulator and is called by every lane (tcgen05.ld is warp-collective);
// `valid` only guards global-memory traffic. Chunk c covers channels [n_first + c * n_step, +16).
// PIPE = false keeps one multiplier run in flight instead of two (the promoted path holds the whole accumulator row in
// registers and has none to spare).
template <int UP, int NCH, class ST, bool PIPE = true, class AccLoader>
__device__ __forceinline__ void epi_bwd_chunks(const EpiDev& e, int H, int W, int Nout, int item, int y, int x,
                                               int n_first, int n_step, bool valid, AccLoader&& load_acc) {
  constexpr int SUBS = UP * UP;
  constexpr bool SC = ST::kScaled;
  const int img = valid ? __ldg(e.img_index + item) : 0;
  const int WW = W * UP;
  const size_t item_pixels = (size_t)H * UP * WW;
  const int NO = e.Gin2 ? 2 * Nout : Nout;
  const size_t gplane = (size_t)H * W;                                   // accumulator pixels per 16-channel run plane
  const size_t gpix = ((size_t)img * (Nout >> 4)) * gplane + (size_t)y * W + x;
  float sc = 1.f;        // SC: accumulator -> stored message (undoes the weight pre-scale, moves to the new item scale)
  unsigned tmax = 0u;    // SC: float bits of the largest |stored value| this thread wrote
  if (SC) {
    const int kin = __ldg(e.kt_in + item);                              // per thread: `item` is in range even when !valid
    const int d = e.out_planar_f32 ? -kin : msg_rescale_exp(__ldg(e.mx_in + item), e.target_exp);
    sc = e.acc_scale * pow2i(d < -100 ? -100 : (d > 100 ? 100 : d));
    if (e.kt_out && valid) e.kt_out[item] = kin + d;   // every lane of the item writes the same value (a warp can span several items)
  }
  for (int pass = 0; pass < (e.Gin2 ? 2 : 1); ++pass) {
    const float* G = pass ? e.Gin2 : e.Gin;
    constexpr int NB = PIPE ? 2 : 1;
    float gg[NB][16];
    unsigned gi[NB] = {};
    if (PIPE && valid) {
      const size_t o0 = gpix + (size_t)(n_first >> 4) * gplane;
      LRPCAP_BOUNDS("multiplier", o0 * 16, 16, e.g_elems);
      load_f32<16>(G + o0 * 16, gg[0]);
      if (UP == 2) gi[0] = __ldg(e.Gidx + o0);
    }
    // Code size: fully unrolled, the up-sampling form is NCH x 4 sub-pixels x 16 channels of convert-and-store code (the
    // N = 256 kernel was 227 KB of SASS) that each warp runs once per tile, so every pass came through the instruction cache
    // cold (ncu: stall_no_inst on the epilogue's ALU instructions, tensor pipe 49 % on the three layers under a pool). The
    // pipelined form keeps two chunks per loop body (the double-buffered multiplier registers need a static index) and
    // loops over the sub-pixels; the promoted form (PIPE = false) is a plain loop whose loader hands out acc[0] and rotates.
    constexpr int kChunkUnroll = !PIPE ? 1 : (UP == 2 && NCH % 2 == 0) ? 2 : NCH;   // !PIPE: the loader rotates its registers
    constexpr int kSubUnroll = 1;   // (one iteration when UP == 1)
#pragma unroll kChunkUnroll
    for (int c = 0; c < NCH; ++c) {
      float v[16];
      __syncwarp();   // reconverge after the predicated global accesses: the accumulator load is .sync.aligned
      load_acc(c, v);
      if (e.relu_acc) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      const int n = n_first + c * n_step;
      if (PIPE) {
        if (c + 1 < NCH && valid) {
          const size_t o1 = gpix + (size_t)((n + n_step) >> 4) * gplane;
          LRPCAP_BOUNDS("multiplier", o1 * 16, 16, e.g_elems);
          load_f32<16>(G + o1 * 16, gg[(c + 1) % NB]);
          if (UP == 2) gi[(c + 1) % NB] = __ldg(e.Gidx + o1);
        }
      } else if (valid) {
        const size_t o1 = gpix + (size_t)(n >> 4) * gplane;
        LRPCAP_BOUNDS("multiplier", o1 * 16, 16, e.g_elems);
        load_f32<16>(G + o1 * 16, gg[0]);
        if (UP == 2) gi[0] = __ldg(e.Gidx + o1);
      }
      if (valid) {
#pragma unroll kSubUnroll
        for (int sub = 0; sub < SUBS; ++sub) {
          float o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = v[i] * g_select<UP>(gg[c % NB][i], gi[c % NB], i, sub);
          if (SC) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              o[i] *= sc;
              tmax = max(tmax, __float_as_uint(o[i]) & 0x7fffffffu);
            }
          }
          const size_t pix = (size_t)(y * UP + sub / UP) * WW + (x * UP + sub % UP);
          epi_store_msg<16, ST>(e, item_pixels, item, pix, NO, pass * Nout + n, o);
        }
      }
    }
  }
  if (SC) {
    __syncwarp();
    // one atomic per item and warp: the lanes of a warp can belong to several items (multi-item tiles of tc_conv.cu)
    const unsigned peers = __match_any_sync(0xffffffffu, item);
    tmax = __reduce_max_sync(peers, tmax);
    if (e.mx_out && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1) && tmax) atomicMax(e.mx_out + item, tmax);
  }
}

// Pulls the multiplier runs a backward epilogue thread is about to read into L2. Issued before the thread waits for its
// accumulator, so the DRAM latency of G overlaps the MMAs instead of serialising with every 16-column chunk.
__device__ __forceinline__ void epi_prefetch_bwd(const EpiDev& e, int H, int W, int Nout, int item, int y, int x, int n) {
  const int img = __ldg(e.img_index + item);
  const size_t goff = (((size_t)img * (Nout >> 4) + (n >> 4)) * H + y) * W + x;
  asm volatile("prefetch.global.L2 [%0];" ::"l"(e.Gin + goff * 16));
  if (e.Gin2) asm volatile("prefetch.global.L2 [%0];" ::"l"(e.Gin2 + goff * 16));
  if (e.Gidx) asm volatile("prefetch.global.L2 [%0];" ::"l"(e.Gidx + goff));
}

// v: accumulator values for channels [n, n+NV) of output pixel (item, y, x) of an H x W x Nout map.
template <int MODE, int NV, class ST>
__device__ __forceinline__ void epi_apply(const EpiDev& e, int H, int W, int Nout, int item, int y, int x, int n,
                                          float (&v)[NV]) {
  if (MODE == EPI_RAW) {
    store_f32<NV>(e.out_f32 + (((size_t)item * H + y) * W + x) * Nout + n, v);
  } else if (MODE == EPI_FWD_TRUE) {
    const size_t off = (((size_t)item * H + y) * W + x) * Nout + n;
    float z[NV], xo[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] *= e.acc_scale;
      z[i] = v[i] + (e.bias ? __ldg(e.bias + n + i) : 0.f);
      xo[i] = fmaxf(z[i], 0.f);
    }
    if (e.overflow) {
      float mx = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) mx = fmaxf(mx, xo[i]);
      if (!(mx < 32768.f)) atomicOr(e.overflow, 1);
    }
    LRPCAP_BOUNDS("activation", off, NV, e.out_elems);
    ST::template store<NV>(e.out, e.out_elems, off, xo);
    if (e.out_f32) {
      LRPCAP_BOUNDS("features", off, NV, e.aux_elems);
      store_f32<NV>(e.out_f32 + off, xo);
    }
    if (e.G || e.Mseed) {
      float gg[NV], mm[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float zr = e.rule_bias ? z[i] : v[i];
        if (e.gmode == G_EPS) {
          const float d = stab_eps(zr, e.eps);
          mm[i] = 1.f / d;
          gg[i] = xo[i] / d;
        } else if (e.gmode == G_Z) {
          const float d = safe_den(zr);
          mm[i] = 1.f / d;
          gg[i] = xo[i] / d;
        } else {
          mm[i] = z[i] > 0.f ? 1.f : 0.f;
          gg[i] = mm[i];
        }
      }
      if (e.G) {
        LRPCAP_BOUNDS("multiplier store", g_offset(item, y, x, n, H, W, Nout, e.g_up), NV, e.g_elems);
        store_f32<NV>(e.G + g_offset(item, y, x, n, H, W, Nout, e.g_up), gg);
      }
      if (e.Mseed) {
        LRPCAP_BOUNDS("seed multiplier", off, NV, e.aux_elems);
        store_f32<NV>(e.Mseed + off, mm);
      }
    }
  } else if (MODE == EPI_FWD_ZACT) {
    const size_t off = (((size_t)item * H + y) * W + x) * Nout + n;
    float xa[NV], gg[NV], mm[NV];
    LRPCAP_BOUNDS("activation read", off, NV, e.x_elems);
    ST::template load<NV>(e.x_act, e.x_elems, off, xa);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float d = safe_den(v[i] * e.acc_scale + ((e.bias && e.rule_bias) ? __ldg(e.bias + n + i) : 0.f));
      mm[i] = 1.f / d;
      gg[i] = xa[i] / d;
    }
    if (e.G) store_f32<NV>(e.G + g_offset(item, y, x, n, H, W, Nout, e.g_up), gg);
    if (e.Mseed) store_f32<NV>(e.Mseed + off, mm);
  } else {  // EPI_BWD
    if (e.relu_acc) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (e.up == 2)
      epi_bwd<2, NV, ST>(e, H, W, Nout, item, y, x, n, v);
    else
      epi_bwd<1, NV, ST>(e, H, W, Nout, item, y, x, n, v);
  }
}

// Fills the device-side epilogue struct from the host-side description; validates per mode.
inline int make_epi_dev(const EpiParams& p, EpiDev* e) {
  e->bias = p.bias;
  e->out_f32 = p.out_f32;
  e->G = p.G;
  e->Mseed = p.Mseed;
  e->gmode = p.gmode;
  e->rule_bias = p.rule_bias;
  e->eps = p.eps;
  e->x_act = p.x_act;
  e->x_elems = p.x_act_elems;
  e->img_index = p.img_index;
  e->Gin = p.Gin;
  e->Gin2 = p.Gin2;
  e->Gidx = p.Gidx;
  e->up = p.up;
  e->relu_acc = p.relu_acc;
  e->g_up = p.g_up;
  e->acc_scale = p.acc_scale;
  e->overflow = p.overflow;
  e->out_planar_f32 = p.out_planar_f32;
  e->mx_in = p.mx_in;
  e->mx_out = p.mx_out;
  e->kt_in = p.kt_in;
  e->kt_out = p.kt_out;
  e->target_exp = p.target_exp;
  e->g_elems = p.g_elems;
  e->aux_elems = p.aux_elems;
  e->out = nullptr;
  e->out_elems = 0;
  switch (p.mode) {
    case EPI_BWD:
      LRPCAP_REQUIRE(p.out_msg && p.Gin && p.img_index && (p.up == 1 || (p.up == 2 && p.Gidx)), kErrInvalidArg,
                     "conv: incomplete backward epilogue");
      e->out = p.out_msg;
      e->out_elems = p.out_msg_elems;
      break;
    case EPI_FWD_TRUE:
      LRPCAP_REQUIRE(p.out_act != nullptr, kErrInvalidArg, "conv: forward epilogue needs out_act");
      e->out = p.out_act;
      e->out_elems = p.out_act_elems;
      break;
    case EPI_FWD_ZACT:
      LRPCAP_REQUIRE(p.x_act != nullptr && (p.G || p.Mseed), kErrInvalidArg, "conv: zact epilogue incomplete");
      break;
    case EPI_RAW:
      LRPCAP_REQUIRE(p.out_f32 != nullptr, kErrInvalidArg, "conv: raw epilogue needs out_f32");
      break;
    default:
      set_last_error("conv: unknown epilogue mode %d", p.mode);
      return kErrInvalidArg;
  }
  return kOk;
}

}  // namespace lrpcap
