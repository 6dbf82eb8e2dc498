// HBM-bound helper kernels of the encoder relevance path: all coalesced, 16-byte vectorised.
#include "encoder_kernels.cuh"
#include "tc_ptx.cuh"
#include <cstdlib>
#include "epilogue.cuh"

namespace lrpcap {

namespace {

inline unsigned grid_for(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  return (unsigned)g;
}

// ------------------------------------------------------------------ weight preparation
__global__ void prep_weights_kernel(const float* __restrict__ w, float* __restrict__ out_f32,
                                    __nv_bfloat16* __restrict__ out_hi, int planes, int cin,
                                    int cout, int fmt, int sign, int taps, float scale) {
  const size_t total = (size_t)taps * cin * cout;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int tap, ci, co;
  if (fmt == WF_SIMT_FWD) {
    co = idx % cout; ci = (idx / cout) % cin; tap = idx / ((size_t)cout * cin);
  } else if (fmt == WF_SIMT_BWD) {
    ci = idx % cin; co = (idx / cin) % cout; tap = taps - 1 - (int)(idx / ((size_t)cout * cin));
  } else if (fmt == WF_TC_FWD) {
    ci = idx % cin; co = (idx / cin) % cout; tap = idx / ((size_t)cout * cin);
  } else {
    co = idx % cout; ci = (idx / cout) % cin; tap = taps - 1 - (int)(idx / ((size_t)cout * cin));
  }
  float v = w[((size_t)tap * cin + ci) * cout + co];
  if (sign == WS_PLUS) v = v >= 0.f ? v : 0.f;
  if (sign == WS_MINUS) v = v < 0.f ? v : 0.f;
  if (fmt == WF_SIMT_FWD || fmt == WF_SIMT_BWD) {
    out_f32[idx] = v;
  } else if (planes == kPlanesH1F8) {    // fp16 high plane [total] + byte plane [total / 64][128] = [e4m3(low part) | e4m3(2^-kResShift * w)]
    __half* oh = reinterpret_cast<__half*>(out_hi);
    uint8_t* o8 = reinterpret_cast<uint8_t*>(out_hi) + total * 2;
    v *= scale;
    const __half h = __float2half_rn(v);
    oh[idx] = h;
    const size_t b = (idx >> 6) * 128 + (idx & 63);      // the K axis (innermost, a multiple of 64) in 64-element blocks
    o8[b] = (uint8_t)__nv_cvt_float_to_fp8(v - __half2float(h), __NV_SATFINITE, __NV_E4M3);
    o8[b + 64] = (uint8_t)__nv_cvt_float_to_fp8(v * (1.f / (float)(1 << kResShift)), __NV_SATFINITE, __NV_E4M3);
  } else if (planes == kPlanesF16x2) {   // two half planes of scale * w (scale = 2^k keeps the low plane normal)
    __half* oh = reinterpret_cast<__half*>(out_hi);
    v *= scale;
    const __half h = __float2half_rn(v);
    oh[idx] = h;
    oh[total + idx] = __float2half_rn(v - __half2float(h));
  } else {
    for (int p = 0; p < planes; ++p) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      out_hi[(size_t)p * total + idx] = h;
      v -= __bfloat162float(h);
    }
  }
}

__global__ void prep_weights_dual_kernel(const float* __restrict__ w, float* __restrict__ out_f32,
                                         __nv_bfloat16* __restrict__ out_hi, int cin, int cout, int fmt, int sign_a,
                                         float scale_a, int sign_b, float scale_b, int half_planes) {
  const size_t total = (size_t)9 * cin * 2 * cout;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int tapf, ci, k;
  if (fmt == WF_SIMT_BWD) {
    ci = idx % cin; k = (idx / cin) % (2 * cout); tapf = idx / ((size_t)cin * 2 * cout);
  } else {
    k = idx % (2 * cout); ci = (idx / (2 * cout)) % cin; tapf = idx / ((size_t)cin * 2 * cout);
  }
  const int co = k % cout;
  const bool second = k >= cout;
  float v = w[((size_t)(8 - tapf) * cin + ci) * cout + co];
  const int sg = second ? sign_b : sign_a;
  if (sg == WS_PLUS) v = v >= 0.f ? v : 0.f;
  if (sg == WS_MINUS) v = v < 0.f ? v : 0.f;
  v *= second ? scale_b : scale_a;
  if (fmt == WF_SIMT_BWD) {
    out_f32[idx] = v;
  } else if (half_planes == 2) {   // fp16 + fp8 layout (see prep_weights_kernel)
    __half* oh = reinterpret_cast<__half*>(out_hi);
    uint8_t* o8 = reinterpret_cast<uint8_t*>(out_hi) + total * 2;
    const __half h = __float2half_rn(v);
    oh[idx] = h;
    const size_t b = (idx >> 6) * 128 + (idx & 63);
    o8[b] = (uint8_t)__nv_cvt_float_to_fp8(v - __half2float(h), __NV_SATFINITE, __NV_E4M3);
    o8[b + 64] = (uint8_t)__nv_cvt_float_to_fp8(v * (1.f / (float)(1 << kResShift)), __NV_SATFINITE, __NV_E4M3);
  } else if (half_planes) {
    __half* oh = reinterpret_cast<__half*>(out_hi);
    const __half h = __float2half_rn(v);
    oh[idx] = h;
    oh[total + idx] = __float2half_rn(v - __half2float(h));
  } else {
    for (int p = 0; p < 2; ++p) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      out_hi[(size_t)p * total + idx] = h;
      v -= __bfloat162float(h);
    }
  }
}

// ------------------------------------------------------------------ max-pool + arg-max masking
template <class ST>
__global__ void pool_mask_kernel(const void* __restrict__ act, size_t act_elems, void* pooled, size_t pooled_elems,
                                 const float* __restrict__ G, float* __restrict__ Gc, unsigned char* __restrict__ Gi,
                                 int items, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const size_t total = (size_t)items * Ho * Wo * C4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (idx % C4) * 4;
  size_t r = idx / C4;
  const int xo = r % Wo; r /= Wo;
  const int yo = r % Ho;
  const int item = r / Ho;
  float v[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
    ST::template load<4>(act, act_elems, (((size_t)item * H + (2 * yo + (p >> 1))) * W + (2 * xo + (p & 1))) * C + c, v[p]);
  float m[4];
  int am[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = v[0][i];
    am[i] = 0;
#pragma unroll
    for (int p = 1; p < 4; ++p)
      if (v[p][i] > m[i]) { m[i] = v[p][i]; am[i] = p; }   // strict '>' keeps the first maximum
  }
  if (pooled) ST::template store<4>(pooled, pooled_elems, (((size_t)item * Ho + yo) * Wo + xo) * C + c, m);
  if (G) {
    // compact multiplier of the pooled layer: the value at the arg-max position (every other entry of the window routes
    // nothing) + its 2-bit position; layout [item][C/16][Ho][Wo][16] / one byte per 4 channels (epilogue.cuh: g_select)
    float g[4][4], sel[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) load_f32<4>(G + g_offset(item, 2 * yo + (p >> 1), 2 * xo + (p & 1), c, H, W, C, 2), g[p]);
#pragma unroll
    for (int i = 0; i < 4; ++i) sel[i] = am[i] == 0 ? g[0][i] : am[i] == 1 ? g[1][i] : am[i] == 2 ? g[2][i] : g[3][i];
    const size_t run = (((size_t)item * (C >> 4) + (c >> 4)) * Ho + yo) * Wo + xo;
    store_f32<4>(Gc + run * 16 + (c & 15), sel);
    if (Gi) Gi[run * 4 + ((c & 15) >> 2)] = (unsigned char)(am[0] | (am[1] << 2) | (am[2] << 4) | (am[3] << 6));
  }
}

// ------------------------------------------------------------------ seed message
// Two-product backward: max |R * M| (and |R * M2|) per item, so that the seed message can be stored as one fp16 plane
// scaled by a power of two per item (epilogue.cuh: kMsgTargetExp). grid = (blocks per item, items).
__global__ void seed_max_kernel(const float* __restrict__ R, const float* __restrict__ M, const float* __restrict__ M2,
                                const int* __restrict__ img_index, unsigned* __restrict__ mx, size_t per_item, int relu) {
  const int item = blockIdx.y;
  const int img = __ldg(img_index + item);
  const float* r0 = R + (size_t)item * per_item;
  const float* m0 = M + (size_t)img * per_item;
  const float* m1 = M2 ? M2 + (size_t)img * per_item : nullptr;
  unsigned best = 0u;
  for (size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < per_item; e += (size_t)gridDim.x * blockDim.x * 4) {
    float r[4], m[4];
    load_f32<4>(r0 + e, r);
    load_f32<4>(m0 + e, m);
#pragma unroll
    for (int i = 0; i < 4; ++i) best = max(best, __float_as_uint((relu ? fmaxf(r[i], 0.f) : r[i]) * m[i]) & 0x7fffffffu);
    if (m1) {
      load_f32<4>(m1 + e, m);
#pragma unroll
      for (int i = 0; i < 4; ++i) best = max(best, __float_as_uint(r[i] * m[i]) & 0x7fffffffu);
    }
  }
  best = __reduce_max_sync(0xffffffffu, best);
  if ((threadIdx.x & 31) == 0 && best) atomicMax(mx + item, best);
}

// mx_true != null (ST = StoreH1): the message is stored as 2^k * value, k = msg_rescale_exp(mx_true[item]); the first
// thread of every item records k (kt_out) and the stored plane's maximum (mx_out) for the next layer's epilogue.
template <class ST>
__global__ void seed_kernel(const float* __restrict__ R, const float* __restrict__ M, const float* __restrict__ M2,
                            const int* __restrict__ img_index, void* msg, size_t msg_elems, int items, size_t per_item,
                            int C, int relu, const unsigned* __restrict__ mx_true, unsigned* __restrict__ mx_out,
                            int* __restrict__ kt_out, int target_exp) {
  constexpr int NV = ST::kScaled ? 8 : 4;
  const size_t totalv = (size_t)items * per_item / NV;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= totalv) return;
  const size_t e = idx * NV;
  const int item = e / per_item;
  const size_t in_item = e - (size_t)item * per_item;
  const int img = __ldg(img_index + item);
  float sc = 1.f;
  if (ST::kScaled) {
    const unsigned mb = __ldg(mx_true + item);
    const int k = msg_rescale_exp(mb, target_exp);
    sc = pow2i(k);
    if (in_item == 0) {
      kt_out[item] = k;
      mx_out[item] = __float_as_uint(__uint_as_float(mb) * sc);
    }
  }
  float r[NV], m[NV], o[NV];
  load_f32<NV>(R + e, r);
  load_f32<NV>(M + (size_t)img * per_item + in_item, m);
#pragma unroll
  for (int i = 0; i < NV; ++i) o[i] = (relu ? fmaxf(r[i], 0.f) : r[i]) * m[i] * sc;
  if (!M2) {
    ST::template store<NV>(msg, msg_elems, e, o);
    return;
  }
  const size_t pix = e / C;               // dual layout: [.., pixel, 2C] = [R*M | R*M2]
  const int c = (int)(e - pix * C);
  ST::template store<NV>(msg, msg_elems, pix * 2 * C + c, o);
  load_f32<NV>(M2 + (size_t)img * per_item + in_item, m);
#pragma unroll
  for (int i = 0; i < NV; ++i) o[i] = r[i] * m[i] * sc;
  ST::template store<NV>(msg, msg_elems, pix * 2 * C + C + c, o);
}

// ------------------------------------------------------------------ last transposed conv (C -> 3) + re-weighting
// fp32-FMA bound (1728 FMA per pixel; 12.85 MB in, 0.6 MB out per word at 224x224). The message arrives as fp32,
// channel-planar [item][C][H][W] (EpiParams::out_planar_f32 of the layer above), so one TMA box {40, 34, 4} IS the
// shared-memory operand: 4 channels of the 32 x 32 tile with its halo, OOB zero-fill = padding, no conversion pass,
// no registers holding staged data. One thread keeps two boxes in flight (double buffer) while all threads run the FMAs:
// a thread owns a 2 x 4 pixel block and per channel loads its 4 x 6 window (4 x (LDS.32, LDS.128, LDS.32)) and the channel's 27
// weights (7 broadcast LDS.128 from a shared-memory copy) for 216 FMAs, which keeps the LSU well below the FMA pipe.
// (History, measured: register-staged 16-channel chunks of the pixel-major split message were DRAM-bound on re-fetched
// lines -- 78 % DRAM, 20 % FMA; a TMA-staged bf16 variant still spent a fifth of its issue slots converting.)
constexpr int kLTY = 32, kLTX = 32;   // tile (rows x cols): 128 threads x (2 x 4) pixels
constexpr int kLThreads = 128;
constexpr int kLC = 4;                // channels per TMA box
constexpr int kLastMaxC = 128;        // 64 channels, or 2 x 64 for the dual (beta != 0) message
constexpr int kLPSX = kLTX + 8, kLPSY = kLTY + 2;   // halo box: columns x0-4 .. x0+35 (a TMA box must start 16 B aligned in
                                                    // its innermost dimension, so the left halo column drags 3 more along)
constexpr int kLBox = kLC * kLPSY * kLPSX;          // floats per box
constexpr int kLWPitch = 28;          // 27 weights per channel (tap-major, 3 colours), padded to 7 x float4

template <bool DUAL, int C>
__global__ void __launch_bounds__(kLThreads, 4)
last_dgrad_kernel(const __grid_constant__ CUtensorMap map, const float* __restrict__ Wa, const float* __restrict__ Wb,
                  const float* __restrict__ images, const int* __restrict__ img_index, float* __restrict__ out, int H,
                  int W, int tiles_x, int tiles_y, int mult) {
  using namespace tcptx;
  extern __shared__ __align__(128) uint8_t last_smem[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(last_smem) + 127) & ~uintptr_t(127));
  float* S = reinterpret_cast<float*>(sm);                  // [2][kLC][kLPSY][kLPSX]
  float* Wsa = S + 2 * kLBox;                               // [C][28]
  float* Wsb = Wsa + C * kLWPitch;                          // [C][28] (DUAL only)
  uint64_t* full = reinterpret_cast<uint64_t*>(Wsb + (DUAL ? C * kLWPitch : 0));   // [2]

  int bid = blockIdx.x;
  const int tiles = tiles_x * tiles_y;
  const int item = bid / tiles;
  bid -= item * tiles;
  const int y0 = (bid / tiles_x) * kLTY, x0 = (bid % tiles_x) * kLTX;
  const int tid = threadIdx.x;
  const int ty = (tid >> 3) * 2, tx = (tid & 7) * 4;
  constexpr int NCH = C / kLC;
  constexpr uint32_t kBoxBytes = (uint32_t)kLBox * sizeof(float);

  if (tid == 0) {
    prefetch_tmap(&map);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      mbar_expect_tx(&full[b], kBoxBytes);
      tma_load_3d(&map, S + b * kLBox, &full[b], x0 - 4, y0 - 1, item * C + b * kLC);
    }
  }
  for (int i = tid; i < 9 * C * 3; i += kLThreads) {   // weights [tap][C][3] -> shared [c][tap * 3 + colour]
    const int tap = i / (C * 3), rem = i - tap * C * 3, c = rem / 3, ci = rem - c * 3;
    Wsa[c * kLWPitch + tap * 3 + ci] = __ldg(Wa + i);
    if (DUAL) Wsb[c * kLWPitch + tap * 3 + ci] = __ldg(Wb + i);
  }

  float ca[2][4][3], cb[2][4][3];
#pragma unroll
  for (int py = 0; py < 2; ++py)
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) ca[py][px][ci] = cb[py][px][ci] = 0.f;
  __syncthreads();   // barriers initialised, weights in place

#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    const int b = ch & 1;
    mbar_wait(&full[b], (uint32_t)((ch >> 1) & 1));
    const float* Sb = S + b * kLBox;
#pragma unroll
    for (int k = 0; k < kLC; ++k) {
      float win[4][6];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float* row = Sb + (k * kLPSY + ty + r) * kLPSX + tx;     // box column j holds image column x0 - 4 + j
        const float4 a = *reinterpret_cast<const float4*>(row + 4);
        win[r][0] = row[3]; win[r][1] = a.x; win[r][2] = a.y; win[r][3] = a.z; win[r][4] = a.w; win[r][5] = row[8];
      }
      const int cc = ch * kLC + k;
      float wa[kLWPitch], wb[kLWPitch];
#pragma unroll
      for (int i = 0; i < kLWPitch / 4; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(Wsa + cc * kLWPitch + 4 * i);
        wa[4 * i] = t.x; wa[4 * i + 1] = t.y; wa[4 * i + 2] = t.z; wa[4 * i + 3] = t.w;
        if (DUAL) {
          const float4 u = *reinterpret_cast<const float4*>(Wsb + cc * kLWPitch + 4 * i);
          wb[4 * i] = u.x; wb[4 * i + 1] = u.y; wb[4 * i + 2] = u.z; wb[4 * i + 3] = u.w;
        }
      }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            const float sv = win[tap / 3 + py][tap % 3 + px];
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
              ca[py][px][ci] = fmaf(sv, wa[tap * 3 + ci], ca[py][px][ci]);
              if (DUAL) cb[py][px][ci] = fmaf(sv, wb[tap * 3 + ci], cb[py][px][ci]);
            }
          }
      }
    }
    fence_proxy_async();   // this thread's reads of buffer b are ordered before the TMA write issued below
    __syncthreads();
    if (tid == 0 && ch + 2 < NCH) {
      mbar_expect_tx(&full[b], kBoxBytes);
      tma_load_3d(&map, S + b * kLBox, &full[b], x0 - 4, y0 - 1, item * C + (ch + 2) * kLC);
    }
  }
  const int img = mult ? __ldg(img_index + item) : 0;
#pragma unroll
  for (int py = 0; py < 2; ++py) {
    const int y = y0 + ty + py;
    if (y >= H) continue;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const int x = x0 + tx + px;
      if (x >= W) continue;
      const size_t pix = (size_t)y * W + x;
      float* o = out + ((size_t)item * H * W + pix) * 3;
      if (mult) {
        const float* xi = images + ((size_t)img * H * W + pix) * 3;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float xv = __ldg(xi + ci);
          o[ci] = DUAL ? (xv >= 0.f ? xv * ca[py][px][ci] : xv * cb[py][px][ci]) : xv * ca[py][px][ci];
        }
      } else {
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) o[ci] = ca[py][px][ci];
      }
    }
  }
}

__global__ void posneg_kernel(const float* __restrict__ x, float* __restrict__ out, size_t pixels) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= pixels) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = x[idx * 3 + c];
    out[idx * 6 + c] = v >= 0.f ? v : 0.f;
    out[idx * 6 + 3 + c] = v < 0.f ? v : 0.f;
  }
}

__global__ void f32_to_split_kernel(const float* __restrict__ in, __nv_bfloat16* hi, int planes, size_t n) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float v = in[idx];
  if (planes == kPlanesH1F8) {   // debug operand: unscaled value in the fp16 + fp8 message layout
    __half* oh = reinterpret_cast<__half*>(hi);
    uint8_t* o8 = reinterpret_cast<uint8_t*>(hi) + n * 2;
    const __half h = __float2half_rn(v);
    oh[idx] = h;
    const size_t b = (idx >> 6) * 128 + (idx & 63);
    o8[b] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(h), __NV_SATFINITE, __NV_E4M3);
    o8[b + 64] = (uint8_t)__nv_cvt_float_to_fp8((v - __half2float(h)) * (float)(1 << kResShift), __NV_SATFINITE, __NV_E4M3);
    return;
  }
  if (planes == kPlanesF16x2) {
    __half* oh = reinterpret_cast<__half*>(hi);
    const __half h = __float2half_rn(v);
    oh[idx] = h;
    oh[n + idx] = __float2half_rn(v - __half2float(h));
    return;
  }
  for (int p = 0; p < planes; ++p) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[(size_t)p * n + idx] = h;
    v -= __bfloat162float(h);
  }
}
__global__ void split_to_f32_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                    float* out, size_t n) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  out[idx] = join_bf16(hi[idx], lo[idx]);
}

}  // namespace

int prep_weights_dual(const float* w_hwio, void* out, int cin, int cout, int fmt, int sign_a, float scale_a, int sign_b,
                      float scale_b, cudaStream_t s, int half_planes) {
  LRPCAP_REQUIRE(fmt == WF_SIMT_BWD || fmt == WF_TC_BWD, kErrInvalidArg, "prep_weights_dual: backward formats only");
  const size_t total = (size_t)9 * cin * 2 * cout;
  prep_weights_dual_kernel<<<grid_for(total, 256), 256, 0, s>>>(w_hwio, reinterpret_cast<float*>(out),
                                                                reinterpret_cast<__nv_bfloat16*>(out), cin, cout, fmt,
                                                                sign_a, scale_a, sign_b, scale_b, half_planes);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int prep_weights(const float* w_hwio, void* out, int cin, int cout, int fmt, int sign, cudaStream_t s, int taps,
                 int planes, float scale) {
  const size_t total = (size_t)taps * cin * cout;
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(out);
  prep_weights_kernel<<<grid_for(total, 256), 256, 0, s>>>(w_hwio, reinterpret_cast<float*>(out), hi, planes, cin, cout,
                                                           fmt, sign, taps, scale);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int pool_mask(const void* act, size_t act_elems, int planes, void* pooled, size_t pooled_elems, const float* G, float* Gc,
              unsigned* Gidx, int items, int H, int W, int C, cudaStream_t s) {
  LRPCAP_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 16 == 0, kErrShape, "pool_mask: H, W must be even and C %% 16 == 0");
  LRPCAP_REQUIRE(!G || Gc, kErrInvalidArg, "pool_mask: a multiplier needs its compact destination");
  const size_t total = (size_t)items * (H / 2) * (W / 2) * (C / 4);
  unsigned char* gi = reinterpret_cast<unsigned char*>(Gidx);
  if (planes == 3)
    pool_mask_kernel<StoreSplit3><<<grid_for(total, 256), 256, 0, s>>>(act, act_elems, pooled, pooled_elems, G, Gc, gi, items, H, W, C);
  else if (planes == 2)
    pool_mask_kernel<StoreSplit><<<grid_for(total, 256), 256, 0, s>>>(act, act_elems, pooled, pooled_elems, G, Gc, gi, items, H, W, C);
  else if (planes == kPlanesF16x2)
    pool_mask_kernel<StoreSplitH><<<grid_for(total, 256), 256, 0, s>>>(act, act_elems, pooled, pooled_elems, G, Gc, gi, items, H, W, C);
  else
    pool_mask_kernel<StoreF32><<<grid_for(total, 256), 256, 0, s>>>(act, act_elems, pooled, pooled_elems, G, Gc, gi, items, H, W, C);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int seed_message(const float* R, const float* M, const float* M2, const int* img_index, void* msg, size_t msg_elems,
                 bool split, int items, int pix, int C, int relu, cudaStream_t s) {
  const size_t per_item = (size_t)pix * C;
  LRPCAP_REQUIRE(per_item % 4 == 0, kErrShape, "seed_message: item size must be a multiple of 4");
  const size_t total4 = (size_t)items * per_item / 4;
  if (split)
    seed_kernel<StoreSplit><<<grid_for(total4, 256), 256, 0, s>>>(R, M, M2, img_index, msg, msg_elems, items, per_item, C, relu,
                                                                  nullptr, nullptr, nullptr, 0);
  else
    seed_kernel<StoreF32><<<grid_for(total4, 256), 256, 0, s>>>(R, M, M2, img_index, msg, msg_elems, items, per_item, C, relu,
                                                                nullptr, nullptr, nullptr, 0);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int seed_message_scaled(const float* R, const float* M, const float* M2, const int* img_index, void* msg, int items, int pix,
                        int C, int relu, unsigned* mx_true, unsigned* mx_out, int* kt_out, int target_exp, cudaStream_t s,
                        int fp8_planes) {
  const size_t per_item = (size_t)pix * C;
  LRPCAP_REQUIRE(per_item % 8 == 0 && C % 8 == 0, kErrShape, "seed_message_scaled: item size must be a multiple of 8");
  int bpi = (int)((per_item / 4 + 255) / 256);
  if (bpi > 32) bpi = 32;
  seed_max_kernel<<<dim3((unsigned)bpi, (unsigned)items), 256, 0, s>>>(R, M, M2, img_index, mx_true, per_item, relu);
  LRPCAP_CUDA(cudaGetLastError());
  const size_t total8 = (size_t)items * per_item / 8;
  const size_t msg_elems = (size_t)items * per_item * (M2 ? 2 : 1);
  if (fp8_planes)
    seed_kernel<StoreH1F8><<<grid_for(total8, 256), 256, 0, s>>>(R, M, M2, img_index, msg, msg_elems, items, per_item, C, relu,
                                                                 mx_true, mx_out, kt_out, target_exp);
  else
    seed_kernel<StoreH1><<<grid_for(total8, 256), 256, 0, s>>>(R, M, M2, img_index, msg, 0, items, per_item, C, relu, mx_true,
                                                               mx_out, kt_out, target_exp);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

template <bool DUAL, int C>
int launch_last(const CUtensorMap& map, const float* Wa, const float* Wb, const float* images, const int* img_index,
                float* out, int H, int W, int tiles_x, int tiles_y, int mult, unsigned grid, cudaStream_t s) {
  const int smem = (2 * kLBox + (DUAL ? 2 : 1) * C * kLWPitch) * (int)sizeof(float) + 16 + 128;
  static int smem_state[kMaxDevices] = {};
  LRPCAP_CUDA(ensure_dynamic_smem(last_dgrad_kernel<DUAL, C>, smem, smem_state));
  last_dgrad_kernel<DUAL, C><<<grid, kLThreads, smem, s>>>(map, Wa, Wb, images, img_index, out, H, W, tiles_x, tiles_y, mult);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int last_dgrad(const float* msg, const float* Wa, const float* Wb, const float* images, const int* img_index, float* out,
               int items, int H, int W, int C, int mult, cudaStream_t s) {
  LRPCAP_REQUIRE(C == 64 || C == 128, kErrShape, "last_dgrad: C must be 64 (or 128 for the dual message), got %d", C);
  const int tiles_x = ceil_div(W, kLTX), tiles_y = ceil_div(H, kLTY);
  const long long blocks = (long long)items * tiles_x * tiles_y;
  LRPCAP_REQUIRE(blocks > 0 && blocks < (1ll << 31), kErrShape, "last_dgrad: grid out of range");
  const unsigned g = (unsigned)blocks;
  CUtensorMap map;
  LRPCAP_TRY(make_map_planar_f32(&map, msg, items * C, H, W, kLPSX, kLPSY, kLC));
  if (Wb) {
    if (C == 64) return launch_last<true, 64>(map, Wa, Wb, images, img_index, out, H, W, tiles_x, tiles_y, mult, g, s);
    return launch_last<true, 128>(map, Wa, Wb, images, img_index, out, H, W, tiles_x, tiles_y, mult, g, s);
  }
  if (C == 64) return launch_last<false, 64>(map, Wa, Wb, images, img_index, out, H, W, tiles_x, tiles_y, mult, g, s);
  return launch_last<false, 128>(map, Wa, Wb, images, img_index, out, H, W, tiles_x, tiles_y, mult, g, s);
}

int make_posneg(const float* x, float* out, size_t pixels, cudaStream_t s) {
  posneg_kernel<<<grid_for(pixels, 256), 256, 0, s>>>(x, out, pixels);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int f32_to_split(const float* in, void* out, size_t n, cudaStream_t s, int planes) {
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(out);
  f32_to_split_kernel<<<grid_for(n, 256), 256, 0, s>>>(in, hi, planes, n);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}
int split_to_f32(const void* in, float* out, size_t n, cudaStream_t s) {
  const __nv_bfloat16* hi = reinterpret_cast<const __nv_bfloat16*>(in);
  split_to_f32_kernel<<<grid_for(n, 256), 256, 0, s>>>(hi, hi + n, out, n);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace lrpcap
