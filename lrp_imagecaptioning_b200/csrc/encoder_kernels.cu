// HBM-bound helper kernels of the encoder relevance path: all coalesced, 16-byte vectorised.
#include "encoder_kernels.cuh"
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
                                    int cout, int fmt, int sign, int taps) {
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
                                         float scale_a, int sign_b, float scale_b) {
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
                                 float* G, int items, int H, int W, int C) {
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
  size_t off[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    off[p] = (((size_t)item * H + (2 * yo + (p >> 1))) * W + (2 * xo + (p & 1))) * C + c;
    ST::template load<4>(act, act_elems, off[p], v[p]);
  }
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
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float g[4];
      load_f32<4>(G + off[p], g);
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i] = (am[i] == p) ? g[i] : 0.f;
      store_f32<4>(G + off[p], g);
    }
  }
}

// ------------------------------------------------------------------ seed message
template <class ST>
__global__ void seed_kernel(const float* __restrict__ R, const float* __restrict__ M, const float* __restrict__ M2,
                            const int* __restrict__ img_index, void* msg, size_t msg_elems, int items, size_t per_item,
                            int C, int relu) {
  const size_t total4 = (size_t)items * per_item / 4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const size_t e = idx * 4;
  const int item = e / per_item;
  const size_t in_item = e - (size_t)item * per_item;
  const int img = __ldg(img_index + item);
  float r[4], m[4], o[4];
  load_f32<4>(R + e, r);
  load_f32<4>(M + (size_t)img * per_item + in_item, m);
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = (relu ? fmaxf(r[i], 0.f) : r[i]) * m[i];
  if (!M2) {
    ST::template store<4>(msg, msg_elems, e, o);
    return;
  }
  const size_t pix = e / C;               // dual layout: [.., pixel, 2C] = [R*M | R*M2]
  const int c = (int)(e - pix * C);
  ST::template store<4>(msg, msg_elems, pix * 2 * C + c, o);
  load_f32<4>(M2 + (size_t)img * per_item + in_item, m);
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = r[i] * m[i];
  ST::template store<4>(msg, msg_elems, pix * 2 * C + C + c, o);
}

// ------------------------------------------------------------------ last transposed conv (C -> 3) + re-weighting
// HBM-bound (12.85 MB in, 0.6 MB out per word at 224x224). Weights sit in constant memory so every FMA takes its
// weight as a uniform constant operand; a thread owns two horizontally adjacent pixels and re-uses its 3x4 window.
constexpr int kLTY = 16, kLTX = 32;   // tile (rows x cols), 256 threads x 2 pixels
constexpr int kLC = 16;               // channel chunk staged in shared memory
constexpr int kLastMaxC = 128;   // 64 channels, or 2 x 64 for the dual (beta != 0) message
__constant__ float c_wlast[2][9 * kLastMaxC * 3];

template <class ST, bool DUAL, int C>
__global__ void __launch_bounds__(256)
last_dgrad_kernel(const void* __restrict__ msg, size_t msg_elems, const float* __restrict__ images,
                  const int* __restrict__ img_index, float* __restrict__ out, int H, int W, int tiles_x,
                  int tiles_y, int mult) {
  constexpr int PSX = kLTX + 2, PSY = kLTY + 2;
  __shared__ float S[kLC][PSX * PSY + 1];

  int bid = blockIdx.x;
  const int tiles = tiles_x * tiles_y;
  const int item = bid / tiles;
  bid -= item * tiles;
  const int y0 = (bid / tiles_x) * kLTY, x0 = (bid % tiles_x) * kLTX;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = (tid & 15) * 2;

  float ca[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, cb[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
  for (int c0 = 0; c0 < C; c0 += kLC) {   // fully unrolled: every constant-bank offset below is an immediate
    __syncthreads();
    for (int pix = tid; pix < PSX * PSY; pix += 256) {     // one halo pixel (16 channels = 4 x 16 B loads) per thread
      const int py = pix / PSX, px = pix - py * PSX;
      const int gy = y0 + py - 1, gx = x0 + px - 1;
      float v[kLC];
#pragma unroll
      for (int i = 0; i < kLC; ++i) v[i] = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        ST::template load<kLC>(msg, msg_elems, (((size_t)item * H + gy) * W + gx) * C + c0, v);
#pragma unroll
      for (int i = 0; i < kLC; ++i) S[i][pix] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kLC; ++k) {
      float win[3][4];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) win[r][q] = S[k][(ty + r) * PSX + tx + q];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float* wa = &c_wlast[0][(tap * C + c0 + k) * 3];
        const float* wb = &c_wlast[1][(tap * C + c0 + k) * 3];
#pragma unroll
        for (int px = 0; px < 2; ++px) {
          const float sv = win[tap / 3][tap % 3 + px];
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) {
            ca[px][ci] = fmaf(sv, wa[ci], ca[px][ci]);
            if (DUAL) cb[px][ci] = fmaf(sv, wb[ci], cb[px][ci]);
          }
        }
      }
    }
  }
  const int y = y0 + ty;
  if (y >= H) return;
#pragma unroll
  for (int px = 0; px < 2; ++px) {
    const int x = x0 + tx + px;
    if (x >= W) continue;
    const size_t pix = (size_t)y * W + x;
    float* o = out + ((size_t)item * H * W + pix) * 3;
    if (mult) {
      const int img = __ldg(img_index + item);
      const float* xi = images + ((size_t)img * H * W + pix) * 3;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float xv = __ldg(xi + ci);
        o[ci] = DUAL ? (xv >= 0.f ? xv * ca[px][ci] : xv * cb[px][ci]) : xv * ca[px][ci];
      }
    } else {
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) o[ci] = ca[px][ci];
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
                      float scale_b, cudaStream_t s) {
  LRPCAP_REQUIRE(fmt == WF_SIMT_BWD || fmt == WF_TC_BWD, kErrInvalidArg, "prep_weights_dual: backward formats only");
  const size_t total = (size_t)9 * cin * 2 * cout;
  prep_weights_dual_kernel<<<grid_for(total, 256), 256, 0, s>>>(w_hwio, reinterpret_cast<float*>(out),
                                                                reinterpret_cast<__nv_bfloat16*>(out), cin, cout, fmt,
                                                                sign_a, scale_a, sign_b, scale_b);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int prep_weights(const float* w_hwio, void* out, int cin, int cout, int fmt, int sign, cudaStream_t s, int taps,
                 int planes) {
  const size_t total = (size_t)taps * cin * cout;
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(out);
  prep_weights_kernel<<<grid_for(total, 256), 256, 0, s>>>(w_hwio, reinterpret_cast<float*>(out), hi, planes, cin, cout,
                                                           fmt, sign, taps);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int pool_mask(const void* act, size_t act_elems, int planes, void* pooled, size_t pooled_elems, float* G, int items,
              int H, int W, int C, cudaStream_t s) {
  LRPCAP_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, kErrShape, "pool_mask: H, W must be even and C %% 4 == 0");
  const size_t total = (size_t)items * (H / 2) * (W / 2) * (C / 4);
  if (planes == 3)
    pool_mask_kernel<StoreSplit3><<<grid_for(total, 256), 256, 0, s>>>(act, act_elems, pooled, pooled_elems, G, items, H, W, C);
  else if (planes == 2)
    pool_mask_kernel<StoreSplit><<<grid_for(total, 256), 256, 0, s>>>(act, act_elems, pooled, pooled_elems, G, items, H, W, C);
  else
    pool_mask_kernel<StoreF32><<<grid_for(total, 256), 256, 0, s>>>(act, act_elems, pooled, pooled_elems, G, items, H, W, C);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int seed_message(const float* R, const float* M, const float* M2, const int* img_index, void* msg, size_t msg_elems,
                 bool split, int items, int pix, int C, int relu, cudaStream_t s) {
  const size_t per_item = (size_t)pix * C;
  LRPCAP_REQUIRE(per_item % 4 == 0, kErrShape, "seed_message: item size must be a multiple of 4");
  const size_t total4 = (size_t)items * per_item / 4;
  if (split)
    seed_kernel<StoreSplit><<<grid_for(total4, 256), 256, 0, s>>>(R, M, M2, img_index, msg, msg_elems, items, per_item, C, relu);
  else
    seed_kernel<StoreF32><<<grid_for(total4, 256), 256, 0, s>>>(R, M, M2, img_index, msg, msg_elems, items, per_item, C, relu);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int last_dgrad(const void* msg, size_t msg_elems, bool split, const float* Wa, const float* Wb, const float* images,
               const int* img_index, float* out, int items, int H, int W, int C, int mult, cudaStream_t s) {
  LRPCAP_REQUIRE(C % kLC == 0 && C <= kLastMaxC, kErrShape, "last_dgrad: C must be a multiple of %d and <= %d", kLC, kLastMaxC);
  const int tiles_x = ceil_div(W, kLTX), tiles_y = ceil_div(H, kLTY);
  const long long blocks = (long long)items * tiles_x * tiles_y;
  LRPCAP_REQUIRE(blocks > 0 && blocks < (1ll << 31), kErrShape, "last_dgrad: grid out of range");
  const unsigned g = (unsigned)blocks;
  const size_t wbytes = (size_t)9 * C * 3 * sizeof(float);
  // stream-ordered refresh of the constant bank (one encoder stream at a time uses it)
  LRPCAP_CUDA(cudaMemcpyToSymbolAsync(c_wlast, Wa, wbytes, 0, cudaMemcpyDeviceToDevice, s));
  if (Wb) LRPCAP_CUDA(cudaMemcpyToSymbolAsync(c_wlast, Wb, wbytes, sizeof(float) * 9 * kLastMaxC * 3, cudaMemcpyDeviceToDevice, s));
  LRPCAP_REQUIRE(C == 64 || C == 128, kErrShape, "last_dgrad: C must be 64 (or 128 for the dual message)");
#define LRPCAP_LAUNCH_LAST(ST, DUAL, CC) \
  last_dgrad_kernel<ST, DUAL, CC><<<g, 256, 0, s>>>(msg, msg_elems, images, img_index, out, H, W, tiles_x, tiles_y, mult)
#define LRPCAP_LAUNCH_LAST_C(ST, DUAL) \
  do { if (C == 64) LRPCAP_LAUNCH_LAST(ST, DUAL, 64); else LRPCAP_LAUNCH_LAST(ST, DUAL, 128); } while (0)
  if (split) {
    if (Wb) LRPCAP_LAUNCH_LAST_C(StoreSplit, true); else LRPCAP_LAUNCH_LAST_C(StoreSplit, false);
  } else {
    if (Wb) LRPCAP_LAUNCH_LAST_C(StoreF32, true); else LRPCAP_LAUNCH_LAST_C(StoreF32, false);
  }
#undef LRPCAP_LAUNCH_LAST_C
#undef LRPCAP_LAUNCH_LAST
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
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
