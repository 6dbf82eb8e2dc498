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
                                    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int cin,
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
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    out_hi[idx] = h;
    out_lo[idx] = l;
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
__global__ void seed_kernel(const float* __restrict__ R, const float* __restrict__ M, const int* __restrict__ img_index,
                            void* msg, size_t msg_elems, int items, size_t per_item, int relu) {
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
  ST::template store<4>(msg, msg_elems, e, o);
}

// ------------------------------------------------------------------ last transposed conv (C -> 3) + re-weighting
constexpr int kLT = 16;   // tile side
constexpr int kLC = 16;   // channel chunk

template <class ST, bool DUAL>
__global__ void __launch_bounds__(kLT * kLT)
last_dgrad_kernel(const void* __restrict__ msg, size_t msg_elems, const float* __restrict__ Wa,
                  const float* __restrict__ Wb, const float* __restrict__ images, const int* __restrict__ img_index,
                  float* __restrict__ out, int H, int W, int C, int tiles_x, int tiles_y, int mult) {
  constexpr int PS = kLT + 2;
  __shared__ float S[kLC][PS * PS + 1];
  __shared__ float wa[9][kLC][3];
  __shared__ float wb[DUAL ? 9 : 1][kLC][3];

  int bid = blockIdx.x;
  const int tiles = tiles_x * tiles_y;
  const int item = bid / tiles;
  bid -= item * tiles;
  const int y0 = (bid / tiles_x) * kLT, x0 = (bid % tiles_x) * kLT;
  const int tid = threadIdx.x;
  const int ty = tid / kLT, tx = tid % kLT;

  float ca[3] = {0.f, 0.f, 0.f}, cb[3] = {0.f, 0.f, 0.f};
  for (int c0 = 0; c0 < C; c0 += kLC) {
    __syncthreads();
    for (int idx = tid; idx < PS * PS * (kLC / 4); idx += kLT * kLT) {
      const int k4 = idx % (kLC / 4);
      const int pix = idx / (kLC / 4);
      const int py = pix / PS, px = pix - py * PS;
      const int gy = y0 + py - 1, gx = x0 + px - 1;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        ST::template load<4>(msg, msg_elems, (((size_t)item * H + gy) * W + gx) * C + c0 + k4 * 4, v);
#pragma unroll
      for (int i = 0; i < 4; ++i) S[k4 * 4 + i][pix] = v[i];
    }
    for (int idx = tid; idx < 9 * kLC * 3; idx += kLT * kLT) {
      const int ci = idx % 3;
      const int k = (idx / 3) % kLC;
      const int tap = idx / (3 * kLC);
      wa[tap][k][ci] = __ldg(Wa + ((size_t)tap * C + c0 + k) * 3 + ci);
      if (DUAL) wb[tap][k][ci] = __ldg(Wb + ((size_t)tap * C + c0 + k) * 3 + ci);
    }
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int p = (ty + tap / 3) * PS + tx + tap % 3;
#pragma unroll
      for (int k = 0; k < kLC; ++k) {
        const float sv = S[k][p];
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          ca[ci] = fmaf(sv, wa[tap][k][ci], ca[ci]);
          if (DUAL) cb[ci] = fmaf(sv, wb[tap][k][ci], cb[ci]);
        }
      }
    }
  }
  const int y = y0 + ty, x = x0 + tx;
  if (y < H && x < W) {
    const size_t pix = (size_t)y * W + x;
    float* o = out + ((size_t)item * H * W + pix) * 3;
    if (mult) {
      const int img = __ldg(img_index + item);
      const float* xi = images + ((size_t)img * H * W + pix) * 3;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float xv = __ldg(xi + ci);
        o[ci] = DUAL ? (xv >= 0.f ? xv * ca[ci] : xv * cb[ci]) : xv * ca[ci];
      }
    } else {
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) o[ci] = ca[ci];
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

__global__ void f32_to_split_kernel(const float* __restrict__ in, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  __nv_bfloat16 h, l;
  split_bf16(in[idx], h, l);
  hi[idx] = h;
  lo[idx] = l;
}
__global__ void split_to_f32_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                    float* out, size_t n) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  out[idx] = join_bf16(hi[idx], lo[idx]);
}

}  // namespace

int prep_weights(const float* w_hwio, void* out, int cin, int cout, int fmt, int sign, cudaStream_t s, int taps) {
  const size_t total = (size_t)taps * cin * cout;
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(out);
  prep_weights_kernel<<<grid_for(total, 256), 256, 0, s>>>(w_hwio, reinterpret_cast<float*>(out), hi, hi + total, cin,
                                                           cout, fmt, sign, taps);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int pool_mask(const void* act, size_t act_elems, bool split, void* pooled, size_t pooled_elems, float* G, int items,
              int H, int W, int C, cudaStream_t s) {
  LRPCAP_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, kErrShape, "pool_mask: H, W must be even and C %% 4 == 0");
  const size_t total = (size_t)items * (H / 2) * (W / 2) * (C / 4);
  if (split)
    pool_mask_kernel<StoreSplit><<<grid_for(total, 256), 256, 0, s>>>(act, act_elems, pooled, pooled_elems, G, items, H, W, C);
  else
    pool_mask_kernel<StoreF32><<<grid_for(total, 256), 256, 0, s>>>(act, act_elems, pooled, pooled_elems, G, items, H, W, C);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int seed_message(const float* R, const float* M, const int* img_index, void* msg, size_t msg_elems, bool split,
                 int items, int pix, int C, int relu, cudaStream_t s) {
  const size_t per_item = (size_t)pix * C;
  LRPCAP_REQUIRE(per_item % 4 == 0, kErrShape, "seed_message: item size must be a multiple of 4");
  const size_t total4 = (size_t)items * per_item / 4;
  if (split)
    seed_kernel<StoreSplit><<<grid_for(total4, 256), 256, 0, s>>>(R, M, img_index, msg, msg_elems, items, per_item, relu);
  else
    seed_kernel<StoreF32><<<grid_for(total4, 256), 256, 0, s>>>(R, M, img_index, msg, msg_elems, items, per_item, relu);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int last_dgrad(const void* msg, size_t msg_elems, bool split, const float* Wa, const float* Wb, const float* images,
               const int* img_index, float* out, int items, int H, int W, int C, int mult, cudaStream_t s) {
  LRPCAP_REQUIRE(C % kLC == 0, kErrShape, "last_dgrad: C must be a multiple of %d", kLC);
  const int tiles_x = ceil_div(W, kLT), tiles_y = ceil_div(H, kLT);
  const long long blocks = (long long)items * tiles_x * tiles_y;
  LRPCAP_REQUIRE(blocks > 0 && blocks < (1ll << 31), kErrShape, "last_dgrad: grid out of range");
  const unsigned g = (unsigned)blocks;
#define LRPCAP_LAUNCH_LAST(ST, DUAL) \
  last_dgrad_kernel<ST, DUAL><<<g, kLT * kLT, 0, s>>>(msg, msg_elems, Wa, Wb, images, img_index, out, H, W, C, tiles_x, tiles_y, mult)
  if (split) {
    if (Wb) LRPCAP_LAUNCH_LAST(StoreSplit, true); else LRPCAP_LAUNCH_LAST(StoreSplit, false);
  } else {
    if (Wb) LRPCAP_LAUNCH_LAST(StoreF32, true); else LRPCAP_LAUNCH_LAST(StoreF32, false);
  }
#undef LRPCAP_LAUNCH_LAST
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int make_posneg(const float* x, float* out, size_t pixels, cudaStream_t s) {
  posneg_kernel<<<grid_for(pixels, 256), 256, 0, s>>>(x, out, pixels);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int f32_to_split(const float* in, void* out, size_t n, cudaStream_t s) {
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(out);
  f32_to_split_kernel<<<grid_for(n, 256), 256, 0, s>>>(in, hi, hi + n, n);
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
