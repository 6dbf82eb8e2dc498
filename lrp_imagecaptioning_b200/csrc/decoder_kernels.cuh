// fp64 kernels of the decoder stage (element-wise LRP rules, attention, LSTM cell, DGEMM).
// Included only by decoder.cu.
#pragma once
#include "common.cuh"

namespace lrpcap {
namespace dk {

constexpr double kEps = 1e-7;   // keras.backend.epsilon(): default eps of explainers.py:157

__device__ __forceinline__ double stabd(double z) { return z + (z >= 0.0 ? kEps : -kEps); }   // explainers.py:141-144
__device__ __forceinline__ double sigm(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double f32r(double x) { return (double)(float)x; }                   // a float32 store

// ------------------------------------------------------------------ DGEMM  C = A B (+ bias), row-major
constexpr int GM = 64, GN = 64, GK = 16;
__global__ void __launch_bounds__(256)
dgemm_kernel(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb, double* __restrict__ C,
             int ldc, int M, int N, int Ktot, const double* __restrict__ bias, int k_per_split) {
  // blockIdx.z = K split: slice z covers [z*k_per_split, min(Ktot, (z+1)*k_per_split)) and writes partial sums to
  // C + z*M*ldc (deterministic two-pass split-K for skinny problems; k_per_split == Ktot when not split)
  __shared__ double As[GK][GM + 1];
  __shared__ double Bs[GK][GN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int kb = blockIdx.z * k_per_split;
  const int K = (kb + k_per_split < Ktot) ? kb + k_per_split : Ktot;
  C += (size_t)blockIdx.z * M * ldc;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  // register prefetch: the global loads of tile k0 + GK are in flight while tile k0 is multiplied
  double ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;          // 64 x 16
      const int mm = idx >> 4, kk = idx & 15;
      const int gm = m0 + mm, gk = k0 + kk;
      ra[r] = (gm < M && gk < K) ? __ldg(A + (size_t)gm * lda + gk) : 0.0;
      const int kr = idx >> 6, nn = idx & 63;  // 16 x 64
      const int gn = n0 + nn, gkb = k0 + kr;
      rb[r] = (gn < N && gkb < K) ? __ldg(B + (size_t)gkb * ldb + gn) : 0.0;
    }
  };
  fetch(kb);
  for (int k0 = kb; k0 < K; k0 += GK) {
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;
      As[idx & 15][idx >> 4] = ra[r];
      Bs[idx >> 6][idx & 63] = rb[r];
    }
    __syncthreads();
    if (k0 + GK < K) fetch(k0 + GK);
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) C[(size_t)gm * ldc + gn] = acc[i][j] + (bias ? bias[gn] : 0.0);
    }
  }
}

__global__ void splitk_reduce_kernel(const double* __restrict__ ws, int splits, double* __restrict__ C, int ldc, int M,
                                     int N, const double* __restrict__ bias) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * N) return;
  const int m = i / N, n = i - (size_t)m * N;
  double acc = 0.0;
  for (int z = 0; z < splits; ++z) acc += ws[((size_t)z * M + m) * N + n];
  C[(size_t)m * ldc + n] = acc + (bias ? bias[n] : 0.0);
}

// ------------------------------------------------------------------ forward: per-image features
__global__ void f32_to_f64_kernel(const float* __restrict__ in, double* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}
// a[n,d] = mean_l F[n,l,d], kept at float32 precision like np.mean of a float32 array (explainers.py:382)
__global__ void mean_feat_kernel(const double* __restrict__ F, double* __restrict__ a, int L, int D) {
  const int n = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double s = 0.0;
    for (int l = 0; l < L; ++l) s += F[((size_t)n * L + l) * D + d];
    a[(size_t)n * D + d] = f32r(s / L);
  }
}
__global__ void f64_to_split_kernel(const double* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  __nv_bfloat16 h, l;
  split_bf16((float)in[i], h, l);
  hi[i] = h;
  lo[i] = l;
}
// as f64_to_split_kernel, the two planes given separately (the destination may be padded beyond n)
__global__ void f64_to_split_strided_kernel(const double* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                            __nv_bfloat16* __restrict__ lo, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  __nv_bfloat16 h, l;
  split_bf16((float)in[i], h, l);
  hi[i] = h;
  lo[i] = l;
}
// rows of a strided fp64 matrix -> three bf16 planes [Mpad, K] (rows >= M zero): fp32-exact GEMM operand
__global__ void f64_rows_to_split3_kernel(const double* __restrict__ A, int lda, int M, int K, int Mpad,
                                          __nv_bfloat16* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)Mpad * K;
  if (i >= n) return;
  const int m = (int)(i / K), k = (int)(i - (size_t)m * K);
  float v = m < M ? (float)A[(size_t)m * lda + k] : 0.f;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[(size_t)p * n + i] = h;
    v -= __bfloat162float(h);
  }
}
// C[m, n] = (double)C32[m, n] (+ bias[n]) for the un-padded part of a tensor-core result
__global__ void f32_to_f64_bias_kernel(const float* __restrict__ C32, int ld32, double* __restrict__ C, int ldc, int M,
                                       int N, const double* __restrict__ bias) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * N) return;
  const int m = (int)(i / N), n = (int)(i - (size_t)m * N);
  C[(size_t)m * ldc + n] = (double)C32[(size_t)m * ld32 + n] + (bias ? bias[n] : 0.0);
}
__global__ void round_f32_kernel(double* x, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = f32r(x[i]);
}
__global__ void relu_copy_kernel(const double* __restrict__ in, double* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fmax(in[i], 0.0);
}

// ------------------------------------------------------------------ forward: one decoder step
// adaptive: XH[n,i] = [emb(tok_prev), g, h_i]                       (explainers.py:400-409)
// grid-TD : XH1[n,i] = [h2_i, g, emb(tok_prev), h1_i]               (explainers.py:1131-1136)
__global__ void build_xh_kernel(double* __restrict__ XH, const double* __restrict__ Emb, const double* __restrict__ gp,
                                const double* __restrict__ h1, const double* __restrict__ h2,
                                const int* __restrict__ tok, int i, int T, int H, int E, int sos, int gridtd) {
  const int n = blockIdx.x;
  const int Kin = gridtd ? (2 * H + 2 * E) : (2 * E + H);
  const int t = (i == 0) ? sos : tok[n * T + i - 1];
  double* x = XH + ((size_t)n * T + i) * Kin;
  const size_t so = ((size_t)n * (T + 1) + i) * H;
  for (int j = threadIdx.x; j < Kin; j += blockDim.x) {
    double v;
    if (!gridtd) {
      if (j < E) v = Emb[(size_t)(t - 1) * E + j];
      else if (j < 2 * E) v = fmax(gp[(size_t)n * E + j - E], 0.0);
      else v = h1[so + j - 2 * E];
    } else {
      if (j < H) v = h2[so + j];
      else if (j < H + E) v = fmax(gp[(size_t)n * E + j - H], 0.0);
      else if (j < H + 2 * E) v = Emb[(size_t)(t - 1) * E + j - H - E];
      else v = h1[so + j - H - 2 * E];
    }
    x[j] = v;
  }
}
// grid-TD: XH2[n,i] = [c_hat_{i+1}, h1_{i+1}, h2_i]                  (explainers.py:1151)
__global__ void build_xh2_kernel(double* __restrict__ XH2, const double* __restrict__ chat,
                                 const double* __restrict__ h1, const double* __restrict__ h2, int i, int T, int H) {
  const int n = blockIdx.x;
  double* x = XH2 + ((size_t)n * T + i) * 3 * H;
  const size_t s0 = ((size_t)n * (T + 1) + i) * H, s1 = s0 + H;
  for (int j = threadIdx.x; j < 3 * H; j += blockDim.x)
    x[j] = (j < H) ? chat[s1 + j] : (j < 2 * H ? h1[s1 + j - H] : h2[s0 + j - 2 * H]);
}
// LSTM cell point-wise part, Keras gate order i, f, c, o              (explainers.py:125-139)
__global__ void lstm_point_kernel(const double* __restrict__ Z, double* __restrict__ h, double* __restrict__ c,
                                  double* __restrict__ zg, double* __restrict__ ia, double* __restrict__ fa,
                                  double* __restrict__ ga, double* __restrict__ oa, int i, int T, int H) {
  const int n = blockIdx.x;
  const size_t s0 = ((size_t)n * (T + 1) + i) * H, s1 = s0 + H;
  const double* z = Z + (size_t)n * 4 * H;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const double i_ = sigm(z[j]), f_ = sigm(z[H + j]), g_ = tanh(z[2 * H + j]), o_ = sigm(z[3 * H + j]);
    const double cn = f_ * c[s0 + j] + i_ * g_;
    c[s1 + j] = cn;
    h[s1 + j] = o_ * tanh(cn);
    zg[s1 + j] = z[2 * H + j];
    ia[s1 + j] = i_;
    fa[s1 + j] = f_;
    ga[s1 + j] = g_;
    oa[s1 + j] = o_;
  }
}
// sentinel: s = tanh(c_{i+1}) * sigmoid(x W_x + h_i W_h)              (explainers.py:415, 1145)
__global__ void sentinel_kernel(const double* __restrict__ sg, const double* __restrict__ c, double* __restrict__ s,
                                int i, int T, int H) {
  const int n = blockIdx.x;
  const size_t s1 = ((size_t)n * (T + 1) + i + 1) * H;
  for (int j = threadIdx.x; j < H; j += blockDim.x) s[s1 + j] = tanh(c[s1 + j]) * sigm(sg[(size_t)n * H + j]);
}
// attention scores e[n,l] = tanh(P[n,l] + hp[n]) . V  (l < L);  e[n,L] = tanh(sp[n] + hp[n]) . V   (explainers.py:413-417)
__global__ void __launch_bounds__(128)
scores_kernel(const double* __restrict__ P, const double* __restrict__ hp, const double* __restrict__ sp,
              const double* __restrict__ Va, double* __restrict__ e, int L, int H) {
  const int l = blockIdx.x, n = blockIdx.y;
  const double* base = (l < L) ? P + ((size_t)n * L + l) * H : sp + (size_t)n * H;
  double acc = 0.0;
  for (int j = threadIdx.x; j < H; j += 128) acc += tanh(base[j] + hp[(size_t)n * H + j]) * Va[j];
  __shared__ double red[128];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 64; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) e[(size_t)n * (L + 1) + l] = red[0];
}
// alpha = softmax(e[0:L]);  beta = softmax(e[0:L+1])[L]              (explainers.py:414, 418)
__global__ void __launch_bounds__(256)
softmax_kernel(const double* __restrict__ e, double* __restrict__ alpha, double* __restrict__ beta, int i, int T, int L) {
  const int n = blockIdx.x;
  const double* en = e + (size_t)n * (L + 1);
  __shared__ double red[256];
  double m = -1e300;
  for (int l = threadIdx.x; l < L; l += 256) m = fmax(m, en[l]);
  red[threadIdx.x] = m;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + st]);
    __syncthreads();
  }
  const double m1 = red[0];
  __syncthreads();
  double s = 0.0;
  for (int l = threadIdx.x; l < L; l += 256) s += exp(en[l] - m1);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  const double s1 = red[0];
  double* an = alpha + ((size_t)n * (T + 1) + i + 1) * L;
  for (int l = threadIdx.x; l < L; l += 256) an[l] = exp(en[l] - m1) / s1;
  if (threadIdx.x == 0) {
    const double m2 = fmax(m1, en[L]);
    beta[(size_t)n * (T + 1) + i + 1] = exp(en[L] - m2) / (s1 * exp(m1 - m2) + exp(en[L] - m2));
  }
}
// ctx = sum_l alpha_l Vf_l ; c_hat = beta s + (1-beta) ctx ; hc = h_last + c_hat (optional)   (explainers.py:419-421)
__global__ void context_kernel(const double* __restrict__ Vf, const double* __restrict__ alpha,
                               const double* __restrict__ beta, const double* __restrict__ s, double* __restrict__ ctx,
                               double* __restrict__ chat, const double* __restrict__ hlast, double* __restrict__ hc,
                               int i, int T, int L, int H) {
  const int n = blockIdx.x;
  const size_t s1 = ((size_t)n * (T + 1) + i + 1) * H;
  const double* an = alpha + ((size_t)n * (T + 1) + i + 1) * L;
  const double b = beta[(size_t)n * (T + 1) + i + 1];
  for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < H; j += gridDim.y * blockDim.x) {
    double acc = 0.0;
#pragma unroll 14   // independent loads in flight; the additions keep their order
    for (int l = 0; l < L; ++l) acc += an[l] * Vf[((size_t)n * L + l) * H + j];
    ctx[s1 + j] = acc;
    const double ch = b * s[s1 + j] + (1.0 - b) * acc;
    chat[s1 + j] = ch;
    if (hc) hc[(size_t)n * H + j] = hlast[s1 + j] + ch;
  }
}
// hc = h2_{i+1} (+ c_hat_{i+1} in keras-logits mode)
__global__ void gridtd_hc_kernel(const double* __restrict__ h2, const double* __restrict__ chat, double* __restrict__ hc,
                                 int i, int T, int H, int add_chat) {
  const int n = blockIdx.x;
  const size_t s1 = ((size_t)n * (T + 1) + i + 1) * H;
  for (int j = threadIdx.x; j < H; j += blockDim.x) hc[(size_t)n * H + j] = h2[s1 + j] + (add_chat ? chat[s1 + j] : 0.0);
}
// teacher-forced: logit of the caption token only (the relevance pass reads nothing else of the logits)
__global__ void __launch_bounds__(128)
logitk_kernel(const double* __restrict__ hc, const double* __restrict__ WoT, const double* __restrict__ bo,
              const int* __restrict__ tok, double* __restrict__ logitk, int i, int T, int H) {
  const int n = blockIdx.x;
  const int k = tok[n * T + i] - 1;
  double acc = 0.0;
  for (int j = threadIdx.x; j < H; j += 128) acc += hc[(size_t)n * H + j] * WoT[(size_t)k * H + j];
  __shared__ double red[128];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 64; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) logitk[n * T + i] = red[0] + bo[k];
}
// greedy: token = arg-max of the logits (first maximum), optionally never the EOS id
__global__ void __launch_bounds__(256)
argmax_kernel(const double* __restrict__ logits, int V, int eos_index, int* __restrict__ tok, double* __restrict__ logitk,
              int i, int T) {
  const int n = blockIdx.x;
  const double* ln = logits + (size_t)n * V;
  double best = -1e300;
  int bi = V;
  for (int v = threadIdx.x; v < V; v += 256) {
    if (v == eos_index) continue;
    const double x = ln[v];
    if (x > best) { best = x; bi = v; }
  }
  __shared__ double rb[256];
  __shared__ int ri[256];
  rb[threadIdx.x] = best;
  ri[threadIdx.x] = bi;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) {
      const double ob = rb[threadIdx.x + st];
      const int oi = ri[threadIdx.x + st];
      if (ob > rb[threadIdx.x] || (ob == rb[threadIdx.x] && oi < ri[threadIdx.x])) {
        rb[threadIdx.x] = ob;
        ri[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    tok[n * T + i] = ri[0] + 1;   // model index -> tokenizer id (explainers.py:92)
    logitk[n * T + i] = rb[0];
  }
}

// ------------------------------------------------------------------ relevance: element-wise rules
// ew(R, a, z) = a * R / stab(z)  (identity-weight helper call);  all kernels: one block per active word.
struct WordRef {
  const int* img;   // [W] image of sorted word p
  const int* t;     // [W] 1-based position of sorted word p
};

// output layer + merge split (explainers.py:569-602 / 1212-1229)
__global__ void lrp_init_kernel(WordRef w, const double* __restrict__ hlast, const double* __restrict__ chat,
                                const double* __restrict__ ctx, const double* __restrict__ s,
                                const double* __restrict__ beta, const double* __restrict__ logitk,
                                const double* __restrict__ WoT, const int* __restrict__ tok, double* __restrict__ Rh,
                                double* __restrict__ Rc, double* __restrict__ Rctx, double* __restrict__ Rchat, int T,
                                int H, int gridtd) {
  const int p = blockIdx.x;
  const int n = w.img[p], t = w.t[p];
  const int k = tok[n * T + t - 1] - 1;
  const double lk = logitk[n * T + t - 1];
  const double coef = lk / stabd(lk);
  const size_t st = ((size_t)n * (T + 1) + t) * H;
  const double b = beta[(size_t)n * (T + 1) + t];
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const double hv = hlast[st + j], cv = chat[st + j];
    const double hc = hv + cv;
    const double r_hc = hc * WoT[(size_t)k * H + j] * coef;
    Rh[(size_t)p * H + j] = hv * r_hc / stabd(hc);
    const double r_chat = cv * r_hc / stabd(hc);
    if (gridtd) {
      Rchat[(size_t)p * H + j] = r_chat;
    } else {
      Rctx[(size_t)p * H + j] = (1.0 - b) * ctx[st + j] * r_chat / stabd(cv);
      Rc[(size_t)p * H + j] = b * s[st + j] * r_chat / stabd(cv);
    }
  }
}
// LSTM cell, gate pass-through (explainers.py:605-619): Rc += Rh (+extra); U = R_g / stab(zg); Rc <- forget branch
__global__ void lrp_cell_kernel(WordRef w, int i, const double* __restrict__ ia, const double* __restrict__ fa,
                                const double* __restrict__ zg, const double* __restrict__ c, double* __restrict__ Rc,
                                const double* __restrict__ Rh, const double* __restrict__ extra,
                                double* __restrict__ U, int T, int H, __nv_bfloat16* __restrict__ Us = nullptr,
                                size_t nUs = 0) {
  // Us != null: U goes straight to the two bf16 planes (hi at Us, lo at Us + nUs) the gate GEMM reads as its A operand
  const int p = blockIdx.x;
  const int n = w.img[p];
  const size_t s0 = ((size_t)n * (T + 1) + i) * H, s1 = s0 + H;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const size_t q = (size_t)p * H + j;
    const double rc = Rc[q] + Rh[q] + (extra ? extra[q] : 0.0);
    const double den = stabd(c[s1 + j]);
    const double rg = ia[s1 + j] * tanh(zg[s1 + j]) * rc / den;
    Rc[q] = fa[s1 + j] * c[s0 + j] * rc / den;
    const double u = rg / stabd(zg[s1 + j]);
    if (Us) {
      __nv_bfloat16 hi, lo;
      split_bf16((float)u, hi, lo);
      Us[q] = hi;
      Us[nUs + q] = lo;
    } else {
      U[q] = u;
    }
  }
}
// adaptive: R_xh = [x_i, h_i] * Y -> word / global / hidden parts (explainers.py:620-630)
template <typename YT>
__global__ void __launch_bounds__(256)
lrp_scatter_adaptive_kernel(WordRef w, int i, const double* __restrict__ XH, const YT* __restrict__ Y,
                            double* __restrict__ Rh, double* __restrict__ Rglob, double* __restrict__ rword, int T,
                            int H, int E) {
  const int p = blockIdx.x;
  const int n = w.img[p];
  const int Kin = 2 * E + H;
  const double* x = XH + ((size_t)n * T + i) * Kin;
  const YT* y = Y + (size_t)p * Kin;
  double wsum = 0.0;
  for (int j = threadIdx.x; j < Kin; j += 256) {
    const double v = x[j] * (double)y[j];
    if (j < E) wsum += v;
    else if (j < 2 * E) Rglob[(size_t)p * E + j - E] += v;
    else Rh[(size_t)p * H + j - 2 * E] = v;
  }
  __shared__ double red[256];
  red[threadIdx.x] = wsum;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) rword[(size_t)p * T + i] = red[0];
}
// grid-TD, language LSTM input split + sentinel/context split (explainers.py:1252-1269)
//   R_x2 = [c_hat_{i+1}, h1_{i+1}, h2_i] * Y2 ; Rchat_i = (i == t-1 ? init : 0) + R_x2[:H] ; Rh1 += R_x2[H:2H] ;
//   Rh2n = R_x2[2H:] ; extra = r_s + Rh1 (added to Rc1 by the next cell kernel) ; Q = R_ctx / stab(ctx_{i+1})
template <typename YT>
__global__ void lrp_scatter_lang_kernel(WordRef w, int i, const double* __restrict__ XH2, const YT* __restrict__ Y2,
                                        const double* __restrict__ chat, const double* __restrict__ ctx,
                                        const double* __restrict__ s, const double* __restrict__ beta,
                                        const double* __restrict__ Rchat_init, double* __restrict__ Rh1,
                                        double* __restrict__ Rh2n, double* __restrict__ extra, double* __restrict__ Q,
                                        int T, int H) {
  const int p = blockIdx.x;
  const int n = w.img[p], t = w.t[p];
  const double* x = XH2 + ((size_t)n * T + i) * 3 * H;
  const YT* y = Y2 + (size_t)p * 3 * H;
  const size_t s1 = ((size_t)n * (T + 1) + i + 1) * H;
  const double b = beta[(size_t)n * (T + 1) + i + 1];
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const size_t q = (size_t)p * H + j;
    const double rchat = ((i == t - 1) ? Rchat_init[q] : 0.0) + x[j] * (double)y[j];
    const double rh1 = Rh1[q] + x[H + j] * (double)y[H + j];
    Rh2n[q] = x[2 * H + j] * (double)y[2 * H + j];
    const double den = stabd(chat[s1 + j]);
    const double r_s = b * s[s1 + j] * rchat / den;
    const double r_ctx = ctx[s1 + j] * (1.0 - b) * rchat / den;
    extra[q] = r_s + rh1;
    Q[((size_t)p * T + i) * H + j] = r_ctx / stabd(ctx[s1 + j]);
  }
}
// grid-TD, top-down LSTM input split (explainers.py:1288-1300): [h2_i, g, emb, h1_i]
template <typename YT>
__global__ void __launch_bounds__(256)
lrp_scatter_td_kernel(WordRef w, int i, const double* __restrict__ XH1, const YT* __restrict__ Y1,
                      const double* __restrict__ Rh2n, double* __restrict__ Rh2, double* __restrict__ Rh1,
                      double* __restrict__ Rglob, double* __restrict__ rword, int T, int H, int E) {
  const int p = blockIdx.x;
  const int n = w.img[p];
  const int Kin = 2 * H + 2 * E;
  const double* x = XH1 + ((size_t)n * T + i) * Kin;
  const YT* y = Y1 + (size_t)p * Kin;
  double wsum = 0.0;
  for (int j = threadIdx.x; j < Kin; j += 256) {
    const double v = x[j] * (double)y[j];
    if (j < H) Rh2[(size_t)p * H + j] = Rh2n[(size_t)p * H + j] + v;
    else if (j < H + E) Rglob[(size_t)p * E + j - H] += v;
    else if (j < H + 2 * E) wsum += v;
    else Rh1[(size_t)p * H + j - H - 2 * E] = v;
  }
  __shared__ double red[256];
  red[threadIdx.x] = wsum;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) rword[(size_t)p * T + i] = red[0];
}

// ------------------------------------------------------------------ relevance: tail (global feature, image features)
__global__ void uglob_kernel(WordRef w, const double* __restrict__ Rglob, const double* __restrict__ gp,
                             double* __restrict__ Ug, int E) {
  const int p = blockIdx.x;
  const int n = w.img[p];
  for (int j = threadIdx.x; j < E; j += blockDim.x) Ug[(size_t)p * E + j] = Rglob[(size_t)p * E + j] / stabd(gp[(size_t)n * E + j]);
}
// ra[p,d] = (a * Ya) / stab(a)   -- r_average_img_feature already divided for the mean-pool split (explainers.py:634-647)
__global__ void ra_kernel(WordRef w, const double* __restrict__ Ya, const double* __restrict__ a, double* __restrict__ ra,
                          int D) {
  const int p = blockIdx.x;
  const int n = w.img[p];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const double av = a[(size_t)n * D + d];
    ra[(size_t)p * D + d] = av * Ya[(size_t)p * D + d] / stabd(av);
  }
}
// adaptive: UV[p,l,:] = float32(Vf_l * alpha_t[l] * R_ctx / stab(ctx_t)) / stab(Vp_l)     (explainers.py:648-659)
__global__ void uv_adaptive_kernel(WordRef w, int p0, const double* __restrict__ Vp, const double* __restrict__ alpha,
                                   const double* __restrict__ ctx, const double* __restrict__ Rctx,
                                   double* __restrict__ UV, int T, int L, int H, __nv_bfloat16* __restrict__ UVs = nullptr,
                                   size_t nUVs = 0) {
  // UVs != null: the rows go straight to the two bf16 planes (hi at UVs, lo at UVs + nUVs) of the image_features GEMM
  const int p = p0 + blockIdx.y, l = blockIdx.x;
  const int n = w.img[p], t = w.t[p];
  const double al = alpha[((size_t)n * (T + 1) + t) * L + l];
  const size_t st = ((size_t)n * (T + 1) + t) * H;
  const double* vp = Vp + ((size_t)n * L + l) * H;
  const size_t o0 = ((size_t)blockIdx.y * L + l) * H;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const double rv = f32r(fmax(vp[j], 0.0) * al * Rctx[(size_t)p * H + j] / stabd(ctx[st + j]));
    const double u = rv / stabd(vp[j]);
    if (UVs) {
      __nv_bfloat16 hi, lo;
      split_bf16((float)u, hi, lo);
      UVs[o0 + j] = hi;
      UVs[nUVs + o0 + j] = lo;
    } else {
      UV[o0 + j] = u;
    }
  }
}
// grid-TD: r_V accumulates over every step <= t in a float32 buffer (explainers.py:1292-1299)
__global__ void uv_gridtd_kernel(WordRef w, int p0, const double* __restrict__ Vp, const double* __restrict__ alpha,
                                 const double* __restrict__ Q, double* __restrict__ UV, int T, int L, int H) {
  const int p = p0 + blockIdx.y, l = blockIdx.x;
  const int n = w.img[p], t = w.t[p];
  const double* vp = Vp + ((size_t)n * L + l) * H;
  double* o = UV + ((size_t)blockIdx.y * L + l) * H;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const double vf = fmax(vp[j], 0.0);
    float acc = 0.f;
    for (int i = t - 1; i >= 0; --i) {
      const double al = alpha[((size_t)n * (T + 1) + i + 1) * L + l];
      acc = (float)((double)acc + vf * al * Q[((size_t)p * T + i) * H + j]);
    }
    o[j] = (double)acc / stabd(vp[j]);
  }
}
// Same arithmetic, one block per (word, group of LG locations): the word's Q rows (T x H, shared by all L locations) are
// held in registers instead of being re-read from L2 for every location.
constexpr int kUvMaxT = 24;   // the switch below enumerates steps 23 .. 0
__global__ void __launch_bounds__(256, 3)
uv_gridtd_rows_kernel(WordRef w, int p0, const double* __restrict__ Vp, const double* __restrict__ alpha,
                      const double* __restrict__ Q, double* __restrict__ UV, int T, int L, int H, int LG,
                      __nv_bfloat16* __restrict__ UVs = nullptr, size_t nUVs = 0) {
  const int p = p0 + blockIdx.y;
  const int n = w.img[p], t = w.t[p];
  const int l0 = blockIdx.x * LG, l1 = (l0 + LG < L) ? l0 + LG : L;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    double q[kUvMaxT];
#pragma unroll
    for (int i = 0; i < kUvMaxT; ++i) q[i] = (i < t) ? Q[((size_t)p * T + i) * H + j] : 0.0;
    for (int l = l0; l < l1; ++l) {
      const double v = Vp[((size_t)n * L + l) * H + j];
      const double vf = fmax(v, 0.0);
      const double* al = alpha + ((size_t)n * (T + 1) + 1) * L + l;
      float acc = 0.f;
      // steps t-1 .. 0 in the reference's order; the fall-through switch enters the unrolled sequence at step t-1 so that
      // q[] keeps static register indices and no skipped step is issued
#define LRPCAP_UV_STEP(I) case (I) + 1: acc = (float)((double)acc + vf * al[(size_t)(I) * L] * q[I]);
      switch (t) {
        LRPCAP_UV_STEP(23) LRPCAP_UV_STEP(22) LRPCAP_UV_STEP(21) LRPCAP_UV_STEP(20) LRPCAP_UV_STEP(19) LRPCAP_UV_STEP(18)
        LRPCAP_UV_STEP(17) LRPCAP_UV_STEP(16) LRPCAP_UV_STEP(15) LRPCAP_UV_STEP(14) LRPCAP_UV_STEP(13) LRPCAP_UV_STEP(12)
        LRPCAP_UV_STEP(11) LRPCAP_UV_STEP(10) LRPCAP_UV_STEP(9) LRPCAP_UV_STEP(8) LRPCAP_UV_STEP(7) LRPCAP_UV_STEP(6)
        LRPCAP_UV_STEP(5) LRPCAP_UV_STEP(4) LRPCAP_UV_STEP(3) LRPCAP_UV_STEP(2) LRPCAP_UV_STEP(1) LRPCAP_UV_STEP(0)
        default: break;
      }
#undef LRPCAP_UV_STEP
      const double u = (double)acc / stabd(v);
      const size_t o = ((size_t)blockIdx.y * L + l) * H + j;
      if (UVs) {
        __nv_bfloat16 hi, lo;
        split_bf16((float)u, hi, lo);
        UVs[o] = hi;
        UVs[nUVs + o] = lo;
      } else {
        UV[o] = u;
      }
    }
  }
}
// The same redistribution in float32 arithmetic: r_V[l] = Vf_l * sum_i alpha_{i+1}[l] * Q_i accumulated with fp32 FMAs
// (the reference accumulates fp64 products into a float32 buffer; both carry ~1e-7 of rounding, six times less work on the
// fp64 / conversion pipes: 840 -> ~100 us per 512 words).  Always writes the GEMM's bf16 planes.
__global__ void __launch_bounds__(256, 3)
uv_gridtd_rows_f32_kernel(WordRef w, int p0, const double* __restrict__ Vp, const double* __restrict__ alpha,
                          const double* __restrict__ Q, int T, int L, int H, int LG, __nv_bfloat16* __restrict__ UVs,
                          size_t nUVs) {
  const int p = p0 + blockIdx.y;
  const int n = w.img[p], t = w.t[p];
  const int l0 = blockIdx.x * LG, l1 = (l0 + LG < L) ? l0 + LG : L;
  extern __shared__ float al_s[];                      // [LG][kUvMaxT]: alpha_{i+1}[l] of this block's locations
  for (int k = threadIdx.x; k < (l1 - l0) * kUvMaxT; k += blockDim.x) {
    const int ll = k / kUvMaxT, i = k - ll * kUvMaxT;
    al_s[k] = i < t ? (float)alpha[((size_t)n * (T + 1) + i + 1) * L + l0 + ll] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    float q[kUvMaxT];
#pragma unroll
    for (int i = 0; i < kUvMaxT; ++i) q[i] = (i < t) ? (float)Q[((size_t)p * T + i) * H + j] : 0.f;
    for (int l = l0; l < l1; ++l) {
      const double v = Vp[((size_t)n * L + l) * H + j];
      const float vf = (float)fmax(v, 0.0);
      const float* al = al_s + (l - l0) * kUvMaxT;
      float acc = 0.f;
#pragma unroll
      for (int i = kUvMaxT - 1; i >= 0; --i) acc = fmaf(vf * al[i], q[i], acc);   // steps beyond t contribute exact zeros
      const float u = (float)((double)acc / stabd(v));
      const size_t o = ((size_t)blockIdx.y * L + l) * H + j;
      __nv_bfloat16 hi, lo;
      split_bf16(u, hi, lo);
      UVs[o] = hi;
      UVs[nUVs + o] = lo;
    }
  }
}

// R_F[word, l, d] = float32( float32(F/L * ra) + F * YF )      (explainers.py:641-659)
// YT = float: YF is the tensor-core GEMM's fp32 result as is (no fp32 -> fp64 pass in between).
// FT = float: the caller's float32 features as they came in (the fp64 copy F holds exactly those values).
template <typename YT, typename FT = double>
__global__ void final_kernel(WordRef w, int p0, const int* __restrict__ order, const FT* __restrict__ F,
                             const double* __restrict__ ra, const YT* __restrict__ YF, float* __restrict__ out, int L,
                             int D) {
  const int p = p0 + blockIdx.y, l = blockIdx.x;
  const int n = w.img[p];
  const FT* f = F + ((size_t)n * L + l) * D;
  const YT* yf = YF + ((size_t)blockIdx.y * L + l) * D;
  float* o = out + ((size_t)order[p] * L + l) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const double fv = (double)f[d];
    const float first = (float)(fv / L * ra[(size_t)p * D + d]);
    o[d] = (float)((double)first + fv * (double)yf[d]);
  }
}

// ------------------------------------------------------------------ frozen-attention gradient decoder
// (explainers.py:780-832 adaptive, :1452-1532 grid-TD).  The reference keeps d_h / d_c / gate gradients in float32
// arrays; f32r() reproduces those stores.

// d_h[t] = d_(h+c_hat) = W_o[:, k]; grid-TD also seeds d_c_hat[t-1] with it.
__global__ void grad_init_kernel(WordRef w, const double* __restrict__ WoT, const int* __restrict__ tok,
                                 double* __restrict__ dh, double* __restrict__ dchat_init, int T, int H) {
  const int p = blockIdx.x;
  const int n = w.img[p], t = w.t[p];
  const int k = tok[n * T + t - 1] - 1;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const double v = WoT[(size_t)k * H + j];
    dh[(size_t)p * H + j] = f32r(v);
    if (dchat_init) dchat_init[(size_t)p * H + j] = v;
  }
}
// One LSTM BPTT step (explainers.py:811-821): gates = [d_i, d_f, d_g, d_o] pre-activation gradients, dc <- d_c[i].
__global__ void grad_cell_kernel(WordRef w, int i, const double* __restrict__ ia, const double* __restrict__ fa,
                                 const double* __restrict__ ga, const double* __restrict__ oa,
                                 const double* __restrict__ c, const double* __restrict__ dh, const double* __restrict__ dh_add,
                                 double* __restrict__ dc, double* __restrict__ gates, int T, int H) {
  const int p = blockIdx.x;
  const int n = w.img[p];
  const size_t s0 = ((size_t)n * (T + 1) + i) * H, s1 = s0 + H;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const size_t q = (size_t)p * H + j;
    const double dhv = dh_add ? f32r(dh[q] + dh_add[q]) : dh[q];
    const double th = tanh(c[s1 + j]);
    const double i_ = ia[s1 + j], f_ = fa[s1 + j], g_ = ga[s1 + j], o_ = oa[s1 + j];
    const double d_oa = f32r(dhv * th);
    const double dcn = f32r(dc[q] + dhv * o_ * (1.0 - th * th));
    const double d_fa = f32r(dcn * c[s0 + j]);
    const double d_ia = f32r(dcn * g_);
    const double d_ga = f32r(dcn * i_);
    dc[q] = f32r(dcn * f_);
    double* gt = gates + (size_t)p * 4 * H;
    gt[j] = f32r(d_ia * i_ * (1.0 - i_));
    gt[H + j] = f32r(d_fa * f_ * (1.0 - f_));
    gt[2 * H + j] = f32r(d_ga * (1.0 - g_ * g_));
    gt[3 * H + j] = f32r(d_oa * o_ * (1.0 - o_));
  }
}
// adaptive: Yg = gates [W_i ; W_h]^T -> d_x (float32 store) and d_h (explainers.py:822-825)
__global__ void __launch_bounds__(256)
grad_scatter_adaptive_kernel(int i, const double* __restrict__ Yg, double* __restrict__ dh, double* __restrict__ dglob,
                             double* __restrict__ dwords, int T, int H, int E) {
  const int p = blockIdx.x;
  const int Kin = 2 * E + H;
  const double* y = Yg + (size_t)p * Kin;
  double wsum = 0.0;
  for (int j = threadIdx.x; j < Kin; j += 256) {
    const double v = f32r(y[j]);
    if (j < E) wsum += v;
    else if (j < 2 * E) dglob[(size_t)p * E + j - E] += v;
    else dh[(size_t)p * H + j - 2 * E] = v;
  }
  __shared__ double red[256];
  red[threadIdx.x] = wsum;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) dwords[(size_t)p * T + i] = red[0];
}
// grid-TD language LSTM input split (explainers.py:1501-1505): Y2 = gates2 [W_i2 ; W_h2]^T over [c_hat, h1, h2]
//   d_c_hat[i] = (i == t-1 ? seed : 0) + Y2[:H]; d_ctx = d_c_hat (1 - beta_{i+1}) -> Q; dh1_add = Y2[H:2H]; dh2 = f32(Y2[2H:])
__global__ void grad_scatter_lang_kernel(WordRef w, int i, const double* __restrict__ Y2, const double* __restrict__ beta,
                                         const double* __restrict__ dchat_init, double* __restrict__ dh1_add,
                                         double* __restrict__ dh2, double* __restrict__ Q, int T, int H) {
  const int p = blockIdx.x;
  const int n = w.img[p], t = w.t[p];
  const double* y = Y2 + (size_t)p * 3 * H;
  const double b = beta[(size_t)n * (T + 1) + i + 1];
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const size_t q = (size_t)p * H + j;
    const double dchat = ((i == t - 1) ? dchat_init[q] : 0.0) + y[j];
    Q[((size_t)p * T + i) * H + j] = dchat * (1.0 - b);
    dh1_add[q] = y[H + j];
    dh2[q] = f32r(y[2 * H + j]);
  }
}
// grid-TD top-down LSTM input split (explainers.py:1517-1523): Y1 over [h2, g, emb, h1]
__global__ void __launch_bounds__(256)
grad_scatter_td_kernel(int i, const double* __restrict__ Y1, double* __restrict__ dh2, double* __restrict__ dh1,
                       double* __restrict__ dglob, double* __restrict__ dwords, int T, int H, int E) {
  const int p = blockIdx.x;
  const int Kin = 2 * H + 2 * E;
  const double* y = Y1 + (size_t)p * Kin;
  double wsum = 0.0;
  for (int j = threadIdx.x; j < Kin; j += 256) {
    const double v = y[j];
    if (j < H) dh2[(size_t)p * H + j] = f32r(dh2[(size_t)p * H + j] + v);
    else if (j < H + E) dglob[(size_t)p * E + j - H] += v;
    else if (j < H + 2 * E) wsum += v;
    else dh1[(size_t)p * H + j - H - 2 * E] = f32r(v);
  }
  __shared__ double red[256];
  red[threadIdx.x] = wsum;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) dwords[(size_t)p * T + i] = red[0];
}
// ReLU mask of the global feature: adaptive tests element 0 only (quirk B3, explainers.py:826); grid-TD element-wise (:1524)
__global__ void grad_glob_mask_kernel(WordRef w, double* __restrict__ dglob, const double* __restrict__ gp, int E,
                                      int scalar_test) {
  const int p = blockIdx.x;
  const int n = w.img[p];
  for (int j = threadIdx.x; j < E; j += blockDim.x) {
    const double ref = scalar_test ? gp[(size_t)n * E] : gp[(size_t)n * E + j];
    if (ref <= 0.0) dglob[(size_t)p * E + j] = 0.0;
  }
}
// d_V rows as the A operand of the image_features GEMM, masked where relu(Vp) <= 0
__global__ void gv_adaptive_kernel(WordRef w, int p0, const double* __restrict__ Vp, const double* __restrict__ alpha,
                                   const double* __restrict__ WoT, const int* __restrict__ tok, double* __restrict__ UV,
                                   int T, int L, int H) {
  const int p = p0 + blockIdx.y, l = blockIdx.x;
  const int n = w.img[p], t = w.t[p];
  const int k = tok[n * T + t - 1] - 1;
  const double al = alpha[((size_t)n * (T + 1) + t) * L + l];
  const double* vp = Vp + ((size_t)n * L + l) * H;
  double* o = UV + ((size_t)blockIdx.y * L + l) * H;
  for (int j = threadIdx.x; j < H; j += blockDim.x) o[j] = vp[j] > 0.0 ? f32r(WoT[(size_t)k * H + j] * al) : 0.0;
}
__global__ void gv_gridtd_kernel(WordRef w, int p0, const double* __restrict__ Vp, const double* __restrict__ alpha,
                                 const double* __restrict__ Q, double* __restrict__ UV, int T, int L, int H) {
  const int p = p0 + blockIdx.y, l = blockIdx.x;
  const int n = w.img[p], t = w.t[p];
  const double* vp = Vp + ((size_t)n * L + l) * H;
  double* o = UV + ((size_t)blockIdx.y * L + l) * H;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    double acc = 0.0;
    for (int i = t - 1; i >= 0; --i) acc += Q[((size_t)p * T + i) * H + j] * alpha[((size_t)n * (T + 1) + i + 1) * L + l];
    o[j] = vp[j] > 0.0 ? acc : 0.0;
  }
}
// d_F[word, l, :] = float32(d_a / L) (+) float32(d_V[l] W_if^T)   (explainers.py:828-830 / 1528-1530)
__global__ void grad_final_kernel(int p0, const int* __restrict__ order, const double* __restrict__ da,
                                  const double* __restrict__ YF, float* __restrict__ out, int L, int D) {
  const int p = p0 + blockIdx.y, l = blockIdx.x;
  const double* yf = YF + ((size_t)blockIdx.y * L + l) * D;
  float* o = out + ((size_t)order[p] * L + l) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    o[d] = (float)((double)(float)(da[(size_t)p * D + d] / L) + (double)(float)yf[d]);
}

}  // namespace dk
}  // namespace lrpcap
