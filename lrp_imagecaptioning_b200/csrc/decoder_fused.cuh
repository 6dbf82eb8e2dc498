// Fused kernels of the decoder forward's fast path (decoder.cu: Decoder::forward, `fused_`): every point-wise stage reads
// the fp32 result of the tensor-core GEMM before it directly (no fp32 -> fp64 pass), adds the bias in fp64, and writes,
// next to the fp64 state the relevance pass reads, the three bf16 planes of the NEXT GEMM's A operand (no separate
// conversion pass).  One decoder step is 9 launches instead of 25 (grid-TD, greedy).
// Reference arithmetic: models/explainers.py:400-421 (adaptive step), :1131-1156 (grid-TD step), :125-139 (LSTM cell).
// Included only by decoder.cu, after decoder_kernels.cuh.
#pragma once
#include "decoder_kernels.cuh"

namespace lrpcap {
namespace dk {

// value -> the three bf16 planes (hi, mid, lo: 24 bits, fp32-exact) of a GEMM operand with `n` elements per plane
__device__ __forceinline__ void put_split3(__nv_bfloat16* __restrict__ A, size_t n, size_t i, double x) {
  float v = (float)x;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    A[(size_t)p * n + i] = h;
    v -= __bfloat162float(h);
  }
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// XH1 row of step i (fp64, kept for the relevance pass) + the A operand of the gate GEMM.  Same layout as build_xh_kernel.
__global__ void __launch_bounds__(256)
fwd_xh_kernel(double* __restrict__ XH, __nv_bfloat16* __restrict__ A, size_t nA, const double* __restrict__ Emb,
              const double* __restrict__ gp, const double* __restrict__ h1, const double* __restrict__ h2,
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
    put_split3(A, nA, (size_t)n * Kin + j, v);
  }
}

// LSTM cell (Keras gate order i, f, c, o) from the gate GEMM's fp32 result Z [*, ldz] (+ bias), and -- when `s` is given --
// the sentinel s = tanh(c') * sigmoid(sg) with sg = Z[:, 4H + j] (the [W_x ; W_h] columns ride in the same GEMM).
// Ahs (optional): planes of [h' ; s'] (rows n and N + n) for the [W_hp | W_ss] GEMM.
// hc (optional, grid-TD language LSTM): hc = h' (+ chat') -> fp64 + planes Ahc for the logit GEMM.
__global__ void __launch_bounds__(256)
fwd_lstm_kernel(const float* __restrict__ Z, int ldz, const double* __restrict__ bias, double* __restrict__ h,
                double* __restrict__ c, double* __restrict__ zg, double* __restrict__ ia, double* __restrict__ fa,
                double* __restrict__ ga, double* __restrict__ oa, double* __restrict__ s, __nv_bfloat16* __restrict__ Ahs,
                size_t nAhs, int N, const double* __restrict__ chat, double* __restrict__ hc,
                __nv_bfloat16* __restrict__ Ahc, size_t nAhc, int add_chat, int i, int T, int H) {
  const int n = blockIdx.x;
  const size_t s0 = ((size_t)n * (T + 1) + i) * H, s1 = s0 + H;
  const float* z = Z + (size_t)n * ldz;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const double zi = (double)z[j] + bias[j], zf = (double)z[H + j] + bias[H + j];
    const double zc = (double)z[2 * H + j] + bias[2 * H + j], zo = (double)z[3 * H + j] + bias[3 * H + j];
    const double i_ = sigm(zi), f_ = sigm(zf), g_ = tanh(zc), o_ = sigm(zo);
    const double cn = f_ * c[s0 + j] + i_ * g_;
    const double th = tanh(cn);
    const double hn = o_ * th;
    c[s1 + j] = cn;
    h[s1 + j] = hn;
    zg[s1 + j] = zc;
    ia[s1 + j] = i_;
    fa[s1 + j] = f_;
    ga[s1 + j] = g_;
    oa[s1 + j] = o_;
    if (s) {
      const double sn = th * sigm((double)z[4 * H + j]);
      s[s1 + j] = sn;
      put_split3(Ahs, nAhs, (size_t)n * H + j, hn);
      put_split3(Ahs, nAhs, (size_t)(N + n) * H + j, sn);
    }
    if (hc) {
      const double v = hn + (add_chat ? chat[s1 + j] : 0.0);
      hc[(size_t)n * H + j] = v;
      put_split3(Ahc, nAhc, (size_t)n * H + j, v);
    }
  }
}

// attention scores, one warp per (image, location): e[n,l] = tanh(P[n,l] + hp[n]) . V ; e[n,L] = tanh(sp[n] + hp[n]) . V
// hp = HS[n, 0:H], sp = HS[N + n, H:2H] (fp32 result of the [h' ; s'] x [W_hp | W_ss] GEMM).
__global__ void __launch_bounds__(256)
fwd_scores_kernel(const double* __restrict__ P, const float* __restrict__ HS, int ldhs, int N,
                  const double* __restrict__ Va, double* __restrict__ e, int L, int H) {
  const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int l = blockIdx.x * 8 + warp;
  if (l > L) return;
  const float* hp = HS + (size_t)n * ldhs;
  const float* sp = HS + (size_t)(N + n) * ldhs + H;
  double acc = 0.0;
  if (l < L) {
    const double2* base = reinterpret_cast<const double2*>(P + ((size_t)n * L + l) * H);
    for (int j2 = lane; j2 < H / 2; j2 += 32) {
      const double2 pv = base[j2];
      const float2 hv = reinterpret_cast<const float2*>(hp)[j2];
      const double2 va = reinterpret_cast<const double2*>(Va)[j2];
      acc += tanh(pv.x + (double)hv.x) * va.x + tanh(pv.y + (double)hv.y) * va.y;
    }
  } else {
    for (int j = lane; j < H; j += 32) acc += tanh((double)sp[j] + (double)hp[j]) * Va[j];
  }
  acc = warp_sum(acc);
  if (lane == 0) e[(size_t)n * (L + 1) + l] = acc;
}

// softmax over the L locations (+ sentinel gate beta), context = sum_l alpha_l Vf_l, c_hat = beta s + (1 - beta) ctx, and
// what the next GEMM consumes.  grid = (N, H / 64); 8 warps split the L locations, a lane owns two adjacent channels.
//   adaptive: hc = h1' + c_hat -> fp64 + planes Ahc (logit GEMM)
//   grid-TD : XH2 row [c_hat', h1', h2_i] -> fp64 + planes Axh2 (language-LSTM gate GEMM)
// Vf32: float copy of relu(Vp) (exact: Vp holds float32 values, explainers.py:383-386).
__global__ void __launch_bounds__(256)
fwd_ctx_kernel(const float* __restrict__ Vf32, const double* __restrict__ e, double* __restrict__ alpha,
               double* __restrict__ beta, const double* __restrict__ s, double* __restrict__ ctx, double* __restrict__ chat,
               const double* __restrict__ h1, const double* __restrict__ h2, double* __restrict__ hc,
               __nv_bfloat16* __restrict__ Ahc, size_t nAhc, double* __restrict__ XH2, __nv_bfloat16* __restrict__ Axh2,
               size_t nAxh2, int i, int T, int L, int H) {
  __shared__ double al[256];
  __shared__ double red[8];
  __shared__ double part[8][64];
  const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double* en = e + (size_t)n * (L + 1);
  // softmax (L <= 256: one location per thread)
  const double ev = tid < L ? en[tid] : -1e300;
  double m = warp_max(ev);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  double m1 = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m1 = fmax(m1, red[w]);
  __syncthreads();
  const double ex = tid < L ? exp(ev - m1) : 0.0;
  double sm = warp_sum(ex);
  if (lane == 0) red[warp] = sm;
  __syncthreads();
  double s1sum = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) s1sum += red[w];
  const double a_l = ex / s1sum;
  al[tid] = a_l;
  const double eL = en[L];
  const double m2 = fmax(m1, eL);
  const double b = exp(eL - m2) / (s1sum * exp(m1 - m2) + exp(eL - m2));
  if (blockIdx.y == 0) {
    if (tid < L) alpha[((size_t)n * (T + 1) + i + 1) * L + tid] = a_l;
    if (tid == 0) beta[(size_t)n * (T + 1) + i + 1] = b;
  }
  __syncthreads();
  // context: channels [64 blockIdx.y, +64)
  const int j0 = blockIdx.y * 64 + 2 * lane;
  double ax = 0.0, ay = 0.0;
  const float* vf = Vf32 + (size_t)n * L * H + j0;
  for (int l = warp; l < L; l += 8) {
    const float2 v = *reinterpret_cast<const float2*>(vf + (size_t)l * H);
    const double a = al[l];
    ax += a * (double)v.x;
    ay += a * (double)v.y;
  }
  part[warp][2 * lane] = ax;
  part[warp][2 * lane + 1] = ay;
  __syncthreads();
  if (tid < 64) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) acc += part[w][tid];
    const int j = blockIdx.y * 64 + tid;
    const size_t s0 = ((size_t)n * (T + 1) + i) * H, s1 = s0 + H;
    ctx[s1 + j] = acc;
    const double ch = b * s[s1 + j] + (1.0 - b) * acc;
    chat[s1 + j] = ch;
    if (hc) {
      const double v = h1[s1 + j] + ch;
      hc[(size_t)n * H + j] = v;
      put_split3(Ahc, nAhc, (size_t)n * H + j, v);
    }
    if (XH2) {
      double* x = XH2 + ((size_t)n * T + i) * 3 * H;
      const double v1 = h1[s1 + j], v2 = h2[s0 + j];
      x[j] = ch;
      x[H + j] = v1;
      x[2 * H + j] = v2;
      const size_t r = (size_t)n * 3 * H;
      put_split3(Axh2, nAxh2, r + j, ch);
      put_split3(Axh2, nAxh2, r + H + j, v1);
      put_split3(Axh2, nAxh2, r + 2 * H + j, v2);
    }
  }
}

// greedy: token = arg-max over (fp32 logit + fp64 bias) (first maximum), optionally never the EOS index
__global__ void __launch_bounds__(256)
fwd_argmax_kernel(const float* __restrict__ C, int ldc, const double* __restrict__ bo, int V, int eos_index,
                  int* __restrict__ tok, double* __restrict__ logitk, int i, int T) {
  const int n = blockIdx.x;
  const float* ln = C + (size_t)n * ldc;
  double best = -1e300;
  int bi = V;
  for (int v = threadIdx.x; v < V; v += 256) {
    if (v == eos_index) continue;
    const double x = (double)ln[v] + bo[v];
    if (x > best) { best = x; bi = v; }
  }
  // warp-level (value, index) arg-max, then across the 8 warps
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  __shared__ double rb[8];
  __shared__ int ri[8];
  if ((threadIdx.x & 31) == 0) { rb[threadIdx.x >> 5] = best; ri[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (rb[w] > best || (rb[w] == best && ri[w] < bi)) { best = rb[w]; bi = ri[w]; }
    tok[n * T + i] = bi + 1;   // model index -> tokenizer id (explainers.py:92)
    logitk[n * T + i] = best;
  }
}

__global__ void relu_f32_copy_kernel(const double* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)fmax(in[i], 0.0);
}

}  // namespace dk
}  // namespace lrpcap
