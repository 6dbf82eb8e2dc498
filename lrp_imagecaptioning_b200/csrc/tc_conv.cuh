// tcgen05 implicit-GEMM 3x3 / 1x1 convolution with split-bf16 ("bf16x3") operands -- interface.
//
// One kernel serves every dense contraction on the hot path:
//   D[pixel, n] = sum_{tap, c} A[item, y+dy(tap), x+dx(tap), c] * B[tap, n, c]
// A: activations / relevance messages, NHWC, split-bf16 planes.   B: prepared weights [taps*Nout, C].
// Products: A_hi*B_hi + A_hi*B_lo + A_lo*B_hi, fp32 accumulation in TMEM.
#pragma once
#include "common.cuh"

namespace lrpcap {

enum EpiMode : int {
  EPI_FWD_TRUE = 0,   // z = acc + b ; x = relu(z) ; write x (split [+fp32]) ; optional G / seed multiplier
  EPI_FWD_ZACT = 1,   // zact = acc (+ b) ; G = x / safe(zact) with x read back from the true forward
  EPI_BWD = 2,        // s_prev = (relu?)(acc) * G[img(item)] ; optional 2x2 up-sampling (fused max-pool routing)
  EPI_RAW = 3,        // out_f32 = acc (debug / generic GEMM)
};

enum GMode : int {
  G_NONE = 0,
  G_EPS = 1,    // G = relu(z) / (z + sgn+(z) eps)      seed multiplier 1 / (z + sgn+(z) eps)
  G_Z = 2,      // G = relu(z) / safe(z)                seed multiplier 1 / safe(z)
  G_MASK = 3,   // G = [z > 0]                          seed multiplier [z > 0]      (gradient family)
};

struct EpiParams {
  int mode = EPI_RAW;
  // forward
  const float* bias = nullptr;      // [Nout] (may be null)
  void* out_act = nullptr;          // split storage [items, H, W, Nout]
  size_t out_act_elems = 0;
  float* out_f32 = nullptr;         // optional fp32 copy [items, H, W, Nout]
  float* G = nullptr;               // fp32 per-image multiplier, channel-tiled layout (epilogue.cuh: g_offset)
  float acc_scale = 1.f;            // forward: accumulator scale (2^-k for weights pre-scaled by 2^k, half planes)
  int* overflow = nullptr;          // forward, half planes: device flag set when an activation >= 32768
  int g_up = 1;                     // 2 when a 2x2 max-pool follows this layer (its G is read by an up-sampling epilogue)
  float* Mseed = nullptr;           // fp32 [items, H, W, Nout] seed multiplier (last layer only)
  int gmode = G_NONE;
  float eps = 0.f;
  int rule_bias = 1;                // 0 for the *IgnoreBias rule variants (the true forward still adds the bias)
  const void* x_act = nullptr;      // FWD_ZACT: split storage of the true activation, same shape as the output
  size_t x_act_elems = 0;
  // backward
  const int* img_index = nullptr;   // [items] -> image
  const float* Gin = nullptr;       // fp32 multiplier of the layer below, [images][Nout/16][H][W][16] on the accumulator
                                    // grid; for up == 2 this is the compact form: the one non-zero of each 2x2 window
  const unsigned* Gidx = nullptr;   // up == 2: [images][Nout/16][H][W] words, bits [2k, 2k+1] = window position
                                    // (sy * 2 + sx) of the arg-max of channel 16 j + k
  const float* Gin2 = nullptr;      // alpha-beta with beta != 0: second multiplier (inhibitor branch); the output then has
                                    // 2*Nout channels: [acc*Gin | acc*Gin2]
  int up = 1;
  int relu_acc = 0;
  void* out_msg = nullptr;          // split storage [items, H*up, W*up, Nout]
  int out_planar_f32 = 0;              // 1: write the message as fp32 [items][NO][H*up][W*up] (read by last_dgrad only)
  size_t out_msg_elems = 0;
  // two-product backward (TcConvArgs::planes == kPlanesH1x2): per-item scale bookkeeping, see epilogue.cuh
  const unsigned* mx_in = nullptr;
  unsigned* mx_out = nullptr;
  const int* kt_in = nullptr;
  int* kt_out = nullptr;
  int target_exp = 4;               // predicted maximum of the outgoing fp16 plane: 2^target_exp
  // logical sizes for the debug build's bounds checks (epilogue.cuh: LRPCAP_DEBUG_BOUNDS); 0 = not checked
  size_t g_elems = 0;               // floats behind G (forward) / Gin, Gin2 (backward)
  size_t aux_elems = 0;             // floats behind out_f32 / Mseed
};

struct TcConvArgs {
  const void* A = nullptr;  // split storage [n_items, H, W, C]
  size_t A_elems = 0;
  int n_items = 0, H = 0, W = 0, C = 0;
  const void* B = nullptr;  // split storage [taps * Nout, C]
  size_t B_elems = 0;
  int taps = 9, Nout = 0;
  int planes = 2;           // 5 (kPlanesH1x2): A = ONE fp16 plane (scaled message), B = two fp16 planes, 2 MMA products (backward /
                            // raw only); 4 (kPlanesF16x2): two IEEE half planes, 3 products, promoted every k-step (forward / raw only);
                            // otherwise bf16 planes of A, B and of the forward epilogue's activation tensors: 2 (hi, lo: 3 MMA
                            // products, 16-bit operands) or 3 (hi, mid, lo: 6 products, fp32-exact operands; forward only)
  int promote_every = 0;    // 2-plane only: > 0 sums partial accumulators in fp32 registers every n k-steps (64 channels
                            // of one tap each); tensor-core accumulation truncates, long chains cost ~1e-5 relative.
                            // 3-plane and half-plane launches always promote (every k-step unless set here).
  EpiParams epi;
};

// Launches on `stream`. Returns a status code (common.cuh).
int tc_conv_launch(const TcConvArgs& args, cudaStream_t stream);

// fp32 SIMT implicit-GEMM convolution with the same epilogues (exact-fp32 precision mode, and the
// 3-channel first layer in both modes). A: fp32 [n_items, H, W, C]; B: fp32 [taps][C][Nout].
struct SimtConvArgs {
  const float* A = nullptr;
  int n_items = 0, H = 0, W = 0, C = 0;
  const float* B = nullptr;
  int taps = 9, Nout = 0;
  int out_planes = 0;       // storage of out_act / out_msg / x_act: 0 = fp32, 2 or 3 = split-bf16 planes
  EpiParams epi;
};
int simt_conv_launch(const SimtConvArgs& args, cudaStream_t stream);

// Tile geometry used for a given map size (exposed for tests / docs).
void tc_conv_tile(int H, int W, int n_items, int* TW, int* TH, int* TI);

}  // namespace lrpcap
