// extern "C" surface declared in include/lrpcap.h: error plumbing, encoder entry points, debug conv.
#include "../../include/lrpcap.h"
#include "handles.cuh"
#include "encoder_kernels.cuh"
#include "tc_conv.cuh"
#include <cmath>
#include <cstdlib>
#include <vector>

namespace lrpcap {
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_err; }
}  // namespace lrpcap

using namespace lrpcap;


extern "C" {

const char* lrpcap_last_error(void) { return get_last_error(); }
int lrpcap_version(void) { return 100; }

int lrpcap_encoder_create(lrpcap_encoder_t** out, const float* const* h_kernels_hwio, const float* const* h_biases,
                          int image_hw, int precision) {
  LRPCAP_REQUIRE(out != nullptr, kErrInvalidArg, "encoder_create: null out");
  Encoder* e = nullptr;
  LRPCAP_TRY(Encoder::create(&e, h_kernels_hwio, h_biases, image_hw, precision));
  *out = new lrpcap_encoder{e};
  return kOk;
}

int lrpcap_encoder_create_arch(lrpcap_encoder_t** out, int arch, const float* const* h_kernels_hwio,
                               const float* const* h_biases, int image_hw, int precision) {
  LRPCAP_REQUIRE(out != nullptr, kErrInvalidArg, "encoder_create_arch: null out");
  Encoder* e = nullptr;
  LRPCAP_TRY(Encoder::create(&e, h_kernels_hwio, h_biases, image_hw, precision, arch));
  *out = new lrpcap_encoder{e};
  return kOk;
}

int lrpcap_encoder_set_weights(lrpcap_encoder_t* enc, const float* const* h_kernels_hwio, const float* const* h_biases) {
  LRPCAP_REQUIRE(enc && enc->impl, kErrInvalidArg, "encoder_set_weights: null handle");
  return enc->impl->set_weights(h_kernels_hwio, h_biases);
}

int lrpcap_encoder_set_weights_device(lrpcap_encoder_t* enc, const float* const* d_kernels_hwio, const float* const* d_biases) {
  LRPCAP_REQUIRE(enc && enc->impl, kErrInvalidArg, "encoder_set_weights_device: null handle");
  return enc->impl->set_weights_device(d_kernels_hwio, d_biases);
}

int lrpcap_encoder_destroy(lrpcap_encoder_t* enc) {
  if (!enc) return kOk;
  delete enc->impl;
  delete enc;
  return kOk;
}

int lrpcap_encoder_forward(lrpcap_encoder_t* enc, const float* d_images, int n_images, int rule, float epsilon,
                           float alpha, float beta, int bias, void* stream) {
  LRPCAP_REQUIRE(enc && enc->impl, kErrInvalidArg, "encoder_forward: null handle");
  EncoderRule r;
  r.kind = rule;
  r.epsilon = epsilon;
  r.alpha = alpha;
  r.beta = beta;
  r.bias = bias;
  return enc->impl->forward(d_images, n_images, r, reinterpret_cast<cudaStream_t>(stream));
}

int lrpcap_encoder_features(lrpcap_encoder_t* enc, float* d_features, void* stream) {
  LRPCAP_REQUIRE(enc && enc->impl && d_features, kErrInvalidArg, "encoder_features: null argument");
  Encoder* e = enc->impl;
  LRPCAP_REQUIRE(e->n_images() > 0, kErrState, "encoder_features: call encoder_forward first");
  const size_t n = (size_t)e->n_images() * e->feature_hw() * e->feature_hw() * 512;
  LRPCAP_CUDA(cudaMemcpyAsync(d_features, e->features(), n * sizeof(float), cudaMemcpyDeviceToDevice,
                              reinterpret_cast<cudaStream_t>(stream)));
  return kOk;
}

int lrpcap_encoder_relevance(lrpcap_encoder_t* enc, const int* h_img_index, const float* d_R_head, int n_words,
                             float* d_R_pix, void* stream) {
  LRPCAP_REQUIRE(enc && enc->impl, kErrInvalidArg, "encoder_relevance: null handle");
  return enc->impl->relevance(h_img_index, d_R_head, n_words, d_R_pix, reinterpret_cast<cudaStream_t>(stream));
}

int lrpcap_encoder_relevance_host(lrpcap_encoder_t* enc, const int* h_img_index, const float* h_R_head, int n_words,
                                  float* h_R_pix, void* stream) {
  LRPCAP_REQUIRE(enc && enc->impl && h_R_head && h_R_pix && n_words > 0, kErrInvalidArg,
                 "encoder_relevance_host: bad argument");
  Encoder* e = enc->impl;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t head = (size_t)n_words * e->feature_hw() * e->feature_hw() * 512;
  const size_t pix = (size_t)n_words * e->image_hw() * e->image_hw() * 3;
  DevBuf dR, dP;
  int st = dR.ensure(head * sizeof(float));
  if (st == kOk) st = dP.ensure(pix * sizeof(float));
  if (st == kOk && cudaMemcpyAsync(dR.p, h_R_head, head * sizeof(float), cudaMemcpyHostToDevice, s) != cudaSuccess) {
    set_last_error("encoder_relevance_host: H2D copy failed");
    st = kErrCuda;
  }
  if (st == kOk) st = e->relevance(h_img_index, dR.as<float>(), n_words, dP.as<float>(), s, h_R_pix);
  if (st == kOk) {
    cudaError_t err = cudaStreamSynchronize(s);
    if (err != cudaSuccess) {
      set_last_error("encoder_relevance_host: %s", cudaGetErrorString(err));
      st = kErrCuda;
    }
  }
  dR.release();
  dP.release();
  return st;
}

int lrpcap_encoder_set_chunk_words(lrpcap_encoder_t* enc, int chunk_words) {
  LRPCAP_REQUIRE(enc && enc->impl && chunk_words > 0, kErrInvalidArg, "encoder_set_chunk_words: bad argument");
  enc->impl->set_chunk_words(chunk_words);
  return kOk;
}

int lrpcap_encoder_set_promote(lrpcap_encoder_t* enc, int every_k_steps) {
  LRPCAP_REQUIRE(enc && enc->impl && every_k_steps >= -1, kErrInvalidArg, "encoder_set_promote: bad argument");
  enc->impl->set_promote(every_k_steps);
  return kOk;
}

long long lrpcap_encoder_launches(lrpcap_encoder_t* enc) { return (enc && enc->impl) ? enc->impl->launches() : 0; }

int lrpcap_encoder_profile(lrpcap_encoder_t* enc, int enable) {
  LRPCAP_REQUIRE(enc && enc->impl, kErrInvalidArg, "encoder_profile: null handle");
  enc->impl->set_profile(enable != 0);
  return kOk;
}

int lrpcap_encoder_profile_read(lrpcap_encoder_t* enc, double* h_out12) {
  LRPCAP_REQUIRE(enc && enc->impl, kErrInvalidArg, "encoder_profile_read: null handle");
  return enc->impl->profile_read(h_out12);
}

int lrpcap_encoder_debug_pool_routes(lrpcap_encoder_t* enc, int layer, unsigned char* h_routes) {
  LRPCAP_REQUIRE(enc && enc->impl, kErrInvalidArg, "encoder_debug_pool_routes: null handle");
  return enc->impl->debug_pool_routes(layer, h_routes);
}

int lrpcap_encoder_debug_multiplier(lrpcap_encoder_t* enc, int layer, int branch, float* h_G) {
  LRPCAP_REQUIRE(enc && enc->impl, kErrInvalidArg, "encoder_debug_multiplier: null handle");
  return enc->impl->debug_multiplier(layer, branch, h_G);
}

int lrpcap_encoder_debug_message_scales(lrpcap_encoder_t* enc, float* h_max, int* h_kt, int cap_words, int* chunk) {
  LRPCAP_REQUIRE(enc && enc->impl, kErrInvalidArg, "encoder_debug_message_scales: null handle");
  return enc->impl->debug_message_scales(h_max, h_kt, cap_words, chunk);
}

int lrpcap_debug_conv_tile(int items, int H, int W, int* tile_w, int* tile_h, int* tile_items) {
  LRPCAP_REQUIRE(tile_w && tile_h && tile_items, kErrInvalidArg, "debug_conv_tile: null argument");
  LRPCAP_REQUIRE(items > 0 && H > 0 && W > 0, kErrShape, "debug_conv_tile: bad shape");
  tc_conv_tile(H, W, items, tile_w, tile_h, tile_items);
  return kOk;
}

int lrpcap_debug_conv(int precision, const float* h_A, int items, int H, int W, int C, const float* h_B, int taps,
                      int Nout, float* h_out) {
  LRPCAP_REQUIRE(h_A && h_B && h_out, kErrInvalidArg, "debug_conv: null argument");
  LRPCAP_REQUIRE(items > 0 && H > 0 && W > 0 && C > 0 && Nout > 0 && (taps == 1 || taps == 9), kErrShape,
                 "debug_conv: bad shape");
  const size_t nA = (size_t)items * H * W * C, nB = (size_t)taps * C * Nout, nO = (size_t)items * H * W * Nout;
  DevBuf dA, dB, dO, sA, sB;
  int st = kOk;
  float wscale = 1.f;
  auto run = [&]() -> int {
    LRPCAP_TRY(dA.ensure(nA * 4));
    LRPCAP_TRY(dB.ensure(nB * 4));
    LRPCAP_TRY(dO.ensure(nO * 4));
    LRPCAP_CUDA(cudaMemcpy(dA.p, h_A, nA * 4, cudaMemcpyHostToDevice));
    LRPCAP_CUDA(cudaMemcpy(dB.p, h_B, nB * 4, cudaMemcpyHostToDevice));
    LRPCAP_CUDA(cudaMemset(dO.p, 0xff, nO * 4));   // NaN-fill: untouched outputs must show up
    EpiParams ep;
    ep.mode = EPI_RAW;
    ep.out_f32 = dO.as<float>();
    if (precision == PREC_BF16X3_TC || precision == 2 || precision == 3 || precision == 4 || precision == 5) {
      // 2: three bf16 planes (the forward pass's arithmetic); 3: two IEEE half planes, promoted (optional forward mode);
      // 4: two-product backward arithmetic (A rounded to ONE fp16 plane x two fp16 weight planes)
      // 5: fp16 + fp8 backward arithmetic (fp16 plane + E4M3 [top bits | residual] plane x fp16 + E4M3 [low | high] weights)
      const int planes = precision == 2 ? 3 : precision == 3 ? kPlanesF16x2 : precision == 4 ? kPlanesH1x2 : precision == 5 ? kPlanesH1F8 : 2;
      const int store_planes = planes == kPlanesH1x2 ? kPlanesF16x2 : planes;
      LRPCAP_TRY(sA.ensure(nA * 2 * 3));
      LRPCAP_TRY(sB.ensure(nB * 2 * 3));
      LRPCAP_TRY(f32_to_split(dA.as<float>(), sA.p, nA, 0, store_planes));
      if (planes == kPlanesH1F8) {   // the byte planes need the weights in the scaled range the encoder uses (Layer::wpow)
        float m = 0.f;
        for (size_t i = 0; i < nB; ++i) m = std::fmax(m, std::fabs(h_B[i]));
        int ex = 0;
        if (m > 0.f && std::isfinite(m)) std::frexp(m, &ex);
        wscale = std::ldexp(1.f, 13 - ex);
      }
      LRPCAP_TRY(prep_weights(dB.as<float>(), sB.p, C, Nout, WF_TC_FWD, WS_ALL, 0, taps, store_planes, wscale));
      TcConvArgs a;
      a.A = sA.p; a.A_elems = nA; a.n_items = items; a.H = H; a.W = W; a.C = C;
      a.B = sB.p; a.B_elems = nB; a.taps = taps; a.Nout = Nout; a.planes = planes; a.epi = ep;
      if (const char* pe = std::getenv("LRPCAP_DEBUG_CONV_PROMOTE")) a.promote_every = std::atoi(pe);   // tests: promoted kernels
      LRPCAP_TRY(tc_conv_launch(a, 0));
    } else {
      SimtConvArgs a;
      a.A = dA.as<float>(); a.n_items = items; a.H = H; a.W = W; a.C = C;
      a.B = dB.as<float>(); a.taps = taps; a.Nout = Nout; a.out_planes = 0; a.epi = ep;
      LRPCAP_TRY(simt_conv_launch(a, 0));
    }
    LRPCAP_CUDA(cudaDeviceSynchronize());
    LRPCAP_CUDA(cudaMemcpy(h_out, dO.p, nO * 4, cudaMemcpyDeviceToHost));
    if (wscale != 1.f)
      for (size_t i = 0; i < nO; ++i) h_out[i] /= wscale;
    return kOk;
  };
  st = run();
  dA.release(); dB.release(); dO.release(); sA.release(); sB.release();
  return st;
}

}  // extern "C"
