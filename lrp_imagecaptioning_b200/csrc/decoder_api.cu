// extern "C" decoder entry points and the whole-path host-buffer call (include/lrpcap.h).
#include "../../include/lrpcap.h"
#include "handles.cuh"
#include <vector>

using namespace lrpcap;


extern "C" {

int lrpcap_decoder_create(lrpcap_decoder_t** out, const lrpcap_decoder_weights_t* w, int sos_token, int keras_logits) {
  LRPCAP_REQUIRE(out != nullptr, kErrInvalidArg, "decoder_create: null out");
  Decoder* d = nullptr;
  LRPCAP_TRY(Decoder::create(&d, w, sos_token, keras_logits));
  *out = new lrpcap_decoder{d};
  return kOk;
}

int lrpcap_decoder_set_weights_device(lrpcap_decoder_t* dec, const lrpcap_decoder_weights_t* d_w) {
  LRPCAP_REQUIRE(dec && dec->impl, kErrInvalidArg, "decoder_set_weights_device: null handle");
  return dec->impl->set_weights_device(d_w);
}

int lrpcap_decoder_destroy(lrpcap_decoder_t* dec) {
  if (!dec) return kOk;
  delete dec->impl;
  delete dec;
  return kOk;
}

int lrpcap_decoder_forward(lrpcap_decoder_t* dec, const float* d_features, int n_images, int L, int* h_captions, int T,
                           int greedy, int eos_token, void* stream) {
  LRPCAP_REQUIRE(dec && dec->impl, kErrInvalidArg, "decoder_forward: null handle");
  return dec->impl->forward(d_features, n_images, L, h_captions, T, greedy, eos_token,
                            reinterpret_cast<cudaStream_t>(stream));
}

int lrpcap_decoder_relevance(lrpcap_decoder_t* dec, const int* h_word_img, const int* h_word_t, int n_words,
                             float* d_R_head, double* h_r_words, float* h_attention, void* stream) {
  LRPCAP_REQUIRE(dec && dec->impl, kErrInvalidArg, "decoder_relevance: null handle");
  return dec->impl->relevance(h_word_img, h_word_t, n_words, d_R_head, h_r_words, h_attention,
                              reinterpret_cast<cudaStream_t>(stream));
}

int lrpcap_decoder_backward(lrpcap_decoder_t* dec, const int* h_word_img, const int* h_word_t, int n_words,
                            float* d_R_head, double* h_r_words, void* stream) {
  LRPCAP_REQUIRE(dec && dec->impl, kErrInvalidArg, "decoder_backward: null handle");
  return dec->impl->backward(h_word_img, h_word_t, n_words, d_R_head, h_r_words, reinterpret_cast<cudaStream_t>(stream));
}

int lrpcap_decoder_caption_logits(lrpcap_decoder_t* dec, double* h_logit) {
  LRPCAP_REQUIRE(dec && dec->impl, kErrInvalidArg, "decoder_caption_logits: null handle");
  return dec->impl->caption_logits(h_logit);
}

int lrpcap_decoder_attention(lrpcap_decoder_t* dec, float* h_alpha, float* h_beta) {
  LRPCAP_REQUIRE(dec && dec->impl, kErrInvalidArg, "decoder_attention: null handle");
  return dec->impl->attention(h_alpha, h_beta);
}

int lrpcap_decoder_last_logits(lrpcap_decoder_t* dec, double* h_logits, void* stream) {
  LRPCAP_REQUIRE(dec && dec->impl, kErrInvalidArg, "decoder_last_logits: null handle");
  return dec->impl->last_logits(h_logits, reinterpret_cast<cudaStream_t>(stream));
}

long long lrpcap_decoder_launches(lrpcap_decoder_t* dec) { return (dec && dec->impl) ? dec->impl->launches() : 0; }

int lrpcap_explain_batch_host(lrpcap_encoder_t* enc, lrpcap_decoder_t* dec, const float* h_images, int n_images,
                              int* h_captions, int T, int greedy, int eos_token, int method, int rule, float epsilon,
                              float alpha, float beta, int bias, float* h_R_pix, void* stream) {
  LRPCAP_REQUIRE(enc && enc->impl && dec && dec->impl, kErrInvalidArg, "explain_batch_host: null handle");
  LRPCAP_REQUIRE(h_images && h_captions && h_R_pix && n_images > 0 && T > 0, kErrInvalidArg,
                 "explain_batch_host: bad argument");
  Encoder* E = enc->impl;
  Decoder* D = dec->impl;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int hw = E->image_hw(), fh = E->feature_hw(), L = fh * fh;
  const size_t img_elems = (size_t)n_images * hw * hw * 3;
  const int W = n_images * T;
  DevBuf &d_img = enc->stage_img, &d_head = enc->stage_head, &d_pix = enc->stage_pix;   // reused across calls
  LRPCAP_TRY(d_img.ensure(img_elems * sizeof(float)));
  LRPCAP_TRY(d_head.ensure((size_t)W * L * 512 * sizeof(float)));
  LRPCAP_TRY(d_pix.ensure((size_t)W * hw * hw * 3 * sizeof(float)));
  LRPCAP_CUDA(cudaMemcpyAsync(d_img.p, h_images, img_elems * sizeof(float), cudaMemcpyHostToDevice, s));
  EncoderRule r;
  r.kind = rule; r.epsilon = epsilon; r.alpha = alpha; r.beta = beta; r.bias = bias;
  LRPCAP_TRY(E->forward(d_img.as<float>(), n_images, r, s));
  LRPCAP_TRY(D->forward(E->features(), n_images, L, h_captions, T, greedy, eos_token, s));
  std::vector<int> wimg(W), wt(W);
  for (int i = 0; i < n_images; ++i)
    for (int t = 1; t <= T; ++t) {
      wimg[(size_t)i * T + t - 1] = i;
      wt[(size_t)i * T + t - 1] = t;
    }
  if (method == 0) LRPCAP_TRY(D->relevance(wimg.data(), wt.data(), W, d_head.as<float>(), nullptr, nullptr, s));
  else LRPCAP_TRY(D->backward(wimg.data(), wt.data(), W, d_head.as<float>(), nullptr, s));
  LRPCAP_TRY(E->relevance(wimg.data(), d_head.as<float>(), W, d_pix.as<float>(), s, h_R_pix));   // chunk-wise D2H overlap
  LRPCAP_CUDA(cudaStreamSynchronize(s));
  return kOk;
}

}  // extern "C"
