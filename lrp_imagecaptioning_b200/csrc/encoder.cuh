// VGG16 (input_1 -> block5_conv3) forward state + batched relevance / gradient backward to pixels.
//
// What the reference does per *word* with one TF session run (analyzer.analyze([img, R]),
// innvestigate/analyzer/base.py:478-520 -> reverse graph built by utils/keras/graph.py:704-942):
// VGG16 forward + per conv layer {forward conv(s), stabilised divide, transposed conv, multiply}.
//
// Here the per-image work is done once (forward(): activations, and per layer one fp32 multiplier tensor
//   G_l = x_l / stab(z_l)   [eps, z rules]     G_l = x_l / safe(z+_l)   [alpha-beta family]     G_l = [z_l > 0]   [gradients]
// with the max-pool arg-max routing folded in as zeros), and the per-word work is a chain of 12 fused
// "transposed conv -> multiply by G" launches over all words at once plus a seed and a 64->3 tail:
//   s_12 = R * M            s_{l-1} = G'_{l-1} * (W_l^T (*) s_l)            R_pix = x_0 * (W_0^T (*) s_0)
// which is algebraically the rule chain of relevance_rule.py (R_in = x * c, s = R_in / stab(z) => s = c * x/stab(z)).
#pragma once
#include <vector>
#include "common.cuh"

namespace lrpcap {

enum Precision : int { PREC_FP32_SIMT = 0, PREC_BF16X3_TC = 1, PREC_F16X2_TC = 2, PREC_TC_AUTO = 3, PREC_H1F8_TC = 4 };

enum RuleKind : int {
  RULE_EPSILON = 0,       // LRPEpsilon            (relevance_analyzer.py:531-552)
  RULE_Z = 1,             // LRPZ
  RULE_ALPHA_BETA = 2,    // LRPAlphaBeta / Alpha1Beta0 / Alpha2Beta1 / ZPlus / SequentialPresetA conv rule
  RULE_ZPLUS_FAST = 3,    // LRPZPlusFast
  RULE_GRADIENT = 4,      // Gradient
  RULE_INPUT_T_GRADIENT = 5,
  RULE_GUIDED_BACKPROP = 6,
};

struct EncoderRule {
  int kind = RULE_EPSILON;
  float epsilon = 1e-7f;
  float alpha = 1.f, beta = 0.f;
  int bias = 1;
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t n);   // grows (never shrinks); contents undefined after growth
  void release();
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

class Encoder {
 public:
  static constexpr int kMaxLayers = 16;
  ~Encoder();
  // arch 0: VGG16 (13 convs, pools after 2/2/3/3), arch 1: VGG19 (16 convs, pools after 2/2/4/4), both up to the last
  // conv of block 5 (models/model.py:419-421: the reference cuts both at their last conv layer; 14 x 14 x 512 head)
  static int create(Encoder** out, const float* const* kernels_hwio, const float* const* biases, int image_hw,
                    int precision, int arch = 0);
  int n_layers() const { return nl_; }
  // Replaces the 13 kernels / biases in place (fine-tuning: the explained model changes every step); keeps every large
  // state / message buffer, drops the prepared weight layouts and the per-image state.
  int set_weights(const float* const* kernels_hwio, const float* const* biases);
  int set_weights_device(const float* const* d_kernels_hwio, const float* const* d_biases);   // the same from device tensors
  // images: device fp32 [n, hw, hw, 3] (already preprocessed). Builds features + per-image rule state.
  int forward(const float* d_images, int n_images, const EncoderRule& rule, cudaStream_t s);
  const float* features() const { return F_.as<float>(); }   // device fp32 [n, hw/16, hw/16, 512]
  int feature_hw() const { return hw_ / 16; }
  int n_images() const { return n_images_; }
  int image_hw() const { return hw_; }
  // h_img_index[w] -> image of word w; d_R_head fp32 [n_words, fh, fh, 512]; d_R_pix fp32 [n_words, hw, hw, 3].
  // h_R_pix (optional, ideally pinned): every chunk of words is copied to the host on a second stream as soon as its
  // last kernel has finished, overlapping the device->host transfer with the next chunk; `s` is made to wait for the
  // last copy, so synchronising `s` covers the host buffer.
  int relevance(const int* h_img_index, const float* d_R_head, int n_words, float* d_R_pix, cudaStream_t s,
                float* h_R_pix = nullptr);
  void set_chunk_words(int n) { chunk_words_ = n > 0 ? n : 1; }
  // k-steps between fp32 promotions of the tensor-core accumulator in the backward GEMMs: 0 = never, n > 0 = every n
  // k-steps on every layer, -1 (default) = once per filter tap on the layers with >= 256 input channels when the rule
  // builds same-sign chains (alpha-beta family, z+), else never
  void set_promote(int every) { bwd_promote_ = every >= -1 ? every : -1; }
  // kernels launched by this object since construction (bench.py's gpu_launches)
  long long launches() const { return launches_; }
  // Kernel timing with CUDA events on the launching stream (bench.py roofline): when enabled every conv launch is
  // bracketed by two events. profile_read() synchronises and returns, per class {tc_bwd, tc_fwd, simt}:
  // out[3*c + 0] = total ms, out[3*c + 1] = algorithmic FLOPs (2*MAC), out[3*c + 2] = launches; then resets.
  void set_profile(bool on) { profile_ = on; }
  int profile_read(double* out9);
  // Test / diagnostics exports of the resident per-image state (synchronous, host outputs):
  //   pool routes: h_out [n_images, H/2, W/2, C] bytes, window position (sy * 2 + sx) the max-pool after conv layer
  //                `layer` routes to (TF MaxPoolGrad's first maximum as this forward pass saw it);
  //   multiplier:  h_out [n_images, H, W, C] fp32, the dense multiplier G_l of conv layer `layer` (< 12) with the pool
  //                routing folded in (zeros away from the arg-max); branch 1 = inhibitor multiplier (beta != 0).
  int debug_pool_routes(int layer, unsigned char* h_out);
  int debug_multiplier(int layer, int branch, float* h_out);
  //   message scales (two-product backward): for the LAST chunk of the last relevance call, h_max [layers + 1][chunk]
  //                = largest |stored fp16 value| of message l per word (row `layers`: the seed's true maximum) and
  //                h_kt [layers][chunk] = log2 of its scale; returns the chunk size through *chunk (0: not two-product)
  int debug_message_scales(float* h_max, int* h_kt, int cap_words, int* chunk);

 private:
  struct ProfRec {
    cudaEvent_t a, b;
    int cls;
    double flops;
  };
  cudaStream_t copy_stream_ = nullptr;
  cudaEvent_t ev_chunk_ = nullptr, ev_copied_ = nullptr;
  bool profile_ = false;
  std::vector<ProfRec> prof_;
  struct Layer {
    int cin, cout, hw;        // hw: spatial size of the conv's input == output
    bool pool_after;
    float* w_hwio = nullptr;  // device fp32 [3,3,cin,cout]
    float* bias = nullptr;    // device fp32 [cout]
    void* prepared[8][3] = {};  // [WeightFormat][WeightSign]
    bool stale[8][3] = {};      // allocation kept, contents older than w_hwio (set_weights_device): re-laid on next use
    bool dual_stale[4] = {};
    int wpow = 0;             // half-plane forward: weights are stored as 2^wpow * w (keeps the low plane out of the subnormals)
    void* dual[4] = {};         // beta != 0: [alpha W+ ; -beta W-] stacked along K, {fp32 SIMT, split-bf16 TC, half-plane TC, fp16 + fp8 TC} backward layouts
  };
  int get_weights(int l, int fmt, int sign, void** out, cudaStream_t s);
  int conv(int l, bool backward, int sign, const void* A, size_t A_elems, int n_items, const struct EpiParams& epi,
           cudaStream_t s, bool dual = false);
  int get_dual_weights(int l, bool tc, void** out, cudaStream_t s);
  bool split() const { return precision_ != PREC_FP32_SIMT; }
  // two-product backward (PREC_F16X2_TC): the relevance message is ONE fp16 plane scaled by a power of two per word and
  // layer, the weights two fp16 planes: 2 MMA products per algorithmic MAC instead of 3, half the message bytes; the
  // message keeps 11 bits instead of 16 (DESIGN.md section 5 for what that costs per rule)
  // PREC_TC_AUTO. The plain two-product mode keeps an 11-bit message: per layer 2e-4 of the GROSS relevance flowing
  // through. The tolerance is 1e-3 of the NET map, so the mode is only picked where little cancels: the purely
  // positive-flow rules (alpha1 beta0 / z+, the reference explainers' default PresetA) on maps of at least 128 x 128 (a
  // head of >= 64 cells: on smaller images every head cell covers the whole image, positive and negative head relevance
  // land on the same pixels and the error relative to what is left grows by sum|R| / |sum R| -- measured 4.8e-3 at
  // 32 x 32, profiles/r02_diag_small_images.jsonl). Everything else (epsilon, z, gradients, alpha-beta with beta > 0,
  // small images) runs the fp16 + fp8 mode below, which carries ~15 bits.
  bool two_product() const {
    if (precision_ == PREC_F16X2_TC) return true;
    if (precision_ != PREC_TC_AUTO || hw_ < 128) return false;
    return rule_.kind == RULE_ZPLUS_FAST || (rule_.kind == RULE_ALPHA_BETA && rule_.beta == 0.f);
  }
  // fp16 + fp8 backward (epilogue.cuh: StoreH1F8): the scaled fp16 message plus an E4M3 plane of [its top bits | its
  // rounding residual] against the weights' fp16 high plane plus an E4M3 plane of [their low part | their high part]: one
  // kind::f16 and one double-rate kind::f8f6f4 product = two product-equivalents with ~15 bits (error per layer 1e-5, the
  // three-product mode's 4e-6). LRPCAP_H1F8=0 sends PREC_TC_AUTO back to three bf16 products.
  bool fp8_mode() const {
    return precision_ == PREC_H1F8_TC || (precision_ == PREC_TC_AUTO && !two_product() && !getenv_off("LRPCAP_H1F8"));
  }
  static bool getenv_off(const char* name);
  bool scaled_messages() const { return two_product() || fp8_mode(); }   // fp16 message planes with per-word scales
  // storage planes of forward activations in tensor-core mode (fp32 otherwise). The per-image forward decides ReLU signs
  // and pool arg-max and forms x/stab(z); 16-bit operands there cost 1e-2-level map errors, so its
  // operands carry >= 22 bits: two IEEE half planes (default, 3 MMA products; falls back for good when an activation
  // leaves the half range) or three bf16 planes (6 products, LRPCAP_FWD_PLANES=3).
  int fwd_planes() const { return split() ? fwd_planes_ : 0; }
  size_t layer_out_elems(int l) const { return (size_t)L_[l].hw * L_[l].hw * L_[l].cout; }

  int hw_ = 0, precision_ = 0, n_images_ = 0, chunk_words_ = 256, bwd_promote_ = -1, fwd_promote_ = 1, fwd_planes_ = kPlanesF16x2;
  int* d_overflow_ = nullptr;   // half-plane forward: set by the epilogue when an activation leaves the half range
  void set_wpow(int l, const float* h_w);
  long long launches_ = 0;
  EncoderRule rule_;
  int nl_ = 13;
  Layer L_[kMaxLayers];
  std::vector<float> w0_host_;       // first-layer kernel (host copy) for the signed-input alpha-beta path
  float* w0_pm_ = nullptr;           // device fp32 [9][6][64]: [W+ ; W-] stacked for the [x+, x-] input
  float* w0_mp_ = nullptr;           // [W- ; W+] (inhibitor branch, beta != 0)
  float *w0_last_a_ = nullptr, *w0_last_b_ = nullptr;   // dual last-layer weights [9][128][3] for x >= 0 / x < 0
  float dual_alpha_ = 0.f, dual_beta_ = 0.f;            // (alpha, beta) the cached dual weights were built for
  DevBuf X0_, F_, Mseed_, G_[kMaxLayers - 1];
  DevBuf Mseed2_, G2_[kMaxLayers - 1];   // inhibitor-branch multipliers (beta != 0)
  // layers followed by a max-pool: compact multipliers (value at the window's arg-max + 2-bit position per channel)
  // read by the up-sampling epilogue; G_[l] / G2_[l] of those layers are scratch between the conv and pool_mask
  DevBuf Gc_[kMaxLayers - 1], Gc2_[kMaxLayers - 1], Gi_[kMaxLayers - 1];
  DevBuf act_[3], posneg_, msg_[2], idx_;
  int msg_target_exp_ = 4;   // two-product backward: predicted maximum of a stored message plane is 2^msg_target_exp_
  int scale_cw_ = 0, scale_m_ = 0;   // chunk stride / words of the last chunk behind scale_
  DevBuf scale_;   // two-product backward: per chunk [layers + 1][chunk] max bits + [layers][chunk] scale exponents (epilogue.cuh)
};

}  // namespace lrpcap
