/* lrpcap -- C ABI of the B200-native LRP engine for attention-LSTM image captioners.
 *
 * The reference (SunJiamei/LRP-ImageCaptioning) has no FFI: its boundary is the Python class surface of
 * models/explainers.py and the iNNvestigate analyzers.  This header is the boundary a binding for that
 * surface attaches to (ctypes stubs in INTEGRATION.md; the in-tree host layer is
 * lrp_imagecaptioning_b200/{_lib,explainers,analyzers}.py).  Conventions:
 *   - every function returns 0 on success or a negative LRPCAP_ERR_* code; the message is in
 *     lrpcap_last_error() (thread-local).  No exceptions cross this boundary.
 *   - "d_" pointers are device memory on the current CUDA device, "h_" pointers are host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous on that
 *     stream unless they return host data.
 *   - tensors are NHWC fp32, channel order as the model sees it (BGR, caffe-mode preprocessing done by the caller).
 *   - one handle per GPU/stream; distinct handles may be used from distinct threads.
 */
#ifndef LRPCAP_H_
#define LRPCAP_H_

#ifdef __cplusplus
extern "C" {
#endif

#define LRPCAP_OK 0
#define LRPCAP_ERR_INVALID_ARG (-1)
#define LRPCAP_ERR_SHAPE (-2)
#define LRPCAP_ERR_CUDA (-3)
#define LRPCAP_ERR_UNSUPPORTED (-4)
#define LRPCAP_ERR_STATE (-5)

/* arithmetic of the dense contractions */
#define LRPCAP_PREC_FP32_SIMT 0 /* fp32 FMA on CUDA cores (exact-fp32 validation mode) */
#define LRPCAP_PREC_BF16X3_TC 1 /* tcgen05 tensor cores, split 16-bit operands (3 products), fp32 accumulate: two bf16 planes
                                 * for the per-word relevance messages, two IEEE half planes (22 bits; automatic fall-back
                                 * to three bf16 planes outside the half range) for the per-image forward */

#define LRPCAP_PREC_F16X2_TC 2  /* as BF16X3_TC, but the per-word relevance messages are ONE fp16 plane scaled by a power of
                                 * two per word and layer, against two fp16 weight planes: 2 products per MAC instead of 3
                                 * and half the message bytes; 11-bit messages (alpha-beta family: 2e-5 maps) */

#define LRPCAP_PREC_TC_AUTO 3   /* tensor cores, arithmetic per rule: the two-product fp16 mode for the positive-flow rules
                                 * (alpha1 beta0 / z+, images of at least 128 x 128), the fp16 + fp8 mode for everything
                                 * else (epsilon, z, gradient family, alpha-beta with beta > 0, small images) */
#define LRPCAP_PREC_H1F8_TC 4   /* messages as a scaled fp16 plane + an E4M3 plane of [top bits | rounding residual], weights as
                                 * an fp16 high plane + an E4M3 plane of [low part | high part]: one kind::f16 and one
                                 * double-rate kind::f8f6f4 product per MAC (two product-equivalents, ~15 bits) */

/* encoder rules; replaces the iNNvestigate analyzer classes constructed in
 * models/explainers.py:32,671,883,928 (LRPSequentialPresetA, Gradient, InputTimesGradient, GuidedBackprop) and
 * innvestigate/analyzer/relevance_based/relevance_analyzer.py:531-721 */
#define LRPCAP_RULE_EPSILON 0          /* LRPEpsilon(epsilon, bias) */
#define LRPCAP_RULE_Z 1                /* LRPZ(bias) */
#define LRPCAP_RULE_ALPHA_BETA 2       /* LRPAlphaBeta(alpha, beta, bias); Alpha1Beta0, ZPlus, SequentialPresetA */
#define LRPCAP_RULE_ZPLUS_FAST 3       /* LRPZPlusFast */
#define LRPCAP_RULE_GRADIENT 4         /* Gradient */
#define LRPCAP_RULE_INPUT_T_GRADIENT 5 /* InputTimesGradient */
#define LRPCAP_RULE_GUIDED_BACKPROP 6  /* GuidedBackprop */

#define LRPCAP_DECODER_ADAPTIVE 0 /* ExplainImgCaptioningAdaptiveAttention*, explainers.py:260-949 */
#define LRPCAP_DECODER_GRIDTD 1   /* ExplainImgCaptioningGridTD*,            explainers.py:995-1653 */

typedef struct lrpcap_encoder lrpcap_encoder_t;
typedef struct lrpcap_decoder lrpcap_decoder_t;

const char* lrpcap_last_error(void);
int lrpcap_version(void);

/* ----------------------------------------------------------------------------------------------- encoder
 * VGG16 input_1 -> block5_conv3 (the `_image_model` of explainers.py:29-30). */

/* h_kernels_hwio[13], h_biases[13]: host fp32, Keras layouts (3,3,Cin,Cout) / (Cout,), layer order
 * block1_conv1 .. block5_conv3.  image_hw: 224 (any multiple of 16 is accepted). */
int lrpcap_encoder_create(lrpcap_encoder_t** out, const float* const* h_kernels_hwio, const float* const* h_biases,
                          int image_hw, int precision);
/* Other encoders the reference offers (models/model.py:419-429, `img_encoder`): arch LRPCAP_ENCODER_VGG16 (13 convs) or
 * LRPCAP_ENCODER_VGG19 (16 convs: block1_conv1 .. block5_conv4), both cut at their last conv layer (14 x 14 x 512 head). */
#define LRPCAP_ENCODER_VGG16 0
#define LRPCAP_ENCODER_VGG19 1
int lrpcap_encoder_create_arch(lrpcap_encoder_t** out, int arch, const float* const* h_kernels_hwio,
                               const float* const* h_biases, int image_hw, int precision);
/* Replaces the weights of an existing handle in place (same layouts as lrpcap_encoder_create) and invalidates the
 * per-image state: what `LRPInferenceLayer*` needs between fine-tuning steps (train.py:569-577), where the reference
 * explains the model that is being trained; the large state and message buffers are kept. Synchronises the device. */
int lrpcap_encoder_set_weights(lrpcap_encoder_t* enc, const float* const* h_kernels_hwio, const float* const* h_biases);
/* The same from DEVICE tensors (arrays of device pointers, same layouts): the fine-tuning step keeps the trained
 * parameters on the device, so the explained model follows them without a host round trip. */
int lrpcap_encoder_set_weights_device(lrpcap_encoder_t* enc, const float* const* d_kernels_hwio, const float* const* d_biases);
int lrpcap_encoder_destroy(lrpcap_encoder_t* enc);

/* Replaces `_image_model.predict(img)` (explainers.py:375, 1097) and the forward half of
 * `analyzer.analyze([X, R])` (innvestigate/analyzer/base.py:478-520): runs the conv stack once per image and keeps,
 * per image and layer, the multiplier tensor the chosen rule needs in the backward pass.
 * d_images: [n_images, hw, hw, 3]. */
int lrpcap_encoder_forward(lrpcap_encoder_t* enc, const float* d_images, int n_images, int rule, float epsilon,
                           float alpha, float beta, int bias, void* stream);
/* d_features: [n_images, hw/16, hw/16, 512] (post-ReLU block5_conv3). */
int lrpcap_encoder_features(lrpcap_encoder_t* enc, float* d_features, void* stream);

/* Replaces `_explain_CNN(X, relevance_value)` = `analyzer.analyze([X, R])` (explainers.py:179-181) for a whole batch
 * of words: word w belongs to image h_img_index[w]; d_R_head[w] is the tensor that seeds the backward pass at
 * block5_conv3's output ('replace' mode, base.py:366-410, graph.py:898-900); d_R_pix[w] is the result at input_1.
 * d_R_head: [n_words, hw/16, hw/16, 512]; d_R_pix: [n_words, hw, hw, 3]. */
int lrpcap_encoder_relevance(lrpcap_encoder_t* enc, const int* h_img_index, const float* d_R_head, int n_words,
                             float* d_R_pix, void* stream);
/* Host-buffer variant (H2D of R_head and D2H of R_pix inside the call, synchronous). */
int lrpcap_encoder_relevance_host(lrpcap_encoder_t* enc, const int* h_img_index, const float* h_R_head, int n_words,
                                  float* h_R_pix, void* stream);
int lrpcap_encoder_set_chunk_words(lrpcap_encoder_t* enc, int chunk_words);
/* Tensor-core mode: the transposed-conv GEMMs hand their TMEM accumulator to fp32 registers every `every_k_steps`
 * k-steps (one k-step = 64 channels of one tap). tcgen05 accumulation rounds toward zero: a chain of n accumulates
 * shrinks every output by ~1.5e-8 n, i.e. -1.6e-5 per 512-channel layer and -1e-4 over the 12 layers -- a uniform scale
 * error that alone breaks the 1e-4 conservation tolerance for the same-sign alpha-beta chains (profiles/r02_diag_parity.jsonl).
 * -1 (default): once per filter tap on the layers with >= 256 input channels for the alpha-beta family and z+ (sum error
 * 1.5e-5, +5 % time), never for the mixed-sign rules (their sums are unaffected: 6e-7); 0: never; n > 0: every n k-steps
 * on every layer. */
int lrpcap_encoder_set_promote(lrpcap_encoder_t* enc, int every_k_steps);
long long lrpcap_encoder_launches(lrpcap_encoder_t* enc);
/* Kernel timing (CUDA events on the launching stream around every convolution launch) for roofline reporting.
 * h_out12: per class {0: tcgen05 transposed conv (relevance), 1: tcgen05 forward conv, 2: fp32 SIMT conv,
 * 3: last transposed conv (64 -> 3 channels) + input re-weighting} [total ms, algorithmic FLOPs, launches];
 * reading synchronises and resets the counters. */
int lrpcap_encoder_profile(lrpcap_encoder_t* enc, int enable);
int lrpcap_encoder_profile_read(lrpcap_encoder_t* enc, double* h_out12);

/* ----------------------------------------------------------------------------------------------- decoder */

typedef struct lrpcap_decoder_weights {
  int kind;          /* LRPCAP_DECODER_* */
  int V, H, E, D;    /* vocabulary, hidden, embedding, CNN feature depth */
  /* shared heads (explainers.py:264-269,278 / 1000-1006,1016); all host fp32, Keras layouts (in, out) */
  const float* image_features_w; /* (D, H)  */
  const float* image_features_b; /* (H)     */
  const float* global_w;         /* (D, E)  */
  const float* global_b;         /* (E)     */
  const float* embedding;        /* (V, E)  */
  const float* output_w;         /* (H, V)  */
  const float* output_b;         /* (V)     */
  /* adaptive attention (explainers.py:270-277): LSTM (2E,4H),(H,4H),(4H); Wv,Wg,Wh,Ws (H,H); Wx (2E,H); V (H,1) */
  const float *lstm_wi, *lstm_wh, *lstm_b, *Wv, *Wg, *Wx, *Wh, *Ws, *Vatt;
  /* grid-TD (explainers.py:1007-1019): language LSTM (2H,4H),(H,4H),(4H); top-down LSTM (H+2E,4H),(H,4H),(4H);
   * W_va,W_ha,W_h,W_s (H,H); W_x (H+2E,H); W_a (H,1) */
  const float *lang_wi, *lang_wh, *lang_b, *td_wi, *td_wh, *td_b, *W_va, *W_ha, *W_a, *W_x, *W_h, *W_s;
} lrpcap_decoder_weights_t;

/* keras_logits != 0: grid-TD logits = (h2 + c_hat) W_o + b as in the Keras model (models/model.py:816);
 * 0 (default): h2 W_o + b as in the reference explainer (explainers.py:1154). */
int lrpcap_decoder_create(lrpcap_decoder_t** out, const lrpcap_decoder_weights_t* w, int sos_token, int keras_logits);
/* Replaces the weights of an existing handle from DEVICE fp32 tensors (the struct's pointers are device memory, same
 * layouts and dimensions as at creation): every derived layout is recomputed in place on the device, the forward state is
 * dropped.  What the LRP-inference fine-tuning step needs between optimizer steps (train.py:569-577). */
int lrpcap_decoder_set_weights_device(lrpcap_decoder_t* dec, const lrpcap_decoder_weights_t* d_w);
int lrpcap_decoder_destroy(lrpcap_decoder_t* dec);

/* Replaces `_forward_beam_search(X, caption)` (explainers.py:370-436, 690-778, 1092-1178, 1344-1450) for a batch:
 * teacher-forced decoder forward over T steps storing every intermediate the relevance pass reads.
 * d_features: [n_images, L, D]; h_captions: [n_images, T] tokenizer ids (model index = id - 1).
 * greedy == 1: h_captions is an OUTPUT -- token t is the arg-max of step t's logits and is fed back as the next input.
 * greedy == 2 ("predict"): h_captions holds the teacher tokens on input (fed as at greedy == 0) and the arg-max of every
 * step's logits on output -- `argmax(model.predict(...))` of the fine-tuning loop (train.py:570, models/model.py:1661);
 * the handle then holds no explainable state (run a greedy == 0 forward on the predicted caption next).
 * eos_token >= 1: that tokenizer id is excluded from the arg-max (every caption then has exactly T words); eos_token <= 0:
 * plain arg-max. Captions are never truncated at EOS: every image has T positions, and a caller that wants the reference's
 * "stop at EOS" restricts the word list it passes to lrpcap_decoder_relevance itself. */
int lrpcap_decoder_forward(lrpcap_decoder_t* dec, const float* d_features, int n_images, int L, int* h_captions, int T,
                           int greedy, int eos_token, void* stream);

/* Replaces `_explain_lstm_single_word_sequence(t)` (explainers.py:537-666, 1180-1321) for a batch of words:
 * word w = (image h_word_img[w], 1-based position h_word_t[w]).
 * d_R_head: [n_words, L, D] fp32 relevance of the CNN grid features.
 * h_r_words: optional [n_words, T] (adaptive: normalised, entry j = reference r_words[j] for j < t-1, rest 0;
 * grid-TD: raw, entries j < t).  h_attention: optional [n_words, L] = attention[t]. */
int lrpcap_decoder_relevance(lrpcap_decoder_t* dec, const int* h_word_img, const int* h_word_t, int n_words,
                             float* d_R_head, double* h_r_words, float* h_attention, void* stream);
/* Replaces `_lstm_decoder_backward(t)` (explainers.py:780-832, 1452-1532): manual BPTT with frozen attention. */
int lrpcap_decoder_backward(lrpcap_decoder_t* dec, const int* h_word_img, const int* h_word_t, int n_words,
                            float* d_R_head, double* h_r_words, void* stream);
/* h_logit: [n_images, T] logit of the caption token at each step (fp64). */
int lrpcap_decoder_caption_logits(lrpcap_decoder_t* dec, double* h_logit);
/* The `.attention` / `.beta` attributes the reference explainers expose after `_forward_beam_search`
 * (explainers.py:429-431): h_alpha [n_images, T+1, L] and h_beta [n_images, T+1], row 0 all zeros. Either may be NULL. */
int lrpcap_decoder_attention(lrpcap_decoder_t* dec, float* h_alpha, float* h_beta);
/* Full logits [n_images, V] (fp64) of the LAST step of the most recent decoder_forward: the quantity beam search ranks
 * (`keras_model.predict_on_batch` + `preds[:, -1]`, explainers.py:73-76).  Synchronous. */
int lrpcap_decoder_last_logits(lrpcap_decoder_t* dec, double* h_logits, void* stream);
long long lrpcap_decoder_launches(lrpcap_decoder_t* dec);

/* ----------------------------------------------------------------------------------------------- whole path
 * One call per batch, host buffers in and out (the end-to-end entry the benchmark times):
 *   images -> encoder forward -> decoder forward (teacher-forced or greedy) -> decoder relevance for every
 *   (image, t = 1..T) word -> encoder relevance -> pixel maps.
 * h_images [n, hw, hw, 3]; h_captions [n, T] (in, or out when greedy); h_R_pix [n*T, hw, hw, 3], word (i, t) at
 * row i*T + (t-1).  method: 0 = LRP (decoder relevance), 1 = gradient family (decoder backward). */
int lrpcap_explain_batch_host(lrpcap_encoder_t* enc, lrpcap_decoder_t* dec, const float* h_images, int n_images,
                              int* h_captions, int T, int greedy, int eos_token, int method, int rule, float epsilon,
                              float alpha, float beta, int bias, float* h_R_pix, void* stream);

/* ----------------------------------------------------------------------------------------------- Grad-CAM
 * Replaces `grad_cam(img_feature, grads)` (explainers.py:939-949, 1643-1653) for a batch of words:
 *   weights = mean_xy grads; cam = sum_k weights_k F_k; pyramid_expand(cam, upscale, sigma); relu; / (max|cam| + 1e-6).
 * d_features [n_images, fh*fh, D]; d_grads [n_words, fh*fh, D] (the decoder gradient); d_cam [n_words, fh*upscale, fh*upscale].
 * The reference uses upscale = 16, sigma = 20.  Synchronous. */
int lrpcap_gradcam(const float* d_features, const int* h_img_index, const float* d_grads, int n_words, int fh, int D,
                   int upscale, float sigma, float* d_cam, void* stream);
/* Guided-Grad-CAM (explainers.py:930-937): d_maps[w, y, x, c] *= d_cam[w, y, x]; d_maps [n_words, hw, hw, 3]. */
int lrpcap_scale_maps(float* d_maps, const float* d_cam, int n_words, int hw, void* stream);

/* ----------------------------------------------------------------------------------------------- LRP-inference
 * Replaces the per-word score of `LRPInferenceLayer{Adaptive,gridTD}.call` (models/model.py:1673-1688, 2045-2060):
 * hp = mean_c(map); hp /= max|hp|; score = mean(hp) (mode 0, 'mean') or mean(relu(hp)) (mode 1, 'pos_mean').
 * d_maps [n_words, hw, hw, 3]; h_scores [n_words].  Synchronous. */
int lrpcap_lrp_inference_scores(const float* d_maps, int n_words, int hw, int mode, float* h_scores, void* stream);

/* ----------------------------------------------------------------------------------------------- evaluation reductions
 * Heat maps the reference derives from a pixel relevance map before it evaluates or plots it.
 *   mode 0: hp = mean_c(postprocess(R))                    exaimin_word.py:96-103, :127-128
 *   mode 1: hp = mean_c(relu(-postprocess(R)))             evaluate_bbox.py:79-83 (negative scores, the committed variant)
 *   mode 2: hp = mean_c(relu(postprocess(R)))              evaluate_bbox.py:80 commented out
 * then `project`: hp / max|hp| (all zeros if the maximum is 0), and, when shift_negative != 0 and some value is
 * negative, (hp + 1) / 2 (evaluate_bbox.py:60-69).  window > 1 first pools window x window tiles of the channel value,
 * pool_type 0 = max, 1 = average (exaimin_word.py:64-77, :143-148; the reference uses 16 on 224 x 224 maps).
 * d_maps [n_words, hw, hw, 3]; d_out [n_words, hw/window, hw/window]; h_means (optional) [n_words] = mean of each
 * projected map (exaimin_word.py:446-447).  Synchronous when h_means is given. */
int lrpcap_heatmaps(const float* d_maps, int n_words, int hw, int mode, int shift_negative, int window, int pool_type,
                    float* d_out, float* h_means, void* stream);
/* Bounding-box "correctness" `_calculate_overlaped_pixels` (evaluate_bbox.py:191-208) for many boxes and thresholds:
 * values <= threshold are dropped; ratio = mass inside the box / total mass, 0 when the total is 0, capped at 1.
 * d_heatmaps [n_maps, hw, hw]; h_boxes [n_boxes, 5] = (map index, x0, y0, x1, y1), rows y0..y1-1, columns x0..x1-1;
 * h_ratio [n_boxes, n_thresholds] (n_thresholds <= 16).  Synchronous. */
int lrpcap_bbox_correctness(const float* d_heatmaps, int n_maps, int hw, const int* h_boxes, int n_boxes,
                            const float* h_thresholds, int n_thresholds, float* h_ratio, void* stream);

/* ----------------------------------------------------------------------------------------------- debug / tests
 * Discrete decisions of the resident forward state, for parity tests: every rule is discontinuous at max-pool arg-max
 * ties (TF MaxPoolGrad routes to the first maximum, innvestigate relevance_analyzer.py:459-480) and the gradient
 * family / Z rule also at ReLU kinks, so a test pins the oracle to the decisions this forward pass took.
 * lrpcap_encoder_debug_pool_routes: conv layer `layer` (0-based; 1, 3, 6, 9 are followed by a pool), h_routes
 * [n_images, H/2, W/2, C] bytes = window position (sy * 2 + sx) each pooled element routes to.
 * lrpcap_encoder_debug_multiplier: h_G [n_images, H, W, C] fp32 = dense per-image multiplier of conv layer `layer` < 12
 * (x/stab(z), x/safe(z+), [z > 0] ... by rule; zero away from the pool arg-max); branch 1 = inhibitor (beta != 0).
 * Both synchronise the device. */
int lrpcap_encoder_debug_pool_routes(lrpcap_encoder_t* enc, int layer, unsigned char* h_routes);
int lrpcap_encoder_debug_multiplier(lrpcap_encoder_t* enc, int layer, int branch, float* h_G);
/* Two-product backward: the power-of-two scale bookkeeping of the LAST chunk of the last lrpcap_encoder_relevance call.
 * h_max [layers + 1][*chunk]: largest |stored fp16 value| of the message entering conv layer l's transposed conv, per
 * word (row `layers`: the seed's true maximum); h_kt [layers][*chunk]: log2 of that message's scale. cap_words: room (in
 * words) of both arrays; *chunk = words of that chunk, 0 when the handle does not run the two-product path. */
int lrpcap_encoder_debug_message_scales(lrpcap_encoder_t* enc, float* h_max, int* h_kt, int cap_words, int* chunk);
/* Host-only: the accumulator tile the generic tcgen05 kernel uses for a [items, H, W] map: tile_w x tile_h pixels of
 * tile_items consecutive items (<= 128 rows). No device work; unit tests of the tile chooser. */
int lrpcap_debug_conv_tile(int items, int H, int W, int* tile_w, int* tile_h, int* tile_items);
/* Single convolution through one implementation, raw accumulator out (unit tests of the GEMM kernels).
 * precision: LRPCAP_PREC_FP32_SIMT; LRPCAP_PREC_BF16X3_TC (two bf16 planes: the backward arithmetic); 2 = three bf16
 * planes, promoted; 3 = two IEEE half planes, promoted (the forward arithmetic); 4 = the two-product backward arithmetic
 * (A rounded to one fp16 plane x two fp16 weight planes); 5 = the fp16 + fp8 backward arithmetic. The environment
 * variable LRPCAP_DEBUG_CONV_PROMOTE=<k-steps> selects the promoted kernels (tests).
 * h_A [items, H, W, C]; h_B [taps][C][Nout] (HWIO for taps = 9); h_out [items, H, W, Nout]. */
int lrpcap_debug_conv(int precision, const float* h_A, int items, int H, int W, int C, const float* h_B, int taps,
                      int Nout, float* h_out);

#ifdef __cplusplus
}
#endif
#endif /* LRPCAP_H_ */
