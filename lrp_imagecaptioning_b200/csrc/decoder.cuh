// Attention-LSTM decoder: batched teacher-forced forward that stores every intermediate, and batched
// word-level relevance (epsilon-LRP) / frozen-attention gradient back to the CNN grid features.
//
// Reference: /root/reference/models/explainers.py
//   adaptive  forward :370-436   LRP :537-666    gradient :690-832
//   grid-TD   forward :1092-1178 LRP :1180-1321  gradient :1344-1532
//   helper    _propagate_relevance_linear_lrp :156-165, _get_sign_stabilizer :141-144 (eps = 1e-7)
//
// The reference runs one word at a time with 6+3t+3L (adaptive) / 4+t(8+L)+2L (grid-TD) helper calls per
// word, each materialising a dense attribution matrix (identity matrices for element-wise steps).  Here all
// (image, word) pairs advance together: the time loop i = T-1..0 runs once for the whole batch (words sorted
// by position so the active ones are a prefix), every element-wise rule is one fused kernel over [words, H],
// and every dense step is one GEMM whose M dimension is the number of active words (or words x L).
// Arithmetic is fp64 (the reference computes this stage in NumPy float64 almost everywhere),
// with the reference's float32 stores (r_V, r_img_feature_input) reproduced.
#pragma once
#include <vector>
#include "common.cuh"
#include "encoder.cuh"   // DevBuf

struct lrpcap_decoder_weights;

namespace lrpcap {

class Decoder {
 public:
  ~Decoder();
  static int create(Decoder** out, const lrpcap_decoder_weights* w, int sos_token, int keras_logits);
  // the same tensors as DEVICE fp32 pointers: re-derives every layout in place (fine-tuning), drops the forward state
  int set_weights_device(const lrpcap_decoder_weights* d_w);
  int forward(const float* d_features, int n_images, int L, int* h_captions, int T, int greedy, int eos_token,
              cudaStream_t s);
  int relevance(const int* h_word_img, const int* h_word_t, int n_words, float* d_R_head, double* h_r_words,
                float* h_attention, cudaStream_t s);
  int backward(const int* h_word_img, const int* h_word_t, int n_words, float* d_R_head, double* h_r_words,
               cudaStream_t s);
  int caption_logits(double* h_logit);
  int attention(float* h_alpha, float* h_beta);
  int last_logits(double* h_logits, cudaStream_t s);   // [N, V] logits of the last forward step
  long long launches() const { return launches_; }
  int n_images() const { return N_; }
  int T() const { return T_; }
  int L() const { return L_; }
  int D() const { return D_; }

 private:
  struct WField {
    const float* lrpcap_decoder_weights::*p;
    size_t n;
  };
  std::vector<WField> fields() const;
  int derive(const lrpcap_decoder_weights* d_w);                             // all derived layouts from device fp32 tensors
  int slot(size_t bytes, void** out);                                        // allocation that set_weights_device() re-uses
  int upload(const float* d, size_t n, double** out);                        // device fp32 -> fp64
  int upload_t(const float* d, int rows, int cols, double** out);            // transposed copy
  int upload_cat(const float* a, int ra, const float* b, int rb, int cols, int col0, int ncols, bool transpose,
                 double** out);
  int gemm(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K,
           const double* bias, cudaStream_t s);
  int features_gemm(int m, cudaStream_t s);
  // C[M,N] (fp64) = A[M,K] (fp64) * B^T with B given as split-bf16 [N][K]: tcgen05 path for the per-step relevance GEMMs
  int gemm_tc(const double* A, int M, int K, const void* Bsplit, int N, double* C, cudaStream_t s);
  // the same contraction with A already in As_ as two bf16 planes of stride nA (written by the producing kernel) and the
  // result left in C32_ as fp32 [*, N] (read by the consuming kernel): one launch instead of three
  int gemm_tc_direct(int M, int K, size_t nA, const void* Bsplit, int N, cudaStream_t s);
  // forward GEMMs on tensor cores with fp32-exact operands (three bf16 planes, accumulator promoted every k-step):
  // C[M, N] = A[M, K] (row stride lda) * B^T (+ bias), B3 = planes of B^T [Npad, K] from split3_weights()
  int gemm_tc3(const double* A, int lda, int M, int K, const void* B3, int Npad, int N, const double* bias, double* C,
               int ldc, cudaStream_t s);
  int split3_weights(const double* d_Wt, int N, int K, int* Npad, void** out);
  // C32[Mpad, Npad] (fp32) = A3 (three bf16 planes [Mpad, K]) * B3^T: the bare tensor-core launch of the fused forward
  int tc3(const void* A3, int Mpad, int K, const void* B3, int Npad, float* C32, cudaStream_t s);
  int upload_split(const float* a, int ra, const float* b, int rb, int cols, int col0, int ncols, void** out);   // YF_[m*L, D] = UV_[m*L, H] * W_if^T (tensor cores when shapes allow)
  int sort_words(const int* h_word_img, const int* h_word_t, int n_words, cudaStream_t s);

  int kind_ = 0, V_ = 0, H_ = 0, E_ = 0, D_ = 0, L_ = 0, N_ = 0, T_ = 0, sos_ = 1, keras_logits_ = 0;
  int Kin1_ = 0, Kin2_ = 0;   // LSTM input widths incl. recurrent part (adaptive: Kin1 = 2E+H)
  long long launches_ = 0;
  bool tc_features_ = false;
  // CUDA graph of the greedy forward (decoder.cu: Decoder::forward); LRPCAP_DECODER_GRAPH=0 disables it
  bool graph_enabled_ = true;
  cudaGraphExec_t fwd_exec_ = nullptr;
  cudaStream_t graph_stream_ = nullptr;
  cudaEvent_t graph_ev_in_ = nullptr, graph_ev_out_ = nullptr;
  uint64_t fwd_key_ = 0, fwd_seen_key_ = 0;
  long long fwd_graph_launches_ = 0;
  bool tc_forward_ = false;                       // gate / logit GEMMs of the forward on tensor cores (gemm_tc3)
  // fused forward (decoder_fused.cuh): point-wise stages read the GEMM results in fp32 and write the next GEMM's operand
  // planes; [W_cat1 | W_sx] and [W_hp | W_ss] are single GEMMs; LRPCAP_DECODER_FUSED=0 restores the unfused sequence
  bool fused_ = false;
  void *W1cat3_ = nullptr, *W2cat3_ = nullptr, *WpTC3_ = nullptr;
  int Npad1_ = 0, Npad2_ = 0, Hpad_ = 0;
  DevBuf Axh_, Ahs_, Axh2_, Ahc_, C1_, C2_, C3_, C4_, Vf32_, pred_, F32_;
  void *WcatB1TC_ = nullptr, *WcatB2TC_ = nullptr;   // [Kin, 4H] split-bf16: B operand of the gradient decoder's GEMMs
  void *Wcat1TC3_ = nullptr, *Wcat2TC3_ = nullptr, *WoTC3_ = nullptr;
  int Vpad_ = 0, G4pad_ = 0;
  DevBuf As3_;
  void *Wgate1TC_ = nullptr, *Wgate2TC_ = nullptr;   // split-bf16 [Kin][H]: g-gate slices as K-major B operands
  DevBuf As_, C32_;
  void* WifTC_ = nullptr;   // split-bf16 [D][H]: K-major B operand of the image_features relevance GEMM
  DevBuf UVs_, YF32_, gemm_ws_;
  std::vector<void*> owned_;
  std::vector<size_t> owned_bytes_;
  bool rederive_ = false;
  size_t slot_cursor_ = 0;
  // weights (fp64, device)
  double *Wif_ = nullptr, *bif_ = nullptr, *WifT_ = nullptr, *Wgf_ = nullptr, *bgf_ = nullptr, *WgfT_ = nullptr;
  double *Emb_ = nullptr, *Wo_ = nullptr, *WoT_ = nullptr, *bo_ = nullptr;
  double *Wcat1_ = nullptr, *b1_ = nullptr, *Wcat2_ = nullptr, *b2_ = nullptr;     // [x;h] -> 4H
  double *Wgate1T_ = nullptr, *Wgate2T_ = nullptr;                                 // g-gate slice, transposed [H, Kin]
  double *Wcat1T_ = nullptr, *Wcat2T_ = nullptr;                                   // [4H, Kin] for the gradient decoder
  double *Wp_ = nullptr, *Whp_ = nullptr, *Wsx_ = nullptr, *Wss_ = nullptr, *Va_ = nullptr;
  // forward state
  DevBuf F_, Vp_, P_, a_, gp_, tok_, logitk_, logits_;
  DevBuf h1_, c1_, zg1_, ia1_, fa1_, ga1_, oa1_, h2_, c2_, zg2_, ia2_, fa2_, ga2_, oa2_;
  DevBuf ctx_, s_, chat_, alpha_, beta_, XH1_, XH2_;
  DevBuf Z_, hp_, sg_, sp_, e_, hc_;
  // word batch
  std::vector<int> order_, wimg_, wt_, nact_;
  DevBuf d_wimg_, d_wt_, d_order_;
  DevBuf Rh1_, Rh2_, Rh2n_, Rc1_, Rc2_, Rchat_, Rctx_, Rglob_, rword_, U_, Y_, Q_, UV_, YF_, ra_;
};

}  // namespace lrpcap
