// Decoder: batched forward with stored state + batched word-level relevance (see decoder.cuh).
#include "decoder.cuh"
#include "decoder_kernels.cuh"
#include "decoder_fused.cuh"
#include "tc_conv.cuh"
#include <cmath>
#include <cstdlib>
#include "../../include/lrpcap.h"
#include <algorithm>
#include <numeric>

namespace lrpcap {

using namespace dk;

namespace {
inline unsigned nblk(size_t n, int b) { return (unsigned)((n + b - 1) / b); }
}  // namespace

Decoder::~Decoder() {
  if (fwd_exec_) cudaGraphExecDestroy(fwd_exec_);
  if (graph_stream_) cudaStreamDestroy(graph_stream_);
  if (graph_ev_in_) cudaEventDestroy(graph_ev_in_);
  if (graph_ev_out_) cudaEventDestroy(graph_ev_out_);
  for (void* p : owned_) cudaFree(p);
  DevBuf* bufs[] = {&F_, &Vp_, &P_, &a_, &gp_, &tok_, &logitk_, &logits_, &h1_, &c1_, &zg1_, &ia1_, &fa1_, &ga1_, &oa1_,
                    &h2_, &c2_, &zg2_, &ia2_, &fa2_, &ga2_, &oa2_, &ctx_, &s_, &chat_, &alpha_, &beta_, &XH1_, &XH2_,
                    &Z_, &hp_, &sg_, &sp_, &e_, &hc_, &d_wimg_, &d_wt_, &d_order_, &Rh1_, &Rh2_, &Rh2n_, &Rc1_, &Rc2_,
                    &Rchat_, &Rctx_, &Rglob_, &rword_, &U_, &Y_, &Q_, &UV_, &YF_, &ra_, &UVs_, &YF32_, &gemm_ws_, &As_, &C32_, &As3_,
                    &Axh_, &Ahs_, &Axh2_, &Ahc_, &C1_, &C2_, &C3_, &C4_, &Vf32_, &pred_, &F32_};
  for (DevBuf* b : bufs) b->release();
}

// ---- weight derivation. Every derived tensor (fp64 copies, transposes, [x ; h] stacks, gate slices, bf16 planes) is
// produced on the device from fp32 source tensors that are already there: staged from the host by create(), or the
// caller's own device tensors in set_weights_device() (fine-tuning: the explained model changes every step, and a host
// round trip of ~30 M parameters plus their re-layout on one core was most of that step).
namespace {
// out (fp64 [R][ncols], or transposed [ncols][R]) = rows of a [ra, cols] stacked over rows of b [rb, cols], columns [col0, col0 + ncols)
__global__ void w_cat_slice_kernel(const float* __restrict__ a, int ra, const float* __restrict__ b, int rb, int cols,
                                   int col0, int ncols, int transpose, double* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int R = ra + rb;
  if (i >= (size_t)R * ncols) return;
  int r, c;
  if (transpose) { c = (int)(i / R); r = (int)(i - (size_t)c * R); }
  else { r = (int)(i / ncols); c = (int)(i - (size_t)r * ncols); }
  const float* src = (r < ra) ? a + (size_t)r * cols : b + (size_t)(r - ra) * cols;
  out[i] = (double)src[col0 + c];
}
// the same selection as two bf16 planes [R][ncols] (hi, then lo at + R * ncols)
__global__ void w_cat_slice_split_kernel(const float* __restrict__ a, int ra, const float* __restrict__ b, int rb, int cols,
                                         int col0, int ncols, __nv_bfloat16* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int R = ra + rb;
  const size_t n = (size_t)R * ncols;
  if (i >= n) return;
  const int r = (int)(i / ncols), c = (int)(i - (size_t)r * ncols);
  const float* src = (r < ra) ? a + (size_t)r * cols : b + (size_t)(r - ra) * cols;
  __nv_bfloat16 hi, lo;
  split_bf16(src[col0 + c], hi, lo);
  out[i] = hi;
  out[n + i] = lo;
}
}  // namespace

// A device allocation that survives set_weights_device(): the first derivation allocates, later ones walk the same list.
int Decoder::slot(size_t bytes, void** out) {
  if (rederive_) {
    LRPCAP_REQUIRE(slot_cursor_ < owned_.size() && owned_bytes_[slot_cursor_] == bytes, kErrState,
                   "decoder_set_weights_device: derived-tensor list changed (slot %zu)", slot_cursor_);
    *out = owned_[slot_cursor_++];
    return kOk;
  }
  void* p = nullptr;
  LRPCAP_CUDA(cudaMalloc(&p, bytes));
  owned_.push_back(p);
  owned_bytes_.push_back(bytes);
  *out = p;
  return kOk;
}

int Decoder::upload(const float* d, size_t n, double** out) {
  return upload_cat(d, (int)n, nullptr, 0, 1, 0, 1, false, out);
}

int Decoder::upload_t(const float* d, int rows, int cols, double** out) {
  return upload_cat(d, rows, nullptr, 0, cols, 0, cols, true, out);
}

// rows of a [ra, cols] stacked over rows of b [rb, cols], column slice [col0, col0+ncols); optionally transposed.
int Decoder::upload_cat(const float* a, int ra, const float* b, int rb, int cols, int col0, int ncols, bool transpose,
                        double** out) {
  LRPCAP_REQUIRE(a && (b || rb == 0), kErrInvalidArg, "decoder: missing weight tensor");
  const size_t n = (size_t)(ra + rb) * ncols;
  void* p = nullptr;
  LRPCAP_TRY(slot(n * sizeof(double), &p));
  w_cat_slice_kernel<<<nblk(n, 256), 256>>>(a, ra, b, rb, cols, col0, ncols, transpose ? 1 : 0, reinterpret_cast<double*>(p));
  LRPCAP_CUDA(cudaGetLastError());
  *out = reinterpret_cast<double*>(p);
  return kOk;
}

// rows of a stacked over rows of b, column slice -> split-bf16 [rows][ncols] (hi plane then lo plane)
int Decoder::upload_split(const float* a, int ra, const float* b, int rb, int cols, int col0, int ncols, void** out) {
  LRPCAP_REQUIRE(a && (b || rb == 0), kErrInvalidArg, "decoder: missing weight tensor");
  const size_t n = (size_t)(ra + rb) * ncols;
  void* p = nullptr;
  LRPCAP_TRY(slot(2 * n * sizeof(__nv_bfloat16), &p));
  w_cat_slice_split_kernel<<<nblk(n, 256), 256>>>(a, ra, b, rb, cols, col0, ncols, reinterpret_cast<__nv_bfloat16*>(p));
  LRPCAP_CUDA(cudaGetLastError());
  *out = p;
  return kOk;
}

int Decoder::gemm_tc(const double* A, int M, int K, const void* Bsplit, int N, double* C, cudaStream_t s) {
  if (M <= 0) return kOk;
  const int Hrows = (M + 15) / 16;             // rows laid out as a [1, Hrows, 16, K] "image" for the 1x1-conv kernel
  const size_t nA = (size_t)Hrows * 16 * K, nC = (size_t)Hrows * 16 * N;
  LRPCAP_TRY(As_.ensure(nA * 4));
  LRPCAP_TRY(C32_.ensure(nC * 4));
  __nv_bfloat16* hi = As_.as<__nv_bfloat16>();
  if ((size_t)M * K < nA) LRPCAP_CUDA(cudaMemsetAsync(As_.p, 0, nA * 4, s));   // zero the padded rows
  f64_to_split_strided_kernel<<<nblk((size_t)M * K, 256), 256, 0, s>>>(A, hi, hi + nA, (size_t)M * K);
  TcConvArgs a;
  a.A = As_.p; a.A_elems = nA; a.n_items = 1; a.H = Hrows; a.W = 16; a.C = K;
  a.B = Bsplit; a.B_elems = (size_t)N * K; a.taps = 1; a.Nout = N;
  a.epi.mode = EPI_RAW;
  a.epi.out_f32 = C32_.as<float>();
  LRPCAP_TRY(tc_conv_launch(a, s));
  f32_to_f64_kernel<<<nblk((size_t)M * N, 256), 256, 0, s>>>(C32_.as<float>(), C, (size_t)M * N);
  launches_ += 3;
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int Decoder::gemm_tc_direct(int M, int K, size_t nA, const void* Bsplit, int N, cudaStream_t s) {
  if (M <= 0) return kOk;
  const int Hrows = (M + 15) / 16;   // rows beyond M hold stale data: every output row depends on its own A row only
  TcConvArgs a;
  a.A = As_.p; a.A_elems = nA; a.n_items = 1; a.H = Hrows; a.W = 16; a.C = K;
  a.B = Bsplit; a.B_elems = (size_t)N * K; a.taps = 1; a.Nout = N;
  a.epi.mode = EPI_RAW;
  a.epi.out_f32 = C32_.as<float>();
  LRPCAP_TRY(tc_conv_launch(a, s));
  ++launches_;
  return kOk;
}

int Decoder::split3_weights(const double* d_Wt, int N, int K, int* Npad, void** out) {
  const int np = (N + 63) / 64 * 64;
  void* p = nullptr;
  LRPCAP_TRY(slot((size_t)3 * np * K * sizeof(__nv_bfloat16), &p));
  f64_rows_to_split3_kernel<<<nblk((size_t)np * K, 256), 256>>>(d_Wt, K, N, K, np, reinterpret_cast<__nv_bfloat16*>(p));
  LRPCAP_CUDA(cudaGetLastError());
  *Npad = np;
  *out = p;
  return kOk;
}

int Decoder::gemm_tc3(const double* A, int lda, int M, int K, const void* B3, int Npad, int N, const double* bias,
                      double* C, int ldc, cudaStream_t s) {
  if (M <= 0) return kOk;
  const int Mpad = (M + 15) / 16 * 16;            // rows laid out as a [1, Mpad/16, 16, K] "image" for the 1x1-conv kernel
  const size_t nA = (size_t)Mpad * K;
  LRPCAP_TRY(As3_.ensure(nA * 3 * sizeof(__nv_bfloat16)));
  LRPCAP_TRY(C32_.ensure((size_t)Mpad * Npad * sizeof(float)));
  f64_rows_to_split3_kernel<<<nblk(nA, 256), 256, 0, s>>>(A, lda, M, K, Mpad, As3_.as<__nv_bfloat16>());
  TcConvArgs a;
  a.A = As3_.p; a.A_elems = nA; a.n_items = 1; a.H = Mpad / 16; a.W = 16; a.C = K;
  a.B = B3; a.B_elems = (size_t)Npad * K; a.taps = 1; a.Nout = Npad;
  a.planes = 3;
  a.epi.mode = EPI_RAW;
  a.epi.out_f32 = C32_.as<float>();
  LRPCAP_TRY(tc_conv_launch(a, s));
  f32_to_f64_bias_kernel<<<nblk((size_t)M * N, 256), 256, 0, s>>>(C32_.as<float>(), Npad, C, ldc, M, N, bias);
  launches_ += 3;
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int Decoder::tc3(const void* A3, int Mpad, int K, const void* B3, int Npad, float* C32, cudaStream_t s) {
  TcConvArgs a;
  a.A = A3; a.A_elems = (size_t)Mpad * K; a.n_items = 1; a.H = Mpad / 16; a.W = 16; a.C = K;
  a.B = B3; a.B_elems = (size_t)Npad * K; a.taps = 1; a.Nout = Npad;
  a.planes = 3;
  a.epi.mode = EPI_RAW;
  a.epi.out_f32 = C32;
  LRPCAP_TRY(tc_conv_launch(a, s));
  ++launches_;
  return kOk;
}

int Decoder::create(Decoder** out, const lrpcap_decoder_weights* w, int sos_token, int keras_logits) {
  LRPCAP_REQUIRE(out && w, kErrInvalidArg, "decoder_create: null argument");
  LRPCAP_REQUIRE(w->kind == LRPCAP_DECODER_ADAPTIVE || w->kind == LRPCAP_DECODER_GRIDTD, kErrInvalidArg,
                 "decoder_create: unknown kind %d", w->kind);
  LRPCAP_REQUIRE(w->V > 0 && w->H > 0 && w->E > 0 && w->D > 0, kErrShape, "decoder_create: bad dimensions");
  LRPCAP_REQUIRE(sos_token >= 1 && sos_token <= w->V, kErrInvalidArg, "decoder_create: SOS id %d out of [1,%d]", sos_token, w->V);
  Decoder* d = new Decoder();
  d->kind_ = w->kind; d->V_ = w->V; d->H_ = w->H; d->E_ = w->E; d->D_ = w->D;
  d->sos_ = sos_token; d->keras_logits_ = keras_logits;
  if (const char* gv = getenv("LRPCAP_DECODER_GRAPH")) d->graph_enabled_ = gv[0] != '0';
  // stage the host tensors on the device as they are (fp32, Keras layouts); every derived layout is built there
  std::vector<WField> fl = d->fields();
  size_t total = 0;
  for (const WField& f : fl) total += (f.n + 3) / 4 * 4;
  DevBuf stage;
  lrpcap_decoder_weights wd = *w;
  int st = stage.ensure(total * sizeof(float));
  size_t off = 0;
  for (size_t k = 0; st == kOk && k < fl.size(); ++k) {
    const float* src = w->*(fl[k].p);
    if (!src) {
      set_last_error("decoder_create: weight tensor %zu missing", k);
      st = kErrInvalidArg;
      break;
    }
    float* dst = stage.as<float>() + off;
    if (cudaMemcpy(dst, src, fl[k].n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
      set_last_error("decoder_create: host -> device copy failed");
      st = kErrCuda;
      break;
    }
    wd.*(fl[k].p) = dst;
    off += (fl[k].n + 3) / 4 * 4;
  }
  if (st == kOk) st = d->derive(&wd);
  if (st == kOk && cudaDeviceSynchronize() != cudaSuccess) {
    set_last_error("decoder_create: weight derivation failed: %s", cudaGetErrorString(cudaGetLastError()));
    st = kErrCuda;
  }
  stage.release();
  if (st != kOk) {
    delete d;
    return st;
  }
  *out = d;
  return kOk;
}

// The weight tensors a decoder of this kind reads, with their element counts (include/lrpcap.h: lrpcap_decoder_weights_t).
std::vector<Decoder::WField> Decoder::fields() const {
  const size_t V = V_, H = H_, E = E_, D = D_;
  typedef lrpcap_decoder_weights W;
  std::vector<WField> f = {{&W::image_features_w, D * H}, {&W::image_features_b, H}, {&W::global_w, D * E}, {&W::global_b, E},
                           {&W::embedding, V * E}, {&W::output_w, H * V}, {&W::output_b, V}};
  if (kind_ == LRPCAP_DECODER_ADAPTIVE) {
    const WField a[] = {{&W::lstm_wi, 2 * E * 4 * H}, {&W::lstm_wh, H * 4 * H}, {&W::lstm_b, 4 * H}, {&W::Wv, H * H}, {&W::Wg, H * H},
                        {&W::Wx, 2 * E * H}, {&W::Wh, H * H}, {&W::Ws, H * H}, {&W::Vatt, H}};
    f.insert(f.end(), a, a + 9);
  } else {
    const WField g[] = {{&W::lang_wi, 2 * H * 4 * H}, {&W::lang_wh, H * 4 * H}, {&W::lang_b, 4 * H}, {&W::td_wi, (H + 2 * E) * 4 * H},
                        {&W::td_wh, H * 4 * H}, {&W::td_b, 4 * H}, {&W::W_va, H * H}, {&W::W_ha, H * H}, {&W::W_a, H},
                        {&W::W_x, (H + 2 * E) * H}, {&W::W_h, H * H}, {&W::W_s, H * H}};
    f.insert(f.end(), g, g + 12);
  }
  return f;
}

// Replaces the weights in place from DEVICE fp32 tensors (same layouts as lrpcap_decoder_create): every derived tensor is
// recomputed into the allocation it already has, so graphs and buffers stay valid; the forward state is dropped.
int Decoder::set_weights_device(const lrpcap_decoder_weights* wd) {
  LRPCAP_REQUIRE(wd != nullptr, kErrInvalidArg, "decoder_set_weights_device: null argument");
  LRPCAP_REQUIRE(wd->kind == kind_ && wd->V == V_ && wd->H == H_ && wd->E == E_ && wd->D == D_, kErrShape,
                 "decoder_set_weights_device: kind / dimensions differ from the handle's");
  for (const WField& f : fields())
    LRPCAP_REQUIRE(wd->*(f.p) != nullptr, kErrInvalidArg, "decoder_set_weights_device: a weight tensor is missing");
  LRPCAP_CUDA(cudaDeviceSynchronize());   // nothing may still read the old values
  rederive_ = true;
  slot_cursor_ = 0;
  const int st = derive(wd);
  rederive_ = false;
  LRPCAP_TRY(st);
  LRPCAP_CUDA(cudaDeviceSynchronize());
  N_ = 0;
  return kOk;
}

// Builds every derived weight layout from device fp32 tensors `w` (pointers are device memory).
int Decoder::derive(const lrpcap_decoder_weights* w) {
  const int V = V_, H = H_, E = E_, D = D_;
  LRPCAP_TRY(upload(w->image_features_w, (size_t)D * H, &Wif_));
  LRPCAP_TRY(upload(w->image_features_b, H, &bif_));
  LRPCAP_TRY(upload_t(w->image_features_w, D, H, &WifT_));
  LRPCAP_TRY(upload(w->global_w, (size_t)D * E, &Wgf_));
  LRPCAP_TRY(upload(w->global_b, E, &bgf_));
  LRPCAP_TRY(upload_t(w->global_w, D, E, &WgfT_));
  LRPCAP_TRY(upload(w->embedding, (size_t)V * E, &Emb_));
  LRPCAP_TRY(upload(w->output_w, (size_t)H * V, &Wo_));
  LRPCAP_TRY(upload_t(w->output_w, H, V, &WoT_));
  LRPCAP_TRY(upload(w->output_b, V, &bo_));
  if (w->kind == LRPCAP_DECODER_ADAPTIVE) {
    Kin1_ = 2 * E + H;
    LRPCAP_TRY(upload_cat(w->lstm_wi, 2 * E, w->lstm_wh, H, 4 * H, 0, 4 * H, false, &Wcat1_));
    LRPCAP_TRY(upload(w->lstm_b, 4 * H, &b1_));
    LRPCAP_TRY(upload_cat(w->lstm_wi, 2 * E, w->lstm_wh, H, 4 * H, 2 * H, H, true, &Wgate1T_));
    LRPCAP_TRY(upload_cat(w->lstm_wi, 2 * E, w->lstm_wh, H, 4 * H, 0, 4 * H, true, &Wcat1T_));
    LRPCAP_TRY(upload(w->Wv, (size_t)H * H, &Wp_));
    LRPCAP_TRY(upload(w->Wg, (size_t)H * H, &Whp_));
    LRPCAP_TRY(upload_cat(w->Wx, 2 * E, w->Wh, H, H, 0, H, false, &Wsx_));
    LRPCAP_TRY(upload(w->Ws, (size_t)H * H, &Wss_));
    LRPCAP_TRY(upload(w->Vatt, H, &Va_));
  } else {
    Kin1_ = 2 * H + 2 * E;
    Kin2_ = 3 * H;
    LRPCAP_TRY(upload_cat(w->td_wi, H + 2 * E, w->td_wh, H, 4 * H, 0, 4 * H, false, &Wcat1_));
    LRPCAP_TRY(upload(w->td_b, 4 * H, &b1_));
    LRPCAP_TRY(upload_cat(w->td_wi, H + 2 * E, w->td_wh, H, 4 * H, 2 * H, H, true, &Wgate1T_));
    LRPCAP_TRY(upload_cat(w->td_wi, H + 2 * E, w->td_wh, H, 4 * H, 0, 4 * H, true, &Wcat1T_));
    LRPCAP_TRY(upload_cat(w->lang_wi, 2 * H, w->lang_wh, H, 4 * H, 0, 4 * H, false, &Wcat2_));
    LRPCAP_TRY(upload(w->lang_b, 4 * H, &b2_));
    LRPCAP_TRY(upload_cat(w->lang_wi, 2 * H, w->lang_wh, H, 4 * H, 2 * H, H, true, &Wgate2T_));
    LRPCAP_TRY(upload_cat(w->lang_wi, 2 * H, w->lang_wh, H, 4 * H, 0, 4 * H, true, &Wcat2T_));
    LRPCAP_TRY(upload(w->W_va, (size_t)H * H, &Wp_));
    LRPCAP_TRY(upload(w->W_ha, (size_t)H * H, &Whp_));
    LRPCAP_TRY(upload_cat(w->W_x, H + 2 * E, w->W_h, H, H, 0, H, false, &Wsx_));
    LRPCAP_TRY(upload(w->W_s, (size_t)H * H, &Wss_));
    LRPCAP_TRY(upload(w->W_a, H, &Va_));
  }
  if (H % 64 == 0 && E % 64 == 0 && !getenv("LRPCAP_DECODER_FP64_GEMM")) {
    if (w->kind == LRPCAP_DECODER_ADAPTIVE) {
      LRPCAP_TRY(upload_split(w->lstm_wi, 2 * E, w->lstm_wh, H, 4 * H, 2 * H, H, &Wgate1TC_));
      LRPCAP_TRY(upload_split(w->lstm_wi, 2 * E, w->lstm_wh, H, 4 * H, 0, 4 * H, &WcatB1TC_));
    } else {
      LRPCAP_TRY(upload_split(w->td_wi, H + 2 * E, w->td_wh, H, 4 * H, 0, 4 * H, &WcatB1TC_));
      LRPCAP_TRY(upload_split(w->lang_wi, 2 * H, w->lang_wh, H, 4 * H, 0, 4 * H, &WcatB2TC_));
      LRPCAP_TRY(upload_split(w->td_wi, H + 2 * E, w->td_wh, H, 4 * H, 2 * H, H, &Wgate1TC_));
      LRPCAP_TRY(upload_split(w->lang_wi, 2 * H, w->lang_wh, H, 4 * H, 2 * H, H, &Wgate2TC_));
    }
  }
  if (H % 64 == 0 && D % 64 == 0 && !getenv("LRPCAP_DECODER_FP64_GEMM")) {
    // B operand of the [words*L, H] x [H, D] relevance GEMM: B[d][h] = W_if[d][h] -- the Keras (D, H) layout as is
    LRPCAP_TRY(upload_split(w->image_features_w, D, nullptr, 0, H, 0, H, &WifTC_));
    tc_features_ = true;
  }
  {
    const char* v = getenv("LRPCAP_DECODER_TC_FWD");
    const bool want = v ? (v[0] != '0') : true;   // LRPCAP_DECODER_TC_FWD=0: all forward GEMMs in fp64
    if (want && H % 64 == 0 && Kin1_ % 64 == 0 && (w->kind == LRPCAP_DECODER_ADAPTIVE || Kin2_ % 64 == 0)) {
      LRPCAP_TRY(split3_weights(Wcat1T_, 4 * H, Kin1_, &G4pad_, &Wcat1TC3_));
      if (w->kind == LRPCAP_DECODER_GRIDTD) LRPCAP_TRY(split3_weights(Wcat2T_, 4 * H, Kin2_, &G4pad_, &Wcat2TC3_));
      LRPCAP_TRY(split3_weights(WoT_, V, H, &Vpad_, &WoTC3_));
      tc_forward_ = true;
      const char* fv = getenv("LRPCAP_DECODER_FUSED");
      if (!(fv && fv[0] == '0') && D % 64 == 0) {
        // rows [W_cat1^T ; W_sx^T] (gates | sentinel gate) and [W_hp^T ; W_ss^T] as single B operands
        const bool ad = w->kind == LRPCAP_DECODER_ADAPTIVE;
        const int K1 = Kin1_;
        double *sxT = nullptr, *hpT = nullptr, *ssT = nullptr, *pT = nullptr, *cat = nullptr;
        LRPCAP_TRY(upload_cat(ad ? w->Wx : w->W_x, K1 - H, ad ? w->Wh : w->W_h, H, H, 0, H, true, &sxT));
        LRPCAP_TRY(upload_t(ad ? w->Wg : w->W_ha, H, H, &hpT));
        LRPCAP_TRY(upload_t(ad ? w->Ws : w->W_s, H, H, &ssT));
        LRPCAP_TRY(upload_t(ad ? w->Wv : w->W_va, H, H, &pT));
        void* catp = nullptr;
        LRPCAP_TRY(slot((size_t)5 * H * K1 * sizeof(double), &catp));
        cat = reinterpret_cast<double*>(catp);
        LRPCAP_CUDA(cudaMemcpy(cat, Wcat1T_, (size_t)4 * H * K1 * sizeof(double), cudaMemcpyDeviceToDevice));
        LRPCAP_CUDA(cudaMemcpy(cat + (size_t)4 * H * K1, sxT, (size_t)H * K1 * sizeof(double), cudaMemcpyDeviceToDevice));
        LRPCAP_TRY(split3_weights(cat, 5 * H, K1, &Npad1_, &W1cat3_));
        LRPCAP_CUDA(cudaMemcpy(cat, hpT, (size_t)H * H * sizeof(double), cudaMemcpyDeviceToDevice));
        LRPCAP_CUDA(cudaMemcpy(cat + (size_t)H * H, ssT, (size_t)H * H * sizeof(double), cudaMemcpyDeviceToDevice));
        LRPCAP_TRY(split3_weights(cat, 2 * H, H, &Npad2_, &W2cat3_));
        LRPCAP_TRY(split3_weights(pT, H, H, &Hpad_, &WpTC3_));
        fused_ = true;
      }
    }
  }
  return kOk;
}

int Decoder::gemm(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K,
                  const double* bias, cudaStream_t s) {
  if (M <= 0) return kOk;
  dim3 grid(nblk(N, GN), nblk(M, GM));
  const int blocks = (int)(grid.x * grid.y);
  int splits = 1;
  if (blocks < 148 && K >= 256) {   // skinny (decoder forward, M = #images): spread K over the SMs
    splits = (296 + blocks - 1) / blocks;
    if (splits > K / 64) splits = K / 64;
  }
  if (splits > 1) {
    const int kper = ((K + splits - 1) / splits + GK - 1) / GK * GK;
    splits = (K + kper - 1) / kper;
    LRPCAP_TRY(gemm_ws_.ensure((size_t)splits * M * N * 8));
    grid.z = splits;
    dgemm_kernel<<<grid, 256, 0, s>>>(A, lda, B, ldb, gemm_ws_.as<double>(), N, M, N, K, nullptr, kper);
    splitk_reduce_kernel<<<nblk((size_t)M * N, 256), 256, 0, s>>>(gemm_ws_.as<double>(), splits, C, ldc, M, N, bias);
    launches_ += 2;
    LRPCAP_CUDA(cudaGetLastError());
    return kOk;
  }
  dgemm_kernel<<<grid, 256, 0, s>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, K);
  ++launches_;
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

int Decoder::forward(const float* d_features, int N, int L, int* h_captions, int T, int greedy, int eos_token,
                     cudaStream_t s) {
  LRPCAP_REQUIRE(d_features && h_captions && N > 0 && L > 0 && T > 0, kErrInvalidArg, "decoder_forward: bad argument");
  const int H = H_, E = E_, D = D_, V = V_;
  const bool td = kind_ == LRPCAP_DECODER_GRIDTD;
  LRPCAP_REQUIRE(greedy >= 0 && greedy <= 2, kErrInvalidArg, "decoder_forward: greedy must be 0, 1 or 2");
  const bool predict = greedy == 2;   // teacher-forced inputs, arg-max outputs (what `model.predict` + argmax gives, train.py:570)
  // grid-TD: the Keras model's logits are (h2 + c_hat) W_o + b (models/model.py:816), the explainer's h2 W_o + b (quirk B1);
  // a prediction is the Keras model's by definition
  const int add_chat = (predict || keras_logits_) ? 1 : 0;
  if (greedy != 1)
    for (int i = 0; i < N * T; ++i)
      LRPCAP_REQUIRE(h_captions[i] >= 1 && h_captions[i] <= V, kErrInvalidArg,
                     "decoder_forward: token id %d at %d outside [1,%d]", h_captions[i], i, V);
  N_ = 0;
  const size_t NL = (size_t)N * L, st = (size_t)N * (T + 1) * H;
  LRPCAP_TRY(F_.ensure(NL * D * 8));
  LRPCAP_TRY(Vp_.ensure(NL * H * 8));
  LRPCAP_TRY(P_.ensure(NL * H * 8));
  LRPCAP_TRY(UV_.ensure(NL * H * 8));   // scratch for relu(Vp) while projecting
  LRPCAP_TRY(a_.ensure((size_t)N * D * 8));
  LRPCAP_TRY(gp_.ensure((size_t)N * E * 8));
  LRPCAP_TRY(tok_.ensure((size_t)N * T * sizeof(int)));
  LRPCAP_TRY(logitk_.ensure((size_t)N * T * 8));
  DevBuf* states[] = {&h1_, &c1_, &zg1_, &ia1_, &fa1_, &ga1_, &oa1_, &ctx_, &s_, &chat_};
  DevBuf* states2[] = {&h2_, &c2_, &zg2_, &ia2_, &fa2_, &ga2_, &oa2_};
  for (DevBuf* b : states) LRPCAP_TRY(b->ensure(st * 8));
  if (td) {
    for (DevBuf* b : states2) LRPCAP_TRY(b->ensure(st * 8));
    LRPCAP_TRY(XH2_.ensure((size_t)N * T * Kin2_ * 8));
  }
  LRPCAP_TRY(alpha_.ensure((size_t)N * (T + 1) * L * 8));
  LRPCAP_TRY(beta_.ensure((size_t)N * (T + 1) * 8));
  LRPCAP_TRY(XH1_.ensure((size_t)N * T * Kin1_ * 8));
  LRPCAP_TRY(Z_.ensure((size_t)N * 4 * H * 8));
  LRPCAP_TRY(hp_.ensure((size_t)N * H * 8));
  LRPCAP_TRY(sg_.ensure((size_t)N * H * 8));
  LRPCAP_TRY(sp_.ensure((size_t)N * H * 8));
  LRPCAP_TRY(e_.ensure((size_t)N * (L + 1) * 8));
  LRPCAP_TRY(hc_.ensure((size_t)N * H * 8));
  const bool fused = fused_ && L <= 256;
  const int Mpad = (N + 15) / 16 * 16, Mpad2 = (2 * N + 15) / 16 * 16;
  const size_t nAxh = (size_t)Mpad * Kin1_, nAhs = (size_t)Mpad2 * H, nAxh2 = (size_t)Mpad * (td ? Kin2_ : 0), nAhc = (size_t)Mpad * H;
  if (fused) {
    LRPCAP_TRY(Axh_.ensure(nAxh * 3 * sizeof(__nv_bfloat16)));
    LRPCAP_TRY(Ahs_.ensure(nAhs * 3 * sizeof(__nv_bfloat16)));
    LRPCAP_TRY(Ahc_.ensure(nAhc * 3 * sizeof(__nv_bfloat16)));
    if (td) LRPCAP_TRY(Axh2_.ensure(nAxh2 * 3 * sizeof(__nv_bfloat16)));
    LRPCAP_TRY(C1_.ensure((size_t)Mpad * Npad1_ * sizeof(float)));
    LRPCAP_TRY(C2_.ensure((size_t)Mpad2 * Npad2_ * sizeof(float)));
    if (td) LRPCAP_TRY(C3_.ensure((size_t)Mpad * G4pad_ * sizeof(float)));
    if (greedy) LRPCAP_TRY(C4_.ensure((size_t)Mpad * Vpad_ * sizeof(float)));
    LRPCAP_TRY(Vf32_.ensure(NL * H * sizeof(float)));
    // the startup GEMMs (M = N * L rows) go through gemm_tc3: size its scratch here, outside any graph capture
    const size_t mp = (NL + 15) / 16 * 16;
    LRPCAP_TRY(As3_.ensure(mp * std::max(D, H) * 3 * sizeof(__nv_bfloat16)));
    LRPCAP_TRY(C32_.ensure(mp * Hpad_ * sizeof(float)));
  }
  if (greedy) LRPCAP_TRY(logits_.ensure((size_t)N * V * 8));
  if (predict) LRPCAP_TRY(pred_.ensure((size_t)N * T * sizeof(int)));
  if (greedy != 1) LRPCAP_CUDA(cudaMemcpyAsync(tok_.p, h_captions, (size_t)N * T * sizeof(int), cudaMemcpyHostToDevice, s));

  // the caller's feature tensor may move between calls: it is converted outside the (pointer-baking) graph
  LRPCAP_TRY(F32_.ensure(NL * D * sizeof(float)));
  LRPCAP_CUDA(cudaMemcpyAsync(F32_.p, d_features, NL * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
  f32_to_f64_kernel<<<nblk(NL * D, 256), 256, 0, s>>>(d_features, F_.as<double>(), NL * D);
  ++launches_;
  // everything else the forward enqueues on the stream (memsets + ~25 small launches per step)
  auto issue = [&](cudaStream_t s) -> int {   // (parameter shadows the caller's stream on purpose)
  for (DevBuf* b : states) LRPCAP_CUDA(cudaMemsetAsync(b->p, 0, st * 8, s));
  if (td)
    for (DevBuf* b : states2) LRPCAP_CUDA(cudaMemsetAsync(b->p, 0, st * 8, s));
  LRPCAP_CUDA(cudaMemsetAsync(alpha_.p, 0, (size_t)N * (T + 1) * L * 8, s));
  LRPCAP_CUDA(cudaMemsetAsync(beta_.p, 0, (size_t)N * (T + 1) * 8, s));
  double *F = F_.as<double>(), *Vp = Vp_.as<double>(), *P = P_.as<double>(), *Vf = UV_.as<double>();
  // image_features / global_img_feature heads (explainers.py:375-388); F was filled from d_features by the caller part
  // Vp stays an fp64 GEMM rounded to float32: its sign is a ReLU decision (gradient decoder: d_V masked where Vf <= 0) and
  // the reference's float32 value is reproduced bit for bit this way; P = Vf W_p only enters tanh(P + hp) and may take the
  // tensor-core path
  LRPCAP_TRY(gemm(F, D, Wif_, H, Vp, H, (int)NL, H, D, bif_, s));
  round_f32_kernel<<<nblk(NL * H, 256), 256, 0, s>>>(Vp, NL * H);   // rows are float32 results in the reference
  relu_copy_kernel<<<nblk(NL * H, 256), 256, 0, s>>>(Vp, Vf, NL * H);
  if (fused) {
    LRPCAP_TRY(gemm_tc3(Vf, H, (int)NL, H, WpTC3_, Hpad_, H, nullptr, P, H, s));
    relu_f32_copy_kernel<<<nblk(NL * H, 256), 256, 0, s>>>(Vp, Vf32_.as<float>(), NL * H);
    ++launches_;
    LRPCAP_CUDA(cudaMemsetAsync(Axh_.p, 0, nAxh * 3 * sizeof(__nv_bfloat16), s));   // rows >= N of the A operands stay zero
    LRPCAP_CUDA(cudaMemsetAsync(Ahs_.p, 0, nAhs * 3 * sizeof(__nv_bfloat16), s));
    LRPCAP_CUDA(cudaMemsetAsync(Ahc_.p, 0, nAhc * 3 * sizeof(__nv_bfloat16), s));
    if (td) LRPCAP_CUDA(cudaMemsetAsync(Axh2_.p, 0, nAxh2 * 3 * sizeof(__nv_bfloat16), s));
  } else {
    LRPCAP_TRY(gemm(Vf, H, Wp_, H, P, H, (int)NL, H, H, nullptr, s));
  }
  mean_feat_kernel<<<N, 256, 0, s>>>(F, a_.as<double>(), L, D);
  LRPCAP_TRY(gemm(a_.as<double>(), D, Wgf_, E, gp_.as<double>(), E, N, E, D, bgf_, s));
  round_f32_kernel<<<nblk((size_t)N * E, 256), 256, 0, s>>>(gp_.as<double>(), (size_t)N * E);
  launches_ += 4;

  int* tok = tok_.as<int>();
  int* tok_out = predict ? pred_.as<int>() : tok;   // greedy feeds its arg-max back; predict keeps the teacher tokens as inputs
  for (int i = 0; fused && i < T; ++i) {   // fused step: 9 launches (grid-TD, greedy); see decoder_fused.cuh
    __nv_bfloat16 *Axh = Axh_.as<__nv_bfloat16>(), *Ahs = Ahs_.as<__nv_bfloat16>(), *Ahc = Ahc_.as<__nv_bfloat16>();
    __nv_bfloat16* Axh2 = td ? Axh2_.as<__nv_bfloat16>() : nullptr;
    fwd_xh_kernel<<<N, 256, 0, s>>>(XH1_.as<double>(), Axh, nAxh, Emb_, gp_.as<double>(), h1_.as<double>(),
                                    td ? h2_.as<double>() : nullptr, tok, i, T, H, E, sos_, td ? 1 : 0);
    LRPCAP_TRY(tc3(Axh, Mpad, Kin1_, W1cat3_, Npad1_, C1_.as<float>(), s));
    fwd_lstm_kernel<<<N, 256, 0, s>>>(C1_.as<float>(), Npad1_, b1_, h1_.as<double>(), c1_.as<double>(), zg1_.as<double>(),
                                      ia1_.as<double>(), fa1_.as<double>(), ga1_.as<double>(), oa1_.as<double>(),
                                      s_.as<double>(), Ahs, nAhs, N, nullptr, nullptr, nullptr, 0, 0, i, T, H);
    LRPCAP_TRY(tc3(Ahs, Mpad2, H, W2cat3_, Npad2_, C2_.as<float>(), s));
    fwd_scores_kernel<<<dim3((L + 1 + 7) / 8, N), 256, 0, s>>>(P, C2_.as<float>(), Npad2_, N, Va_, e_.as<double>(), L, H);
    fwd_ctx_kernel<<<dim3(N, H / 64), 256, 0, s>>>(Vf32_.as<float>(), e_.as<double>(), alpha_.as<double>(), beta_.as<double>(),
                                                   s_.as<double>(), ctx_.as<double>(), chat_.as<double>(), h1_.as<double>(),
                                                   td ? h2_.as<double>() : nullptr, td ? nullptr : hc_.as<double>(),
                                                   td ? nullptr : Ahc, nAhc, td ? XH2_.as<double>() : nullptr, Axh2, nAxh2,
                                                   i, T, L, H);
    launches_ += 4;
    if (td) {
      LRPCAP_TRY(tc3(Axh2, Mpad, Kin2_, Wcat2TC3_, G4pad_, C3_.as<float>(), s));
      fwd_lstm_kernel<<<N, 256, 0, s>>>(C3_.as<float>(), G4pad_, b2_, h2_.as<double>(), c2_.as<double>(), zg2_.as<double>(),
                                        ia2_.as<double>(), fa2_.as<double>(), ga2_.as<double>(), oa2_.as<double>(), nullptr,
                                        nullptr, 0, N, chat_.as<double>(), hc_.as<double>(), Ahc, nAhc, add_chat, i, T, H);
      ++launches_;
    }
    if (greedy) {
      LRPCAP_TRY(tc3(Ahc, Mpad, H, WoTC3_, Vpad_, C4_.as<float>(), s));
      fwd_argmax_kernel<<<N, 256, 0, s>>>(C4_.as<float>(), Vpad_, bo_, V, eos_token >= 1 ? eos_token - 1 : -1, tok_out,
                                          logitk_.as<double>(), i, T);
    } else {
      logitk_kernel<<<N, 128, 0, s>>>(hc_.as<double>(), WoT_, bo_, tok, logitk_.as<double>(), i, T, H);
    }
    ++launches_;
  }
  for (int i = 0; !fused && i < T; ++i) {
    build_xh_kernel<<<N, 256, 0, s>>>(XH1_.as<double>(), Emb_, gp_.as<double>(), h1_.as<double>(),
                                      td ? h2_.as<double>() : nullptr, tok, i, T, H, E, sos_, td ? 1 : 0);
    const double* xh = XH1_.as<double>() + (size_t)i * Kin1_;
    if (tc_forward_) LRPCAP_TRY(gemm_tc3(xh, T * Kin1_, N, Kin1_, Wcat1TC3_, G4pad_, 4 * H, b1_, Z_.as<double>(), 4 * H, s));
    else LRPCAP_TRY(gemm(xh, T * Kin1_, Wcat1_, 4 * H, Z_.as<double>(), 4 * H, N, 4 * H, Kin1_, b1_, s));
    lstm_point_kernel<<<N, 256, 0, s>>>(Z_.as<double>(), h1_.as<double>(), c1_.as<double>(), zg1_.as<double>(),
                                        ia1_.as<double>(), fa1_.as<double>(), ga1_.as<double>(), oa1_.as<double>(), i, T, H);
    const double* h_new = h1_.as<double>() + (size_t)(i + 1) * H;
    LRPCAP_TRY(gemm(h_new, (T + 1) * H, Whp_, H, hp_.as<double>(), H, N, H, H, nullptr, s));
    LRPCAP_TRY(gemm(xh, T * Kin1_, Wsx_, H, sg_.as<double>(), H, N, H, Kin1_, nullptr, s));
    sentinel_kernel<<<N, 256, 0, s>>>(sg_.as<double>(), c1_.as<double>(), s_.as<double>(), i, T, H);
    const double* s_new = s_.as<double>() + (size_t)(i + 1) * H;
    LRPCAP_TRY(gemm(s_new, (T + 1) * H, Wss_, H, sp_.as<double>(), H, N, H, H, nullptr, s));
    scores_kernel<<<dim3(L + 1, N), 128, 0, s>>>(P, hp_.as<double>(), sp_.as<double>(), Va_, e_.as<double>(), L, H);
    softmax_kernel<<<N, 256, 0, s>>>(e_.as<double>(), alpha_.as<double>(), beta_.as<double>(), i, T, L);
    context_kernel<<<dim3(N, (H + 127) / 128), 128, 0, s>>>(Vf, alpha_.as<double>(), beta_.as<double>(), s_.as<double>(), ctx_.as<double>(),
                                     chat_.as<double>(), h1_.as<double>(), td ? nullptr : hc_.as<double>(), i, T, L, H);
    launches_ += 6;
    if (td) {
      build_xh2_kernel<<<N, 256, 0, s>>>(XH2_.as<double>(), chat_.as<double>(), h1_.as<double>(), h2_.as<double>(), i, T, H);
      if (tc_forward_)
        LRPCAP_TRY(gemm_tc3(XH2_.as<double>() + (size_t)i * Kin2_, T * Kin2_, N, Kin2_, Wcat2TC3_, G4pad_, 4 * H, b2_,
                            Z_.as<double>(), 4 * H, s));
      else
        LRPCAP_TRY(gemm(XH2_.as<double>() + (size_t)i * Kin2_, T * Kin2_, Wcat2_, 4 * H, Z_.as<double>(), 4 * H, N, 4 * H,
                        Kin2_, b2_, s));
      lstm_point_kernel<<<N, 256, 0, s>>>(Z_.as<double>(), h2_.as<double>(), c2_.as<double>(), zg2_.as<double>(),
                                          ia2_.as<double>(), fa2_.as<double>(), ga2_.as<double>(), oa2_.as<double>(), i, T, H);
      gridtd_hc_kernel<<<N, 256, 0, s>>>(h2_.as<double>(), chat_.as<double>(), hc_.as<double>(), i, T, H, add_chat);
      launches_ += 3;
    }
    if (greedy) {
      if (tc_forward_) LRPCAP_TRY(gemm_tc3(hc_.as<double>(), H, N, H, WoTC3_, Vpad_, V, bo_, logits_.as<double>(), V, s));
      else LRPCAP_TRY(gemm(hc_.as<double>(), H, Wo_, V, logits_.as<double>(), V, N, V, H, bo_, s));
      argmax_kernel<<<N, 256, 0, s>>>(logits_.as<double>(), V, eos_token >= 1 ? eos_token - 1 : -1, tok_out,
                                      logitk_.as<double>(), i, T);
    } else {
      logitk_kernel<<<N, 128, 0, s>>>(hc_.as<double>(), WoT_, bo_, tok, logitk_.as<double>(), i, T, H);
    }
    ++launches_;
  }
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
  };

  // The greedy loop is a fixed sequence of ~500 small launches for a given (N, T, buffers): the first call with a
  // configuration runs it eagerly (and sizes every scratch buffer), the second captures it into a CUDA graph, later ones
  // replay the graph -- the launch gaps between its latency-bound kernels are a third of this phase.
  bool done = false;
  if (greedy && graph_enabled_) {
    uint64_t key = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { key = (key ^ v) * 1099511628211ull; };
    mix((uint64_t)N); mix((uint64_t)T); mix((uint64_t)L); mix((uint64_t)(int64_t)eos_token); mix((uint64_t)greedy);
    mix((uint64_t)(uintptr_t)s);
    DevBuf* all[] = {&F_, &Vp_, &P_, &UV_, &a_, &gp_, &tok_, &logitk_, &h1_, &c1_, &zg1_, &ia1_, &fa1_, &ga1_, &oa1_, &ctx_, &s_,
                     &chat_, &h2_, &c2_, &zg2_, &ia2_, &fa2_, &ga2_, &oa2_, &XH1_, &XH2_, &alpha_, &beta_, &Z_, &hp_, &sg_,
                     &sp_, &e_, &hc_, &logits_, &pred_, &gemm_ws_, &As3_, &C32_, &Axh_, &Ahs_, &Axh2_, &Ahc_, &C1_, &C2_, &C3_, &C4_, &Vf32_};
    for (DevBuf* b : all) mix((uint64_t)(uintptr_t)b->p);
    // the legacy default stream cannot be captured: graphs run on a private stream ordered after / before it by events
    cudaStream_t gs = s;
    if (!s) {
      if (!graph_stream_) {
        LRPCAP_CUDA(cudaStreamCreateWithFlags(&graph_stream_, cudaStreamNonBlocking));
        LRPCAP_CUDA(cudaEventCreateWithFlags(&graph_ev_in_, cudaEventDisableTiming));
        LRPCAP_CUDA(cudaEventCreateWithFlags(&graph_ev_out_, cudaEventDisableTiming));
      }
      gs = graph_stream_;
    }
    auto launch_graph = [&]() -> int {
      if (!s) {
        LRPCAP_CUDA(cudaEventRecord(graph_ev_in_, s));
        LRPCAP_CUDA(cudaStreamWaitEvent(gs, graph_ev_in_, 0));
      }
      LRPCAP_CUDA(cudaGraphLaunch(fwd_exec_, gs));
      if (!s) {
        LRPCAP_CUDA(cudaEventRecord(graph_ev_out_, gs));
        LRPCAP_CUDA(cudaStreamWaitEvent(s, graph_ev_out_, 0));
      }
      return kOk;
    };
    if (fwd_exec_ && key == fwd_key_) {
      LRPCAP_TRY(launch_graph());
      launches_ += fwd_graph_launches_;
      done = true;
    } else if (key == fwd_seen_key_) {
      if (fwd_exec_) { cudaGraphExecDestroy(fwd_exec_); fwd_exec_ = nullptr; }
      const long long l0 = launches_;
      cudaGraph_t g = nullptr;
      if (cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int ist = issue(gs);
        const cudaError_t ce = cudaStreamEndCapture(gs, &g);
        if (ist == kOk && ce == cudaSuccess && g && cudaGraphInstantiate(&fwd_exec_, g, 0) == cudaSuccess) {
          fwd_key_ = key;
          fwd_graph_launches_ = launches_ - l0;
          LRPCAP_TRY(launch_graph());
          done = true;
        } else {
          fwd_exec_ = nullptr;
          launches_ = l0;
          cudaGetLastError();      // a failed capture leaves a sticky-free error behind; fall back to eager launches
          graph_enabled_ = false;
        }
        if (g) cudaGraphDestroy(g);
      }
    } else {
      fwd_seen_key_ = key;
    }
  }
  if (!done) LRPCAP_TRY(issue(s));
  if (greedy) {
    LRPCAP_CUDA(cudaMemcpyAsync(h_captions, predict ? pred_.p : tok_.p, (size_t)N * T * sizeof(int), cudaMemcpyDeviceToHost, s));
    LRPCAP_CUDA(cudaStreamSynchronize(s));
  }
  // predict mode: the stored state belongs to the teacher tokens while logit_k holds the arg-max logits -- not a state to
  // explain from; the caller runs a teacher-forced forward on the predicted caption next (models/model.py:1661-1672)
  N_ = predict ? 0 : N; T_ = T; L_ = L;
  return kOk;
}

int Decoder::caption_logits(double* h_logit) {
  LRPCAP_REQUIRE(N_ > 0 && h_logit, kErrState, "decoder_caption_logits: call decoder_forward first");
  LRPCAP_CUDA(cudaMemcpy(h_logit, logitk_.p, (size_t)N_ * T_ * 8, cudaMemcpyDeviceToHost));
  return kOk;
}

int Decoder::last_logits(double* h_logits, cudaStream_t s) {
  LRPCAP_REQUIRE(N_ > 0 && h_logits, kErrState, "decoder_last_logits: call decoder_forward first");
  LRPCAP_TRY(logits_.ensure((size_t)N_ * V_ * 8));
  LRPCAP_TRY(gemm(hc_.as<double>(), H_, Wo_, V_, logits_.as<double>(), V_, N_, V_, H_, bo_, s));
  LRPCAP_CUDA(cudaMemcpyAsync(h_logits, logits_.p, (size_t)N_ * V_ * 8, cudaMemcpyDeviceToHost, s));
  LRPCAP_CUDA(cudaStreamSynchronize(s));
  return kOk;
}

int Decoder::attention(float* h_alpha, float* h_beta) {
  LRPCAP_REQUIRE(N_ > 0, kErrState, "decoder_attention: call decoder_forward first");
  if (h_alpha) {
    std::vector<double> al((size_t)N_ * (T_ + 1) * L_);
    LRPCAP_CUDA(cudaMemcpy(al.data(), alpha_.p, al.size() * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < al.size(); ++i) h_alpha[i] = (float)al[i];
  }
  if (h_beta) {
    std::vector<double> be((size_t)N_ * (T_ + 1));
    LRPCAP_CUDA(cudaMemcpy(be.data(), beta_.p, be.size() * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < be.size(); ++i) h_beta[i] = (float)be[i];
  }
  return kOk;
}

int Decoder::features_gemm(int m, cudaStream_t s) {
  const int H = H_, D = D_, L = L_;
  const int fh = (int)std::lround(std::sqrt((double)L));
  if (tc_features_ && fh * fh == L) {
    const size_t nA = (size_t)m * L * H, nO = (size_t)m * L * D;
    LRPCAP_TRY(UVs_.ensure(nA * 4));
    LRPCAP_TRY(YF32_.ensure(nO * 4));
    __nv_bfloat16* hi = UVs_.as<__nv_bfloat16>();
    f64_to_split_kernel<<<nblk(nA, 256), 256, 0, s>>>(UV_.as<double>(), hi, hi + nA, nA);
    TcConvArgs a;
    a.A = UVs_.p; a.A_elems = nA; a.n_items = m; a.H = fh; a.W = fh; a.C = H;
    a.B = WifTC_; a.B_elems = (size_t)D * H; a.taps = 1; a.Nout = D;
    a.epi.mode = EPI_RAW;
    a.epi.out_f32 = YF32_.as<float>();
    LRPCAP_TRY(tc_conv_launch(a, s));
    f32_to_f64_kernel<<<nblk(nO, 256), 256, 0, s>>>(YF32_.as<float>(), YF_.as<double>(), nO);
    launches_ += 3;
    LRPCAP_CUDA(cudaGetLastError());
    return kOk;
  }
  return gemm(UV_.as<double>(), H, WifT_, D, YF_.as<double>(), D, m * L, D, H, nullptr, s);
}

// Words sorted by position (descending) so that at time-step i the active words (t > i) are a prefix.
int Decoder::sort_words(const int* h_word_img, const int* h_word_t, int W, cudaStream_t s) {
  LRPCAP_REQUIRE(N_ > 0, kErrState, "decoder: call decoder_forward first");
  LRPCAP_REQUIRE(h_word_img && h_word_t && W > 0, kErrInvalidArg, "decoder: bad word list");
  for (int w = 0; w < W; ++w) {
    LRPCAP_REQUIRE(h_word_img[w] >= 0 && h_word_img[w] < N_, kErrInvalidArg, "decoder: word %d image %d out of range", w, h_word_img[w]);
    // reference: NotImplementedError('index out of range of captions') when t > len(xt) (explainers.py:538-539)
    LRPCAP_REQUIRE(h_word_t[w] >= 1 && h_word_t[w] <= T_, kErrInvalidArg, "decoder: index out of range of captions (t=%d, T=%d)", h_word_t[w], T_);
  }
  order_.resize(W);
  std::iota(order_.begin(), order_.end(), 0);
  std::stable_sort(order_.begin(), order_.end(), [&](int a, int b) { return h_word_t[a] > h_word_t[b]; });
  wimg_.resize(W); wt_.resize(W);
  for (int p = 0; p < W; ++p) { wimg_[p] = h_word_img[order_[p]]; wt_[p] = h_word_t[order_[p]]; }
  nact_.assign(T_, 0);
  for (int i = 0; i < T_; ++i) {
    int c = 0;
    while (c < W && wt_[c] > i) ++c;
    nact_[i] = c;
  }
  LRPCAP_TRY(d_wimg_.ensure((size_t)W * sizeof(int)));
  LRPCAP_TRY(d_wt_.ensure((size_t)W * sizeof(int)));
  LRPCAP_TRY(d_order_.ensure((size_t)W * sizeof(int)));
  LRPCAP_CUDA(cudaMemcpyAsync(d_wimg_.p, wimg_.data(), (size_t)W * sizeof(int), cudaMemcpyHostToDevice, s));
  LRPCAP_CUDA(cudaMemcpyAsync(d_wt_.p, wt_.data(), (size_t)W * sizeof(int), cudaMemcpyHostToDevice, s));
  LRPCAP_CUDA(cudaMemcpyAsync(d_order_.p, order_.data(), (size_t)W * sizeof(int), cudaMemcpyHostToDevice, s));
  return kOk;
}

int Decoder::relevance(const int* h_word_img, const int* h_word_t, int W, float* d_R_head, double* h_r_words,
                       float* h_attention, cudaStream_t s) {
  LRPCAP_REQUIRE(d_R_head != nullptr, kErrInvalidArg, "decoder_relevance: null output");
  LRPCAP_TRY(sort_words(h_word_img, h_word_t, W, s));
  const int H = H_, E = E_, D = D_, L = L_, T = T_;
  const bool td = kind_ == LRPCAP_DECODER_GRIDTD;
  const size_t WH = (size_t)W * H;
  DevBuf* rb[] = {&Rh1_, &Rc1_, &Rctx_};
  for (DevBuf* b : rb) LRPCAP_TRY(b->ensure(WH * 8));
  LRPCAP_TRY(U_.ensure((size_t)W * std::max(H, E) * 8));
  LRPCAP_TRY(Rglob_.ensure((size_t)W * E * 8));
  LRPCAP_TRY(rword_.ensure((size_t)W * T * 8));
  LRPCAP_TRY(Y_.ensure((size_t)W * std::max(Kin1_, std::max(Kin2_, D)) * 8));
  LRPCAP_TRY(ra_.ensure((size_t)W * D * 8));
  LRPCAP_CUDA(cudaMemsetAsync(Rglob_.p, 0, (size_t)W * E * 8, s));
  LRPCAP_CUDA(cudaMemsetAsync(rword_.p, 0, (size_t)W * T * 8, s));
  LRPCAP_CUDA(cudaMemsetAsync(Rc1_.p, 0, WH * 8, s));
  if (td) {
    DevBuf* rb2[] = {&Rh2_, &Rh2n_, &Rc2_, &Rchat_};
    for (DevBuf* b : rb2) LRPCAP_TRY(b->ensure(WH * 8));
    LRPCAP_TRY(Q_.ensure((size_t)W * T * H * 8));
    LRPCAP_CUDA(cudaMemsetAsync(Rc2_.p, 0, WH * 8, s));
    LRPCAP_CUDA(cudaMemsetAsync(Rh1_.p, 0, WH * 8, s));
  }
  WordRef wr{d_wimg_.as<int>(), d_wt_.as<int>()};
  const int* tok = tok_.as<int>();
  // direct path: the cell kernel writes the gate GEMM's bf16 planes, the scatter kernels read its fp32 result
  const bool direct = Wgate1TC_ && (!td || Wgate2TC_) && !getenv("LRPCAP_DECODER_FUSED_OFF");
  const size_t nAs = (size_t)((W + 15) / 16) * 16 * H;
  if (direct) {
    LRPCAP_TRY(As_.ensure(nAs * 4));
    LRPCAP_TRY(C32_.ensure((size_t)((W + 15) / 16) * 16 * std::max(Kin1_, Kin2_) * 4));
  }
  __nv_bfloat16* Us = direct ? As_.as<__nv_bfloat16>() : nullptr;

  if (!td) {
    lrp_init_kernel<<<W, 256, 0, s>>>(wr, h1_.as<double>(), chat_.as<double>(), ctx_.as<double>(), s_.as<double>(),
                                      beta_.as<double>(), logitk_.as<double>(), WoT_, tok, Rh1_.as<double>(),
                                      Rc1_.as<double>(), Rctx_.as<double>(), nullptr, T, H, 0);
    ++launches_;
    for (int i = T - 1; i >= 0; --i) {
      const int na = nact_[i];
      if (na == 0) continue;
      const size_t nA = (size_t)((na + 15) / 16) * 16 * H;
      lrp_cell_kernel<<<na, 256, 0, s>>>(wr, i, ia1_.as<double>(), fa1_.as<double>(), zg1_.as<double>(), c1_.as<double>(),
                                         Rc1_.as<double>(), Rh1_.as<double>(), nullptr, U_.as<double>(), T, H, Us, nA);
      if (direct) {
        LRPCAP_TRY(gemm_tc_direct(na, H, nA, Wgate1TC_, Kin1_, s));
        lrp_scatter_adaptive_kernel<float><<<na, 256, 0, s>>>(wr, i, XH1_.as<double>(), C32_.as<float>(), Rh1_.as<double>(),
                                                              Rglob_.as<double>(), rword_.as<double>(), T, H, E);
      } else {
        if (Wgate1TC_) LRPCAP_TRY(gemm_tc(U_.as<double>(), na, H, Wgate1TC_, Kin1_, Y_.as<double>(), s));
        else LRPCAP_TRY(gemm(U_.as<double>(), H, Wgate1T_, Kin1_, Y_.as<double>(), Kin1_, na, Kin1_, H, nullptr, s));
        lrp_scatter_adaptive_kernel<double><<<na, 256, 0, s>>>(wr, i, XH1_.as<double>(), Y_.as<double>(), Rh1_.as<double>(),
                                                               Rglob_.as<double>(), rword_.as<double>(), T, H, E);
      }
      launches_ += 2;
    }
  } else {
    lrp_init_kernel<<<W, 256, 0, s>>>(wr, h2_.as<double>(), chat_.as<double>(), ctx_.as<double>(), s_.as<double>(),
                                      beta_.as<double>(), logitk_.as<double>(), WoT_, tok, Rh2_.as<double>(), nullptr,
                                      nullptr, Rchat_.as<double>(), T, H, 1);
    ++launches_;
    for (int i = T - 1; i >= 0; --i) {
      const int na = nact_[i];
      if (na == 0) continue;
      const size_t nA = (size_t)((na + 15) / 16) * 16 * H;
      lrp_cell_kernel<<<na, 256, 0, s>>>(wr, i, ia2_.as<double>(), fa2_.as<double>(), zg2_.as<double>(), c2_.as<double>(),
                                         Rc2_.as<double>(), Rh2_.as<double>(), nullptr, U_.as<double>(), T, H, Us, nA);
      // Rctx_ doubles as the "extra" (r_s + R_h1) buffer of the top-down cell
      if (direct) {
        LRPCAP_TRY(gemm_tc_direct(na, H, nA, Wgate2TC_, Kin2_, s));
        lrp_scatter_lang_kernel<float><<<na, 256, 0, s>>>(wr, i, XH2_.as<double>(), C32_.as<float>(), chat_.as<double>(),
                                                          ctx_.as<double>(), s_.as<double>(), beta_.as<double>(),
                                                          Rchat_.as<double>(), Rh1_.as<double>(), Rh2n_.as<double>(),
                                                          Rctx_.as<double>(), Q_.as<double>(), T, H);
      } else {
        if (Wgate2TC_) LRPCAP_TRY(gemm_tc(U_.as<double>(), na, H, Wgate2TC_, Kin2_, Y_.as<double>(), s));
        else LRPCAP_TRY(gemm(U_.as<double>(), H, Wgate2T_, Kin2_, Y_.as<double>(), Kin2_, na, Kin2_, H, nullptr, s));
        lrp_scatter_lang_kernel<double><<<na, 256, 0, s>>>(wr, i, XH2_.as<double>(), Y_.as<double>(), chat_.as<double>(),
                                                           ctx_.as<double>(), s_.as<double>(), beta_.as<double>(),
                                                           Rchat_.as<double>(), Rh1_.as<double>(), Rh2n_.as<double>(),
                                                           Rctx_.as<double>(), Q_.as<double>(), T, H);
      }
      // top-down cell: Rc1 += extra (Rh1 is already folded into extra, so pass a zero-free "Rh" = extra, extra = null)
      lrp_cell_kernel<<<na, 256, 0, s>>>(wr, i, ia1_.as<double>(), fa1_.as<double>(), zg1_.as<double>(), c1_.as<double>(),
                                         Rc1_.as<double>(), Rctx_.as<double>(), nullptr, U_.as<double>(), T, H, Us, nA);
      if (direct) {
        LRPCAP_TRY(gemm_tc_direct(na, H, nA, Wgate1TC_, Kin1_, s));
        lrp_scatter_td_kernel<float><<<na, 256, 0, s>>>(wr, i, XH1_.as<double>(), C32_.as<float>(), Rh2n_.as<double>(),
                                                        Rh2_.as<double>(), Rh1_.as<double>(), Rglob_.as<double>(),
                                                        rword_.as<double>(), T, H, E);
      } else {
        if (Wgate1TC_) LRPCAP_TRY(gemm_tc(U_.as<double>(), na, H, Wgate1TC_, Kin1_, Y_.as<double>(), s));
        else LRPCAP_TRY(gemm(U_.as<double>(), H, Wgate1T_, Kin1_, Y_.as<double>(), Kin1_, na, Kin1_, H, nullptr, s));
        lrp_scatter_td_kernel<double><<<na, 256, 0, s>>>(wr, i, XH1_.as<double>(), Y_.as<double>(), Rh2n_.as<double>(),
                                                         Rh2_.as<double>(), Rh1_.as<double>(), Rglob_.as<double>(),
                                                         rword_.as<double>(), T, H, E);
      }
      launches_ += 4;
    }
  }
  // global feature -> average feature (explainers.py:634-639)
  uglob_kernel<<<W, 256, 0, s>>>(wr, Rglob_.as<double>(), gp_.as<double>(), U_.as<double>(), E);
  LRPCAP_TRY(gemm(U_.as<double>(), E, WgfT_, D, Y_.as<double>(), D, W, D, E, nullptr, s));
  ra_kernel<<<W, 256, 0, s>>>(wr, Y_.as<double>(), a_.as<double>(), ra_.as<double>(), D);
  launches_ += 2;
  // image_features dense over all grid cells, in word chunks (explainers.py:641-659)
  const int CH = std::min(W, 512);
  LRPCAP_TRY(UV_.ensure(std::max((size_t)CH, (size_t)N_) * L * H * 8));
  LRPCAP_TRY(YF_.ensure((size_t)CH * L * D * 8));
  // direct tail: the redistribution kernels write the GEMM's bf16 planes, final_kernel reads its fp32 result
  const int fh = (int)std::lround(std::sqrt((double)L));
  const bool tail_direct = direct && tc_features_ && fh * fh == L;
  const bool exact_uv = getenv("LRPCAP_DECODER_EXACT_UV") != nullptr;   // the float32-store emulation of explainers.py:1292-1299
  if (tail_direct) {
    LRPCAP_TRY(UVs_.ensure((size_t)CH * L * H * 4));
    LRPCAP_TRY(YF32_.ensure((size_t)CH * L * D * 4));
  }
  for (int p0 = 0; p0 < W; p0 += CH) {
    const int m = std::min(CH, W - p0);
    const size_t nUV = (size_t)m * L * H;
    __nv_bfloat16* UVs = tail_direct ? UVs_.as<__nv_bfloat16>() : nullptr;
    if (!td)
      uv_adaptive_kernel<<<dim3(L, m), 256, 0, s>>>(wr, p0, Vp_.as<double>(), alpha_.as<double>(), ctx_.as<double>(),
                                                    Rctx_.as<double>(), UV_.as<double>(), T, L, H, UVs, nUV);
    else if (T <= kUvMaxT) {
      constexpr int kLG = 14;
      if (tail_direct && !exact_uv)
        uv_gridtd_rows_f32_kernel<<<dim3((L + kLG - 1) / kLG, m), 256, kLG * kUvMaxT * sizeof(float), s>>>(
            wr, p0, Vp_.as<double>(), alpha_.as<double>(), Q_.as<double>(), T, L, H, kLG, UVs, nUV);
      else
        uv_gridtd_rows_kernel<<<dim3((L + kLG - 1) / kLG, m), 256, 0, s>>>(wr, p0, Vp_.as<double>(), alpha_.as<double>(),
                                                                           Q_.as<double>(), UV_.as<double>(), T, L, H, kLG, UVs, nUV);
    } else {
      if (tail_direct) {   // rare long captions: rows in fp64, then the ordinary conversion
        uv_gridtd_kernel<<<dim3(L, m), 256, 0, s>>>(wr, p0, Vp_.as<double>(), alpha_.as<double>(), Q_.as<double>(),
                                                    UV_.as<double>(), T, L, H);
        f64_to_split_kernel<<<nblk(nUV, 256), 256, 0, s>>>(UV_.as<double>(), UVs, UVs + nUV, nUV);
        ++launches_;
      } else {
        uv_gridtd_kernel<<<dim3(L, m), 256, 0, s>>>(wr, p0, Vp_.as<double>(), alpha_.as<double>(), Q_.as<double>(),
                                                    UV_.as<double>(), T, L, H);
      }
    }
    if (tail_direct) {
      TcConvArgs a;
      a.A = UVs_.p; a.A_elems = nUV; a.n_items = m; a.H = fh; a.W = fh; a.C = H;
      a.B = WifTC_; a.B_elems = (size_t)D * H; a.taps = 1; a.Nout = D;
      a.epi.mode = EPI_RAW;
      a.epi.out_f32 = YF32_.as<float>();
      LRPCAP_TRY(tc_conv_launch(a, s));
      ++launches_;
      if (F32_.p)
        final_kernel<float, float><<<dim3(L, m), 256, 0, s>>>(wr, p0, d_order_.as<int>(), F32_.as<float>(), ra_.as<double>(),
                                                              YF32_.as<float>(), d_R_head, L, D);
      else
        final_kernel<float><<<dim3(L, m), 256, 0, s>>>(wr, p0, d_order_.as<int>(), F_.as<double>(), ra_.as<double>(),
                                                       YF32_.as<float>(), d_R_head, L, D);
    } else {
      LRPCAP_TRY(features_gemm(m, s));
      final_kernel<double><<<dim3(L, m), 256, 0, s>>>(wr, p0, d_order_.as<int>(), F_.as<double>(), ra_.as<double>(),
                                                      YF_.as<double>(), d_R_head, L, D);
    }
    launches_ += 2;
  }
  LRPCAP_CUDA(cudaGetLastError());

  if (h_r_words || h_attention) {
    LRPCAP_CUDA(cudaStreamSynchronize(s));
    if (h_r_words) {
      std::vector<double> rw((size_t)W * T);
      LRPCAP_CUDA(cudaMemcpy(rw.data(), rword_.p, rw.size() * 8, cudaMemcpyDeviceToHost));
      for (int p = 0; p < W; ++p) {
        double* o = h_r_words + (size_t)order_[p] * T;
        const double* r = rw.data() + (size_t)p * T;
        const int t = wt_[p];
        for (int j = 0; j < T; ++j) o[j] = 0.0;
        if (td) {
          for (int j = 0; j < t; ++j) o[j] = r[j];                 // raw (explainers.py:1320)
        } else {                                                    // explainers.py:660-665
          double mx = 0.0;
          for (int j = 1; j < t; ++j) mx = std::max(mx, std::fabs(r[j]));
          for (int j = 1; j < t; ++j) o[j - 1] = (mx != 0.0) ? r[j] / mx : r[j];
        }
      }
    }
    if (h_attention) {
      std::vector<double> al((size_t)N_ * (T + 1) * L);
      LRPCAP_CUDA(cudaMemcpy(al.data(), alpha_.p, al.size() * 8, cudaMemcpyDeviceToHost));
      for (int w = 0; w < W; ++w)
        for (int l = 0; l < L; ++l)
          h_attention[(size_t)w * L + l] = (float)al[((size_t)h_word_img[w] * (T + 1) + h_word_t[w]) * L + l];
    }
  }
  return kOk;
}

int Decoder::backward(const int* h_word_img, const int* h_word_t, int W, float* d_R_head, double* h_r_words,
                      cudaStream_t s) {
  LRPCAP_REQUIRE(d_R_head != nullptr, kErrInvalidArg, "decoder_backward: null output");
  LRPCAP_TRY(sort_words(h_word_img, h_word_t, W, s));
  const int H = H_, E = E_, D = D_, L = L_, T = T_;
  const bool td = kind_ == LRPCAP_DECODER_GRIDTD;
  const size_t WH = (size_t)W * H;
  // buffer roles: Rh1_/Rh2_ = d_h1/d_h2, Rc1_/Rc2_ = d_c1/d_c2, U_ = gate gradients [W,4H], Rchat_ = d_c_hat seed,
  // Rctx_ = d_h1 increment from the language LSTM, Q_ = d_ctx per step, Rglob_ = d_global, rword_ = d_words
  DevBuf* rb[] = {&Rh1_, &Rc1_, &Rctx_, &Rh2_, &Rc2_, &Rchat_};
  for (DevBuf* b : rb) {
    LRPCAP_TRY(b->ensure(WH * 8));
    LRPCAP_CUDA(cudaMemsetAsync(b->p, 0, WH * 8, s));
  }
  LRPCAP_TRY(U_.ensure((size_t)W * std::max(4 * H, E) * 8));
  LRPCAP_TRY(Rglob_.ensure((size_t)W * E * 8));
  LRPCAP_TRY(rword_.ensure((size_t)W * T * 8));
  LRPCAP_TRY(Y_.ensure((size_t)W * std::max(Kin1_, std::max(Kin2_, D)) * 8));
  LRPCAP_TRY(ra_.ensure((size_t)W * D * 8));
  LRPCAP_CUDA(cudaMemsetAsync(Rglob_.p, 0, (size_t)W * E * 8, s));
  LRPCAP_CUDA(cudaMemsetAsync(rword_.p, 0, (size_t)W * T * 8, s));
  if (td) LRPCAP_TRY(Q_.ensure((size_t)W * T * H * 8));
  WordRef wr{d_wimg_.as<int>(), d_wt_.as<int>()};
  const int* tok = tok_.as<int>();

  grad_init_kernel<<<W, 256, 0, s>>>(wr, WoT_, tok, td ? Rh2_.as<double>() : Rh1_.as<double>(),
                                     td ? Rchat_.as<double>() : nullptr, T, H);
  ++launches_;
  for (int i = T - 1; i >= 0; --i) {
    const int na = nact_[i];
    if (na == 0) continue;
    if (!td) {
      grad_cell_kernel<<<na, 256, 0, s>>>(wr, i, ia1_.as<double>(), fa1_.as<double>(), ga1_.as<double>(), oa1_.as<double>(),
                                          c1_.as<double>(), Rh1_.as<double>(), nullptr, Rc1_.as<double>(), U_.as<double>(), T, H);
      if (WcatB1TC_) LRPCAP_TRY(gemm_tc(U_.as<double>(), na, 4 * H, WcatB1TC_, Kin1_, Y_.as<double>(), s));
      else LRPCAP_TRY(gemm(U_.as<double>(), 4 * H, Wcat1T_, Kin1_, Y_.as<double>(), Kin1_, na, Kin1_, 4 * H, nullptr, s));
      grad_scatter_adaptive_kernel<<<na, 256, 0, s>>>(i, Y_.as<double>(), Rh1_.as<double>(), Rglob_.as<double>(),
                                                      rword_.as<double>(), T, H, E);
      launches_ += 2;
    } else {
      grad_cell_kernel<<<na, 256, 0, s>>>(wr, i, ia2_.as<double>(), fa2_.as<double>(), ga2_.as<double>(), oa2_.as<double>(),
                                          c2_.as<double>(), Rh2_.as<double>(), nullptr, Rc2_.as<double>(), U_.as<double>(), T, H);
      if (WcatB2TC_) LRPCAP_TRY(gemm_tc(U_.as<double>(), na, 4 * H, WcatB2TC_, Kin2_, Y_.as<double>(), s));
      else LRPCAP_TRY(gemm(U_.as<double>(), 4 * H, Wcat2T_, Kin2_, Y_.as<double>(), Kin2_, na, Kin2_, 4 * H, nullptr, s));
      grad_scatter_lang_kernel<<<na, 256, 0, s>>>(wr, i, Y_.as<double>(), beta_.as<double>(), Rchat_.as<double>(),
                                                  Rctx_.as<double>(), Rh2_.as<double>(), Q_.as<double>(), T, H);
      grad_cell_kernel<<<na, 256, 0, s>>>(wr, i, ia1_.as<double>(), fa1_.as<double>(), ga1_.as<double>(), oa1_.as<double>(),
                                          c1_.as<double>(), Rh1_.as<double>(), Rctx_.as<double>(), Rc1_.as<double>(),
                                          U_.as<double>(), T, H);
      if (WcatB1TC_) LRPCAP_TRY(gemm_tc(U_.as<double>(), na, 4 * H, WcatB1TC_, Kin1_, Y_.as<double>(), s));
      else LRPCAP_TRY(gemm(U_.as<double>(), 4 * H, Wcat1T_, Kin1_, Y_.as<double>(), Kin1_, na, Kin1_, 4 * H, nullptr, s));
      grad_scatter_td_kernel<<<na, 256, 0, s>>>(i, Y_.as<double>(), Rh2_.as<double>(), Rh1_.as<double>(), Rglob_.as<double>(),
                                                rword_.as<double>(), T, H, E);
      launches_ += 4;
    }
  }
  grad_glob_mask_kernel<<<W, 256, 0, s>>>(wr, Rglob_.as<double>(), gp_.as<double>(), E, td ? 0 : 1);
  LRPCAP_TRY(gemm(Rglob_.as<double>(), E, WgfT_, D, ra_.as<double>(), D, W, D, E, nullptr, s));
  ++launches_;
  const int CH = std::min(W, 512);
  LRPCAP_TRY(UV_.ensure(std::max((size_t)CH, (size_t)N_) * L * H * 8));
  LRPCAP_TRY(YF_.ensure((size_t)CH * L * D * 8));
  for (int p0 = 0; p0 < W; p0 += CH) {
    const int m = std::min(CH, W - p0);
    if (!td)
      gv_adaptive_kernel<<<dim3(L, m), 256, 0, s>>>(wr, p0, Vp_.as<double>(), alpha_.as<double>(), WoT_, tok,
                                                    UV_.as<double>(), T, L, H);
    else
      gv_gridtd_kernel<<<dim3(L, m), 256, 0, s>>>(wr, p0, Vp_.as<double>(), alpha_.as<double>(), Q_.as<double>(),
                                                  UV_.as<double>(), T, L, H);
    LRPCAP_TRY(features_gemm(m, s));
    grad_final_kernel<<<dim3(L, m), 256, 0, s>>>(p0, d_order_.as<int>(), ra_.as<double>(), YF_.as<double>(), d_R_head, L, D);
    launches_ += 2;
  }
  LRPCAP_CUDA(cudaGetLastError());
  if (h_r_words) {
    LRPCAP_CUDA(cudaStreamSynchronize(s));
    std::vector<double> rw((size_t)W * T);
    LRPCAP_CUDA(cudaMemcpy(rw.data(), rword_.p, rw.size() * 8, cudaMemcpyDeviceToHost));
    for (int p = 0; p < W; ++p) {
      double* o = h_r_words + (size_t)order_[p] * T;
      for (int j = 0; j < T; ++j) o[j] = (j < wt_[p]) ? rw[(size_t)p * T + j] : 0.0;   // raw sums (explainers.py:831, 1527)
    }
  }
  return kOk;
}

}  // namespace lrpcap
