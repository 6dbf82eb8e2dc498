// Encoder: per-image forward state and batched per-word relevance backward (see encoder.cuh).
#include "encoder.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "encoder_kernels.cuh"
#include "tc_conv.cuh"

namespace lrpcap {

int DevBuf::ensure(size_t n) {
  if (n <= bytes) return kOk;
  release();
  LRPCAP_CUDA(cudaMalloc(&p, n));
  bytes = n;
  return kOk;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}

namespace {
struct Arch {
  int layers;
  int cin[16], cout[16], shift[16];
  bool pool[16];
};
const Arch kArch[2] = {
    {13, {3, 64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512}, {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512},
     {0, 0, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4}, {false, true, false, true, false, false, true, false, false, true, false, false, false}},
    {16, {3, 64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512, 512, 512},
     {64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512, 512, 512, 512}, {0, 0, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4},
     {false, true, false, true, false, false, false, true, false, false, false, true, false, false, false, false}},
};
constexpr int kForwardChunk = 64;   // images per forward pass: fills the 148 SMs on the 14x14 layers (1.2 GB per activation buffer)
}  // namespace

Encoder::~Encoder() {
  for (auto& l : L_) {
    if (l.w_hwio) cudaFree(l.w_hwio);
    if (l.bias) cudaFree(l.bias);
    for (auto& f : l.prepared)
      for (auto& p : f)
        if (p) cudaFree(p);
  }
  if (d_overflow_) cudaFree(d_overflow_);
  if (copy_stream_) cudaStreamDestroy(copy_stream_);
  if (ev_chunk_) cudaEventDestroy(ev_chunk_);
  if (ev_copied_) cudaEventDestroy(ev_copied_);
  if (w0_pm_) cudaFree(w0_pm_);
  if (w0_mp_) cudaFree(w0_mp_);
  if (w0_last_a_) cudaFree(w0_last_a_);
  if (w0_last_b_) cudaFree(w0_last_b_);
  Mseed2_.release();
  for (auto& g : G2_) g.release();
  for (auto& l : L_)
    for (auto& p : l.dual)
      if (p) cudaFree(p);
  X0_.release(); F_.release(); Mseed_.release(); posneg_.release(); idx_.release(); scale_.release();
  for (auto& g : G_) g.release();
  for (auto& g : Gc_) g.release();
  for (auto& g : Gc2_) g.release();
  for (auto& g : Gi_) g.release();
  for (auto& a : act_) a.release();
  for (auto& m : msg_) m.release();
}

int Encoder::create(Encoder** out, const float* const* kernels_hwio, const float* const* biases, int image_hw,
                    int precision, int arch) {
  LRPCAP_REQUIRE(arch == 0 || arch == 1, kErrInvalidArg, "encoder_create: unknown architecture %d (0 = VGG16, 1 = VGG19)", arch);
  const Arch& A = kArch[arch];
  LRPCAP_REQUIRE(out && kernels_hwio && biases, kErrInvalidArg, "encoder_create: null argument");
  LRPCAP_REQUIRE(image_hw >= 16 && image_hw % 16 == 0, kErrShape, "encoder_create: image size %d must be a multiple of 16", image_hw);
  LRPCAP_REQUIRE(precision >= PREC_FP32_SIMT && precision <= PREC_H1F8_TC, kErrInvalidArg,
                 "encoder_create: unknown precision %d", precision);
  Encoder* e = new Encoder();
  e->hw_ = image_hw;
  e->precision_ = precision;
  e->nl_ = A.layers;
  const int kLayers = A.layers;
  if (const char* v = std::getenv("LRPCAP_FWD_PLANES")) {    // knob: forward operand storage: 3 bf16 planes (default), 2 bf16
    const int n = std::atoi(v);                               // planes, or 4 = two IEEE half planes (kPlanesF16x2)
    if (n == 2 || n == 3 || n == kPlanesF16x2) e->fwd_planes_ = n;
  }
  if (const char* v = std::getenv("LRPCAP_MSG_TARGET_EXP")) {   // experiment knob: two-product message scale target
    const int n = std::atoi(v);
    if (n >= -8 && n <= 15) e->msg_target_exp_ = n;
  }
  if (const char* v = std::getenv("LRPCAP_FWD_PROMOTE")) {   // experiment knob: k-steps per accumulator hand-over, forward
    const int n = std::atoi(v);
    if (n >= 1 && n <= 64) e->fwd_promote_ = n;
  }
  for (int l = 0; l < kLayers; ++l) {
    Layer& L = e->L_[l];
    L.cin = A.cin[l];
    L.cout = A.cout[l];
    L.hw = image_hw >> A.shift[l];
    L.pool_after = A.pool[l];
    const size_t nw = (size_t)9 * L.cin * L.cout;
    if (!kernels_hwio[l] || !biases[l]) {
      delete e;
      set_last_error("encoder_create: layer %d weights missing", l);
      return kErrInvalidArg;
    }
    cudaError_t err = cudaMalloc(&L.w_hwio, nw * sizeof(float));
    if (err == cudaSuccess) err = cudaMalloc(&L.bias, L.cout * sizeof(float));
    if (err == cudaSuccess) err = cudaMemcpy(L.w_hwio, kernels_hwio[l], nw * sizeof(float), cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(L.bias, biases[l], L.cout * sizeof(float), cudaMemcpyHostToDevice);
    if (err != cudaSuccess) {
      delete e;
      set_last_error("encoder_create: device allocation/copy failed: %s", cudaGetErrorString(err));
      return kErrCuda;
    }
  }
  for (int l = 0; l < kLayers; ++l) e->set_wpow(l, kernels_hwio[l]);
  e->w0_host_.assign(kernels_hwio[0], kernels_hwio[0] + 9 * 3 * 64);
  *out = e;
  return kOk;
}

// 2^wpow * max|w| in [2^12, 2^13): the high half plane stays far from overflow, the low one (2^-11 of it) normal
void Encoder::set_wpow(int l, const float* h_w) {
  Layer& L = L_[l];
  float m = 0.f;
  const size_t n = (size_t)9 * L.cin * L.cout;
  for (size_t i = 0; i < n; ++i) m = std::fmax(m, std::fabs(h_w[i]));
  int ex = 0;
  if (m > 0.f && std::isfinite(m)) std::frexp(m, &ex);   // m = f * 2^ex, f in [0.5, 1)
  L.wpow = 13 - ex;
}

int Encoder::set_weights(const float* const* kernels_hwio, const float* const* biases) {
  const int kLayers = nl_;
  LRPCAP_REQUIRE(kernels_hwio && biases, kErrInvalidArg, "encoder_set_weights: null argument");
  for (int l = 0; l < kLayers; ++l)
    LRPCAP_REQUIRE(kernels_hwio[l] && biases[l], kErrInvalidArg, "encoder_set_weights: layer %d weights missing", l);
  LRPCAP_CUDA(cudaDeviceSynchronize());   // nothing may still read the old layouts
  for (int l = 0; l < kLayers; ++l) {
    Layer& L = L_[l];
    LRPCAP_CUDA(cudaMemcpy(L.w_hwio, kernels_hwio[l], (size_t)9 * L.cin * L.cout * sizeof(float), cudaMemcpyHostToDevice));
    LRPCAP_CUDA(cudaMemcpy(L.bias, biases[l], L.cout * sizeof(float), cudaMemcpyHostToDevice));
    for (int f = 0; f < 8; ++f)
      for (int sg = 0; sg < 3; ++sg) L.stale[f][sg] = L.prepared[f][sg] != nullptr;
    for (int d = 0; d < 4; ++d) L.dual_stale[d] = L.dual[d] != nullptr;
  }
  float** small[] = {&w0_pm_, &w0_mp_, &w0_last_a_, &w0_last_b_};
  for (float** p : small)
    if (*p) { cudaFree(*p); *p = nullptr; }
  for (int l = 0; l < kLayers; ++l) set_wpow(l, kernels_hwio[l]);
  w0_host_.assign(kernels_hwio[0], kernels_hwio[0] + 9 * 3 * 64);
  n_images_ = 0;
  return kOk;
}

namespace {
__global__ void absmax_kernel(const float* __restrict__ w, size_t n, unsigned* __restrict__ out) {
  unsigned best = 0u;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    best = max(best, __float_as_uint(w[i]) & 0x7fffffffu);
  best = __reduce_max_sync(0xffffffffu, best);
  if ((threadIdx.x & 31) == 0 && best) atomicMax(out, best);
}
}  // namespace

// As set_weights, from DEVICE tensors (fine-tuning keeps the trained parameters on the device: no host round trip).
int Encoder::set_weights_device(const float* const* d_kernels_hwio, const float* const* d_biases) {
  LRPCAP_REQUIRE(d_kernels_hwio && d_biases, kErrInvalidArg, "encoder_set_weights_device: null argument");
  for (int l = 0; l < nl_; ++l)
    LRPCAP_REQUIRE(d_kernels_hwio[l] && d_biases[l], kErrInvalidArg, "encoder_set_weights_device: layer %d weights missing", l);
  LRPCAP_CUDA(cudaDeviceSynchronize());   // nothing may still read the old layouts
  DevBuf mx;
  LRPCAP_TRY(mx.ensure(kMaxLayers * sizeof(unsigned)));
  LRPCAP_CUDA(cudaMemset(mx.p, 0, kMaxLayers * sizeof(unsigned)));
  for (int l = 0; l < nl_; ++l) {
    Layer& L = L_[l];
    const size_t nw = (size_t)9 * L.cin * L.cout;
    LRPCAP_CUDA(cudaMemcpy(L.w_hwio, d_kernels_hwio[l], nw * sizeof(float), cudaMemcpyDeviceToDevice));
    LRPCAP_CUDA(cudaMemcpy(L.bias, d_biases[l], L.cout * sizeof(float), cudaMemcpyDeviceToDevice));
    absmax_kernel<<<64, 256>>>(L.w_hwio, nw, mx.as<unsigned>() + l);
    // the derived layouts keep their allocations (cudaFree + cudaMalloc of ~40 buffers cost the fine-tune step 0.1-0.25 s
    // of host time) and are re-laid from the new w_hwio on their next use
    for (int f = 0; f < 8; ++f)
      for (int sg = 0; sg < 3; ++sg) L.stale[f][sg] = L.prepared[f][sg] != nullptr;
    for (int d = 0; d < 4; ++d) L.dual_stale[d] = L.dual[d] != nullptr;
  }
  LRPCAP_CUDA(cudaGetLastError());
  float** small[] = {&w0_pm_, &w0_mp_, &w0_last_a_, &w0_last_b_};
  for (float** p : small)
    if (*p) { cudaFree(*p); *p = nullptr; }
  float hmax[kMaxLayers];
  LRPCAP_CUDA(cudaMemcpy(hmax, mx.p, kMaxLayers * sizeof(float), cudaMemcpyDeviceToHost));
  mx.release();
  for (int l = 0; l < nl_; ++l) {
    int ex = 0;
    if (hmax[l] > 0.f && std::isfinite(hmax[l])) std::frexp(hmax[l], &ex);
    L_[l].wpow = 13 - ex;
  }
  w0_host_.resize((size_t)9 * 3 * 64);
  LRPCAP_CUDA(cudaMemcpy(w0_host_.data(), L_[0].w_hwio, w0_host_.size() * sizeof(float), cudaMemcpyDeviceToHost));
  n_images_ = 0;
  return kOk;
}

bool Encoder::getenv_off(const char* name) {
  const char* v = std::getenv(name);
  return v && v[0] == '0';
}

int Encoder::get_weights(int l, int fmt, int sign, void** out, cudaStream_t s) {
  Layer& L = L_[l];
  if (!L.prepared[fmt][sign] || L.stale[fmt][sign]) {
    void* p = L.prepared[fmt][sign];
    if (!p) LRPCAP_CUDA(cudaMalloc(&p, (size_t)9 * L.cin * L.cout * (fmt == WF_TC_FWD3 ? 6 : 4)));
    L.prepared[fmt][sign] = p;
    L.stale[fmt][sign] = false;
    if (fmt == WF_TC_FWD3) LRPCAP_TRY(prep_weights(L.w_hwio, p, L.cin, L.cout, WF_TC_FWD, sign, s, 9, 3));
    else if (fmt == WF_TC_BWDH)
      LRPCAP_TRY(prep_weights(L.w_hwio, p, L.cin, L.cout, WF_TC_BWD, sign, s, 9, kPlanesF16x2, std::ldexp(1.f, L.wpow)));
    else if (fmt == WF_TC_BWDF8)
      LRPCAP_TRY(prep_weights(L.w_hwio, p, L.cin, L.cout, WF_TC_BWD, sign, s, 9, kPlanesH1F8, std::ldexp(1.f, L.wpow)));
    else if (fmt == WF_TC_FWDH)
      LRPCAP_TRY(prep_weights(L.w_hwio, p, L.cin, L.cout, WF_TC_FWD, sign, s, 9, kPlanesF16x2, std::ldexp(1.f, L.wpow)));
    else LRPCAP_TRY(prep_weights(L.w_hwio, p, L.cin, L.cout, fmt, sign, s));
    ++launches_;
  }
  *out = L.prepared[fmt][sign];
  return kOk;
}

int Encoder::get_dual_weights(int l, bool tc, void** out, cudaStream_t s) {
  Layer& L = L_[l];
  const bool f8 = tc && fp8_mode();
  const bool h = tc && (two_product() || f8);
  const int di = tc ? (f8 ? 3 : h ? 2 : 1) : 0;
  void*& slot = L.dual[di];
  if (!slot || L.dual_stale[di]) {
    if (!slot) LRPCAP_CUDA(cudaMalloc(&slot, (size_t)9 * L.cin * 2 * L.cout * sizeof(float)));
    L.dual_stale[di] = false;
    const float ws = h ? std::ldexp(1.f, L.wpow) : 1.f;
    LRPCAP_TRY(prep_weights_dual(L.w_hwio, slot, L.cin, L.cout, tc ? WF_TC_BWD : WF_SIMT_BWD, WS_PLUS, rule_.alpha * ws, WS_MINUS,
                                 -rule_.beta * ws, s, f8 ? 2 : h ? 1 : 0));
    ++launches_;
  }
  *out = slot;
  return kOk;
}

int Encoder::conv(int l, bool backward, int sign, const void* A, size_t A_elems, int n_items, const EpiParams& epi,
                  cudaStream_t s, bool dual) {
  const Layer& L = L_[l];
  const int C = backward ? (dual ? 2 * L.cout : L.cout) : L.cin;
  const int Nout = backward ? L.cin : L.cout;
  void* B = nullptr;
  ++launches_;
  const bool tc = split() && C % 64 == 0 && Nout % 64 == 0;
  if (!tc) LRPCAP_REQUIRE(!split() || l == 0, kErrState, "encoder: layer %d has no tensor-core shape", l);
  if (dual) LRPCAP_TRY(get_dual_weights(l, tc, &B, s));
  else LRPCAP_TRY(get_weights(l, tc ? (backward ? (fp8_mode() ? WF_TC_BWDF8 : two_product() ? WF_TC_BWDH : WF_TC_BWD) : (fwd_planes_ == 3 ? WF_TC_FWD3 : fwd_planes_ == kPlanesF16x2 ? WF_TC_FWDH : WF_TC_FWD)) : (backward ? WF_SIMT_BWD : WF_SIMT_FWD), sign, &B, s));
  ProfRec rec{};
  if (profile_) {
    LRPCAP_CUDA(cudaEventCreate(&rec.a));
    LRPCAP_CUDA(cudaEventCreate(&rec.b));
    rec.cls = tc ? (backward ? 0 : 1) : 2;
    rec.flops = 2.0 * 9.0 * (double)n_items * L.hw * L.hw * (double)L.cin * L.cout * (dual ? 2.0 : 1.0);
    LRPCAP_CUDA(cudaEventRecord(rec.a, s));
  }
  int st;
  if (tc) {
    TcConvArgs a;
    a.A = A; a.A_elems = A_elems; a.n_items = n_items; a.H = L.hw; a.W = L.hw; a.C = C;
    a.B = B; a.B_elems = (size_t)9 * L.cin * L.cout * (dual ? 2 : 1); a.taps = 9; a.Nout = Nout;
    a.planes = backward ? (fp8_mode() ? kPlanesH1F8 : two_product() ? kPlanesH1x2 : 2) : fwd_planes_;
    // backward: tensor-core fp32 accumulation rounds toward zero, which shrinks every output of a chain of n accumulates
    // by ~1.5e-8 n (measured, tools/diag_parity.py trunc: -1.6e-5 at K = 4608, -1e-4 over the 12 layers). A uniform 1e-4
    // scale error is inside the per-pixel tolerance of every rule, but for the same-sign chains of the alpha-beta family
    // it IS the conservation-sum error (mixed-sign rules: 6e-7). Default (-1): for those rules the deep layers
    // (K >= 2304) hand their accumulator to fp32 registers once per filter tap (sum error 1.5e-5, +5 % time); the
    // shallow layers (< 3e-6 each) and the other rules keep the faster path.
    int pe = bwd_promote_;
    if (pe < 0) {
      const bool same_sign = rule_.kind == RULE_ALPHA_BETA || rule_.kind == RULE_ZPLUS_FAST;
      pe = (same_sign && C >= 256) ? (C / 64 < 8 ? C / 64 : 8) : 0;
    }
    a.promote_every = backward ? pe : fwd_promote_;
    a.epi = epi;
    if (backward && scaled_messages()) a.epi.acc_scale = std::ldexp(1.f, -L.wpow);   // the fp16 weight planes hold 2^wpow * w
    if (!backward && fwd_planes_ == kPlanesF16x2) {
      a.epi.acc_scale = std::ldexp(1.f, -L.wpow);
      a.epi.overflow = d_overflow_;
    }
    st = tc_conv_launch(a, s);
  } else {
    SimtConvArgs a;
    a.A = reinterpret_cast<const float*>(A); a.n_items = n_items; a.H = L.hw; a.W = L.hw; a.C = C;
    a.B = reinterpret_cast<const float*>(B); a.taps = 9; a.Nout = Nout;
    LRPCAP_REQUIRE(!(backward && scaled_messages()), kErrState, "encoder: layer %d backward has no tensor-core shape", l);
    a.out_planes = split() ? (backward ? 2 : fwd_planes_) : 0;
    a.epi = epi;
    if (!backward && split() && fwd_planes_ == kPlanesF16x2) a.epi.overflow = d_overflow_;
    st = simt_conv_launch(a, s);
  }
  if (profile_) {
    LRPCAP_CUDA(cudaEventRecord(rec.b, s));
    prof_.push_back(rec);
  }
  return st;
}

int Encoder::profile_read(double* out) {
  LRPCAP_REQUIRE(out != nullptr, kErrInvalidArg, "profile_read: null output");
  for (int i = 0; i < 12; ++i) out[i] = 0.0;
  for (ProfRec& r : prof_) {
    LRPCAP_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    LRPCAP_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    out[3 * r.cls + 0] += ms;
    out[3 * r.cls + 1] += r.flops;
    out[3 * r.cls + 2] += 1.0;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  prof_.clear();
  return kOk;
}

int Encoder::debug_pool_routes(int l, unsigned char* h_out) {
  const int kLayers = nl_;
  LRPCAP_REQUIRE(n_images_ > 0, kErrState, "debug_pool_routes: call encoder_forward first");
  LRPCAP_REQUIRE(l >= 0 && l < kLayers - 1 && L_[l].pool_after && h_out, kErrInvalidArg,
                 "debug_pool_routes: layer %d is not followed by a max-pool", l);
  const Layer& L = L_[l];
  const int Ho = L.hw / 2, C = L.cout;
  const size_t words = (size_t)n_images_ * (C / 16) * Ho * Ho;
  std::vector<unsigned> gi(words);
  LRPCAP_CUDA(cudaDeviceSynchronize());
  LRPCAP_CUDA(cudaMemcpy(gi.data(), Gi_[l].p, words * sizeof(unsigned), cudaMemcpyDeviceToHost));
  for (int img = 0; img < n_images_; ++img)
    for (int j = 0; j < C / 16; ++j)
      for (int y = 0; y < Ho; ++y)
        for (int x = 0; x < Ho; ++x) {
          const unsigned w = gi[(((size_t)img * (C / 16) + j) * Ho + y) * Ho + x];
          unsigned char* o = h_out + (((size_t)img * Ho + y) * Ho + x) * C + 16 * j;
          for (int k = 0; k < 16; ++k) o[k] = (unsigned char)((w >> (2 * k)) & 3u);
        }
  return kOk;
}

int Encoder::debug_message_scales(float* h_max, int* h_kt, int cap_words, int* chunk) {
  LRPCAP_REQUIRE(h_max && h_kt && chunk, kErrInvalidArg, "debug_message_scales: null argument");
  *chunk = 0;
  if (!scale_.p || scale_m_ == 0) return kOk;
  LRPCAP_REQUIRE(cap_words >= scale_m_, kErrInvalidArg, "debug_message_scales: room for %d words needed", scale_m_);
  LRPCAP_CUDA(cudaDeviceSynchronize());
  std::vector<unsigned> mx((size_t)(nl_ + 1) * scale_cw_);
  std::vector<int> kt((size_t)nl_ * scale_cw_);
  LRPCAP_CUDA(cudaMemcpy(mx.data(), scale_.p, mx.size() * 4, cudaMemcpyDeviceToHost));
  LRPCAP_CUDA(cudaMemcpy(kt.data(), scale_.as<unsigned>() + mx.size(), kt.size() * 4, cudaMemcpyDeviceToHost));
  for (int l = 0; l <= nl_; ++l)
    for (int w = 0; w < scale_m_; ++w) {
      float f;
      std::memcpy(&f, &mx[(size_t)l * scale_cw_ + w], 4);
      h_max[(size_t)l * scale_m_ + w] = f;
      if (l < nl_) h_kt[(size_t)l * scale_m_ + w] = kt[(size_t)l * scale_cw_ + w];
    }
  *chunk = scale_m_;
  return kOk;
}

int Encoder::debug_multiplier(int l, int branch, float* h_out) {
  const int kLayers = nl_;
  LRPCAP_REQUIRE(n_images_ > 0, kErrState, "debug_multiplier: call encoder_forward first");
  LRPCAP_REQUIRE(l >= 0 && l < kLayers - 1 && h_out && (branch == 0 || branch == 1), kErrInvalidArg,
                 "debug_multiplier: layer %d / branch %d out of range", l, branch);
  const Layer& L = L_[l];
  const int H = L.hw, C = L.cout;
  const bool pooled = L.pool_after;
  const DevBuf& src = pooled ? (branch ? Gc2_[l] : Gc_[l]) : (branch ? G2_[l] : G_[l]);
  const int Hs = pooled ? H / 2 : H;
  const size_t n = (size_t)n_images_ * C * Hs * Hs;
  LRPCAP_REQUIRE(src.p && src.bytes >= n * sizeof(float), kErrState, "debug_multiplier: this rule keeps no such multiplier");
  std::vector<float> g(n);
  std::vector<unsigned> gi;
  LRPCAP_CUDA(cudaDeviceSynchronize());
  LRPCAP_CUDA(cudaMemcpy(g.data(), src.p, n * sizeof(float), cudaMemcpyDeviceToHost));
  if (pooled) {
    gi.resize(n / 16);
    LRPCAP_CUDA(cudaMemcpy(gi.data(), Gi_[l].p, gi.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
    std::fill(h_out, h_out + (size_t)n_images_ * H * H * C, 0.f);
  }
  for (int img = 0; img < n_images_; ++img)
    for (int j = 0; j < C / 16; ++j)
      for (int y = 0; y < Hs; ++y)
        for (int x = 0; x < Hs; ++x) {
          const size_t run = (((size_t)img * (C / 16) + j) * Hs + y) * Hs + x;
          for (int k = 0; k < 16; ++k) {
            int yy = y, xx = x;
            if (pooled) {
              const unsigned pos = (gi[run] >> (2 * k)) & 3u;
              yy = 2 * y + (int)(pos >> 1);
              xx = 2 * x + (int)(pos & 1u);
            }
            h_out[(((size_t)img * H + yy) * H + xx) * C + 16 * j + k] = g[run * 16 + k];
          }
        }
  return kOk;
}

int Encoder::forward(const float* d_images, int n, const EncoderRule& rule, cudaStream_t s) {
  const int kLayers = nl_;
  LRPCAP_REQUIRE(d_images && n > 0, kErrInvalidArg, "encoder_forward: no images");
  switch (rule.kind) {
    case RULE_EPSILON:
      LRPCAP_REQUIRE(rule.epsilon > 0.f, kErrInvalidArg, "encoder_forward: epsilon must be > 0");
      break;
    case RULE_ALPHA_BETA:
      LRPCAP_REQUIRE(rule.alpha >= 1.f && rule.beta >= 0.f && fabsf(rule.alpha - rule.beta - 1.f) < 1e-6f,
                     kErrInvalidArg, "encoder_forward: need alpha >= 1, beta >= 0, alpha - beta = 1");
      break;
    case RULE_Z: case RULE_ZPLUS_FAST: case RULE_GRADIENT: case RULE_INPUT_T_GRADIENT: case RULE_GUIDED_BACKPROP:
      break;
    default:
      set_last_error("encoder_forward: unknown rule %d", rule.kind);
      return kErrInvalidArg;
  }
  n_images_ = 0;
  rule_ = rule;
  if (split() && fwd_planes_ == kPlanesF16x2) {
    if (!d_overflow_) LRPCAP_CUDA(cudaMalloc(&d_overflow_, sizeof(int)));
    LRPCAP_CUDA(cudaMemsetAsync(d_overflow_, 0, sizeof(int), s));
  }
  const size_t img_elems = (size_t)hw_ * hw_ * 3;
  LRPCAP_TRY(X0_.ensure((size_t)n * img_elems * sizeof(float)));
  if (d_images != X0_.p)   // (the half-range fallback below re-enters with the kept copy)
    LRPCAP_CUDA(cudaMemcpyAsync(X0_.p, d_images, (size_t)n * img_elems * sizeof(float), cudaMemcpyDeviceToDevice, s));
  for (int l = 0; l < kLayers - 1; ++l) {
    LRPCAP_TRY(G_[l].ensure((size_t)n * layer_out_elems(l) * sizeof(float)));
    if (L_[l].pool_after) {
      LRPCAP_TRY(Gc_[l].ensure((size_t)n * layer_out_elems(l) / 4 * sizeof(float)));
      LRPCAP_TRY(Gi_[l].ensure((size_t)n * layer_out_elems(l) / 64 * sizeof(unsigned)));
    }
  }
  LRPCAP_TRY(Mseed_.ensure((size_t)n * layer_out_elems(kLayers - 1) * sizeof(float)));
  LRPCAP_TRY(F_.ensure((size_t)n * layer_out_elems(kLayers - 1) * sizeof(float)));
  const int FC = n < kForwardChunk ? n : kForwardChunk;
  const size_t act_bytes = (size_t)FC * hw_ * hw_ * 64 * (split() ? 6 : 4);
  for (auto& a : act_) LRPCAP_TRY(a.ensure(act_bytes));

  const bool ab = rule.kind == RULE_ALPHA_BETA, zpf = rule.kind == RULE_ZPLUS_FAST;
  const bool inh = ab && rule.beta != 0.f;   // inhibitor branch f(W-, W+, x+, x-) (relevance_rule.py:314-320)
  if (inh) {
    for (int l = 0; l < kLayers - 1; ++l) {
      LRPCAP_TRY(G2_[l].ensure((size_t)n * layer_out_elems(l) * sizeof(float)));
      if (L_[l].pool_after) LRPCAP_TRY(Gc2_[l].ensure((size_t)n * layer_out_elems(l) / 4 * sizeof(float)));
    }
    LRPCAP_TRY(Mseed2_.ensure((size_t)n * layer_out_elems(kLayers - 1) * sizeof(float)));
    if (dual_alpha_ != rule.alpha || dual_beta_ != rule.beta) {   // cached stacked weights depend on (alpha, beta)
      for (auto& L : L_)
        for (int d = 0; d < 4; ++d) L.dual_stale[d] = L.dual[d] != nullptr;
      if (w0_last_a_) { cudaFree(w0_last_a_); w0_last_a_ = nullptr; }
      if (w0_last_b_) { cudaFree(w0_last_b_); w0_last_b_ = nullptr; }
      dual_alpha_ = rule.alpha;
      dual_beta_ = rule.beta;
    }
  }
  int gmode = G_MASK;
  if (rule.kind == RULE_EPSILON) gmode = G_EPS;
  else if (rule.kind == RULE_Z) gmode = G_Z;
  else if (ab || zpf) gmode = G_NONE;

  if (ab && !w0_pm_) {   // [W+ ; W-] stacked along the input-channel axis for the [x+, x-] first-layer input
    std::vector<float> pm((size_t)9 * 6 * 64);
    for (int tap = 0; tap < 9; ++tap)
      for (int ci = 0; ci < 3; ++ci)
        for (int co = 0; co < 64; ++co) {
          const float w = w0_host_[((size_t)tap * 3 + ci) * 64 + co];
          pm[((size_t)tap * 6 + ci) * 64 + co] = w >= 0.f ? w : 0.f;
          pm[((size_t)tap * 6 + 3 + ci) * 64 + co] = w < 0.f ? w : 0.f;
        }
    LRPCAP_CUDA(cudaMalloc(&w0_pm_, pm.size() * sizeof(float)));
    LRPCAP_CUDA(cudaMemcpy(w0_pm_, pm.data(), pm.size() * sizeof(float), cudaMemcpyHostToDevice));
    for (int tap = 0; tap < 9; ++tap)      // [W- ; W+]: z_inh = W- * x+ + W+ * x-
      for (int ci = 0; ci < 3; ++ci)
        for (int co = 0; co < 64; ++co) {
          const float w = w0_host_[((size_t)tap * 3 + ci) * 64 + co];
          pm[((size_t)tap * 6 + ci) * 64 + co] = w < 0.f ? w : 0.f;
          pm[((size_t)tap * 6 + 3 + ci) * 64 + co] = w >= 0.f ? w : 0.f;
        }
    LRPCAP_CUDA(cudaMalloc(&w0_mp_, pm.size() * sizeof(float)));
    LRPCAP_CUDA(cudaMemcpy(w0_mp_, pm.data(), pm.size() * sizeof(float), cudaMemcpyHostToDevice));
  }

  for (int i0 = 0; i0 < n; i0 += FC) {
    const int m = (n - i0) < FC ? (n - i0) : FC;
    const float* img = X0_.as<float>() + (size_t)i0 * img_elems;
    const void* X = img;
    size_t X_elems = (size_t)m * img_elems;
    int bx = -1;  // act_ buffer currently holding X (-1: the image)
    for (int l = 0; l < kLayers; ++l) {
      const Layer& L = L_[l];
      const size_t oe = layer_out_elems(l);
      int by = 0;
      while (by == bx) ++by;
      void* Y = act_[by].p;
      float* Gl = (l < kLayers - 1) ? G_[l].as<float>() + (size_t)i0 * oe : nullptr;
      float* Ml = (l == kLayers - 1) ? Mseed_.as<float>() + (size_t)i0 * oe : nullptr;

      EpiParams ep;
      ep.mode = EPI_FWD_TRUE;
      ep.bias = L.bias;
      ep.out_act = Y;
      ep.out_act_elems = (size_t)m * oe;
      ep.out_f32 = (l == kLayers - 1) ? F_.as<float>() + (size_t)i0 * oe : nullptr;
      ep.gmode = gmode;
      ep.eps = rule.epsilon;
      ep.rule_bias = rule.bias;
      ep.g_up = L.pool_after ? 2 : 1;
      if (gmode != G_NONE) { ep.G = Gl; ep.Mseed = Ml; }
      ep.g_elems = (size_t)m * oe;
      ep.aux_elems = (size_t)m * oe;
      LRPCAP_TRY(conv(l, false, WS_ALL, X, X_elems, m, ep, s));

      if (ab || zpf) {
        EpiParams ez;
        ez.mode = EPI_FWD_ZACT;
        ez.bias = L.bias;
        ez.rule_bias = ab ? rule.bias : 0;
        ez.x_act = Y;
        ez.x_act_elems = (size_t)m * oe;
        ez.G = Gl;
        ez.Mseed = Ml;
        ez.g_elems = (size_t)m * oe;
        ez.aux_elems = (size_t)m * oe;
        ez.g_up = L.pool_after ? 2 : 1;
        if (l == 0 && ab) {
          LRPCAP_TRY(posneg_.ensure((size_t)m * hw_ * hw_ * 6 * sizeof(float)));
          LRPCAP_TRY(make_posneg(img, posneg_.as<float>(), (size_t)m * hw_ * hw_, s));
          SimtConvArgs a;
          a.A = posneg_.as<float>(); a.n_items = m; a.H = hw_; a.W = hw_; a.C = 6;
          a.B = w0_pm_; a.taps = 9; a.Nout = 64; a.out_planes = fwd_planes();
          a.epi = ez;
          LRPCAP_TRY(simt_conv_launch(a, s));
          launches_ += 2;
        } else {
          LRPCAP_TRY(conv(l, false, WS_PLUS, X, X_elems, m, ez, s));
        }
        if (inh) {   // second multiplier: x / safe(z_inh), z_inh = W- * x+ + W+ * x- + b
          ez.G = (l < kLayers - 1) ? G2_[l].as<float>() + (size_t)i0 * oe : nullptr;
          ez.Mseed = (l == kLayers - 1) ? Mseed2_.as<float>() + (size_t)i0 * oe : nullptr;
          if (l == 0) {
            SimtConvArgs a;
            a.A = posneg_.as<float>(); a.n_items = m; a.H = hw_; a.W = hw_; a.C = 6;
            a.B = w0_mp_; a.taps = 9; a.Nout = 64; a.out_planes = fwd_planes();
            a.epi = ez;
            LRPCAP_TRY(simt_conv_launch(a, s));
            ++launches_;
          } else {
            LRPCAP_TRY(conv(l, false, WS_MINUS, X, X_elems, m, ez, s));
          }
        }
      }

      if (L.pool_after) {
        int bp = 0;
        while (bp == bx || bp == by) ++bp;
        LRPCAP_TRY(pool_mask(Y, (size_t)m * oe, fwd_planes(), act_[bp].p, (size_t)m * oe / 4, Gl,
                             Gc_[l].as<float>() + (size_t)i0 * oe / 4, Gi_[l].as<unsigned>() + (size_t)i0 * oe / 64, m, L.hw,
                             L.hw, L.cout, s));
        ++launches_;
        if (inh) {
          LRPCAP_TRY(pool_mask(Y, (size_t)m * oe, fwd_planes(), nullptr, 0, G2_[l].as<float>() + (size_t)i0 * oe,
                               Gc2_[l].as<float>() + (size_t)i0 * oe / 4, nullptr, m, L.hw, L.hw, L.cout, s));
          ++launches_;
        }
        X = act_[bp].p;
        X_elems = (size_t)m * oe / 4;
        bx = bp;
      } else {
        X = Y;
        X_elems = (size_t)m * oe;
        bx = by;
      }
    }
  }
  if (split() && fwd_planes_ == kPlanesF16x2) {
    // an activation outside the half range would turn into inf: fall back to the bf16 planes for this handle, for good
    int flag = 0;
    LRPCAP_CUDA(cudaMemcpyAsync(&flag, d_overflow_, sizeof(int), cudaMemcpyDeviceToHost, s));
    LRPCAP_CUDA(cudaStreamSynchronize(s));
    if (flag) {
      fwd_planes_ = 3;
      return forward(X0_.as<float>(), n, rule, s);
    }
  }
  n_images_ = n;
  return kOk;
}

int Encoder::relevance(const int* h_img_index, const float* d_R_head, int n_words, float* d_R_pix, cudaStream_t s,
                       float* h_R_pix) {
  const int kLayers = nl_;
  LRPCAP_REQUIRE(n_images_ > 0, kErrState, "encoder_relevance: call encoder_forward first");
  LRPCAP_REQUIRE(h_img_index && d_R_head && d_R_pix && n_words > 0, kErrInvalidArg, "encoder_relevance: bad argument");
  for (int w = 0; w < n_words; ++w)
    LRPCAP_REQUIRE(h_img_index[w] >= 0 && h_img_index[w] < n_images_, kErrInvalidArg,
                   "encoder_relevance: img_index[%d]=%d out of range [0,%d)", w, h_img_index[w], n_images_);
  LRPCAP_TRY(idx_.ensure((size_t)n_words * sizeof(int)));
  LRPCAP_CUDA(cudaMemcpyAsync(idx_.p, h_img_index, (size_t)n_words * sizeof(int), cudaMemcpyHostToDevice, s));
  const int CW = n_words < chunk_words_ ? n_words : chunk_words_;
  const bool inh = rule_.kind == RULE_ALPHA_BETA && rule_.beta != 0.f;
  const size_t msg_bytes = (size_t)CW * hw_ * hw_ * 64 * sizeof(float) * (inh ? 2 : 1);
  for (auto& m : msg_) LRPCAP_TRY(m.ensure(msg_bytes));

  const bool ab = rule_.kind == RULE_ALPHA_BETA, zpf = rule_.kind == RULE_ZPLUS_FAST;
  const bool guided = rule_.kind == RULE_GUIDED_BACKPROP;
  const int sign = (ab || zpf) ? WS_PLUS : WS_ALL;
  const int mult = (rule_.kind == RULE_GRADIENT || guided) ? 0 : 1;
  const int fh = hw_ / 16;
  const size_t head_elems = layer_out_elems(kLayers - 1);
  const size_t pix_elems = (size_t)hw_ * hw_ * 3;

  void *Wa = nullptr, *Wb = nullptr;
  if (inh) {   // x >= 0: alpha W+^T s_a - beta W-^T s_i ;  x < 0: alpha W-^T s_a - beta W+^T s_i
    if (!w0_last_a_) {
      LRPCAP_CUDA(cudaMalloc(&w0_last_a_, (size_t)9 * 128 * 3 * sizeof(float)));
      LRPCAP_CUDA(cudaMalloc(&w0_last_b_, (size_t)9 * 128 * 3 * sizeof(float)));
      LRPCAP_TRY(prep_weights_dual(L_[0].w_hwio, w0_last_a_, 3, 64, WF_SIMT_BWD, WS_PLUS, rule_.alpha, WS_MINUS, -rule_.beta, s));
      LRPCAP_TRY(prep_weights_dual(L_[0].w_hwio, w0_last_b_, 3, 64, WF_SIMT_BWD, WS_MINUS, rule_.alpha, WS_PLUS, -rule_.beta, s));
      launches_ += 2;
    }
    Wa = w0_last_a_;
    Wb = w0_last_b_;
  } else {
    LRPCAP_TRY(get_weights(0, WF_SIMT_BWD, sign, &Wa, s));
    if (ab) LRPCAP_TRY(get_weights(0, WF_SIMT_BWD, WS_MINUS, &Wb, s));
  }
  const int mul = inh ? 2 : 1;
  const bool tp = scaled_messages();
  if (tp) LRPCAP_TRY(scale_.ensure((size_t)(2 * kLayers + 1) * CW * sizeof(unsigned)));
  unsigned* mxb = scale_.as<unsigned>();   // [layers + 1][CW]: row l = stored max of message l, last row = the seed's true max
  int* ktb = reinterpret_cast<int*>(mxb + (size_t)(kLayers + 1) * CW);   // [layers][CW]: log2 scale of message l

  for (int w0 = 0, m = 0; w0 < n_words; w0 += m) {
    m = (n_words - w0) < CW ? (n_words - w0) : CW;
    // streaming to the host: the copy of the last chunk is not overlapped with anything, so make that chunk small
    if (h_R_pix && w0 + m == n_words && m >= 128) m = (m / 2 + 31) / 32 * 32;
    const int* idx = idx_.as<int>() + w0;
    int cur = 0;
    if (tp) {
      scale_cw_ = CW;
      scale_m_ = m;
      LRPCAP_CUDA(cudaMemsetAsync(mxb, 0, (size_t)(kLayers + 1) * CW * sizeof(unsigned), s));
      LRPCAP_TRY(seed_message_scaled(d_R_head + (size_t)w0 * head_elems, Mseed_.as<float>(), inh ? Mseed2_.as<float>() : nullptr,
                                     idx, msg_[cur].p, m, fh * fh, 512, guided ? 1 : 0, mxb + (size_t)kLayers * CW,
                                     mxb + (size_t)(kLayers - 1) * CW, ktb + (size_t)(kLayers - 1) * CW, msg_target_exp_, s,
                                     fp8_mode() ? 1 : 0));
      ++launches_;
    } else {
      LRPCAP_TRY(seed_message(d_R_head + (size_t)w0 * head_elems, Mseed_.as<float>(), inh ? Mseed2_.as<float>() : nullptr, idx,
                              msg_[cur].p, (size_t)m * head_elems * mul, split(), m, fh * fh, 512, guided ? 1 : 0, s));
    }
    ++launches_;
    for (int l = kLayers - 1; l >= 1; --l) {
      EpiParams ep;
      ep.mode = EPI_BWD;
      ep.img_index = idx;
      ep.up = L_[l - 1].pool_after ? 2 : 1;
      ep.Gin = ep.up == 2 ? Gc_[l - 1].as<float>() : G_[l - 1].as<float>();
      ep.Gidx = ep.up == 2 ? Gi_[l - 1].as<unsigned>() : nullptr;
      ep.g_elems = (size_t)n_images_ * layer_out_elems(l - 1) / (ep.up == 2 ? 4 : 1);
      ep.relu_acc = guided ? 1 : 0;
      ep.out_msg = msg_[cur ^ 1].p;
      ep.Gin2 = inh ? (ep.up == 2 ? Gc2_[l - 1].as<float>() : G2_[l - 1].as<float>()) : nullptr;
      ep.out_msg_elems = (size_t)m * layer_out_elems(l - 1) * mul;
      ep.out_planar_f32 = (l == 1) ? 1 : 0;   // the last message is read by last_dgrad only: fp32, channel-planar
      if (tp) {
        ep.mx_in = mxb + (size_t)l * CW;
        ep.kt_in = ktb + (size_t)l * CW;
        ep.mx_out = l > 1 ? mxb + (size_t)(l - 1) * CW : nullptr;
        ep.kt_out = l > 1 ? ktb + (size_t)(l - 1) * CW : nullptr;
        ep.target_exp = msg_target_exp_;
      }
      LRPCAP_TRY(conv(l, true, sign, msg_[cur].p, (size_t)m * layer_out_elems(l) * mul, m, ep, s, inh));
      cur ^= 1;
    }
    ProfRec rec{};
    if (profile_) {
      LRPCAP_CUDA(cudaEventCreate(&rec.a));
      LRPCAP_CUDA(cudaEventCreate(&rec.b));
      rec.cls = 3;
      rec.flops = 2.0 * 9.0 * (double)m * hw_ * hw_ * 64.0 * mul * 3.0;
      LRPCAP_CUDA(cudaEventRecord(rec.a, s));
    }
    LRPCAP_TRY(last_dgrad(msg_[cur].as<float>(), reinterpret_cast<const float*>(Wa), reinterpret_cast<const float*>(Wb),
                          X0_.as<float>(), idx, d_R_pix + (size_t)w0 * pix_elems, m, hw_, hw_, 64 * mul, mult, s));
    if (profile_) {
      LRPCAP_CUDA(cudaEventRecord(rec.b, s));
      prof_.push_back(rec);
    }
    ++launches_;
    if (h_R_pix) {   // stream this chunk's maps to the host while the next chunk computes
      if (!copy_stream_) {
        LRPCAP_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
        LRPCAP_CUDA(cudaEventCreateWithFlags(&ev_chunk_, cudaEventDisableTiming));
        LRPCAP_CUDA(cudaEventCreateWithFlags(&ev_copied_, cudaEventDisableTiming));
      }
      LRPCAP_CUDA(cudaEventRecord(ev_chunk_, s));
      LRPCAP_CUDA(cudaStreamWaitEvent(copy_stream_, ev_chunk_, 0));
      LRPCAP_CUDA(cudaMemcpyAsync(h_R_pix + (size_t)w0 * pix_elems, d_R_pix + (size_t)w0 * pix_elems,
                                  (size_t)m * pix_elems * sizeof(float), cudaMemcpyDeviceToHost, copy_stream_));
    }
  }
  if (h_R_pix && copy_stream_) {
    LRPCAP_CUDA(cudaEventRecord(ev_copied_, copy_stream_));
    LRPCAP_CUDA(cudaStreamWaitEvent(s, ev_copied_, 0));
  }
  return kOk;
}

}  // namespace lrpcap
