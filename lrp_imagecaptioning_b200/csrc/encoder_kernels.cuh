// HBM-bound helper kernels of the encoder relevance path (first/last layer).
#pragma once
#include "common.cuh"

namespace lrpcap {

enum WeightFormat : int {
  WF_SIMT_FWD = 0,  // fp32 [tap][cin][cout]                 (= HWIO)
  WF_SIMT_BWD = 1,  // fp32 [tap'][cout][cin], tap' = flipped (transposed conv)
  WF_TC_FWD = 2,    // split-bf16 [tap][cout][cin]           (K-major B operand of the forward GEMM)
  WF_TC_BWD = 3,    // split-bf16 [tap'][cin][cout]          (K-major B operand of the dgrad GEMM)
  WF_TC_FWD3 = 4,   // as WF_TC_FWD with three bf16 planes (fp32-exact forward operands)
  WF_TC_FWDH = 5,   // as WF_TC_FWD with two IEEE half planes of 2^k * w (Layer::wpow): the default forward operands
  WF_TC_BWDH = 6,   // as WF_TC_BWD with two IEEE half planes of 2^k * w: B operand of the two-product backward
  WF_TC_BWDF8 = 7,  // as WF_TC_BWD in the fp16 + fp8 layout (high fp16 plane + E4M3 [low | high] byte plane) of 2^k * w
};
enum WeightSign : int { WS_ALL = 0, WS_PLUS = 1, WS_MINUS = 2 };

// w_hwio: device fp32 [3,3,cin,cout]. out: 9*cin*cout elements in `fmt`.
// planes: bf16 planes written for the tensor-core formats (2 or 3; ignored for the fp32 formats), or kPlanesF16x2:
// two IEEE half planes of `scale` * w.
int prep_weights(const float* w_hwio, void* out, int cin, int cout, int fmt, int sign, cudaStream_t s, int taps = 9,
                 int planes = 2, float scale = 1.f);

// Backward weights of the alpha-beta rule with beta != 0, stacked along K (2*cout input channels):
//   k <  cout : scale_a * sign_a(W)      k >= cout : scale_b * sign_b(W)
// fmt: WF_SIMT_BWD -> fp32 [tap'][2*cout][cin]; WF_TC_BWD -> split-bf16 [tap'][cin][2*cout].
// half_planes (WF_TC_BWD only): 1 = two IEEE half planes instead of two bf16 planes (fold 2^wpow into the scales);
// 2 = the fp16 + fp8 layout.
int prep_weights_dual(const float* w_hwio, void* out, int cin, int cout, int fmt, int sign_a, float scale_a, int sign_b,
                      float scale_b, cudaStream_t s, int half_planes = 0);

// 2x2/2 max-pool of `act` [items,H,W,C] (storage-typed). If `pooled` != null writes [items,H/2,W/2,C];
// if `G` != null (the layer's multiplier in the g_offset layout with up = 2) also writes its compact form: `Gc`
// [items][C/16][H/2][W/2][16] = the value at the first maximum of each 2x2 window (TF MaxPoolGrad routing,
// innvestigate relevance_analyzer.py:459-480 -- every other entry routes nothing) and, if given, `Gidx`
// [items][C/16][H/2][W/2] words with 2 bits per channel = window position (sy * 2 + sx) of that maximum.
int pool_mask(const void* act, size_t act_elems, int planes /*0 = fp32, 2, 3*/, void* pooled, size_t pooled_elems, const float* G,
              float* Gc, unsigned* Gidx, int items, int H, int W, int C, cudaStream_t s);

// msg[item] = (relu?)(R[item]) * M[img_index[item]]   ([items, hw, hw, C]); msg storage-typed.
// M2 != null: dual message with 2*C channels [R*M | R*M2].
int seed_message(const float* R, const float* M, const float* M2, const int* img_index, void* msg, size_t msg_elems, bool split,
                 int items, int pix, int C, int relu, cudaStream_t s);

// Two-product backward: the seed as ONE fp16 plane scaled by a power of two per item. mx_true [items] must be zero on
// entry (receives max |value| per item); mx_out [items] receives the stored plane's maximum, kt_out [items] log2(scale).
int seed_message_scaled(const float* R, const float* M, const float* M2, const int* img_index, void* msg, int items, int pix,
                        int C, int relu, unsigned* mx_true, unsigned* mx_out, int* kt_out, int target_exp, cudaStream_t s,
                        int fp8_planes = 0);   // fp8_planes: the fp16 + fp8 layout (epilogue.cuh: StoreH1F8) instead of one fp16 plane

// Last transposed conv (64 -> 3 channels) + input re-weighting:
//   c_a = Wa^T (*) s,  c_b = Wb^T (*) s (only if Wb != null)
//   out = mult ? (x >= 0 ? x * c_a : x * c_b) : c_a            x = image[img_index[item]]
// Wa/Wb: fp32 [tap'][C][3] (WF_SIMT_BWD of the first layer).
int last_dgrad(const float* msg, const float* Wa, const float* Wb, const float* images, const int* img_index, float* out,
               int items, int H, int W, int C, int mult, cudaStream_t s);

// [x] (3 ch) -> [x+, x-] (6 ch), fp32 NHWC.
int make_posneg(const float* x, float* out, size_t pixels, cudaStream_t s);

int f32_to_split(const float* in, void* out, size_t n, cudaStream_t s, int planes = 2);
int split_to_f32(const void* in, float* out, size_t n, cudaStream_t s);

}  // namespace lrpcap
