// fp32 SIMT implicit-GEMM 3x3 / 1x1 convolution (exact-fp32 precision mode; also the 3-channel first
// VGG16 layer in the tensor-core mode, whose K = 27 is not a tensor-core shape).
// Same epilogues as the tcgen05 kernel (epilogue.cuh). Tile: 8x8 pixels x 64 output channels per CTA,
// 256 threads, 4 pixels x 4 channels per thread, 8-channel K chunks staged in shared memory with halo.
#include "epilogue.cuh"

namespace lrpcap {

namespace {

constexpr int kTP = 8;    // tile side in pixels
constexpr int kTN = 64;   // output channels per CTA
constexpr int kKC = 8;    // K chunk (input channels)

template <int MODE, class ST>
__global__ void __launch_bounds__(256)
simt_conv_kernel(const float* __restrict__ A, const float* __restrict__ B, int H, int W, int C, int taps, int Nout,
                 int tiles_x, int tiles_y, const EpiDev e) {
  __shared__ float As[(kTP + 2) * (kTP + 2)][kKC];
  __shared__ __align__(16) float Bs[9][kKC][kTN];

  const int R = (taps == 9) ? 1 : 0;
  const int PS = kTP + 2 * R;
  int bid = blockIdx.x;
  const int tiles = tiles_x * tiles_y;
  const int item = bid / tiles;
  bid -= item * tiles;
  const int y0 = (bid / tiles_x) * kTP;
  const int x0 = (bid % tiles_x) * kTP;
  const int n0 = blockIdx.y * kTN;

  const int tid = threadIdx.x;
  const int cg = tid & 15;
  const int pg = tid >> 4;
  const int prow = pg >> 1;
  const int pcol0 = (pg & 1) * 4;

  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

  for (int c0 = 0; c0 < C; c0 += kKC) {
    __syncthreads();
    for (int idx = tid; idx < PS * PS * kKC; idx += 256) {
      const int k = idx % kKC;
      const int pix = idx / kKC;
      const int py = pix / PS, px = pix - py * PS;
      const int gy = y0 + py - R, gx = x0 + px - R;
      const int c = c0 + k;
      float val = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W && c < C)
        val = __ldg(A + (((size_t)item * H + gy) * W + gx) * C + c);
      As[pix][k] = val;
    }
    for (int idx = tid; idx < taps * kKC * kTN; idx += 256) {
      const int n = idx % kTN;
      const int k = (idx / kTN) % kKC;
      const int tap = idx / (kTN * kKC);
      const int c = c0 + k;
      float val = 0.f;
      if (c < C && n0 + n < Nout) val = __ldg(B + ((size_t)tap * C + c) * Nout + n0 + n);
      Bs[tap][k][n] = val;
    }
    __syncthreads();
    // two-level summation: this 8-channel chunk (<= 72 terms) is summed on its own and then added to the running
    // total, so rounding error grows with the number of chunks rather than with K
    float part[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) part[j][i] = 0.f;
    for (int tap = 0; tap < taps; ++tap) {
      const int dy = (taps == 9) ? tap / 3 - 1 : 0;
      const int dx = (taps == 9) ? tap % 3 - 1 : 0;
      const int rbase = (prow + R + dy) * PS + pcol0 + R + dx;
#pragma unroll
      for (int k = 0; k < kKC; ++k) {
        const float4 b = *reinterpret_cast<const float4*>(&Bs[tap][k][cg * 4]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = As[rbase + j][k];
          part[j][0] = fmaf(a, b.x, part[j][0]);
          part[j][1] = fmaf(a, b.y, part[j][1]);
          part[j][2] = fmaf(a, b.z, part[j][2]);
          part[j][3] = fmaf(a, b.w, part[j][3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] += part[j][i];
  }

  const int y = y0 + prow;
  const int n = n0 + cg * 4;
  if (y < H && n < Nout) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = x0 + pcol0 + j;
      if (x < W) epi_apply<MODE, 4, ST>(e, H, W, Nout, item, y, x, n, acc[j]);
    }
  }
}

template <int MODE, class ST>
int launch_t(const SimtConvArgs& a, const EpiDev& e, cudaStream_t stream) {
  const int tiles_x = ceil_div(a.W, kTP), tiles_y = ceil_div(a.H, kTP);
  const long long bx = (long long)a.n_items * tiles_x * tiles_y;
  LRPCAP_REQUIRE(bx > 0 && bx < (1ll << 31), kErrShape, "simt_conv: grid of %lld blocks out of range", bx);
  dim3 grid((unsigned)bx, (unsigned)ceil_div(a.Nout, kTN));
  simt_conv_kernel<MODE, ST><<<grid, 256, 0, stream>>>(a.A, a.B, a.H, a.W, a.C, a.taps, a.Nout, tiles_x, tiles_y, e);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

template <class ST>
int launch_mode(const SimtConvArgs& a, const EpiDev& e, cudaStream_t stream) {
  switch (a.epi.mode) {
    case EPI_FWD_TRUE: return launch_t<EPI_FWD_TRUE, ST>(a, e, stream);
    case EPI_FWD_ZACT: return launch_t<EPI_FWD_ZACT, ST>(a, e, stream);
    case EPI_BWD: return launch_t<EPI_BWD, ST>(a, e, stream);
    case EPI_RAW: return launch_t<EPI_RAW, ST>(a, e, stream);
  }
  set_last_error("simt_conv: unknown epilogue mode %d", a.epi.mode);
  return kErrInvalidArg;
}

}  // namespace

int simt_conv_launch(const SimtConvArgs& a, cudaStream_t stream) {
  LRPCAP_REQUIRE(a.A && a.B, kErrInvalidArg, "simt_conv: null operand");
  LRPCAP_REQUIRE(a.taps == 9 || a.taps == 1, kErrShape, "simt_conv: taps must be 1 or 9");
  LRPCAP_REQUIRE(a.n_items > 0 && a.H > 0 && a.W > 0 && a.C > 0, kErrShape, "simt_conv: empty problem");
  LRPCAP_REQUIRE(a.Nout > 0 && a.Nout % 4 == 0, kErrShape, "simt_conv: Nout=%d must be a multiple of 4", a.Nout);
  EpiDev e;
  LRPCAP_TRY(make_epi_dev(a.epi, &e));
  if (a.out_planes == 2) return launch_mode<StoreSplit>(a, e, stream);
  if (a.out_planes == 3) return launch_mode<StoreSplit3>(a, e, stream);
  if (a.out_planes == kPlanesF16x2) return launch_mode<StoreSplitH>(a, e, stream);
  return launch_mode<StoreF32>(a, e, stream);
}

}  // namespace lrpcap
