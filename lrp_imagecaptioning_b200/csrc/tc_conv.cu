// tcgen05 implicit-GEMM convolution, split-bf16 operands, fp32 accumulation in TMEM (sm_100a).
//
// Replaces, for the VGG16 encoder of the reference, the TensorFlow ops emitted by
//   keras Conv2D forward                      (/root/reference/models/explainers.py:375 via _image_model.predict)
//   iNNvestigate GradientWRT on a conv layer  (innvestigate/layers.py:138-157 -> utils/keras/backend.py:58-60)
// and, with taps == 1, the dense contractions of the decoder relevance (explainers.py:156-165).
//
// Structure per CTA (128 threads, one 128-pixel x BN-channel output tile):
//   thread 0      : TMA producer. Per k-step (tap, 64-channel block) four cp.async.bulk.tensor loads
//                   (A_hi, A_lo as 4-D boxes shifted by the tap offset -- OOB rows/cols are zero-filled,
//                   which *is* the 'same' padding -- and B_hi, B_lo as 2-D boxes), 128B-swizzled.
//   thread 32     : MMA issuer. 4 K-slices x 3 tcgen05.mma (hi*hi, hi*lo, lo*hi) per k-step into one
//                   TMEM accumulator (128 lanes x BN fp32 columns); tcgen05.commit frees the smem stage.
//   all 4 warps   : epilogue. tcgen05.ld 32x32b.x16 -> registers -> fused rule arithmetic -> global.
#include "epilogue.cuh"
#include <cuda.h>

namespace lrpcap {

namespace {

constexpr int kBlockK = 64;                       // channels per k-step: 64 bf16 = 128 B = one swizzle row
constexpr int kATileBytes = 128 * 128;            // 128 rows x 128 B
constexpr uint32_t kSpinLimit = 1u << 26;

struct Geom {
  int H, W, TW, TH, tiles_x, tiles_y, cblocks, taps, Nout, n_items, n_tiles_n;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (8 rows * 128 B = 1024)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

template <int BN>
struct Cfg {
  static constexpr int kBTileBytes = BN * 128;
  static constexpr int kStageBytes = 2 * kATileBytes + 2 * kBTileBytes;
  static constexpr int kStages = (BN == 256) ? 2 : (BN == 128 ? 3 : 4);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// ------------------------------------------------------------------ the kernel
template <int BN, int MODE>
__global__ void __launch_bounds__(128, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, const Geom g,
               const EpiDev e) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* accum_bar = empty_bar + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;

  // tile coordinates: n-tile fastest, then spatial tile, then item
  int bid = blockIdx.x;
  const int n_tile = bid % g.n_tiles_n;
  bid /= g.n_tiles_n;
  const int tiles_per_item = g.tiles_x * g.tiles_y;
  const int item = bid / tiles_per_item;
  const int t_in = bid - item * tiles_per_item;
  const int x0 = (t_in % g.tiles_x) * g.TW;
  const int y0 = (t_in / g.tiles_x) * g.TH;
  const int n0 = n_tile * BN;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA_hi);
    prefetch_tmap(&tmA_lo);
    prefetch_tmap(&tmB_hi);
    prefetch_tmap(&tmB_lo);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_k = g.taps * g.cblocks;
  const uint32_t stage_tx = 2u * (uint32_t)(g.TW * g.TH) * 128u + 2u * (uint32_t)C::kBTileBytes;

  if (threadIdx.x == 0) {
    // ---------------- TMA producer ----------------
    for (int it = 0; it < num_k; ++it) {
      const int s = it % C::kStages;
      const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      uint8_t* st = smem + s * C::kStageBytes;
      mbar_expect_tx(&full_bar[s], stage_tx);
      const int tap = it / g.cblocks;
      const int cb = it - tap * g.cblocks;
      int dy = 0, dx = 0;
      if (g.taps == 9) {
        dy = tap / 3 - 1;
        dx = tap % 3 - 1;
      }
      tma_load_4d(&tmA_hi, st, &full_bar[s], cb * kBlockK, x0 + dx, y0 + dy, item);
      tma_load_4d(&tmA_lo, st + kATileBytes, &full_bar[s], cb * kBlockK, x0 + dx, y0 + dy, item);
      tma_load_2d(&tmB_hi, st + 2 * kATileBytes, &full_bar[s], cb * kBlockK, tap * g.Nout + n0);
      tma_load_2d(&tmB_lo, st + 2 * kATileBytes + C::kBTileBytes, &full_bar[s], cb * kBlockK, tap * g.Nout + n0);
    }
  } else if (threadIdx.x == 32) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = make_idesc(128, BN);
    for (int it = 0; it < num_k; ++it) {
      const int s = it % C::kStages;
      const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t a_hi = smem_u32(smem + s * C::kStageBytes);
      const uint64_t da_hi = make_desc_sw128(a_hi);
      const uint64_t da_lo = make_desc_sw128(a_hi + kATileBytes);
      const uint64_t db_hi = make_desc_sw128(a_hi + 2 * kATileBytes);
      const uint64_t db_lo = make_desc_sw128(a_hi + 2 * kATileBytes + C::kBTileBytes);
#pragma unroll
      for (int k = 0; k < kBlockK / 16; ++k) {
        const uint64_t adv = (uint64_t)(k * 2);  // 16 bf16 = 32 B = 2 x 16 B units inside the swizzle row
        umma_bf16(tmem_base, da_hi + adv, db_hi + adv, idesc, (it | k) != 0 ? 1u : 0u);
        umma_bf16(tmem_base, da_hi + adv, db_lo + adv, idesc, 1u);
        umma_bf16(tmem_base, da_lo + adv, db_hi + adv, idesc, 1u);
      }
      umma_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above have read it
    }
    umma_commit(accum_bar);        // accumulator complete
  }
  __syncwarp();

  // ---------------- epilogue (all 4 warps; warp w owns TMEM lanes [32w, 32w+32)) ----------------
  mbar_wait(accum_bar, 0);
  tc_fence_after();

  const int r = threadIdx.x;
  const int ty = r / g.TW;
  const int tx = r - ty * g.TW;
  const int y = y0 + ty, x = x0 + tx;
  const bool valid = (r < g.TW * g.TH) && (y < g.H) && (x < g.W);
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);

#pragma unroll 1
  for (int c = 0; c < BN / 16; ++c) {
    float v[16];
    __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the predicated stores below
    tmem_ld16(lane_base + (uint32_t)(c * 16), v);
    if (valid) epi_apply<MODE, 16, StoreSplit>(e, g.H, g.W, g.Nout, item, y, x, n0 + c * 16, v);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_map_act(CUtensorMap* m, const void* base, int n_items, int H, int W, int C, int TW, int TH) {
  EncodeTiledFn fn = get_encode_fn();
  LRPCAP_REQUIRE(fn != nullptr, kErrCuda, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_items};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LRPCAP_REQUIRE(r == CUDA_SUCCESS, kErrCuda, "cuTensorMapEncodeTiled(activation) failed: %d", (int)r);
  return kOk;
}

int make_map_w(CUtensorMap* m, const void* base, int rows, int C, int BN) {
  EncodeTiledFn fn = get_encode_fn();
  LRPCAP_REQUIRE(fn != nullptr, kErrCuda, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LRPCAP_REQUIRE(r == CUDA_SUCCESS, kErrCuda, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return kOk;
}

template <int BN, int MODE>
int launch_t(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
             const Geom& g, const EpiDev& e, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    LRPCAP_CUDA(cudaFuncSetAttribute(tc_conv_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::kSmemBytes));
    configured = true;
  }
  const long long blocks = (long long)g.n_items * g.tiles_x * g.tiles_y * g.n_tiles_n;
  LRPCAP_REQUIRE(blocks > 0 && blocks < (1ll << 31), kErrShape, "tc_conv: grid of %lld blocks out of range", blocks);
  tc_conv_kernel<BN, MODE><<<(unsigned)blocks, 128, C::kSmemBytes, stream>>>(a_hi, a_lo, b_hi, b_lo, g, e);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

template <int BN>
int launch_mode(int mode, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                const CUtensorMap& b_lo, const Geom& g, const EpiDev& e, cudaStream_t stream) {
  switch (mode) {
    case EPI_FWD_TRUE: return launch_t<BN, EPI_FWD_TRUE>(a_hi, a_lo, b_hi, b_lo, g, e, stream);
    case EPI_FWD_ZACT: return launch_t<BN, EPI_FWD_ZACT>(a_hi, a_lo, b_hi, b_lo, g, e, stream);
    case EPI_BWD: return launch_t<BN, EPI_BWD>(a_hi, a_lo, b_hi, b_lo, g, e, stream);
    case EPI_RAW: return launch_t<BN, EPI_RAW>(a_hi, a_lo, b_hi, b_lo, g, e, stream);
  }
  set_last_error("tc_conv: unknown epilogue mode %d", mode);
  return kErrInvalidArg;
}

}  // namespace

void tc_conv_tile(int H, int W, int* TW, int* TH) {
  if (W >= 16 && W % 16 == 0 && H >= 8) {
    *TW = 16;
    *TH = 8;
    return;
  }
  if (W <= 128) {
    int max_th = 128 / W;
    if (max_th > H) max_th = H;
    int th = max_th;
    while (th > 1 && H % th != 0) --th;
    *TW = W;
    *TH = th;
    return;
  }
  *TW = 16;
  *TH = 8;
}

int tc_conv_launch(const TcConvArgs& a, cudaStream_t stream) {
  LRPCAP_REQUIRE(a.A && a.B, kErrInvalidArg, "tc_conv: null operand");
  LRPCAP_REQUIRE(a.C > 0 && a.C % kBlockK == 0, kErrShape, "tc_conv: C=%d must be a positive multiple of 64", a.C);
  LRPCAP_REQUIRE(a.Nout > 0 && a.Nout % 64 == 0, kErrShape, "tc_conv: Nout=%d must be a positive multiple of 64", a.Nout);
  LRPCAP_REQUIRE(a.taps == 9 || a.taps == 1, kErrShape, "tc_conv: taps must be 1 or 9");
  LRPCAP_REQUIRE(a.n_items > 0 && a.H > 0 && a.W > 0, kErrShape, "tc_conv: empty problem");
  const int BN = (a.Nout % 256 == 0) ? 256 : (a.Nout % 128 == 0 ? 128 : 64);
  Geom g;
  g.H = a.H;
  g.W = a.W;
  tc_conv_tile(a.H, a.W, &g.TW, &g.TH);
  g.tiles_x = ceil_div(a.W, g.TW);
  g.tiles_y = ceil_div(a.H, g.TH);
  g.cblocks = a.C / kBlockK;
  g.taps = a.taps;
  g.Nout = a.Nout;
  g.n_items = a.n_items;
  g.n_tiles_n = a.Nout / BN;

  const __nv_bfloat16* A_hi = reinterpret_cast<const __nv_bfloat16*>(a.A);
  const __nv_bfloat16* B_hi = reinterpret_cast<const __nv_bfloat16*>(a.B);
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  LRPCAP_TRY(make_map_act(&ma_hi, A_hi, a.n_items, a.H, a.W, a.C, g.TW, g.TH));
  LRPCAP_TRY(make_map_act(&ma_lo, A_hi + a.A_elems, a.n_items, a.H, a.W, a.C, g.TW, g.TH));
  LRPCAP_TRY(make_map_w(&mb_hi, B_hi, a.taps * a.Nout, a.C, BN));
  LRPCAP_TRY(make_map_w(&mb_lo, B_hi + a.B_elems, a.taps * a.Nout, a.C, BN));

  const EpiParams& p = a.epi;
  EpiDev e;
  LRPCAP_TRY(make_epi_dev(p, &e));

  switch (BN) {
    case 256: return launch_mode<256>(p.mode, ma_hi, ma_lo, mb_hi, mb_lo, g, e, stream);
    case 128: return launch_mode<128>(p.mode, ma_hi, ma_lo, mb_hi, mb_lo, g, e, stream);
    default: return launch_mode<64>(p.mode, ma_hi, ma_lo, mb_hi, mb_lo, g, e, stream);
  }
}

}  // namespace lrpcap
