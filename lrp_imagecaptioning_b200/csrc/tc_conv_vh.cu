// tcgen05 implicit-GEMM 3x3 (transposed) convolution, "vertical halo" variant for the wide, shallow layers
// (224x224 and 112x112 maps, 64 / 128 output channels) where the generic kernel (tc_conv.cu) is bound by
// L2 -> shared-memory operand traffic: it re-fetches the activation tile for each of the 9 taps and the weight tile
// for every 128-pixel tile.
//
// Same contraction and the same fused epilogues as tc_conv.cu (reference: iNNvestigate GradientWRT on a conv layer,
// innvestigate/layers.py:138-157 -> utils/keras/backend.py:58-60), different operand staging:
//   * a CTA tile is 16 x 16 pixels = two 128-row MMA tiles side by side; both use every weight tile once
//     it is in shared memory (weight traffic per pixel halves);
//   * an MMA tile is 8 x 16 pixels (8 wide, 16 tall); per (64-channel block, dx, MMA tile) ONE activation patch of
//     8 x 18 pixels is loaded (TMA box shifted by dx, rows y0-1 .. y0+16, OOB zero-fill = 'same' padding). A patch row
//     is 8 pixels = one swizzle group of 8 x 128 B, so the operand for tap (dy, dx) is the same patch at byte offset
//     dy * 1024 -- a plain SWIZZLE_128B K-major descriptor. 3 patches of 18 rows replace 9 tiles of 16 rows: 2.7x less
//     activation traffic. Patches are 18 / 36 KB ring slots (one or two planes; 4 to 8 of them, vh_rings()): with one
//     72 KB patch per dx and 2 slots the ring was shallower than the TMA latency and the tensor pipe idled 45 % of the time;
//   * NCAT (64 output channels): the weight planes [hi ; lo] sit back to back in shared memory and are issued as ONE
//     N = 128 operand against A_hi (accumulator columns [0,64) = hi*hi, [64,128) = hi*lo) plus one N = 64 MMA
//     A_lo * B_hi; the epilogue adds the two column blocks. Two MMAs instead of three, and a third less
//     shared-memory operand bandwidth, which is what bounds N = 64 MMAs.
// Rings: activation patches (NA slots) and weight taps (NB slots) are separate mbarrier rings fed by the TMA warp
// (converged, one elected lane issues: tc_ptx.cuh).
#include "epilogue.cuh"
#include "tc_ptx.cuh"
#include <cstdlib>
#include <type_traits>

namespace lrpcap {

namespace {

using namespace tcptx;

constexpr int kEpiWarps = 8;
constexpr int kEpiWarp0 = 4;                            // warps 0..3: producer warpgroup (TMA thread, MMA thread, two idle warps)
constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);
constexpr int kProducerRegs = 96, kEpilogueRegs = 200;  // setmaxnreg: the epilogue warpgroups take the producers' registers
constexpr int kTW = 16, kTH = 16, kTM = 2;              // CTA tile: 16 x 16 pixels = kTM MMA tiles of 8 (wide) x 16
constexpr int kMW = kTW / kTM;                          // MMA tile width: 8 pixels = one 1024 B swizzle group per row
constexpr int kPatchRows = kTH + 2;                     // 18 pixel rows of 8 pixels
constexpr int kAPlane = kPatchRows * kMW * 128;         // 18,432 B per bf16 plane (multiple of 1024)
constexpr int kNAMax = 8;
constexpr int kMaxNB = 8;
constexpr int kSmemLimit = 227 * 1024;

struct VhGeom {
  int H, W, tiles_x, tiles_y, cblocks, Nout, n_items, n_tiles_n, NA, NB;   // NA / NB: patch / weight-tap ring slots
};
struct VhMaps {
  CUtensorMap a[2];
  CUtensorMap b[2];
};

struct VhTile {
  int item, x0, y0, n0;
};
// `m_tile` = index over (item, tile row, tile column); CTA pairs take two consecutive ones per channel tile
__device__ __forceinline__ VhTile vh_tile(const VhGeom& g, int m, int n_tile, int BN) {
  VhTile t;
  const int per_item = g.tiles_x * g.tiles_y;
  t.item = m / per_item;
  m -= t.item * per_item;
  t.x0 = (m % g.tiles_x) * kTW;
  t.y0 = (m / g.tiles_x) * kTH;
  t.n0 = n_tile * BN;
  return t;
}

// A1: two-product mode (tc_conv.cu): ONE fp16 message plane x [B_hi ; B_lo]. NCAT: a single N = 128 MMA per K slice;
// BN = 128: two N = 128 MMAs into the same accumulator. The patch slots are half as large, so the ring is deeper.
// F8: fp16 + fp8 mode (tc_conv.cu, epilogue.cuh: StoreH1F8): plane 0 of the patches and weight taps is fp16, plane 1 the
// E4M3 byte plane (same tile sizes: 128 bytes per pixel and 64-channel block); per K slice one kind::f16 and one
// kind::f8f6f4 MMA.
// SM2: CTA pairs (cta_group::2, clusters of two; tc_conv.cu has the protocol): each CTA owns one 16 x 16 pixel tile and
// stages its own patches, the pair shares every weight tap -- each CTA stages HALF of its operand rows (NCAT: CTA 0 the
// high plane, CTA 1 the low plane; otherwise rows [rank BN/2, +BN/2) of both planes) and the leader issues M = 256 MMAs.
// Half-size weight slots are what lets the rings be deep enough: one CTA alone has room for only two 128-channel taps.
//
// MMA order inside one (channel block, dx) group: tap by tap over both MMA tiles (each weight slot is released after its
// tap, the two patches after the group). Issuing tile by tile with per-patch release (a deeper patch pipeline) was measured
// and makes no difference once the issue loop runs on the uniform datapath (tc_ptx.cuh) -- before that, the ~110 cycles the
// issuing thread spent per MMA, not the memory pipeline, were what paced this kernel (ncu: tensor pipe 29 % busy).
template <int BN, int MODE, bool NCAT, bool A1 = false, bool F8 = false, bool SM2 = false>
__global__ void __launch_bounds__(kThreads, 1)
tc_conv_vh_kernel(const __grid_constant__ VhMaps tm, const VhGeom g, const EpiDev e, const int total_tiles) {
  static_assert(!F8 || (!NCAT && !A1), "fp16 + fp8 mode: plain two-plane staging");
  static_assert(!SM2 || A1 || F8, "CTA pairs: two-product and fp16 + fp8 modes");
  constexpr int AP = A1 ? 1 : 2;
  constexpr int kASlot = AP * kAPlane;
  using ST = typename std::conditional<F8, StoreH1F8, typename std::conditional<A1, StoreH1, StoreSplit>::type>::type;
  constexpr int kBRows = SM2 ? (NCAT ? BN : BN / 2) : BN;   // weight rows per plane staged by this CTA
  constexpr int kBPlanes = (SM2 && NCAT) ? 1 : 2;           // planes staged by this CTA
  constexpr int kBPlane = kBRows * 128;
  constexpr int kBSlot = kBPlanes * kBPlane;
  constexpr int ACC = NCAT ? 2 * BN : BN;               // accumulator columns per MMA tile
  constexpr int kTmemCols = 2 * kTM * ACC;              // double-buffered
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM allocation must be a power of two <= 512");

  const int NA = g.NA, NB = g.NB;
  const uint32_t rank = SM2 ? cluster_ctarank() : 0u;
  const int tile0 = SM2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = SM2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto coord = [&](int tile) {
    const int n_tile = tile % g.n_tiles_n, m = tile / g.n_tiles_n;
    return vh_tile(g, SM2 ? 2 * m + (int)rank : m, n_tile, BN);
  };

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + NA * kASlot;
  uint64_t* afull = reinterpret_cast<uint64_t*>(b_ring + NB * kBSlot);
  uint64_t* aempty = afull + kNAMax;
  uint64_t* bfull = aempty + kNAMax;
  uint64_t* bempty = bfull + kMaxNB;
  uint64_t* tfull = bempty + kMaxNB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform: the role branches stay on the uniform datapath
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int p = 0; p < 2; ++p) {
      prefetch_tmap(&tm.a[p]);
      prefetch_tmap(&tm.b[p]);
    }
    for (int s = 0; s < NA; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&aempty[s], 1);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(&bfull[s], 1);
      mbar_init(&bempty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], SM2 ? 2 * kEpiWarps : kEpiWarps);   // pair: the leader's barrier counts both CTAs' warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (SM2) cluster_sync_all();
  if (warp == 1) {
    if (SM2) tmem_alloc2(tmem_slot, kTmemCols);
    else tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarp0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
  if (warp == 0) {
    {   // the whole warp, converged: the TMA wrappers elect the issuing lane (tc_ptx.cuh)
      // ---------------- TMA producer: patches and weight taps in the order the MMA thread waits for them ----------------
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0;   // ring slot and phase bit, advanced by hand (NA / NB are run-time values)
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const VhTile tc = coord(tile);
        for (int cb = 0; cb < g.cblocks; ++cb) {
          for (int dxi = 0; dxi < 3; ++dxi) {
            for (int j = 0; j < kTM; ++j) {
              mbar_wait(&aempty[sa], pha ^ 1u);
              if (!SM2) mbar_expect_tx_e(&afull[sa], (uint32_t)kASlot);
              else if (rank == 0) mbar_expect_tx_e(&afull[sa], 2u * (uint32_t)kASlot);   // both CTAs' bytes land on the leader's barrier
              uint8_t* ap = a_ring + sa * kASlot;
#pragma unroll
              for (int p = 0; p < AP; ++p) {   // byte planes (F8, plane 1) count the innermost coordinate in bytes
                const int c0 = cb * ((F8 && p == 1) ? 128 : kBlockK);
                if (SM2) tma2_load_4d(&tm.a[p], ap + p * kAPlane, &afull[sa], c0, tc.x0 + kMW * j + dxi - 1, tc.y0 - 1, tc.item);
                else tma_load_4d_e(&tm.a[p], ap + p * kAPlane, &afull[sa], c0, tc.x0 + kMW * j + dxi - 1, tc.y0 - 1, tc.item);
              }
              if (j == kTM - 1) {   // the three weight taps of this dx, used by both MMA tiles
                for (int dyi = 0; dyi < 3; ++dyi) {
                  mbar_wait(&bempty[sb], phb ^ 1u);
                  if (!SM2) mbar_expect_tx_e(&bfull[sb], (uint32_t)kBSlot);
                  else if (rank == 0) mbar_expect_tx_e(&bfull[sb], 2u * (uint32_t)kBSlot);
                  uint8_t* bp = b_ring + sb * kBSlot;
                  const int tap = dyi * 3 + dxi;
#pragma unroll
                  for (int p = 0; p < 2; ++p) {
                    const int c0 = cb * ((F8 && p == 1) ? 128 : kBlockK);
                    if (!SM2) {
                      tma_load_2d_e(&tm.b[p], bp + p * kBPlane, &bfull[sb], c0, tap * g.Nout + tc.n0);
                    } else if (NCAT) {   // CTA `rank` stages plane `rank`: rows [0, BN) of the pair's N = 2 BN operand are the high plane
                      if (p == (int)rank) tma2_load_2d(&tm.b[p], bp, &bfull[sb], c0, tap * g.Nout + tc.n0);
                    } else {
                      tma2_load_2d(&tm.b[p], bp + p * kBPlane, &bfull[sb], c0, tap * g.Nout + tc.n0 + (int)rank * (BN / 2));
                    }
                  }
                  if (++sb == (uint32_t)NB) {
                    sb = 0;
                    phb ^= 1u;
                  }
                }
              }
              if (++sa == (uint32_t)NA) {
                sa = 0;
                pha ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ---------------- MMA issuer (pairs: the leader's, for both CTAs) ----------------
      // the whole warp runs this loop converged; the MMA / commit wrappers elect the issuing lane (tc_ptx.cuh)
      constexpr uint32_t idesc_n = make_idesc(SM2 ? 256 : 128, BN, A1 || F8);
      constexpr uint32_t idesc_cat = make_idesc(SM2 ? 256 : 128, NCAT ? 2 * BN : BN, A1 || F8);
      auto mma16 = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
        if (SM2) umma2_bf16(d, a, b, id, acc);
        else umma_bf16(d, a, b, id, acc);
      };
      auto mma8 = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
        if (SM2) umma2_f8(d, a, b, id, acc);
        else umma_f8(d, a, b, id, acc);
      };
      auto commit = [](uint64_t* bar) {
        if (SM2) umma_commit2_mc(bar, (uint16_t)3);
        else umma_commit(bar);
      };
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0, tl = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++tl) {
        const uint32_t buf = tl & 1u;
        mbar_wait(&tempty[buf], ((tl >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * (kTM * ACC);
        for (int cb = 0; cb < g.cblocks; ++cb) {
          for (int dxi = 0; dxi < 3; ++dxi) {
            const uint32_t accum = (cb != 0 || dxi != 0) ? 1u : 0u;
            uint32_t as_[kTM], ap_[kTM], bs_[3], bp_[3];   // slots and phase bits of this group's patches and taps
#pragma unroll
            for (int j = 0; j < kTM; ++j) {
              as_[j] = sa;
              ap_[j] = pha;
              if (++sa == (uint32_t)NA) {
                sa = 0;
                pha ^= 1u;
              }
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              bs_[d] = sb;
              bp_[d] = phb;
              if (++sb == (uint32_t)NB) {
                sb = 0;
                phb ^= 1u;
              }
            }
            // K slice k of weight tap dyi for MMA tile j
            auto slice = [&](int j, int dyi, int k) {
              const uint32_t bbase = smem_u32(b_ring + bs_[dyi] * kBSlot);
              const uint64_t db_hi = make_desc_sw128(bbase);                 // NCAT: the same start, N = 2 BN rows
              const uint64_t db_lo = make_desc_sw128(bbase + kBPlane);
              const uint32_t abase = smem_u32(a_ring + as_[j] * kASlot);
              const uint64_t da_hi = make_desc_sw128(abase + (uint32_t)dyi * 1024u);
              const uint64_t da_lo = make_desc_sw128(abase + (A1 ? 0 : kAPlane) + (uint32_t)dyi * 1024u);
              const uint32_t d = tmem_d + j * ACC;
              const uint64_t adv = (uint64_t)(k * 2);
              const uint32_t acc_k = (accum != 0u || dyi != 0 || k != 0) ? 1u : 0u;   // the tile's first MMA overwrites
              if (F8) {
                mma16(d, da_hi + adv, db_hi + adv, idesc_n, acc_k);     // fp16 message x fp16 high weights
                mma8(d, da_lo + adv, db_lo + adv, idesc_n, 1u);         // [top bits | residual] x [low | high], E4M3
              } else if (A1) {
                if (NCAT) {
                  mma16(d, da_hi + adv, db_hi + adv, idesc_cat, acc_k);   // [A*hi | A*lo]
                } else {
                  mma16(d, da_hi + adv, db_lo + adv, idesc_n, acc_k);
                  mma16(d, da_hi + adv, db_hi + adv, idesc_n, 1u);
                }
              } else if (NCAT) {
                mma16(d, da_hi + adv, db_hi + adv, idesc_cat, acc_k);   // [hi*hi | hi*lo]
                mma16(d, da_lo + adv, db_hi + adv, idesc_n, 1u);        // += lo*hi into the first block
              } else {
                mma16(d, da_hi + adv, db_lo + adv, idesc_n, acc_k);
                mma16(d, da_lo + adv, db_hi + adv, idesc_n, 1u);
                mma16(d, da_hi + adv, db_hi + adv, idesc_n, 1u);
              }
            };
            auto issue = [&](int j, int dyi) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) slice(j, dyi, k);
            };
            auto wait_a = [&](int j) {
              mbar_wait(&afull[as_[j]], ap_[j]);
              tc_fence_after();
            };
            auto wait_b = [&](int dyi) {
              mbar_wait(&bfull[bs_[dyi]], bp_[dyi]);
              tc_fence_after();
            };
            for (int j = 0; j < kTM; ++j) wait_a(j);
            for (int dyi = 0; dyi < 3; ++dyi) {
              wait_b(dyi);
              if (g.cblocks == 1) {   // measured: alternating the two accumulators per K slice helps the K = 576 layer only
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
#pragma unroll
                  for (int j = 0; j < kTM; ++j) slice(j, dyi, k);
              } else {
#pragma unroll
                for (int j = 0; j < kTM; ++j) issue(j, dyi);
              }
              commit(&bempty[bs_[dyi]]);
            }
            for (int j = 0; j < kTM; ++j) commit(&aempty[as_[j]]);
          }
        }
        commit(&tfull[buf]);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ---------------- epilogue: warp w owns TMEM lanes [32 (w % 4), +32) of MMA tile (w - 4) / 4 ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpilogueRegs));
    const int q = warp & 3;
    const int j = (warp - kEpiWarp0) >> 2;
    const int r = q * 32 + lane;
    const int ty = r >> 3, tx = kMW * j + (r & 7);
    uint32_t tl = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++tl) {
      const VhTile tc = coord(tile);
      const int y = tc.y0 + ty, x = tc.x0 + tx;
      const bool valid = (y < g.H) && (x < g.W);
      const uint32_t buf = tl & 1u;
      if (MODE == EPI_BWD) {   // multipliers of the NEXT tile (and of the first one) -> L2, one tile period ahead of use
        for (int pt = (tl == 0 ? tile : tile + tile_step); pt <= tile + tile_step && pt < total_tiles; pt += tile_step) {
          const VhTile nt = coord(pt);
          if (nt.y0 + ty < g.H && nt.x0 + tx < g.W)
            for (int c = 0; c < BN / 16; ++c)
              epi_prefetch_bwd(e, g.H, g.W, g.Nout, nt.item, nt.y0 + ty, nt.x0 + tx, nt.n0 + c * 16);
        }
      }
      mbar_wait(&tfull[buf], (tl >> 1) & 1u);
      tc_fence_after();
      const uint32_t lane_base = tmem_base + buf * (kTM * ACC) + j * ACC + ((uint32_t)(q * 32) << 16);
      auto load_acc = [&](int c, float (&v)[16]) {
        if (NCAT) {
          float w[16];
          tmem_ld16x2(lane_base + (uint32_t)(c * 16), lane_base + (uint32_t)(BN + c * 16), v, w);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += w[i];
        } else {
          tmem_ld16(lane_base + (uint32_t)(c * 16), v);
        }
      };
      if constexpr (MODE == EPI_BWD) {
        if (e.up == 2)
          epi_bwd_chunks<2, BN / 16, ST>(e, g.H, g.W, g.Nout, tc.item, y, x, tc.n0, 16, valid, load_acc);
        else
          epi_bwd_chunks<1, BN / 16, ST>(e, g.H, g.W, g.Nout, tc.item, y, x, tc.n0, 16, valid, load_acc);
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 16; ++c) {
          float v[16];
          __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge after the predicated stores below
          load_acc(c, v);
          if (valid) epi_apply<MODE, 16, ST>(e, g.H, g.W, g.Nout, tc.item, y, x, tc.n0 + c * 16, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (SM2 && rank != 0) mbar_arrive_cluster(&tempty[buf], 0u);
        else mbar_arrive(&tempty[buf]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (SM2) cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer can still reach its barriers / accumulator
  if (warp == 1) {
    if (SM2) tmem_dealloc2(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Ring depths for the shared memory one CTA has: at least 4 patch slots, 4 weight-tap slots when they fit, the rest to patches.
static void vh_rings(int a_slot, int b_slot, int* NA, int* NB) {
  const int room = kSmemLimit - 1024 - 512;
  int na = (room - 4 * b_slot) / a_slot;
  if (na > kNAMax) na = kNAMax;
  if (na < 4) na = 4;
  int nb = (room - na * a_slot) / b_slot;
  if (nb > kMaxNB) nb = kMaxNB;
  *NA = na;
  *NB = nb;
}

template <int BN, int MODE, bool NCAT, bool A1, bool F8, bool SM2>
int launch_vh(const VhMaps& tm, VhGeom g, const EpiDev& e, cudaStream_t stream) {
  constexpr int kBSlot = SM2 ? BN * 128 : 2 * BN * 128;   // pairs: each CTA stages half of the tap's operand rows
  constexpr int kASlot = (A1 ? 1 : 2) * kAPlane;
  vh_rings(kASlot, kBSlot, &g.NA, &g.NB);
  LRPCAP_REQUIRE(g.NB >= 2, kErrUnsupported, "tc_conv_vh: no room for a weight ring (BN=%d)", BN);
  const int smem = g.NA * kASlot + g.NB * kBSlot + 1024 + 512;
  auto kern = tc_conv_vh_kernel<BN, MODE, NCAT, A1, F8, SM2>;
  static int smem_state[kMaxDevices] = {};
  LRPCAP_CUDA(ensure_dynamic_smem(kern, smem, smem_state));
  const long long tiles_m = (long long)g.n_items * g.tiles_x * g.tiles_y;
  const long long tiles = (SM2 ? tiles_m / 2 : tiles_m) * g.n_tiles_n;   // pairs: pair-tiles
  LRPCAP_REQUIRE(tiles > 0 && tiles < (1ll << 30), kErrShape, "tc_conv_vh: %lld tiles out of range", tiles);
  const int num_sms = device_sm_count();
  if constexpr (!SM2) {
    const unsigned grid = (unsigned)(tiles < num_sms ? tiles : num_sms);
    kern<<<grid, kThreads, smem, stream>>>(tm, g, e, (int)tiles);
    LRPCAP_CUDA(cudaGetLastError());
    return kOk;
  } else {
  LRPCAP_REQUIRE(tiles_m % 2 == 0, kErrShape, "tc_conv_vh: %lld pixel tiles do not pair up", tiles_m);
  const int pairs_max = num_sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2u * (unsigned)(tiles < pairs_max ? tiles : pairs_max));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LRPCAP_CUDA(cudaLaunchKernelEx(&cfg, kern, tm, g, e, (int)tiles));
  return kOk;
  }
}

template <int BN, bool NCAT, bool A1, bool F8 = false, bool SM2 = false>
int launch_vh_mode(int mode, const VhMaps& tm, const VhGeom& g, const EpiDev& e, cudaStream_t stream) {
  switch (mode) {
    case EPI_BWD: return launch_vh<BN, EPI_BWD, NCAT, A1, F8, SM2>(tm, g, e, stream);
    case EPI_RAW: return launch_vh<BN, EPI_RAW, NCAT, A1, F8, SM2>(tm, g, e, stream);
  }
  set_last_error("tc_conv_vh: epilogue mode %d not instantiated", mode);
  return kErrUnsupported;
}

bool vh_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* v = std::getenv("LRPCAP_TC_VH");
    on = (v && v[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

}  // namespace

bool tc_conv_vh_eligible(const TcConvArgs& a, int BN) {
  return vh_enabled() && a.taps == 9 && (a.planes == 2 || a.planes == kPlanesH1x2 || a.planes == kPlanesH1F8) && a.promote_every <= 0 &&
         (BN == 64 || BN == 128) && a.W % kTW == 0 && a.H % kTH == 0 && (a.epi.mode == EPI_BWD || a.epi.mode == EPI_RAW);
}

int tc_conv_vh_launch(const TcConvArgs& a, int BN, cudaStream_t stream) {
  LRPCAP_REQUIRE(tc_conv_vh_eligible(a, BN), kErrUnsupported, "tc_conv_vh: shape not eligible");
  VhGeom g;
  g.H = a.H;
  g.W = a.W;
  g.tiles_x = a.W / kTW;
  g.tiles_y = a.H / kTH;
  g.cblocks = a.C / kBlockK;
  g.Nout = a.Nout;
  g.n_items = a.n_items;
  g.n_tiles_n = a.Nout / BN;
  g.NA = g.NB = 0;
  const bool a1 = a.planes == kPlanesH1x2, f8 = a.planes == kPlanesH1F8;
  // CTA pairs when the 16 x 16 pixel tiles pair up (two-product and fp16 + fp8 modes)
  static const bool vh_pairs = [] { const char* v = std::getenv("LRPCAP_VH_2SM"); return !(v && v[0] == '0'); }();
  const bool pair = tc_pair_enabled() && vh_pairs && (a1 || f8) && ((long long)a.n_items * g.tiles_x * g.tiles_y) % 2 == 0;
  const bool ncat = a1 && BN == 64;
  const int brows = pair && !ncat ? BN / 2 : BN;   // weight rows per TMA box
  const __nv_bfloat16* A0 = reinterpret_cast<const __nv_bfloat16*>(a.A);
  const __nv_bfloat16* B0 = reinterpret_cast<const __nv_bfloat16*>(a.B);
  VhMaps tm;
  if (f8) {   // plane 0: fp16 [.., C]; plane 1: bytes [.., 2 C] right behind it
    LRPCAP_TRY(make_map_act(&tm.a[0], A0, a.n_items, a.H, a.W, a.C, kMW, kPatchRows));
    LRPCAP_TRY(make_map_act_u8(&tm.a[1], reinterpret_cast<const uint8_t*>(A0) + a.A_elems * 2, a.n_items, a.H, a.W, 2 * a.C, kMW,
                               kPatchRows));
    LRPCAP_TRY(make_map_w(&tm.b[0], B0, a.taps * a.Nout, a.C, brows));
    LRPCAP_TRY(make_map_w_u8(&tm.b[1], reinterpret_cast<const uint8_t*>(B0) + a.B_elems * 2, a.taps * a.Nout, 2 * a.C, brows));
  } else
  for (int pl = 0; pl < 2; ++pl) {
    LRPCAP_TRY(make_map_act(&tm.a[pl], A0 + (size_t)(a1 ? 0 : pl) * a.A_elems, a.n_items, a.H, a.W, a.C, kMW, kPatchRows));
    LRPCAP_TRY(make_map_w(&tm.b[pl], B0 + (size_t)pl * a.B_elems, a.taps * a.Nout, a.C, brows));
  }
  EpiDev e;
  LRPCAP_TRY(make_epi_dev(a.epi, &e));
  if (f8) {
    if (BN == 64) return pair ? launch_vh_mode<64, false, false, true, true>(a.epi.mode, tm, g, e, stream)
                              : launch_vh_mode<64, false, false, true, false>(a.epi.mode, tm, g, e, stream);
    return pair ? launch_vh_mode<128, false, false, true, true>(a.epi.mode, tm, g, e, stream)
                : launch_vh_mode<128, false, false, true, false>(a.epi.mode, tm, g, e, stream);
  }
  if (a1) {
    if (BN == 64) return pair ? launch_vh_mode<64, true, true, false, true>(a.epi.mode, tm, g, e, stream)
                              : launch_vh_mode<64, true, true, false, false>(a.epi.mode, tm, g, e, stream);
    return pair ? launch_vh_mode<128, false, true, false, true>(a.epi.mode, tm, g, e, stream)
                : launch_vh_mode<128, false, true, false, false>(a.epi.mode, tm, g, e, stream);
  }
  if (BN == 64) return launch_vh_mode<64, true, false>(a.epi.mode, tm, g, e, stream);
  return launch_vh_mode<128, false, false>(a.epi.mode, tm, g, e, stream);
}

}  // namespace lrpcap
