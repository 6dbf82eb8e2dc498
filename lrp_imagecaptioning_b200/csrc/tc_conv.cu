// tcgen05 implicit-GEMM convolution, split 16-bit operands (bf16 or IEEE half planes), fp32 accumulation in TMEM (sm_100a).
//
// Replaces, for the VGG16 encoder of the reference, the TensorFlow ops emitted by
//   keras Conv2D forward                      (/root/reference/models/explainers.py:375 via _image_model.predict)
//   iNNvestigate GradientWRT on a conv layer  (innvestigate/layers.py:138-157 -> utils/keras/backend.py:58-60)
// and, with taps == 1, the dense contractions of the decoder relevance (explainers.py:156-165).
//
// Structure per CTA (persistent, 384 threads = one producer warpgroup + two epilogue warpgroups, tiles of 128 pixels x BN channels):
//   warp 0        : TMA producer (the warp runs the loop converged, one elected lane issues: tc_ptx.cuh). Per k-step
//                   (tap, 64-channel block) up to four cp.async.bulk.tensor loads (the A planes as 4-D boxes shifted by
//                   the tap offset -- OOB rows/cols are zero-filled, which *is* the 'same' padding -- and the B planes
//                   as 2-D boxes), 128B-swizzled.
//   warp 1        : MMA issuer (converged, elected lane). 4 K-slices x 3 tcgen05.mma (hi*hi, hi*lo, lo*hi; 6 with three
//                   planes; 2 in the two-product and fp16 + fp8 modes) per k-step into one of two TMEM accumulators
//                   (128 lanes x BN fp32 columns each); tcgen05.commit frees the smem stage.
//   warps 2, 3    : idle (they complete the producer warpgroup, which gives most of its registers away: setmaxnreg)
//   warps 4..11   : epilogue (overlaps the next tile's MMAs). tcgen05.ld 32x32b.x16 -> registers -> fused rule arithmetic
//                   -> global; with PROMO the partial accumulator of every k-step group is added in fp32 registers
//                   (a whole 128-column accumulator row per thread, hence the register hand-over).
// The wide shallow layers of the backward pass take the vertical-halo variant in tc_conv_vh.cu instead.
#include "epilogue.cuh"
#include "tc_ptx.cuh"
#include <cstdlib>
#include <type_traits>

namespace lrpcap {

namespace {

using namespace tcptx;
constexpr int kATileBytes = 128 * 128;            // 128 rows x 128 B

struct Geom {
  int H, W, TW, TH, TI, tiles_x, tiles_y, cblocks, taps, Nout, n_items, n_tiles_n, group;   // TI: items per tile; group: k-steps per accumulator hand-over
};

// NS = number of bf16 planes per operand: 2 -> products (0,0)(0,1)(1,0); 3 -> additionally (0,2)(2,0)(1,1).
// AP = number of A planes staged: NS, or 1 for the two-product backward (one fp16 message plane x [B_hi ; B_lo]).
// SM2: CTA-pair mode, each CTA stages half of the B rows (the pair's MMA has N = BN).
template <int BN, int NS, int AP = NS, bool SM2 = false>
struct Cfg {
  static constexpr int kBTileBytes = (SM2 ? BN / 2 : BN) * 128;
  static constexpr int kStageBytes = AP * kATileBytes + NS * kBTileBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes;
  static_assert(kStages >= 2, "tile too large for a double-buffered shared-memory ring");
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};
struct Maps {
  CUtensorMap a[3];
  CUtensorMap b[3];
};

// ------------------------------------------------------------------ the kernel
// Persistent, warp-specialised: grid = #SMs, each CTA walks tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...
//   warp 0 (one lane) : TMA producer      -- smem ring runs across tile boundaries, never drains
//   warp 1 (one lane) : MMA issuer        -- accumulators double-buffered in TMEM (2 x BN columns)
//   warps 4..11       : epilogue          -- tcgen05.ld -> fused rule arithmetic -> global, overlapping the next tile's MMAs
constexpr int kEpiWarps = 8;                       // two warps per TMEM lane quarter, interleaved over column chunks
constexpr int kEpiWarp0 = 4;                       // first epilogue warp: warps 0..3 are the producer warpgroup
constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);
constexpr int kProducerRegs = 40, kEpilogueRegs = 232;   // 128 * 48 + 256 * 224 <= 64 K registers


struct TileCoord {
  int item, x0, y0, n0;
};
__device__ __forceinline__ TileCoord tile_coord(const Geom& g, int tile, int BN) {
  TileCoord t;
  const int n_tile = tile % g.n_tiles_n;
  int m = tile / g.n_tiles_n;
  const int tiles_per_item = g.tiles_x * g.tiles_y;   // per group of TI items
  const int grp = m / tiles_per_item;
  t.item = grp * g.TI;                                // first item of the tile
  m -= grp * tiles_per_item;
  t.x0 = (m % g.tiles_x) * g.TW;
  t.y0 = (m / g.tiles_x) * g.TH;
  t.n0 = n_tile * BN;
  return t;
}

// CTA-pair mode: pair-tile q covers two consecutive 128-pixel tiles (2 mp, 2 mp + 1) of one channel tile; CTA `rank` takes
// the pixels of tile 2 mp + rank.
__device__ __forceinline__ TileCoord tile_coord_pair(const Geom& g, int q, int BN, int rank) {
  TileCoord t;
  const int n_tile = q % g.n_tiles_n;
  int m = 2 * (q / g.n_tiles_n) + rank;
  const int tiles_per_item = g.tiles_x * g.tiles_y;   // per group of TI items
  const int grp = m / tiles_per_item;
  t.item = grp * g.TI;                                // first item of the tile
  m -= grp * tiles_per_item;
  t.x0 = (m % g.tiles_x) * g.TW;
  t.y0 = (m / g.tiles_x) * g.TH;
  t.n0 = n_tile * BN;
  return t;
}

// A1: two-product mode. A is ONE fp16 plane (the scaled relevance message), B two fp16 planes [hi ; lo] that sit back to
// back in shared memory. BN <= 128: one MMA with N = 2 BN per K slice (accumulator columns [0, BN) = A*hi, [BN, 2 BN) =
// A*lo, summed by the epilogue); BN = 256: two N = 256 MMAs into the same accumulator.
// F8: fp16 + fp8 mode (epilogue.cuh: StoreH1F8). Plane 0 of A and B are fp16 (message, high weight part), plane 1 are E4M3
// byte planes of the same tile size (128 bytes per row = [top bits | residual] of the message, [low part | high part] of
// the weights): per k-step four kind::f16 MMAs (K = 16) and four kind::f8f6f4 MMAs (K = 32) into one accumulator.
// SM2: CTA-pair mode (cta_group::2, launched as clusters of two): the pair runs one M = 256 x N = BN MMA per K slice. Each
// CTA stages its own 128 pixels of A and HALF of the B rows (the operand bytes through each SM's shared memory drop by a
// third to a half -- the shared-memory pipe, shared by TMA writes and MMA operand reads, is what bounds the one-CTA kernel);
// its TMEM holds the accumulator rows of its pixels, its epilogue warps drain them. TMA loads of both CTAs complete on
// the leader's full barrier, the leader's MMA thread issues for both and its commits arrive on both CTAs' barriers.
template <int BN, int MODE, int NS, bool PROMO, bool F16 = false, bool A1 = false, bool F8 = false, bool SM2 = false>
__global__ void __launch_bounds__(kThreads, 1)
tc_conv_kernel(const __grid_constant__ Maps tm, const Geom g, const EpiDev e, const int total_tiles) {
  static_assert(!A1 || (NS == 2 && F16), "two-product mode: one fp16 A plane x two fp16 B planes");
  static_assert(!F8 || (NS == 2 && F16 && !A1), "fp16 + fp8 mode: one fp16 and one byte plane per operand");
  static_assert(!SM2 || BN >= 128, "CTA-pair mode: N = 128 or 256, half of the B rows per CTA");
  constexpr int AP = A1 ? 1 : NS;
  constexpr bool BCAT = A1 && BN <= 128 && !SM2;   // pairs split the operand rows between the CTAs: no [hi ; lo] concatenation
  constexpr int ACC = BCAT ? 2 * BN : BN;             // TMEM columns per accumulator buffer
  using C = Cfg<BN, NS, AP, SM2>;
  const uint32_t rank = SM2 ? cluster_ctarank() : 0u;   // 0 = leader
  const int tile0 = SM2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;            // first (pair-)tile of this CTA (pair)
  const int tile_step = SM2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto coord = [&](int tile) { return SM2 ? tile_coord_pair(g, tile, BN, (int)rank) : tile_coord(g, tile, BN); };
  using ST = typename std::conditional<F8, StoreH1F8, typename std::conditional<A1, StoreH1, typename std::conditional<NS == 3, StoreSplit3,
                                       typename std::conditional<F16, StoreSplitH, StoreSplit>::type>::type>::type>::type;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;           // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform: the role branches stay on the uniform datapath
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int p = 0; p < NS; ++p) {
      prefetch_tmap(&tm.a[p]);
      prefetch_tmap(&tm.b[p]);
    }
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], SM2 ? 2 * kEpiWarps : kEpiWarps);   // one arrival per epilogue warp (of both CTAs: the leader's counts)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (SM2) cluster_sync_all();   // both CTAs' barriers exist before any remote arrival / completion
  if (warp == 1) {
    if (SM2) tmem_alloc2(tmem_slot, 2 * ACC);
    else tmem_alloc(tmem_slot, 2 * ACC);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_k = g.taps * g.cblocks;
  const uint32_t stage_tx = (uint32_t)AP * (uint32_t)(g.TW * g.TH * g.TI) * 128u + (uint32_t)NS * (uint32_t)C::kBTileBytes;

  if (warp < kEpiWarp0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
  if (warp == 0) {
    {   // the whole warp, converged: the TMA wrappers elect the issuing lane (tc_ptx.cuh)
      // ---------------- TMA producer ----------------
      uint32_t s = 0, ph = 0;   // ring stage and phase bit
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const TileCoord tc = coord(tile);
        for (int kk = 0; kk < num_k; ++kk) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* st = smem + s * C::kStageBytes;
          if (!SM2) mbar_expect_tx_e(&full_bar[s], stage_tx);
          else if (rank == 0) mbar_expect_tx_e(&full_bar[s], 2u * stage_tx);   // the bytes of both CTAs land on the leader's barrier
          const int tap = kk / g.cblocks;
          const int cb = kk - tap * g.cblocks;
          int dy = 0, dx = 0;
          if (g.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
#pragma unroll
          for (int p = 0; p < AP; ++p) {   // byte planes (F8, plane 1) count their innermost coordinate in bytes: 128 per block
            const int c0 = cb * ((F8 && p == 1) ? 128 : kBlockK);
            if (SM2) tma2_load_4d(&tm.a[p], st + p * kATileBytes, &full_bar[s], c0, tc.x0 + dx, tc.y0 + dy, tc.item);
            else tma_load_4d_e(&tm.a[p], st + p * kATileBytes, &full_bar[s], c0, tc.x0 + dx, tc.y0 + dy, tc.item);
          }
#pragma unroll
          for (int p = 0; p < NS; ++p) {
            const int c0 = cb * ((F8 && p == 1) ? 128 : kBlockK);
            uint8_t* dstb = st + AP * kATileBytes + p * C::kBTileBytes;
            if (SM2) tma2_load_2d(&tm.b[p], dstb, &full_bar[s], c0, tap * g.Nout + tc.n0 + (int)rank * (BN / 2));
            else tma_load_2d_e(&tm.b[p], dstb, &full_bar[s], c0, tap * g.Nout + tc.n0);
          }
          if (++s == (uint32_t)C::kStages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ---------------- MMA issuer (CTA-pair mode: the leader's, for both CTAs) ----------------
      // the whole warp runs this loop converged; the MMA / commit wrappers elect the issuing lane (tc_ptx.cuh)
      constexpr uint32_t idesc = make_idesc(SM2 ? 256 : 128, BN, F16);
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
      constexpr uint32_t idesc_cat = make_idesc(128, BCAT ? 2 * BN : BN, F16);
      uint32_t s = 0, ph = 0, tl = 0;   // tl counts accumulator hand-overs: one per tile, or one per `group` k-steps (promotion)
      const int gsz = PROMO ? g.group : num_k;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        for (int kk0 = 0; kk0 < num_k; kk0 += gsz, ++tl) {
          const uint32_t buf = tl & 1u;
          mbar_wait(&tempty_bar[buf], ((tl >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * ACC;
          const int kk1 = (kk0 + gsz < num_k) ? kk0 + gsz : num_k;
          for (int kk = kk0; kk < kk1; ++kk) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t sbase = smem_u32(smem + s * C::kStageBytes);
            uint64_t da[NS], db[NS];
#pragma unroll
            for (int p = 0; p < NS; ++p) {
              da[p] = make_desc_sw128(sbase + (p < AP ? p : 0) * kATileBytes);
              db[p] = make_desc_sw128(sbase + AP * kATileBytes + p * C::kBTileBytes);
            }
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              const uint64_t adv = (uint64_t)(k * 2);  // 16 bf16 = 32 B = 2 x 16 B units inside the swizzle row
              const uint32_t first = (kk != kk0 || k != 0) ? 1u : 0u;   // a fresh accumulator starts every group
              if (F8) {   // fp16 slice k, then the byte planes' 32-byte slice k (together: every product of the k-step once)
                mma16(tmem_d, da[0] + adv, db[0] + adv, idesc, first);
                mma8(tmem_d, da[1] + adv, db[1] + adv, idesc, 1u);
              } else if (A1) {
                if (BCAT) {
                  mma16(tmem_d, da[0] + adv, db[0] + adv, idesc_cat, first);   // [A*hi | A*lo]
                } else {
                  mma16(tmem_d, da[0] + adv, db[1] + adv, idesc, first);
                  mma16(tmem_d, da[0] + adv, db[0] + adv, idesc, 1u);
                }
              } else if (NS == 3) {   // small terms first
                mma16(tmem_d, da[1] + adv, db[1] + adv, idesc, first);
                mma16(tmem_d, da[0] + adv, db[NS - 1] + adv, idesc, 1u);
                mma16(tmem_d, da[NS - 1] + adv, db[0] + adv, idesc, 1u);
                mma16(tmem_d, da[0] + adv, db[1] + adv, idesc, 1u);
                mma16(tmem_d, da[1] + adv, db[0] + adv, idesc, 1u);
                mma16(tmem_d, da[0] + adv, db[0] + adv, idesc, 1u);
              } else {
                mma16(tmem_d, da[0] + adv, db[1] + adv, idesc, first);
                mma16(tmem_d, da[1] + adv, db[0] + adv, idesc, 1u);
                mma16(tmem_d, da[0] + adv, db[0] + adv, idesc, 1u);
              }
            }
            commit(&empty_bar[s]);  // frees this smem stage (in both CTAs of a pair) once the MMAs above have read it
            if (++s == (uint32_t)C::kStages) {
              s = 0;
              ph ^= 1u;
            }
          }
          commit(&tfull_bar[buf]);  // (partial) accumulator complete
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ---------------- epilogue warps (warp w owns TMEM lanes [32 (w % 4), +32)) ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpilogueRegs));
    const int q = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    const int r = q * 32 + lane;
    // accumulator row r = pixel (ti, ty, tx) of the tile: TI items x TH rows x TW pixels (TI > 1 on the 14 / 28 / 56-wide maps,
    // where one item's rows fill only 98 or 112 of the 128 MMA rows: 14 pixels x 1 row x 9 items fill 126)
    const int ti = r / (g.TW * g.TH);
    const int rr = r - ti * (g.TW * g.TH);
    const int ty = rr / g.TW;
    const int tx = rr - ty * g.TW;
    const bool row_ok = r < g.TW * g.TH * g.TI;
    uint32_t tl = 0;
    constexpr int kChunks = BN / 16 / (kEpiWarps / 4);   // 16-column chunks owned by this warp
    auto release_acc = [&](uint32_t buf) {   // this warp has drained accumulator buffer `buf`
      if (lane == 0) {
        if (SM2 && rank != 0) mbar_arrive_cluster(&tempty_bar[buf], 0u);
        else mbar_arrive(&tempty_bar[buf]);
      }
    };
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      const TileCoord tc = coord(tile);
      const int y = tc.y0 + ty, x = tc.x0 + tx;
      const bool valid = row_ok && (y < g.H) && (x < g.W) && (tc.item + ti < g.n_items);
      const int item = (tc.item + ti < g.n_items) ? tc.item + ti : g.n_items - 1;   // clamped: per-item tables are read unguarded
      if constexpr (PROMO) {
        // promotion: sum the per-group partial accumulators in registers (IEEE fp32 adds); tensor-core accumulation
        // drops low bits on every accumulate, which over hundreds of MMAs costs ~1e-5 relative
        float acc[kChunks][16];
#pragma unroll
        for (int ci = 0; ci < kChunks; ++ci)
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[ci][i] = 0.f;
        if (MODE == EPI_BWD && row_ok) {   // multipliers of this tile -> L2 while its MMAs run
          if (y < g.H && x < g.W)
            for (int c = half; c < BN / 16; c += kEpiWarps / 4)
              epi_prefetch_bwd(e, g.H, g.W, g.Nout, item, y, x, tc.n0 + c * 16);
        }
        for (int kk0 = 0; kk0 < num_k; kk0 += g.group, ++tl) {
          const uint32_t buf = tl & 1u;
          mbar_wait(&tfull_bar[buf], (tl >> 1) & 1u);
          tc_fence_after();
          const uint32_t lane_base = tmem_base + buf * ACC + ((uint32_t)(q * 32) << 16);
          // the partial accumulator comes out in batches of up to 64 columns per wait: a wait per 16 columns left the
          // epilogue, not the tensor pipe, pacing the promoted kernels (the load -> wait latency was paid kChunks times)
          constexpr int kBatch = BCAT ? 2 : 4;              // chunks per wait (BCAT reads two column blocks per chunk)
#pragma unroll
          for (int c0 = 0; c0 < kChunks; c0 += kBatch) {
            uint32_t r[kBatch][16], w[BCAT ? kBatch : 1][16];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
              if (c0 + b < kChunks) {
                const uint32_t col = (uint32_t)((half + (c0 + b) * (kEpiWarps / 4)) * 16);
                tmem_ld16_issue(lane_base + col, r[b]);
                if (BCAT) tmem_ld16_issue(lane_base + BN + col, w[b]);
              }
            }
            tmem_ld_wait();
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
              if (c0 + b < kChunks) {
                tmem_ld_fence16(r[b]);
                if (BCAT) tmem_ld_fence16(w[b]);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  float v = __uint_as_float(r[b][i]);
                  if (BCAT) v += __uint_as_float(w[b][i]);
                  acc[c0 + b][i] += v;
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          release_acc(buf);
        }
        // Code size: the tile-end epilogue runs once per tile and warp; unrolled over the chunks it was 110-210 KB of SASS
        // per kernel and came through the instruction cache cold every time (the backward kernels showed the same: see
        // epilogue.cuh). The chunk loops below are real loops: they always read acc[0] and rotate the register array.
        auto rotate_acc = [&]() {
          float t0[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) t0[i] = acc[0][i];
#pragma unroll
          for (int j = 0; j + 1 < kChunks; ++j)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[j][i] = acc[j + 1][i];
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[kChunks - 1][i] = t0[i];
        };
        if constexpr (MODE == EPI_BWD) {
          constexpr int kStep = kEpiWarps / 4;
          auto load_acc = [&](int, float (&v)[16]) {   // called once per chunk, in order (and once more per pass for beta != 0)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = acc[0][i];
            rotate_acc();
          };
          if (e.up == 2)
            epi_bwd_chunks<2, kChunks, ST, false>(e, g.H, g.W, g.Nout, item, y, x, tc.n0 + half * 16, 16 * kStep, valid, load_acc);
          else
            epi_bwd_chunks<1, kChunks, ST, false>(e, g.H, g.W, g.Nout, item, y, x, tc.n0 + half * 16, 16 * kStep, valid, load_acc);
        } else {
#pragma unroll 1
          for (int ci = 0; ci < kChunks; ++ci) {
            if (valid) epi_apply<MODE, 16, ST>(e, g.H, g.W, g.Nout, item, y, x, tc.n0 + (half + ci * (kEpiWarps / 4)) * 16, acc[0]);
            rotate_acc();
          }
        }
        __syncwarp();
      } else {
        const uint32_t buf = tl & 1u;
        if (MODE == EPI_BWD && row_ok) {   // multipliers of the NEXT tile (and of the first one) -> L2, a tile period ahead
          for (int pt = (tl == 0 ? tile : tile + tile_step); pt <= tile + tile_step && pt < total_tiles; pt += tile_step) {
            const TileCoord nt = coord(pt);
            if (nt.y0 + ty < g.H && nt.x0 + tx < g.W && nt.item + ti < g.n_items)
              for (int c = half; c < BN / 16; c += kEpiWarps / 4)
                epi_prefetch_bwd(e, g.H, g.W, g.Nout, nt.item + ti, nt.y0 + ty, nt.x0 + tx, nt.n0 + c * 16);
          }
        }
        mbar_wait(&tfull_bar[buf], (tl >> 1) & 1u);
        tc_fence_after();
        const uint32_t lane_base = tmem_base + buf * ACC + ((uint32_t)(q * 32) << 16);
        auto ld_chunk = [&](uint32_t col, float (&v)[16]) {
          if (BCAT) {
            float w[16];
            tmem_ld16x2(lane_base + col, lane_base + BN + col, v, w);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += w[i];
          } else {
            tmem_ld16(lane_base + col, v);
          }
        };
        if constexpr (MODE == EPI_BWD) {
          constexpr int kStep = kEpiWarps / 4;
          auto load_acc = [&](int c, float (&v)[16]) { ld_chunk((uint32_t)((half + c * kStep) * 16), v); };
          if (e.up == 2)
            epi_bwd_chunks<2, kChunks, ST>(e, g.H, g.W, g.Nout, item, y, x, tc.n0 + half * 16, 16 * kStep, valid, load_acc);
          else
            epi_bwd_chunks<1, kChunks, ST>(e, g.H, g.W, g.Nout, item, y, x, tc.n0 + half * 16, 16 * kStep, valid, load_acc);
        } else {
#pragma unroll 1
          for (int c = half; c < BN / 16; c += kEpiWarps / 4) {
            float v[16];
            __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the predicated stores below
            ld_chunk((uint32_t)(c * 16), v);
            if (valid) epi_apply<MODE, 16, ST>(e, g.H, g.W, g.Nout, item, y, x, tc.n0 + c * 16, v);
          }
        }
        tc_fence_before();
        __syncwarp();
        release_acc(buf);
        ++tl;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (SM2) cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer can still reach its barriers / accumulator
  if (warp == 1) {
    if (SM2) tmem_dealloc2(tmem_base, 2 * ACC);
    else tmem_dealloc(tmem_base, 2 * ACC);
  }
}

}  // namespace

// ------------------------------------------------------------------ host side (shared with tc_conv_vh.cu)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

int make_map_act(CUtensorMap* m, const void* base, int n_items, int H, int W, int C, int TW, int TH, int TI) {
  EncodeTiledFn fn = get_encode_fn();
  LRPCAP_REQUIRE(fn != nullptr, kErrCuda, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_items};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TI};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LRPCAP_REQUIRE(r == CUDA_SUCCESS, kErrCuda, "cuTensorMapEncodeTiled(activation) failed: %d", (int)r);
  return kOk;
}

int make_map_act_u8(CUtensorMap* m, const void* base, int n_items, int H, int W, int Cb, int TW, int TH, int TI) {
  EncodeTiledFn fn = get_encode_fn();
  LRPCAP_REQUIRE(fn != nullptr, kErrCuda, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)Cb, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_items};
  cuuint64_t strides[3] = {(cuuint64_t)Cb, (cuuint64_t)W * Cb, (cuuint64_t)H * W * Cb};
  cuuint32_t box[4] = {128, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TI};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LRPCAP_REQUIRE(r == CUDA_SUCCESS, kErrCuda, "cuTensorMapEncodeTiled(byte activation) failed: %d", (int)r);
  return kOk;
}

int make_map_w_u8(CUtensorMap* m, const void* base, int rows, int Cb, int BN) {
  EncodeTiledFn fn = get_encode_fn();
  LRPCAP_REQUIRE(fn != nullptr, kErrCuda, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)Cb, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Cb};
  cuuint32_t box[2] = {128, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LRPCAP_REQUIRE(r == CUDA_SUCCESS, kErrCuda, "cuTensorMapEncodeTiled(byte weights) failed: %d", (int)r);
  return kOk;
}

int make_map_planar_f32(CUtensorMap* m, const void* base, int n_planes, int H, int W, int box_w, int box_h, int box_c) {
  EncodeTiledFn fn = get_encode_fn();
  LRPCAP_REQUIRE(fn != nullptr, kErrCuda, "cuTensorMapEncodeTiled entry point unavailable");
  LRPCAP_REQUIRE(W % 4 == 0 && box_w % 4 == 0, kErrShape, "planar message: row pitch must be a multiple of 16 bytes");
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_planes};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_c};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LRPCAP_REQUIRE(r == CUDA_SUCCESS, kErrCuda, "cuTensorMapEncodeTiled(planar message) failed: %d", (int)r);
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

namespace {

template <int BN, int MODE, int NS, bool PROMO, bool F16 = false, bool A1 = false, bool F8 = false>
int launch_t(const Maps& tm, const Geom& g, const EpiDev& e, cudaStream_t stream) {
  using C = Cfg<BN, NS, A1 ? 1 : NS>;
  static int smem_state[kMaxDevices] = {};
  LRPCAP_CUDA(ensure_dynamic_smem(tc_conv_kernel<BN, MODE, NS, PROMO, F16, A1, F8>, C::kSmemBytes, smem_state));
  const long long tiles = (long long)ceil_div(g.n_items, g.TI) * g.tiles_x * g.tiles_y * g.n_tiles_n;
  LRPCAP_REQUIRE(tiles > 0 && tiles < (1ll << 31), kErrShape, "tc_conv: %lld tiles out of range", tiles);
  const int num_sms = device_sm_count();
  const unsigned grid = (unsigned)(tiles < num_sms ? tiles : num_sms);   // persistent: one CTA per SM
  tc_conv_kernel<BN, MODE, NS, PROMO, F16, A1, F8><<<grid, kThreads, C::kSmemBytes, stream>>>(tm, g, e, (int)tiles);
  LRPCAP_CUDA(cudaGetLastError());
  return kOk;
}

// CTA-pair launch (BN = 128 / 256): clusters of two CTAs, one pair per two SMs; `tm.b` boxes hold BN / 2 rows.
template <int BN, int MODE, int NS, bool PROMO, bool F16, bool A1, bool F8>
int launch_pair(const Maps& tm, const Geom& g, const EpiDev& e, cudaStream_t stream) {
  using C = Cfg<BN, NS, A1 ? 1 : NS, true>;
  auto kern = tc_conv_kernel<BN, MODE, NS, PROMO, F16, A1, F8, true>;
  static int smem_state[kMaxDevices] = {};
  LRPCAP_CUDA(ensure_dynamic_smem(kern, C::kSmemBytes, smem_state));
  const long long tiles_m = (long long)ceil_div(g.n_items, g.TI) * g.tiles_x * g.tiles_y;
  const long long pair_tiles = tiles_m / 2 * g.n_tiles_n;
  LRPCAP_REQUIRE(tiles_m % 2 == 0 && pair_tiles > 0 && pair_tiles < (1ll << 30), kErrShape, "tc_conv: %lld pixel tiles do not pair up", tiles_m);
  const int pairs_max = device_sm_count() / 2;
  const unsigned grid = 2u * (unsigned)(pair_tiles < pairs_max ? pair_tiles : pairs_max);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LRPCAP_CUDA(cudaLaunchKernelEx(&cfg, kern, tm, g, e, (int)pair_tiles));
  return kOk;
}

}  // namespace

// -1 = not asked yet; LRPCAP_TC_2SM=0 disables the CTA-pair kernels (tc_conv_vh.cu asks too)
bool tc_pair_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* v = std::getenv("LRPCAP_TC_2SM");
    on = (v && v[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

namespace {

// Instantiated combinations: 2 planes -> every epilogue, BN in {64,128,256}, with and without promotion;
// 3 bf16 planes or 2 half planes (forward / raw only, always promoted) -> BN in {64,128}.
template <int BN, bool PROMO>
int launch_mode2(int mode, const Maps& tm, const Geom& g, const EpiDev& e, cudaStream_t stream) {
  switch (mode) {
    case EPI_FWD_TRUE: return launch_t<BN, EPI_FWD_TRUE, 2, PROMO>(tm, g, e, stream);
    case EPI_FWD_ZACT: return launch_t<BN, EPI_FWD_ZACT, 2, PROMO>(tm, g, e, stream);
    case EPI_BWD: return launch_t<BN, EPI_BWD, 2, PROMO>(tm, g, e, stream);
    case EPI_RAW: return launch_t<BN, EPI_RAW, 2, PROMO>(tm, g, e, stream);
  }
  set_last_error("tc_conv: unknown epilogue mode %d", mode);
  return kErrInvalidArg;
}

// CTA-pair dispatch for the N = 128 / 256 backward / raw launches (the caller made `tm.b` with BN / 2-row boxes).
template <int BN>
int launch_pair_mode(int mode, int planes, const Maps& tm, const Geom& g, const EpiDev& e, cudaStream_t stream) {
  const bool promo = g.group > 0;
#define LRPCAP_PAIR(MODE)                                                                                              \
  do {                                                                                                                 \
    if (planes == kPlanesH1F8)                                                                                         \
      return promo ? launch_pair<BN, MODE, 2, true, true, false, true>(tm, g, e, stream) : launch_pair<BN, MODE, 2, false, true, false, true>(tm, g, e, stream); \
    if (planes == kPlanesH1x2)                                                                                         \
      return promo ? launch_pair<BN, MODE, 2, true, true, true, false>(tm, g, e, stream) : launch_pair<BN, MODE, 2, false, true, true, false>(tm, g, e, stream); \
    return promo ? launch_pair<BN, MODE, 2, true, false, false, false>(tm, g, e, stream) : launch_pair<BN, MODE, 2, false, false, false, false>(tm, g, e, stream); \
  } while (0)
  if (mode == EPI_BWD) LRPCAP_PAIR(EPI_BWD);
  if (mode == EPI_RAW) LRPCAP_PAIR(EPI_RAW);
#undef LRPCAP_PAIR
  set_last_error("tc_conv: the CTA-pair kernels cover backward / raw epilogues only (mode %d)", mode);
  return kErrUnsupported;
}

template <int BN>
int launch_mode(int mode, int planes, const Maps& tm, const Geom& g, const EpiDev& e, cudaStream_t stream) {
  if (planes == kPlanesH1F8) {   // fp16 + fp8 backward
    if (mode == EPI_BWD)
      return g.group > 0 ? launch_t<BN, EPI_BWD, 2, true, true, false, true>(tm, g, e, stream)
                         : launch_t<BN, EPI_BWD, 2, false, true, false, true>(tm, g, e, stream);
    if (mode == EPI_RAW)
      return g.group > 0 ? launch_t<BN, EPI_RAW, 2, true, true, false, true>(tm, g, e, stream)
                         : launch_t<BN, EPI_RAW, 2, false, true, false, true>(tm, g, e, stream);
    set_last_error("tc_conv: the fp16 + fp8 mode supports backward / raw epilogues only (mode %d)", mode);
    return kErrUnsupported;
  }
  if (planes == kPlanesH1x2) {   // two-product backward: one fp16 message plane x two fp16 weight planes
    if (mode == EPI_BWD)
      return g.group > 0 ? launch_t<BN, EPI_BWD, 2, true, true, true>(tm, g, e, stream)
                         : launch_t<BN, EPI_BWD, 2, false, true, true>(tm, g, e, stream);
    if (mode == EPI_RAW)
      return g.group > 0 ? launch_t<BN, EPI_RAW, 2, true, true, true>(tm, g, e, stream)
                         : launch_t<BN, EPI_RAW, 2, false, true, true>(tm, g, e, stream);
    set_last_error("tc_conv: the two-product mode supports backward / raw epilogues only (mode %d)", mode);
    return kErrUnsupported;
  }
  if (planes == kPlanesF16x2) {   // two half planes, always promoted
    switch (mode) {
      case EPI_FWD_TRUE: return launch_t<BN, EPI_FWD_TRUE, 2, true, true>(tm, g, e, stream);
      case EPI_FWD_ZACT: return launch_t<BN, EPI_FWD_ZACT, 2, true, true>(tm, g, e, stream);
      case EPI_RAW: return launch_t<BN, EPI_RAW, 2, true, true>(tm, g, e, stream);
    }
    set_last_error("tc_conv: half-plane operands support forward / raw epilogues only (mode %d)", mode);
    return kErrUnsupported;
  }
  if (planes == 3) {
    if constexpr (BN <= 128) {
      switch (mode) {
        case EPI_FWD_TRUE: return launch_t<BN, EPI_FWD_TRUE, 3, true>(tm, g, e, stream);
        case EPI_FWD_ZACT: return launch_t<BN, EPI_FWD_ZACT, 3, true>(tm, g, e, stream);
        case EPI_RAW: return launch_t<BN, EPI_RAW, 3, true>(tm, g, e, stream);
      }
    }
    set_last_error("tc_conv: 3-plane operands support forward / raw epilogues with BN <= 128 only (mode %d, BN %d)", mode, BN);
    return kErrUnsupported;
  }
  return g.group > 0 ? launch_mode2<BN, true>(mode, tm, g, e, stream) : launch_mode2<BN, false>(mode, tm, g, e, stream);
}

}  // namespace

// Tile of <= 128 accumulator rows: TW x TH pixels of TI consecutive items. Maps whose width is a multiple of 16 take 16 x 8;
// other widths up to 128 take the (TW | W, TH | H, TI) with the most MMA rows in use -- the 14 / 28 / 56-wide VGG maps
// tile as 14 x 1 x 9 items = 126 rows instead of 98 / 112 rows of one item (launches on those maps are tensor-bound).
void tc_conv_tile(int H, int W, int n_items, int* TW, int* TH, int* TI) {
  *TI = 1;
  if ((W >= 16 && W % 16 == 0 && H >= 8) || W > 128) {
    *TW = 16;
    *TH = 8;
    return;
  }
  long long best_rows = -1;
  int bw = W, bh = 1, bi = 1;
  for (int tw = W; tw >= 1; --tw) {
    if (W % tw != 0 || tw > 128) continue;
    if (tw < 7 && tw != W) continue;             // short row segments waste TMA requests
    for (int th = 1; th <= H && tw * th <= 128; ++th) {
      if (H % th != 0) continue;
      for (int ti = 1; ti <= n_items && ti <= 32 && tw * th * ti <= 128; ++ti) {
        const long long tiles = (long long)((n_items + ti - 1) / ti) * (H / th) * (W / tw);
        // fewest tiles wins; ties: fewer items per tile, then wider tiles (loops run wide -> narrow, 1 -> many)
        if (best_rows < 0 || tiles < best_rows) {
          best_rows = tiles;
          bw = tw; bh = th; bi = ti;
        }
      }
    }
  }
  *TW = bw;
  *TH = bh;
  *TI = bi;
}

int tc_conv_launch(const TcConvArgs& a, cudaStream_t stream) {
  LRPCAP_REQUIRE(a.A && a.B, kErrInvalidArg, "tc_conv: null operand");
  LRPCAP_REQUIRE(a.C > 0 && a.C % kBlockK == 0, kErrShape, "tc_conv: C=%d must be a positive multiple of 64", a.C);
  LRPCAP_REQUIRE(a.Nout > 0 && a.Nout % 64 == 0, kErrShape, "tc_conv: Nout=%d must be a positive multiple of 64", a.Nout);
  LRPCAP_REQUIRE(a.taps == 9 || a.taps == 1, kErrShape, "tc_conv: taps must be 1 or 9");
  LRPCAP_REQUIRE(a.n_items > 0 && a.H > 0 && a.W > 0, kErrShape, "tc_conv: empty problem");
  LRPCAP_REQUIRE(a.planes == 2 || a.planes == 3 || a.planes == kPlanesF16x2 || a.planes == kPlanesH1x2 || a.planes == kPlanesH1F8,
                 kErrInvalidArg, "tc_conv: planes must be 2, 3, 4 (two half planes), 5 (two-product) or 6 (fp16 + fp8)");
  // N = 256 for the backward modes (half the shared-memory operand bytes per MMA flop of N = 128). The forward modes
  // (3 planes / two half planes, promoted every k-step) stay at N <= 128: with N = 256 the per-k-step accumulator drain
  // (256 columns) outweighed the operand saving (measured: forward 15.0 -> 18.0 ms per 64 images).
  const int BN = ((a.planes == 2 || a.planes == kPlanesH1x2 || a.planes == kPlanesH1F8) && a.Nout % 256 == 0) ? 256
                                                                                        : (a.Nout % 128 == 0 ? 128 : 64);
  if (tc_conv_vh_eligible(a, BN)) return tc_conv_vh_launch(a, BN, stream);   // wide shallow layers: tc_conv_vh.cu
  Geom g;
  g.H = a.H;
  g.W = a.W;
  tc_conv_tile(a.H, a.W, a.n_items, &g.TW, &g.TH, &g.TI);
  g.tiles_x = ceil_div(a.W, g.TW);
  g.tiles_y = ceil_div(a.H, g.TH);
  const long long tiles_m_all = (long long)ceil_div(a.n_items, g.TI) * g.tiles_x * g.tiles_y;   // pixel tiles of the launch
  g.cblocks = a.C / kBlockK;
  g.taps = a.taps;
  g.Nout = a.Nout;
  g.n_items = a.n_items;
  g.n_tiles_n = a.Nout / BN;
  g.group = (a.planes != 2 && a.planes != kPlanesH1x2 && a.planes != kPlanesH1F8) ? (a.promote_every > 0 ? a.promote_every : 1)
                                                                                   : (a.promote_every > 0 ? a.promote_every : 0);

  const __nv_bfloat16* A0 = reinterpret_cast<const __nv_bfloat16*>(a.A);
  const __nv_bfloat16* B0 = reinterpret_cast<const __nv_bfloat16*>(a.B);
  // CTA pairs for the N >= 128 backward launches whose pixel tiles pair up (every full chunk of the encoder chain)
  const bool pair = BN >= 128 && tc_pair_enabled() && (a.planes == 2 || a.planes == kPlanesH1x2 || a.planes == kPlanesH1F8) &&
                    (a.epi.mode == EPI_BWD || a.epi.mode == EPI_RAW) && tiles_m_all % 2 == 0;
  const int brows = pair ? BN / 2 : BN;   // B rows staged per CTA
  Maps tm;
  const int n_planes = a.planes == 3 ? 3 : 2;
  const int a_planes = a.planes == kPlanesH1x2 ? 1 : n_planes;
  if (a.planes == kPlanesH1F8) {   // plane 0: fp16 [.., C]; plane 1: bytes [.., 2 C] right behind it (both 2 B per element)
    const uint8_t* A8 = reinterpret_cast<const uint8_t*>(A0) + a.A_elems * 2;
    const uint8_t* B8 = reinterpret_cast<const uint8_t*>(B0) + a.B_elems * 2;
    LRPCAP_TRY(make_map_act(&tm.a[0], A0, a.n_items, a.H, a.W, a.C, g.TW, g.TH, g.TI));
    LRPCAP_TRY(make_map_act_u8(&tm.a[1], A8, a.n_items, a.H, a.W, 2 * a.C, g.TW, g.TH, g.TI));
    LRPCAP_TRY(make_map_w(&tm.b[0], B0, a.taps * a.Nout, a.C, brows));
    LRPCAP_TRY(make_map_w_u8(&tm.b[1], B8, a.taps * a.Nout, 2 * a.C, brows));
    tm.a[2] = tm.a[0];
    tm.b[2] = tm.b[0];
  } else
  for (int pl = 0; pl < 3; ++pl) {
    const int q = pl < n_planes ? pl : 0;   // unused third slot aliases plane 0
    LRPCAP_TRY(make_map_act(&tm.a[pl], A0 + (size_t)(q < a_planes ? q : 0) * a.A_elems, a.n_items, a.H, a.W, a.C, g.TW, g.TH, g.TI));
    LRPCAP_TRY(make_map_w(&tm.b[pl], B0 + (size_t)q * a.B_elems, a.taps * a.Nout, a.C, brows));
  }

  const EpiParams& p = a.epi;
  EpiDev e;
  LRPCAP_TRY(make_epi_dev(p, &e));
  if (pair) return BN == 256 ? launch_pair_mode<256>(p.mode, a.planes, tm, g, e, stream) : launch_pair_mode<128>(p.mode, a.planes, tm, g, e, stream);

  switch (BN) {
    case 256: return launch_mode<256>(p.mode, a.planes, tm, g, e, stream);
    case 128: return launch_mode<128>(p.mode, a.planes, tm, g, e, stream);
    default: return launch_mode<64>(p.mode, a.planes, tm, g, e, stream);
  }
}

}  // namespace lrpcap
