// tcgen05 engine for the shared-MLP layer (see gemm.cuh): Y = act(scale * (X W^T) + shift) [+ 64-row max-pool].
//
// fp32 parity (<= 1e-4 rel, BASELINE.json north_star) on bf16 tensor cores: every fp32 operand is carried as two
// bf16 planes x = hi + lo (16 significant bits, fp32 exponent range) and each k-block issues three products into
// the SAME fp32 TMEM accumulator:  hi*hi + lo*hi + hi*lo  (the dropped lo*lo term is ~2^-18 relative).
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer : cp.async.bulk.tensor (128B swizzle) of X_hi/X_lo [128 x 64] and W_hi/W_lo [BN x 64]
//                              tiles into a STAGES-deep shared-memory ring, completion on `full` mbarriers
//   warp 1      MMA issuer   : one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16), 12 per
//                              k-block; tcgen05.commit releases the stage (`empty`) and publishes the accumulator
//                              (`tmem_full`).  Also owns the TMEM allocation (2 x BN fp32 columns, double buffered
//                              so the epilogue of tile i overlaps the main loop of tile i+1)
//   warps 2..5  epilogue     : tcgen05.ld 32x32b.x32 (thread = one position row, 32 channels per load), BN
//                              scale/shift + activation in registers, then either vectorised fp32 / bf16 hi+lo
//                              row stores, or the 64-row max-pool via redux.sync on an order-preserving integer
//                              image of the floats (one instruction per channel per warp)
// Rows (positions) live on the 128 TMEM lanes, channels on the columns, so a tile is 128 positions x BN channels
// and both operands are K-major -- the canonical TMA/UMMA SWIZZLE_128B layout.
#include <cuda.h>

#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace regnet {

using namespace tc;

namespace {

constexpr int BM = 128;      // positions per tile = TMEM lanes
constexpr int BK = 64;       // bf16 per k-block row = 128 bytes = one swizzle span
constexpr int NTHREADS = 192;
constexpr int EPI_THREADS = 128;

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;              // one plane of X
  static constexpr int B_BYTES = BN * BK * 2;              // one plane of W
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = BN == 128 ? 3 : 2;         // shared-memory ring depth (192 KB of tiles)
  static constexpr int OFF_BAR = STAGES * STAGE_BYTES;     // mbarriers + tmem pointer
  static constexpr int OFF_SCALE = OFF_BAR + 256;          // scale[BN], shift[BN]
  static constexpr int OFF_PART = OFF_SCALE + 2 * BN * 4;  // pooled partials [4][BN] as uint32
  static constexpr int OFF_STAGE = (OFF_PART + 4 * BN * 4 + 1023) / 1024 * 1024;  // epilogue store staging:
  static constexpr int STAGE_OUT_BYTES = 4 * 2 * 2048;     //   4 warps x {hi, lo} x [32 rows x 64 B], 64B-swizzled
  static constexpr int SMEM_BYTES = OFF_STAGE + STAGE_OUT_BYTES + 1024;  // + slack for manual 1024B alignment
  static constexpr int TMEM_COLS = 2 * BN;                 // 256 or 512: power of two >= 32
};

template <int ACT>
__device__ __forceinline__ float act_fn(float v) {
  if (ACT == 1) return fmaxf(v, 0.f);
  if (ACT == 2) return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-v)));
  return v;
}

// POOL / ACT are compile-time so that each instance carries only its own epilogue (a runtime switch inlined the
// sigmoid's division subroutine 256 times and the unrolled epilogue overflowed the instruction cache).
// GATHER: the A operand is not a materialised (P, K) matrix.  Row p of the operand is
//   [ table[g(p), 0..C) | xyz_rel[p, 0..3) | 0 ... ],   g(p) = nbr[p] + (p / rows_per_cloud) * n_prev,
// i.e. the grouped [feature | xyz - centroid] row of pn2_utils/modules.py:44-52 -- ball-query grouping fused into the
// GEMM's TMA producer: the C feature columns are fetched straight from the point-major feature table (bf16 hi/lo planes,
// L2-resident) with cp.async.bulk.tensor tile::gather4, four neighbour rows per instruction, one instruction per lane and
// plane; the last k-block (the three xyz_rel columns, K = 16) comes from a small (P, 16) side matrix.
// map_xhi / map_xlo then describe the feature table (box 64 x 1), map_x2hi / map_x2lo the xyz_rel planes (SWIZZLE_32B).
struct GatherA {
  const int32_t* nbr = nullptr;   // (P) neighbour index inside the cloud
  uint32_t rows_per_cloud = 1;    // M * 64 operand rows per cloud (multiple of 4)
  uint32_t n_prev = 0;            // table rows per cloud
  int n_feat_kb = 0;              // C / 64 gathered k-blocks
};

template <int BN, bool POOL, int ACT, bool GATHER>
__global__ void __maxnreg__(88)   // one CTA per SM by shared memory; 88 registers leave room for a co-resident FPS CTA
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_xhi, const __grid_constant__ CUtensorMap map_xlo,
               const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
               const __grid_constant__ CUtensorMap map_ohi, const __grid_constant__ CUtensorMap map_olo,
               const __grid_constant__ CUtensorMap map_x2hi, const __grid_constant__ CUtensorMap map_x2lo, int64_t P,
               int K, int cout, Epilogue ep, GatherA ga) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  // bars[0..STAGES) full, [STAGES..2*STAGES) empty, then tmem_full[2], tmem_empty[2]
  const uint32_t bar_full = smem_base + C::OFF_BAR;
  const uint32_t bar_empty = bar_full + 8 * STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * STAGES;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t bar_sfull = bar_tempty + 16;   // tile-scheduler ring: 4 slots, producer -> {MMA, epilogue}
  const uint32_t bar_sempty = bar_sfull + 32;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + C::OFF_BAR + 8 * (2 * STAGES + 4 + 8));
  volatile int* sched_ring = reinterpret_cast<volatile int*>(tmem_holder + 1);
  float* s_scale = reinterpret_cast<float*>(smem + C::OFF_SCALE);
  float* s_shift = s_scale + BN;
  uint32_t* s_part = reinterpret_cast<uint32_t*>(smem + C::OFF_PART);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ctile = (cout + BN - 1) / BN;
  const int64_t n_ptile = (P + BM - 1) / BM;
  const int64_t n_tiles = n_ptile * n_ctile;
  const int n_kblk = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    if (GATHER) {
      tma_prefetch_desc(&map_x2hi);
      tma_prefetch_desc(&map_x2lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4);
    }
    for (int r = 0; r < 4; ++r) {
      mbar_init(bar_sfull + 8 * r, 1);
      mbar_init(bar_sempty + 8 * r, 5);   // MMA lane + one lane of each epilogue warp
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_xhi);
    tma_prefetch_desc(&map_xlo);
    tma_prefetch_desc(&map_whi);
    tma_prefetch_desc(&map_wlo);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  (void)bars;

  // Tile scheduling.  With ep.tile_counter the CTAs draw tiles from a global counter (a CTA that becomes resident
  // late, or shares its SM with another kernel, simply takes fewer tiles); the producer lane draws, and hands the
  // tile index to the MMA lane and the epilogue warps through a 4-slot shared-memory ring.  Without it: static
  // round-robin over gridDim.x.  -1 ends the loop.
  unsigned int* const tile_counter = ep.tile_counter;
  auto draw_tile = [&](uint32_t it) -> int {        // producer lane only
    int64_t t;
    if (tile_counter) {
      const uint32_t slot = it & 3;
      if (it >= 4) mbar_wait(bar_sempty + 8 * slot, ((it >> 2) - 1) & 1);
      t = (int64_t)atomicAdd(tile_counter, 1u);
      sched_ring[slot] = t < n_tiles ? (int)t : -1;
      mbar_arrive(bar_sfull + 8 * slot);
    } else {
      t = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
    }
    return t < n_tiles ? (int)t : -1;
  };
  auto take_tile = [&](uint32_t it, bool release) -> int {   // MMA lane / every epilogue thread
    if (!tile_counter) {
      const int64_t t = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      return t < n_tiles ? (int)t : -1;
    }
    const uint32_t slot = it & 3;
    mbar_wait(bar_sfull + 8 * slot, (it >> 2) & 1);
    const int t = sched_ring[slot];
    if (release) mbar_arrive(bar_sempty + 8 * slot);
    return t;
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (GATHER) {
      // every lane owns four consecutive rows of the tile = one gather4 per plane and k-block
      uint32_t stage = 0, phase = 0;
      for (uint32_t pit = 0;; ++pit) {
        int tile = 0;
        if (lane == 0) tile = draw_tile(pit);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile < 0) break;
        const int row0 = (tile / n_ctile) * BM;
        const int col0 = (tile % n_ctile) * BN;
        const int64_t r4 = (int64_t)row0 + 4 * lane;
        int g[4] = {0, 0, 0, 0};
        if (r4 + 3 < P) {
          const int4 j = *reinterpret_cast<const int4*>(ga.nbr + r4);
          const int base = (int)((uint32_t)r4 / ga.rows_per_cloud) * (int)ga.n_prev;
          g[0] = j.x + base; g[1] = j.y + base; g[2] = j.z + base; g[3] = j.w + base;
        }
        for (int kb = 0; kb < n_kblk; ++kb) {
          const uint32_t sA = smem_base + stage * C::STAGE_BYTES;
          const uint32_t full = bar_full + 8 * stage;
          if (lane == 0) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            if (kb < ga.n_feat_kb) {
              mbar_arrive_expect_tx(full, C::STAGE_BYTES);
            } else {
              mbar_arrive_expect_tx(full, 2 * BM * 32 + 2 * C::B_BYTES);
              tma_load_2d(sA, &map_x2hi, full, 0, row0);
              tma_load_2d(sA + C::A_BYTES, &map_x2lo, full, 0, row0);
            }
            tma_load_2d(sA + 2 * C::A_BYTES, &map_whi, full, kb * BK, col0);
            tma_load_2d(sA + 2 * C::A_BYTES + C::B_BYTES, &map_wlo, full, kb * BK, col0);
          }
          __syncwarp();   // the stage is free and its barrier armed before any lane's gather can complete on it
          if (kb < ga.n_feat_kb) {
            tma_gather4_2d(sA + lane * 512, &map_xhi, full, kb * BK, g[0], g[1], g[2], g[3]);
            tma_gather4_2d(sA + C::A_BYTES + lane * 512, &map_xlo, full, kb * BK, g[0], g[1], g[2], g[3]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (lane == 0) {
      // Dynamic scheduling draws ONE TILE AHEAD: the atomic for tile pit + 1 is issued before tile pit's loads and its
      // value is only touched (published to the ring) after them, so its ~1 us round trip never stalls the 2-stage ring
      // (a draw in front of the loads cost 1.4 - 2 us per tile, 20 - 35 % of a 4-k-block tile).
      uint32_t stage = 0, phase = 0;
      int tile = draw_tile(0);
      for (uint32_t pit = 0; tile >= 0; ++pit) {
        unsigned fut = 0;
        if (tile_counter) fut = atomicAdd(tile_counter, 1u);
        const int row0 = (tile / n_ctile) * BM;
        const int col0 = (tile % n_ctile) * BN;
        for (int kb = 0; kb < n_kblk; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sA = smem_base + stage * C::STAGE_BYTES;
          const uint32_t full = bar_full + 8 * stage;
          mbar_arrive_expect_tx(full, C::STAGE_BYTES);
          tma_load_2d(sA, &map_xhi, full, kb * BK, row0);
          tma_load_2d(sA + C::A_BYTES, &map_xlo, full, kb * BK, row0);
          tma_load_2d(sA + 2 * C::A_BYTES, &map_whi, full, kb * BK, col0);
          tma_load_2d(sA + 2 * C::A_BYTES + C::B_BYTES, &map_wlo, full, kb * BK, col0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (tile_counter) {
          const uint32_t it = pit + 1, slot = it & 3;
          if (it >= 4) mbar_wait(bar_sempty + 8 * slot, ((it >> 2) - 1) & 1);
          tile = (int64_t)fut < n_tiles ? (int)fut : -1;
          sched_ring[slot] = tile;
          mbar_arrive(bar_sfull + 8 * slot);
        } else {
          const int64_t t = (int64_t)blockIdx.x + (int64_t)(pit + 1) * gridDim.x;
          tile = t < n_tiles ? (int)t : -1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      uint32_t stage = 0, phase = 0;
      for (uint32_t it = 0;; ++it) {
        if (take_tile(it, true) < 0) break;
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < n_kblk; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sA = smem_base + stage * C::STAGE_BYTES;
          const bool xyz_kb = GATHER && kb >= ga.n_feat_kb;   // [128 x 16] SWIZZLE_32B operand of the xyz_rel columns
          const uint64_t a_hi = xyz_kb ? make_sdesc_k16(sA) : make_sdesc(sA);
          const uint64_t a_lo = xyz_kb ? make_sdesc_k16(sA + C::A_BYTES) : make_sdesc(sA + C::A_BYTES);
          const uint64_t b_hi = make_sdesc(sA + 2 * C::A_BYTES), b_lo = make_sdesc(sA + 2 * C::A_BYTES + C::B_BYTES);
          const int rem = K - kb * BK;
          const int ksteps = xyz_kb ? 1 : rem >= BK ? BK / 16 : (rem + 15) / 16;
          for (int k = 0; k < ksteps; ++k) {  // +2 per k-step: 32 bytes >> 4 inside the 128B swizzle span
            umma_f16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb | k) != 0);
          }
          for (int k = 0; k < ksteps; ++k) umma_f16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1);
          for (int k = 0; k < ksteps; ++k) umma_f16(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
          umma_commit(bar_empty + 8 * stage);            // stage reusable once these MMAs have read it
          if (kb == n_kblk - 1) umma_commit(bar_tfull + 8 * acc);  // accumulator complete
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue (warps 2..5; TMEM lane quarter = warp % 4) =================
    const int q = warp & 3;
    const int et = threadIdx.x - 64;  // 0..127
    for (uint32_t it = 0;; ++it) {
      const int tile = take_tile(it, false);
      __syncwarp();
      if (tile_counter && lane == 0) mbar_arrive(bar_sempty + 8 * (it & 3));
      if (tile < 0) break;
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int64_t row0 = (int64_t)(tile / n_ctile) * BM;
      const int col0 = (tile % n_ctile) * BN;
      // stage this channel tile's scale / shift (previous tile's readers are past the trailing barrier)
      const bool do_dot = !POOL && ep.dot_w != nullptr;   // fused 1-output head (Epilogue::dot_*), weights in s_part
      float* s_dotw = reinterpret_cast<float*>(s_part);
      for (int c = et; c < BN; c += EPI_THREADS) {
        const int gc = col0 + c;
        s_scale[c] = (ep.scale && gc < cout) ? ep.scale[gc] : 1.f;
        s_shift[c] = (ep.shift && gc < cout) ? ep.shift[gc] : 0.f;
        if (do_dot) s_dotw[c] = gc < cout ? ep.dot_w[gc] : 0.f;
      }
      float dacc = 0.f;
      epi_bar_sync();
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const int64_t row = row0 + q * 32 + lane;
      const bool row_ok = row < P;
      if (POOL) {
        // 64-row max-pool.  The BN scale of a pooled layer is made non-negative when the weights are uploaded
        // (split_rows_launch row_sign), and every activation is non-decreasing, so act(scale * max(acc) + shift) ==
        // max(act(scale * acc + shift)): the max is taken on the RAW accumulators (4 rows per thread in registers, then
        // 3 shuffle levels -- tc_ptx.cuh), the affine map and the activation are applied once per pooled value below.
        const uint32_t t_lo = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN, t_hi = t_lo + (16u << 16);
        const int my_col = colmax_column(lane);
        uint32_t va[16], vb[16];
        tmem_ld_16x256b_x4_async(t_lo, va);
        tmem_ld_16x256b_x4_async(t_hi, vb);
        tmem_ld_wait();
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          float m[8];
          colmax_rows4(va, vb, m);                       // 4 rows per thread, in registers
          if (ch + 1 < BN / 32) {                        // next chunk's loads fly during the shuffles
            tmem_ld_16x256b_x4_async(t_lo + (ch + 1) * 32, va);
            tmem_ld_16x256b_x4_async(t_hi + (ch + 1) * 32, vb);
          }
          s_part[q * BN + ch * 32 + my_col] = __float_as_uint(colmax_lanes8(m, lane));
          if (ch + 1 < BN / 32) tmem_ld_wait();
        }
      }
#pragma unroll 1
      for (int ch = 0; !POOL && ch < BN / 32; ++ch) {
        // 32 channels per trip, as two 16-column TMEM loads: keeps the epilogue at ~80 registers so that the
        // register-resident FPS kernel of the next step fits on the same SM (profiles/README.md)
        const int c0 = col0 + ch * 32;
        const uint32_t st_hi = smem_base + C::OFF_STAGE + (warp - 2) * 4096, st_lo = st_hi + 2048;
        const bool stage_out = !POOL && ep.out_hi && c0 < cout;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + ch * 32 + half * 16, v);
          float y[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + ch * 32 + half * 16 + 4 * j4);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + ch * 32 + half * 16 + 4 * j4);
            y[4 * j4 + 0] = act_fn<ACT>(fmaf(__uint_as_float(v[4 * j4 + 0]), sc.x, sh.x));
            y[4 * j4 + 1] = act_fn<ACT>(fmaf(__uint_as_float(v[4 * j4 + 1]), sc.y, sh.y));
            y[4 * j4 + 2] = act_fn<ACT>(fmaf(__uint_as_float(v[4 * j4 + 2]), sc.z, sh.z));
            y[4 * j4 + 3] = act_fn<ACT>(fmaf(__uint_as_float(v[4 * j4 + 3]), sc.w, sh.w));
          }
          if (do_dot) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dacc = fmaf(y[j], s_dotw[ch * 32 + half * 16 + j], dacc);
          }
          {
            if (stage_out) {
              // bf16 hi/lo planes: stage this warp's [32 rows x 32 ch] block in shared memory (64B-swizzled rows,
              // conflict-free STS.128) and let TMA write full lines; rows >= P / channels >= cout are clipped by TMA
              uint32_t hi[8], lo[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) split_bf16_pair(y[2 * j], y[2 * j + 1], hi[j], lo[j]);
              if (half == 0) {
                if (lane == 0) bulk_wait_read();  // the previous block's TMA stores have finished reading the staging
                __syncwarp();
              }
              const uint32_t swz = (uint32_t)(lane >> 1) & 3u;
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint32_t off = (uint32_t)lane * 64u + (((uint32_t)(half * 2 + j) ^ swz) << 4);
                sts_v4(st_hi + off, hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                sts_v4(st_lo + off, lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
              }
              if (half == 1) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&map_ohi, st_hi, c0, (int)(row0 + q * 32));
                  tma_store_2d(&map_olo, st_lo, c0, (int)(row0 + q * 32));
                  bulk_commit();
                }
              }
            }
            if (ep.out_f32 && row_ok) {
              const int ch0 = c0 + half * 16;
              if (ch0 + 16 <= cout) {
                float4* dst = reinterpret_cast<float4*>(ep.out_f32 + row * ep.ld_f32 + ch0);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (ch0 + j < cout) ep.out_f32[row * ep.ld_f32 + ch0 + j] = y[j];
              }
            }
          }
        }
      }
      if (do_dot && row_ok) {
        const float v = fmaf(dacc, ep.dot_scale ? ep.dot_scale[0] : 1.f, ep.dot_shift ? ep.dot_shift[0] : 0.f);
        ep.dot_out[row] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-v)));
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      if (POOL) {
        epi_bar_sync();
        for (int e = et; e < 2 * BN; e += EPI_THREADS) {
          const int g = e / BN, c = e - g * BN;
          const int64_t grow = row0 / 64 + g;
          if (grow * 64 < P && col0 + c < cout) {
            const float m = fmaxf(__uint_as_float(s_part[(2 * g) * BN + c]), __uint_as_float(s_part[(2 * g + 1) * BN + c]));
            const float y = act_fn<ACT>(fmaf(m, s_scale[c], s_shift[c]));
            ep.out_f32[grow * ep.ld_f32 + col0 + c] = y;
            if (ep.pool_hi) {   // the same values as bf16 hi/lo planes: the gather table of the next level's first layer
              __nv_bfloat16 h, l;
              split_bf16(y, h, l);
              ep.pool_hi[grow * ep.ld_pool + col0 + c] = h;
              ep.pool_lo[grow * ep.ld_pool + col0 + c] = l;
            }
          }
        }
      }
      epi_bar_sync();  // s_scale / s_part free for the next tile
    }
    if (lane == 0) bulk_wait_all();  // outstanding TMA stores complete before the CTA (and its smem) goes away
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
}  // namespace (kernels)

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  }
  return fn;
}

}  // namespace

int tc_make_map(CUtensorMap* map, const void* base, int64_t rows, int cols, int ld, int box_rows, int box_cols,
                int swizzle_bytes) {
  const CUtensorMapSwizzle swizzle = swizzle_bytes == 32   ? CU_TENSOR_MAP_SWIZZLE_32B
                                     : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                           : CU_TENSOR_MAP_SWIZZLE_128B;
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("gemm_tc: cuTensorMapEncodeTiled is not available from this driver");
    return REGNET_ECUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_tc: cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%d ld=%d)", (int)r,
              (long long)rows, cols, ld);
    return REGNET_ECUDA;
  }
  return REGNET_OK;
}

int tc_driver_ok(void) { return encode_fn() != nullptr; }

// 3-D map over a (d2, d1, d0) array with d0 contiguous (torch's (B, C, L) tensors: d0 = L, d1 = C, d2 = B);
// box = box0 x box1 x 1, SWIZZLE_128B when box0 * elem_bytes == 128, otherwise no swizzle.
int tc_make_map3(CUtensorMap* map, const void* base, int elem_bytes, int64_t d0, int64_t d1, int64_t d2, int64_t stride1,
                 int64_t stride2, int box0, int box1) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("tc_make_map3: cuTensorMapEncodeTiled is not available from this driver");
    return REGNET_ECUDA;
  }
  RN_CHECK_ARG(elem_bytes == 2 || elem_bytes == 4, "tc_make_map3: element size %d", elem_bytes);
  RN_CHECK_ARG((stride1 * elem_bytes) % 16 == 0 && (stride2 * elem_bytes) % 16 == 0 &&
                   (reinterpret_cast<uintptr_t>(base) & 15) == 0,
               "tc_make_map3: rows must start on 16-byte boundaries (innermost extent %lld x %d bytes)", (long long)stride1,
               elem_bytes);
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)stride1 * elem_bytes, (cuuint64_t)stride2 * elem_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle swz = box0 * elem_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tc_make_map3: cuTensorMapEncodeTiled failed with CUresult %d (dims %lld x %lld x %lld)", (int)r, (long long)d0,
              (long long)d1, (long long)d2);
    return REGNET_ECUDA;
  }
  return REGNET_OK;
}

namespace {


struct Maps {
  CUtensorMap xh, xl, wh, wl, oh, ol, x2h, x2l;
};

template <int BN, bool POOL, int ACT, bool GATHER>
int launch(const Maps& m, int64_t P, int K, int cout, const Epilogue& ep, const GatherA& ga, cudaStream_t stream) {
  int dev = 0, sms = 0;
  RN_CUDA(cudaGetDevice(&dev));
  RN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  RN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, POOL, ACT, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Cfg<BN>::SMEM_BYTES));
  const int64_t n_tiles = ((P + BM - 1) / BM) * ((cout + BN - 1) / BN);
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  gemm_tc_kernel<BN, POOL, ACT, GATHER><<<grid, NTHREADS, Cfg<BN>::SMEM_BYTES, stream>>>(
      m.xh, m.xl, m.wh, m.wl, m.oh, m.ol, m.x2h, m.x2l, P, K, cout, ep, ga);
  RN_LAUNCH_CHECK("gemm_tc_kernel");
  return REGNET_OK;
}

template <int BN, bool GATHER>
int launch_bn(const Maps& m, int64_t P, int K, int cout, const Epilogue& ep, const GatherA& ga, cudaStream_t stream) {
  if (GATHER) {   // only the instances the SA first layers use: ReLU, no pooling
    if (!ep.pool && ep.act == 1) return launch<BN, false, 1, GATHER>(m, P, K, cout, ep, ga, stream);
    set_error("gemm_tc: the gathered-operand variant supports act = relu without pooling only");
    return REGNET_EINVAL;
  }
  if (ep.pool) {
    if (ep.act == 1) return launch<BN, true, 1, false>(m, P, K, cout, ep, ga, stream);
    if (ep.act == 0) return launch<BN, true, 0, false>(m, P, K, cout, ep, ga, stream);
    return launch<BN, true, 2, false>(m, P, K, cout, ep, ga, stream);
  }
  if (ep.act == 1) return launch<BN, false, 1, false>(m, P, K, cout, ep, ga, stream);
  if (ep.act == 0) return launch<BN, false, 0, false>(m, P, K, cout, ep, ga, stream);
  return launch<BN, false, 2, false>(m, P, K, cout, ep, ga, stream);
}

}  // namespace

int gemm_tc_supported(void) { return tc_driver_ok(); }

static int check_epilogue(const Epilogue& ep, int64_t P) {
  RN_CHECK_ARG(ep.pool == 0 || (ep.pool == 64 && P % 64 == 0 && ep.out_f32 && !ep.out_hi),
               "gemm_tc: pooled epilogue needs pool == 64, P %% 64 == 0 and an fp32 output");
  RN_CHECK_ARG(!ep.out_f32 || ep.pool || ep.ld_f32 % 4 == 0, "gemm_tc: fp32 output leading dimension must be a multiple of 4");
  RN_CHECK_ARG(!ep.out_hi || ep.ld_split % 8 == 0, "gemm_tc: split output leading dimension must be a multiple of 8");
  RN_CHECK_ARG(P < (1LL << 31), "gemm_tc: too many rows");
  RN_CHECK_ARG(ep.act >= 0 && ep.act <= 2, "gemm_tc: unknown activation %d", ep.act);
  RN_CHECK_ARG(!ep.dot_w || (ep.dot_out && !ep.pool), "gemm_tc: the fused 1-output head needs dot_out and no pooling");
  return REGNET_OK;
}

int gemm_tc_launch(const __nv_bfloat16* Xhi, const __nv_bfloat16* Xlo, int ldx, const __nv_bfloat16* Whi,
                   const __nv_bfloat16* Wlo, int ldw, int64_t P, int K, int cout, const Epilogue& ep,
                   cudaStream_t stream) {
  RN_CHECK_ARG(ldx % 8 == 0 && ldw % 8 == 0 && ldx >= K && ldw >= K,
               "gemm_tc: leading dimensions must be multiples of 8 and >= K (K=%d ldx=%d ldw=%d)", K, ldx, ldw);
  RN_TRY(check_epilogue(ep, P));
  RN_CHECK_ARG(!ep.dot_w || cout <= 128, "gemm_tc: the fused 1-output head needs cout <= 128 (one column tile)");
  if (P == 0) return REGNET_OK;
  const int bn = cout > 128 ? 256 : 128;
  Maps m;
  RN_TRY(tc_make_map(&m.xh, Xhi, P, K, ldx, BM, BK, 128));
  RN_TRY(tc_make_map(&m.xl, Xlo, P, K, ldx, BM, BK, 128));
  RN_TRY(tc_make_map(&m.wh, Whi, cout, K, ldw, bn, BK, 128));
  RN_TRY(tc_make_map(&m.wl, Wlo, cout, K, ldw, bn, BK, 128));
  m.oh = m.xh; m.ol = m.xl;  // placeholders when there is no split output
  if (ep.out_hi) {
    RN_TRY(tc_make_map(&m.oh, ep.out_hi, P, cout, ep.ld_split, 32, 32, 64));
    RN_TRY(tc_make_map(&m.ol, ep.out_lo, P, cout, ep.ld_split, 32, 32, 64));
  }
  m.x2h = m.xh; m.x2l = m.xl;
  const GatherA ga;
  if (bn == 256) return launch_bn<256, false>(m, P, K, cout, ep, ga, stream);
  return launch_bn<128, false>(m, P, K, cout, ep, ga, stream);
}

int gemm_tc_gather_launch(const __nv_bfloat16* Thi, const __nv_bfloat16* Tlo, int64_t table_rows, int C, int ldt,
                          const int32_t* nbr, int rows_per_cloud, int n_prev, const __nv_bfloat16* Zhi,
                          const __nv_bfloat16* Zlo, const __nv_bfloat16* Whi, const __nv_bfloat16* Wlo, int ldw, int64_t P,
                          int cout, const Epilogue& ep, cudaStream_t stream) {
  const int K = C + 3;
  RN_CHECK_ARG(C > 0 && C % BK == 0 && ldt % 8 == 0 && ldt >= C, "gemm_tc gather: the table needs C %% 64 == 0 columns (C=%d ldt=%d)", C, ldt);
  RN_CHECK_ARG(ldw % 8 == 0 && ldw >= K, "gemm_tc gather: weight leading dimension (%d) must be a multiple of 8 and >= C + 3", ldw);
  RN_CHECK_ARG(rows_per_cloud > 0 && rows_per_cloud % 4 == 0 && n_prev > 0 && table_rows < (1LL << 31),
               "gemm_tc gather: bad grouping geometry");
  RN_CHECK_ARG((reinterpret_cast<uintptr_t>(nbr) & 15) == 0, "gemm_tc gather: the neighbour index must be 16-byte aligned");
  RN_TRY(check_epilogue(ep, P));
  if (P == 0) return REGNET_OK;
  const int bn = cout > 128 ? 256 : 128;
  Maps m;
  RN_TRY(tc_make_map(&m.xh, Thi, table_rows, C, ldt, 1, BK, 128));     // gather4: box = 64 columns x ONE row
  RN_TRY(tc_make_map(&m.xl, Tlo, table_rows, C, ldt, 1, BK, 128));
  RN_TRY(tc_make_map(&m.x2h, Zhi, P, 16, 16, BM, 16, 32));            // xyz_rel planes (P, 16), SWIZZLE_32B
  RN_TRY(tc_make_map(&m.x2l, Zlo, P, 16, 16, BM, 16, 32));
  RN_TRY(tc_make_map(&m.wh, Whi, cout, K, ldw, bn, BK, 128));
  RN_TRY(tc_make_map(&m.wl, Wlo, cout, K, ldw, bn, BK, 128));
  m.oh = m.wh; m.ol = m.wl;
  if (ep.out_hi) {
    RN_TRY(tc_make_map(&m.oh, ep.out_hi, P, cout, ep.ld_split, 32, 32, 64));
    RN_TRY(tc_make_map(&m.ol, ep.out_lo, P, cout, ep.ld_split, 32, 32, 64));
  }
  GatherA ga;
  ga.nbr = nbr; ga.rows_per_cloud = (uint32_t)rows_per_cloud; ga.n_prev = (uint32_t)n_prev; ga.n_feat_kb = C / BK;
  if (bn == 256) return launch_bn<256, true>(m, P, K, cout, ep, ga, stream);
  return launch_bn<128, true>(m, P, K, cout, ep, ga, stream);
}

}  // namespace regnet
