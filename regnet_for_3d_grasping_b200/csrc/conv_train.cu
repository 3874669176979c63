// Training-path 1x1 convolutions on the tcgen05 engine, in torch's channel-major (B, C, L) layout.
//
// Replaces what torch runs for nn.Conv1d / nn.Conv2d(kernel 1, no bias) of the shared MLPs in train mode
// (/root/reference/multi_model/utils/pn2_utils/nn/modules/conv.py:24-36,64-76 forward, and autograd's dgrad / wgrad):
//
//   fprop   Z[b, co, l]  = sum_ci W[co, ci]  X[b, ci, l]      conv1x1_tc_kernel, A = W   (co x ci, K-major)
//   dgrad   dX[b, ci, l] = sum_co W[co, ci]  dZ[b, co, l]     conv1x1_tc_kernel, A = W^T (ci x co, K-major)
//   wgrad   dW[co, ci]   = sum_b,l dZ[b, co, l] X[b, ci, l]   wgrad_tc_kernel (split-K over (b, l), both operands K-major)
//
// Same numerics as the eval engine (gemm_tc.cu): every fp32 operand is two bf16 planes x = hi + lo and each k-block
// issues hi*hi + lo*hi + hi*lo into one fp32 TMEM accumulator (passes = 3), or hi*hi only (passes = 1, plain bf16).
// Activations stay in the layout torch's modules use, so the position axis l is the CONTIGUOUS one: in fprop / dgrad the
// activation is the N operand of the MMA and is MN-major (tcgen05 takes either major; instruction-descriptor bit 16),
// loaded by TMA as {64 positions x 64 channels} boxes of a 3-D (L, C, B) map; the accumulator holds channels on the
// 128 TMEM lanes and 256 positions on the columns, so every epilogue thread owns one channel: its 32 consecutive
// positions are one 128-byte row of the TMA store, and the per-channel batch-norm moments of the layer (the sums
// BatchNorm needs in train mode) accumulate in that thread's registers -- no separate statistics pass over Z.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "internal.cuh"
#include "tc_ptx.cuh"
#include "train_common.cuh"

namespace regnet {

using namespace tc;

namespace {

constexpr int CT_BM = 128;    // output rows (channels) per tile = TMEM lanes
constexpr int CT_BN = 256;    // positions per tile = TMEM columns of one accumulator
constexpr int CT_BK = 64;     // reduction depth per k-block
constexpr int CT_STAGES = 2;
constexpr int CT_A_BYTES = CT_BM * CT_BK * 2;                      // one plane of the weight tile
constexpr int CT_B_BYTES = CT_BN * CT_BK * 2;                      // one plane of the activation tile
constexpr int CT_STAGE_BYTES = 2 * CT_A_BYTES + 2 * CT_B_BYTES;    // 96 KB
constexpr int CT_OFF_BAR = CT_STAGES * CT_STAGE_BYTES;             // 192 KB of tiles, then barriers
constexpr int CT_OFF_OUT = CT_OFF_BAR + 1024;                      // store staging, 1024-byte aligned
// 4 warps x [32 rows x 128 B].  One buffer per warp, not two: the kernel then leaves 17 KB of the SM's shared memory, enough
// for a co-resident FPS CTA of the next step's prefetched geometry chain (with 226 KB taken, an FPS launch holding 120 SMs
// made every convolution of the backward pass run on the remaining 28)
constexpr int CT_OUT_BYTES = 4 * 4096;
constexpr int CT_SMEM = CT_OFF_OUT + CT_OUT_BYTES + 1024;          // + slack for the manual alignment
constexpr int CT_THREADS = 192;

struct ConvDims {
  int rows;      // output channels (fprop: cout, dgrad: cin)
  int K;         // reduction depth (fprop: cin, dgrad: cout)
  int L;         // positions per batch entry
  int B;
  int passes;    // 3 = split-bf16 (fp32 parity), 1 = plain bf16
  uint32_t lbo, sbo;   // MN-major descriptor offsets of the activation operand (bytes)
};

// dgrad with the reduction pass of the PREVIOUS block's BatchNorm backward fused into the epilogue: the output tile is the
// gradient dy w.r.t. that block's activation y = dropout(relu(bn(z))); with z (the block's convolution output) read next to
// it, every epilogue thread -- one channel -- accumulates sum(g) and sum(g * xhat), g = dy * dropout * [bn(z) > 0], which
// is all the batch-norm backward needs before its apply pass (train_ops.cu bn_bwd_reduce_ex_kernel does the same in a
// separate pass over dy and z).
struct BnReduce {
  const float* z = nullptr;        // (B, rows, L) convolution output of the block whose gradient is being produced
  const float* mean = nullptr;     // (rows) batch statistics / affine map of that block's BatchNorm
  const float* invstd = nullptr;
  const float* scale = nullptr;
  const float* shift = nullptr;
  int relu = 1;
  DropCfg drop = {0, 0, 1.f};
  double* sums = nullptr;          // (rows, 2): sum g, sum g * xhat (zeroed by the launcher)
};

constexpr int MODE_PLAIN = 0, MODE_MOMENTS = 1, MODE_BNREDUCE = 2;

template <int MODE>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv1x1_tc_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                  const __grid_constant__ CUtensorMap map_xhi, const __grid_constant__ CUtensorMap map_xlo,
                  const __grid_constant__ CUtensorMap map_out, ConvDims d, double* __restrict__ moments, BnReduce br) {
  constexpr bool STATS = MODE == MODE_MOMENTS;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_base + CT_OFF_BAR;
  const uint32_t bar_empty = bar_full + 8 * CT_STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * CT_STAGES;
  const uint32_t bar_tempty = bar_tfull + 16;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + CT_OFF_BAR + 8 * (2 * CT_STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_mt = (d.rows + CT_BM - 1) / CT_BM;
  const int n_lt = (d.L + CT_BN - 1) / CT_BN;
  const int64_t n_tiles = (int64_t)n_mt * n_lt * d.B;
  const int n_kblk = (d.K + CT_BK - 1) / CT_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < CT_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_ahi);
    tma_prefetch_desc(&map_xhi);
    tma_prefetch_desc(&map_out);
    if (d.passes == 3) {
      tma_prefetch_desc(&map_alo);
      tma_prefetch_desc(&map_xlo);
    }
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), 2 * CT_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // tile t -> (channel tile fastest, then position tile, then batch entry): CTAs running at the same time work on
  // the same activation tile with different weight rows, so the activation is read from HBM once and from L2 after
  auto tile_coords = [&](int64_t t, int& mt, int& lt, int& b) {
    mt = (int)(t % n_mt);
    const int64_t r = t / n_mt;
    lt = (int)(r % n_lt);
    b = (int)(r / n_lt);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t stage_tx = d.passes == 3 ? CT_STAGE_BYTES : CT_A_BYTES + CT_B_BYTES;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int mt, lt, b;
        tile_coords(t, mt, lt, b);
        for (int kb = 0; kb < n_kblk; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sA = smem_base + stage * CT_STAGE_BYTES;
          const uint32_t sB = sA + 2 * CT_A_BYTES;
          const uint32_t full = bar_full + 8 * stage;
          mbar_arrive_expect_tx(full, stage_tx);
          tma_load_2d(sA, &map_ahi, full, kb * CT_BK, mt * CT_BM);
#pragma unroll
          for (int j = 0; j < CT_BN / 64; ++j)
            tma_load_3d(sB + j * (CT_BK * 128), &map_xhi, full, lt * CT_BN + j * 64, kb * CT_BK, b);
          if (d.passes == 3) {
            tma_load_2d(sA + CT_A_BYTES, &map_alo, full, kb * CT_BK, mt * CT_BM);
#pragma unroll
            for (int j = 0; j < CT_BN / 64; ++j)
              tma_load_3d(sB + CT_B_BYTES + j * (CT_BK * 128), &map_xlo, full, lt * CT_BN + j * 64, kb * CT_BK, b);
          }
          if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(CT_BM, CT_BN) | IDESC_B_MN_MAJOR;
      uint32_t stage = 0, phase = 0, it = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * CT_BN;
        for (int kb = 0; kb < n_kblk; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sA = smem_base + stage * CT_STAGE_BYTES;
          const uint32_t sB = sA + 2 * CT_A_BYTES;
          const uint64_t a_hi = make_sdesc(sA), a_lo = make_sdesc(sA + CT_A_BYTES);
          const uint64_t b_hi = make_sdesc_mn(sB, d.lbo, d.sbo), b_lo = make_sdesc_mn(sB + CT_B_BYTES, d.lbo, d.sbo);
          const int rem = d.K - kb * CT_BK;
          const int ksteps = rem >= CT_BK ? CT_BK / 16 : (rem + 15) / 16;
          // A (K-major, 128-byte rows): +32 bytes per 16-deep k-step; B (MN-major): two 8-row k groups = +2048 bytes
          for (int k = 0; k < ksteps; ++k) umma_f16(d_tmem, a_hi + 2 * k, b_hi + 128 * k, idesc, (kb | k) != 0);
          if (d.passes == 3) {
            for (int k = 0; k < ksteps; ++k) umma_f16(d_tmem, a_lo + 2 * k, b_hi + 128 * k, idesc, 1);
            for (int k = 0; k < ksteps; ++k) umma_f16(d_tmem, a_hi + 2 * k, b_lo + 128 * k, idesc, 1);
          }
          umma_commit(bar_empty + 8 * stage);
          if (kb == n_kblk - 1) umma_commit(bar_tfull + 8 * acc);
          if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ---- epilogue: warp q owns TMEM lanes [32q, 32q + 32) = 32 output channels; thread = one channel ----
    const int q = warp & 3;
    const uint32_t stg = smem_base + CT_OFF_OUT + (warp - 2) * 4096;
    uint32_t it = 0, nstore = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      int mt, lt, b;
      tile_coords(t, mt, lt, b);
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int row0 = mt * CT_BM + q * 32;
      const int l0 = lt * CT_BN;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      float pivot = 0.f, s1 = 0.f, s2 = 0.f;
      // MODE_BNREDUCE: this thread's channel of the previous block
      const int brow = row0 + lane;
      const bool bok = MODE == MODE_BNREDUCE && brow < d.rows;
      float bmu = 0.f, bis = 0.f, bsc = 0.f, bsh = 0.f;
      const float* __restrict__ zrow = nullptr;
      if (bok) {
        bmu = br.mean[brow]; bis = br.invstd[brow]; bsc = br.scale[brow]; bsh = br.shift[brow];
        zrow = br.z + ((int64_t)b * d.rows + brow) * d.L + l0;
      }
#pragma unroll 1
      for (int ch = 0; ch < CT_BN / 32; ++ch) {
        uint32_t v[32];
        const int ncols = min(32, d.L - (l0 + ch * 32));
        float4 zv[8];
        if (MODE == MODE_BNREDUCE) {   // z loads fly while the accumulator chunk is read from tensor memory
#pragma unroll
          for (int j = 0; j < 8; ++j)
            zv[j] = (bok && 4 * j + 4 <= ncols) ? __ldg(reinterpret_cast<const float4*>(zrow + ch * 32) + j)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * CT_BN + ch * 32, v);
        if (MODE == MODE_BNREDUCE) {
          const int64_t e0 = ((int64_t)b * d.rows + brow) * d.L + l0 + ch * 32;   // element index of this chunk (dropout mask)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float g0 = __uint_as_float(v[4 * j]), g1 = __uint_as_float(v[4 * j + 1]), g2 = __uint_as_float(v[4 * j + 2]),
                  g3 = __uint_as_float(v[4 * j + 3]);
            if (!(bok && 4 * j + 4 <= ncols)) g0 = g1 = g2 = g3 = 0.f;    // L is a multiple of 8: whole vectors only
            if (br.drop.thr) {
              const float4 m = drop_mult((uint64_t)(e0 + 4 * j) >> 2, br.drop.seed, br.drop.thr, br.drop.inv_keep);
              g0 *= m.x; g1 *= m.y; g2 *= m.z; g3 *= m.w;
            }
            const float4 z = zv[j];
            if (br.relu) {
              g0 = fmaf(z.x, bsc, bsh) > 0.f ? g0 : 0.f; g1 = fmaf(z.y, bsc, bsh) > 0.f ? g1 : 0.f;
              g2 = fmaf(z.z, bsc, bsh) > 0.f ? g2 : 0.f; g3 = fmaf(z.w, bsc, bsh) > 0.f ? g3 : 0.f;
            }
            s1 += (g0 + g1) + (g2 + g3);
            s2 = fmaf(g0, (z.x - bmu) * bis, fmaf(g1, (z.y - bmu) * bis, fmaf(g2, (z.z - bmu) * bis, fmaf(g3, (z.w - bmu) * bis, s2))));
          }
        }
        if (STATS) {
          // deviations from the first value of the tile row: the sum of squares stays small, no cancellation
          if (ch == 0) pivot = __uint_as_float(v[0]);
          if (ncols == 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float e = __uint_as_float(v[j]) - pivot;
              s1 += e;
              s2 = fmaf(e, e, s2);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float e = j < ncols ? __uint_as_float(v[j]) - pivot : 0.f;
              s1 += e;
              s2 = fmaf(e, e, s2);
            }
          }
        }
        if (ncols > 0 && row0 < d.rows) {   // warp-uniform
          const uint32_t buf = stg;
          if (lane == 0) bulk_wait_read();    // the previous chunk's store has read the buffer (it overlapped the TMEM load)
          __syncwarp();
          const uint32_t rowaddr = buf + (uint32_t)lane * 128u;
          const uint32_t sw = (uint32_t)lane & 7u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts_v4(rowaddr + (((uint32_t)j ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_out, buf, l0 + ch * 32, row0, b);
            bulk_commit();
          }
          ++nstore;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      if (MODE == MODE_BNREDUCE && bok) {
        atomicAdd(br.sums + 2 * brow, (double)s1);
        atomicAdd(br.sums + 2 * brow + 1, (double)s2);
      }
      if (STATS) {
        const int row = row0 + lane;
        const int n = min(CT_BN, d.L - l0);
        if (row < d.rows && n > 0) {
          // tile row -> (count, mean, M2); merged across tiles as raw moments in fp64 (sum, sum of squares)
          const double nd = (double)n;
          const double mean = (double)pivot + (double)s1 / nd;
          const double m2 = fmax((double)s2 - (double)s1 * (double)s1 / nd, 0.0);
          atomicAdd(moments + 2 * row, nd * mean);
          atomicAdd(moments + 2 * row + 1, m2 + nd * mean * mean);
        }
      }
    }
    if (lane == 0) bulk_wait_all();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * CT_BN);
  }
}

// ---- wgrad: dW[co, ci] = sum over (b, l) of dZ[b, co, l] * X[b, ci, l]; both operands K-major (l contiguous) ----------
struct WgradDims {
  int Co, Ci, L, B;
  int n_split;
  int passes;
};

constexpr int WG_OFF_BAR = CT_STAGES * CT_STAGE_BYTES;
constexpr int WG_SMEM = WG_OFF_BAR + 1024 + 1024;

__global__ void __launch_bounds__(CT_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_ghi, const __grid_constant__ CUtensorMap map_glo,
                const __grid_constant__ CUtensorMap map_xhi, const __grid_constant__ CUtensorMap map_xlo, WgradDims d,
                float* __restrict__ partial) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_base + WG_OFF_BAR;
  const uint32_t bar_empty = bar_full + 8 * CT_STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * CT_STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + WG_OFF_BAR + 8 * (2 * CT_STAGES + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_mt = (d.Co + CT_BM - 1) / CT_BM;
  const int n_nt = (d.Ci + CT_BN - 1) / CT_BN;
  const int n_tiles = n_mt * n_nt;
  const int tile = blockIdx.x % n_tiles, split = blockIdx.x / n_tiles;   // tile fastest: co-running CTAs share the k range
  const int mt = tile % n_mt, nt = tile / n_mt;
  const int lblocks = (d.L + CT_BK - 1) / CT_BK;
  const int64_t total_kb = (int64_t)d.B * lblocks;
  const int64_t kb0 = total_kb * split / d.n_split, kb1 = total_kb * (split + 1) / d.n_split;
  const int n_mma = min(CT_BN, (d.Ci - nt * CT_BN + 15) / 16 * 16);   // N of the instruction: this tile's input channels

  if (threadIdx.x == 0) {
    for (int s = 0; s < CT_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_ghi);
    tma_prefetch_desc(&map_xhi);
    if (d.passes == 3) {
      tma_prefetch_desc(&map_glo);
      tma_prefetch_desc(&map_xlo);
    }
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), CT_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t stage_tx = d.passes == 3 ? CT_STAGE_BYTES : CT_A_BYTES + CT_B_BYTES;
      for (int64_t kb = kb0; kb < kb1; ++kb) {
        const int b = (int)(kb / lblocks), l0 = (int)(kb % lblocks) * CT_BK;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        const uint32_t sA = smem_base + stage * CT_STAGE_BYTES;
        const uint32_t sB = sA + 2 * CT_A_BYTES;
        const uint32_t full = bar_full + 8 * stage;
        mbar_arrive_expect_tx(full, stage_tx);
        tma_load_3d(sA, &map_ghi, full, l0, mt * CT_BM, b);
        tma_load_3d(sB, &map_xhi, full, l0, nt * CT_BN, b);
        if (d.passes == 3) {
          tma_load_3d(sA + CT_A_BYTES, &map_glo, full, l0, mt * CT_BM, b);
          tma_load_3d(sB + CT_B_BYTES, &map_xlo, full, l0, nt * CT_BN, b);
        }
        if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(CT_BM, n_mma);
      uint32_t stage = 0, phase = 0;
      for (int64_t kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sA = smem_base + stage * CT_STAGE_BYTES;
        const uint32_t sB = sA + 2 * CT_A_BYTES;
        const uint64_t a_hi = make_sdesc(sA), a_lo = make_sdesc(sA + CT_A_BYTES);
        const uint64_t b_hi = make_sdesc(sB), b_lo = make_sdesc(sB + CT_B_BYTES);
        const int l0 = (int)(kb % lblocks) * CT_BK;
        const int rem = d.L - l0;
        const int ksteps = rem >= CT_BK ? CT_BK / 16 : (rem + 15) / 16;
        for (int k = 0; k < ksteps; ++k) umma_f16(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb != kb0) || k != 0);
        if (d.passes == 3) {
          for (int k = 0; k < ksteps; ++k) umma_f16(tmem_base, a_lo + 2 * k, b_hi + 2 * k, idesc, 1);
          for (int k = 0; k < ksteps; ++k) umma_f16(tmem_base, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
        }
        umma_commit(bar_empty + 8 * stage);
        if (kb == kb1 - 1) umma_commit(bar_tfull);
        if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int co = mt * CT_BM + q * 32 + lane;
    if (kb1 > kb0) {
      mbar_wait(bar_tfull, 0);
      tc_fence_after();
    }
    float* out = partial + ((int64_t)split * d.Co + co) * d.Ci;
#pragma unroll 1
    for (int ch = 0; ch * 32 < n_mma; ++ch) {
      uint32_t v[32];
      if (kb1 > kb0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ch * 32, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      const int ci0 = nt * CT_BN + ch * 32;
      if (co < d.Co) {
        if ((d.Ci & 3) == 0 && ci0 + 32 <= d.Ci) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(out + ci0 + 4 * j) =
                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                            __uint_as_float(v[4 * j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ci0 + j < d.Ci) out[ci0 + j] = __uint_as_float(v[j]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, CT_BN);
  }
}

__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int n_split, int64_t size, float* __restrict__ dW) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < size; i += (int64_t)gridDim.x * 256) {
    float acc = 0.f;
    for (int s = 0; s < n_split; ++s) acc += partial[(int64_t)s * size + i];   // fixed order: deterministic
    dW[i] = acc;
  }
}

// ---- operand preparation ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ x, int64_t n4, int64_t n, __nv_bfloat16* __restrict__ hi,
                    __nv_bfloat16* __restrict__ lo) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    __nv_bfloat16 h[4], l[4];
    split_bf16(v.x, h[0], l[0]);
    split_bf16(v.y, h[1], l[1]);
    split_bf16(v.z, h[2], l[2]);
    split_bf16(v.w, h[3], l[3]);
    reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
    reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
  }
  if (blockIdx.x == 0) {
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += 256) split_bf16(x[i], hi[i], lo[i]);
  }
}

// W (rows, cols) fp32 row-major -> planes (rows, ld_out) [transpose == 0] or the planes of W^T (cols, ld_out)
// [transpose == 1], zero padded to ld_out columns
__global__ void __launch_bounds__(256)
split_weight_kernel(const float* __restrict__ W, int rows, int cols, int ld_out, int transpose,
                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int out_rows = transpose ? cols : rows, out_cols = transpose ? rows : cols;
  const int64_t total = (int64_t)out_rows * ld_out;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int r = (int)(i / ld_out), c = (int)(i % ld_out);
    float v = 0.f;
    if (c < out_cols) v = transpose ? W[(int64_t)c * cols + r] : W[(int64_t)r * cols + c];
    split_bf16(v, hi[i], lo[i]);
  }
}

}  // namespace

}  // namespace regnet

using namespace regnet;

extern "C" {

int regnet_split_planes(const float* x, int64_t n, void* hi, void* lo, void* stream) {
  RN_CHECK_ARG(x && hi && lo, "split_planes: null argument");
  if (n <= 0) return REGNET_OK;
  RN_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(hi) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(lo) & 7) == 0, "split_planes: misaligned buffers");
  const int64_t n4 = n / 4;
  const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>((n4 + 255) / 256, 148LL * 16));
  split_planes_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n4, n, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  RN_LAUNCH_CHECK("split_planes_kernel");
  return REGNET_OK;
}

int regnet_split_weight(const float* W, int rows, int cols, int transpose, int ld_out, void* hi, void* lo, void* stream) {
  RN_CHECK_ARG(W && hi && lo && rows > 0 && cols > 0, "split_weight: null or empty argument");
  RN_CHECK_ARG(ld_out >= (transpose ? rows : cols) && ld_out % 8 == 0, "split_weight: ld_out must be a multiple of 8 covering the row");
  const int64_t total = (int64_t)(transpose ? cols : rows) * ld_out;
  const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, 148LL * 8));
  split_weight_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(W, rows, cols, ld_out, transpose,
                                                                          (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  RN_LAUNCH_CHECK("split_weight_kernel");
  return REGNET_OK;
}

static int conv1x1_launch(const void* x_hi, const void* x_lo, int B, int K, int64_t L, const void* a_hi, const void* a_lo,
                          int rows, int lda, float* out, double* moments, const BnReduce* br, int passes,
                          cudaStream_t stream) {
  RN_CHECK_ARG(x_hi && x_lo && a_hi && a_lo && out, "conv1x1_train: null argument");
  RN_CHECK_ARG(B > 0 && K > 0 && L > 0 && rows > 0, "conv1x1_train: empty problem");
  RN_CHECK_ARG(passes == 1 || passes == 3, "conv1x1_train: passes must be 1 or 3");
  RN_CHECK_ARG(L % 8 == 0, "conv1x1_train: the position extent (%lld) must be a multiple of 8", (long long)L);
  RN_CHECK_ARG(lda % 8 == 0 && lda >= K, "conv1x1_train: weight leading dimension %d must be a multiple of 8 and >= %d", lda, K);
  RN_CHECK_ARG(L < (1LL << 31), "conv1x1_train: too many positions");
  RN_CHECK_ARG(!(moments && br), "conv1x1_train: moments and the fused batch-norm reduction are exclusive");
  if (!gemm_tc_supported()) {
    set_error("conv1x1_train: tensor-map driver entry point unavailable");
    return REGNET_ECUDA;
  }
  CUtensorMap m_ahi, m_alo, m_xhi, m_xlo, m_out;
  RN_TRY(tc_make_map(&m_ahi, a_hi, rows, K, lda, CT_BM, CT_BK, 128));
  RN_TRY(tc_make_map(&m_alo, a_lo, rows, K, lda, CT_BM, CT_BK, 128));
  RN_TRY(tc_make_map3(&m_xhi, x_hi, 2, L, K, B, L, (int64_t)K * L, 64, CT_BK));
  RN_TRY(tc_make_map3(&m_xlo, x_lo, 2, L, K, B, L, (int64_t)K * L, 64, CT_BK));
  RN_TRY(tc_make_map3(&m_out, out, 4, L, rows, B, L, (int64_t)rows * L, 32, 32));
  ConvDims d;
  d.rows = rows; d.K = K; d.L = (int)L; d.B = B; d.passes = passes;
  d.lbo = CT_BK * 128; d.sbo = 1024;
  int dev = 0, sms = 0;
  RN_CUDA(cudaGetDevice(&dev));
  RN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = (int64_t)((rows + CT_BM - 1) / CT_BM) * ((L + CT_BN - 1) / CT_BN) * B;
  const int grid = (int)std::min<int64_t>(n_tiles, sms);
  const BnReduce none;
  if (moments) {
    RN_CUDA(cudaMemsetAsync(moments, 0, sizeof(double) * 2 * rows, stream));
    RN_CUDA(cudaFuncSetAttribute(conv1x1_tc_kernel<MODE_MOMENTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
    conv1x1_tc_kernel<MODE_MOMENTS><<<grid, CT_THREADS, CT_SMEM, stream>>>(m_ahi, m_alo, m_xhi, m_xlo, m_out, d, moments, none);
  } else if (br) {
    RN_CUDA(cudaMemsetAsync(br->sums, 0, sizeof(double) * 2 * rows, stream));
    RN_CUDA(cudaFuncSetAttribute(conv1x1_tc_kernel<MODE_BNREDUCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
    conv1x1_tc_kernel<MODE_BNREDUCE><<<grid, CT_THREADS, CT_SMEM, stream>>>(m_ahi, m_alo, m_xhi, m_xlo, m_out, d, nullptr, *br);
  } else {
    RN_CUDA(cudaFuncSetAttribute(conv1x1_tc_kernel<MODE_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
    conv1x1_tc_kernel<MODE_PLAIN><<<grid, CT_THREADS, CT_SMEM, stream>>>(m_ahi, m_alo, m_xhi, m_xlo, m_out, d, nullptr, none);
  }
  RN_LAUNCH_CHECK("conv1x1_tc_kernel");
  return REGNET_OK;
}

int regnet_conv1x1_train(const void* x_hi, const void* x_lo, int B, int K, int64_t L, const void* a_hi, const void* a_lo,
                         int rows, int lda, float* out, double* moments, int passes, void* stream_) {
  return conv1x1_launch(x_hi, x_lo, B, K, L, a_hi, a_lo, rows, lda, out, moments, nullptr, passes, (cudaStream_t)stream_);
}

int regnet_conv1x1_train_dgrad_bnreduce(const void* g_hi, const void* g_lo, int B, int K, int64_t L, const void* a_hi,
                                        const void* a_lo, int rows, int lda, float* out, const float* z_prev,
                                        const float* mean, const float* invstd, const float* scale, const float* shift,
                                        int relu, float drop_p, uint64_t drop_seed, double* sums, int passes, void* stream_) {
  RN_CHECK_ARG(z_prev && mean && invstd && scale && shift && sums, "conv1x1_train_dgrad_bnreduce: null argument");
  RN_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "conv1x1_train_dgrad_bnreduce: dropout probability %f", drop_p);
  RN_CHECK_ARG((reinterpret_cast<uintptr_t>(z_prev) & 15) == 0, "conv1x1_train_dgrad_bnreduce: z_prev must be 16-byte aligned");
  BnReduce br;
  br.z = z_prev; br.mean = mean; br.invstd = invstd; br.scale = scale; br.shift = shift; br.relu = relu;
  br.drop = make_drop(drop_p, drop_seed); br.sums = sums;
  return conv1x1_launch(g_hi, g_lo, B, K, L, a_hi, a_lo, rows, lda, out, nullptr, &br, passes, (cudaStream_t)stream_);
}

static int wgrad_splits(int Co, int Ci, int64_t total_kb, int sms) {
  const int n_tiles = ((Co + CT_BM - 1) / CT_BM) * ((Ci + CT_BN - 1) / CT_BN);
  int64_t s = std::max(1, sms / n_tiles);
  s = std::min<int64_t>(s, std::max<int64_t>(1, total_kb));
  return (int)s;
}

int64_t regnet_conv1x1_wgrad_workspace_bytes(int B, int Co, int Ci, int64_t L) {
  if (B <= 0 || Co <= 0 || Ci <= 0 || L <= 0) return 0;
  const int64_t total_kb = (int64_t)B * ((L + CT_BK - 1) / CT_BK);
  return (int64_t)wgrad_splits(Co, Ci, total_kb, 148) * Co * Ci * (int64_t)sizeof(float) + 256;
}

int regnet_conv1x1_train_wgrad(const void* g_hi, const void* g_lo, const void* x_hi, const void* x_lo, int B, int Co,
                               int Ci, int64_t L, float* dW, void* workspace, int64_t workspace_bytes, int passes,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RN_CHECK_ARG(g_hi && g_lo && x_hi && x_lo && dW && workspace, "conv1x1_train_wgrad: null argument");
  RN_CHECK_ARG(B > 0 && Co > 0 && Ci > 0 && L > 0, "conv1x1_train_wgrad: empty problem");
  RN_CHECK_ARG(passes == 1 || passes == 3, "conv1x1_train_wgrad: passes must be 1 or 3");
  RN_CHECK_ARG(L % 8 == 0 && L < (1LL << 31), "conv1x1_train_wgrad: the position extent (%lld) must be a multiple of 8", (long long)L);
  RN_CHECK_ARG(workspace_bytes >= regnet_conv1x1_wgrad_workspace_bytes(B, Co, Ci, L), "conv1x1_train_wgrad: workspace too small");
  if (!gemm_tc_supported()) {
    set_error("conv1x1_train_wgrad: tensor-map driver entry point unavailable");
    return REGNET_ECUDA;
  }
  CUtensorMap m_ghi, m_glo, m_xhi, m_xlo;
  RN_TRY(tc_make_map3(&m_ghi, g_hi, 2, L, Co, B, L, (int64_t)Co * L, CT_BK, CT_BM));
  RN_TRY(tc_make_map3(&m_glo, g_lo, 2, L, Co, B, L, (int64_t)Co * L, CT_BK, CT_BM));
  RN_TRY(tc_make_map3(&m_xhi, x_hi, 2, L, Ci, B, L, (int64_t)Ci * L, CT_BK, CT_BN));
  RN_TRY(tc_make_map3(&m_xlo, x_lo, 2, L, Ci, B, L, (int64_t)Ci * L, CT_BK, CT_BN));
  WgradDims d;
  d.Co = Co; d.Ci = Ci; d.L = (int)L; d.B = B; d.passes = passes;
  const int64_t total_kb = (int64_t)B * ((L + CT_BK - 1) / CT_BK);
  d.n_split = wgrad_splits(Co, Ci, total_kb, 148);
  const int n_tiles = ((Co + CT_BM - 1) / CT_BM) * ((Ci + CT_BN - 1) / CT_BN);
  RN_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
  float* partial = reinterpret_cast<float*>(workspace);
  wgrad_tc_kernel<<<n_tiles * d.n_split, CT_THREADS, WG_SMEM, stream>>>(m_ghi, m_glo, m_xhi, m_xlo, d, partial);
  RN_LAUNCH_CHECK("wgrad_tc_kernel");
  const int64_t size = (int64_t)Co * Ci;
  const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>((size + 255) / 256, 148LL * 8));
  wgrad_reduce_kernel<<<(unsigned)blocks, 256, 0, stream>>>(partial, d.n_split, size, dW);
  RN_LAUNCH_CHECK("wgrad_reduce_kernel");
  return REGNET_OK;
}

}  // extern "C"
