// Second shared-MLP layer of set-abstraction levels 1 and 2 with its A operand PRODUCED INSIDE THE KERNEL.
//
// Reference chain (pn2_utils/modules.py:44-52 grouping + centring + concat, nn/modules/conv.py:64-76 first block):
//   a[p, :] = relu(bn0(W0 [feature[g(p)] | xyz[g(p)] - centre(p)]))          p = grouped position, g(p) its source point
// Linear-first (scorenet.cu): the feature part W_f f is a GEMM over the SOURCE points; folding the xyz term the same way,
//   Z'[j, :] = scale0 * (W_f f_j + W_x xyz_j)      per source point   (fold_kernel below, fp32 FMA on the GEMM's output)
//   T[m, :]  = shift0 - scale0 * (W_x centre_m)    per centroid
//   a[p, :]  = relu(Z'[g(p), :] + T[p / 64, :])
// the grouped activation is ONE ADD per element away from two small L2-resident tables.  Materialising it cost a 1 GB write
// and a 1 GB read at level 1 (sa_operand.1 0.26 ms + an HBM-bound layer at 48 % tensor); here eight producer warps gather
// the rows, add, split to bf16 hi/lo and write the SWIZZLE_128B K-major operand tile straight into shared memory, where
// the tensor pipe consumes it.
//
// Roles (448 threads, one CTA per SM, tiles of 128 positions x 256 output channels drawn from a global counter):
//   warp 0        scheduler (draws tiles one ahead, 4-slot ring) + TMA loads of the W k-blocks (hi | lo, 64 KB per slot)
//   warp 1        MMA issuer: 12 tcgen05.mma (split-bf16 triple) per k-block into one of two TMEM accumulators
//   warps 2..5    epilogue: tcgen05.ld, BN + ReLU, bf16 hi/lo, 64B-swizzled staging, TMA store (as gemm_tc.cu)
//   warps 6..13   producers: thread = (row r + 32 i, 8-channel chunk c); the 32-byte gathers of the NEXT k-block (also across
//                 a tile boundary: row indices and T rows of the next tile are staged one tile ahead) are in flight in
//                 registers while the current one is converted, so the L2 latency never meets the slot wait
// Shared memory: 2 slots x (A 32 KB + W 64 KB) + 16 KB store staging + T / row staging = 220 KB (an FPS CTA still fits).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "gemm.cuh"
#include "internal.cuh"
#include "tc_ptx.cuh"

namespace regnet {

using namespace tc;

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int KMAX = 512;
constexpr int A_PLANE = BM * BK * 2;                  // 16 KB
constexpr int W_PLANE = BN * BK * 2;                  // 32 KB
constexpr int NSLOT = 2;
constexpr int OFF_A = 0;                              // [slot][hi | lo]
constexpr int OFF_W = OFF_A + NSLOT * 2 * A_PLANE;    // [slot][hi | lo]
constexpr int OFF_OUT = OFF_W + NSLOT * 2 * W_PLANE;  // 4 epilogue warps x {hi, lo} x [32 rows x 64 B]
constexpr int OFF_TV = OFF_OUT + 4 * 4096;            // [2 tiles][2 centroids][KMAX] fp32
constexpr int OFF_ROW = OFF_TV + 2 * 2 * KMAX * 4;    // [2 tiles][128] source row
constexpr int OFF_SCALE = OFF_ROW + 2 * BM * 4;       // scale[256], shift[256]
constexpr int OFF_BAR = OFF_SCALE + 2 * BN * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;      // + slack for the manual 1024-byte alignment
constexpr int NTHREADS = 448;
constexpr int PROD_THREADS = 256, PROD_WARPS = 8;
static_assert(SMEM_BYTES <= 227 * 1024, "gemm_fused_a: shared memory budget");

__device__ __forceinline__ void prod_bar_sync() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

struct FusedAArgs {
  const float* Z; int ldz;            // (table rows, K): scale0 * (W_f f + W_x xyz) per source point
  const float* T;                     // (P / 64, K): shift0 - scale0 * W_x centre per centroid
  const int32_t* nbr;                 // (P) neighbour index inside the cloud
  uint32_t rows_per_cloud, n_prev;
  uint32_t P; int K, cout;
  const float* scale; const float* shift;
  unsigned int* tile_counter;
  int variant;                        // bring-up experiments (REGNET_FUSED_A_VARIANT): 1 = no gathers, 2 = no conversion, 4 = no output stores
};

__global__ void __maxnreg__(104)
gemm_fused_a_kernel(const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
                    const __grid_constant__ CUtensorMap map_ohi, const __grid_constant__ CUtensorMap map_olo, FusedAArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_base + OFF_BAR;          // [2]: 8 producer warps + the W loader's expect_tx
  const uint32_t bar_empty = bar_full + 16;               // [2]: tcgen05.commit
  const uint32_t bar_tfull = bar_empty + 16;              // [2]
  const uint32_t bar_tempty = bar_tfull + 16;             // [2]: 4 epilogue warps
  const uint32_t bar_sfull = bar_tempty + 16;             // [4]
  const uint32_t bar_sempty = bar_sfull + 32;             // [4]: MMA lane + 4 epilogue warps + 8 producer warps
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 128);
  volatile int* sched_ring = reinterpret_cast<volatile int*>(tmem_holder + 1);
  float* s_scale = reinterpret_cast<float*>(smem + OFF_SCALE);
  float* s_shift = s_scale + BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ctile = (a.cout + BN - 1) / BN;
  const int n_tiles = (int)(a.P / BM) * n_ctile;
  const int n_kblk = a.K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(bar_full + 8 * s, PROD_WARPS + 1);
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 4);
    }
    for (int r = 0; r < 4; ++r) {
      mbar_init(bar_sfull + 8 * r, 1);
      mbar_init(bar_sempty + 8 * r, 1 + 4 + PROD_WARPS);
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_whi);
    tma_prefetch_desc(&map_wlo);
    tma_prefetch_desc(&map_ohi);
    tma_prefetch_desc(&map_olo);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // The scheduler lane draws tiles from the global counter TWO tiles ahead: the atomic's round trip (~1 us) is in flight
  // while the lane issues a whole tile of W loads, and its value is published (4-slot ring) afterwards.  A draw that sat in
  // front of the loads left the 2-slot W ring dry once per tile (measured: 2 us per tile, 35 % of a 4-k-block tile).
  auto publish_tile = [&](uint32_t it, unsigned t) -> int {          // scheduler lane
    const uint32_t slot = it & 3;
    if (it >= 4) mbar_wait(bar_sempty + 8 * slot, ((it >> 2) - 1) & 1);
    const int tile = t < (unsigned)n_tiles ? (int)t : -1;
    sched_ring[slot] = tile;
    mbar_arrive(bar_sfull + 8 * slot);
    return tile;
  };
  auto take_tile = [&](uint32_t it) -> int {           // one lane per consumer warp; releases the slot at once
    const uint32_t slot = it & 3;
    mbar_wait(bar_sfull + 8 * slot, (it >> 2) & 1);
    const int t = sched_ring[slot];
    mbar_arrive(bar_sempty + 8 * slot);
    return t;
  };

  if (warp == 0) {
    // ================= scheduler + W loader =================
    if (lane == 0) {
      uint32_t job = 0;
      int tile = publish_tile(0, atomicAdd(a.tile_counter, 1u));
      int next = publish_tile(1, atomicAdd(a.tile_counter, 1u));   // the producers stage a tile's rows one tile ahead
      for (uint32_t pit = 0; tile >= 0; ++pit) {
        const unsigned fut = atomicAdd(a.tile_counter, 1u);         // tile pit + 2: consumed after this tile's loads
        const int col0 = (tile % n_ctile) * BN;
        for (int kb = 0; kb < n_kblk; ++kb, ++job) {
          const uint32_t slot = job & 1;
          mbar_wait(bar_empty + 8 * slot, ((job >> 1) & 1) ^ 1);
          const uint32_t sW = smem_base + OFF_W + slot * 2 * W_PLANE;
          mbar_arrive_expect_tx(bar_full + 8 * slot, 2 * W_PLANE);
          tma_load_2d(sW, &map_whi, bar_full + 8 * slot, kb * BK, col0);
          tma_load_2d(sW + W_PLANE, &map_wlo, bar_full + 8 * slot, kb * BK, col0);
        }
        tile = next;
        if (tile >= 0) next = publish_tile(pit + 2, fut);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      uint32_t job = 0;
      for (uint32_t it = 0;; ++it) {
        if (take_tile(it) < 0) break;
        const uint32_t acc = it & 1;
        mbar_wait(bar_tempty + 8 * acc, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < n_kblk; ++kb, ++job) {
          const uint32_t slot = job & 1;
          mbar_wait(bar_full + 8 * slot, (job >> 1) & 1);
          tc_fence_after();
          const uint32_t sA = smem_base + OFF_A + slot * 2 * A_PLANE, sW = smem_base + OFF_W + slot * 2 * W_PLANE;
          const uint64_t a_hi = make_sdesc(sA), a_lo = make_sdesc(sA + A_PLANE);
          const uint64_t b_hi = make_sdesc(sW), b_lo = make_sdesc(sW + W_PLANE);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb | k) != 0);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
          umma_commit(bar_empty + 8 * slot);
          if (kb == n_kblk - 1) umma_commit(bar_tfull + 8 * acc);
        }
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ================= epilogue (TMEM lane quarter = warp % 4) =================
    const int q = warp & 3;
    const int et = threadIdx.x - 64;
    for (uint32_t it = 0;; ++it) {
      int tile = 0;
      if (lane == 0) tile = take_tile(it);
      tile = __shfl_sync(FULL, tile, 0);
      if (tile < 0) break;
      const uint32_t acc = it & 1;
      const int row0 = (tile / n_ctile) * BM;
      const int col0 = (tile % n_ctile) * BN;
      for (int c = et; c < BN; c += 128) {
        const int gc = col0 + c;
        s_scale[c] = (a.scale && gc < a.cout) ? a.scale[gc] : 1.f;
        s_shift[c] = (a.shift && gc < a.cout) ? a.shift[gc] : 0.f;
      }
      epi_bar_sync();
      mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      const uint32_t st_hi = smem_base + OFF_OUT + (warp - 2) * 4096, st_lo = st_hi + 2048;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        const int c0 = col0 + ch * 32;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + ch * 32 + half * 16, v);
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + ch * 32 + half * 16 + 4 * j4);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + ch * 32 + half * 16 + 4 * j4);
            const float y0 = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 0]), sc.x, sh.x), 0.f);
            const float y1 = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 1]), sc.y, sh.y), 0.f);
            const float y2 = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 2]), sc.z, sh.z), 0.f);
            const float y3 = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 3]), sc.w, sh.w), 0.f);
            split_bf16_pair(y0, y1, hi[2 * j4], lo[2 * j4]);
            split_bf16_pair(y2, y3, hi[2 * j4 + 1], lo[2 * j4 + 1]);
          }
          if (half == 0) {
            if (lane == 0) bulk_wait_read();
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
            if (lane == 0 && c0 < a.cout && !(a.variant & 4)) {
              tma_store_2d(&map_ohi, st_hi, c0, row0 + q * 32);
              tma_store_2d(&map_olo, st_lo, c0, row0 + q * 32);
              bulk_commit();
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      epi_bar_sync();   // s_scale free for the next tile
    }
    if (lane == 0) bulk_wait_all();
    __syncwarp();
  } else {
    // ================= A producers =================
    const int pt = threadIdx.x - 192;            // 0..255
    const int c = pt & 7;                        // 8-channel chunk inside the 64-channel k-block
    const int rb = pt >> 3;                      // rows rb + 32 i
    float* s_tv = reinterpret_cast<float*>(smem + OFF_TV);
    int* s_row = reinterpret_cast<int*>(smem + OFF_ROW);

    // One tile's source rows (thread < 128: one row) and its two T rows (2 * K / 4 float4 over the 256 threads: one each for
    // K = 512): the global loads are issued into registers here and stored to shared memory a k-block later.
    auto stage_rows_issue = [&](int tile, int& g, float4& tv) {
      const int row0 = (tile / n_ctile) * BM;
      g = 0;
      if (pt < BM) {
        const uint32_t p = (uint32_t)row0 + pt;
        g = a.nbr[p] + (int)(p / a.rows_per_cloud) * (int)a.n_prev;
      }
      tv = make_float4(0.f, 0.f, 0.f, 0.f);
      const int f4_per_row = a.K / 4;
      if (pt < 2 * f4_per_row) {
        const int cg = pt / f4_per_row, off = pt - cg * f4_per_row;
        tv = *reinterpret_cast<const float4*>(a.T + (int64_t)(row0 / 64 + cg) * a.K + 4 * off);
      }
    };
    auto stage_rows_store = [&](int buf, int g, const float4& tv) {
      if (pt < BM) s_row[buf * BM + pt] = g;
      const int f4_per_row = a.K / 4;
      if (pt < 2 * f4_per_row) {
        const int cg = pt / f4_per_row, off = pt - cg * f4_per_row;
        *reinterpret_cast<float4*>(s_tv + (buf * 2 + cg) * KMAX + 4 * off) = tv;
      }
    };

    int cur_tile = 0, next_tile = -1;
    if (lane == 0) cur_tile = take_tile(0);
    cur_tile = __shfl_sync(FULL, cur_tile, 0);
    int buf = 0;
    // two register sets, alternating by k-block parity (n_kblk is even): the set being filled by the prefetch is never
    // copied, so nothing waits for those loads before the next k-block's conversion
    float4 ra[8], rb2[8];
    const float* zrow[4];
    if (cur_tile >= 0) {
      int g; float4 tv;
      stage_rows_issue(cur_tile, g, tv);
      stage_rows_store(0, g, tv);
      prod_bar_sync();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        zrow[i] = a.Z + (int64_t)s_row[rb + 32 * i] * a.ldz + 8 * c;
        ra[2 * i] = __ldg(reinterpret_cast<const float4*>(zrow[i]));
        ra[2 * i + 1] = __ldg(reinterpret_cast<const float4*>(zrow[i]) + 1);
      }
    }
    uint32_t job = 0;
    int pre_g = 0;
    float4 pre_tv = make_float4(0.f, 0.f, 0.f, 0.f);
    auto do_job = [&](const float4 (&cur)[8], float4 (&nxt)[8], int kb) {
      if (kb == 1) {                       // the loads issued at the top of the tile have had a k-block to land
        prod_bar_sync();                   // every producer is past the previous tile: its staging buffer is free
        if (next_tile >= 0) stage_rows_store(buf ^ 1, pre_g, pre_tv);
      } else if (kb == 2) {
        prod_bar_sync();                   // staged rows visible before the last k-block prefetches across the boundary
      }
      // ---- prefetch the next job's rows into the other register set ----
      if (kb + 1 < n_kblk) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4* src = reinterpret_cast<const float4*>(zrow[i] + (kb + 1) * BK);
          if (!(a.variant & 1)) {
            nxt[2 * i] = __ldg(src);
            nxt[2 * i + 1] = __ldg(src + 1);
          }
        }
      } else if (next_tile >= 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          zrow[i] = a.Z + (int64_t)s_row[(buf ^ 1) * BM + rb + 32 * i] * a.ldz + 8 * c;
          nxt[2 * i] = __ldg(reinterpret_cast<const float4*>(zrow[i]));
          nxt[2 * i + 1] = __ldg(reinterpret_cast<const float4*>(zrow[i]) + 1);
        }
      }
      // ---- convert the current job ----
      const uint32_t slot = job & 1;
      if (lane == 0) mbar_wait(bar_empty + 8 * slot, ((job >> 1) & 1) ^ 1);
      __syncwarp();
      const uint32_t sA = smem_base + OFF_A + slot * 2 * A_PLANE;
      const float* tv0 = s_tv + (buf * 2) * KMAX + kb * BK + 8 * c;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (a.variant & 2) break;
        const float* tvp = tv0 + (i >> 1) * KMAX;           // rows < 64: first centroid of the tile
        const float4 t0 = *reinterpret_cast<const float4*>(tvp), t1 = *reinterpret_cast<const float4*>(tvp + 4);
        const float4 z0 = cur[2 * i], z1 = cur[2 * i + 1];
        uint32_t h[4], l[4];
        relu_split_bf16_pair(__fadd_rn(z0.x, t0.x), __fadd_rn(z0.y, t0.y), h[0], l[0]);
        relu_split_bf16_pair(__fadd_rn(z0.z, t0.z), __fadd_rn(z0.w, t0.w), h[1], l[1]);
        relu_split_bf16_pair(__fadd_rn(z1.x, t1.x), __fadd_rn(z1.y, t1.y), h[2], l[2]);
        relu_split_bf16_pair(__fadd_rn(z1.z, t1.z), __fadd_rn(z1.w, t1.w), h[3], l[3]);
        const uint32_t r = (uint32_t)(rb + 32 * i);
        const uint32_t off = r * 128u + (((uint32_t)c ^ (r & 7u)) << 4);
        sts_v4(sA + off, h[0], h[1], h[2], h[3]);
        sts_v4(sA + A_PLANE + off, l[0], l[1], l[2], l[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * slot);
      ++job;
    };
    for (uint32_t it = 0; cur_tile >= 0; ++it) {
      if (lane == 0) next_tile = take_tile(it + 1);
      next_tile = __shfl_sync(FULL, next_tile, 0);
      if (next_tile >= 0) stage_rows_issue(next_tile, pre_g, pre_tv);
      for (int kb = 0; kb < n_kblk; kb += 2) {
        do_job(ra, rb2, kb);
        do_job(rb2, ra, kb + 1);
      }
      cur_tile = next_tile;
      buf ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// Z'[j, c] = scale[c] * (Z[j, c] + Wx[c, :] . xyz[j]) in place (rows = B * n_prev source points),
// T[m, c]  = shift[c] - scale[c] * (Wx[c, :] . centre[m])       (rows = B * M centroids).
// A thread keeps one channel quad (its 12 weights, scale and shift in registers) and walks rows.
__global__ void __launch_bounds__(256)
fold_kernel(float* __restrict__ Z, int ldz, uint32_t n_prev, const float* __restrict__ xyz, Strides3 xst,
            const float* __restrict__ new_xyz, uint32_t M, const float* __restrict__ Wx, int ldw, const float* __restrict__ scale,
            const float* __restrict__ shift, int C, uint32_t z_rows, uint32_t t_rows, float* __restrict__ T) {
  const int c4 = C / 4;                                   // <= 256 and a divisor of 256 (host check)
  const int c = (threadIdx.x % c4) * 4;
  const uint32_t rows_per_iter = 256 / c4, sub = threadIdx.x / c4;
  float w[4][3];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int d = 0; d < 3; ++d) w[u][d] = Wx[(c + u) * ldw + d];
  const float4 sc = *reinterpret_cast<const float4*>(scale + c);
  const float4 sh = *reinterpret_cast<const float4*>(shift + c);
  for (uint32_t row = blockIdx.x * rows_per_iter + sub; row < z_rows + t_rows; row += gridDim.x * rows_per_iter) {
    const bool is_t = row >= z_rows;
    float p[3];
    if (!is_t) {
      const uint32_t b = row / n_prev, j = row - b * n_prev;
#pragma unroll
      for (int d = 0; d < 3; ++d) p[d] = xyz[(int64_t)b * xst.b + d * xst.c + (int64_t)j * xst.n];
    } else {
      const uint32_t r = row - z_rows, b = r / M, m = r - b * M;
#pragma unroll
      for (int d = 0; d < 3; ++d) p[d] = new_xyz[((int64_t)b * 3 + d) * M + m];
    }
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = fmaf(w[u][2], p[2], fmaf(w[u][1], p[1], __fmul_rn(w[u][0], p[0])));
    if (!is_t) {
      float4* zp = reinterpret_cast<float4*>(Z + (int64_t)row * ldz + c);
      const float4 z = *zp;
      *zp = make_float4(__fmul_rn(sc.x, __fadd_rn(z.x, v[0])), __fmul_rn(sc.y, __fadd_rn(z.y, v[1])),
                        __fmul_rn(sc.z, __fadd_rn(z.z, v[2])), __fmul_rn(sc.w, __fadd_rn(z.w, v[3])));
    } else {
      *reinterpret_cast<float4*>(T + (int64_t)(row - z_rows) * C + c) =
          make_float4(fmaf(-sc.x, v[0], sh.x), fmaf(-sc.y, v[1], sh.y), fmaf(-sc.z, v[2], sh.z), fmaf(-sc.w, v[3], sh.w));
    }
  }
}

// The same operand MATERIALISED as bf16 hi/lo planes (rows, C): out[p, :] = relu(Z'[g(p), :] + T[p / 64, :]) with the very
// roundings of the producer warps above, for the plan's co-running mode (scorenet.cu: next to a prefetched FPS the plain
// GEMM is the better neighbour) -- both forms feed identical bits to identical MMA sequences.
// One warp = 32 consecutive rows (one centroid), lane = 8 channels of a 256-channel slab (blockIdx.y).
__global__ void __launch_bounds__(256)
gather_add_kernel(const float* __restrict__ Z, int ldz, uint32_t n_prev, const float* __restrict__ T,
                  const int32_t* __restrict__ nbr, uint32_t rows_per_cloud, int C, uint32_t rows,
                  __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const uint32_t lane = threadIdx.x & 31;
  const int c0 = (int)(blockIdx.y * 256 + lane * 8);
  const bool live = c0 < C;
  const int cc = live ? c0 : C - 8;
  const uint32_t nwarps = gridDim.x * 8;
  for (uint32_t base = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 32; base < rows; base += nwarps * 32) {
    const int j = nbr[base + lane] + (int)(base / rows_per_cloud) * (int)n_prev;
    const float* __restrict__ tp = T + (int64_t)(base >> 6) * C + cc;
    const float4 t0 = *reinterpret_cast<const float4*>(tp), t1 = *reinterpret_cast<const float4*>(tp + 4);
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const int jr = __shfl_sync(0xffffffffu, j, r);
      const float4 z0 = __ldg(reinterpret_cast<const float4*>(Z + (int64_t)jr * ldz + cc));
      const float4 z1 = __ldg(reinterpret_cast<const float4*>(Z + (int64_t)jr * ldz + cc) + 1);
      uint32_t h[4], l[4];
      relu_split_bf16_pair(__fadd_rn(z0.x, t0.x), __fadd_rn(z0.y, t0.y), h[0], l[0]);
      relu_split_bf16_pair(__fadd_rn(z0.z, t0.z), __fadd_rn(z0.w, t0.w), h[1], l[1]);
      relu_split_bf16_pair(__fadd_rn(z1.x, t1.x), __fadd_rn(z1.y, t1.y), h[2], l[2]);
      relu_split_bf16_pair(__fadd_rn(z1.z, t1.z), __fadd_rn(z1.w, t1.w), h[3], l[3]);
      if (live) {
        const int64_t off = (int64_t)(base + r) * C + c0;
        *reinterpret_cast<uint4*>(out_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(out_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
  }
}

}  // namespace

int sa_gather_add_launch(const float* Z, int ldz, int n_prev, const float* T, const int32_t* nbr, int rows_per_cloud, int C,
                         int64_t rows, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t stream) {
  RN_CHECK_ARG(Z && T && nbr && out_hi && out_lo, "sa_gather_add: null argument");
  RN_CHECK_ARG(C % 8 == 0 && ldz % 4 == 0 && rows % 64 == 0 && rows < (1LL << 31) && rows_per_cloud % 64 == 0,
               "sa_gather_add: bad shape");
  if (rows == 0) return REGNET_OK;
  const unsigned yw = (unsigned)ceil_div(C, 256);
  const unsigned gx = (unsigned)std::min<int64_t>(ceil_div((int)(rows / 32), 8), std::max(1, 148 * 8 / (int)yw));
  RN_PREFER_MAX_SMEM(gather_add_kernel);
  gather_add_kernel<<<dim3(gx, yw), 256, 0, stream>>>(Z, ldz, (uint32_t)n_prev, T, nbr, (uint32_t)rows_per_cloud, C,
                                                       (uint32_t)rows, out_hi, out_lo);
  RN_LAUNCH_CHECK("sa gather_add_kernel");
  return REGNET_OK;
}

int sa_fold_launch(float* Z, int ldz, int n_prev, const float* xyz, Strides3 xst, const float* new_xyz, int M, const float* Wx,
                   int ldw, const float* scale, const float* shift, int B, int C, float* T, cudaStream_t stream) {
  RN_CHECK_ARG(Z && xyz && new_xyz && Wx && scale && shift && T, "sa_fold: null argument");
  RN_CHECK_ARG(C % 4 == 0 && C <= 1024 && 256 % (C / 4) == 0 && ldz % 4 == 0 && B > 0 && n_prev > 0 && M > 0,
               "sa_fold: bad shape (C / 4 must divide 256)");
  const int64_t iters = ((int64_t)B * (n_prev + M) + 256 / (C / 4) - 1) / (256 / (C / 4));
  const unsigned grid = (unsigned)std::min<int64_t>(iters, 148 * 8);
  fold_kernel<<<grid, 256, 0, stream>>>(Z, ldz, (uint32_t)n_prev, xyz, xst, new_xyz, (uint32_t)M, Wx, ldw, scale, shift, C,
                                        (uint32_t)(B * n_prev), (uint32_t)(B * M), T);
  RN_LAUNCH_CHECK("sa fold_kernel");
  return REGNET_OK;
}

int gemm_fused_a_supported(int64_t P, int K, int rows_per_cloud) {
  return gemm_tc_supported() && P > 0 && P % BM == 0 && P < (1LL << 31) && K % (2 * BK) == 0 && K >= 4 * BK && K <= KMAX &&
         rows_per_cloud % 64 == 0;
}

int gemm_fused_a_launch(const float* Z, int ldz, const float* T, const int32_t* nbr, int rows_per_cloud, int n_prev,
                        const __nv_bfloat16* Whi, const __nv_bfloat16* Wlo, int ldw, int64_t P, int K, int cout,
                        const float* scale, const float* shift, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int ld_out,
                        unsigned int* tile_counter, cudaStream_t stream) {
  RN_CHECK_ARG(Z && T && nbr && Whi && Wlo && out_hi && out_lo && tile_counter, "gemm_fused_a: null argument");
  RN_CHECK_ARG(gemm_fused_a_supported(P, K, rows_per_cloud), "gemm_fused_a: unsupported shape (P=%lld K=%d)", (long long)P, K);
  RN_CHECK_ARG(ldz % 4 == 0 && ldz >= K && ldw % 8 == 0 && ldw >= K && ld_out % 8 == 0 && cout > 0,
               "gemm_fused_a: bad leading dimensions");
  RN_CHECK_ARG((reinterpret_cast<uintptr_t>(Z) & 15) == 0 && (reinterpret_cast<uintptr_t>(T) & 15) == 0,
               "gemm_fused_a: tables must be 16-byte aligned");
  CUtensorMap wh, wl, oh, ol;
  RN_TRY(tc_make_map(&wh, Whi, cout, K, ldw, BN, BK, 128));
  RN_TRY(tc_make_map(&wl, Wlo, cout, K, ldw, BN, BK, 128));
  RN_TRY(tc_make_map(&oh, out_hi, P, cout, ld_out, 32, 32, 64));
  RN_TRY(tc_make_map(&ol, out_lo, P, cout, ld_out, 32, 32, 64));
  FusedAArgs a;
  a.Z = Z; a.ldz = ldz; a.T = T; a.nbr = nbr;
  a.rows_per_cloud = (uint32_t)rows_per_cloud; a.n_prev = (uint32_t)n_prev;
  a.P = (uint32_t)P; a.K = K; a.cout = cout; a.scale = scale; a.shift = shift; a.tile_counter = tile_counter;
  a.variant = 0;
  if (const char* e = getenv("REGNET_FUSED_A_VARIANT")) a.variant = atoi(e);
  int dev = 0, sms = 0;
  RN_CUDA(cudaGetDevice(&dev));
  RN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  RN_CUDA(cudaFuncSetAttribute(gemm_fused_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const int64_t n_tiles = (P / BM) * ((cout + BN - 1) / BN);
  const int grid = (int)std::min<int64_t>(n_tiles, sms);
  gemm_fused_a_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(wh, wl, oh, ol, a);
  RN_LAUNCH_CHECK("gemm_fused_a_kernel");
  return REGNET_OK;
}

}  // namespace regnet
