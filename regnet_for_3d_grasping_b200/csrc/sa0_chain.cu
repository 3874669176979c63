// Whole shared-MLP chain of set-abstraction level 0 in ONE kernel: gather + centre + 6 -> 128 -> 128 -> 256 + max-pool.
//
// Reference chain (pn2_utils/modules.py:44-52, three nn/modules/conv.py:64-76 blocks, modules.py:245): group xyz / rgb,
// subtract the centroid, concat, 3 x (1x1 conv + BN + ReLU), max over the 64 neighbours -- for the B=15 batch that is
// 4.9 M positions and, unfused, 2 x 2.5 GB of 128-wide activations through HBM per layer boundary.  Here no
// activation ever leaves the SM: the three layers are chained THROUGH TENSOR MEMORY.
//
//   layer 0 (K = 6, padded to 16)  SS-MMA : A0 = gathered rows (smem, written by the producer warps), B = W0 (smem)
//   layer 1 (K = 128)              TS-MMA : A1 = layer-0 activations, bf16 hi/lo, IN TMEM          , B = W1 (smem)
//   layer 2 (K = 128)              TS-MMA : A2 = layer-1 activations, bf16 hi/lo, IN TMEM          , B = W2 (smem)
//
// Every product is the split-bf16 triple hi*hi + lo*hi + hi*lo with fp32 accumulation (gemm_tc.cu).  W1 (64 KB) and W2
// (128 KB) are loaded once per CTA by TMA and stay resident; that leaves no room for 64 KB activation tiles in shared
// memory, hence tcgen05.mma with the A operand in TMEM: the converter warps read a 16-column slice of the fp32
// accumulator (tcgen05.ld), apply BN + ReLU, split to bf16 hi/lo and write the 8 + 8 packed columns back IN PLACE
// (tcgen05.st) -- the accumulator region of layer l becomes the A operand of layer l+1.
//
// TMEM map (512 columns, all of it):  X = [0,128)  acc0 -> A1     Y = [128,256)  acc1 -> A2     Z = [256,512)  acc2
//
// Warp roles (448 threads, one CTA per SM, tiles of 128 positions = 2 centroids drawn from a global counter):
//   warp 0       scheduler + producer: draws the tile, gathers its 128 rows (neighbour index -> rgb, xyz - centroid) as
//                bf16 hi/lo into the A0 stage, up to two tiles ahead; one-time TMA load of W1 / W2
//   warp 1       MMA issuer (one lane); issue order L1(i), L0(i+1), L2(i): the tensor pipe runs tile i+1's first layer
//                and tile i's second layer while the workers drain Z of tile i-1; hazards on X / Y are ordered by the
//                pipe itself
//   warps 2..13  workers, three per TMEM lane quarter: conversions acc -> A (in 16-channel k-steps, each 32-channel chunk
//                released to the MMA lane as soon as its two k-steps are written) and the pool epilogue acc2 -> max over
//                the 64 rows of each centroid -> BN + ReLU -> (B*M, 256) fp32; TMEM loads are software pipelined
#include <cuda.h>

#include "gemm.cuh"
#include "internal.cuh"
#include "tc_ptx.cuh"

namespace regnet {

using namespace tc;

namespace {

constexpr int BM = 128;                        // positions per tile = TMEM lanes
constexpr int C1 = 128;                        // width of layers 0 and 1
constexpr int C2 = 256;                        // width of layer 2
constexpr int BK = 64;
constexpr int W1_PLANE = C1 * BK * 2;          // 16 KB: one bf16 plane of one k-block of W1
constexpr int W2_PLANE = C2 * BK * 2;          // 32 KB
constexpr int A0_PLANE = BM * 32;              // 4 KB: [128 rows x 16 bf16]
constexpr int OFF_W1 = 0;                      // [kb][hi|lo]  64 KB
constexpr int OFF_W2 = OFF_W1 + 4 * W1_PLANE;  // [kb][hi|lo] 128 KB
constexpr int OFF_A0 = OFF_W2 + 4 * W2_PLANE;  // 2 stages x [hi|lo]
constexpr int OFF_W0 = OFF_A0 + 4 * A0_PLANE;  // [hi|lo]
constexpr int OFF_SC = OFF_W0 + 2 * A0_PLANE;  // scale0, shift0, scale1, shift1 [128]; scale2, shift2 [256]
constexpr int OFF_PART = OFF_SC + 4096;        // pooled partials [4][256] u32
constexpr int OFF_BAR = OFF_PART + 4096;
// No alignment slack and a 256-byte barrier block: the kernel must leave ~2.5 KB of the SM's 228 KB for a co-resident
// FPS CTA of the next step (1.3 KB of records + its 1 KB system reservation), whichever of the two arrives first --
// with 1.5 KB left an FPS launch that met a running sa0_chain waited for the whole kernel (timeline: fps.2 0.2 -> 1.2 ms).
constexpr int SMEM_BYTES = OFF_BAR + 256;
constexpr int NTHREADS = 448;
constexpr uint32_t TM_X = 0, TM_Y = 128, TM_Z = 256;
static_assert(SMEM_BYTES <= 232448, "sa0_chain: shared memory budget");

// tcgen05.mma with the A operand in tensor memory (cute SM100_MMA_F16BF16_TS): A is [128 lanes x 16 bf16], two
// consecutive k per 32-bit column, i.e. 8 columns per instruction.
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// two fp32 -> one packed bf16x2 hi word + one lo word (same rounding as split_bf16: hi = rn(x), lo = rn(x - hi));
// the even element sits in the low half of the word = the lower k index of the TMEM / shared-memory operand
__device__ __forceinline__ void split_pair(float y0, float y1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);   // .x (low half) = y0
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(__fsub_rn(y0, h0), __fsub_rn(y1, h1));
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// ReLU folded into the conversions: hi = bf16_rz(max(y, 0)) -- rounding toward zero keeps y - hi >= 0 for y >= 0, so the
// second relu-conversion lo = bf16_rn(max(y - hi, 0)) is exact for y >= 0 and gives hi = lo = 0 for y < 0.  No FMNMX;
// hi + lo still carries 16 significant bits (residual <= 2^-17 |y|).
__device__ __forceinline__ void relu_split_pair(float y0, float y1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y1), "f"(y0));   // d = {hi half: a, lo half: b}
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(__fsub_rn(y1, h1)), "f"(__fsub_rn(y0, h0)));
}

// Streaming read-only global loads that do not allocate in L1: with 231 KB of shared memory configured the L1 is ~28 KB,
// and the 128-row gathers of every tile (~250 lines) would evict the few local-memory lines of the register-starved
// workers (ncu: 29 % L1 hit rate on local loads before).
__device__ __forceinline__ float ldg_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ldg_stream_s32(const int32_t* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

struct Sa0ChainArgs {
  const float* xyz; Strides3 xst;
  const float* new_xyz;
  const float* feat; int64_t feat_bstride; int feat_ld;
  const int32_t* nbr;
  const float* W0; int ldw0;                       // (128, >= 6) fp32, operand order [feature(3) | xyz_rel(3)]
  const float* scale0; const float* shift0;
  const float* scale1; const float* shift1;
  const float* scale2; const float* shift2;        // scale2 must be >= 0 (see the pool epilogue)
  float* out; int ld_out;                          // (rows / 64, 256)
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;    // optional: the same values as bf16 hi/lo planes (rows / 64, 256)
  float* dbg;                                      // MODE 1: (rows, 256): raw acc0 | raw acc1
  uint32_t M, rows;
  unsigned int* tile_counter;
  int variant;
  unsigned long long* timing;                      // TIMING kernels: (grid, 5 roles, 8 counters)
};

// MODE 0: production.  1: also dumps the raw accumulators of layers 0 and 1 to a.dbg (tests).  2: per-role wait
// counters (clock64 around every mbarrier wait) to a.timing (scripts/sa0_chain_timing.py).
template <int MODE>
__global__ void __maxnreg__(64)   // 448 x 64 registers leave room for the co-resident FPS CTA of the next step (4 warps x 227)
sa0_chain_kernel(const __grid_constant__ CUtensorMap map_w1hi, const __grid_constant__ CUtensorMap map_w1lo,
                 const __grid_constant__ CUtensorMap map_w2hi, const __grid_constant__ CUtensorMap map_w2lo,
                 const Sa0ChainArgs a) {
  constexpr bool TIMING = MODE == 2;
  // the dynamic region starts right behind the 1 KB the system reserves per CTA (this kernel has no static shared
  // memory), i.e. 1024-byte aligned as the SWIZZLE_128B operands need; checked, not assumed
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn;
  const uint32_t sb = smem_u32(smem);
  if (sb & 1023u) __trap();
  const uint32_t bar_w = sb + OFF_BAR;            // W1 / W2 landed
  const uint32_t bar_a0full = bar_w + 8;          // [2] producers -> MMA
  const uint32_t bar_a0empty = bar_a0full + 16;   // [2] MMA -> producers
  const uint32_t bar_acc0 = bar_a0empty + 16;     // L0 complete -> converters
  const uint32_t bar_acc1 = bar_acc0 + 8;         // L1 complete -> converters
  const uint32_t bar_acc2 = bar_acc1 + 8;         // L2 complete -> workers (pool)
  const uint32_t bar_z_empty = bar_acc2 + 8;      // workers -> MMA: Z drained
  const uint32_t bar_a1 = bar_z_empty + 8;        // [4] workers -> MMA, per 32-channel chunk of A1
  const uint32_t bar_a2 = bar_a1 + 32;            // [4] same for A2
  const uint32_t bar_sfull = bar_a2 + 32;         // [4] scheduler ring
  const uint32_t bar_sempty = bar_sfull + 32;     // [4]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * 26);
  volatile int* ring = reinterpret_cast<volatile int*>(tmem_holder + 1);
  float* s_sc = reinterpret_cast<float*>(smem + OFF_SC);   // [0,128) scale0 [128,256) shift0 [256,384) scale1 [384,512) shift1
  float* s_sc2 = s_sc + 512;                               // [0,256) scale2 [256,512) shift2

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_tiles = (a.rows + BM - 1) / BM;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_a0full + 8 * s, 32);
      mbar_init(bar_a0empty + 8 * s, 1);
    }
    mbar_init(bar_acc0, 1);
    mbar_init(bar_acc1, 1);
    mbar_init(bar_acc2, 1);
    mbar_init(bar_z_empty, 12);            // one lane of each worker warp
    for (int c = 0; c < 4; ++c) {
      mbar_init(bar_a1 + 8 * c, 256);      // a chunk = two k-steps, converted by two different worker warps per quarter
      mbar_init(bar_a2 + 8 * c, 256);
      mbar_init(bar_sfull + 8 * c, 1);
      mbar_init(bar_sempty + 8 * c, 13);   // MMA lane + one lane of each worker warp
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_w1hi);
    tma_prefetch_desc(&map_w1lo);
    tma_prefetch_desc(&map_w2hi);
    tma_prefetch_desc(&map_w2lo);
  }
  // one-time shared-memory set-up by all threads: A0 stages zeroed (columns 6..15 stay zero for ever), W0 as a
  // K-major bf16 hi/lo operand, BN scale / shift of the three layers
  for (int i = threadIdx.x; i < 6 * A0_PLANE / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(smem + OFF_A0)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (threadIdx.x < C1) {
    const int c = threadIdx.x;
    uint32_t hi[3], lo[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      split_pair(a.W0[c * a.ldw0 + 2 * j], a.W0[c * a.ldw0 + 2 * j + 1], hi[j], lo[j]);
    const uint32_t off = k16_offset((uint32_t)c, 0);
    *reinterpret_cast<uint4*>(smem + OFF_W0 + off) = make_uint4(hi[0], hi[1], hi[2], 0u);
    *reinterpret_cast<uint4*>(smem + OFF_W0 + A0_PLANE + off) = make_uint4(lo[0], lo[1], lo[2], 0u);
    s_sc[c] = a.scale0[c];
    s_sc[128 + c] = a.shift0[c];
    s_sc[256 + c] = a.scale1[c];
    s_sc[384 + c] = a.shift1[c];
  } else if (threadIdx.x < C1 + C2) {
    const int c = threadIdx.x - C1;
    s_sc2[c] = a.scale2[c];
    s_sc2[256 + c] = a.shift2[c];
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  unsigned long long tw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_begin = TIMING ? clock64() : 0;
  auto wait_t = [&](uint32_t bar, uint32_t parity, int k) {   // mbar_wait, accounted to counter k when TIMING
    if (TIMING) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      tw[k] += (unsigned long long)(clock64() - t0);
    } else {
      mbar_wait(bar, parity);
    }
  };
  auto flush_t = [&](int role) {
    if (TIMING && lane == 0 && a.timing) {
      tw[7] = (unsigned long long)(clock64() - t_begin);
      for (int k = 0; k < 8; ++k) a.timing[((size_t)blockIdx.x * 5 + role) * 8 + k] = tw[k];
    }
  };
  auto take_tile = [&](uint32_t it) -> int {
    const uint32_t slot = it & 3;
    wait_t(bar_sfull + 8 * slot, (it >> 2) & 1, 0);
    return ring[slot];
  };

  if (warp == 0) {
    // ================= scheduler + A0 producer + one-time weight load =================
    // lane 0 draws the tile and publishes it; then the whole warp gathers the tile's 128 rows (4 per lane: neighbour
    // index -> rgb, xyz - centroid), splits them to bf16 hi/lo and writes the A0 stage.  The ring lets this warp run
    // up to 4 tiles ahead of the slowest consumer, the two A0 stages up to 2 tiles ahead of layer 0.
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, 4 * W1_PLANE + 4 * W2_PLANE);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_2d(sb + OFF_W1 + kb * 2 * W1_PLANE, &map_w1hi, bar_w, kb * BK, 0);
        tma_load_2d(sb + OFF_W1 + kb * 2 * W1_PLANE + W1_PLANE, &map_w1lo, bar_w, kb * BK, 0);
        tma_load_2d(sb + OFF_W2 + kb * 2 * W2_PLANE, &map_w2hi, bar_w, kb * BK, 0);
        tma_load_2d(sb + OFF_W2 + kb * 2 * W2_PLANE + W2_PLANE, &map_w2lo, bar_w, kb * BK, 0);
      }
    }
    for (uint32_t it = 0;; ++it) {
      uint32_t t = 0;
      if (lane == 0) {
        const uint32_t slot = it & 3;
        if (it >= 4) wait_t(bar_sempty + 8 * slot, ((it >> 2) - 1) & 1, 1);
        t = a.tile_counter ? atomicAdd(a.tile_counter, 1u) : blockIdx.x + it * gridDim.x;
        ring[slot] = t < n_tiles ? (int)t : -1;
        mbar_arrive(bar_sfull + 8 * slot);
      }
      t = __shfl_sync(FULL, t, 0);
      if (t >= n_tiles) break;
      int jn[4];
      float c3[4][3];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t row = t * BM + u * 32 + lane;
        const bool ok = row < a.rows;
        const uint32_t bm = ok ? row >> 6 : 0;      // 64 neighbours per centroid
        const uint32_t b = bm / a.M, m = bm - b * a.M;
        jn[u] = ok ? ldg_stream_s32(a.nbr + row) : -1;
#pragma unroll
        for (int x = 0; x < 3; ++x) c3[u][x] = ldg_stream_f32(a.new_xyz + ((int64_t)b * 3 + x) * a.M + m);
      }
      float v[4][6];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t row = t * BM + u * 32 + lane;
        const uint32_t b = (row >> 6) / a.M;
        const int j = jn[u];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          v[u][c] = j >= 0 ? ldg_stream_f32(a.feat + (int64_t)b * a.feat_bstride + (int64_t)j * a.feat_ld + c) : 0.f;
#pragma unroll
        for (int x = 0; x < 3; ++x)
          v[u][3 + x] = j >= 0 ? ldg_stream_f32(a.xyz + (int64_t)b * a.xst.b + x * a.xst.c + (int64_t)j * a.xst.n) : 0.f;
      }
      const uint32_t s = it & 1, ph = (it >> 1) & 1;
      wait_t(bar_a0empty + 8 * s, ph ^ 1, 2);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t hi[3], lo[3];
        const bool ok = jn[u] >= 0;
        split_pair(v[u][0], v[u][1], hi[0], lo[0]);
        split_pair(v[u][2], ok ? __fsub_rn(v[u][3], c3[u][0]) : 0.f, hi[1], lo[1]);
        split_pair(ok ? __fsub_rn(v[u][4], c3[u][1]) : 0.f, ok ? __fsub_rn(v[u][5], c3[u][2]) : 0.f, hi[2], lo[2]);
        const uint32_t a0 = sb + OFF_A0 + s * 2 * A0_PLANE + k16_offset((uint32_t)(u * 32 + lane), 0);
        sts_v4(a0, hi[0], hi[1], hi[2], 0u);
        sts_v4(a0 + A0_PLANE, lo[0], lo[1], lo[2], 0u);
      }
      fence_proxy_async();
      mbar_arrive(bar_a0full + 8 * s);
    }
    flush_t(0);
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(BM, C1), idesc2 = make_idesc(BM, C2);
      const uint64_t w0_hi = make_sdesc_k16(sb + OFF_W0), w0_lo = make_sdesc_k16(sb + OFF_W0 + A0_PLANE);
      auto issue_l0 = [&](uint32_t it) {
        const uint32_t s = it & 1, ph = (it >> 1) & 1;
        wait_t(bar_a0full + 8 * s, ph, 1);
        tc_fence_after();
        const uint32_t a0 = sb + OFF_A0 + s * 2 * A0_PLANE;
        const uint64_t a_hi = make_sdesc_k16(a0), a_lo = make_sdesc_k16(a0 + A0_PLANE);
        umma_f16(tmem_base + TM_X, a_hi, w0_hi, idesc1, 0);
        umma_f16(tmem_base + TM_X, a_lo, w0_hi, idesc1, 1);
        umma_f16(tmem_base + TM_X, a_hi, w0_lo, idesc1, 1);
        umma_commit(bar_a0empty + 8 * s);
        umma_commit(bar_acc0);
      };
      // one 32-channel chunk (2 k-steps) of a TS layer: A hi/lo at columns 16*ks / 16*ks + 8 of `a_region`.  (Layer 2
      // stays ONE N = 256 instruction per k-step: the issue path costs ~80 cycles per tcgen05.mma -- elect + 4 R2UR --
      // which an N = 128 instruction (64 cycles of tensor work) cannot hide.)
      auto issue_chunk = [&](uint32_t d, uint32_t a_region, uint32_t w_base, uint32_t w_plane, uint32_t idesc, int ch) {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ks = 2 * ch + h;
            const uint32_t wk = w_base + (ks >> 2) * 2 * w_plane;
            const uint64_t b_hi = make_sdesc(wk) + 2 * (ks & 3), b_lo = make_sdesc(wk + w_plane) + 2 * (ks & 3);
            const uint32_t a_hi = a_region + 16 * ks, a_lo = a_hi + 8;
            if (pass == 0) umma_f16_ts(d, a_hi, b_hi, idesc, ks != 0);
            if (pass == 1) umma_f16_ts(d, a_lo, b_hi, idesc, 1);
            if (pass == 2) umma_f16_ts(d, a_hi, b_lo, idesc, 1);
          }
        }
      };
      int tile = take_tile(0);
      mbar_arrive(bar_sempty);
      if (tile >= 0) {
        issue_l0(0);
        wait_t(bar_w, 0, 2);
      }
      for (uint32_t it = 0; tile >= 0; ++it) {
        const uint32_t ph = it & 1;
        for (int ch = 0; ch < 4; ++ch) {
          wait_t(bar_a1 + 8 * ch, ph, 3 + (ch ? 1 : 0));
          tc_fence_after();
          issue_chunk(tmem_base + TM_Y, tmem_base + TM_X, sb + OFF_W1, W1_PLANE, idesc1, ch);
        }
        umma_commit(bar_acc1);
        const int next = take_tile(it + 1);
        mbar_arrive(bar_sempty + 8 * ((it + 1) & 3));
        if (next >= 0) issue_l0(it + 1);   // overwrites X after L1(it) in pipe order
        if (it > 0) {
          wait_t(bar_z_empty, (it - 1) & 1, 6);
          tc_fence_after();
        }
        for (int ch = 0; ch < 4; ++ch) {
          wait_t(bar_a2 + 8 * ch, ph, 5);
          tc_fence_after();
          issue_chunk(tmem_base + TM_Z, tmem_base + TM_Y, sb + OFF_W2, W2_PLANE, idesc2, ch);
        }
        umma_commit(bar_acc2);
        tile = next;
      }
      flush_t(1);
    }
    __syncwarp();
  } else {
    // ================= workers (warps 2..13): conversions and the pool epilogue =================
    // Three warps per TMEM lane quarter q = warp % 4; worker w = 0, 1, 2 of a quarter takes the items w, w+3, w+6 of every
    // job -- the 8 k-steps (16 channels) of a conversion, the 8 column chunks (32 channels) of the pool.  Per tile the
    // jobs come in the order in which the tensor pipe produces their inputs:
    //     acc1(it) -> A2 (layer 2 waits for it)   acc0(it+1) -> A1   acc2(it) -> pooled output
    // Consecutive k-steps of a conversion belong to different workers, so the first 32-channel chunk reaches the MMA
    // lane after ONE k-step time and every chunk barrier collects 2 x 128 arrivals.
    // Thread coordinates are re-derived from a volatile %tid.x read inside every job instead of being kept live across
    // the tile loop: with 64 registers (the budget that lets the next step's FPS CTA share the scheduler's register
    // file) two loop-carried address registers were spilled, and with ~4 KB of L1 left a spill reload is an L2 round trip.
    auto tid_now = []() -> uint32_t {
      uint32_t t;
      asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
      return t;
    };
    const int q = warp & 3;

    // accumulator -> BN + ReLU -> bf16 hi/lo A operand, in place in TMEM; region X (layer 0 -> 1) or Y (layer 1 -> 2)
    auto convert = [&](bool second, float* dbg_row) {
      const uint32_t t = tid_now();
      const int w = (int)((t >> 5) - 2) >> 2;
      const uint32_t lane_base = tmem_base + ((((t >> 5) & 3u) * 32u) << 16) + (second ? TM_Y : TM_X);
      const uint32_t sc = sb + OFF_SC + (second ? 1024 : 0), sh = sc + 512;   // byte addresses of scale / shift
      const uint32_t bar_ready = second ? bar_a2 : bar_a1;
      const long long tc0 = TIMING ? clock64() : 0;
      // No software pipelining: the three workers of a quarter hide each other's TMEM latencies.  Any second buffer
      // (or a load issued into v while the permuted outputs are still live) spills, and with 231 KB of shared memory
      // configured the L1 is ~28 KB: local-memory traffic goes to L2, every spill costs hundreds of cycles (ncu counted
      // 38 M local accesses in an earlier version that was 2x slower).
      auto convert_step = [&](int ks) {
        uint32_t v[16];
        tmem_ld16(lane_base + 16 * ks, v);
        if (MODE == 1 && dbg_row) {
#pragma unroll
          for (int j = 0; j < 16; ++j) dbg_row[16 * ks + j] = __uint_as_float(v[j]);
        }
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {     // converted in place: v[4 j4 ..] <- h h l l
          const float4 s4 = lds_f4(sc + 4 * (16 * ks + 4 * j4));
          const float4 t4 = lds_f4(sh + 4 * (16 * ks + 4 * j4));
          const float y0 = fmaf(__uint_as_float(v[4 * j4 + 0]), s4.x, t4.x);
          const float y1 = fmaf(__uint_as_float(v[4 * j4 + 1]), s4.y, t4.y);
          const float y2 = fmaf(__uint_as_float(v[4 * j4 + 2]), s4.z, t4.z);
          const float y3 = fmaf(__uint_as_float(v[4 * j4 + 3]), s4.w, t4.w);
          uint32_t h0, l0, h1, l1;
          relu_split_pair(y0, y1, h0, l0);
          relu_split_pair(y2, y3, h1, l1);
          v[4 * j4 + 0] = h0; v[4 * j4 + 1] = h1; v[4 * j4 + 2] = l0; v[4 * j4 + 3] = l1;
        }
        // v = [h h l l | h h l l | h h l l | h h l l] -> the TMEM layout [8 hi words | 8 lo words]
        const uint32_t o[16] = {v[0], v[1], v[4], v[5], v[8], v[9], v[12], v[13], v[2], v[3], v[6], v[7], v[10], v[11], v[14], v[15]};
        tmem_st16(lane_base + 16 * ks, o);
        tmem_st_wait();                     // this k-step is in TMEM: count it on its chunk's barrier
        tc_fence_before();
        mbar_arrive(bar_ready + 8 * (ks >> 1));
      };
      convert_step(w);
      convert_step(w + 3);
      if (w + 6 < 8) convert_step(w + 6);
      if (TIMING) tw[2] += (unsigned long long)(clock64() - tc0);
    };
    auto dbg_row_of = [&](int tile, bool second) -> float* {
      if (MODE != 1) return nullptr;
      const uint32_t row = (uint32_t)tile * BM + (tid_now() & 127u);   // quarter * 32 + lane
      return (MODE == 1 && a.dbg && row < a.rows) ? a.dbg + (size_t)row * 256 + (second ? 128 : 0) : nullptr;
    };
    // acc2 -> max over each centroid's 64 rows -> BN + ReLU -> (B*M, 256) fp32.  scale2 >= 0 (the launcher's contract:
    // rows of W2 with a negative BN scale are negated when the weights are prepared), so relu(scale * max(acc) + shift)
    // == max(relu(scale * acc + shift)): the max is taken on the raw accumulators -- 4 rows per thread in registers
    // (16x256b TMEM loads), then 3 shuffle levels (tc_ptx.cuh); the affine map + ReLU once per pooled value.  16 columns
    // per item (.x2 loads): the 32-column version needs 32 live registers and spilled.
    auto pool_drain = [&]() {
      const long long tp0 = TIMING ? clock64() : 0;
      const uint32_t t = tid_now();
      const int lane = (int)(t & 31u), q = (int)((t >> 5) & 3u), w = (int)((t >> 5) - 2) >> 2;
      const int my_col = colmax16_column(lane);
      const uint32_t s_part_addr = sb + OFF_PART;
      const uint32_t t_lo = tmem_base + ((uint32_t)(q * 32) << 16) + TM_Z, t_hi = t_lo + (16u << 16);
      uint32_t va[8], vb[8];
      tmem_ld_16x256b_x2_async(t_lo + 16 * w, va);
      tmem_ld_16x256b_x2_async(t_hi + 16 * w, vb);
      asm volatile("bar.sync 2, 384;" ::: "memory");   // every worker is done reading the previous tile's partials
#pragma unroll 1
      for (int c = w; c < C2 / 16; c += 3) {           // this worker's items: 16 columns each (8 + 8 registers)
        tmem_ld_wait();
        float m[4];
        colmax16_rows4(va, vb, m);
        if (c + 3 < C2 / 16) {                         // next item's loads fly during the shuffles
          tmem_ld_16x256b_x2_async(t_lo + 16 * (c + 3), va);
          tmem_ld_16x256b_x2_async(t_hi + 16 * (c + 3), vb);
        }
        const float r = colmax16_lanes8(m, lane);
        if (!(lane & 4)) sts_u32(s_part_addr + 4 * (q * C2 + c * 16 + my_col), __float_as_uint(r));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_z_empty);          // Z may be overwritten by the next tile's layer 2
      if (TIMING) tw[3] += (unsigned long long)(clock64() - tp0);
    };
    // second half of the pool job, off the tensor pipe's critical path (it runs after the next tile's acc1 conversion):
    // combine the two row-halves of each centroid, BN + ReLU, store
    auto pool_combine = [&](int tile) {
      const long long tp1 = TIMING ? clock64() : 0;
      const int wt = (int)tid_now() - 64;      // 0..383 among the workers
      const uint32_t s_sc2_addr = sb + OFF_SC + 2048, s_part_addr = sb + OFF_PART;
      asm volatile("bar.sync 2, 384;" ::: "memory");   // all partials written
      const long long tp2 = TIMING ? clock64() : 0;
#pragma unroll
      for (int u = 0; u < 2; ++u) {           // 512 pooled values (2 centroids x 256 channels) over 384 threads
        const int e = wt + 384 * u;
        if (e < 2 * C2) {
          const int g = e >> 8, c = e & (C2 - 1);
          const uint32_t grow = (uint32_t)tile * 2 + g;
          const float m = fmaxf(__uint_as_float(lds_u32(s_part_addr + 4 * ((2 * g) * C2 + c))),
                                __uint_as_float(lds_u32(s_part_addr + 4 * ((2 * g + 1) * C2 + c))));
          if (grow * 64u < a.rows) {
            const float y = fmaxf(fmaf(m, lds_f32(s_sc2_addr + 4 * c), lds_f32(s_sc2_addr + 4 * (256 + c))), 0.f);
            a.out[(size_t)grow * a.ld_out + c] = y;
            if (a.out_hi) {   // gather table of the next level's first layer (gemm_tc.cu GatherA)
              __nv_bfloat16 h, l;
              split_bf16(y, h, l);
              a.out_hi[(size_t)grow * C2 + c] = h;
              a.out_lo[(size_t)grow * C2 + c] = l;
            }
          }
        }
      }
      if (TIMING) {
        tw[4] += (unsigned long long)(tp2 - tp1);
        tw[5] += (unsigned long long)(clock64() - tp2);
      }
    };

    // The tile number is NOT kept in a register across the iteration (it was the value the 64-register budget spilled):
    // a worker releases ring slot `it` only after the combine of tile `it`, so the slot stays valid and the tile is simply
    // re-read from shared memory where it is needed.  The scheduler can still run 2 tiles ahead (= the A0 stages).
    const uint32_t ring_addr = smem_u32(const_cast<int*>(ring));
    // `opaque` stops the compiler from pre-computing (and then spilling) values derived from the loop counter
    auto opaque = [](uint32_t x) -> uint32_t {
      asm volatile("" : "+r"(x));
      return x;
    };
    auto tile_of = [&](uint32_t it) -> int { return (int)lds_u32(ring_addr + 4 * (opaque(it) & 3)); };
    auto finish_tile = [&](uint32_t it) {     // combine + store tile `it`, then let the scheduler recycle its slot
      pool_combine(tile_of(it));
      __syncwarp();
      if ((tid_now() & 31u) == 0) mbar_arrive(bar_sempty + 8 * (opaque(it) & 3));
    };
    bool have = take_tile(0) >= 0;
    if (have) {
      wait_t(bar_acc0, 0, 1);
      tc_fence_after();
      convert(false, dbg_row_of(tile_of(0), false));
    }
    uint32_t it = 0;
    for (; have; ++it) {
      wait_t(bar_acc1, it & 1, 1);          // layer 2 of this tile is waiting for A2: first
      tc_fence_after();
      convert(true, dbg_row_of(tile_of(it), true));
      if (it > 0) finish_tile(it - 1);      // (its partials were drained before; hidden under layer 2)
      const bool have_next = take_tile(it + 1) >= 0;
      if (have_next) {
        wait_t(bar_acc0, (it + 1) & 1, 1);
        tc_fence_after();
        convert(false, dbg_row_of(tile_of(it + 1), false));
      }
      wait_t(bar_acc2, it & 1, 6);
      tc_fence_after();
      pool_drain();                         // frees Z for the next tile's layer 2
      have = have_next;
    }
    if (it > 0) finish_tile(it - 1);
    if ((warp & 3) == 0) flush_t(2 + ((warp - 2) >> 2));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int sa0_chain_launch(const float* xyz, Strides3 xst, const float* new_xyz, const float* feat, int64_t feat_bstride,
                     int feat_ld, const int32_t* nbr, const float* W0, int ldw0, const float* scale0,
                     const float* shift0, const __nv_bfloat16* W1hi, const __nv_bfloat16* W1lo, int ldw1,
                     const float* scale1, const float* shift1, const __nv_bfloat16* W2hi, const __nv_bfloat16* W2lo,
                     int ldw2, const float* scale2, const float* shift2, int B, int M, float* out, int ld_out,
                     __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, float* dbg, unsigned int* tile_counter, int variant,
                     cudaStream_t stream) {
  const int64_t rows64 = (int64_t)B * M * 64;
  RN_CHECK_ARG(rows64 > 0 && rows64 < (1LL << 31), "sa0_chain: bad row count");
  RN_CHECK_ARG(ldw1 % 8 == 0 && ldw1 >= C1 && ldw2 % 8 == 0 && ldw2 >= C1 && ldw0 >= 6, "sa0_chain: bad leading dimensions");
  CUtensorMap m1h, m1l, m2h, m2l;
  RN_TRY(tc_make_map(&m1h, W1hi, C1, C1, ldw1, C1, BK, 128));
  RN_TRY(tc_make_map(&m1l, W1lo, C1, C1, ldw1, C1, BK, 128));
  RN_TRY(tc_make_map(&m2h, W2hi, C2, C1, ldw2, C2, BK, 128));
  RN_TRY(tc_make_map(&m2l, W2lo, C2, C1, ldw2, C2, BK, 128));
  int dev = 0, sms = 0;
  RN_CUDA(cudaGetDevice(&dev));
  RN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  Sa0ChainArgs a;
  a.xyz = xyz; a.xst = xst; a.new_xyz = new_xyz; a.feat = feat; a.feat_bstride = feat_bstride; a.feat_ld = feat_ld;
  a.nbr = nbr; a.W0 = W0; a.ldw0 = ldw0; a.scale0 = scale0; a.shift0 = shift0; a.scale1 = scale1; a.shift1 = shift1;
  a.scale2 = scale2; a.shift2 = shift2; a.out = out; a.ld_out = ld_out; a.out_hi = out_hi; a.out_lo = out_lo;
  a.dbg = dbg; a.M = (uint32_t)M;
  a.rows = (uint32_t)rows64; a.tile_counter = tile_counter; a.variant = variant;
  const int64_t n_tiles = (rows64 + BM - 1) / BM;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  // variant: 0 production; 1 = dbg receives the raw accumulators of layers 0 / 1, (rows, 256) fp32; 2 = dbg receives the
  // per-role wait counters, (grid, 5 roles, 8) u64
  a.timing = nullptr;
  if (variant == 2 && dbg) {
    a.timing = reinterpret_cast<unsigned long long*>(dbg);
    a.dbg = nullptr;
    RN_CUDA(cudaFuncSetAttribute(sa0_chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    sa0_chain_kernel<2><<<grid, NTHREADS, SMEM_BYTES, stream>>>(m1h, m1l, m2h, m2l, a);
  } else if (variant == 1 && dbg) {
    RN_CUDA(cudaFuncSetAttribute(sa0_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    sa0_chain_kernel<1><<<grid, NTHREADS, SMEM_BYTES, stream>>>(m1h, m1l, m2h, m2l, a);
  } else {
    a.dbg = nullptr;
    RN_CUDA(cudaFuncSetAttribute(sa0_chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    sa0_chain_kernel<0><<<grid, NTHREADS, SMEM_BYTES, stream>>>(m1h, m1l, m2h, m2l, a);
  }
  RN_LAUNCH_CHECK("sa0_chain_kernel");
  return REGNET_OK;
}

}  // namespace regnet
