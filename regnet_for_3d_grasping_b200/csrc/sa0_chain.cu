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
//   warp 0       scheduler + one-time TMA load of W1 / W2
//   warp 1       MMA issuer (one lane); issue order L1(i), L0(i+1), L2(i) keeps the tensor pipe busy while the
//                converters work on the other region; hazards on X / Y are ordered by the pipe itself
//   warps 2..5   converters: acc0 -> A1, acc1 -> A2, in four 32-channel chunks, each chunk released to the MMA lane
//                as soon as it is written (the next layer starts after a quarter of the conversion)
//   warps 6..9   pool epilogue: acc2 -> BN + ReLU -> max over the 64 rows of each centroid -> (B*M, 256) fp32
//   warps 10..13 producers: neighbour index -> rgb, xyz - centroid -> bf16 hi/lo rows of A0 (prefetched one tile ahead)
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
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
constexpr int NTHREADS = 448;
constexpr uint32_t TM_X = 0, TM_Y = 128, TM_Z = 256;
static_assert(SMEM_BYTES <= 232448, "sa0_chain: shared memory budget");

// K-major operand with 32-byte rows (K = 16 bf16).  variant bit 0 = 0: SWIZZLE_32B canonical layout
// ((8,n),2):((2,SBO),1) in 16-byte units, 16-byte chunk index XOR bit 2 of the row; = 1: no swizzle ("interleave"),
// core matrices of 8 rows x 16 bytes, the two K chunks LBO = 128 bytes apart, 8-row groups SBO = 256 bytes apart.
__device__ __forceinline__ uint32_t k16_offset(uint32_t row, uint32_t chunk, int noswz) {
  if (noswz) return (row >> 3) * 256u + chunk * 128u + (row & 7u) * 16u;
  return row * 32u + ((chunk ^ ((row >> 2) & 1u)) << 4);
}
__device__ __forceinline__ uint64_t make_sdesc_k16(uint32_t smem_addr, int noswz) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)(noswz ? (128 >> 4) : 1) << 16;   // LBO
  d |= (uint64_t)(256 >> 4) << 32;                 // SBO
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(noswz ? 0 : 6) << 61;            // SWIZZLE_NONE / SWIZZLE_32B
  return d;
}

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

// two fp32 -> one packed bf16x2 hi word + one lo word (same rounding as split_bf16: hi = rn(x), lo = rn(x - hi))
__device__ __forceinline__ void split_pair(float y0, float y1, uint32_t& hi, uint32_t& lo, int swap) {
  if (swap) { const float t = y0; y0 = y1; y1 = t; }
  const __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);   // .x (low half) = y0
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(__fsub_rn(y0, h0), __fsub_rn(y1, h1));
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

struct Sa0ChainArgs {
  const float* xyz; Strides3 xst;
  const float* new_xyz;
  const float* feat; int64_t feat_bstride; int feat_ld;
  const int32_t* nbr;
  const float* W0; int ldw0;                       // (128, >= 6) fp32, operand order [feature(3) | xyz_rel(3)]
  const float* scale0; const float* shift0;
  const float* scale1; const float* shift1;
  const float* scale2; const float* shift2;
  float* out; int ld_out;                          // (rows / 64, 256)
  float* dbg;                                      // optional (rows, 256): raw acc0 | raw acc1
  uint32_t M, rows;
  unsigned int* tile_counter;
  int variant;
};

__global__ void __maxnreg__(64)
sa0_chain_kernel(const __grid_constant__ CUtensorMap map_w1hi, const __grid_constant__ CUtensorMap map_w1lo,
                 const __grid_constant__ CUtensorMap map_w2hi, const __grid_constant__ CUtensorMap map_w2lo,
                 const Sa0ChainArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar_w = sb + OFF_BAR;            // W1 / W2 landed
  const uint32_t bar_a0full = bar_w + 8;          // [2] producers -> MMA
  const uint32_t bar_a0empty = bar_a0full + 16;   // [2] MMA -> producers
  const uint32_t bar_acc0 = bar_a0empty + 16;     // L0 complete -> converters
  const uint32_t bar_acc1 = bar_acc0 + 8;         // L1 complete -> converters
  const uint32_t bar_acc2 = bar_acc1 + 8;         // L2 complete -> pool warps
  const uint32_t bar_z_empty = bar_acc2 + 8;      // pool warps -> MMA
  const uint32_t bar_a1 = bar_z_empty + 8;        // [4] converters -> MMA, per 32-channel chunk of A1
  const uint32_t bar_a2 = bar_a1 + 32;            // [4] same for A2
  const uint32_t bar_sfull = bar_a2 + 32;         // [4] scheduler ring
  const uint32_t bar_sempty = bar_sfull + 32;     // [4]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * 26);
  volatile int* ring = reinterpret_cast<volatile int*>(tmem_holder + 1);
  float* s_sc = reinterpret_cast<float*>(smem + OFF_SC);   // [0,128) scale0 [128,256) shift0 [256,384) scale1 [384,512) shift1
  float* s_sc2 = s_sc + 512;                               // [0,256) scale2 [256,512) shift2
  uint32_t* s_part = reinterpret_cast<uint32_t*>(smem + OFF_PART);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_tiles = (a.rows + BM - 1) / BM;
  const int noswz = a.variant & 1, swap = (a.variant >> 1) & 1;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_a0full + 8 * s, 128);
      mbar_init(bar_a0empty + 8 * s, 1);
    }
    mbar_init(bar_acc0, 1);
    mbar_init(bar_acc1, 1);
    mbar_init(bar_acc2, 1);
    mbar_init(bar_z_empty, 4);
    for (int c = 0; c < 4; ++c) {
      mbar_init(bar_a1 + 8 * c, 128);
      mbar_init(bar_a2 + 8 * c, 128);
      mbar_init(bar_sfull + 8 * c, 1);
      mbar_init(bar_sempty + 8 * c, 13);   // MMA lane + one lane of each converter / pool / producer warp
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
      split_pair(a.W0[c * a.ldw0 + 2 * j], a.W0[c * a.ldw0 + 2 * j + 1], hi[j], lo[j], 0);
    const uint32_t off = k16_offset((uint32_t)c, 0, noswz);
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

  auto take_tile = [&](uint32_t it) -> int {
    const uint32_t slot = it & 3;
    mbar_wait(bar_sfull + 8 * slot, (it >> 2) & 1);
    return ring[slot];
  };

  if (warp == 0) {
    // ================= scheduler + one-time weight load =================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, 4 * W1_PLANE + 4 * W2_PLANE);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_2d(sb + OFF_W1 + kb * 2 * W1_PLANE, &map_w1hi, bar_w, kb * BK, 0);
        tma_load_2d(sb + OFF_W1 + kb * 2 * W1_PLANE + W1_PLANE, &map_w1lo, bar_w, kb * BK, 0);
        tma_load_2d(sb + OFF_W2 + kb * 2 * W2_PLANE, &map_w2hi, bar_w, kb * BK, 0);
        tma_load_2d(sb + OFF_W2 + kb * 2 * W2_PLANE + W2_PLANE, &map_w2lo, bar_w, kb * BK, 0);
      }
      for (uint32_t it = 0;; ++it) {
        const uint32_t slot = it & 3;
        if (it >= 4) mbar_wait(bar_sempty + 8 * slot, ((it >> 2) - 1) & 1);
        const uint32_t t = a.tile_counter ? atomicAdd(a.tile_counter, 1u) : blockIdx.x + it * gridDim.x;
        ring[slot] = t < n_tiles ? (int)t : -1;
        mbar_arrive(bar_sfull + 8 * slot);
        if (t >= n_tiles) break;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(BM, C1), idesc2 = make_idesc(BM, C2);
      const uint64_t w0_hi = make_sdesc_k16(sb + OFF_W0, noswz), w0_lo = make_sdesc_k16(sb + OFF_W0 + A0_PLANE, noswz);
      auto issue_l0 = [&](uint32_t it) {
        const uint32_t s = it & 1, ph = (it >> 1) & 1;
        mbar_wait(bar_a0full + 8 * s, ph);
        tc_fence_after();
        const uint32_t a0 = sb + OFF_A0 + s * 2 * A0_PLANE;
        const uint64_t a_hi = make_sdesc_k16(a0, noswz), a_lo = make_sdesc_k16(a0 + A0_PLANE, noswz);
        umma_f16(tmem_base + TM_X, a_hi, w0_hi, idesc1, 0);
        umma_f16(tmem_base + TM_X, a_lo, w0_hi, idesc1, 1);
        umma_f16(tmem_base + TM_X, a_hi, w0_lo, idesc1, 1);
        umma_commit(bar_a0empty + 8 * s);
        umma_commit(bar_acc0);
      };
      // one 32-channel chunk (2 k-steps) of a TS layer: A hi/lo at columns 16*ks / 16*ks + 8 of `a_region`
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
        mbar_wait(bar_w, 0);
      }
      for (uint32_t it = 0; tile >= 0; ++it) {
        const uint32_t ph = it & 1;
        for (int ch = 0; ch < 4; ++ch) {
          mbar_wait(bar_a1 + 8 * ch, ph);
          tc_fence_after();
          issue_chunk(tmem_base + TM_Y, tmem_base + TM_X, sb + OFF_W1, W1_PLANE, idesc1, ch);
        }
        umma_commit(bar_acc1);
        const int next = take_tile(it + 1);
        mbar_arrive(bar_sempty + 8 * ((it + 1) & 3));
        if (next >= 0) issue_l0(it + 1);   // overwrites X after L1(it) in pipe order
        if (it > 0) {
          mbar_wait(bar_z_empty, (it - 1) & 1);
          tc_fence_after();
        }
        for (int ch = 0; ch < 4; ++ch) {
          mbar_wait(bar_a2 + 8 * ch, ph);
          tc_fence_after();
          issue_chunk(tmem_base + TM_Z, tmem_base + TM_Y, sb + OFF_W2, W2_PLANE, idesc2, ch);
        }
        umma_commit(bar_acc2);
        tile = next;
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ================= converters: accumulator -> BN + ReLU -> bf16 hi/lo A operand, in place in TMEM =================
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    auto convert = [&](uint32_t region, const float* sc, const float* sh, uint32_t bar_ready, float* dbg_row) {
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ks = 2 * ch + h;
          uint32_t v[16], o[16];
          tmem_ld16(lane_base + region + 16 * ks, v);
          if (dbg_row) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dbg_row[16 * ks + j] = __uint_as_float(v[j]);
          }
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 s4 = *reinterpret_cast<const float4*>(sc + 16 * ks + 4 * j4);
            const float4 t4 = *reinterpret_cast<const float4*>(sh + 16 * ks + 4 * j4);
            const float y0 = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 0]), s4.x, t4.x), 0.f);
            const float y1 = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 1]), s4.y, t4.y), 0.f);
            const float y2 = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 2]), s4.z, t4.z), 0.f);
            const float y3 = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 3]), s4.w, t4.w), 0.f);
            split_pair(y0, y1, o[2 * j4], o[8 + 2 * j4], swap);
            split_pair(y2, y3, o[2 * j4 + 1], o[8 + 2 * j4 + 1], swap);
          }
          tmem_st16(lane_base + region + 16 * ks, o);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_ready + 8 * ch);
      }
    };
    for (uint32_t it = 0;; ++it) {
      const int tile = take_tile(it);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sempty + 8 * (it & 3));
      if (tile < 0) break;
      const uint32_t ph = it & 1;
      const uint32_t row = (uint32_t)tile * BM + (warp & 3) * 32 + lane;
      float* dbg_row = (a.dbg && row < a.rows) ? a.dbg + (size_t)row * 256 : nullptr;
      mbar_wait(bar_acc0, ph);
      tc_fence_after();
      convert(TM_X, s_sc, s_sc + 128, bar_a1, dbg_row);
      mbar_wait(bar_acc1, ph);
      tc_fence_after();
      convert(TM_Y, s_sc + 256, s_sc + 384, bar_a2, dbg_row ? dbg_row + 128 : nullptr);
    }
  } else if (warp < 10) {
    // ================= pool epilogue: acc2 -> BN + ReLU -> max over each centroid's 64 rows =================
    const int q = warp & 3;
    const int et = threadIdx.x - 6 * 32;   // 0..127
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + TM_Z;
    for (uint32_t it = 0;; ++it) {
      const int tile = take_tile(it);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sempty + 8 * (it & 3));
      if (tile < 0) break;
      mbar_wait(bar_acc2, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < C2 / 32; ++ch) {
        uint32_t mine = 0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[16];
          tmem_ld16(lane_base + ch * 32 + half * 16, v);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 s4 = *reinterpret_cast<const float4*>(s_sc2 + ch * 32 + half * 16 + 4 * j4);
            const float4 t4 = *reinterpret_cast<const float4*>(s_sc2 + 256 + ch * 32 + half * 16 + 4 * j4);
            const float y[4] = {fmaf(__uint_as_float(v[4 * j4 + 0]), s4.x, t4.x), fmaf(__uint_as_float(v[4 * j4 + 1]), s4.y, t4.y),
                                fmaf(__uint_as_float(v[4 * j4 + 2]), s4.z, t4.z), fmaf(__uint_as_float(v[4 * j4 + 3]), s4.w, t4.w)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // ReLU and the float order in one integer max: negative floats are negative ints, and non-negative
              // floats order like their bit patterns
              const int yi = max(__float_as_int(y[j]), 0);
              const int m = __reduce_max_sync(FULL, yi);
              if (lane == half * 16 + 4 * j4 + j) mine = (uint32_t)m;
            }
          }
        }
        s_part[q * C2 + ch * 32 + lane] = mine;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_z_empty);
      asm volatile("bar.sync 2, 128;" ::: "memory");
      for (int e = et; e < 2 * C2; e += 128) {
        const int g = e / C2, c = e - g * C2;
        const uint32_t grow = (uint32_t)tile * 2 + g;
        if (grow * 64u < a.rows) {
          const uint32_t m = max(s_part[(2 * g) * C2 + c], s_part[(2 * g + 1) * C2 + c]);
          a.out[(size_t)grow * a.ld_out + c] = __uint_as_float(m);
        }
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
    }
  } else {
    // ================= producers: gather the 6 input channels of a row, bf16 hi/lo, into the A0 stage =================
    const uint32_t r = threadIdx.x - 10 * 32;   // row of the tile
    auto gather = [&](int tile, float (&v)[6]) {
      const uint32_t row = (uint32_t)tile * BM + r;
      if (row < a.rows) {
        const uint32_t bm = row >> 6;         // 64 neighbours per centroid
        const uint32_t b = bm / a.M, m = bm - b * a.M;
        const int j = a.nbr[row];
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = a.feat[(int64_t)b * a.feat_bstride + (int64_t)j * a.feat_ld + c];
#pragma unroll
        for (int x = 0; x < 3; ++x)
          v[3 + x] = __fsub_rn(a.xyz[(int64_t)b * a.xst.b + x * a.xst.c + (int64_t)j * a.xst.n],
                               a.new_xyz[((int64_t)b * 3 + x) * a.M + m]);
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) v[c] = 0.f;
      }
    };
    const uint32_t off = k16_offset(r, 0, noswz);
    float vn[6];
    int tile = take_tile(0);
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_sempty);
    if (tile >= 0) gather(tile, vn);
    for (uint32_t it = 0; tile >= 0; ++it) {
      uint32_t hi[3], lo[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) split_pair(vn[2 * j], vn[2 * j + 1], hi[j], lo[j], 0);
      const int next = take_tile(it + 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sempty + 8 * ((it + 1) & 3));
      if (next >= 0) gather(next, vn);
      const uint32_t s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(bar_a0empty + 8 * s, ph ^ 1);
      const uint32_t a0 = sb + OFF_A0 + s * 2 * A0_PLANE + off;
      sts_v4(a0, hi[0], hi[1], hi[2], 0u);
      sts_v4(a0 + A0_PLANE, lo[0], lo[1], lo[2], 0u);
      fence_proxy_async();
      mbar_arrive(bar_a0full + 8 * s);
      tile = next;
    }
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
                     float* dbg, unsigned int* tile_counter, int variant, cudaStream_t stream) {
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
  RN_CUDA(cudaFuncSetAttribute(sa0_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  Sa0ChainArgs a;
  a.xyz = xyz; a.xst = xst; a.new_xyz = new_xyz; a.feat = feat; a.feat_bstride = feat_bstride; a.feat_ld = feat_ld;
  a.nbr = nbr; a.W0 = W0; a.ldw0 = ldw0; a.scale0 = scale0; a.shift0 = shift0; a.scale1 = scale1; a.shift1 = shift1;
  a.scale2 = scale2; a.shift2 = shift2; a.out = out; a.ld_out = ld_out; a.dbg = dbg; a.M = (uint32_t)M;
  a.rows = (uint32_t)rows64; a.tile_counter = tile_counter; a.variant = variant;
  const int64_t n_tiles = (rows64 + BM - 1) / BM;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  sa0_chain_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(m1h, m1l, m2h, m2l, a);
  RN_LAUNCH_CHECK("sa0_chain_kernel");
  return REGNET_OK;
}

}  // namespace regnet
