// First two shared-MLP layers of set-abstraction level 0 in ONE kernel (sa_modules.0.mlp.{0,1}: 6 -> 128 -> 128).
//
// Reference chain (pn2_utils/modules.py:44-52 + nn/modules/conv.py:64-76, twice): group xyz / rgb, subtract the centroid,
// concat, conv 6->128 + BN + ReLU, conv 128->128 + BN + ReLU -- five kernels and 2 x 2.5 GB of activations for the
// B=15 batch.  Here the 128-wide activation of the first layer never leaves the SM:
//   warps 6..13  A-producers : gather the 6 input channels of a position (its neighbour index -> rgb, xyz - centroid),
//                              compute the 6->128 layer on the FP32 pipes, split to bf16 hi/lo and write the result
//                              straight into shared memory in the K-major SWIZZLE_128B layout tcgen05 reads
//   warp 1       MMA issuer  : the 128->128 layer as tcgen05.mma (3 products hi*hi + lo*hi + hi*lo, fp32 in TMEM);
//                              its weights (64 KB as bf16 hi/lo) are loaded once per CTA by TMA and stay resident
//   warps 2..5   epilogue    : TMEM -> BN/ReLU -> bf16 hi/lo -> swizzled staging -> TMA store (operand of layer 3)
//   warp 0       scheduler   : draws tiles from a global counter, publishes them through a shared-memory ring
// A-tiles are double buffered (2 x 64 KB), accumulators too, so gather/FFMA, tensor work and stores overlap.
#include <cuda.h>

#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace regnet {

using namespace tc;

namespace {

constexpr int BM = 128, CH = 128, BK = 64;
constexpr int PLANE = BM * BK * 2;               // 16 KB: one bf16 plane of one k-block
constexpr int OFF_W = 0;                         // W1: [kb][hi|lo] = 4 planes
constexpr int OFF_A = 4 * PLANE;                 // A stages: 2 x [kb][hi|lo]
constexpr int OFF_OUT = OFF_A + 2 * 4 * PLANE;   // epilogue staging 4 warps x {hi,lo} x 2 KB
constexpr int OFF_L0W = OFF_OUT + 16384;         // layer-0 weights packed [128][8] floats: w0..w5, scale, shift
constexpr int OFF_SC = OFF_L0W + CH * 8 * 4;     // layer-1 scale[128], shift[128]
constexpr int OFF_BAR = OFF_SC + 2 * CH * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int N_PROD_WARPS = 8;
constexpr int NTHREADS = 32 * (6 + N_PROD_WARPS);

// 64 registers x 448 threads: leaves the register file room for the co-resident FPS CTA of the next step
__global__ void __maxnreg__(64)
sa0_front_kernel(const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
                 const __grid_constant__ CUtensorMap map_ohi, const __grid_constant__ CUtensorMap map_olo,
                 const float* __restrict__ xyz, Strides3 xst, const float* __restrict__ new_xyz,
                 const float* __restrict__ feat, int64_t feat_bstride, int feat_ld, const int32_t* __restrict__ nbr,
                 const float* __restrict__ W0, int ldw0, const float* __restrict__ scale0, const float* __restrict__ shift0,
                 const float* __restrict__ scale1, const float* __restrict__ shift1, uint32_t M, uint32_t rows,
                 unsigned int* tile_counter) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar_w = sb + OFF_BAR;            // weights landed
  const uint32_t bar_afull = bar_w + 8;           // [2] producers -> MMA
  const uint32_t bar_aempty = bar_afull + 16;     // [2] MMA -> producers
  const uint32_t bar_tfull = bar_aempty + 16;     // [2] MMA -> epilogue
  const uint32_t bar_tempty = bar_tfull + 16;     // [2] epilogue -> MMA
  const uint32_t bar_sfull = bar_tempty + 16;     // [4] scheduler ring
  const uint32_t bar_sempty = bar_sfull + 32;     // [4]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * 17);
  volatile int* ring = reinterpret_cast<volatile int*>(tmem_holder + 1);
  float4* l0w = reinterpret_cast<float4*>(smem + OFF_L0W);
  float* s_scale = reinterpret_cast<float*>(smem + OFF_SC);
  float* s_shift = s_scale + CH;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_tiles = (rows + BM - 1) / BM;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_afull + 8 * s, N_PROD_WARPS * 32);
      mbar_init(bar_aempty + 8 * s, 1);
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 4);
    }
    for (int r = 0; r < 4; ++r) {
      mbar_init(bar_sfull + 8 * r, 1);
      mbar_init(bar_sempty + 8 * r, N_PROD_WARPS + 1 + 4);   // one lane per producer / epilogue warp + the MMA lane
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_whi);
    tma_prefetch_desc(&map_wlo);
  }
  // layer-0 weights and layer-1 scale/shift into shared memory (every thread helps)
  for (int c = threadIdx.x; c < CH; c += NTHREADS) {
    l0w[2 * c] = make_float4(W0[c * ldw0 + 0], W0[c * ldw0 + 1], W0[c * ldw0 + 2], W0[c * ldw0 + 3]);
    l0w[2 * c + 1] = make_float4(W0[c * ldw0 + 4], W0[c * ldw0 + 5], scale0[c], shift0[c]);
    s_scale[c] = scale1[c];
    s_shift[c] = shift1[c];
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), 256);
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
      mbar_arrive_expect_tx(bar_w, 4 * PLANE);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_2d(sb + OFF_W + kb * 2 * PLANE, &map_whi, bar_w, kb * BK, 0);
        tma_load_2d(sb + OFF_W + kb * 2 * PLANE + PLANE, &map_wlo, bar_w, kb * BK, 0);
      }
      for (uint32_t it = 0;; ++it) {
        const uint32_t slot = it & 3;
        if (it >= 4) mbar_wait(bar_sempty + 8 * slot, ((it >> 2) - 1) & 1);
        const uint32_t t = tile_counter ? atomicAdd(tile_counter, 1u) : blockIdx.x + it * gridDim.x;
        ring[slot] = t < n_tiles ? (int)t : -1;
        mbar_arrive(bar_sfull + 8 * slot);
        if (t >= n_tiles) break;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, CH);
      mbar_wait(bar_w, 0);
      for (uint32_t it = 0;; ++it) {
        const int tile = take_tile(it);
        mbar_arrive(bar_sempty + 8 * (it & 3));
        if (tile < 0) break;
        const uint32_t s = it & 1, ph = (it >> 1) & 1;
        mbar_wait(bar_tempty + 8 * s, ph ^ 1);
        mbar_wait(bar_afull + 8 * s, ph);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + s * CH;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t a0 = sb + OFF_A + s * 4 * PLANE + kb * 2 * PLANE, w0 = sb + OFF_W + kb * 2 * PLANE;
          const uint64_t a_hi = make_sdesc(a0), a_lo = make_sdesc(a0 + PLANE);
          const uint64_t b_hi = make_sdesc(w0), b_lo = make_sdesc(w0 + PLANE);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb | k) != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
        }
        umma_commit(bar_aempty + 8 * s);
        umma_commit(bar_tfull + 8 * s);
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ================= epilogue: layer-1 BN + ReLU, bf16 hi/lo planes out through TMA =================
    const int q = warp & 3;
    const uint32_t st_hi = sb + OFF_OUT + (warp - 2) * 4096, st_lo = st_hi + 2048;
    for (uint32_t it = 0;; ++it) {
      const int tile = take_tile(it);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sempty + 8 * (it & 3));
      if (tile < 0) break;
      const uint32_t s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(bar_tfull + 8 * s, ph);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < CH / 32; ++ch) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + s * CH + ch * 32 + half * 16, v);
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = ch * 32 + half * 16 + 2 * j;
            const float y0 = fmaxf(fmaf(__uint_as_float(v[2 * j]), s_scale[c], s_shift[c]), 0.f);
            const float y1 = fmaxf(fmaf(__uint_as_float(v[2 * j + 1]), s_scale[c + 1], s_shift[c + 1]), 0.f);
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(y0, h0, l0);
            split_bf16(y1, h1, l1);
            __nv_bfloat162 hh = __halves2bfloat162(h0, h1), ll = __halves2bfloat162(l0, l1);
            hi[j] = *reinterpret_cast<uint32_t*>(&hh);
            lo[j] = *reinterpret_cast<uint32_t*>(&ll);
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
            if (lane == 0) {
              tma_store_2d(&map_ohi, st_hi, ch * 32, tile * BM + q * 32);
              tma_store_2d(&map_olo, st_lo, ch * 32, tile * BM + q * 32);
              bulk_commit();
            }
          }
        }
      }
      tc_fence_before();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * s);
    }
    if (lane == 0) bulk_wait_all();
    __syncwarp();
  } else {
    // ================= A-producers: gather + layer 0 (6 -> 128) on the FP32 pipes =================
    const int pt = threadIdx.x - 6 * 32;      // 0..255
    const int r = pt & 127;                   // row of the tile
    const int h = pt >> 7;                    // which 64-channel half == which k-block of the A tile
    auto gather = [&](int tile, float (&v)[6]) {
      const uint32_t row = (uint32_t)tile * BM + r;
      if (row < rows) {
        const uint32_t bm = row >> 6;         // 64 neighbours per centroid
        const uint32_t b = bm / M, m = bm - b * M;
        const int j = nbr[row];
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = feat[(int64_t)b * feat_bstride + (int64_t)j * feat_ld + c];
#pragma unroll
        for (int a = 0; a < 3; ++a)
          v[3 + a] = __fsub_rn(xyz[(int64_t)b * xst.b + a * xst.c + (int64_t)j * xst.n],
                               new_xyz[((int64_t)b * 3 + a) * M + m]);
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) v[c] = 0.f;
      }
    };
    float vn[6];
    int tile = take_tile(0);
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_sempty);
    if (tile >= 0) gather(tile, vn);
    for (uint32_t it = 0; tile >= 0; ++it) {
      float v[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) v[c] = vn[c];
      const int next = take_tile(it + 1);     // the ring is at most 4 ahead; taking it early lets the next gather fly
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sempty + 8 * ((it + 1) & 3));
      if (next >= 0) gather(next, vn);
      const uint32_t s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(bar_aempty + 8 * s, ph ^ 1);
      const uint32_t a_hi = sb + OFF_A + s * 4 * PLANE + h * 2 * PLANE, a_lo = a_hi + PLANE;
#pragma unroll 2
      for (int chunk = 0; chunk < 8; ++chunk) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float y[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = h * 64 + chunk * 8 + 2 * u + e;
            const float4 wa = l0w[2 * c], wb = l0w[2 * c + 1];
            float acc = fmaf(wa.x, v[0], 0.f);
            acc = fmaf(wa.y, v[1], acc);
            acc = fmaf(wa.z, v[2], acc);
            acc = fmaf(wa.w, v[3], acc);
            acc = fmaf(wb.x, v[4], acc);
            acc = fmaf(wb.y, v[5], acc);
            y[e] = fmaxf(fmaf(acc, wb.z, wb.w), 0.f);
          }
          __nv_bfloat16 h0, l0, h1, l1;
          split_bf16(y[0], h0, l0);
          split_bf16(y[1], h1, l1);
          __nv_bfloat162 hh = __halves2bfloat162(h0, h1), ll = __halves2bfloat162(l0, l1);
          hi[u] = *reinterpret_cast<uint32_t*>(&hh);
          lo[u] = *reinterpret_cast<uint32_t*>(&ll);
        }
        // K-major SWIZZLE_128B: row r at r*128 bytes, 16-byte chunk index XOR (r mod 8)
        const uint32_t off = (uint32_t)r * 128u + (((uint32_t)chunk ^ ((uint32_t)r & 7u)) << 4);
        sts_v4(a_hi + off, hi[0], hi[1], hi[2], hi[3]);
        sts_v4(a_lo + off, lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core's async proxy
      mbar_arrive(bar_afull + 8 * s);
      tile = next;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

int sa0_front_launch(const float* xyz, Strides3 xst, const float* new_xyz, const float* feat, int64_t feat_bstride,
                     int feat_ld, const int32_t* nbr, const float* W0, int ldw0, const float* scale0,
                     const float* shift0, const __nv_bfloat16* W1hi, const __nv_bfloat16* W1lo, int ldw1,
                     const float* scale1, const float* shift1, int B, int M, __nv_bfloat16* out_hi,
                     __nv_bfloat16* out_lo, int ld_out, unsigned int* tile_counter, cudaStream_t stream) {
  const int64_t rows64 = (int64_t)B * M * 64;
  RN_CHECK_ARG(rows64 < (1LL << 31) && ld_out % 8 == 0 && ldw1 % 8 == 0 && ldw1 >= CH, "sa0_front: bad sizes");
  CUtensorMap mwh, mwl, moh, mol;
  RN_TRY(tc_make_map(&mwh, W1hi, CH, CH, ldw1, CH, BK, 128));
  RN_TRY(tc_make_map(&mwl, W1lo, CH, CH, ldw1, CH, BK, 128));
  RN_TRY(tc_make_map(&moh, out_hi, rows64, CH, ld_out, 32, 32, 64));
  RN_TRY(tc_make_map(&mol, out_lo, rows64, CH, ld_out, 32, 32, 64));
  int dev = 0, sms = 0;
  RN_CUDA(cudaGetDevice(&dev));
  RN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  RN_CUDA(cudaFuncSetAttribute(sa0_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const int64_t n_tiles = (rows64 + BM - 1) / BM;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  sa0_front_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(mwh, mwl, moh, mol, xyz, xst, new_xyz, feat, feat_bstride,
                                                          feat_ld, nbr, W0, ldw0, scale0, shift0, scale1, shift1,
                                                          (uint32_t)M, (uint32_t)rows64, tile_counter);
  RN_LAUNCH_CHECK("sa0_front_kernel");
  return REGNET_OK;
}

}  // namespace regnet
