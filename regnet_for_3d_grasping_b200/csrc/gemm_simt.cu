// fp32 FFMA engine for the shared-MLP layer (see gemm.cuh).  128x128x16 tiles, 256 threads, 8x8 micro-tiles,
// register-prefetched global loads, fused scale/shift/activation (+ 64-row max-pool) epilogue.
// This is the bring-up / cross-check engine; the product engine is gemm_tc.cu.  Both are CUDA.
#include "gemm.cuh"

namespace regnet {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int LDS_A = BM + 4, LDS_B = BN + 4;

__global__ void __launch_bounds__(NT, 2)
gemm_simt_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw, int64_t P, int kpad,
                 int cout, Epilogue ep) {
  __shared__ __align__(16) float As[BK][LDS_A];
  __shared__ __align__(16) float Bs[BK][LDS_B];

  const int t = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;
  const int lr = t >> 2;         // 0..63: row within half tile
  const int lk = (t & 3) << 2;   // 0,4,8,12
  const int ty = t >> 4, tx = t & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 pa[2], pb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = row0 + lr + h * 64;
      pa[h] = (r < P) ? *reinterpret_cast<const float4*>(X + r * ldx + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
      const int c = col0 + lr + h * 64;
      pb[h] = (c < cout) ? *reinterpret_cast<const float4*>(W + (int64_t)c * ldw + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + h * 64;
      As[lk + 0][r] = pa[h].x; As[lk + 1][r] = pa[h].y; As[lk + 2][r] = pa[h].z; As[lk + 3][r] = pa[h].w;
      Bs[lk + 0][r] = pb[h].x; Bs[lk + 1][r] = pb[h].y; Bs[lk + 2][r] = pb[h].z; Bs[lk + 3][r] = pb[h].w;
    }
  };

  gload(0);
  for (int k0 = 0; k0 < kpad; k0 += BK) {
    __syncthreads();
    sstore();
    __syncthreads();
    if (k0 + BK < kpad) gload(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

  // ---- epilogue --------------------------------------------------------------------------------------------
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = col0 + tx * 8 + j;
    sc[j] = (ep.scale && c < cout) ? ep.scale[c] : 1.f;
    sh[j] = (ep.shift && c < cout) ? ep.shift[c] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = apply_act(fmaf(acc[i][j], sc[j], sh[j]), ep.act);

  if (ep.pool == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t r = row0 + ty * 8 + i;
      if (r >= P) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = col0 + tx * 8 + j;
        if (c >= cout) continue;
        if (ep.out_f32) ep.out_f32[r * ep.ld_f32 + c] = acc[i][j];
        if (ep.out_hi) {
          __nv_bfloat16 h, l;
          split_bf16(acc[i][j], h, l);
          ep.out_hi[r * ep.ld_split + c] = h;
          ep.out_lo[r * ep.ld_split + c] = l;
        }
      }
    }
  } else {
    // pool == 64: the 128-row tile holds two groups; thread rows ty*8.. belong to group ty/8
    float* part = &As[0][0];  // 16 x 128 floats = 8 KB <= sizeof(As)
    static_assert(sizeof(float) * 16 * BN <= sizeof(float) * BK * LDS_A, "partial buffer fits in As");
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float m = acc[0][j];
#pragma unroll
      for (int i = 1; i < 8; ++i) m = fmaxf(m, acc[i][j]);
      part[ty * BN + tx * 8 + j] = m;
    }
    __syncthreads();
    const int g = t >> 7, c = t & 127;
    float m = part[(g * 8) * BN + c];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, part[(g * 8 + i) * BN + c]);
    const int64_t grow = row0 / 64 + g;
    if (grow * 64 < P && col0 + c < cout) ep.out_f32[grow * ep.ld_f32 + col0 + c] = m;
  }
}

}  // namespace

int gemm_simt_launch(const float* X, int ldx, const float* W, int ldw, int64_t P, int kpad, int cout,
                     const Epilogue& ep, cudaStream_t stream) {
  RN_CHECK_ARG(kpad % BK == 0 && ldx % 4 == 0 && ldw % 4 == 0 && ldx >= kpad && ldw >= kpad,
               "gemm_simt: K padding/leading dimensions must be multiples of 16/4 (kpad=%d ldx=%d ldw=%d)", kpad, ldx, ldw);
  RN_CHECK_ARG(ep.pool == 0 || (ep.pool == 64 && P % 64 == 0 && ep.out_f32 && !ep.out_hi),
               "gemm_simt: pooled epilogue needs pool == 64, P %% 64 == 0 and an fp32 output");
  if (P == 0) return REGNET_OK;
  dim3 grid((unsigned)((P + BM - 1) / BM), (unsigned)ceil_div(cout, BN));
  gemm_simt_kernel<<<grid, NT, 0, stream>>>(X, ldx, W, ldw, P, kpad, cout, ep);
  RN_LAUNCH_CHECK("gemm_simt_kernel");
  return REGNET_OK;
}

}  // namespace regnet
