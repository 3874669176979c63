// Operand producers of the training path: the grouped / interpolated input of a shared MLP written DIRECTLY as the bf16
// hi/lo planes the tensor-core convolutions read (conv_train.cu), in torch's channel-major layout.
//
//   sa_group_planes    QueryGrouper.forward of pn2_utils/modules.py:39-56: group_points(xyz) - new_xyz, group_points(feature),
//                      concat [xyz_rel | feature] along the channels -> planes (B, 3 + C, M, K).  Replaces two gathers, a
//                      subtraction, torch.cat and the split pass (five passes over the 1.3 GB level-1 operand).
//   fp_interp_planes   FeatureInterpolator.forward of modules.py:104-131: 3-NN weighted interpolation of the sparse features,
//                      concat [interpolated | dense] -> planes (B, C2 + C1, Nd).
//   *_backward_strided the scatter-adds of grouping_kernel.cu:54-149 / interpolate_kernel.cu:239-337 reading a channel
//                      slice of the (B, Ctot, L) input gradient in place (no .contiguous() copy of the slice).
#include <algorithm>
#include <cstdlib>

#include "internal.cuh"
#include "train_common.cuh"

namespace regnet {

namespace {

constexpr int TG = 256;

// one thread = four consecutive positions of one (b, c) row
__global__ void __launch_bounds__(TG)
sa_group_planes_kernel(const float* __restrict__ xyz, Strides3 xst, const float* __restrict__ new_xyz,
                       const float* __restrict__ feat, Strides3 fst, const int64_t* __restrict__ index, int C, int N, int M,
                       int K, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int* __restrict__ oob) {
  const int c = blockIdx.y, b = blockIdx.z;
  const int64_t MK = (int64_t)M * K;
  const float* __restrict__ src = c < 3 ? xyz + (int64_t)b * xst.b + (int64_t)c * xst.c
                                        : feat + (int64_t)b * fst.b + (int64_t)(c - 3) * fst.c;
  const int64_t sn = c < 3 ? xst.n : fst.n;
  const int64_t* __restrict__ idx = index + (int64_t)b * MK;
  const float* __restrict__ ctr = new_xyz + ((int64_t)b * 3 + c) * M;       // only read when c < 3
  const int64_t out0 = ((int64_t)b * (C + 3) + c) * MK;
  for (int64_t e = ((int64_t)blockIdx.x * TG + threadIdx.x) * 4; e < MK; e += (int64_t)gridDim.x * TG * 4) {
    const longlong2 j01 = *reinterpret_cast<const longlong2*>(idx + e), j23 = *reinterpret_cast<const longlong2*>(idx + e + 2);
    const int64_t j[4] = {j01.x, j01.y, j23.x, j23.y};
    float v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (j[t] < 0 || j[t] >= N) { *oob = 1; v[t] = 0.f; continue; }
      v[t] = src[j[t] * sn];
    }
    if (c < 3) {   // K is a multiple of 4: the four positions share one centroid
      const float cv = ctr[e / K];
#pragma unroll
      for (int t = 0; t < 4; ++t) v[t] = __fsub_rn(v[t], cv);
    }
    store_planes4(hi, lo, out0 + e, make_float4(v[0], v[1], v[2], v[3]));
  }
}

__global__ void __launch_bounds__(TG)
fp_interp_planes_kernel(const float* __restrict__ sparse, Strides3 sst, const float* __restrict__ dense, Strides3 dst,
                        const int64_t* __restrict__ index, const float* __restrict__ weight, int C2, int C1, int Ns, int Nd,
                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int* __restrict__ oob) {
  // grid.y = C2 + C1 (one channel per row) or C2 / 4 + C1 (four interpolated channels per row, then the dense channels)
  const bool four = (int)gridDim.y != C2 + C1;
  const int c = four ? ((int)blockIdx.y < C2 / 4 ? (int)blockIdx.y * 4 : (int)blockIdx.y - C2 / 4 + C2) : (int)blockIdx.y;
  const int b = blockIdx.z;
  const int64_t out0 = ((int64_t)b * (C2 + C1) + c) * Nd;
  if (c >= C2) {   // dense (skip) channels: a strided copy
    const float* __restrict__ src = dense + (int64_t)b * dst.b + (int64_t)(c - C2) * dst.c;
    for (int64_t n = ((int64_t)blockIdx.x * TG + threadIdx.x) * 4; n < Nd; n += (int64_t)gridDim.x * TG * 4) {
      float4 v;
      if (dst.n == 1 && ((reinterpret_cast<uintptr_t>(src + n) & 15) == 0)) v = *reinterpret_cast<const float4*>(src + n);
      else v = make_float4(src[n * dst.n], src[(n + 1) * dst.n], src[(n + 2) * dst.n], src[(n + 3) * dst.n]);
      store_planes4(hi, lo, out0 + n, v);
    }
    return;
  }
  const float* __restrict__ src = sparse + (int64_t)b * sst.b + (int64_t)c * sst.c;
  const int64_t* __restrict__ idx = index + (int64_t)b * Nd * 3;
  const float* __restrict__ w = weight + (int64_t)b * Nd * 3;
  // CPB interpolated channels per CTA row (host: 4 when C2 % 4 == 0, else 1): the 36 bytes of index and weight per dense
  // point are read once for all of them instead of once per channel
  const int cpb = (int)gridDim.y == C2 + C1 ? 1 : 4;
  if (cpb == 4) {
    const int c4 = blockIdx.y * 4;           // the host lays the 4-channel rows first: blockIdx.y < C2 / 4
    const float* __restrict__ s4 = sparse + (int64_t)b * sst.b + (int64_t)c4 * sst.c;
    const int64_t o4 = ((int64_t)b * (C2 + C1) + c4) * Nd;
    for (int64_t n = ((int64_t)blockIdx.x * TG + threadIdx.x) * 4; n < Nd; n += (int64_t)gridDim.x * TG * 4) {
      float v[4][4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int64_t o = (n + t) * 3;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int64_t j = idx[o + k];
          if (j < 0 || j >= Ns) { *oob = 1; continue; }
          const float wk = w[o + k];
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) acc[ch] = fmaf(s4[(int64_t)ch * sst.c + j * sst.n], wk, acc[ch]);
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) v[ch][t] = acc[ch];
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        store_planes4(hi, lo, o4 + (int64_t)ch * Nd + n, make_float4(v[ch][0], v[ch][1], v[ch][2], v[ch][3]));
    }
    return;
  }
  for (int64_t n = ((int64_t)blockIdx.x * TG + threadIdx.x) * 4; n < Nd; n += (int64_t)gridDim.x * TG * 4) {
    float v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int64_t o = (n + t) * 3;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {   // the reference's accumulation: acc += in[j] * w, left to right (fma chain)
        const int64_t j = idx[o + k];
        if (j < 0 || j >= Ns) { *oob = 1; continue; }
        acc = fmaf(src[j * sst.n], w[o + k], acc);
      }
      v[t] = acc;
    }
    store_planes4(hi, lo, out0 + n, make_float4(v[0], v[1], v[2], v[3]));
  }
}

constexpr int ROWS_MAX = 12288;   // floats of one source row kept in shared memory (48 KB) by the strided scatter-adds
constexpr int ROWS_BIG = 53248;   // linear-first gather / scatter: up to 208 KB of dynamic shared memory (opt-in)

// four channels per CTA (the *4 kernels): REGNET_TRAIN_GATHER_CH4=0 keeps one row per CTA
bool use_ch4(int C0, int n_rows);

template <typename Kern>
int allow_big_rows(Kern kernel, int n_floats) {
  if (n_floats * 4 > 48 * 1024)
    RN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, n_floats * 4));
  return REGNET_OK;
}

// gin[b, c, :] = scatter-add of gout[b, c0 + c, :] (row stride MK, batch stride gbs) by index -- one CTA per (b, c) row,
// the row accumulated in shared memory and written once
__global__ void __launch_bounds__(TG)
group_backward_strided_kernel(const float* __restrict__ gout, int64_t gbs, int c0, const int64_t* __restrict__ index, int C,
                              int N, int64_t MK, float* __restrict__ gin, int* __restrict__ oob) {
  extern __shared__ float acc[];
  const int c = blockIdx.x, b = blockIdx.y;
  for (int j = threadIdx.x; j < N; j += TG) acc[j] = 0.f;
  __syncthreads();
  const float* __restrict__ g = gout + (int64_t)b * gbs + (int64_t)(c0 + c) * MK;
  const int64_t* __restrict__ idx = index + (int64_t)b * MK;
  for (int64_t e = threadIdx.x; e < MK; e += TG) {
    const int64_t j = idx[e];
    if (j < 0 || j >= N) { *oob = 1; continue; }
    atomicAdd(acc + j, g[e]);
  }
  __syncthreads();
  float* __restrict__ o = gin + ((int64_t)b * C + c) * N;
  for (int j = threadIdx.x; j < N; j += TG) o[j] = acc[j];
}

__global__ void __launch_bounds__(TG)
interp_backward_strided_kernel(const float* __restrict__ gout, int64_t gbs, int c0, const int64_t* __restrict__ index,
                               const float* __restrict__ weight, int C, int Ns, int Nd, float* __restrict__ gin,
                               int* __restrict__ oob) {
  extern __shared__ float acc[];
  const int b = blockIdx.y;
  const int64_t* __restrict__ idx = index + (int64_t)b * Nd * 3;
  const float* __restrict__ w = weight + (int64_t)b * Nd * 3;
  if ((int)gridDim.x != C) {
    // four channels per CTA (host: C % 4 == 0 and four planar rows of at most 48 KB): the 36 bytes of index and weight per
    // dense point are read once for all four; the rows stay planar so that the four atomics of a point hit different banks
    const int c4 = blockIdx.x * 4;
    for (int j = threadIdx.x; j < 4 * Ns; j += TG) acc[j] = 0.f;
    __syncthreads();
    const float* __restrict__ g4 = gout + (int64_t)b * gbs + (int64_t)(c0 + c4) * Nd;
    for (int n = threadIdx.x; n < Nd; n += TG) {
      const float gv[4] = {g4[n], g4[Nd + n], g4[2 * (int64_t)Nd + n], g4[3 * (int64_t)Nd + n]};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int64_t j = idx[(int64_t)n * 3 + k];
        if (j < 0 || j >= Ns) { *oob = 1; continue; }
        const float wk = w[(int64_t)n * 3 + k];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) atomicAdd(acc + ch * Ns + j, __fmul_rn(gv[ch], wk));
      }
    }
    __syncthreads();
    float* __restrict__ o4 = gin + ((int64_t)b * C + c4) * Ns;
    for (int j = threadIdx.x; j < 4 * Ns; j += TG) o4[j] = acc[j];
    return;
  }
  const int c = blockIdx.x;
  for (int j = threadIdx.x; j < Ns; j += TG) acc[j] = 0.f;
  __syncthreads();
  const float* __restrict__ g = gout + (int64_t)b * gbs + (int64_t)(c0 + c) * Nd;
  for (int n = threadIdx.x; n < Nd; n += TG) {
    const float gv = g[n];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int64_t j = idx[(int64_t)n * 3 + k];
      if (j < 0 || j >= Ns) { *oob = 1; continue; }
      atomicAdd(acc + j, __fmul_rn(gv, w[(int64_t)n * 3 + k]));
    }
  }
  __syncthreads();
  float* __restrict__ o = gin + ((int64_t)b * C + c) * Ns;
  for (int j = threadIdx.x; j < Ns; j += TG) o[j] = acc[j];
}

// ---- first convolution applied per SOURCE point (linear operations commute with grouping / interpolation) ------------------
// W [x_rel ; f_j] = W_x x_rel + (W_f f)_j: the feature part of a set-abstraction module's first 1x1 convolution is evaluated
// once per point of the previous level (Y = W_f f, a GEMM over N points instead of M*64 grouped positions) and the grouped
// pre-activation is a gather of Y plus the three-term coordinate part; likewise W [interp(s) ; d] = interp(W_s s) + W_d d
// for a feature-propagation module.  The eval plan does the same (gather.cu sa_gather_affine / fp_interp_affine); here the
// result is the convolution OUTPUT Z0 in fp32 together with its batch moments, i.e. exactly what conv1x1_tc_kernel<MOMENTS>
// would have produced, and the backward is a scatter-add of dZ0 (plus tiny reductions for W_x / W_d).
// One CTA per (b, c) row: the row of Y (<= 12 288 floats) is staged in shared memory.
__device__ __forceinline__ void block_moments(float s1, float s2, float pivot, int64_t n, double* __restrict__ mom) {
  __shared__ float red[2][TG / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < TG / 32; ++i) { a += red[0][i]; b += red[1][i]; }
    // sums of deviations from the pivot -> raw moments in fp64
    const double nd = (double)n, mean = (double)pivot + a / nd;
    const double m2 = fmax(b - a * a / nd, 0.0);
    atomicAdd(mom, nd * mean);
    atomicAdd(mom + 1, m2 + nd * mean * mean);
  }
}

__global__ void __launch_bounds__(TG)
sa_gather_linear_kernel(const float* __restrict__ Y, const int64_t* __restrict__ index, const float* __restrict__ xr,
                        const float* __restrict__ Wx, int ldwx, int C0, int N, int64_t MK, float* __restrict__ Z,
                        double* __restrict__ moments, int* __restrict__ oob) {
  extern __shared__ float row[];
  const int c = blockIdx.x, b = blockIdx.y;
  const float* __restrict__ y = Y + ((int64_t)b * C0 + c) * N;
  for (int j = threadIdx.x; j < N; j += TG) row[j] = y[j];
  __syncthreads();
  const float w0 = Wx[(int64_t)c * ldwx], w1 = Wx[(int64_t)c * ldwx + 1], w2 = Wx[(int64_t)c * ldwx + 2];
  const int64_t* __restrict__ idx = index + (int64_t)b * MK;
  const float* __restrict__ x0 = xr + (int64_t)b * 3 * MK;
  float* __restrict__ z = Z + ((int64_t)b * C0 + c) * MK;
  float s1 = 0.f, s2 = 0.f;
  const float pivot = row[min((int64_t)N - 1, max((int64_t)0, idx[0]))];
  for (int64_t e = (int64_t)threadIdx.x * 4; e < MK; e += TG * 4) {
    const longlong2 j01 = *reinterpret_cast<const longlong2*>(idx + e), j23 = *reinterpret_cast<const longlong2*>(idx + e + 2);
    const float4 a0 = *reinterpret_cast<const float4*>(x0 + e), a1 = *reinterpret_cast<const float4*>(x0 + MK + e),
                 a2 = *reinterpret_cast<const float4*>(x0 + 2 * MK + e);
    const int64_t j[4] = {j01.x, j01.y, j23.x, j23.y};
    const float ax[4] = {a0.x, a0.y, a0.z, a0.w}, ay[4] = {a1.x, a1.y, a1.z, a1.w}, az[4] = {a2.x, a2.y, a2.z, a2.w};
    float v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float g = 0.f;
      if (j[t] < 0 || j[t] >= N) *oob = 1; else g = row[j[t]];
      v[t] = fmaf(w2, az[t], fmaf(w1, ay[t], fmaf(w0, ax[t], g)));
      const float d = v[t] - pivot;
      s1 += d;
      s2 = fmaf(d, d, s2);
    }
    *reinterpret_cast<float4*>(z + e) = make_float4(v[0], v[1], v[2], v[3]);
  }
  block_moments(s1, s2, pivot, MK, moments + 2 * c);
}

// dY[b, c, :] = scatter-add of dZ[b, c, :] by index;  dWx_part[b, c, d] = sum_e dZ[b, c, e] * xr[b, d, e]
__global__ void __launch_bounds__(TG)
sa_scatter_linear_kernel(const float* __restrict__ dZ, const int64_t* __restrict__ index, const float* __restrict__ xr, int C0,
                         int N, int64_t MK, float* __restrict__ dY, float* __restrict__ dwx_part, int* __restrict__ oob) {
  extern __shared__ float acc[];
  __shared__ float red[3][TG / 32];
  const int c = blockIdx.x, b = blockIdx.y;
  for (int j = threadIdx.x; j < N; j += TG) acc[j] = 0.f;
  __syncthreads();
  const float* __restrict__ g = dZ + ((int64_t)b * C0 + c) * MK;
  const int64_t* __restrict__ idx = index + (int64_t)b * MK;
  const float* __restrict__ x0 = xr + (int64_t)b * 3 * MK;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f;
  for (int64_t e = threadIdx.x; e < MK; e += TG) {
    const float gv = g[e];
    const int64_t j = idx[e];
    if (j < 0 || j >= N) *oob = 1; else atomicAdd(acc + j, gv);
    t0 = fmaf(gv, x0[e], t0);
    t1 = fmaf(gv, x0[MK + e], t1);
    t2 = fmaf(gv, x0[2 * MK + e], t2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    t0 += __shfl_xor_sync(0xffffffffu, t0, o);
    t1 += __shfl_xor_sync(0xffffffffu, t1, o);
    t2 += __shfl_xor_sync(0xffffffffu, t2, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = t0; red[1][threadIdx.x >> 5] = t1; red[2][threadIdx.x >> 5] = t2; }
  __syncthreads();
  float* __restrict__ o = dY + ((int64_t)b * C0 + c) * N;
  for (int j = threadIdx.x; j < N; j += TG) o[j] = acc[j];
  if (threadIdx.x < 3) {
    float a = 0.f;
    for (int i = 0; i < TG / 32; ++i) a += red[threadIdx.x][i];
    dwx_part[((int64_t)b * C0 + c) * 3 + threadIdx.x] = a;
  }
}

// ---- four channels per CTA ------------------------------------------------------------------------------------------------
// The one-row kernels above re-read the index (8 B) and the coordinate / weight rows (12 B) of every position once per
// CHANNEL: 20 B x 983 040 positions x 256 channels = 5 GB of L2 traffic per level-1 launch, more than the 1 GB they write
// (ncu: DRAM 18 %, profiles/r02_ncu_full_train_gather.txt).  With four channels per CTA the rows sit interleaved in shared
// memory (one float4 per source point: one LDS.128 gathers all four) and that traffic drops 4x.  Used by the two gathers
// when C0 % 4 == 0 and the four rows fit 100 KB (two CTAs per SM); same arithmetic per element, so the values are
// bit-identical.  Level-1 gather 0.70 -> 0.26 ms, level 2 0.30 -> 0.26 ms total 1.00 -> 0.51 ms, FP gather 0.78 -> 0.64 ms.
constexpr int CH4_MAX_ROWS = 6400;

__device__ __forceinline__ void block_moments4(const float (&s1)[4], const float (&s2)[4], const float (&pivot)[4], int64_t n,
                                               double* __restrict__ mom) {
  __shared__ float red4[8][TG / 32];
  __shared__ float piv4[4];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    if (threadIdx.x == 0) piv4[ch] = pivot[ch];
    float a = s1[ch], b = s2[ch];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { red4[2 * ch][threadIdx.x >> 5] = a; red4[2 * ch + 1][threadIdx.x >> 5] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const int ch = threadIdx.x;
    double a = 0.0, b = 0.0;
    for (int i = 0; i < TG / 32; ++i) { a += red4[2 * ch][i]; b += red4[2 * ch + 1][i]; }
    const double nd = (double)n, mean = (double)piv4[ch] + a / nd;
    const double m2 = fmax(b - a * a / nd, 0.0);
    atomicAdd(mom + 2 * ch, nd * mean);
    atomicAdd(mom + 2 * ch + 1, m2 + nd * mean * mean);
  }
}

// stage rows c0 .. c0 + 3 of a (C, n) matrix as float4 per column
__device__ __forceinline__ void stage_rows4(const float* __restrict__ y, int64_t ld, int n, float4* __restrict__ row4) {
  for (int j = threadIdx.x; j < n; j += TG) row4[j] = make_float4(y[j], y[ld + j], y[2 * ld + j], y[3 * ld + j]);
  __syncthreads();
}

__global__ void __launch_bounds__(TG)
sa_gather_linear4_kernel(const float* __restrict__ Y, const int64_t* __restrict__ index, const float* __restrict__ xr,
                         const float* __restrict__ Wx, int ldwx, int C0, int N, int64_t MK, float* __restrict__ Z,
                         double* __restrict__ moments, int* __restrict__ oob) {
  extern __shared__ float4 row4[];
  const int c0 = blockIdx.x * 4, b = blockIdx.y;
  stage_rows4(Y + ((int64_t)b * C0 + c0) * N, N, N, row4);
  float w[4][3];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch)
#pragma unroll
    for (int d = 0; d < 3; ++d) w[ch][d] = Wx[(int64_t)(c0 + ch) * ldwx + d];
  const int64_t* __restrict__ idx = index + (int64_t)b * MK;
  const float* __restrict__ x0 = xr + (int64_t)b * 3 * MK;
  float* __restrict__ z = Z + ((int64_t)b * C0 + c0) * MK;
  const float4 pv = row4[min((int64_t)N - 1, max((int64_t)0, idx[0]))];
  const float pivot[4] = {pv.x, pv.y, pv.z, pv.w};
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t e = (int64_t)threadIdx.x * 4; e < MK; e += TG * 4) {
    const longlong2 j01 = *reinterpret_cast<const longlong2*>(idx + e), j23 = *reinterpret_cast<const longlong2*>(idx + e + 2);
    const float4 a0 = *reinterpret_cast<const float4*>(x0 + e), a1 = *reinterpret_cast<const float4*>(x0 + MK + e),
                 a2 = *reinterpret_cast<const float4*>(x0 + 2 * MK + e);
    const int64_t j[4] = {j01.x, j01.y, j23.x, j23.y};
    const float ax[4] = {a0.x, a0.y, a0.z, a0.w}, ay[4] = {a1.x, a1.y, a1.z, a1.w}, az[4] = {a2.x, a2.y, a2.z, a2.w};
    float v[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j[t] < 0 || j[t] >= N) *oob = 1; else g = row4[j[t]];
      const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        v[ch][t] = fmaf(w[ch][2], az[t], fmaf(w[ch][1], ay[t], fmaf(w[ch][0], ax[t], gg[ch])));
        const float d = v[ch][t] - pivot[ch];
        s1[ch] += d;
        s2[ch] = fmaf(d, d, s2[ch]);
      }
    }
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      *reinterpret_cast<float4*>(z + ch * MK + e) = make_float4(v[ch][0], v[ch][1], v[ch][2], v[ch][3]);
  }
  block_moments4(s1, s2, pivot, MK, moments + 2 * c0);
}

__global__ void __launch_bounds__(TG)
fp_gather_linear4_kernel(const float* __restrict__ Ys, const int64_t* __restrict__ index, const float* __restrict__ weight,
                         const float* __restrict__ dense, Strides3 dst, int nd, const float* __restrict__ Wd, int ldwd, int C0,
                         int Ns, int Nd, float* __restrict__ Z, double* __restrict__ moments, int* __restrict__ oob) {
  extern __shared__ float4 row4[];
  const int c0 = blockIdx.x * 4, b = blockIdx.y;
  stage_rows4(Ys + ((int64_t)b * C0 + c0) * Ns, Ns, Ns, row4);
  float wd[4][4];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch)
#pragma unroll
    for (int d = 0; d < 4; ++d) wd[ch][d] = d < nd ? Wd[(int64_t)(c0 + ch) * ldwd + d] : 0.f;
  const int64_t* __restrict__ idx = index + (int64_t)b * Nd * 3;
  const float* __restrict__ w = weight + (int64_t)b * Nd * 3;
  const float* __restrict__ dn = dense + (int64_t)b * dst.b;
  float* __restrict__ z = Z + ((int64_t)b * C0 + c0) * Nd;
  const float4 pv = row4[min((int64_t)Ns - 1, max((int64_t)0, idx[0]))];
  const float pivot[4] = {pv.x, pv.y, pv.z, pv.w};
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n = threadIdx.x; n < Nd; n += TG) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int64_t j = idx[(int64_t)n * 3 + k];
      if (j < 0 || j >= Ns) { *oob = 1; continue; }
      const float4 g = row4[j];
      const float wk = w[(int64_t)n * 3 + k];
      acc[0] = fmaf(g.x, wk, acc[0]);
      acc[1] = fmaf(g.y, wk, acc[1]);
      acc[2] = fmaf(g.z, wk, acc[2]);
      acc[3] = fmaf(g.w, wk, acc[3]);
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      if (d < nd) {
        const float dv = dn[(int64_t)d * dst.c + (int64_t)n * dst.n];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) acc[ch] = fmaf(wd[ch][d], dv, acc[ch]);
      }
    }
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      z[(int64_t)ch * Nd + n] = acc[ch];
      const float dv = acc[ch] - pivot[ch];
      s1[ch] += dv;
      s2[ch] = fmaf(dv, dv, s2[ch]);
    }
  }
  block_moments4(s1, s2, pivot, Nd, moments + 2 * c0);
}

// Scatter with TWO planar rows per CTA: index and coordinates read once per two channels, and 40 KB of rows still leave five
// CTAs per SM (four rows -- 80 KB, two CTAs -- were measured slower than one: 1.10 -> 1.23 ms).
__global__ void __launch_bounds__(TG)
sa_scatter_linear2_kernel(const float* __restrict__ dZ, const int64_t* __restrict__ index, const float* __restrict__ xr, int C0,
                          int N, int64_t MK, float* __restrict__ dY, float* __restrict__ dwx_part, int* __restrict__ oob) {
  extern __shared__ float acc[];
  __shared__ float red[6][TG / 32];
  const int c0 = blockIdx.x * 2, b = blockIdx.y;
  for (int j = threadIdx.x; j < 2 * N; j += TG) acc[j] = 0.f;
  __syncthreads();
  const float* __restrict__ g = dZ + ((int64_t)b * C0 + c0) * MK;
  const int64_t* __restrict__ idx = index + (int64_t)b * MK;
  const float* __restrict__ x0 = xr + (int64_t)b * 3 * MK;
  float t[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  for (int64_t e = threadIdx.x; e < MK; e += TG) {
    const int64_t j = idx[e];
    const float xa = x0[e], xb = x0[MK + e], xc = x0[2 * MK + e];
    const float g0 = g[e], g1 = g[MK + e];
    if (j < 0 || j >= N) *oob = 1;
    else { atomicAdd(acc + j, g0); atomicAdd(acc + N + j, g1); }
    t[0][0] = fmaf(g0, xa, t[0][0]); t[0][1] = fmaf(g0, xb, t[0][1]); t[0][2] = fmaf(g0, xc, t[0][2]);
    t[1][0] = fmaf(g1, xa, t[1][0]); t[1][1] = fmaf(g1, xb, t[1][1]); t[1][2] = fmaf(g1, xc, t[1][2]);
  }
#pragma unroll
  for (int ch = 0; ch < 2; ++ch)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float a = t[ch][d];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if ((threadIdx.x & 31) == 0) red[3 * ch + d][threadIdx.x >> 5] = a;
    }
  __syncthreads();
  float* __restrict__ o = dY + ((int64_t)b * C0 + c0) * N;
  for (int j = threadIdx.x; j < 2 * N; j += TG) o[j] = acc[j];
  if (threadIdx.x < 6) {
    float a = 0.f;
    for (int i = 0; i < TG / 32; ++i) a += red[threadIdx.x][i];
    dwx_part[((int64_t)b * C0 + c0) * 3 + threadIdx.x] = a;
  }
}

// Z[b, c, n] = sum_k w[b,n,k] * Ys[b, c, idx[b,n,k]] + sum_d Wd[c, d] * dense[b, d, n]   (nd dense channels, <= 4)
__global__ void __launch_bounds__(TG)
fp_gather_linear_kernel(const float* __restrict__ Ys, const int64_t* __restrict__ index, const float* __restrict__ weight,
                        const float* __restrict__ dense, Strides3 dst, int nd, const float* __restrict__ Wd, int ldwd, int C0,
                        int Ns, int Nd, float* __restrict__ Z, double* __restrict__ moments, int* __restrict__ oob) {
  extern __shared__ float row[];
  const int c = blockIdx.x, b = blockIdx.y;
  const float* __restrict__ y = Ys + ((int64_t)b * C0 + c) * Ns;
  for (int j = threadIdx.x; j < Ns; j += TG) row[j] = y[j];
  __syncthreads();
  float wd[4] = {0.f, 0.f, 0.f, 0.f};
  for (int d = 0; d < nd; ++d) wd[d] = Wd[(int64_t)c * ldwd + d];
  const int64_t* __restrict__ idx = index + (int64_t)b * Nd * 3;
  const float* __restrict__ w = weight + (int64_t)b * Nd * 3;
  const float* __restrict__ dn = dense + (int64_t)b * dst.b;
  float* __restrict__ z = Z + ((int64_t)b * C0 + c) * Nd;
  float s1 = 0.f, s2 = 0.f;
  const float pivot = row[min((int64_t)Ns - 1, max((int64_t)0, idx[0]))];
  for (int n = threadIdx.x; n < Nd; n += TG) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int64_t j = idx[(int64_t)n * 3 + k];
      if (j < 0 || j >= Ns) { *oob = 1; continue; }
      acc = fmaf(row[j], w[(int64_t)n * 3 + k], acc);
    }
    for (int d = 0; d < nd; ++d) acc = fmaf(wd[d], dn[(int64_t)d * dst.c + (int64_t)n * dst.n], acc);
    z[n] = acc;
    const float dv = acc - pivot;
    s1 += dv;
    s2 = fmaf(dv, dv, s2);
  }
  block_moments(s1, s2, pivot, Nd, moments + 2 * c);
}

// dWd_part[b, c, d] = sum_n dZ[b, c, n] * dense[b, d, n]
__global__ void __launch_bounds__(TG)
fp_dense_wgrad_kernel(const float* __restrict__ dZ, const float* __restrict__ dense, Strides3 dst, int nd, int C0, int Nd,
                      float* __restrict__ part) {
  __shared__ float red[4][TG / 32];
  const int c = blockIdx.x, b = blockIdx.y;
  const float* __restrict__ g = dZ + ((int64_t)b * C0 + c) * Nd;
  const float* __restrict__ dn = dense + (int64_t)b * dst.b;
  float t[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n = threadIdx.x; n < Nd; n += TG) {
    const float gv = g[n];
    for (int d = 0; d < nd; ++d) t[d] = fmaf(gv, dn[(int64_t)d * dst.c + (int64_t)n * dst.n], t[d]);
  }
#pragma unroll
  for (int d = 0; d < 4; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t[d] += __shfl_xor_sync(0xffffffffu, t[d], o);
    if ((threadIdx.x & 31) == 0) red[d][threadIdx.x >> 5] = t[d];
  }
  __syncthreads();
  if (threadIdx.x < nd) {
    float a = 0.f;
    for (int i = 0; i < TG / 32; ++i) a += red[threadIdx.x][i];
    part[((int64_t)b * C0 + c) * nd + threadIdx.x] = a;
  }
}

// ---- set-abstraction level 0, first block, without its 2.5 GB pre-activation --------------------------------------------------
// The first block of level 0 maps 6 inputs per grouped position ([xyz - centre | rgb]) to 128 channels over 4.9 M positions:
// as a convolution it writes Z0 (2.5 GB), BatchNorm reads it twice, and the backward reads and writes it again.  Z0 is a
// LINEAR function of 6 numbers, so
//   * its batch moments follow from the 6 sums and 21 products of the inputs:  E[z_c] = w_c . E[x],  E[z_c^2] = w_c' E[x x'] w_c
//     (sa0_input_moments_kernel: one pass over the inputs, 27 numbers);
//   * the activation y = relu(bn(z)) is recomputed from the inputs wherever it is needed (sa0_apply_planes_kernel writes it
//     once, as the next block's operand planes);
//   * the whole backward of the block needs 7 sums per channel over g = dy * [bn(z) > 0]:  G0_c = sum g,  G_cd = sum g x_d
//     (sa0_backward_sums_kernel: one pass over dy, z recomputed for the mask); with them
//       dbeta = G0,  dgamma = is (w_c . G_c - mu G0),  dW_cd = sc (G_cd - m1 Sx_d - m2 is (w_c' Sxx_d - mu Sx_d)),
//     m1 = G0 / n, m2 = dgamma / n  (conv_train.py) -- the inputs need no gradient, so dZ0 is never formed.
// Every thread owns 8 consecutive positions (one centroid: K is a multiple of 8) and loops over the channels.
constexpr int SA0_PPT = 8;

struct Sa0In {
  const float* xyz; Strides3 xst; const float* new_xyz; const float* feat; Strides3 fst; const int64_t* index;
  int N, M, K;
};

__device__ __forceinline__ void sa0_load_inputs(const Sa0In& a, int b, int64_t e, float (&x)[SA0_PPT][6], int* oob) {
  const int64_t MK = (int64_t)a.M * a.K;
  const int64_t* __restrict__ idx = a.index + (int64_t)b * MK + e;
  const int m = (int)(e / a.K);
  const float c0 = a.new_xyz[((int64_t)b * 3 + 0) * a.M + m], c1 = a.new_xyz[((int64_t)b * 3 + 1) * a.M + m],
              c2 = a.new_xyz[((int64_t)b * 3 + 2) * a.M + m];
  const float* __restrict__ px = a.xyz + (int64_t)b * a.xst.b;
  const float* __restrict__ pf = a.feat + (int64_t)b * a.fst.b;
#pragma unroll
  for (int t = 0; t < SA0_PPT; ++t) {
    int64_t j = idx[t];
    if (j < 0 || j >= a.N) { *oob = 1; j = 0; }
    x[t][0] = __fsub_rn(px[j * a.xst.n], c0);
    x[t][1] = __fsub_rn(px[j * a.xst.n + a.xst.c], c1);
    x[t][2] = __fsub_rn(px[j * a.xst.n + 2 * a.xst.c], c2);
    x[t][3] = pf[j * a.fst.n];
    x[t][4] = pf[j * a.fst.n + a.fst.c];
    x[t][5] = pf[j * a.fst.n + 2 * a.fst.c];
  }
}

__device__ __forceinline__ float sa0_z(const float (&w)[6], const float (&x)[6]) {
  float z = __fmul_rn(w[0], x[0]);
#pragma unroll
  for (int d = 1; d < 6; ++d) z = __fmaf_rn(w[d], x[d], z);
  return z;
}

// sums[0..5] = sum x_d, sums[6 + pair(d1 <= d2)] = sum x_d1 x_d2   (fp64, accumulated over the whole batch)
__global__ void __launch_bounds__(TG)
sa0_input_moments_kernel(Sa0In a, int B, double* __restrict__ sums, int* __restrict__ oob) {
  const int64_t MK = (int64_t)a.M * a.K, per_b = MK / SA0_PPT, total = per_b * B;
  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.f;
  for (int64_t g = (int64_t)blockIdx.x * TG + threadIdx.x; g < total; g += (int64_t)gridDim.x * TG) {
    const int b = (int)(g / per_b);
    const int64_t e = (g % per_b) * SA0_PPT;
    float x[SA0_PPT][6];
    sa0_load_inputs(a, b, e, x, oob);
#pragma unroll
    for (int t = 0; t < SA0_PPT; ++t) {
      int p = 6;
#pragma unroll
      for (int d1 = 0; d1 < 6; ++d1) {
        acc[d1] += x[t][d1];
#pragma unroll
        for (int d2 = d1; d2 < 6; ++d2, ++p) acc[p] = fmaf(x[t][d1], x[t][d2], acc[p]);
      }
    }
  }
  __shared__ float red[27][TG / 32];
#pragma unroll
  for (int i = 0; i < 27; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 27) {
    double v = 0.0;
    for (int i = 0; i < TG / 32; ++i) v += red[threadIdx.x][i];
    atomicAdd(sums + threadIdx.x, v);
  }
}

// y[b, c, e] = relu(fma(W0[c,:] . x[b,e,:], scale[c], shift[c])) as bf16 hi/lo planes (B, C0, M*K)
__global__ void __launch_bounds__(TG)
sa0_apply_planes_kernel(Sa0In a, int B, const float* __restrict__ W0, const float* __restrict__ scale,
                        const float* __restrict__ shift, int C0, int relu, __nv_bfloat16* __restrict__ hi,
                        __nv_bfloat16* __restrict__ lo, int* __restrict__ oob) {
  extern __shared__ float sw[];        // [C0][8]: w0..w5, scale, shift
  for (int i = threadIdx.x; i < C0 * 8; i += TG) {
    const int c = i >> 3, d = i & 7;
    sw[i] = d < 6 ? W0[c * 6 + d] : d == 6 ? scale[c] : shift[c];
  }
  __syncthreads();
  const int64_t MK = (int64_t)a.M * a.K, per_b = MK / SA0_PPT, total = per_b * B;
  for (int64_t g = (int64_t)blockIdx.x * TG + threadIdx.x; g < total; g += (int64_t)gridDim.x * TG) {
    const int b = (int)(g / per_b);
    const int64_t e = (g % per_b) * SA0_PPT;
    float x[SA0_PPT][6];
    sa0_load_inputs(a, b, e, x, oob);
    const int64_t out0 = (int64_t)b * C0 * MK + e;
#pragma unroll 2
    for (int c = 0; c < C0; ++c) {
      const float4 wa = *reinterpret_cast<const float4*>(sw + c * 8), wb = *reinterpret_cast<const float4*>(sw + c * 8 + 4);
      const float w[6] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y};
      float y[SA0_PPT];
#pragma unroll
      for (int t = 0; t < SA0_PPT; ++t) {
        const float v = __fmaf_rn(sa0_z(w, x[t]), wb.z, wb.w);
        y[t] = relu ? fmaxf(v, 0.f) : v;
      }
      // one 16-byte store per plane: the warp writes 512 contiguous bytes (two 8-byte stores left every sector half written
      // per instruction)
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) split_bf16_pair(y[2 * u], y[2 * u + 1], ph[u], pl[u]);
      *reinterpret_cast<uint4*>(hi + out0 + (int64_t)c * MK) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      *reinterpret_cast<uint4*>(lo + out0 + (int64_t)c * MK) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
  }
}

// G[c][0] = sum g, G[c][1 + d] = sum g x_d with g = dy * [fma(z, scale, shift) > 0]   (fp64 accumulators (C0, 7))
__global__ void __launch_bounds__(TG)
sa0_backward_sums_kernel(Sa0In a, int B, const float* __restrict__ dy, const float* __restrict__ W0,
                         const float* __restrict__ scale, const float* __restrict__ shift, int C0, int relu,
                         double* __restrict__ G, int* __restrict__ oob) {
  extern __shared__ float sm[];        // [C0][8] weights, then [C0][7] per-CTA accumulators
  float* sw = sm;
  float* sg = sm + C0 * 8;
  for (int i = threadIdx.x; i < C0 * 8; i += TG) {
    const int c = i >> 3, d = i & 7;
    sw[i] = d < 6 ? W0[c * 6 + d] : d == 6 ? scale[c] : shift[c];
  }
  for (int i = threadIdx.x; i < C0 * 7; i += TG) sg[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t MK = (int64_t)a.M * a.K, per_b = MK / SA0_PPT, total = per_b * B;
  // warp-uniform trip count: the tail of the last warp works on a clamped position with zero weight
  const int64_t g0 = (int64_t)blockIdx.x * TG + (threadIdx.x & ~31);
  for (int64_t gw = g0; gw < total; gw += (int64_t)gridDim.x * TG) {
    const int64_t g = gw + lane;
    const bool live = g < total;
    const int64_t gc = live ? g : total - 1;
    const int b = (int)(gc / per_b);
    const int64_t e = (gc % per_b) * SA0_PPT;
    float x[SA0_PPT][6];
    sa0_load_inputs(a, b, e, x, oob);
    const float* __restrict__ dyb = dy + (int64_t)b * C0 * MK + e;
#pragma unroll 1
    for (int c = 0; c < C0; ++c) {
      const float4 wa = *reinterpret_cast<const float4*>(sw + c * 8), wb = *reinterpret_cast<const float4*>(sw + c * 8 + 4);
      const float w[6] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y};
      const float4 d0 = *reinterpret_cast<const float4*>(dyb + (int64_t)c * MK), d1 = *reinterpret_cast<const float4*>(dyb + (int64_t)c * MK + 4);
      const float dv[SA0_PPT] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < SA0_PPT; ++t) {
        const float v = __fmaf_rn(sa0_z(w, x[t]), wb.z, wb.w);
        const float gq = (live && (!relu || v > 0.f)) ? dv[t] : 0.f;
        acc[0] += gq;
#pragma unroll
        for (int d = 0; d < 6; ++d) acc[1 + d] = fmaf(gq, x[t][d], acc[1 + d]);
      }
      // transpose-reduce of the 7 (+1 pad) partial sums: every exchange halves the values a lane carries -- 4 + 2 + 1
      // shuffles, then two more on the single value left -- 9 shuffles instead of 35 butterflies; lanes 4k hold sum k
      {
        const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
        float w4[4], w2[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float hi_v = i + 4 < 7 ? acc[i + 4] : 0.f;
          const float send = h16 ? acc[i] : hi_v, keep = h16 ? hi_v : acc[i];
          w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float send = h8 ? w4[i] : w4[i + 2], keep = h8 ? w4[i + 2] : w4[i];
          w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        const float send = h4 ? w2[0] : w2[1], keep = h4 ? w2[1] : w2[0];
        float tot = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        tot += __shfl_xor_sync(0xffffffffu, tot, 2);
        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
        const int k = (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0);
        if ((lane & 3) == 0 && k < 7) atomicAdd(sg + c * 7 + k, tot);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C0 * 7; i += TG) atomicAdd(G + i, (double)sg[i]);
}

// closed forms of the level-0 first block (see above): tri(d1, d2) indexes the 21 upper-triangle sums behind the 6 first sums
__device__ __forceinline__ double sa0_sxx(const double* __restrict__ sums, int d1, int d2) {
  if (d1 > d2) { const int t = d1; d1 = d2; d2 = t; }
  return sums[6 + d1 * 6 - d1 * (d1 - 1) / 2 + (d2 - d1)];
}

// moments[c] = (sum z_c, sum z_c^2) = (w_c . Sx, w_c' Sxx w_c)
__global__ void __launch_bounds__(128)
sa0_moments_from_sums_kernel(const double* __restrict__ sums, const float* __restrict__ W0, int C0, double* __restrict__ moments) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= C0) return;
  double w[6];
#pragma unroll
  for (int d = 0; d < 6; ++d) w[d] = (double)W0[c * 6 + d];
  double m1 = 0.0, m2 = 0.0;
#pragma unroll
  for (int d1 = 0; d1 < 6; ++d1) {
    m1 += w[d1] * sums[d1];
#pragma unroll
    for (int d2 = 0; d2 < 6; ++d2) m2 += w[d1] * w[d2] * sa0_sxx(sums, d1, d2);
  }
  moments[2 * c] = m1;
  moments[2 * c + 1] = m2;
}

// dbeta = G0, dgamma = is (w . G - mu G0), dW_d = sc (G_d - (G0 / n) Sx_d - (dgamma is / n) (w' Sxx_d - mu Sx_d))
__global__ void __launch_bounds__(128)
sa0_backward_finalize_kernel(const double* __restrict__ G, const double* __restrict__ sums, const float* __restrict__ W0,
                             const float* __restrict__ invstd, const float* __restrict__ scale, int C0, double count,
                             float* __restrict__ dW0, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= C0) return;
  double w[6];
#pragma unroll
  for (int d = 0; d < 6; ++d) w[d] = (double)W0[c * 6 + d];
  double mu = 0.0, wg = 0.0;
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    mu += w[d] * sums[d];
    wg += w[d] * G[c * 7 + 1 + d];
  }
  mu /= count;
  const double g0 = G[c * 7], is = (double)invstd[c], sc = (double)scale[c];
  const double dg = is * (wg - mu * g0);
  dbeta[c] = (float)g0;
  dgamma[c] = (float)dg;
  if (dW0) {
#pragma unroll
    for (int d = 0; d < 6; ++d) {
      double zx = -mu * sums[d];
#pragma unroll
      for (int e = 0; e < 6; ++e) zx += w[e] * sa0_sxx(sums, e, d);
      dW0[c * 6 + d] = (float)(sc * (G[c * 7 + 1 + d] - g0 / count * sums[d] - dg * is / count * zx));
    }
  }
}

unsigned grid_x(int64_t elems4) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>((elems4 + TG - 1) / TG, 64));
}

bool use_ch4(int C0, int n_rows) {
  static const int on = getenv("REGNET_TRAIN_GATHER_CH4") ? atoi(getenv("REGNET_TRAIN_GATHER_CH4")) : 1;
  return on && C0 % 4 == 0 && n_rows <= CH4_MAX_ROWS;
}

}  // namespace

}  // namespace regnet

using namespace regnet;

extern "C" {

int regnet_sa_group_planes(const float* xyz, int64_t xsb, int64_t xsc, int64_t xsn, const float* new_xyz, const float* feature,
                           int64_t fsb, int64_t fsc, int64_t fsn, const int64_t* index, int B, int C, int N, int M, int K,
                           void* out_hi, void* out_lo, void* stream) {
  RN_CHECK_ARG(xyz && new_xyz && index && out_hi && out_lo && (feature || C == 0), "sa_group_planes: null argument");
  RN_CHECK_ARG(B > 0 && N > 0 && M > 0 && K > 0 && C >= 0, "sa_group_planes: empty input");
  RN_CHECK_ARG(K % 4 == 0, "sa_group_planes: the neighbour count (%d) must be a multiple of 4", K);
  RN_CHECK_ARG(B <= 65535 && C + 3 <= 65535, "sa_group_planes: grid limit");
  const int64_t MK = (int64_t)M * K;
  sa_group_planes_kernel<<<dim3(grid_x(MK / 4), C + 3, B), TG, 0, (cudaStream_t)stream>>>(
      xyz, Strides3{xsb, xsc, xsn}, new_xyz, feature, Strides3{fsb, fsc, fsn}, index, C, N, M, K, (__nv_bfloat16*)out_hi,
      (__nv_bfloat16*)out_lo, oob_flag());
  RN_LAUNCH_CHECK("sa_group_planes_kernel");
  return REGNET_OK;
}

int regnet_fp_interp_planes(const float* sparse, int64_t ssb, int64_t ssc, int64_t ssn, const float* dense, int64_t dsb,
                            int64_t dsc, int64_t dsn, const int64_t* index, const float* weight, int B, int C2, int C1, int Ns,
                            int Nd, void* out_hi, void* out_lo, void* stream) {
  RN_CHECK_ARG(sparse && index && weight && out_hi && out_lo && (dense || C1 == 0), "fp_interp_planes: null argument");
  RN_CHECK_ARG(B > 0 && C2 > 0 && C1 >= 0 && Ns > 0 && Nd > 0, "fp_interp_planes: empty input");
  RN_CHECK_ARG(Nd % 4 == 0, "fp_interp_planes: the dense point count (%d) must be a multiple of 4", Nd);
  RN_CHECK_ARG(B <= 65535 && C2 + C1 <= 65535, "fp_interp_planes: grid limit");
  const int rows_y = (C2 % 4 == 0 && use_ch4(4, 0)) ? C2 / 4 + C1 : C2 + C1;
  fp_interp_planes_kernel<<<dim3(grid_x(Nd / 4), rows_y, B), TG, 0, (cudaStream_t)stream>>>(
      sparse, Strides3{ssb, ssc, ssn}, dense, Strides3{dsb, dsc, dsn}, index, weight, C2, C1, Ns, Nd, (__nv_bfloat16*)out_hi,
      (__nv_bfloat16*)out_lo, oob_flag());
  RN_LAUNCH_CHECK("fp_interp_planes_kernel");
  return REGNET_OK;
}

int regnet_group_points_backward_strided(const float* grad_out, int64_t batch_stride, int c0, const int64_t* index, int B,
                                         int C, int N, int M, int K, float* grad_in, void* stream) {
  RN_CHECK_ARG(grad_out && index && grad_in, "group_points_backward_strided: null argument");
  RN_CHECK_ARG(B > 0 && C > 0 && N > 0 && M > 0 && K > 0 && c0 >= 0, "group_points_backward_strided: empty input");
  RN_CHECK_ARG(N <= ROWS_MAX && B <= 65535, "group_points_backward_strided: at most %d source points per cloud", ROWS_MAX);
  group_backward_strided_kernel<<<dim3(C, B), TG, sizeof(float) * (size_t)N, (cudaStream_t)stream>>>(
      grad_out, batch_stride, c0, index, C, N, (int64_t)M * K, grad_in, oob_flag());
  RN_LAUNCH_CHECK("group_backward_strided_kernel");
  return REGNET_OK;
}

int regnet_interpolate_backward_strided(const float* grad_out, int64_t batch_stride, int c0, const int64_t* index,
                                        const float* weight, int B, int C, int Ns, int Nd, float* grad_in, void* stream) {
  RN_CHECK_ARG(grad_out && index && weight && grad_in, "interpolate_backward_strided: null argument");
  RN_CHECK_ARG(B > 0 && C > 0 && Ns > 0 && Nd > 0 && c0 >= 0, "interpolate_backward_strided: empty input");
  RN_CHECK_ARG(Ns <= ROWS_MAX && B <= 65535, "interpolate_backward_strided: at most %d sparse points per cloud", ROWS_MAX);
  if (C % 4 == 0 && 4 * Ns <= ROWS_MAX && use_ch4(4, 0)) {
    interp_backward_strided_kernel<<<dim3(C / 4, B), TG, sizeof(float) * 4 * (size_t)Ns, (cudaStream_t)stream>>>(
        grad_out, batch_stride, c0, index, weight, C, Ns, Nd, grad_in, oob_flag());
    RN_LAUNCH_CHECK("interp_backward_strided_kernel");
    return REGNET_OK;
  }
  interp_backward_strided_kernel<<<dim3(C, B), TG, sizeof(float) * (size_t)Ns, (cudaStream_t)stream>>>(
      grad_out, batch_stride, c0, index, weight, C, Ns, Nd, grad_in, oob_flag());
  RN_LAUNCH_CHECK("interp_backward_strided_kernel");
  return REGNET_OK;
}


int regnet_sa_gather_linear(const float* Y, const int64_t* index, const float* xyz_rel, const float* Wx, int ldwx, int B, int C0,
                            int N, int M, int K, float* Z, double* moments, void* stream) {
  RN_CHECK_ARG(Y && index && xyz_rel && Wx && Z && moments, "sa_gather_linear: null argument");
  RN_CHECK_ARG(B > 0 && C0 > 0 && N > 0 && M > 0 && K > 0 && K % 4 == 0, "sa_gather_linear: bad shape");
  RN_CHECK_ARG(N <= ROWS_BIG && B <= 65535, "sa_gather_linear: at most %d source points per cloud", ROWS_BIG);
  cudaStream_t s = (cudaStream_t)stream;
  RN_CUDA(cudaMemsetAsync(moments, 0, sizeof(double) * 2 * C0, s));
  if (use_ch4(C0, N)) {
    RN_TRY(allow_big_rows(sa_gather_linear4_kernel, 4 * N));
    sa_gather_linear4_kernel<<<dim3(C0 / 4, B), TG, sizeof(float4) * (size_t)N, s>>>(Y, index, xyz_rel, Wx, ldwx, C0, N,
                                                                                     (int64_t)M * K, Z, moments, oob_flag());
    RN_LAUNCH_CHECK("sa_gather_linear4_kernel");
    return REGNET_OK;
  }
  RN_TRY(allow_big_rows(sa_gather_linear_kernel, N));
  sa_gather_linear_kernel<<<dim3(C0, B), TG, sizeof(float) * (size_t)N, s>>>(Y, index, xyz_rel, Wx, ldwx, C0, N, (int64_t)M * K, Z,
                                                                            moments, oob_flag());
  RN_LAUNCH_CHECK("sa_gather_linear_kernel");
  return REGNET_OK;
}

int regnet_sa_scatter_linear(const float* dZ, const int64_t* index, const float* xyz_rel, int B, int C0, int N, int M, int K,
                             float* dY, float* dwx_part, void* stream) {
  RN_CHECK_ARG(dZ && index && xyz_rel && dY && dwx_part, "sa_scatter_linear: null argument");
  RN_CHECK_ARG(B > 0 && C0 > 0 && N > 0 && M > 0 && K > 0, "sa_scatter_linear: bad shape");
  RN_CHECK_ARG(N <= ROWS_BIG && B <= 65535, "sa_scatter_linear: at most %d source points per cloud", ROWS_BIG);
  if (C0 % 2 == 0 && N <= CH4_MAX_ROWS && use_ch4(4, 0)) {
    sa_scatter_linear2_kernel<<<dim3(C0 / 2, B), TG, sizeof(float) * 2 * (size_t)N, (cudaStream_t)stream>>>(
        dZ, index, xyz_rel, C0, N, (int64_t)M * K, dY, dwx_part, oob_flag());
    RN_LAUNCH_CHECK("sa_scatter_linear2_kernel");
    return REGNET_OK;
  }
  RN_TRY(allow_big_rows(sa_scatter_linear_kernel, N));
  sa_scatter_linear_kernel<<<dim3(C0, B), TG, sizeof(float) * (size_t)N, (cudaStream_t)stream>>>(
      dZ, index, xyz_rel, C0, N, (int64_t)M * K, dY, dwx_part, oob_flag());
  RN_LAUNCH_CHECK("sa_scatter_linear_kernel");
  return REGNET_OK;
}

int regnet_fp_gather_linear(const float* Ys, const int64_t* index, const float* weight, const float* dense, int64_t dsb,
                            int64_t dsc, int64_t dsn, int nd, const float* Wd, int ldwd, int B, int C0, int Ns, int Nd, float* Z,
                            double* moments, void* stream) {
  RN_CHECK_ARG(Ys && index && weight && Z && moments && (nd == 0 || (dense && Wd)), "fp_gather_linear: null argument");
  RN_CHECK_ARG(B > 0 && C0 > 0 && Ns > 0 && Nd > 0 && nd >= 0 && nd <= 4, "fp_gather_linear: bad shape (at most 4 dense channels)");
  RN_CHECK_ARG(Ns <= ROWS_BIG && B <= 65535, "fp_gather_linear: at most %d sparse points per cloud", ROWS_BIG);
  cudaStream_t s = (cudaStream_t)stream;
  RN_CUDA(cudaMemsetAsync(moments, 0, sizeof(double) * 2 * C0, s));
  if (use_ch4(C0, Ns)) {
    RN_TRY(allow_big_rows(fp_gather_linear4_kernel, 4 * Ns));
    fp_gather_linear4_kernel<<<dim3(C0 / 4, B), TG, sizeof(float4) * (size_t)Ns, s>>>(
        Ys, index, weight, dense, Strides3{dsb, dsc, dsn}, nd, Wd, ldwd, C0, Ns, Nd, Z, moments, oob_flag());
    RN_LAUNCH_CHECK("fp_gather_linear4_kernel");
    return REGNET_OK;
  }
  RN_TRY(allow_big_rows(fp_gather_linear_kernel, Ns));
  fp_gather_linear_kernel<<<dim3(C0, B), TG, sizeof(float) * (size_t)Ns, s>>>(Ys, index, weight, dense, Strides3{dsb, dsc, dsn}, nd,
                                                                             Wd, ldwd, C0, Ns, Nd, Z, moments, oob_flag());
  RN_LAUNCH_CHECK("fp_gather_linear_kernel");
  return REGNET_OK;
}

int regnet_fp_dense_wgrad(const float* dZ, const float* dense, int64_t dsb, int64_t dsc, int64_t dsn, int nd, int B, int C0, int Nd,
                          float* part, void* stream) {
  RN_CHECK_ARG(dZ && dense && part, "fp_dense_wgrad: null argument");
  RN_CHECK_ARG(B > 0 && C0 > 0 && Nd > 0 && nd > 0 && nd <= 4 && B <= 65535, "fp_dense_wgrad: bad shape");
  fp_dense_wgrad_kernel<<<dim3(C0, B), TG, 0, (cudaStream_t)stream>>>(dZ, dense, Strides3{dsb, dsc, dsn}, nd, C0, Nd, part);
  RN_LAUNCH_CHECK("fp_dense_wgrad_kernel");
  return REGNET_OK;
}


static int sa0_args(const char* who, const float* xyz, const float* new_xyz, const float* feature, const int64_t* index, int B,
                    int N, int M, int K) {
  if (!(xyz && new_xyz && feature && index)) { set_error("%s: null argument", who); return REGNET_EINVAL; }
  if (!(B > 0 && N > 0 && M > 0 && K > 0 && K % SA0_PPT == 0)) {
    set_error("%s: bad shape (the neighbour count must be a multiple of %d)", who, SA0_PPT);
    return REGNET_EINVAL;
  }
  return REGNET_OK;
}

static unsigned sa0_grid(int64_t threads_needed) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>((threads_needed + TG - 1) / TG, 148LL * 8));
}

int regnet_sa0_input_moments(const float* xyz, int64_t xsb, int64_t xsc, int64_t xsn, const float* new_xyz, const float* feature,
                             int64_t fsb, int64_t fsc, int64_t fsn, const int64_t* index, int B, int N, int M, int K,
                             const float* W0, int C0, double* sums27, double* moments, void* stream) {
  RN_TRY(sa0_args("sa0_input_moments", xyz, new_xyz, feature, index, B, N, M, K));
  RN_CHECK_ARG(sums27 != nullptr && (moments == nullptr || (W0 != nullptr && C0 > 0)), "sa0_input_moments: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  RN_CUDA(cudaMemsetAsync(sums27, 0, sizeof(double) * 27, s));
  const Sa0In a{xyz, Strides3{xsb, xsc, xsn}, new_xyz, feature, Strides3{fsb, fsc, fsn}, index, N, M, K};
  sa0_input_moments_kernel<<<sa0_grid((int64_t)B * M * K / SA0_PPT), TG, 0, s>>>(a, B, sums27, oob_flag());
  RN_LAUNCH_CHECK("sa0_input_moments_kernel");
  if (moments) {
    sa0_moments_from_sums_kernel<<<(C0 + 127) / 128, 128, 0, s>>>(sums27, W0, C0, moments);
    RN_LAUNCH_CHECK("sa0_moments_from_sums_kernel");
  }
  return REGNET_OK;
}

int regnet_sa0_backward_finalize(const double* G, const double* sums27, const float* W0, const float* invstd, const float* scale,
                                 int C0, double count, float* dW0, float* dgamma, float* dbeta, void* stream) {
  RN_CHECK_ARG(G && sums27 && W0 && invstd && scale && dgamma && dbeta && C0 > 0 && count > 0.0,
               "sa0_backward_finalize: null argument or bad shape");
  sa0_backward_finalize_kernel<<<(C0 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(G, sums27, W0, invstd, scale, C0, count, dW0,
                                                                                    dgamma, dbeta);
  RN_LAUNCH_CHECK("sa0_backward_finalize_kernel");
  return REGNET_OK;
}

int regnet_sa0_apply_planes(const float* xyz, int64_t xsb, int64_t xsc, int64_t xsn, const float* new_xyz, const float* feature,
                            int64_t fsb, int64_t fsc, int64_t fsn, const int64_t* index, int B, int N, int M, int K,
                            const float* W0, const float* scale, const float* shift, int C0, int relu, void* y_hi, void* y_lo,
                            void* stream) {
  RN_TRY(sa0_args("sa0_apply_planes", xyz, new_xyz, feature, index, B, N, M, K));
  RN_CHECK_ARG(W0 && scale && shift && y_hi && y_lo && C0 > 0 && C0 <= 1024, "sa0_apply_planes: null argument or bad width");
  const Sa0In a{xyz, Strides3{xsb, xsc, xsn}, new_xyz, feature, Strides3{fsb, fsc, fsn}, index, N, M, K};
  sa0_apply_planes_kernel<<<sa0_grid((int64_t)B * M * K / SA0_PPT), TG, sizeof(float) * 8 * C0, (cudaStream_t)stream>>>(
      a, B, W0, scale, shift, C0, relu, (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo, oob_flag());
  RN_LAUNCH_CHECK("sa0_apply_planes_kernel");
  return REGNET_OK;
}

int regnet_sa0_backward_sums(const float* xyz, int64_t xsb, int64_t xsc, int64_t xsn, const float* new_xyz, const float* feature,
                             int64_t fsb, int64_t fsc, int64_t fsn, const int64_t* index, int B, int N, int M, int K,
                             const float* dy, const float* W0, const float* scale, const float* shift, int C0, int relu,
                             double* G, void* stream) {
  RN_TRY(sa0_args("sa0_backward_sums", xyz, new_xyz, feature, index, B, N, M, K));
  RN_CHECK_ARG(dy && W0 && scale && shift && G && C0 > 0 && C0 <= 1024, "sa0_backward_sums: null argument or bad width");
  RN_CHECK_ARG((reinterpret_cast<uintptr_t>(dy) & 15) == 0, "sa0_backward_sums: dy must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  RN_CUDA(cudaMemsetAsync(G, 0, sizeof(double) * 7 * C0, s));
  const Sa0In a{xyz, Strides3{xsb, xsc, xsn}, new_xyz, feature, Strides3{fsb, fsc, fsn}, index, N, M, K};
  sa0_backward_sums_kernel<<<sa0_grid((int64_t)B * M * K / SA0_PPT), TG, sizeof(float) * 15 * C0, s>>>(
      a, B, dy, W0, scale, shift, C0, relu, G, oob_flag());
  RN_LAUNCH_CHECK("sa0_backward_sums_kernel");
  return REGNET_OK;
}

}  // extern "C"
