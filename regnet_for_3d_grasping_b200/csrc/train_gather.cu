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
  const int c = blockIdx.y, b = blockIdx.z;
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

constexpr int ROWS_MAX = 12288;   // floats of one source row kept in shared memory (48 KB)

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
  const int c = blockIdx.x, b = blockIdx.y;
  for (int j = threadIdx.x; j < Ns; j += TG) acc[j] = 0.f;
  __syncthreads();
  const float* __restrict__ g = gout + (int64_t)b * gbs + (int64_t)(c0 + c) * Nd;
  const int64_t* __restrict__ idx = index + (int64_t)b * Nd * 3;
  const float* __restrict__ w = weight + (int64_t)b * Nd * 3;
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

unsigned grid_x(int64_t elems4) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>((elems4 + TG - 1) / TG, 64));
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
  fp_interp_planes_kernel<<<dim3(grid_x(Nd / 4), C2 + C1, B), TG, 0, (cudaStream_t)stream>>>(
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
  interp_backward_strided_kernel<<<dim3(C, B), TG, sizeof(float) * (size_t)Ns, (cudaStream_t)stream>>>(
      grad_out, batch_stride, c0, index, weight, C, Ns, Nd, grad_in, oob_flag());
  RN_LAUNCH_CHECK("interp_backward_strided_kernel");
  return REGNET_OK;
}

}  // extern "C"
