// Shared-MLP layer contraction: Y = act(scale * (X W^T) + shift) [+ max over runs of 64 rows].
// One description, two engines (gemm_simt.cu: fp32 FFMA tiles; gemm_tc.cu: tcgen05 split-bf16).
//
// This is the reference's Conv1d/Conv2d(k=1, bias=False) + BatchNorm + ReLU of
// multi_model/utils/pn2_utils/nn/modules/conv.py:24-36,64-76 (eval mode: BN folded to scale/shift), and with
// pool = 64 also the `torch.max(new_feature, 3)` of pn2_utils/modules.py:245 fused into the epilogue.
// Rows are positions (point-major), columns are channels.
#pragma once
#include "common.cuh"

struct CUtensorMap_st;  // <cuda.h>

namespace regnet {

struct Epilogue {
  const float* scale = nullptr;   // (cout) or nullptr = 1
  const float* shift = nullptr;   // (cout) or nullptr = 0
  int act = 1;                    // 0 none, 1 relu, 2 sigmoid
  int pool = 0;                   // 0, or 64: rows are reduced by max in runs of 64 (P % 64 == 0)
  float* out_f32 = nullptr;       // (P or P/pool, ld_f32) fp32
  int ld_f32 = 0;
  __nv_bfloat16* out_hi = nullptr;  // (P, ld_split) bf16 planes of the same values (pool == 0 only)
  __nv_bfloat16* out_lo = nullptr;
  int ld_split = 0;
  unsigned int* tile_counter = nullptr;  // tcgen05 engine: zeroed device counter => dynamic tile scheduling
  // tcgen05 engine, pool == 0, cout <= 128: a fused 1-output head on the activated row, the reference's
  // conv_score (cout -> 1, with bias) + bn_score + sigmoid of pointnet2.py:82-84,117-119:
  //   dot_out[p] = sigmoid(dot_scale[0] * (y[p,:] . dot_w) + dot_shift[0])
  const float* dot_w = nullptr;
  const float* dot_scale = nullptr;
  const float* dot_shift = nullptr;
  float* dot_out = nullptr;
  __nv_bfloat16* pool_hi = nullptr;      // tcgen05 engine, pool != 0: the pooled values also as bf16 hi/lo planes
  __nv_bfloat16* pool_lo = nullptr;      //   (P / pool, ld_pool) -- the gather table of the next level's first layer
  int ld_pool = 0;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-v)));
  return v;
}

// X (P, ldx) fp32 with columns >= K zero, W (cout, ldw) fp32 with columns >= K zero; kpad = padded K (mult of 16)
int gemm_simt_launch(const float* X, int ldx, const float* W, int ldw, int64_t P, int kpad, int cout,
                     const Epilogue& ep, cudaStream_t stream);

// X as bf16 hi/lo planes (P, ldx), W as bf16 hi/lo planes (cout, ldw); ldx, ldw multiples of 8; K = logical depth
int gemm_tc_launch(const __nv_bfloat16* Xhi, const __nv_bfloat16* Xlo, int ldx, const __nv_bfloat16* Whi,
                   const __nv_bfloat16* Wlo, int ldw, int64_t P, int K, int cout, const Epilogue& ep,
                   cudaStream_t stream);
int gemm_tc_supported(void);
// Same contraction with a GATHERED A operand (grouping fused into the TMA producer, gemm_tc.cu GatherA): row p =
// [table[nbr[p] + (p / rows_per_cloud) * n_prev, 0..C) | Z[p, 0..3)], K = C + 3; table = bf16 hi/lo planes (table_rows, ldt)
// with C %% 64 == 0, Z = (P, 16) bf16 hi/lo planes whose first three columns are xyz - centroid; W (cout, ldw) with the
// columns in that [feature | xyz_rel] order.  act = relu, no pooling.
int gemm_tc_gather_launch(const __nv_bfloat16* Thi, const __nv_bfloat16* Tlo, int64_t table_rows, int C, int ldt,
                          const int32_t* nbr, int rows_per_cloud, int n_prev, const __nv_bfloat16* Zhi,
                          const __nv_bfloat16* Zlo, const __nv_bfloat16* Whi, const __nv_bfloat16* Wlo, int ldw, int64_t P,
                          int cout, const Epilogue& ep, cudaStream_t stream);

// 2-D bf16 tensor map over a (rows, cols) row-major matrix with leading dimension ld (elements); box = box_cols x
// box_rows, swizzle_bytes in {64, 128}.  CUtensorMap is passed opaquely so that this header needs no <cuda.h>.
int tc_make_map(::CUtensorMap_st* map, const void* base, int64_t rows, int cols, int ld, int box_rows, int box_cols,
                int swizzle_bytes);
int tc_make_map3(::CUtensorMap_st* map, const void* base, int elem_bytes, int64_t d0, int64_t d1, int64_t d2,
                 int64_t stride1, int64_t stride2, int box0, int box1);
int tc_driver_ok(void);  // 1 when the driver entry point for tensor maps is available on this box

}  // namespace regnet
