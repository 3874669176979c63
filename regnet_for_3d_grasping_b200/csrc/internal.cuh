// Internal launch functions shared between translation units of libregnet_b200.
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace regnet {

int fps_block_log2(int N);
// single_pick: one pick per cluster exchange (the lighter instruction stream, for a launch that co-runs with tensor
// kernels) instead of the multi-pick rounds (faster alone) -- identical indices either way
int fps_launch(const float* pts, Strides3 st, int B, int N, int M, int64_t* idx64, int32_t* idx32, float* new_xyz,
               int cluster_size, int threads, cudaStream_t stream, bool single_pick = false);
int ball_query_launch(const float* pts, Strides3 pst, const float* ctr, Strides3 cst, int B, int N, int M,
                      float radius, int K, int64_t* index, int64_t* count, int32_t* index32, cudaStream_t stream);
int three_nn_launch(const float* qry, Strides3 qst, const float* key, Strides3 kst, int B, int Nq, int Nk,
                    int64_t* index, float* dist, int32_t* index32, float* weight, cudaStream_t stream);

int group_forward_launch(const float* in, Strides3 st, const int64_t* index, int B, int C, int N, int M, int K,
                         float* out, int* d_oob, cudaStream_t stream);
int group_backward_launch(const float* gout, const int64_t* index, int B, int C, int N, int M, int K, float* gin,
                          int* d_oob, cudaStream_t stream);
int interp_forward_launch(const float* in, Strides3 st, const int64_t* index, const float* weight, int B, int C,
                          int Ns, int Nd, float* out, int* d_oob, cudaStream_t stream);
int interp_backward_launch(const float* gout, const int64_t* index, const float* weight, int B, int C, int Ns,
                           int Nd, float* gin, int* d_oob, cudaStream_t stream);

int sa_operand_launch(const float* xyz, Strides3 xst, const float* new_xyz, const float* feat, int64_t feat_bstride,
                      int feat_ld, int C, const int32_t* nbr, int B, int N, int M, int K, int kpad, float* out_f32,
                      __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t stream);
int sa0_fused_launch(const float* xyz, Strides3 xst, const float* new_xyz, const float* feat, int64_t feat_bstride,
                     int feat_ld, const int32_t* nbr, const float* W, int ldw, const float* scale, const float* shift,
                     int cout, int B, int M, int K, float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                     int ld_out, cudaStream_t stream);
// SA level 0, layers 0+1 fused (sa0_front.cu, tcgen05 engine only)
int sa0_front_launch(const float* xyz, Strides3 xst, const float* new_xyz, const float* feat, int64_t feat_bstride,
                     int feat_ld, const int32_t* nbr, const float* W0, int ldw0, const float* scale0,
                     const float* shift0, const __nv_bfloat16* W1hi, const __nv_bfloat16* W1lo, int ldw1,
                     const float* scale1, const float* shift1, int B, int M, __nv_bfloat16* out_hi,
                     __nv_bfloat16* out_lo, int ld_out, unsigned int* tile_counter, cudaStream_t stream);
// SA level 0, whole chain 6 -> 128 -> 128 -> 256 + 64-neighbour max-pool, activations chained through TMEM (sa0_chain.cu)
int sa0_chain_launch(const float* xyz, Strides3 xst, const float* new_xyz, const float* feat, int64_t feat_bstride,
                     int feat_ld, const int32_t* nbr, const float* W0, int ldw0, const float* scale0,
                     const float* shift0, const __nv_bfloat16* W1hi, const __nv_bfloat16* W1lo, int ldw1,
                     const float* scale1, const float* shift1, const __nv_bfloat16* W2hi, const __nv_bfloat16* W2lo,
                     int ldw2, const float* scale2, const float* shift2, int B, int M, float* out, int ld_out,
                     __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, float* dbg, unsigned int* tile_counter, int variant,
                     cudaStream_t stream);
int fp_operand_launch(const float* sparse, int64_t sparse_bstride, int sparse_ld, int C2, const float* dense,
                      int64_t dense_bstride, int dense_ld, int C1, const int32_t* idx, const float* w, int B, int Nd,
                      int kpad, float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t stream);
int sa_gather_affine_launch(const float* Z, int ldz, int n_prev, const float* xyz, Strides3 xst, const float* new_xyz,
                            const float* Wx, int ldw, const int32_t* nbr, const float* scale, const float* shift, int B,
                            int M, int cout, float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                            cudaStream_t stream);
// SA levels 1, 2, second layer with its operand produced in the kernel (gemm_fused_a.cu): sa_fold_launch turns the per-point
// GEMM output Z into Z' = scale * (Z + W_x xyz) in place and writes T = shift - scale * W_x centre per centroid; the GEMM then
// builds a[p] = relu(Z'[g(p)] + T[p / 64]) in shared memory.
int sa_fold_launch(float* Z, int ldz, int n_prev, const float* xyz, Strides3 xst, const float* new_xyz, int M, const float* Wx,
                   int ldw, const float* scale, const float* shift, int B, int C, float* T, cudaStream_t stream);
// the same operand materialised as bf16 hi/lo planes (rows, C), bit-identical to what gemm_fused_a builds in shared memory
int sa_gather_add_launch(const float* Z, int ldz, int n_prev, const float* T, const int32_t* nbr, int rows_per_cloud, int C,
                         int64_t rows, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t stream);
int gemm_fused_a_supported(int64_t P, int K, int rows_per_cloud);
int gemm_fused_a_launch(const float* Z, int ldz, const float* T, const int32_t* nbr, int rows_per_cloud, int n_prev,
                        const __nv_bfloat16* Whi, const __nv_bfloat16* Wlo, int ldw, int64_t P, int K, int cout,
                        const float* scale, const float* shift, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int ld_out,
                        unsigned int* tile_counter, cudaStream_t stream);
int fp_interp_affine_launch(const float* Y, int64_t y_bstride, int ldy, const float* D, int ldd, const float* dense3,
                            int64_t dense3_bstride, int dense3_ld, const float* Wd3, int ldw3, const int32_t* idx,
                            const float* w, const float* scale, const float* shift, int B, int Nd, int cout,
                            float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t stream);
// dst[:, c] = src[:, (c + rot) mod cols] for c < cols, zero padding up to kpad.  row_sign (optional, one per row):
// rows with row_sign[r] < 0 are negated -- used with abs_copy_launch to make the BN scale of a max-pooled layer
// non-negative (scale * (w . x) == |scale| * ((sign(scale) w) . x)), so that the max can be taken on raw accumulators
int split_rows_launch(const float* src, int64_t rows, int cols, int ld_src, int kpad, __nv_bfloat16* hi,
                      __nv_bfloat16* lo, float* f32, cudaStream_t stream, int rot = 0, const float* row_sign = nullptr);
int abs_copy_launch(const float* src, float* dst, int n, cudaStream_t stream);
int score_head_launch(const float* X, int ldx, const float* w, const float* scale, const float* shift, int64_t P,
                      int cin, float* score, cudaStream_t stream);

// uniform-grid neighbour search (grid.cu); ws = grid_workspace_bytes(B, N) bytes of device scratch per grid
constexpr int GMAX = 64;                 // at most GMAX x GMAX cells
constexpr int MAX_CELLS = GMAX * GMAX;

struct GridHeader {                      // one per cloud
  float x0, y0, inv_h;
  int gx, gy;
  float h;
  int pad[2];
};

struct GridPtrs {                        // views into a grid workspace
  GridHeader* hdr;                       // [B]
  int* cell_start;                       // [B][MAX_CELLS + 1]: cell c's run is sorted[cell_start[c] .. cell_start[c + 1])
  float4* sorted;                        // [B][N] = {x, y, z, bits(original index)}
};

__device__ __forceinline__ int cell_coord(float v, float v0, float inv_h, int g) {
  int c = (int)floorf((v - v0) * inv_h);
  return min(g - 1, max(0, c));
}

GridPtrs grid_carve(void* ws, int B, int N);
int64_t grid_workspace_bytes(int B, int N);
int grid_build_launch(const float* pts, Strides3 st, int B, int N, float min_cell, void* ws, cudaStream_t stream,
                      bool stable = false);   // stable: ascending original index inside every cell
int ball_query_grid_launch(const float* pts, Strides3 pst, const float* ctr, Strides3 cst, int B, int N, int M,
                           float radius, const void* ws, int32_t* index32, cudaStream_t stream,
                           int64_t* index64 = nullptr, int64_t* count64 = nullptr);
int three_nn_grid_launch(const float* qry, Strides3 qst, const float* key, Strides3 kst, int B, int Nq, int Nk,
                         const void* ws, int32_t* index32, float* weight, cudaStream_t stream,
                         int64_t* index64 = nullptr, float* dist = nullptr);

int* oob_flag();  // per-device lazily allocated device int, set by kernels that meet an out-of-range index

}  // namespace regnet
