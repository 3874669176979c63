/*
 * regnet_b200.h -- C ABI of libregnet_b200.so: REGNet's point-cloud hot path as sm_100a CUDA.
 *
 * Plain pointers and sizes only (no torch types).  All pointers are DEVICE pointers unless stated otherwise;
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Every function returns 0 on
 * success and a non-zero REGNET_E* code on failure; regnet_last_error() then returns a thread-local message.
 * Outputs are caller-allocated (the reference allocates inside the op: sampling_kernel.cu:140,
 * ball_query_kernel.cu:107-109, interpolate_kernel.cu:109-110,196,301, grouping_kernel.cu:122 -- the Python
 * shim regnet_for_3d_grasping_b200/pn2_ext.py allocates with torch so the caching allocator / stream
 * semantics are kept).  Inputs are borrowed and never written.
 *
 * Section 1 replaces, one entry point each, the 7 functions of the reference's pybind module `pn2_ext`
 * (multi_model/utils/pn2_utils/csrc/main.cpp:6-14) and the 2 of `dgcnn_ext`
 * (multi_model/utils/pn2_utils/functions/csrc/main.cpp:3-6).  Tensors that the reference takes as
 * (B, 3, N) "any stride" at::Tensors are passed as base pointer + three element strides, so the permuted
 * views REGNet actually passes (score_network.py:46, pointnet2.py:89-90) need no copy.
 *
 * Section 2 is the fused path behind the module-level drop-in (multi_model.score_network.ScoreNetwork):
 * a native plan that runs the whole PointNet2Seg forward (utils/pointnet2.py:86-121) on device.
 *
 * Paths cited are relative to /root/reference/multi_model/utils/pn2_utils/ unless they start elsewhere.
 */
#ifndef REGNET_B200_H_
#define REGNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REGNET_OK 0
#define REGNET_EINVAL 1   /* bad argument (the reference's TORCH_CHECK / CHECK_EQ failures) */
#define REGNET_ECUDA 2    /* CUDA runtime / launch error */
#define REGNET_ELIMIT 3   /* size outside what this build supports (message says which) */

/* ---- library ------------------------------------------------------------------------------------------ */
const char* regnet_last_error(void);          /* thread-local, valid until the next failing call */
int regnet_abi_version(void);                 /* bumped on any signature change */
int regnet_device_arch(int* sm_major, int* sm_minor, int* sm_count);  /* of the current device */
/* The reference device-asserts on out-of-range gather/scatter indices (grouping_kernel.cu:89,
 * interpolate_kernel.cu:168,278).  Here such elements are skipped and a device flag is raised; this call
 * synchronises, returns REGNET_EINVAL if the flag was raised since the last call, and clears it. */
int regnet_check_index_errors(void);

/* ---- 1. pn2_ext operator ABI --------------------------------------------------------------------------- */

/* csrc/sampling.h:7-9 FarthestPointSample (kernel csrc/sampling_kernel.cu:47-117).
 * points: logical (B,3,N) fp32 with element strides (sb,sc,sn).  index: (B,M) int64, first pick = 0.
 * new_xyz (optional, may be NULL): (B,3,M) contiguous, the sampled coordinates (function.py:11-26 gather_points
 * fused).  Bit-exact with the reference including its tie rule.  EINVAL unless 0 < M <= N.  Any N: clouds of up to
 * 65 536 points stay register-resident in a thread-block cluster, larger ones take a generic one-CTA-per-cloud kernel
 * with the min-distance array in global memory (scratch from cudaMallocAsync on `stream`). */
int regnet_farthest_point_sample(const float* points, int64_t sb, int64_t sc, int64_t sn, int B, int N, int M,
                                 int64_t* index, float* new_xyz, void* stream);
/* same, with explicit tuning (cluster_size in {0=auto,1,2,4,8}, threads in {0=auto,256,512,1024}; a negative
 * thread count selects the barrier.cluster exchange instead of st.async+mbarrier, for A/B measurements) and an
 * optional int32 copy of the indices for the fused path.  Shapes without a compiled instance (the library carries the
 * ones its policy selects) are served by the generic kernel: same indices. */
int regnet_farthest_point_sample_ex(const float* points, int64_t sb, int64_t sc, int64_t sn, int B, int N, int M,
                                    int64_t* index64, int32_t* index32, float* new_xyz, int cluster_size,
                                    int threads, void* stream);

/* csrc/ball_query.h:7-11 BallQuery (kernel csrc/ball_query_kernel.cu:31-74).
 * points (B,3,N), centroids (B,3,M) strided as above; radius is a C float squared in fp32; strict d < r*r;
 * first K hits in ascending index order, first hit replicated into unused slots, no hit -> zeros.
 * index (B,M,K) int64, count (B,M) int64; index32 (optional) int32 copy.  Any K (K > 128 takes a generic kernel). */
int regnet_ball_query(const float* points, int64_t psb, int64_t psc, int64_t psn, const float* centroids,
                      int64_t csb, int64_t csc, int64_t csn, int B, int N, int M, float radius, int K,
                      int64_t* index, int64_t* count, int32_t* index32, void* stream);

/* Grid-accelerated forms of ball_query / point_search: identical results (indices, counts, squared distances), the
 * caller lends regnet_search_workspace_bytes(B, N) bytes of device scratch (N = points per cloud for ball_query, keys
 * per cloud for point_search).  Used when K == 64 (ball query), 4096 <= N <= 65536 and the workspace is big enough;
 * otherwise they fall through to regnet_ball_query / regnet_point_search.  Replace the same reference entry points
 * (csrc/ball_query_kernel.cu:87-131, csrc/interpolate_kernel.cu:88-128). */
int64_t regnet_search_workspace_bytes(int B, int N);
int regnet_ball_query_ws(const float* points, int64_t psb, int64_t psc, int64_t psn, const float* centroids,
                         int64_t csb, int64_t csc, int64_t csn, int B, int N, int M, float radius, int K,
                         int64_t* index, int64_t* count, void* workspace, int64_t workspace_bytes, void* stream);
int regnet_point_search_ws(const float* query, int64_t qsb, int64_t qsc, int64_t qsn, const float* key, int64_t ksb,
                           int64_t ksc, int64_t ksn, int B, int Nq, int Nk, int k, int64_t* index, float* distance,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* Double-precision instantiations of the three search operators (the reference dispatches over float and double:
 * sampling_kernel.cu:149, ball_query_kernel.cu:112, interpolate_kernel.cu:113).  Same semantics, arithmetic in double in
 * the reference's order; generic kernels (REGNet only passes float32, which the tuned kernels above serve). */
int regnet_farthest_point_sample_f64(const double* points, int64_t sb, int64_t sc, int64_t sn, int B, int N, int M,
                                     int64_t* index, void* stream);
int regnet_ball_query_f64(const double* points, int64_t psb, int64_t psc, int64_t psn, const double* centroids, int64_t csb,
                          int64_t csc, int64_t csn, int B, int N, int M, double radius, int K, int64_t* index,
                          int64_t* count, void* stream);
int regnet_point_search_f64(const double* query, int64_t qsb, int64_t qsc, int64_t qsn, const double* key, int64_t ksb,
                            int64_t ksc, int64_t ksn, int B, int Nq, int Nk, int k, int64_t* index, double* distance,
                            void* stream);

/* csrc/grouping.h:7-14 GroupPointsForward / GroupPointsBackward (csrc/grouping_kernel.cu:29-51, 54-149).
 * input (B,C,N) strided; index (B,M,K) int64 contiguous; out (B,C,M,K) contiguous.
 * backward: grad_out (B,C,M,K) contiguous -> grad_in (B,C,N) contiguous, zero-filled then scatter-added. */
int regnet_group_points_forward(const float* input, int64_t sb, int64_t sc, int64_t sn, const int64_t* index,
                                int B, int C, int N, int M, int K, float* out, void* stream);
int regnet_group_points_backward(const float* grad_out, const int64_t* index, int B, int C, int N, int M, int K,
                                 float* grad_in, void* stream);

/* csrc/interpolate.h:8-22 PointSearch (kernel csrc/interpolate_kernel.cu:28-77): exact 3-NN, strict '<'
 * insertion (earliest index wins ties), distances SQUARED.  EINVAL unless k == 3 and Nk >= 3.
 * index (B,Nq,3) int64, distance (B,Nq,3) fp32. */
int regnet_point_search(const float* query, int64_t qsb, int64_t qsc, int64_t qsn, const float* key, int64_t ksb,
                        int64_t ksc, int64_t ksn, int B, int Nq, int Nk, int k, int64_t* index, float* distance,
                        void* stream);

/* InterpolateForward / InterpolateBackward (csrc/interpolate_kernel.cu:134-232, 239-337).
 * input (B,C,Ns) strided, index/weight (B,Nd,3) contiguous -> out (B,C,Nd) contiguous;
 * backward: grad_out (B,C,Nd) contiguous -> grad_in (B,C,Ns) contiguous. */
int regnet_interpolate_forward(const float* input, int64_t sb, int64_t sc, int64_t sn, const int64_t* index,
                               const float* weight, int B, int C, int Ns, int Nd, float* out, void* stream);
int regnet_interpolate_backward(const float* grad_out, const int64_t* index, const float* weight, int B, int C,
                                int Ns, int Nd, float* grad_in, void* stream);

/* dgcnn_ext (functions/csrc/gather_knn.h:7-13): same maths as group_points with N' == N. */
int regnet_gather_knn_forward(const float* input, int64_t sb, int64_t sc, int64_t sn, const int64_t* index, int B,
                              int C, int N, int M, int K, float* out, void* stream);
int regnet_gather_knn_backward(const float* grad_out, const int64_t* index, int B, int C, int N, int M, int K,
                               float* grad_in, void* stream);

/* ---- 2. fused ScoreNet forward -------------------------------------------------------------------------- */

/* Engine of the shared-MLP contraction (nn/modules/mlp.py:55-106 = chained 1x1 conv + BN + ReLU):
 *   REGNET_ENGINE_TC    tcgen05 tensor cores, operands split into bf16 hi+lo planes, 3 products, fp32 accumulate
 *                       in TMEM (the product path)
 *   REGNET_ENGINE_SIMT  fp32 FFMA tiles (bring-up / cross-check engine, also CUDA, never a CPU path) */
#define REGNET_ENGINE_TC 0
#define REGNET_ENGINE_SIMT 1

#define REGNET_MAX_LAYERS 4

typedef struct regnet_scorenet_config {
  int32_t batch;               /* B */
  int32_t num_points;          /* N (25600 in REGNet, train.py:70) */
  int32_t num_centroids[3];    /* utils/pointnet2.py:40  (5120,1024,256) */
  float radius[3];             /* utils/pointnet2.py:41  (0.02,0.08,0.32) */
  int32_t num_neighbours[3];   /* utils/pointnet2.py:42  (64,64,64); must be 64 for the pooled epilogue */
  int32_t engine;              /* REGNET_ENGINE_* */
  int32_t use_side_stream;     /* 0: one stream; 1: geometry chain (FPS/ball query/3-NN) on an internal second stream;
                                  2: only FPS there (co-resides with the GEMM CTAs), the rest on the caller's stream;
                                  3: as 2, plus ball query of levels 1-2 and the 3-NN searches behind their FPS on a
                                     further internal stream (level 0's ball query stays on the caller's stream) */
} regnet_scorenet_config;

typedef struct regnet_scorenet regnet_scorenet;   /* opaque plan */

/* Create / destroy a plan.  The plan owns its workspace (cudaMalloc) and prepared weights. */
int regnet_scorenet_create(const regnet_scorenet_config* cfg, regnet_scorenet** out);
int regnet_scorenet_destroy(regnet_scorenet* plan);
int64_t regnet_scorenet_workspace_bytes(const regnet_scorenet* plan);

/* Fold + upload one conv/BN layer (eval mode: y = relu(scale * (W x) + shift)).
 * stage: 0-2 = sa_modules[i], 3-5 = fp_modules[i-3], 6 = seg mlp, 7 = score head (conv_score+bn_score, sigmoid).
 * weight: DEVICE fp32 (cout, cin) row-major (the conv weight with its trailing 1x1 dims dropped);
 * scale/shift: DEVICE fp32 (cout) = gamma/sqrt(var+eps), beta - mean*scale (+ scale*bias for the score head). */
int regnet_scorenet_set_layer(regnet_scorenet* plan, int stage, int layer, int cin, int cout, const float* weight,
                              const float* scale, const float* shift, void* stream);

/* Run the forward.  pc: (B,N,6) fp32 contiguous (xyz,rgb), as score_network.py:31-46 receives it.
 * all_feature: (B,N,256) fp32 contiguous (= the reference's all_feature values; the reference returns them as a
 * transposed view of (B,256,N), score_network.py:48).  score: (B,N) fp32.  Asynchronous on `stream`. */
int regnet_scorenet_forward(regnet_scorenet* plan, const float* pc, float* all_feature, float* score, void* stream);

/* Optional software pipelining across forwards (throughput mode).  The geometry chain of a forward (FPS, ball query,
 * 3-NN: xyz only, ~40 % of a step, dominated by FPS's sequential dependency) is enqueued on the plan's side stream for
 * the NEXT input while the MLPs of the current one run: call prefetch(pc_next) BEFORE forward(pc_current); the
 * following forward(pc_next) finds its geometry ready.  `pc` must stay valid and unchanged until that forward has
 * completed.  At most two prefetches may be outstanding.  Without a prefetch, forward computes the geometry itself. */
int regnet_scorenet_prefetch(regnet_scorenet* plan, const float* pc, void* stream);
/* Make `stream` wait for every outstanding prefetch (their side-stream work), e.g. before timing or reusing `pc`. */
int regnet_scorenet_join_prefetch(regnet_scorenet* plan, void* stream);

/* Plan options.  "defer_prefetch" (default 0): 1 parks a prefetch until the next forward has launched its level-0 kernel
 * (or until its results are needed) instead of enqueueing it at once -- an experiment switch, see csrc/scorenet.cu.  "dynamic_tiles": tile scheduling of the GEMM launches.
 * "sa_fused_a" (default 2): how the second layer of set-abstraction levels 1-2 gets its operand relu(Z'[g] + T) --
 * 1 = built inside the GEMM by producer warps (csrc/gemm_fused_a.cu), 3 = materialised by a gather-add pass (identical
 * bits), 2 = the former unless a prefetched FPS is co-running, 0 = the round-1 gather-affine producer (no folded tables). */
int regnet_scorenet_set_option(regnet_scorenet* plan, const char* name, int value);

/* The geometry chain alone (FPS, ball query and 3-NN of every level), for callers that run the MLPs themselves -- the
 * training path.  Consumes a matching regnet_scorenet_prefetch or computes the chain now; on return `stream` is ordered
 * behind every result, readable through regnet_scorenet_intermediate ("fps*", "xyz*", "bq*", "nn*", "nnw*") until the next
 * forward / geometry call that reuses the slot (two slots alternate). */
int regnet_scorenet_geometry(regnet_scorenet* plan, const float* pc, void* stream);

/* Read back intermediates of the last forward for parity tests (device pointers into the plan's workspace,
 * valid until the next forward).  what: "fps0".."fps2" int32 (B,M_i); "bq0".."bq2" int32 (B,M_i,64);
 * "nn0".."nn2" int32 (B,Nd_i,3); "nnw0".."nnw2" fp32 (B,Nd_i,3) interpolation weights; "xyz0".."xyz2" fp32 (B,3,M_i) sampled
 * coordinates; "sa0".."sa2" fp32 (B,M_i,C) point-major; "fp0".."fp2" fp32 point-major. */
int regnet_scorenet_intermediate(regnet_scorenet* plan, const char* what, void** ptr, int64_t* numel);

/* Per-launch timing for bench.py's roofline block.  With profiling on, the forward runs every kernel on the
 * caller's stream (no side stream) bracketed by CUDA events; regnet_scorenet_profile synchronises and writes one
 * "label milliseconds start_offset_ms\n" line per launch of the LAST forward into buf (labels: fps.i, ball_query.i, three_nn.i,
 * sa_operand.i, fp_operand.i, gemm.<stage>.l<j>[pool], score_head). */
/* on = 1: as above.  on = 2: timeline mode -- streams and prefetch stay as configured, events are recorded around
 * every launch and accumulate over forwards until regnet_scorenet_profile reads them (overlap diagnostics). */
int regnet_scorenet_set_profiling(regnet_scorenet* plan, int on);
int regnet_scorenet_profile(regnet_scorenet* plan, char* buf, int64_t buf_bytes);

/* Number of kernels launched by the last regnet_scorenet_forward (for bench.py's gpu_launches). */
int regnet_scorenet_launch_count(const regnet_scorenet* plan);

/* ---- 3. region stage ----------------------------------------------------- */

/* dataset_utils/get_regiondataset.py:354-434 _select_score_center, on device and batched.
 * pc (B,N,6), score (B,N).  Per cloud: positives = score > score_thre; more than center_num of them -> farthest point
 * sampling over the positives in index order (bit-exact with the reference's per-cloud pn2_ext call), mapped back to
 * cloud indices; 1..center_num -> all of them, then uniform repeats; none -> center_num distinct random points.
 * center_index (B,center_num) int64; center_pc (B,center_num,6); positive_count (B) int32, optional.
 * workspace: regnet_select_score_center_workspace(B,N,center_num) bytes of device scratch. */
int64_t regnet_select_score_center_workspace(int B, int N, int center_num);
int regnet_select_score_center(const float* pc, const float* score, int B, int N, int center_num, float score_thre,
                               uint64_t seed, int64_t* center_index, float* center_pc, int32_t* positive_count,
                               void* workspace, int64_t workspace_bytes, void* stream);

/* get_regiondataset.py:279-352 _get_local_points_batch + _get_group_pc: for every centre the points with
 * sqrt(dx^2+dy^2+dz^2) <= radius (non-strict, products rounded separately as torch does), then exactly group_num of
 * them: without replacement when count >= group_num, with replacement when 0 < count < group_num, -1 rows when empty.
 * index (B,NC,G) int64, group (B,NC,G,6) fp32, count (B,NC) int32 optional (number of points inside each ball). */
int regnet_ball_crop_sample(const float* pc, const float* center_pc, int B, int N, int NC, float radius, int group_num,
                            uint64_t seed, int64_t* index, float* group, int32_t* count, void* stream);

/* Grid form of regnet_ball_crop_sample: same membership test, counts and sampling distribution, candidates taken from a
 * uniform (x, y) grid of the cloud instead of all N points (the no-replacement picks are not in ascending index order).
 * workspace: regnet_ball_crop_workspace_bytes(B, N) bytes of device scratch; falls through to the scan for small clouds. */
int64_t regnet_ball_crop_workspace_bytes(int B, int N);
int regnet_ball_crop_sample_ws(const float* pc, const float* center_pc, int B, int N, int NC, float radius, int group_num,
                               uint64_t seed, int64_t* index, float* group, int32_t* count, void* workspace,
                               int64_t workspace_bytes, void* stream);

/* Closing-box membership of every point of every grasp's crop (gripper_region_network.py:505-528): points (M, G, C)
 * fp32 with xyz first, centre (M, 3), rot (M, 3, 3) = rows approach / axis_y / minor normal; a point is inside iff
 * 0 < x' < xlim, |y'| < ylim, |z'| < zlim for p' = rot (p - centre); xlim / ylim per grasp (M floats) when the row
 * pointers are non-null, else the scalars.  mask (M, G) bytes.  M <= 65535 per call. */
int regnet_closing_box_mask(const float* points, int M, int G, int C, const float* centre, const float* rot,
                            const float* xlim_row, const float* ylim_row, float xlim, float ylim, float zlim,
                            uint8_t* mask, void* stream);

/* multi_model/gripper_region_network.py:532-544: per-row sampler over a (rows,G) byte mask: more than K set -> K
 * without replacement (ascending), more than min_count -> K with replacement, else the row is rejected (-1).
 * index (rows,K) int64; count (rows) int32 optional. */
int regnet_mask_sample(const uint8_t* mask, int rows, int G, int K, int min_count, uint64_t seed, int64_t* index,
                       int32_t* count, void* stream);

/* gripper_region_network.py:389-395 + utils/pointnet2.py:161,167,223,232: out[b,c,:] = max over the G rows
 * feat[idx[b,c,g] + b*N, :] of a point-major (B*N, C) feature tensor (negative indices wrap as in the reference).
 * index (B,NC,G) int64; out (B,NC,C).  C % 4 == 0. */
int regnet_gather_max(const float* feat, const int64_t* index, int B, int N, int NC, int G, int C, float* out,
                      void* stream);

/* ---- 3b. training-mode pieces of the shared MLP ----------------------------------------------------------------------
 * BatchNorm with BATCH statistics + ReLU, forward and backward, over (B, C, L) fp32 contiguous tensors (torch's NCHW /
 * NCL layout; L = M*K for the 2-D blocks).  Replaces nn.BatchNorm{1,2}d(training) + nn.ReLU(inplace) of
 * nn/modules/conv.py:24-36,64-76 (cuDNN bn_fw_tr / bn_bw kernels + two elementwise passes in torch).
 *   forward : y = [relu](gamma * (x - mean_c) * invstd_c + beta), mean / biased variance over (B, L) per channel;
 *             running_mean / running_var (nullable) are updated in place like torch: (1-momentum)*old + momentum*new,
 *             unbiased variance; save_mean, save_invstd, scale = gamma*invstd, shift = beta - mean*scale (C floats each)
 *             are outputs the backward needs.  L must be a multiple of 4, B*C <= 65535; B*L == 1 -> EINVAL (torch:
 *             "Expected more than 1 value per channel when training").
 *   backward: dx, dgamma, dbeta from dy (gradient w.r.t. y) and the saved tensors; the ReLU mask is recomputed.
 * workspace: regnet_bn_workspace_bytes(B, C, L) bytes of device scratch, contents undefined afterwards. */
int64_t regnet_bn_workspace_bytes(int B, int C, int64_t L);
int regnet_bn_relu_train_forward(const float* x, int B, int C, int64_t L, const float* gamma, const float* beta, float eps,
                                 float momentum, int relu, float* running_mean, float* running_var, float* y,
                                 float* save_mean, float* save_invstd, float* scale, float* shift, void* workspace,
                                 int64_t workspace_bytes, void* stream);
int regnet_bn_relu_train_backward(const float* dy, const float* x, int B, int C, int64_t L, const float* save_mean,
                                  const float* save_invstd, const float* scale, const float* shift, int relu, float* dx,
                                  float* dgamma, float* dbeta, void* workspace, int64_t workspace_bytes, void* stream);
/* The pooled (last) block of a set-abstraction MLP: batch-statistics BN + ReLU + max over the 64 neighbours in one go
 * (conv.py:64-76 followed by modules.py:245).  x (B, C, M, 64) fp32 contiguous -> out (B, C, M), argmax (B, C, M) bytes;
 * the (B, C, M, 64) activation is never written, and backward reads its gradient from (dout, argmax) instead of a dense
 * tensor.  Other arguments as regnet_bn_relu_train_*; workspace regnet_bn_workspace_bytes(B, C, 64 * M). */
int regnet_bn_relu_max64_train_forward(const float* x, int B, int C, int64_t M, const float* gamma, const float* beta,
                                       float eps, float momentum, int relu, float* running_mean, float* running_var,
                                       float* out, uint8_t* argmax, float* save_mean, float* save_invstd, float* scale,
                                       float* shift, void* workspace, int64_t workspace_bytes, void* stream);
int regnet_bn_relu_max64_train_backward(const float* dout, const uint8_t* argmax, const float* x, int B, int C, int64_t M,
                                        const float* save_mean, const float* save_invstd, const float* scale,
                                        const float* shift, int relu, float* dx, float* dgamma, float* dbeta,
                                        void* workspace, int64_t workspace_bytes, void* stream);
/* Max over the innermost 64 elements of (rows, 64) fp32 (torch.max(x, 3)[0] of modules.py:245 for K = 64) with the
 * arg-max kept as one byte per row, and its backward (dx = dout at the arg-max, 0 elsewhere; every element written). */
int regnet_maxpool64_forward(const float* x, int64_t rows, float* out, uint8_t* argmax, void* stream);
int regnet_maxpool64_backward(const float* dout, const uint8_t* argmax, int64_t rows, float* dx, void* stream);

/* ---- 3c. training-path 1x1 convolutions on the tcgen05 engine (csrc/conv_train.cu) -------------------------------------
 * Replace what torch runs for nn.Conv1d / nn.Conv2d(kernel 1, bias=False) inside nn/modules/conv.py:24-36,64-76 in train
 * mode -- the cuDNN / cutlass forward, dgrad and wgrad kernels -- in torch's own (B, C, L) layout (L = M*K for 2-D blocks).
 * Operands are bf16 hi/lo PLANES (x ~= hi + lo, what regnet_split_planes / regnet_bn_apply_ex / regnet_bn_backward_ex
 * write): passes = 3 issues hi*hi + lo*hi + hi*lo (fp32 parity, ~1e-5), passes = 1 issues hi*hi only (plain bf16).
 * hi / lo arguments are device pointers to bf16 arrays of the stated shape. */

/* x (n) fp32 -> hi, lo (n) bf16. */
int regnet_split_planes(const float* x, int64_t n, void* hi, void* lo, void* stream);
/* W (rows, cols) fp32 row-major -> planes (rows, ld_out), or with transpose != 0 the planes of W^T (cols, ld_out);
 * zero padded, ld_out a multiple of 8.  (A conv weight (cout, cin, 1[, 1]) is W with rows = cout, cols = cin.) */
int regnet_split_weight(const float* W, int rows, int cols, int transpose, int ld_out, void* hi, void* lo, void* stream);
/* out[b, r, l] = sum_k A[r, k] x[b, k, l]:  x planes (B, K, L), A planes (rows, lda) with K valid columns, out (B, rows, L)
 * fp32.  fprop: A = W (rows = cout, K = cin); dgrad: A = W^T (rows = cin, K = cout) and x = the output gradient.
 * moments (nullable): (rows, 2) doubles receiving per-row sum and sum of squares of `out` over (b, l) -- the batch
 * statistics BatchNorm needs, accumulated in the epilogue (zeroed by this call).  L % 8 == 0, lda % 8 == 0. */
int regnet_conv1x1_train(const void* x_hi, const void* x_lo, int B, int K, int64_t L, const void* a_hi, const void* a_lo,
                         int rows, int lda, float* out, double* moments, int passes, void* stream);
/* dW[co, ci] = sum_{b,l} g[b, co, l] x[b, ci, l]: g planes (B, Co, L), x planes (B, Ci, L), dW (Co, Ci) fp32 (overwritten).
 * Split-K over (b, l) across the SMs, partial tiles reduced in a fixed order (deterministic).
 * workspace: regnet_conv1x1_wgrad_workspace_bytes(B, Co, Ci, L) bytes of device scratch. */
int64_t regnet_conv1x1_wgrad_workspace_bytes(int B, int Co, int Ci, int64_t L);
int regnet_conv1x1_train_wgrad(const void* g_hi, const void* g_lo, const void* x_hi, const void* x_lo, int B, int Co,
                               int Ci, int64_t L, float* dW, void* workspace, int64_t workspace_bytes, int passes,
                               void* stream);
/* BatchNorm (batch statistics) pieces for the chained MLP: statistics from the convolution's moments; apply with the
 * result as fp32 (y) and / or planes (y_hi, y_lo) and an optional dropout (F.dropout of nn/modules/mlp.py:101-105: a
 * counter-based mask regenerated from drop_seed in the backward, keep probability 1 - drop_p, kept values scaled by
 * 1 / (1 - drop_p)); backward with the gradient w.r.t. the convolution output as fp32 (dz) and / or planes. */
int regnet_bn_finalize_moments(const double* moments, int C, double count, const float* gamma, const float* beta, float eps,
                               float momentum, float* running_mean, float* running_var, float* save_mean,
                               float* save_invstd, float* scale, float* shift, void* stream);
int regnet_bn_apply_ex(const float* z, int B, int C, int64_t L, const float* scale, const float* shift, int relu,
                       float drop_p, uint64_t drop_seed, float* y, void* y_hi, void* y_lo, void* stream);
int regnet_bn_backward_ex(const float* dy, const float* z, int B, int C, int64_t L, const float* save_mean,
                          const float* save_invstd, const float* scale, const float* shift, int relu, float drop_p,
                          uint64_t drop_seed, float* dz, void* dz_hi, void* dz_lo, float* dgamma, float* dbeta,
                          void* workspace, int64_t workspace_bytes, void* stream);
/* dgrad with the reduction pass of the previous block's BatchNorm backward fused into its epilogue: out = W^T g as
 * regnet_conv1x1_train (rows = that block's channels), and sums (rows, 2) doubles = per-channel sum(h), sum(h * xhat) with
 * h = out * dropout-mask * [bn(z_prev) > 0] and xhat = (z_prev - mean) * invstd -- what regnet_bn_backward_ex computes in
 * a separate pass over out and z_prev.  regnet_bn_backward_from_sums then finishes that block's backward (dgamma, dbeta,
 * the gradient w.r.t. z_prev as fp32 and / or planes) with the apply pass only; workspace >= 2 * C floats. */
int regnet_conv1x1_train_dgrad_bnreduce(const void* g_hi, const void* g_lo, int B, int K, int64_t L, const void* a_hi,
                                        const void* a_lo, int rows, int lda, float* out, const float* z_prev,
                                        const float* mean, const float* invstd, const float* scale, const float* shift,
                                        int relu, float drop_p, uint64_t drop_seed, double* sums, int passes, void* stream);
int regnet_bn_backward_from_sums(const float* dy, const float* z, int B, int C, int64_t L, const float* save_mean,
                                 const float* save_invstd, const float* scale, const float* shift, int relu, float drop_p,
                                 uint64_t drop_seed, const double* sums, float* dz, void* dz_hi, void* dz_lo, float* dgamma,
                                 float* dbeta, void* workspace, int64_t workspace_bytes, void* stream);
/* the pooled block with given statistics: out[b,c,m] = max_k [relu](fma(z[b,c,m,k], scale, shift)), and its backward */
int regnet_bn_apply_max64(const float* z, int B, int C, int64_t M, const float* scale, const float* shift, int relu,
                          float* out, uint8_t* argmax, void* stream);
int regnet_bn_max64_backward_ex(const float* dout, const uint8_t* argmax, const float* z, int B, int C, int64_t M,
                                const float* save_mean, const float* save_invstd, const float* scale, const float* shift,
                                int relu, float* dz, void* dz_hi, void* dz_lo, float* dgamma, float* dbeta, void* workspace,
                                int64_t workspace_bytes, void* stream);

/* Operand producers of the training path (csrc/train_gather.cu): the input of a shared MLP written directly as planes.
 * regnet_sa_group_planes: QueryGrouper.forward (modules.py:39-56) -- out[b, 0:3, m, k] = xyz[b, :, index[b,m,k]] -
 * new_xyz[b, :, m], out[b, 3:, m, k] = feature[b, :, index[b,m,k]]; xyz (B,3,N) and feature (B,C,N) strided, new_xyz (B,3,M)
 * contiguous, index (B,M,K) int64, K % 4 == 0; planes (B, 3 + C, M, K).
 * regnet_fp_interp_planes: FeatureInterpolator.forward (modules.py:104-131) -- out[b, 0:C2, n] = sum_k sparse[b, :,
 * index[b,n,k]] * weight[b,n,k], out[b, C2:, n] = dense[b, :, n]; planes (B, C2 + C1, Nd), Nd % 4 == 0.
 * The two *_backward_strided calls are regnet_group_points_backward / regnet_interpolate_backward reading channels
 * [c0, c0 + C) of a (B, Ctot, L) gradient in place (batch_stride = Ctot * L elements); source rows of at most 12 288 points. */
int regnet_sa_group_planes(const float* xyz, int64_t xsb, int64_t xsc, int64_t xsn, const float* new_xyz, const float* feature,
                           int64_t fsb, int64_t fsc, int64_t fsn, const int64_t* index, int B, int C, int N, int M, int K,
                           void* out_hi, void* out_lo, void* stream);
int regnet_fp_interp_planes(const float* sparse, int64_t ssb, int64_t ssc, int64_t ssn, const float* dense, int64_t dsb,
                            int64_t dsc, int64_t dsn, const int64_t* index, const float* weight, int B, int C2, int C1, int Ns,
                            int Nd, void* out_hi, void* out_lo, void* stream);
int regnet_group_points_backward_strided(const float* grad_out, int64_t batch_stride, int c0, const int64_t* index, int B,
                                         int C, int N, int M, int K, float* grad_in, void* stream);
int regnet_interpolate_backward_strided(const float* grad_out, int64_t batch_stride, int c0, const int64_t* index,
                                        const float* weight, int B, int C, int Ns, int Nd, float* grad_in, void* stream);

/* First convolution of a set-abstraction / feature-propagation MLP applied per SOURCE point (linear operations commute
 * with grouping and interpolation; the eval plan does the same): the grouped pre-activation
 *   Z[b,c,m,k] = Y[b,c,index[b,m,k]] + sum_d Wx[c,d] * xyz_rel[b,d,m,k]            (Y = W_f f, a GEMM over the N source points)
 *   Z[b,c,n]   = sum_k weight[b,n,k] * Ys[b,c,index[b,n,k]] + sum_d Wd[c,d] * dense[b,d,n]   (Ys = W_s s; nd <= 4 dense channels)
 * in fp32 with its batch moments (C0, 2) fp64 -- what regnet_conv1x1_train would have produced on the grouped operand -- and
 * the matching backward pieces: dY = scatter-add of dZ by index with dWx_part[b,c,d] = sum dZ * xyz_rel, and
 * dWd_part[b,c,d] = sum_n dZ * dense.  Y, Z, dZ, dY contiguous (B, C0, .); xyz_rel (B, 3, M*K); rows of <= 12 288 points. */
int regnet_sa_gather_linear(const float* Y, const int64_t* index, const float* xyz_rel, const float* Wx, int ldwx, int B, int C0,
                            int N, int M, int K, float* Z, double* moments, void* stream);
int regnet_sa_scatter_linear(const float* dZ, const int64_t* index, const float* xyz_rel, int B, int C0, int N, int M, int K,
                             float* dY, float* dwx_part, void* stream);
int regnet_fp_gather_linear(const float* Ys, const int64_t* index, const float* weight, const float* dense, int64_t dsb,
                            int64_t dsc, int64_t dsn, int nd, const float* Wd, int ldwd, int B, int C0, int Ns, int Nd, float* Z,
                            double* moments, void* stream);
int regnet_fp_dense_wgrad(const float* dZ, const float* dense, int64_t dsb, int64_t dsc, int64_t dsn, int nd, int B, int C0, int Nd,
                          float* part, void* stream);

/* First block of a set-abstraction MLP whose grouped input has 6 channels ([xyz - centre | 3 feature channels], level 0 of
 * ScoreNet) WITHOUT materialising its pre-activation: Z0 is linear in 6 numbers per position, so (csrc/train_gather.cu)
 *   regnet_sa0_input_moments   sums27 = [sum x_d (6) | sum x_d1 x_d2, d1 <= d2 (21)] over all grouped positions, fp64, and
 *                              (moments != NULL) the batch moments (C0, 2) = (sum z, sum z^2) of Z0 that follow from them;
 *   regnet_sa0_apply_planes    y = [relu](scale * (W0 x) + shift) recomputed from the inputs, written as planes (B, C0, M*K);
 *   regnet_sa0_backward_sums   G (C0, 7) fp64 = [sum g | sum g x_d] with g = dy * [scale * (W0 x) + shift > 0]: everything the
 *                              block's backward needs (dgamma, dbeta, dW0) in closed form; its inputs need no gradient.
 * xyz (B,3,N), feature (B,3,N) strided; new_xyz (B,3,M) contiguous; index (B,M,K) int64, K % 8 == 0; W0 (C0, 6) row-major in
 * [xyz | feature] order; dy (B, C0, M*K) fp32. */
int regnet_sa0_input_moments(const float* xyz, int64_t xsb, int64_t xsc, int64_t xsn, const float* new_xyz, const float* feature,
                             int64_t fsb, int64_t fsc, int64_t fsn, const int64_t* index, int B, int N, int M, int K,
                             const float* W0, int C0, double* sums27, double* moments, void* stream);
/* dbeta, dgamma and dW0 (C0, 6; may be null) of that block in closed form from G (regnet_sa0_backward_sums), the 27 input sums,
 * the weights and the batch statistics (invstd, scale = gamma * invstd); count = B * M * K. */
int regnet_sa0_backward_finalize(const double* G, const double* sums27, const float* W0, const float* invstd, const float* scale,
                                 int C0, double count, float* dW0, float* dgamma, float* dbeta, void* stream);
int regnet_sa0_apply_planes(const float* xyz, int64_t xsb, int64_t xsc, int64_t xsn, const float* new_xyz, const float* feature,
                            int64_t fsb, int64_t fsc, int64_t fsn, const int64_t* index, int B, int N, int M, int K,
                            const float* W0, const float* scale, const float* shift, int C0, int relu, void* y_hi, void* y_lo,
                            void* stream);
int regnet_sa0_backward_sums(const float* xyz, int64_t xsb, int64_t xsc, int64_t xsn, const float* new_xyz, const float* feature,
                             int64_t fsb, int64_t fsc, int64_t fsn, const int64_t* index, int B, int N, int M, int K,
                             const float* dy, const float* W0, const float* scale, const float* shift, int C0, int relu,
                             double* G, void* stream);

/* ---- 4. building blocks exposed for tests / micro-benchmarks ------------------------------------------- */

/* Y = act(scale * (X W^T) + shift), X (P,cin) fp32 row-major, W (cout,cin) fp32 row-major.
 * pool == 0: Y (P,cout) fp32 row-major.  pool > 0 (must be 64): Y (P/pool,cout) = max over each run of `pool` rows.
 * act: 0 none, 1 relu, 2 sigmoid.  engine as above.  Stand-alone (allocates + frees scratch internally; not for
 * hot loops -- the plan uses pre-split operands). */
int regnet_mlp_layer(const float* X, const float* W, const float* scale, const float* shift, int64_t P, int cin,
                     int cout, int pool, int act, int engine, float* Y, void* stream);

/* The same contraction on caller-owned bf16 hi/lo planes, without allocation or synchronisation (tcgen05 engine):
 * X planes (P, ldx) with K valid columns, W planes (cout, ldw); Y = act(scale * (X W^T) + shift) as fp32 (P, ld_f32) and / or
 * planes (P, ld_split); ldx, ldw, ld_split multiples of 8, ld_f32 a multiple of 4.  Used by the region / refine heads
 * (multi_model/utils/pointnet2.py:165-197, 227-254 with BatchNorm folded in eval mode). */
int regnet_linear_planes(const void* x_hi, const void* x_lo, int ldx, int64_t P, int K, const void* w_hi, const void* w_lo,
                         int ldw, int cout, const float* scale, const float* shift, int act, float* out_f32, int ld_f32,
                         void* out_hi, void* out_lo, int ld_split, void* stream);

/* Set-abstraction level 0 as ONE kernel (csrc/sa0_chain.cu): group rgb / xyz by `nbr`, subtract the centroid, the three
 * 1x1 conv + BN + ReLU blocks 6 -> 128 -> 128 -> 256 and the max over each centroid's 64 neighbours
 * (pn2_utils/modules.py:44-52,241-245 with the channel plan of pointnet2.py:43).  pc (B,N,6) fp32 [xyz|rgb];
 * new_xyz (B,3,M) planar; nbr (B,M,64) int32; W0 (128,6) in OPERAND order [rgb | xyz - centroid], W1 (128,128),
 * W2 (256,128) row-major fp32; scale/shift = folded BatchNorm; out (B*M,256).  dbg: NULL, or (B*M*64, 256) receiving the
 * raw accumulators of layers 0 and 1 when variant == 1 (tests), or (148,5,8) int64 wait counters when variant == 2
 * (scripts/sa0_chain_timing.py); variant 0 = production instance.  Stand-alone and synchronising (tests); the plan
 * calls the kernel with pre-split weights. */
int regnet_sa0_chain(const float* pc, const float* new_xyz, const int32_t* nbr, int B, int N, int M, const float* W0,
                     const float* scale0, const float* shift0, const float* W1, const float* scale1,
                     const float* shift1, const float* W2, const float* scale2, const float* shift2, float* out,
                     float* dbg, int variant, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REGNET_B200_H_ */
