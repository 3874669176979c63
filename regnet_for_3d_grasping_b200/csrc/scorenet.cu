// Native plan for the whole ScoreNet (PointNet2Seg) forward -- section 2 of include/regnet_b200.h.
//
// Restates, B200-first, the call chain of /root/reference/multi_model/utils/pointnet2.py:86-121 :
//   3 x PointNetSAModule (pn2_utils/modules.py:210-246): FPS -> gather -> ball query -> group/concat -> SharedMLP -> max
//   3 x PointnetFPModule (pn2_utils/modules.py:500-509, 104-131): 3-NN -> weights -> interpolate/concat -> SharedMLP
//   seg SharedMLP + conv_score/bn_score/sigmoid (pointnet2.py:116-119)
// Differences from the reference's execution (not its results):
//   * activations are point-major (rows = points/positions, channels contiguous) so every gather reads a row;
//   * the geometry chain (FPS / ball query / 3-NN depends on xyz only) runs on a second stream, overlapped with
//     the MLP GEMMs of the previous level;
//   * group+concat and interpolate+concat write the GEMM operand once, already split for the tensor-core engine;
//   * BN (eval) + ReLU + the 64-neighbour max-pool are GEMM epilogues; nothing else touches the activations.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "internal.cuh"

using namespace regnet;

namespace {

const int SA_CH[3][3] = {{128, 128, 256}, {256, 256, 512}, {512, 512, 1024}};  // pointnet2.py:43
const int FP_CH[3][3] = {{1024, 1024, 0}, {512, 512, 0}, {256, 256, 256}};       // pointnet2.py:44
const int FP_NL[3] = {2, 2, 3};
const int SEG_CH[4] = {512, 256, 256, 128};                                       // pointnet2.py:46

const char* const GEMM_LABEL[7][4] = {
    {"gemm.sa0.l0", "gemm.sa0.l1", "gemm.sa0.l2pool", ""}, {"gemm.sa1.l0", "gemm.sa1.l1", "gemm.sa1.l2pool", ""},
    {"gemm.sa2.l0", "gemm.sa2.l1", "gemm.sa2.l2pool", ""}, {"gemm.fp0.l0", "gemm.fp0.l1", "", ""},
    {"gemm.fp1.l0", "gemm.fp1.l1", "", ""},                {"gemm.fp2.l0", "gemm.fp2.l1", "gemm.fp2.l2", ""},
    {"gemm.seg.l0", "gemm.seg.l1", "gemm.seg.l2", "gemm.seg.l3"}};
const char* const FPS_LABEL[3] = {"fps.0", "fps.1", "fps.2"};
const char* const BQ_LABEL[3] = {"ball_query.0", "ball_query.1", "ball_query.2"};
const char* const NN_LABEL[3] = {"three_nn.0", "three_nn.1", "three_nn.2"};
const char* const SAOP_LABEL[3] = {"sa_operand.0", "sa_operand.1", "sa_operand.2"};
const char* const FPOP_LABEL[3] = {"fp_operand.0", "fp_operand.1", "fp_operand.2"};

struct Layer {
  int cin = 0, cout = 0, kpad = 0;
  float* w_f32 = nullptr;           // (cout, kpad) zero padded
  __nv_bfloat16* w_hi = nullptr;    // (cout, kpad)
  __nv_bfloat16* w_lo = nullptr;
  float* scale = nullptr;
  float* shift = nullptr;
  bool set = false;
};

struct Act {  // an activation matrix (rows, ld) in the layout of the selected engine
  float* f32 = nullptr;
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int ld = 0;
};

}  // namespace

struct regnet_scorenet {
  struct Cfg : regnet_scorenet_config { int fuse_sa0 = 3; int sa0_variant = 0; int gather_a = 1; int fp_linear_first = 1; int sa_linear_first = 1; int dynamic_tiles = 1; int use_grid = 1; int corun_cs = 8; int corun_threads = 128; int corun_small = 1; int corun1_cs = 8; int corun1_threads = 128; int corun_single = 0; int sa_fused_a = 2; } cfg;
  void* grid_ws[2] = {nullptr, nullptr};   // [0]: level-0 points, [1]: level-1 points (rebuilt per use)
  unsigned int* tile_counters = nullptr;   // one zeroed counter per GEMM launch of a forward (dynamic tile scheduling)
  int gemm_idx = 0;
  int pool_planes_level = -1;              // >= 0 while the pooled layer of that SA level is being launched
  const void* score_head = nullptr;        // Layer* of the score head while the last seg layer is being launched
  float* score_out = nullptr;
  int B = 0, N = 0;
  int M[3] = {0, 0, 0};
  Layer layers[8][REGNET_MAX_LAYERS];
  int nlayers[8] = {3, 3, 3, 2, 2, 3, 4, 1};
  // geometry results, double buffered: geometry of forward i+1 (side stream) may run while the MLPs of forward i
  // (caller's stream) still read slot i%2 -- see regnet_scorenet_prefetch
  struct Geom {
    int32_t* fps_idx[3] = {nullptr, nullptr, nullptr};
    float* new_xyz[3] = {nullptr, nullptr, nullptr};   // (B,3,M_i) planar
    int32_t* nbr[3] = {nullptr, nullptr, nullptr};     // (B,M_i,64)
    int32_t* nn_idx[3] = {nullptr, nullptr, nullptr};  // (B,Nd_i,3)
    float* nn_w[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_bq[3] = {nullptr, nullptr, nullptr}, ev_nn = nullptr;
    const float* pc = nullptr;   // input this slot was computed for (prefetched and not yet consumed)
    bool pending = false;
  } geom[2];
  // defer_prefetch = 1 (default): a prefetch is not enqueued at once but parked until the next forward has launched its
  // level-0 kernel (sa0_chain needs all but 2.8 KB of an SM's shared memory, the multi-pick FPS needs 5 KB -- they cannot
  // share an SM), or until something needs its results (flush_deferred); 2 / 3 park it until SA level 1 / 2 is enqueued,
  // 0 enqueues at once.  With the round-2 multi-pick FPS (1.5 ms alone) the parked multi-pick prefetch beats the one-pick
  // FPS co-running from the start of the step: 6.77 vs 7.01 ms per step (before the replay loop was tightened, 7.24 vs 7.26).
  const float* deferred_pc = nullptr;
  int deferred_slot = -1;
  int defer_prefetch = 1;
  int next_slot = 0;   // slot the next geometry pass writes
  int last_slot = 0;   // slot the last forward consumed (regnet_scorenet_intermediate)
  // features (fp32, point-major)
  float* sa_out[3] = {nullptr, nullptr, nullptr};    // (B,M_i,C_i)
  // tcgen05 engine: sa_out[0], sa_out[1] also as bf16 hi/lo planes = gather tables of the next level's first layer, and
  // the (P,16) xyz - centroid planes of the level being computed (gemm_tc.cu GatherA)
  __nv_bfloat16* sa_hi[3] = {nullptr, nullptr, nullptr};
  __nv_bfloat16* sa_lo[3] = {nullptr, nullptr, nullptr};
  // FP modules with the first convolution applied before the interpolation (gather.cu fp_interp_affine_kernel):
  // fp_out[0], fp_out[1] as planes (the next module's sparse operand), Y = sparse W_s^T and D = dense W_d^T in fp32
  __nv_bfloat16* fp_hi[2] = {nullptr, nullptr};
  __nv_bfloat16* fp_lo[2] = {nullptr, nullptr};
  float* fp_y = nullptr;
  float* fp_d = nullptr;
  float* sa_t = nullptr;   // SA levels 1, 2: T = shift0 - scale0 * W_x centre per centroid (gemm_fused_a.cu)
  float* sa_z = nullptr;   // SA levels 1, 2: Z = previous level's features x W_f^T, per point (gather.cu sa_gather_affine_kernel)
  __nv_bfloat16* xyzrel_hi = nullptr;
  __nv_bfloat16* xyzrel_lo = nullptr;
  float* fp_out[2] = {nullptr, nullptr};             // fp0 (B,M1,1024), fp1 (B,M0,512); fp2 is the caller's buffer
  float* last_allfeat = nullptr;
  // ping-pong activation arenas
  unsigned char* arena[2] = {nullptr, nullptr};
  size_t arena_bytes = 0;
  std::vector<void*> allocs;
  size_t total_bytes = 0;
  cudaStream_t side = nullptr;
  cudaStream_t side2 = nullptr;            // FPS of levels 1, 2 in FPS-only mode: off the level-0 chain (see geometry_enqueue)
  cudaEvent_t ev_start = nullptr;
  int launches = 0;
  int prefetch_launches = 0;
  // optional per-launch timing (regnet_scorenet_set_profiling): events around every launch, serial execution
  int profiling = 0;   // 0 off; 1 serial per-launch timing; 2 timeline (keeps streams / prefetch, records events)
  struct Rec { const char* label; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  size_t nrec = 0;
};

namespace {

int dalloc(regnet_scorenet* p, void** out, size_t bytes) {
  bytes = (bytes + 255) / 256 * 256;
  RN_CUDA(cudaMalloc(out, bytes));
  p->allocs.push_back(*out);
  p->total_bytes += bytes;
  return REGNET_OK;
}

// profiling brackets: no-ops unless enabled
void prof_begin(regnet_scorenet* p, const char* label, cudaStream_t s) {
  if (!p->profiling) return;
  if (p->nrec == p->recs.size()) {
    regnet_scorenet::Rec r;
    r.label = label;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    p->recs.push_back(r);
  }
  p->recs[p->nrec].label = label;
  cudaEventRecord(p->recs[p->nrec].a, s);
}
void prof_end(regnet_scorenet* p, cudaStream_t s) {
  if (!p->profiling) return;
  cudaEventRecord(p->recs[p->nrec].b, s);
  ++p->nrec;
}

Act make_act(const regnet_scorenet* p, int which, int64_t rows, int ld) {
  Act a;
  a.ld = ld;
  if (p->cfg.engine == REGNET_ENGINE_SIMT) {
    a.f32 = reinterpret_cast<float*>(p->arena[which]);
  } else {
    a.hi = reinterpret_cast<__nv_bfloat16*>(p->arena[which]);
    a.lo = a.hi + (size_t)rows * ld;
  }
  return a;
}

// one conv+BN+act layer: in -> (out_act and/or out_f32), optional 64-row pooling
int run_layer_impl(regnet_scorenet* p, const Layer& L, const Act& in, int64_t P, int act, int pool, const Act* out_act,
                   float* out_f32, int ld_f32, cudaStream_t s);

int run_layer(regnet_scorenet* p, const char* label, const Layer& L, const Act& in, int64_t P, int act, int pool,
              const Act* out_act, float* out_f32, int ld_f32, cudaStream_t s) {
  prof_begin(p, label, s);
  const int rc = run_layer_impl(p, L, in, P, act, pool, out_act, out_f32, ld_f32, s);
  prof_end(p, s);
  return rc;
}

int run_layer_impl(regnet_scorenet* p, const Layer& L, const Act& in, int64_t P, int act, int pool, const Act* out_act,
                   float* out_f32, int ld_f32, cudaStream_t s) {
  if (!L.set) {
    set_error("scorenet: a layer (cin=%d) was never given weights (regnet_scorenet_set_layer)", L.cin);
    return REGNET_EINVAL;
  }
  Epilogue ep;
  ep.scale = L.scale; ep.shift = L.shift; ep.act = act; ep.pool = pool;
  ep.out_f32 = out_f32; ep.ld_f32 = ld_f32;
  ++p->launches;
  if (p->cfg.engine == REGNET_ENGINE_SIMT) {
    if (out_act) {
      if (out_f32 && out_f32 != out_act->f32) {
        set_error("scorenet: SIMT engine cannot write two fp32 outputs");
        return REGNET_EINVAL;
      }
      ep.out_f32 = out_act->f32; ep.ld_f32 = out_act->ld;
    }
    return gemm_simt_launch(in.f32, in.ld, L.w_f32, L.kpad, P, L.kpad, L.cout, ep, s);
  }
  if (out_act) { ep.out_hi = out_act->hi; ep.out_lo = out_act->lo; ep.ld_split = out_act->ld; }
  if (p->score_head) {
    const Layer* H = static_cast<const Layer*>(p->score_head);
    ep.dot_w = H->w_f32; ep.dot_scale = H->scale; ep.dot_shift = H->shift; ep.dot_out = p->score_out;
  }
  if (pool && p->pool_planes_level >= 0 && p->sa_hi[p->pool_planes_level]) {
    ep.pool_hi = p->sa_hi[p->pool_planes_level];
    ep.pool_lo = p->sa_lo[p->pool_planes_level];
    ep.ld_pool = L.cout;
  }
  if (p->cfg.dynamic_tiles && p->gemm_idx < 64) ep.tile_counter = p->tile_counters + (p->gemm_idx++);
  return gemm_tc_launch(in.hi, in.lo, in.ld, L.w_hi, L.w_lo, L.kpad, P, L.cin, L.cout, ep, s);
}

}  // namespace

extern "C" {

int regnet_scorenet_create(const regnet_scorenet_config* cfg, regnet_scorenet** out) {
  RN_CHECK_ARG(cfg && out, "scorenet_create: null argument");
  RN_CHECK_ARG(cfg->batch > 0 && cfg->num_points > 0, "scorenet_create: empty batch");
  RN_CHECK_ARG(cfg->engine == REGNET_ENGINE_TC || cfg->engine == REGNET_ENGINE_SIMT, "scorenet_create: unknown engine");
  int prev = cfg->num_points;
  for (int i = 0; i < 3; ++i) {
    RN_CHECK_ARG(cfg->num_centroids[i] > 0 && cfg->num_centroids[i] <= prev,
                 "scorenet_create: num_centroids[%d]=%d must be in (0, %d]", i, cfg->num_centroids[i], prev);
    RN_CHECK_ARG(cfg->num_neighbours[i] == 64, "scorenet_create: num_neighbours must be 64 (pooled epilogue)");
    RN_CHECK_ARG(cfg->radius[i] > 0.f, "scorenet_create: radius must be > 0");
    prev = cfg->num_centroids[i];
  }
  RN_CHECK_ARG(cfg->num_centroids[2] >= 3, "scorenet_create: the coarsest level needs >= 3 points for 3-NN");
  if (cfg->engine == REGNET_ENGINE_TC && !gemm_tc_supported()) {
    set_error("scorenet_create: tensor-map driver entry point unavailable; the tcgen05 engine cannot run here");
    return REGNET_ECUDA;
  }
  regnet_scorenet* p = new regnet_scorenet();
  static_cast<regnet_scorenet_config&>(p->cfg) = *cfg;
  if (const char* e = getenv("REGNET_FUSE_SA0")) p->cfg.fuse_sa0 = atoi(e);
  if (const char* e = getenv("REGNET_SA0_VARIANT")) p->cfg.sa0_variant = atoi(e);
  if (const char* e = getenv("REGNET_GATHER_A")) p->cfg.gather_a = atoi(e);
  if (const char* e = getenv("REGNET_FP_LINEAR_FIRST")) p->cfg.fp_linear_first = atoi(e);
  if (const char* e = getenv("REGNET_SA_LINEAR_FIRST")) p->cfg.sa_linear_first = atoi(e);
  if (const char* e = getenv("REGNET_DYNAMIC_TILES")) p->cfg.dynamic_tiles = atoi(e);
  if (const char* e = getenv("REGNET_SA_FUSED_A")) p->cfg.sa_fused_a = atoi(e);
  if (const char* e = getenv("REGNET_USE_GRID")) p->cfg.use_grid = atoi(e);
  if (const char* e = getenv("REGNET_FPS_CORUN")) sscanf(e, "%d,%d", &p->cfg.corun_cs, &p->cfg.corun_threads);
  if (const char* e = getenv("REGNET_FPS_CORUN_SMALL")) p->cfg.corun_small = atoi(e);
  if (const char* e = getenv("REGNET_FPS_CORUN1")) sscanf(e, "%d,%d", &p->cfg.corun1_cs, &p->cfg.corun1_threads);
  if (const char* e = getenv("REGNET_DEFER_PREFETCH")) p->defer_prefetch = atoi(e);
  if (const char* e = getenv("REGNET_FPS_CORUN_SINGLE")) p->cfg.corun_single = atoi(e);
  if (const char* e = getenv("REGNET_SIDE_MODE")) { if (p->cfg.use_side_stream >= 2) p->cfg.use_side_stream = atoi(e); }
  p->B = cfg->batch;
  p->N = cfg->num_points;
  for (int i = 0; i < 3; ++i) p->M[i] = cfg->num_centroids[i];
  const int B = p->B, N = p->N;
  const int* M = p->M;
  int rc = REGNET_OK;
  auto A = [&](void** ptr, size_t bytes) { if (!rc) rc = dalloc(p, ptr, bytes); };
  const int nd[3] = {M[1], M[0], N};  // dense point counts of fp0, fp1, fp2
  for (int g = 0; g < 2; ++g) {
    regnet_scorenet::Geom& G = p->geom[g];
    for (int i = 0; i < 3; ++i) {
      A((void**)&G.fps_idx[i], sizeof(int32_t) * (size_t)B * M[i]);
      A((void**)&G.new_xyz[i], sizeof(float) * (size_t)B * 3 * M[i]);
      A((void**)&G.nbr[i], sizeof(int32_t) * (size_t)B * M[i] * 64);
      A((void**)&G.nn_idx[i], sizeof(int32_t) * (size_t)B * nd[i] * 3);
      A((void**)&G.nn_w[i], sizeof(float) * (size_t)B * nd[i] * 3);
    }
  }
  for (int i = 0; i < 3; ++i) A((void**)&p->sa_out[i], sizeof(float) * (size_t)B * M[i] * SA_CH[i][2]);
  if (p->cfg.engine == REGNET_ENGINE_TC) {
    for (int i = 0; i < 3; ++i) {
      A((void**)&p->sa_hi[i], sizeof(__nv_bfloat16) * (size_t)B * M[i] * SA_CH[i][2]);
      A((void**)&p->sa_lo[i], sizeof(__nv_bfloat16) * (size_t)B * M[i] * SA_CH[i][2]);
    }
    A((void**)&p->fp_hi[0], sizeof(__nv_bfloat16) * (size_t)B * M[1] * 1024);
    A((void**)&p->fp_lo[0], sizeof(__nv_bfloat16) * (size_t)B * M[1] * 1024);
    A((void**)&p->fp_hi[1], sizeof(__nv_bfloat16) * (size_t)B * M[0] * 512);
    A((void**)&p->fp_lo[1], sizeof(__nv_bfloat16) * (size_t)B * M[0] * 512);
    // Y: (B*M2,1024) | (B*M1,512) | (B*M0,256);  D: (B*M1,1024) | (B*M0,512)
    const size_t ymax = std::max({(size_t)B * M[2] * 1024, (size_t)B * M[1] * 512, (size_t)B * M[0] * 256});
    const size_t dmax = std::max((size_t)B * M[1] * 1024, (size_t)B * M[0] * 512);
    A((void**)&p->fp_y, sizeof(float) * ymax);
    A((void**)&p->fp_d, sizeof(float) * dmax);
    A((void**)&p->sa_z, sizeof(float) * std::max((size_t)B * M[0] * SA_CH[1][0], (size_t)B * M[1] * SA_CH[2][0]));
    A((void**)&p->sa_t, sizeof(float) * std::max((size_t)B * M[1] * SA_CH[1][0], (size_t)B * M[2] * SA_CH[2][0]));
    const size_t pmax = (size_t)B * std::max(M[1], M[2]) * 64;
    A((void**)&p->xyzrel_hi, sizeof(__nv_bfloat16) * pmax * 16);
    A((void**)&p->xyzrel_lo, sizeof(__nv_bfloat16) * pmax * 16);
  }
  A((void**)&p->tile_counters, sizeof(unsigned int) * 64);
  A(&p->grid_ws[0], (size_t)grid_workspace_bytes(B, N));
  A(&p->grid_ws[1], (size_t)grid_workspace_bytes(B, M[0]));
  A((void**)&p->fp_out[0], sizeof(float) * (size_t)B * M[1] * 1024);
  A((void**)&p->fp_out[1], sizeof(float) * (size_t)B * M[0] * 512);
  // activation arenas: the widest (rows x ld) any layer reads or writes, 4 bytes per element in both engines
  size_t need = 0;
  auto upd = [&](int64_t rows, int ld) { need = std::max(need, (size_t)rows * (size_t)ld * 4); };
  for (int i = 0; i < 3; ++i) {
    const int64_t P = (int64_t)B * M[i] * 64;
    const int cin = (i == 0 ? 3 : SA_CH[i - 1][2]) + 3;
    upd(P, round_up(cin, 16));
    upd(P, SA_CH[i][0]);
    upd(P, SA_CH[i][1]);
  }
  upd((int64_t)B * M[1], 1536); upd((int64_t)B * M[1], 1024);
  upd((int64_t)B * M[0], 1280); upd((int64_t)B * M[0], 512);
  upd((int64_t)B * N, round_up(515, 16)); upd((int64_t)B * N, 512);
  p->arena_bytes = need;
  A((void**)&p->arena[0], need);
  A((void**)&p->arena[1], need);
  if (!rc && p->cfg.use_side_stream) {
    cudaError_t e = cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking);
    if (e == cudaSuccess && p->cfg.use_side_stream >= 2 && !getenv("REGNET_NO_SIDE2"))
      e = cudaStreamCreateWithFlags(&p->side2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming);
    for (int g = 0; g < 2; ++g) {
      for (int i = 0; i < 3 && e == cudaSuccess; ++i)
        e = cudaEventCreateWithFlags(&p->geom[g].ev_bq[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->geom[g].ev_nn, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "stream/event creation");
  }
  if (rc) {
    regnet_scorenet_destroy(p);
    return rc;
  }
  *out = p;
  return REGNET_OK;
}

int regnet_scorenet_destroy(regnet_scorenet* p) {
  if (!p) return REGNET_OK;
  for (void* a : p->allocs) cudaFree(a);
  if (p->ev_start) cudaEventDestroy(p->ev_start);
  for (int g = 0; g < 2; ++g) {
    for (int i = 0; i < 3; ++i) if (p->geom[g].ev_bq[i]) cudaEventDestroy(p->geom[g].ev_bq[i]);
    if (p->geom[g].ev_nn) cudaEventDestroy(p->geom[g].ev_nn);
  }
  if (p->side) cudaStreamDestroy(p->side);
  if (p->side2) cudaStreamDestroy(p->side2);
  for (auto& r : p->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  delete p;
  return REGNET_OK;
}

int64_t regnet_scorenet_workspace_bytes(const regnet_scorenet* p) { return p ? (int64_t)p->total_bytes : 0; }

int regnet_scorenet_launch_count(const regnet_scorenet* p) { return p ? p->launches : 0; }

int regnet_scorenet_set_layer(regnet_scorenet* p, int stage, int layer, int cin, int cout, const float* weight,
                              const float* scale, const float* shift, void* stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  RN_CHECK_ARG(p && weight && scale && shift, "scorenet_set_layer: null argument");
  RN_CHECK_ARG(stage >= 0 && stage < 8 && layer >= 0 && layer < p->nlayers[stage],
               "scorenet_set_layer: no layer %d in stage %d", layer, stage);
  int ecin, ecout;
  if (stage < 3) {
    ecin = layer == 0 ? (stage == 0 ? 3 : SA_CH[stage - 1][2]) + 3 : SA_CH[stage][layer - 1];
    ecout = SA_CH[stage][layer];
  } else if (stage < 6) {
    const int f = stage - 3;
    const int sparse_c = f == 0 ? SA_CH[2][2] : FP_CH[f - 1][FP_NL[f - 1] - 1];
    const int dense_c = f == 0 ? SA_CH[1][2] : f == 1 ? SA_CH[0][2] : 3;
    ecin = layer == 0 ? sparse_c + dense_c : FP_CH[f][layer - 1];
    ecout = FP_CH[f][layer];
  } else if (stage == 6) {
    ecin = layer == 0 ? 256 : SEG_CH[layer - 1];
    ecout = SEG_CH[layer];
  } else {
    ecin = 128;
    ecout = 1;
  }
  RN_CHECK_ARG(cin == ecin && cout == ecout, "scorenet_set_layer: stage %d layer %d expects (%d -> %d), got (%d -> %d)",
               stage, layer, ecin, ecout, cin, cout);
  Layer& L = p->layers[stage][layer];
  if (!L.w_f32) {
    L.cin = cin; L.cout = cout; L.kpad = round_up(cin, 16);
    RN_TRY(dalloc(p, (void**)&L.w_f32, sizeof(float) * (size_t)cout * L.kpad));
    RN_TRY(dalloc(p, (void**)&L.w_hi, sizeof(__nv_bfloat16) * (size_t)cout * L.kpad));
    RN_TRY(dalloc(p, (void**)&L.w_lo, sizeof(__nv_bfloat16) * (size_t)cout * L.kpad));
    RN_TRY(dalloc(p, (void**)&L.scale, sizeof(float) * cout));
    RN_TRY(dalloc(p, (void**)&L.shift, sizeof(float) * cout));
  }
  // SA first layers: the operand is [feature | xyz_rel] (gather.cu), the reference's conv weight is [xyz_rel | feature]
  const int rot = (stage < 3 && layer == 0) ? 3 : 0;
  // max-pooled layers (last layer of every SA stage): rows with a negative BN scale are negated together with the scale,
  // scale * (w . x) == |scale| * ((-w) . x); with scale >= 0 the pooled epilogues take the max on raw accumulators
  const bool pooled = stage < 3 && layer == 2;
  RN_TRY(split_rows_launch(weight, cout, cin, cin, L.kpad, L.w_hi, L.w_lo, L.w_f32, s, rot, pooled ? scale : nullptr));
  if (pooled) RN_TRY(abs_copy_launch(scale, L.scale, cout, s));
  else RN_CUDA(cudaMemcpyAsync(L.scale, scale, sizeof(float) * cout, cudaMemcpyDeviceToDevice, s));
  RN_CUDA(cudaMemcpyAsync(L.shift, shift, sizeof(float) * cout, cudaMemcpyDeviceToDevice, s));
  L.set = true;
  return REGNET_OK;
}

// Geometry chain of one forward (depends on xyz only): FPS -> ball query per level, then the three 3-NN searches.
// Runs on the side stream when there is one (after everything already queued on `ms`, which orders it behind the
// previous reader of this slot), else on `ms`.
// Level tables of one forward: xyz pointer / strides / point count of levels 0..3 for geometry slot G.
struct Levels {
  const float* xyz[4];
  Strides3 st[4];
  int n[4];
};
static Levels make_levels(const regnet_scorenet* p, const regnet_scorenet::Geom& G, const float* pc) {
  const int N = p->N;
  const int* M = p->M;
  Levels L;
  L.xyz[0] = pc; L.xyz[1] = G.new_xyz[0]; L.xyz[2] = G.new_xyz[1]; L.xyz[3] = G.new_xyz[2];
  L.st[0] = Strides3{(int64_t)N * 6, 1, 6};  // pc (B,N,6): the (B,3,N) view of score_network.py:46 without a copy
  for (int i = 0; i < 3; ++i) L.st[i + 1] = Strides3{(int64_t)3 * M[i], M[i], 1};
  L.n[0] = N; L.n[1] = M[0]; L.n[2] = M[1]; L.n[3] = M[2];
  return L;
}

static int ball_query_level(regnet_scorenet* p, regnet_scorenet::Geom& G, const Levels& L, int i, cudaStream_t s) {
  const float r = p->cfg.radius[i];
  if (p->cfg.use_grid && i < 2 && L.n[i] <= 65536) {
    // levels 0 and 1: bin the points into an (x,y) grid of cell edge >= r, then look at 3x3 cells per centroid
    prof_begin(p, i == 0 ? "grid_build.0" : "grid_build.1", s);
    RN_TRY(grid_build_launch(L.xyz[i], L.st[i], p->B, L.n[i], r * 1.001f + 1e-7f, p->grid_ws[i], s));
    prof_end(p, s);
    prof_begin(p, BQ_LABEL[i], s);
    RN_TRY(ball_query_grid_launch(L.xyz[i], L.st[i], L.xyz[i + 1], L.st[i + 1], p->B, L.n[i], p->M[i], r, p->grid_ws[i],
                                  G.nbr[i], s));
    prof_end(p, s);
    p->launches += 2;
    return REGNET_OK;
  }
  prof_begin(p, BQ_LABEL[i], s);
  RN_TRY(ball_query_launch(L.xyz[i], L.st[i], L.xyz[i + 1], L.st[i + 1], p->B, L.n[i], p->M[i], r, 64, nullptr, nullptr,
                           G.nbr[i], s));
  prof_end(p, s);
  ++p->launches;
  return REGNET_OK;
}

static int three_nn_all(regnet_scorenet* p, regnet_scorenet::Geom& G, const Levels& L, cudaStream_t s) {
  for (int f = 0; f < 3; ++f) {  // fp f: dense level 2-f, sparse level 3-f
    const int dl = 2 - f, sl = 3 - f;
    if (p->cfg.use_grid && f == 2) {
      // the big one (N queries against M0 keys): finest grid the cell budget allows over the keys
      prof_begin(p, "grid_build.nn", s);
      RN_TRY(grid_build_launch(L.xyz[sl], L.st[sl], p->B, L.n[sl], 0.f, p->grid_ws[1], s));
      prof_end(p, s);
      prof_begin(p, NN_LABEL[f], s);
      RN_TRY(three_nn_grid_launch(L.xyz[dl], L.st[dl], L.xyz[sl], L.st[sl], p->B, L.n[dl], L.n[sl], p->grid_ws[1],
                                  G.nn_idx[f], G.nn_w[f], s));
      prof_end(p, s);
      p->launches += 2;
      continue;
    }
    prof_begin(p, NN_LABEL[f], s);
    RN_TRY(three_nn_launch(L.xyz[dl], L.st[dl], L.xyz[sl], L.st[sl], p->B, L.n[dl], L.n[sl], nullptr, nullptr,
                           G.nn_idx[f], G.nn_w[f], s));
    prof_end(p, s);
    ++p->launches;
  }
  return REGNET_OK;
}

// Side-stream part of one forward's geometry (depends on xyz only).
//   use_side_stream == 1: the whole chain (FPS -> ball query per level, then the 3-NN searches) runs on the side stream;
//   use_side_stream == 2: only the three FPS launches do (register-resident, ~1 KB of shared memory: they co-reside
//                         with the one-CTA-per-SM GEMM kernels); ball query / 3-NN need tens of KB of shared memory
//                         per CTA, cannot share an SM with a GEMM CTA and stay on the caller's stream;
//   no side stream / profiling: everything on `ms`.
// It is ordered behind everything already queued on `ms`, i.e. behind the previous reader of this slot.
static int geometry_enqueue(regnet_scorenet* p, const float* pc, int slot, cudaStream_t ms, bool overlapped) {
  regnet_scorenet::Geom& G = p->geom[slot];
  const bool fork = p->side != nullptr && p->profiling != 1;  // serial profiling puts everything on `ms`
  // mode 2: only FPS leaves the caller's stream.  mode 3 (needs side2): ball query of levels 1-2 and the 3-NN searches
  // follow their FPS on the second side stream (they have > 1 ms of slack before the MLP chain needs them and are
  // latency-bound kernels that fill the tails of the persistent GEMM launches); level 0's ball query stays on the
  // caller's stream, it is the first thing the MLP chain needs.
  const bool mode3 = fork && p->cfg.use_side_stream == 3 && p->side2 != nullptr;
  const bool fps_only = fork && (p->cfg.use_side_stream == 2 || (p->cfg.use_side_stream == 3 && !mode3));
  cudaStream_t gs = fork ? p->side : ms;
  if (fork) {
    RN_CUDA(cudaEventRecord(p->ev_start, ms));
    RN_CUDA(cudaStreamWaitEvent(gs, p->ev_start, 0));
  }
  const Levels L = make_levels(p, G, pc);
  for (int i = 0; i < 3; ++i) {
    // FPS-only modes: levels 1 and 2 go to a second side stream behind level 0's event.  The side stream then carries
    // nothing but the level-0 launches of consecutive forwards, so the pipelined period is max(fps.0, MLP chain)
    // instead of fps.0 + fps.1 + fps.2 (timeline: 6.0 + 1.0 + 0.2 ms co-running against a 7.0 ms MLP chain).
    cudaStream_t gs_i = ((fps_only || mode3) && i > 0 && p->side2) ? p->side2 : gs;
    if (gs_i != gs && i == 1) RN_CUDA(cudaStreamWaitEvent(gs_i, G.ev_bq[0], 0));
    prof_begin(p, FPS_LABEL[i], gs_i);
    // a prefetched FPS shares its SMs with the previous step's GEMM CTAs: 4 warps (one per scheduler partition,
    // 227 registers) is the shape whose register-file footprint leaves room for them (profiles/README.md)
    // The smaller levels get 4-warp CTAs as well when prefetched: a 256- or 512-thread CTA cannot become resident next
    // to a tensor kernel's CTA and waits for a whole GEMM launch to drain (timeline: fps.2 0.10 ms alone, 1.15 ms
    // co-running) -- and the FPS chain is what bounds the pipelined step.
    const bool corun = overlapped && fork && (L.n[i] > 12288 || p->cfg.corun_small);
    int cs = 0, th = 0;
    if (corun) {
      cs = L.n[i] > 12288 ? p->cfg.corun_cs : L.n[i] > 2048 ? p->cfg.corun1_cs : 4;
      th = L.n[i] > 12288 ? p->cfg.corun_threads : L.n[i] > 2048 ? p->cfg.corun1_threads : 128;
    }
    // corun_single = 1: a co-running FPS takes the one-pick-per-exchange kernel (1.3 KB of shared memory: fits next to
    // sa0_chain, so it may start with the step).  Default 0: multi-pick rounds everywhere -- since the replay loop was
    // tightened they are 2.3x faster alone (1.53 vs 3.56 ms) and, parked behind sa0_chain (defer_prefetch), also give the
    // shorter pipelined step (profiles/README.md)
    RN_TRY(fps_launch(L.xyz[i], L.st[i], p->B, L.n[i], p->M[i], nullptr, G.fps_idx[i], G.new_xyz[i], cs, th, gs_i,
                      corun && p->cfg.corun_single));
    prof_end(p, gs_i);
    ++p->launches;
    if (!fps_only && !(mode3 && i == 0)) RN_TRY(ball_query_level(p, G, L, i, gs_i));
    if (fork) RN_CUDA(cudaEventRecord(G.ev_bq[i], gs_i));  // level i ready (FPS only, or FPS + ball query)
  }
  cudaStream_t ns = mode3 ? p->side2 : gs;
  if (!fps_only) RN_TRY(three_nn_all(p, G, L, ns));
  if (fork) RN_CUDA(cudaEventRecord(G.ev_nn, ns));
  return REGNET_OK;
}

// enqueue a parked prefetch now, ordered behind everything already queued on `ms`
static int flush_deferred(regnet_scorenet* p, cudaStream_t ms, bool overlapped) {
  if (p->deferred_slot < 0) return REGNET_OK;
  const int slot = p->deferred_slot;
  const float* pc = p->deferred_pc;
  p->deferred_slot = -1;
  p->deferred_pc = nullptr;
  const int saved = p->launches;
  p->launches = 0;
  RN_TRY(geometry_enqueue(p, pc, slot, ms, overlapped));
  p->prefetch_launches = p->launches;
  p->launches = saved;
  return REGNET_OK;
}

int regnet_scorenet_prefetch(regnet_scorenet* p, const float* pc, void* stream_) {
  RN_CHECK_ARG(p && pc, "scorenet_prefetch: null argument");
  const int slot = p->next_slot;
  if (p->geom[slot].pending) {
    set_error("scorenet_prefetch: both geometry slots hold unconsumed prefetches; call regnet_scorenet_forward first");
    return REGNET_EINVAL;
  }
  RN_TRY(flush_deferred(p, (cudaStream_t)stream_, true));      // at most one prefetch is parked
  p->geom[slot].pc = pc;
  p->geom[slot].pending = true;
  p->next_slot ^= 1;
  if (p->defer_prefetch && p->side != nullptr && p->profiling != 1) {
    p->deferred_pc = pc;
    p->deferred_slot = slot;
    p->prefetch_launches = 0;
    return REGNET_OK;
  }
  const int saved = p->launches;
  p->launches = 0;
  RN_TRY(geometry_enqueue(p, pc, slot, (cudaStream_t)stream_, true));
  p->prefetch_launches = p->launches;
  p->launches = saved;
  return REGNET_OK;
}

int regnet_scorenet_join_prefetch(regnet_scorenet* p, void* stream_) {
  RN_CHECK_ARG(p, "scorenet_join_prefetch: null plan");
  RN_TRY(flush_deferred(p, (cudaStream_t)stream_, true));
  if (!p->side) return REGNET_OK;   // single-stream plans: prefetches are already in `stream` order
  for (int g = 0; g < 2; ++g) {
    if (!p->geom[g].pending) continue;
    for (int i = 0; i < 3; ++i) RN_CUDA(cudaStreamWaitEvent((cudaStream_t)stream_, p->geom[g].ev_bq[i], 0));
    RN_CUDA(cudaStreamWaitEvent((cudaStream_t)stream_, p->geom[g].ev_nn, 0));
  }
  return REGNET_OK;
}

int regnet_scorenet_forward(regnet_scorenet* p, const float* pc, float* all_feature, float* score, void* stream_) {
  cudaStream_t ms = (cudaStream_t)stream_;
  RN_CHECK_ARG(p && pc && all_feature && score, "scorenet_forward: null argument");
  const int B = p->B, N = p->N;
  const int* M = p->M;
  const bool fork = p->side != nullptr && p->profiling != 1;
  if (p->profiling != 2) p->nrec = 0;  // timeline mode accumulates across forwards until read
  p->launches = 0;
  p->last_allfeat = all_feature;
  // geometry: consume the oldest prefetch if it was made for this input, otherwise compute it now
  if (p->deferred_slot >= 0 && (p->deferred_pc == pc || p->profiling == 1)) RN_TRY(flush_deferred(p, ms, false));
  int slot = p->next_slot;
  const int oldest = p->geom[p->next_slot].pending ? p->next_slot : (p->next_slot ^ 1);
  if (p->geom[oldest].pending && p->geom[oldest].pc == pc && p->profiling != 1) {
    slot = oldest;
    p->launches = p->prefetch_launches;
  } else {
    if (p->geom[slot].pending) slot ^= 1;
    if (p->geom[slot].pending) {
      set_error("scorenet_forward: input does not match the pending prefetches");
      return REGNET_EINVAL;
    }
    RN_TRY(geometry_enqueue(p, pc, slot, ms, false));
    if (slot == p->next_slot) p->next_slot ^= 1;
  }
  regnet_scorenet::Geom& G = p->geom[slot];
  G.pending = false;
  p->last_slot = slot;
  p->gemm_idx = 0;
  // a prefetch for the next batch is outstanding: its FPS CTAs will share the SMs with this forward's kernels
  const bool corunning = p->geom[slot ^ 1].pending || p->deferred_slot >= 0;
  if (p->cfg.engine == REGNET_ENGINE_TC && p->cfg.dynamic_tiles)
    RN_CUDA(cudaMemsetAsync(p->tile_counters, 0, sizeof(unsigned int) * 64, ms));
  const bool mode3 = fork && p->cfg.use_side_stream == 3 && p->side2 != nullptr;
  const bool fps_only = fork && (p->cfg.use_side_stream == 2 || (p->cfg.use_side_stream == 3 && !mode3));
  const Levels L = make_levels(p, G, pc);
  const float* const* lvl_xyz = L.xyz;
  const Strides3* lvl_st = L.st;
  const int* lvl_n = L.n;

  // ---- set abstraction MLPs -----------------------------------------------------------------------------
  const float* feat = pc + 3;      // level-0 features = rgb, rows of stride 6
  int64_t feat_bs = (int64_t)N * 6;
  int feat_ld = 6, feat_c = 3;
  for (int i = 0; i < 3; ++i) {
    if (fork) RN_CUDA(cudaStreamWaitEvent(ms, G.ev_bq[i], 0));
    if (fps_only || (mode3 && i == 0)) RN_TRY(ball_query_level(p, G, L, i, ms));
    const int64_t P = (int64_t)B * M[i] * 64;
    const int kpad = round_up(feat_c + 3, 16);
    Act a1 = make_act(p, 1, P, SA_CH[i][0]);
    Act a2 = make_act(p, 0, P, SA_CH[i][1]);
    bool need_l1 = true;
    if (i == 0 && p->cfg.fuse_sa0 >= 3 && p->cfg.engine == REGNET_ENGINE_TC) {
      // level 0: the whole chain (gather, 6 -> 128 -> 128 -> 256, max-pool) in one kernel, activations chained through TMEM
      const Layer& L0 = p->layers[0][0];
      const Layer& L1 = p->layers[0][1];
      const Layer& L2 = p->layers[0][2];
      if (!L0.set || !L1.set || !L2.set) {
        set_error("scorenet: sa_modules.0.mlp.{0,1,2} were never given weights");
        return REGNET_EINVAL;
      }
      unsigned int* counter = (p->cfg.dynamic_tiles && p->gemm_idx < 64) ? p->tile_counters + (p->gemm_idx++) : nullptr;
      prof_begin(p, "sa0_chain", ms);
      RN_TRY(sa0_chain_launch(lvl_xyz[0], lvl_st[0], G.new_xyz[0], feat, feat_bs, feat_ld, G.nbr[0], L0.w_f32, L0.kpad,
                              L0.scale, L0.shift, L1.w_hi, L1.w_lo, L1.kpad, L1.scale, L1.shift, L2.w_hi, L2.w_lo, L2.kpad,
                              L2.scale, L2.shift, B, M[0], p->sa_out[0], SA_CH[0][2], p->sa_hi[0], p->sa_lo[0], nullptr,
                              counter, p->cfg.sa0_variant, ms));
      prof_end(p, ms);
      ++p->launches;
      if (p->defer_prefetch <= 1) {   // the next step's geometry chain starts here, behind the level-0 kernel (see deferred_pc)
        const int keep = p->launches;
        RN_TRY(flush_deferred(p, ms, true));
        p->launches = keep;
      }
      feat = p->sa_out[i];
      feat_c = feat_ld = SA_CH[i][2];
      feat_bs = (int64_t)M[i] * feat_c;
      continue;
    }
    if (i == 0 && p->cfg.fuse_sa0 >= 2 && p->cfg.engine == REGNET_ENGINE_TC) {
      // level 0: gather + centre + layers 0 and 1 (6 -> 128 -> 128) in one kernel; the first activation stays on chip
      const Layer& L0 = p->layers[0][0];
      const Layer& L1 = p->layers[0][1];
      if (!L0.set || !L1.set) {
        set_error("scorenet: sa_modules.0.mlp.{0,1} were never given weights");
        return REGNET_EINVAL;
      }
      unsigned int* counter = (p->cfg.dynamic_tiles && p->gemm_idx < 64) ? p->tile_counters + (p->gemm_idx++) : nullptr;
      prof_begin(p, "sa0_front", ms);
      RN_TRY(sa0_front_launch(lvl_xyz[0], lvl_st[0], G.new_xyz[0], feat, feat_bs, feat_ld, G.nbr[0], L0.w_f32, L0.kpad,
                              L0.scale, L0.shift, L1.w_hi, L1.w_lo, L1.kpad, L1.scale, L1.shift, B, M[0], a2.hi, a2.lo,
                              a2.ld, counter, ms));
      prof_end(p, ms);
      ++p->launches;
      need_l1 = false;
    } else if (i == 0 && p->cfg.fuse_sa0) {
      // level 0: gather + centre + first layer (6 -> 128) in one SIMT pass, no 6-channel operand round trip
      const Layer& L0 = p->layers[0][0];
      if (!L0.set) {
        set_error("scorenet: sa_modules.0.mlp.0 was never given weights");
        return REGNET_EINVAL;
      }
      prof_begin(p, "sa0_fused", ms);
      RN_TRY(sa0_fused_launch(lvl_xyz[0], lvl_st[0], G.new_xyz[0], feat, feat_bs, feat_ld, G.nbr[0], L0.w_f32, L0.kpad,
                              L0.scale, L0.shift, L0.cout, B, M[0], 64, a1.f32, a1.hi, a1.lo, a1.ld, ms));
      prof_end(p, ms);
      ++p->launches;
    } else if (i > 0 && p->cfg.sa_linear_first && p->cfg.engine == REGNET_ENGINE_TC && feat_c % 64 == 0 && p->sa_hi[i - 1]) {
      // levels 1, 2: the feature part of the first convolution is applied per POINT of the previous level (Z, a GEMM over
      // N_prev rows instead of M*64 grouped rows), the grouping then gathers rows of Z and adds the fp32 xyz term
      const Layer& L0 = p->layers[i][0];
      if (!L0.set) {
        set_error("scorenet: sa_modules.%d.mlp.0 was never given weights", i);
        return REGNET_EINVAL;
      }
      Epilogue ez;
      ez.act = 0; ez.out_f32 = p->sa_z; ez.ld_f32 = L0.cout;
      if (p->cfg.dynamic_tiles && p->gemm_idx < 64) ez.tile_counter = p->tile_counters + (p->gemm_idx++);
      prof_begin(p, GEMM_LABEL[i][0], ms);
      RN_TRY(gemm_tc_launch(p->sa_hi[i - 1], p->sa_lo[i - 1], feat_c, L0.w_hi, L0.w_lo, L0.kpad, (int64_t)B * lvl_n[i], feat_c,
                            L0.cout, ez, ms));
      prof_end(p, ms);
      ++p->launches;
      const Layer& L1 = p->layers[i][1];
      if (p->cfg.sa_fused_a && p->cfg.dynamic_tiles && p->gemm_idx < 64 && L1.set && L1.cin == L0.cout &&
          gemm_fused_a_supported(P, L0.cout, M[i] * 64)) {
        // Fold the xyz term into the per-point table and a per-centroid vector (gemm_fused_a.cu): the grouped activation
        // is then relu(Z'[g] + T).  sa_fused_a = 1: the second layer's GEMM builds that operand tile in shared memory and
        // it is never materialised; 3: it is materialised by a gather-add pass and fed to the plain GEMM -- identical bits;
        // 2 (default): the fused GEMM unless a prefetched FPS is co-running.  Measured (B = 15, 25 600 points): the fused
        // kernel's eight producer warps take a single forward from 10.0 to 8.9 ms, but next to a co-resident FPS they
        // cost it more issue slots than the saved pass returns (pipelined step 7.1 -> 7.8 ms).
        const bool fused = p->cfg.sa_fused_a == 1 || (p->cfg.sa_fused_a == 2 && !corunning);
        prof_begin(p, i == 1 ? "sa_fold.1" : "sa_fold.2", ms);
        RN_TRY(sa_fold_launch(p->sa_z, L0.cout, lvl_n[i], lvl_xyz[i], lvl_st[i], G.new_xyz[i], M[i], L0.w_f32 + feat_c, L0.kpad,
                              L0.scale, L0.shift, B, L0.cout, p->sa_t, ms));
        prof_end(p, ms);
        ++p->launches;
        if (fused) {
          prof_begin(p, GEMM_LABEL[i][1], ms);
          RN_TRY(gemm_fused_a_launch(p->sa_z, L0.cout, p->sa_t, G.nbr[i], M[i] * 64, lvl_n[i], L1.w_hi, L1.w_lo, L1.kpad, P,
                                     L1.cin, L1.cout, L1.scale, L1.shift, a2.hi, a2.lo, a2.ld,
                                     p->tile_counters + (p->gemm_idx++), ms));
          prof_end(p, ms);
          ++p->launches;
          need_l1 = false;
        } else {
          prof_begin(p, SAOP_LABEL[i], ms);
          RN_TRY(sa_gather_add_launch(p->sa_z, L0.cout, lvl_n[i], p->sa_t, G.nbr[i], M[i] * 64, L0.cout, P, a1.hi, a1.lo, ms));
          prof_end(p, ms);
          ++p->launches;
        }
      } else {
        prof_begin(p, SAOP_LABEL[i], ms);
        RN_TRY(sa_gather_affine_launch(p->sa_z, L0.cout, lvl_n[i], lvl_xyz[i], lvl_st[i], G.new_xyz[i], L0.w_f32 + feat_c,
                                       L0.kpad, G.nbr[i], L0.scale, L0.shift, B, M[i], L0.cout, nullptr, a1.hi, a1.lo, ms));
        prof_end(p, ms);
        ++p->launches;
      }
    } else if (i > 0 && (p->cfg.gather_a >> (i - 1) & 1) && p->cfg.engine == REGNET_ENGINE_TC && feat_c % 64 == 0 &&
               p->sa_hi[i - 1]) {
      // the grouping is fused into the first layer's TMA producer (tile::gather4 from the previous level's pooled
      // features, kept as bf16 hi/lo planes); only the (P,16) xyz - centroid columns are materialised.  cfg.gather_a is a
      // bit mask over levels 1, 2; default = level 1 only: gather4 moves one 128-byte row per ~24 cycles per SM, which
      // beats the operand round trip of level 1 (1.04 -> 0.71 ms) but not of level 2, whose two 256-channel column tiles
      // gather every row twice (0.54 -> 0.64 ms).
      const Layer& L0 = p->layers[i][0];
      if (!L0.set) {
        set_error("scorenet: sa_modules.%d.mlp.0 was never given weights", i);
        return REGNET_EINVAL;
      }
      prof_begin(p, i == 1 ? "sa_xyzrel.1" : "sa_xyzrel.2", ms);
      RN_TRY(sa_operand_launch(lvl_xyz[i], lvl_st[i], G.new_xyz[i], nullptr, 0, 0, 0, G.nbr[i], B, lvl_n[i], M[i], 64, 16,
                               nullptr, p->xyzrel_hi, p->xyzrel_lo, ms));
      prof_end(p, ms);
      ++p->launches;
      Epilogue ep;
      ep.scale = L0.scale; ep.shift = L0.shift; ep.act = 1; ep.pool = 0;
      ep.out_hi = a1.hi; ep.out_lo = a1.lo; ep.ld_split = a1.ld;
      if (p->cfg.dynamic_tiles && p->gemm_idx < 64) ep.tile_counter = p->tile_counters + (p->gemm_idx++);
      prof_begin(p, GEMM_LABEL[i][0], ms);
      RN_TRY(gemm_tc_gather_launch(p->sa_hi[i - 1], p->sa_lo[i - 1], (int64_t)B * lvl_n[i], feat_c, feat_c, G.nbr[i],
                                   M[i] * 64, lvl_n[i], p->xyzrel_hi, p->xyzrel_lo, L0.w_hi, L0.w_lo, L0.kpad, P, L0.cout,
                                   ep, ms));
      prof_end(p, ms);
      ++p->launches;
    } else {
      Act a0 = make_act(p, 0, P, kpad);
      prof_begin(p, SAOP_LABEL[i], ms);
      RN_TRY(sa_operand_launch(lvl_xyz[i], lvl_st[i], G.new_xyz[i], feat, feat_bs, feat_ld, feat_c, G.nbr[i], B,
                               lvl_n[i], M[i], 64, kpad, a0.f32, a0.hi, a0.lo, ms));
      prof_end(p, ms);
      ++p->launches;
      RN_TRY(run_layer(p, GEMM_LABEL[i][0], p->layers[i][0], a0, P, 1, 0, &a1, nullptr, 0, ms));
    }
    if (need_l1) RN_TRY(run_layer(p, GEMM_LABEL[i][1], p->layers[i][1], a1, P, 1, 0, &a2, nullptr, 0, ms));
    p->pool_planes_level = p->cfg.engine == REGNET_ENGINE_TC ? i : -1;   // picked up by run_layer_impl
    RN_TRY(run_layer(p, GEMM_LABEL[i][2], p->layers[i][2], a2, P, 1, 64, nullptr, p->sa_out[i], SA_CH[i][2], ms));
    p->pool_planes_level = -1;
    if (i == (p->defer_prefetch <= 1 ? 0 : p->defer_prefetch - 1) || i == 2) {   // defer_prefetch = 2, 3: behind SA level 1, 2
      const int keep = p->launches;
      RN_TRY(flush_deferred(p, ms, true));
      p->launches = keep;
    }
    feat = p->sa_out[i];
    feat_c = feat_ld = SA_CH[i][2];
    feat_bs = (int64_t)M[i] * feat_c;
  }
  // ---- feature propagation --------------------------------------------------------------------------------
  if (fork) RN_CUDA(cudaStreamWaitEvent(ms, G.ev_nn, 0));
  if (fps_only) RN_TRY(three_nn_all(p, G, L, ms));
  const float* sparse = p->sa_out[2];
  int sparse_c = SA_CH[2][2];
  int sparse_n = M[2];
  Act seg_in;
  for (int f = 0; f < 3; ++f) {
    const int dl = 2 - f;
    const int nd = lvl_n[dl];
    const float* dense = dl == 0 ? pc + 3 : p->sa_out[dl - 1];
    const int dense_c = dl == 0 ? 3 : SA_CH[dl - 1][2];
    const int dense_ld = dl == 0 ? 6 : dense_c;
    const int64_t dense_bs = (int64_t)nd * dense_ld;
    const int64_t P = (int64_t)B * nd;
    const int kpad = round_up(sparse_c + dense_c, 16);
    Act cur = make_act(p, 0, P, kpad);
    int which = 1, first_layer = 0;
    const bool linear_first = p->cfg.fp_linear_first && p->cfg.engine == REGNET_ENGINE_TC;
    if (linear_first) {
      // The first convolution commutes with the (linear) 3-NN interpolation and with the concat: apply it at the Ns
      // sparse points (Y) and to the dense features (D), then interpolate Y, add D, BN + ReLU.  5x fewer GEMM rows for
      // fp2 (76.8 K instead of 384 K), and the (Nd, C2 + C1) operand is never built.
      const Layer& L0 = p->layers[3 + f][0];
      if (!L0.set) {
        set_error("scorenet: fp_modules.%d.mlp.0 was never given weights", f);
        return REGNET_EINVAL;
      }
      const int cout0 = L0.cout;
      const __nv_bfloat16* sp_hi = f == 0 ? p->sa_hi[2] : p->fp_hi[f - 1];
      const __nv_bfloat16* sp_lo = f == 0 ? p->sa_lo[2] : p->fp_lo[f - 1];
      Epilogue ey;
      ey.act = 0; ey.out_f32 = p->fp_y; ey.ld_f32 = cout0;
      if (p->cfg.dynamic_tiles && p->gemm_idx < 64) ey.tile_counter = p->tile_counters + (p->gemm_idx++);
      prof_begin(p, GEMM_LABEL[3 + f][0], ms);
      RN_TRY(gemm_tc_launch(sp_hi, sp_lo, sparse_c, L0.w_hi, L0.w_lo, L0.kpad, (int64_t)B * sparse_n, sparse_c, cout0, ey, ms));
      prof_end(p, ms);
      ++p->launches;
      if (dl > 0) {
        Epilogue ed;
        ed.act = 0; ed.out_f32 = p->fp_d; ed.ld_f32 = cout0;
        if (p->cfg.dynamic_tiles && p->gemm_idx < 64) ed.tile_counter = p->tile_counters + (p->gemm_idx++);
        prof_begin(p, f == 0 ? "gemm.fp0.l0d" : "gemm.fp1.l0d", ms);
        RN_TRY(gemm_tc_launch(p->sa_hi[dl - 1], p->sa_lo[dl - 1], dense_c, L0.w_hi + sparse_c, L0.w_lo + sparse_c, L0.kpad, P,
                              dense_c, cout0, ed, ms));
        prof_end(p, ms);
        ++p->launches;
      }
      Act nxt = make_act(p, which, P, cout0);
      prof_begin(p, FPOP_LABEL[f], ms);
      RN_TRY(fp_interp_affine_launch(p->fp_y, (int64_t)sparse_n * cout0, cout0, dl > 0 ? p->fp_d : nullptr, cout0,
                                     dl == 0 ? pc + 3 : nullptr, (int64_t)N * 6, 6, dl == 0 ? L0.w_f32 + sparse_c : nullptr,
                                     L0.kpad, G.nn_idx[f], G.nn_w[f], L0.scale, L0.shift, B, nd, cout0, nullptr, nxt.hi,
                                     nxt.lo, ms));
      prof_end(p, ms);
      ++p->launches;
      cur = nxt;
      which ^= 1;
      first_layer = 1;
    } else {
      prof_begin(p, FPOP_LABEL[f], ms);
      RN_TRY(fp_operand_launch(sparse, (int64_t)sparse_n * sparse_c, sparse_c, sparse_c, dense, dense_bs, dense_ld, dense_c,
                               G.nn_idx[f], G.nn_w[f], B, nd, kpad, cur.f32, cur.hi, cur.lo, ms));
      prof_end(p, ms);
      ++p->launches;
    }
    for (int l = first_layer; l < FP_NL[f]; ++l) {
      const bool last = l == FP_NL[f] - 1;
      const int cout = FP_CH[f][l];
      if (!last) {
        Act nxt = make_act(p, which, P, cout);
        RN_TRY(run_layer(p, GEMM_LABEL[3 + f][l], p->layers[3 + f][l], cur, P, 1, 0, &nxt, nullptr, 0, ms));
        cur = nxt;
        which ^= 1;
      } else if (f < 2) {
        Act planes;   // tcgen05 engine: also as planes, the sparse operand of the next module's Y GEMM
        planes.hi = p->fp_hi[f]; planes.lo = p->fp_lo[f]; planes.ld = cout;
        RN_TRY(run_layer(p, GEMM_LABEL[3 + f][l], p->layers[3 + f][l], cur, P, 1, 0, linear_first ? &planes : nullptr,
                         p->fp_out[f], cout, ms));
        sparse = p->fp_out[f];
      } else {
        // fp2's last layer: all_feature for the caller, and the seg head's operand
        if (p->cfg.engine == REGNET_ENGINE_SIMT) {
          RN_TRY(run_layer(p, GEMM_LABEL[3 + f][l], p->layers[3 + f][l], cur, P, 1, 0, nullptr, all_feature, cout, ms));
          seg_in.f32 = all_feature;
          seg_in.ld = cout;
        } else {
          seg_in = make_act(p, which, P, cout);
          RN_TRY(run_layer(p, GEMM_LABEL[3 + f][l], p->layers[3 + f][l], cur, P, 1, 0, &seg_in, all_feature, cout, ms));
          which ^= 1;
        }
      }
    }
    sparse_c = FP_CH[f][FP_NL[f] - 1];
    sparse_n = nd;
  }
  // ---- seg head + score -------------------------------------------------------------------------------------
  {
    const int64_t P = (int64_t)B * N;
    Act cur = seg_in;
    int which = (cur.hi == reinterpret_cast<__nv_bfloat16*>(p->arena[0]) || cur.f32 == reinterpret_cast<float*>(p->arena[0])) ? 1 : 0;
    for (int l = 0; l < 3; ++l) {
      Act nxt = make_act(p, which, P, SEG_CH[l]);
      RN_TRY(run_layer(p, GEMM_LABEL[6][l], p->layers[6][l], cur, P, 1, 0, &nxt, nullptr, 0, ms));
      cur = nxt;
      which ^= 1;
    }
    const Layer& H = p->layers[7][0];
    if (!H.set) {
      set_error("scorenet: score head weights were never set");
      return REGNET_EINVAL;
    }
    if (p->cfg.engine == REGNET_ENGINE_SIMT) {
      float* last = reinterpret_cast<float*>(p->arena[which]);  // (P,128) fp32
      Act o; o.f32 = last; o.ld = 128;
      RN_TRY(run_layer(p, GEMM_LABEL[6][3], p->layers[6][3], cur, P, 1, 0, &o, nullptr, 0, ms));
      prof_begin(p, "score_head", ms);
      RN_TRY(score_head_launch(last, 128, H.w_f32, H.scale, H.shift, P, 128, score, ms));
      prof_end(p, ms);
      ++p->launches;
    } else {
      // last seg layer (256 -> 128) with the score head (128 -> 1, BN, sigmoid) fused into its epilogue: the (P,128)
      // activation is consumed in registers and never written
      p->score_head = &H;
      p->score_out = score;
      RN_TRY(run_layer(p, GEMM_LABEL[6][3], p->layers[6][3], cur, P, 1, 0, nullptr, nullptr, 0, ms));
      p->score_head = nullptr;
    }
  }
  return REGNET_OK;
}

// Geometry only (FPS, ball query, 3-NN of all levels) for callers that run the MLPs themselves -- the training path
// (pointnet2.py _forward_modules): consumes a matching prefetch or computes the chain now, leaves `stream` ordered behind
// every result, and exposes them through regnet_scorenet_intermediate ("fps*", "xyz*", "bq*", "nn*", "nnw*").
int regnet_scorenet_geometry(regnet_scorenet* p, const float* pc, void* stream_) {
  cudaStream_t ms = (cudaStream_t)stream_;
  RN_CHECK_ARG(p && pc, "scorenet_geometry: null argument");
  const bool fork = p->side != nullptr && p->profiling != 1;
  p->launches = 0;
  RN_TRY(flush_deferred(p, ms, p->deferred_pc != pc));
  int slot = p->next_slot;
  const int oldest = p->geom[p->next_slot].pending ? p->next_slot : (p->next_slot ^ 1);
  if (p->geom[oldest].pending && p->geom[oldest].pc == pc && p->profiling != 1) {
    slot = oldest;
    p->launches = p->prefetch_launches;
  } else {
    if (p->geom[slot].pending) slot ^= 1;
    if (p->geom[slot].pending) {
      set_error("scorenet_geometry: input does not match the pending prefetches");
      return REGNET_EINVAL;
    }
    RN_TRY(geometry_enqueue(p, pc, slot, ms, false));
    if (slot == p->next_slot) p->next_slot ^= 1;
  }
  regnet_scorenet::Geom& G = p->geom[slot];
  G.pending = false;
  p->last_slot = slot;
  const bool mode3 = fork && p->cfg.use_side_stream == 3 && p->side2 != nullptr;
  const bool fps_only = fork && (p->cfg.use_side_stream == 2 || (p->cfg.use_side_stream == 3 && !mode3));
  const Levels L = make_levels(p, G, pc);
  for (int i = 0; i < 3; ++i) {
    if (fork) RN_CUDA(cudaStreamWaitEvent(ms, G.ev_bq[i], 0));
    if (fps_only || (mode3 && i == 0)) RN_TRY(ball_query_level(p, G, L, i, ms));
  }
  if (fork) RN_CUDA(cudaStreamWaitEvent(ms, G.ev_nn, 0));
  if (fps_only) RN_TRY(three_nn_all(p, G, L, ms));
  return REGNET_OK;
}

int regnet_scorenet_set_option(regnet_scorenet* p, const char* name, int value) {
  RN_CHECK_ARG(p && name, "scorenet_set_option: null argument");
  const std::string k(name);
  if (k == "defer_prefetch") p->defer_prefetch = value;
  else if (k == "dynamic_tiles") p->cfg.dynamic_tiles = value;
  else if (k == "sa_fused_a") p->cfg.sa_fused_a = value;
  else {
    set_error("scorenet_set_option: unknown option '%s'", name);
    return REGNET_EINVAL;
  }
  return REGNET_OK;
}

int regnet_scorenet_set_profiling(regnet_scorenet* p, int on) {
  RN_CHECK_ARG(p != nullptr, "scorenet_set_profiling: null plan");
  p->profiling = on;
  p->nrec = 0;
  return REGNET_OK;
}

int regnet_scorenet_profile(regnet_scorenet* p, char* buf, int64_t buf_bytes) {
  RN_CHECK_ARG(p && buf && buf_bytes > 0, "scorenet_profile: null argument");
  RN_CUDA(cudaDeviceSynchronize());
  int64_t off = 0;
  buf[0] = 0;
  for (size_t i = 0; i < p->nrec; ++i) {
    float ms = 0.f, t0 = 0.f;
    RN_CUDA(cudaEventElapsedTime(&ms, p->recs[i].a, p->recs[i].b));
    RN_CUDA(cudaEventElapsedTime(&t0, p->recs[0].a, p->recs[i].a));
    const int n = snprintf(buf + off, (size_t)(buf_bytes - off), "%s %.6f %.6f\n", p->recs[i].label, ms, t0);
    if (n < 0 || off + n >= buf_bytes) break;
    off += n;
  }
  return REGNET_OK;
}

int regnet_scorenet_intermediate(regnet_scorenet* p, const char* what, void** ptr, int64_t* numel) {
  RN_CHECK_ARG(p && what && ptr && numel, "scorenet_intermediate: null argument");
  const size_t n = strlen(what);
  RN_CHECK_ARG(n >= 3, "scorenet_intermediate: unknown name '%s'", what);
  const int i = what[n - 1] - '0';
  RN_CHECK_ARG(i >= 0 && i < 3, "scorenet_intermediate: unknown name '%s'", what);
  const std::string k(what, n - 1);
  const int B = p->B;
  const int nd[3] = {p->M[1], p->M[0], p->N};
  const regnet_scorenet::Geom& G = p->geom[p->last_slot];
  if (k == "fps") { *ptr = G.fps_idx[i]; *numel = (int64_t)B * p->M[i]; }
  else if (k == "xyz") { *ptr = G.new_xyz[i]; *numel = (int64_t)B * 3 * p->M[i]; }
  else if (k == "bq") { *ptr = G.nbr[i]; *numel = (int64_t)B * p->M[i] * 64; }
  else if (k == "nn") { *ptr = G.nn_idx[i]; *numel = (int64_t)B * nd[i] * 3; }
  else if (k == "nnw") { *ptr = G.nn_w[i]; *numel = (int64_t)B * nd[i] * 3; }
  else if (k == "sa") { *ptr = p->sa_out[i]; *numel = (int64_t)B * p->M[i] * SA_CH[i][2]; }
  else if (k == "fp" && i < 2) { *ptr = p->fp_out[i]; *numel = (int64_t)B * nd[i] * (i == 0 ? 1024 : 512); }
  else if (k == "fp" && i == 2) { *ptr = p->last_allfeat; *numel = (int64_t)B * p->N * 256; }
  else {
    set_error("scorenet_intermediate: unknown name '%s'", what);
    return REGNET_EINVAL;
  }
  return REGNET_OK;
}

}  // extern "C"
