// Region stage on device: grasp-centre selection, ball crops with random
// fixed-size sampling, grouped feature max-pool, and the per-row masked sampler used by the closing-box crop.
//
// Replaces host-bound Python of the reference:
//   R1 dataset_utils/get_regiondataset.py:354-434  _select_score_center  (score.cpu(), per-cloud FPS launches)
//   R2 dataset_utils/get_regiondataset.py:279-352  _get_local_points_batch + _get_group_pc
//      (materialises [N_C*N,3] repeats, then B*N_C Python iterations of nonzero + np.random.choice + index copy)
//   R3 multi_model/gripper_region_network.py:389-395 + utils/pointnet2.py:161,167 (row gather + MaxPool1d)
//   R6 multi_model/gripper_region_network.py:532-544 (per-grasp Python loop of nonzero + np.random.choice)
// Deterministic parts are exact (which points are positive / inside the ball / inside the box, FPS over the
// positives, ">= group_num" vs "> region_num" thresholds, the -1 fills).  The random parts use a counter-based
// generator instead of numpy's Mersenne twister: same distributions (uniform subsets without replacement via
// selection sampling, i.i.d. uniform draws with replacement), different streams -- the reference itself seeds from
// the wall clock (train.py:59, test.py:56), so its draws are not reproducible either.
#include <algorithm>
#include <stdlib.h>

#include "internal.cuh"

namespace regnet {

int fps_launch_var(const float* pts, Strides3 st, int B, int Nmax, int M, const int32_t* n_per_cloud, int32_t* idx32,
                   cudaStream_t stream);

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ uint32_t rand_u32(uint64_t seed, uint64_t a, uint64_t b) {
  return (uint32_t)(mix64(mix64(seed ^ (a * 0x2545f4914f6cdd1dull)) + b) >> 32);
}
__device__ __forceinline__ float rand_u01(uint64_t seed, uint64_t a, uint64_t b) {
  return (float)(rand_u32(seed, a, b) >> 8) * (1.0f / 16777216.0f);  // [0,1), 24 bits
}

// ---- R1a: ordered compaction of the positive points of each cloud ----------------------------------------------------
// map_index[b, 0..P) = ascending indices with score > thre (torch.nonzero order); cxyz = their xyz, planar (B,3,N).
__global__ void __launch_bounds__(1024)
compact_positives_kernel(const float* __restrict__ score, float thre, const float* __restrict__ pc, int N,
                         int32_t* __restrict__ map_index, float* __restrict__ cxyz, int32_t* __restrict__ count) {
  __shared__ int warp_sum[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ipt = (N + 1023) / 1024;
  const int lo = min(N, tid * ipt), hi = min(N, lo + ipt);
  const float* __restrict__ s = score + (int64_t)b * N;
  int c = 0;
  for (int j = lo; j < hi; ++j) c += s[j] > thre;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, w, o);
      if (lane >= o) w += v;
    }
    warp_sum[lane] = w;
  }
  __syncthreads();
  int pos = incl - c + (warp > 0 ? warp_sum[warp - 1] : 0);
  for (int j = lo; j < hi; ++j) {
    if (s[j] > thre) {
      map_index[(int64_t)b * N + pos] = j;
      const float* __restrict__ q = pc + ((int64_t)b * N + j) * 6;
      cxyz[((int64_t)b * 3 + 0) * N + pos] = q[0];
      cxyz[((int64_t)b * 3 + 1) * N + pos] = q[1];
      cxyz[((int64_t)b * 3 + 2) * N + pos] = q[2];
      ++pos;
    }
  }
  if (tid == 1023) count[b] = warp_sum[31];
}

// 4-round Feistel permutation of [0, 2^bits), cycle-walked onto [0, n): a seeded sample without replacement
__device__ __forceinline__ uint32_t feistel_perm(uint32_t x, int bits, uint32_t n, uint64_t seed) {
  const int hb = bits >> 1, lb = bits - hb;  // low half has lb bits
  const uint32_t lmask = (1u << lb) - 1u, hmask = (1u << hb) - 1u;
  do {
    uint32_t l = x & lmask, h = x >> lb;
#pragma unroll
    for (int r = 0; r < 4; ++r) {  // alternate which half is mixed so both widths are respected
      if ((r & 1) == 0) h = (h ^ rand_u32(seed, r, l)) & hmask;
      else l = (l ^ rand_u32(seed, r, h)) & lmask;
    }
    x = (h << lb) | l;
  } while (x >= n);
  return x;
}

// ---- R1c: centre indices + centre rows ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
finalize_centers_kernel(const float* __restrict__ pc, const int32_t* __restrict__ map_index,
                        const int32_t* __restrict__ count, const int32_t* __restrict__ fps_idx, int N, int M,
                        uint64_t seed, int64_t* __restrict__ center_index, float* __restrict__ center_pc) {
  const int b = blockIdx.x;
  const int P = count[b];
  int bits = 1;
  while ((1u << bits) < (uint32_t)N) ++bits;
  if (bits < 2) bits = 2;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    int src;
    if (P > M) {                      // get_regiondataset.py:379-383 / 409-415: FPS over the positives
      src = map_index[(int64_t)b * N + fps_idx[(int64_t)b * M + i]];
    } else if (P > 0) {               // :385-389 / 417-424: all positives, then uniform repeats
      const int k = i < P ? i : (int)(rand_u32(seed, (uint64_t)b, (uint64_t)i) % (uint32_t)P);
      src = map_index[(int64_t)b * N + k];
    } else {                          // :391-393 / 426-428: no positive at all -> M distinct random points
      src = (int)feistel_perm((uint32_t)i, bits, (uint32_t)N, seed + 0x51ed27ull * (uint64_t)(b + 1));
    }
    center_index[(int64_t)b * M + i] = src;
    const float* __restrict__ q = pc + ((int64_t)b * N + src) * 6;
    float* o = center_pc + ((int64_t)b * M + i) * 6;
#pragma unroll
    for (int c = 0; c < 6; ++c) o[c] = q[c];
  }
}

// ---- R2: ball crop + fixed-size random sample ---------------------------------------------------------------------------
// Inside test exactly as get_regiondataset.py:289-294: d = sqrt((dx*dx + dy*dy) + dz*dz) with separately rounded
// products (torch.mul / + / torch.sqrt), kept iff d <= r (non-strict), r rounded to fp32.
__device__ __forceinline__ bool in_ball(float4 q, float cx, float cy, float cz, float r) {
  const float dx = __fsub_rn(q.x, cx), dy = __fsub_rn(q.y, cy), dz = __fsub_rn(q.z, cz);
  const float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  return __fsqrt_rn(s) <= r;
}

constexpr int CROP_WARPS = 8;
constexpr int CROP_TILE = 2048;

__global__ void __launch_bounds__(CROP_WARPS * 32)
ball_crop_sample_kernel(const float* __restrict__ pc, const float* __restrict__ center_pc, int N, int NC, float radius,
                        int G, uint64_t seed, int64_t* __restrict__ index, float* __restrict__ group,
                        int32_t* __restrict__ count_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tile = reinterpret_cast<float4*>(smem_raw);
  int* lists = reinterpret_cast<int*>(smem_raw + sizeof(float4) * CROP_TILE);  // [CROP_WARPS][G]
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * CROP_WARPS + warp;
  const bool live = c < NC;
  const float* __restrict__ P = pc + (int64_t)b * N * 6;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (live) {
    const float* q = center_pc + ((int64_t)b * NC + c) * 6;
    cx = q[0]; cy = q[1]; cz = q[2];
  }
  auto stage = [&](int base, int n) {
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
      const float* q = P + (int64_t)(base + t) * 6;
      tile[t] = make_float4(q[0], q[1], q[2], 0.f);
    }
    __syncthreads();
  };
  // pass 1: how many points fall inside the ball
  int cnt = 0;
  for (int base = 0; base < N; base += CROP_TILE) {
    const int n = min(CROP_TILE, N - base);
    stage(base, n);
    if (live)
      for (int t = lane; t < n; t += 32) cnt += in_ball(tile[t], cx, cy, cz, radius);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
  // pass 2: selection sampling (no replacement) when there are enough points, else collect the hit list
  const uint64_t cid = (uint64_t)b * NC + (live ? c : 0);
  int64_t* out_i = index + cid * G;
  float* out_g = group + cid * (int64_t)G * 6;
  int* list = lists + warp * G;
  int seen = 0, chosen = 0;
  for (int base = 0; base < N; base += CROP_TILE) {
    const int n = min(CROP_TILE, N - base);
    stage(base, n);
    if (!live || cnt == 0) continue;
    for (int t0 = 0; t0 < n; t0 += 32) {
      const int t = t0 + lane;
      const bool hit = t < n && in_ball(tile[t], cx, cy, cz, radius);
      unsigned m = __ballot_sync(FULL, hit);
      if (cnt >= G) {                                   // :334-335  np.random.choice(len, G, replace=False)
        while (m && chosen < G) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          const int j = base + t0 + bit;
          const float u = rand_u01(seed, cid, (uint64_t)seen);
          if (u * (float)(cnt - seen) < (float)(G - chosen)) {
            if (lane == 0) out_i[chosen] = j;
            if (lane < 6) out_g[(int64_t)chosen * 6 + lane] = P[(int64_t)j * 6 + lane];
            ++chosen;
          }
          ++seen;
        }
      } else if (hit) {
        list[seen + __popc(m & ((1u << lane) - 1u))] = base + t;
      }
      if (cnt < G) seen += __popc(m);
    }
  }
  if (!live) return;
  __syncwarp();
  if (cnt == 0) {                                       // row keeps the reference's -1 fill (:324-325)
    for (int k = lane; k < G; k += 32) out_i[k] = -1;
    for (int k = lane; k < G * 6; k += 32) out_g[k] = -1.0f;
  } else if (cnt < G) {                                 // :336-337  np.random.choice(len, G, replace=True)
    for (int k = lane; k < G; k += 32) {
      int h = (int)(rand_u01(seed, cid, (uint64_t)k) * (float)cnt);
      h = min(h, cnt - 1);
      const int j = list[h];
      out_i[k] = j;
#pragma unroll
      for (int q = 0; q < 6; ++q) out_g[(int64_t)k * 6 + q] = P[(int64_t)j * 6 + q];
    }
  }
  if (lane == 0 && count_out) count_out[cid] = cnt;
}

// Grid form of the ball crop: the candidates of a centre are the three cell-row runs of its 3x3 cell block in the
// uniform (x, y) grid of the cloud (cell edge >= radius, grid.cu) instead of all N points -- ~70 candidates for the
// 8 mm crop and ~1 400 for the 64 mm crop of a 25 600-point table-top cloud.  Same membership test, same counts, same
// uniform sampling (Algorithm S does not care in which order the members are visited); the grid is built with the
// STABLE scatter, so the visiting order -- and with it every draw -- is reproducible for a given seed; the picks of
// the no-replacement case come out in cell order instead of ascending index order.
__global__ void __launch_bounds__(CROP_WARPS * 32)
ball_crop_grid_kernel(const float* __restrict__ pc, const float* __restrict__ center_pc, int N, int NC, float radius,
                      int G, uint64_t seed, const GridHeader* __restrict__ hdr, const int* __restrict__ cell_start,
                      const float4* __restrict__ sorted, int64_t* __restrict__ index, float* __restrict__ group,
                      int32_t* __restrict__ count_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* lists = reinterpret_cast<int*>(smem_raw);  // [CROP_WARPS][G]
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * CROP_WARPS + warp;
  if (c >= NC) return;
  const float* __restrict__ P = pc + (int64_t)b * N * 6;
  const float* q = center_pc + ((int64_t)b * NC + c) * 6;
  const float cx = q[0], cy = q[1], cz = q[2];
  const GridHeader g = hdr[b];
  const int* __restrict__ cs = cell_start + (int64_t)b * (MAX_CELLS + 1);
  const float4* __restrict__ sp = sorted + (int64_t)b * N;
  const int gx0 = cell_coord(cx, g.x0, g.inv_h, g.gx), gy0 = cell_coord(cy, g.y0, g.inv_h, g.gy);
  int beg[3], end[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int ry = gy0 + d - 1;
    if (ry < 0 || ry >= g.gy) { beg[d] = end[d] = 0; continue; }
    beg[d] = cs[ry * g.gx + max(0, gx0 - 1)];
    end[d] = cs[ry * g.gx + min(g.gx - 1, gx0 + 1) + 1];
  }
  // pass 1: count
  int cnt = 0;
#pragma unroll
  for (int d = 0; d < 3; ++d)
    for (int t = beg[d] + lane; t < end[d]; t += 32) cnt += in_ball(sp[t], cx, cy, cz, radius);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
  const uint64_t cid = (uint64_t)b * NC + c;
  int64_t* out_i = index + cid * G;
  float* out_g = group + cid * (int64_t)G * 6;
  int* list = lists + warp * G;
  int seen = 0, chosen = 0;
  if (cnt > 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      for (int t0 = beg[d]; t0 < end[d]; t0 += 32) {
        const int t = t0 + lane;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        bool hit = false;
        if (t < end[d]) { v = sp[t]; hit = in_ball(v, cx, cy, cz, radius); }
        const int j = __float_as_int(v.w);
        unsigned m = __ballot_sync(FULL, hit);
        if (cnt >= G) {                                   // np.random.choice(len, G, replace=False): Algorithm S
          while (m && chosen < G) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int jb = __shfl_sync(FULL, j, bit);
            const float u = rand_u01(seed, cid, (uint64_t)seen);
            if (u * (float)(cnt - seen) < (float)(G - chosen)) {
              if (lane == 0) out_i[chosen] = jb;
              if (lane < 6) out_g[(int64_t)chosen * 6 + lane] = P[(int64_t)jb * 6 + lane];
              ++chosen;
            }
            ++seen;
          }
        } else {
          if (hit) list[seen + __popc(m & ((1u << lane) - 1u))] = j;
          seen += __popc(m);
        }
      }
    }
  }
  __syncwarp();
  if (cnt == 0) {                                       // row keeps the reference's -1 fill (:324-325)
    for (int k = lane; k < G; k += 32) out_i[k] = -1;
    for (int k = lane; k < G * 6; k += 32) out_g[k] = -1.0f;
  } else if (cnt < G) {                                 // :336-337  np.random.choice(len, G, replace=True)
    for (int k = lane; k < G; k += 32) {
      int h = (int)(rand_u01(seed, cid, (uint64_t)k) * (float)cnt);
      h = min(h, cnt - 1);
      const int j = list[h];
      out_i[k] = j;
#pragma unroll
      for (int qq = 0; qq < 6; ++qq) out_g[(int64_t)k * 6 + qq] = P[(int64_t)j * 6 + qq];
    }
  }
  if (lane == 0 && count_out) count_out[cid] = cnt;
}

// ---- R6: closing-box membership -------------------------------------------------------------------------------------------
// mask[m, g] = point g of grasp m's crop lies in the gripper's closing box: p' = R_m (p - c_m), then the six strict
// half-space tests 0 < x' < depth/2, |y'| < width/2, |z'| < height/2 (gripper_region_network.py:505-528: a bmm over the
// whole (M, G, 3) crop followed by six comparison passes; here one pass, nothing but the mask is written).
__global__ void __launch_bounds__(256)
closing_box_mask_kernel(const float* __restrict__ pts, int C, int G, const float* __restrict__ centre,
                        const float* __restrict__ rot, const float* __restrict__ xlim_row,
                        const float* __restrict__ ylim_row, float xlim, float ylim, float zlim,
                        uint8_t* __restrict__ mask) {
  const int m = blockIdx.y;
  const float cx = centre[m * 3], cy = centre[m * 3 + 1], cz = centre[m * 3 + 2];
  float r[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) r[i] = rot[m * 9 + i];
  const float xl = xlim_row ? xlim_row[m] : xlim, yl = ylim_row ? ylim_row[m] : ylim;
  const float* __restrict__ row = pts + (int64_t)m * G * C;
  for (int g = blockIdx.x * 256 + threadIdx.x; g < G; g += gridDim.x * 256) {
    const float dx = __fsub_rn(row[(int64_t)g * C], cx), dy = __fsub_rn(row[(int64_t)g * C + 1], cy),
                dz = __fsub_rn(row[(int64_t)g * C + 2], cz);
    const float x = fmaf(r[2], dz, fmaf(r[1], dy, __fmul_rn(r[0], dx)));
    const float y = fmaf(r[5], dz, fmaf(r[4], dy, __fmul_rn(r[3], dx)));
    const float z = fmaf(r[8], dz, fmaf(r[7], dy, __fmul_rn(r[6], dx)));
    mask[(int64_t)m * G + g] = (x > 0.f) & (x < xl) & (y > -yl) & (y < yl) & (z > -zlim) & (z < zlim);
  }
}

// ---- R6: per-row masked sampler ------------------------------------------------------------------------------------
// rows x G mask -> K indices per row: more than K set -> K without replacement; more than `min_count` -> K with
// replacement; otherwise the row is rejected and keeps -1 (gripper_region_network.py:532-544: "> region_num",
// "> 5").  Output indices ascending in the no-replacement case.
__global__ void __launch_bounds__(256)
mask_sample_kernel(const uint8_t* __restrict__ mask, int rows, int G, int K, int min_count, uint64_t seed,
                   int64_t* __restrict__ idx_out, int32_t* __restrict__ count_out) {
  extern __shared__ int lists[];  // [8][K]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const uint8_t* __restrict__ m = mask + (int64_t)r * G;
  int cnt = 0;
  for (int t = lane; t < G; t += 32) cnt += m[t] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
  int64_t* out = idx_out + (int64_t)r * K;
  int* list = lists + warp * K;
  if (count_out && lane == 0) count_out[r] = cnt;
  if (cnt <= min_count) {
    for (int k = lane; k < K; k += 32) out[k] = -1;
    return;
  }
  int seen = 0, chosen = 0;
  for (int t0 = 0; t0 < G; t0 += 32) {
    const int t = t0 + lane;
    const bool hit = t < G && m[t] != 0;
    unsigned bits = __ballot_sync(FULL, hit);
    if (cnt > K) {
      while (bits && chosen < K) {
        const int bit = __ffs(bits) - 1;
        bits &= bits - 1;
        const float u = rand_u01(seed, (uint64_t)r, (uint64_t)seen);
        if (u * (float)(cnt - seen) < (float)(K - chosen)) {
          if (lane == 0) out[chosen] = t0 + bit;
          ++chosen;
        }
        ++seen;
      }
    } else {
      if (hit) list[seen + __popc(bits & ((1u << lane) - 1u))] = t;
      seen += __popc(bits);
    }
  }
  if (cnt <= K) {
    __syncwarp();
    for (int k = lane; k < K; k += 32) {
      int h = (int)(rand_u01(seed, (uint64_t)r, (uint64_t)k) * (float)cnt);
      out[k] = list[min(h, cnt - 1)];
    }
  }
}

// ---- R3: grouped feature max-pool -----------------------------------------------------------------------------------
// out[b,c,:] = max_g feat[row(b, idx[b,c,g]), :], feat point-major (B*N, C).  Negative indices wrap like the
// reference's `all_feature.view(-1,C)[idx + b*N]` does (python negative indexing of the flattened tensor).
__global__ void __launch_bounds__(128)
gather_max_kernel(const float* __restrict__ feat, const int64_t* __restrict__ idx, int B, int N, int NC, int G, int C,
                  float* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t cid = (int64_t)blockIdx.x * 4 + warp;
  if (cid >= (int64_t)B * NC) return;
  const int b = (int)(cid / NC);
  const int64_t* __restrict__ ix = idx + cid * G;
  const int64_t total = (int64_t)B * N;
  for (int c0 = lane * 4; c0 < C; c0 += 128) {
    float4 acc = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    int g = 0;
    for (; g + 4 <= G; g += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int64_t row = ix[g + u] + (int64_t)b * N;
        if (row < 0) row += total;
        v[u] = *reinterpret_cast<const float4*>(feat + row * C + c0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x = fmaxf(acc.x, v[u].x); acc.y = fmaxf(acc.y, v[u].y);
        acc.z = fmaxf(acc.z, v[u].z); acc.w = fmaxf(acc.w, v[u].w);
      }
    }
    for (; g < G; ++g) {
      int64_t row = ix[g] + (int64_t)b * N;
      if (row < 0) row += total;
      const float4 v = *reinterpret_cast<const float4*>(feat + row * C + c0);
      acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
    }
    *reinterpret_cast<float4*>(out + cid * C + c0) = acc;
  }
}

}  // namespace
}  // namespace regnet

using namespace regnet;

extern "C" {

int regnet_select_score_center(const float* pc, const float* score, int B, int N, int center_num, float score_thre,
                               uint64_t seed, int64_t* center_index, float* center_pc, int32_t* positive_count,
                               void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  RN_CHECK_ARG(pc && score && center_index && center_pc && workspace, "select_score_center: null argument");
  RN_CHECK_ARG(B > 0 && N > 0 && center_num > 0 && center_num <= N, "select_score_center: need 0 < center_num <= N");
  // workspace: map_index (B,N) i32 | cxyz (B,3,N) f32 | count (B) i32 | fps_idx (B,M) i32
  const int64_t need = (int64_t)B * N * 4 + (int64_t)B * 3 * N * 4 + (int64_t)B * 4 + (int64_t)B * center_num * 4 + 64;
  RN_CHECK_ARG(workspace_bytes >= need, "select_score_center: workspace too small (%lld < %lld bytes)",
               (long long)workspace_bytes, (long long)need);
  int32_t* map_index = (int32_t*)workspace;
  float* cxyz = (float*)(map_index + (int64_t)B * N);
  int32_t* count = (int32_t*)(cxyz + (int64_t)B * 3 * N);
  int32_t* fps_idx = count + ((B + 3) / 4) * 4;
  compact_positives_kernel<<<B, 1024, 0, s>>>(score, score_thre, pc, N, map_index, cxyz, count);
  RN_LAUNCH_CHECK("compact_positives_kernel");
  RN_TRY(fps_launch_var(cxyz, Strides3{(int64_t)3 * N, N, 1}, B, N, center_num, count, fps_idx, s));
  finalize_centers_kernel<<<B, 256, 0, s>>>(pc, map_index, count, fps_idx, N, center_num, seed, center_index, center_pc);
  RN_LAUNCH_CHECK("finalize_centers_kernel");
  if (positive_count) RN_CUDA(cudaMemcpyAsync(positive_count, count, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, s));
  return REGNET_OK;
}

int64_t regnet_select_score_center_workspace(int B, int N, int center_num) {
  return (int64_t)B * N * 4 + (int64_t)B * 3 * N * 4 + (int64_t)B * 4 + (int64_t)B * center_num * 4 + 64;
}

int regnet_ball_crop_sample(const float* pc, const float* center_pc, int B, int N, int NC, float radius, int group_num,
                            uint64_t seed, int64_t* index, float* group, int32_t* count, void* stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  RN_CHECK_ARG(pc && center_pc && index && group, "ball_crop_sample: null argument");
  RN_CHECK_ARG(B > 0 && N > 0 && NC > 0 && group_num > 0, "ball_crop_sample: empty problem");
  if (group_num > 4096) {
    set_error("ball_crop_sample: group_num=%d exceeds the supported maximum of 4096", group_num);
    return REGNET_ELIMIT;
  }
  const size_t smem = sizeof(float4) * CROP_TILE + sizeof(int) * (size_t)CROP_WARPS * group_num;
  RN_CUDA(cudaFuncSetAttribute(ball_crop_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(sizeof(float4) * CROP_TILE + sizeof(int) * CROP_WARPS * 4096)));
  dim3 grid(ceil_div(NC, CROP_WARPS), B);
  ball_crop_sample_kernel<<<grid, CROP_WARPS * 32, smem, s>>>(pc, center_pc, N, NC, radius, group_num, seed, index, group,
                                                              count);
  RN_LAUNCH_CHECK("ball_crop_sample_kernel");
  return REGNET_OK;
}

int64_t regnet_ball_crop_workspace_bytes(int B, int N) { return grid_workspace_bytes(B, N); }

int regnet_ball_crop_sample_ws(const float* pc, const float* center_pc, int B, int N, int NC, float radius, int group_num,
                               uint64_t seed, int64_t* index, float* group, int32_t* count, void* workspace,
                               int64_t workspace_bytes, void* stream_) {
  if (!workspace || workspace_bytes < grid_workspace_bytes(B, N) || N < 2048 || !(radius > 0.f) || getenv("REGNET_API_BRUTE"))
    return regnet_ball_crop_sample(pc, center_pc, B, N, NC, radius, group_num, seed, index, group, count, stream_);
  cudaStream_t s = (cudaStream_t)stream_;
  RN_CHECK_ARG(pc && center_pc && index && group, "ball_crop_sample: null argument");
  RN_CHECK_ARG(B > 0 && N > 0 && NC > 0 && group_num > 0, "ball_crop_sample: empty problem");
  if (group_num > 4096) {
    set_error("ball_crop_sample: group_num=%d exceeds the supported maximum of 4096", group_num);
    return REGNET_ELIMIT;
  }
  RN_TRY(grid_build_launch(pc, Strides3{(int64_t)N * 6, 1, 6}, B, N, radius * 1.001f + 1e-7f, workspace, s, true));
  const GridPtrs g = grid_carve(workspace, B, N);
  RN_CUDA(cudaFuncSetAttribute(ball_crop_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(sizeof(int) * CROP_WARPS * 4096)));
  dim3 grid(ceil_div(NC, CROP_WARPS), B);
  ball_crop_grid_kernel<<<grid, CROP_WARPS * 32, sizeof(int) * (size_t)CROP_WARPS * group_num, s>>>(
      pc, center_pc, N, NC, radius, group_num, seed, g.hdr, g.cell_start, g.sorted, index, group, count);
  RN_LAUNCH_CHECK("ball_crop_grid_kernel");
  return REGNET_OK;
}

int regnet_closing_box_mask(const float* points, int M, int G, int C, const float* centre, const float* rot,
                            const float* xlim_row, const float* ylim_row, float xlim, float ylim, float zlim,
                            uint8_t* mask, void* stream_) {
  RN_CHECK_ARG(points && centre && rot && mask, "closing_box_mask: null argument");
  RN_CHECK_ARG(M >= 0 && G > 0 && C >= 3 && M <= 65535 * 64, "closing_box_mask: bad sizes");
  if (M == 0) return REGNET_OK;
  RN_CHECK_ARG(M <= 65535, "closing_box_mask: more than 65535 grasps per call");
  closing_box_mask_kernel<<<dim3((unsigned)std::min(8, ceil_div(G, 256)), (unsigned)M), 256, 0, (cudaStream_t)stream_>>>(
      points, C, G, centre, rot, xlim_row, ylim_row, xlim, ylim, zlim, mask);
  RN_LAUNCH_CHECK("closing_box_mask_kernel");
  return REGNET_OK;
}

int regnet_mask_sample(const uint8_t* mask, int rows, int G, int K, int min_count, uint64_t seed, int64_t* index,
                       int32_t* count, void* stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  RN_CHECK_ARG(mask && index, "mask_sample: null argument");
  RN_CHECK_ARG(rows >= 0 && G > 0 && K > 0 && K <= 1024, "mask_sample: need G > 0 and 0 < K <= 1024");
  if (rows == 0) return REGNET_OK;
  mask_sample_kernel<<<ceil_div(rows, 8), 256, sizeof(int) * 8 * (size_t)K, s>>>(mask, rows, G, K, min_count, seed, index,
                                                                               count);
  RN_LAUNCH_CHECK("mask_sample_kernel");
  return REGNET_OK;
}

int regnet_gather_max(const float* feat, const int64_t* index, int B, int N, int NC, int G, int C, float* out,
                      void* stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  RN_CHECK_ARG(feat && index && out, "gather_max: null argument");
  RN_CHECK_ARG(B > 0 && N > 0 && NC > 0 && G > 0 && C > 0 && C % 4 == 0, "gather_max: need C %% 4 == 0 and non-empty sizes");
  const int64_t centers = (int64_t)B * NC;
  gather_max_kernel<<<(unsigned)((centers + 3) / 4), 128, 0, s>>>(feat, index, B, N, NC, G, C, out);
  RN_LAUNCH_CHECK("gather_max_kernel");
  return REGNET_OK;
}

}  // extern "C"
