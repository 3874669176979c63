// Uniform-grid acceleration of ball query and 3-NN (fused ScoreNet plan only).
//
// The reference scans every point for every centroid (csrc/ball_query_kernel.cu:57-72, csrc/interpolate_kernel.cu:51-70).
// The answers it defines do not depend on the scan order:
//   ball query = the K smallest ORIGINAL indices among the points with d < r^2 (first hit replicated, zeros if none);
//   3-NN       = the three smallest keys under the lexicographic order (d, index)   (strict '<' insertion while scanning
//                in ascending index order keeps the earlier index on equal distances).
// So the candidates may come from a spatial index, as long as every point within the radius is among them.  Table-top
// clouds are thin in z, hence a 2-D (x,y) grid: points are binned per cloud into cells of edge >= r (one CTA per cloud:
// bounding box, shared-memory histogram, scan, scatter); a centroid looks at its 3x3 cell block = three contiguous runs
// of the binned array.  Distances use the reference's exact rounding (sqdist_ref), so results stay bit-identical --
// tests/test_gpu_scorenet.py compares them with the brute-force oracle on every level.
#include "internal.cuh"

namespace regnet {

namespace {

constexpr unsigned FULL = 0xffffffffu;
// ---- build: bin the points of each cloud --------------------------------------------------------------------------
// sorted[b][pos] = {x, y, z, bits(original index)}; cell_start[b][c] .. cell_start[b][c+1] is cell c's run.
template <bool STABLE>
__global__ void __launch_bounds__(1024)
grid_build_kernel(const float* __restrict__ pts, Strides3 st, int N, float min_cell, GridHeader* __restrict__ hdr,
                  int* __restrict__ cell_start, float4* __restrict__ sorted) {
  __shared__ int hist[MAX_CELLS + 1];
  __shared__ float red[4][32];
  __shared__ int wsum[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ p = pts + (int64_t)b * st.b;
  // bounding box in x, y
  float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
  for (int j = tid; j < N; j += 1024) {
    const float x = p[(int64_t)j * st.n], y = p[(int64_t)j * st.n + st.c];
    xmin = fminf(xmin, x); xmax = fmaxf(xmax, x); ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xmin = fminf(xmin, __shfl_xor_sync(FULL, xmin, o)); xmax = fmaxf(xmax, __shfl_xor_sync(FULL, xmax, o));
    ymin = fminf(ymin, __shfl_xor_sync(FULL, ymin, o)); ymax = fmaxf(ymax, __shfl_xor_sync(FULL, ymax, o));
  }
  if (lane == 0) { red[0][warp] = xmin; red[1][warp] = xmax; red[2][warp] = ymin; red[3][warp] = ymax; }
  __syncthreads();
  xmin = red[0][lane]; xmax = red[1][lane]; ymin = red[2][lane]; ymax = red[3][lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xmin = fminf(xmin, __shfl_xor_sync(FULL, xmin, o)); xmax = fmaxf(xmax, __shfl_xor_sync(FULL, xmax, o));
    ymin = fminf(ymin, __shfl_xor_sync(FULL, ymin, o)); ymax = fmaxf(ymax, __shfl_xor_sync(FULL, ymax, o));
  }
  // cell edge: at least min_cell (>= search radius, with margin), enlarged until the grid fits GMAX x GMAX
  float h = fmaxf(min_cell, 1e-12f);
  h = fmaxf(h, fmaxf(xmax - xmin, ymax - ymin) * (1.0f / (GMAX - 1)));
  const float inv_h = 1.0f / h;
  const int gx = min(GMAX, (int)floorf((xmax - xmin) * inv_h) + 1);
  const int gy = min(GMAX, (int)floorf((ymax - ymin) * inv_h) + 1);
  const int cells = gx * gy;
  if (tid == 0) {
    GridHeader g;
    g.x0 = xmin; g.y0 = ymin; g.inv_h = inv_h; g.gx = gx; g.gy = gy; g.h = h; g.pad[0] = g.pad[1] = 0;
    hdr[b] = g;
  }
  for (int c = tid; c <= MAX_CELLS; c += 1024) hist[c] = 0;
  __syncthreads();
  for (int j = tid; j < N; j += 1024) {
    const float x = p[(int64_t)j * st.n], y = p[(int64_t)j * st.n + st.c];
    atomicAdd(&hist[cell_coord(y, ymin, inv_h, gy) * gx + cell_coord(x, xmin, inv_h, gx)], 1);
  }
  __syncthreads();
  // exclusive scan of hist[0..cells) (4 cells per thread), result in place; hist[cells] = N
  int v[4], s = 0;
#pragma unroll
  for (int u = 0; u < 4; ++u) { v[u] = hist[tid * 4 + u]; s += v[u]; }
  int incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, w, o);
      if (lane >= o) w += t;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  int run = incl - s + (warp > 0 ? wsum[warp - 1] : 0);
#pragma unroll
  for (int u = 0; u < 4; ++u) { hist[tid * 4 + u] = run; run += v[u]; }
  __syncthreads();
  int* __restrict__ cs = cell_start + (int64_t)b * (MAX_CELLS + 1);
  for (int c = tid; c <= cells; c += 1024) cs[c] = c < cells ? hist[c] : N;
  __syncthreads();
  float4* __restrict__ out = sorted + (int64_t)b * N;
  if (STABLE) {
    // Stable scatter: inside a cell the points keep ascending original index, so whatever enumerates a cell run sees a
    // reproducible order (the ball crop's random draws depend on it).  Chunks of 1024 consecutive points; within a
    // warp the same-cell lanes are ranked with match.any, and the 32 warps of a chunk reserve their slots in turn.
    for (int base = 0; base < N; base += 1024) {
      const int j = base + tid;
      int c = -1;
      float x = 0.f, y = 0.f, z = 0.f;
      if (j < N) {
        x = p[(int64_t)j * st.n]; y = p[(int64_t)j * st.n + st.c]; z = p[(int64_t)j * st.n + 2 * st.c];
        c = cell_coord(y, ymin, inv_h, gy) * gx + cell_coord(x, xmin, inv_h, gx);
      }
      const unsigned same = __match_any_sync(FULL, c);
      const int rank = __popc(same & ((1u << lane) - 1u)), leader = __ffs(same) - 1, group = __popc(same);
      for (int w = 0; w < 32; ++w) {
        if (warp == w) {
          int pos0 = 0;
          if (lane == leader && c >= 0) { pos0 = hist[c]; hist[c] = pos0 + group; }
          pos0 = __shfl_sync(FULL, pos0, leader);
          if (c >= 0) out[pos0 + rank] = make_float4(x, y, z, __int_as_float(j));
        }
        __syncthreads();
      }
    }
    return;
  }
  // scatter (order inside a cell is arbitrary; the queries re-establish index order themselves)
  for (int j = tid; j < N; j += 1024) {
    const float x = p[(int64_t)j * st.n], y = p[(int64_t)j * st.n + st.c], z = p[(int64_t)j * st.n + 2 * st.c];
    const int c = cell_coord(y, ymin, inv_h, gy) * gx + cell_coord(x, xmin, inv_h, gx);
    const int pos = atomicAdd(&hist[c], 1);
    out[pos] = make_float4(x, y, z, __int_as_float(j));
  }
}

// ---- ball query over the grid: one warp per centroid --------------------------------------------------------------------
constexpr int BQG_WARPS = 8;
constexpr int BQG_LIST = 1024;   // hits kept per centroid; more -> exact fallback scan in index order

__device__ __forceinline__ void bitonic64(uint32_t& a, uint32_t& b, int lane) {
  // ascending sort of 64 keys held as (a = element lane, b = element lane + 32)
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j == 32) {
        const bool up = (k == 64) ? true : false;
        const uint32_t lo = min(a, b), hi = max(a, b);
        a = up ? lo : hi;
        b = up ? hi : lo;
      } else {
        const uint32_t pa = __shfl_xor_sync(FULL, a, j), pb = __shfl_xor_sync(FULL, b, j);
        const bool lower = (lane & j) == 0;
        const bool up_a = ((lane & k) == 0) || k == 64;          // element index = lane      -> bit k of index
        const bool up_b = (((lane + 32) & k) == 0) || k == 64;   // element index = lane + 32
        a = (lower == up_a) ? min(a, pa) : max(a, pa);
        b = (lower == up_b) ? min(b, pb) : max(b, pb);
      }
    }
  }
}

__global__ void __launch_bounds__(BQG_WARPS * 32)
ball_query_grid_kernel(const float* __restrict__ pts, Strides3 pst, const float* __restrict__ ctr, Strides3 cst, int N,
                       int M, float radius, const GridHeader* __restrict__ hdr, const int* __restrict__ cell_start,
                       const float4* __restrict__ sorted, int32_t* __restrict__ index32, int64_t* __restrict__ index64,
                       int64_t* __restrict__ count64) {
  // outputs: index32 (the fused plan) and / or index64 + count64 (the pn2_ext operator surface), whichever is non-null
  constexpr int K = 64;
  __shared__ uint16_t lists[BQG_WARPS][BQG_LIST];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x * BQG_WARPS + warp;
  if (m >= M) return;
  const GridHeader g = hdr[b];
  const int* __restrict__ cs = cell_start + (int64_t)b * (MAX_CELLS + 1);
  const float4* __restrict__ sp = sorted + (int64_t)b * N;
  const float* __restrict__ c = ctr + (int64_t)b * cst.b;
  const float x1 = c[(int64_t)m * cst.n], y1 = c[(int64_t)m * cst.n + cst.c], z1 = c[(int64_t)m * cst.n + 2 * cst.c];
  const float r2 = __fmul_rn(radius, radius);
  const int cx = cell_coord(x1, g.x0, g.inv_h, g.gx), cy = cell_coord(y1, g.y0, g.inv_h, g.gy);
  uint16_t* list = lists[warp];
  int cnt = 0;
  bool overflow = false;
  for (int dy = -1; dy <= 1 && !overflow; ++dy) {
    const int ry = cy + dy;
    if (ry < 0 || ry >= g.gy) continue;
    const int xa = max(0, cx - 1), xb = min(g.gx - 1, cx + 1);
    const int beg = cs[ry * g.gx + xa], end = cs[ry * g.gx + xb + 1];
    for (int t0 = beg; t0 < end; t0 += 32) {
      const int t = t0 + lane;
      bool hit = false;
      int j = 0;
      if (t < end) {
        const float4 q = sp[t];
        hit = sqdist_ref(x1, y1, z1, q.x, q.y, q.z) < r2;
        j = __float_as_int(q.w);
      }
      const unsigned bits = __ballot_sync(FULL, hit);
      const int n = __popc(bits);
      if (cnt + n > BQG_LIST) { overflow = true; break; }
      if (hit) list[cnt + __popc(bits & ((1u << lane) - 1u))] = (uint16_t)j;
      cnt += n;
    }
  }
  __syncwarp();
  const int64_t obase = ((int64_t)b * M + m) * K;
  auto put = [&](int k, int v) {
    if (index32) index32[obase + k] = v;
    if (index64) index64[obase + k] = v;
  };
  if (overflow) {
    // more hits than the list holds (very dense ball): exact scan of the original order, first K hits
    const float* __restrict__ p = pts + (int64_t)b * pst.b;
    int got = 0, first = 0;
    for (int t0 = 0; t0 < N && got < K; t0 += 32) {
      const int j = t0 + lane;
      const bool hit = j < N && sqdist_ref(x1, y1, z1, p[(int64_t)j * pst.n], p[(int64_t)j * pst.n + pst.c],
                                           p[(int64_t)j * pst.n + 2 * pst.c]) < r2;
      const unsigned bits = __ballot_sync(FULL, hit);
      if (got == 0 && bits) first = t0 + __ffs(bits) - 1;
      const int slot = got + __popc(bits & ((1u << lane) - 1u));
      if (hit && slot < K) put(slot, j);
      got += __popc(bits);
    }
    got = min(got, K);
    for (int k = got + lane; k < K; k += 32) put(k, first);
    if (count64 && lane == 0) count64[(int64_t)b * M + m] = got;
    return;
  }
  // keep the K smallest original indices: find the K-th smallest by bisection on the index value
  uint32_t limit = 0xffffffffu;  // keep entries < limit
  if (cnt > K) {
    int lo = 0, hi = N;          // smallest T with count(idx < T) >= K
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      int below = 0;
      for (int i = lane; i < cnt; i += 32) below += list[i] < mid;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(FULL, below, o);
      if (below >= K) hi = mid; else lo = mid + 1;
    }
    limit = (uint32_t)lo;        // indices are distinct, so exactly K entries are < lo
  }
  // gather the (at most 64) survivors into two registers per lane, sort ascending, write
  uint32_t a = 0xffffffffu, bb = 0xffffffffu;
  int kept = 0;
  for (int i0 = 0; i0 < cnt; i0 += 32) {
    const int i = i0 + lane;
    const uint32_t v = i < cnt ? (uint32_t)list[i] : 0xffffffffu;
    const bool keep = v < limit;
    const unsigned bits = __ballot_sync(FULL, keep);
    const int slot = kept + __popc(bits & ((1u << lane) - 1u));
    // route value to (register, lane) = (slot / 32, slot % 32) through shared memory (the list is no longer needed
    // below index i0, and slot <= i always holds)
    __syncwarp();
    if (keep) list[slot] = (uint16_t)v;
    __syncwarp();
    kept += __popc(bits);
  }
  __syncwarp();
  const int n = min(kept, K);
  a = lane < n ? (uint32_t)list[lane] : 0xffffffffu;
  bb = lane + 32 < n ? (uint32_t)list[lane + 32] : 0xffffffffu;
  bitonic64(a, bb, lane);
  const uint32_t first = n > 0 ? __shfl_sync(FULL, a, 0) : 0u;
  put(lane, (int)(lane < n ? a : first));
  put(lane + 32, (int)(lane + 32 < n ? bb : first));
  if (count64 && lane == 0) count64[(int64_t)b * M + m] = n;
}

// ---- 3-NN over the grid of KEYS: one thread per query --------------------------------------------------------------
// Search the (2R+1)^2 cell block around the query for growing R until the third best distance is certified
// (d3 <= (R*h - overshoot)^2 means nothing outside the block can beat or tie it with a smaller index... ties outside
// would need d == d3 exactly, which the strict bound excludes), with a brute-force fallback after R = 3.
__device__ __forceinline__ void top3_insert(float d, int j, float& d0, float& d1, float& d2, int& i0, int& i1, int& i2) {
  // lexicographic (d, j): the order-independent statement of the reference's strict-'<' insertion scan
  const bool lt2 = d < d2 || (d == d2 && j < i2);
  if (!lt2) return;
  const bool lt1 = d < d1 || (d == d1 && j < i1);
  const bool lt0 = d < d0 || (d == d0 && j < i0);
  if (lt0) { d2 = d1; i2 = i1; d1 = d0; i1 = i0; d0 = d; i0 = j; }
  else if (lt1) { d2 = d1; i2 = i1; d1 = d; i1 = j; }
  else { d2 = d; i2 = j; }
}

__global__ void __launch_bounds__(128)
three_nn_grid_kernel(const float* __restrict__ qry, Strides3 qst, const float* __restrict__ key, Strides3 kst, int Nq,
                     int Nk, const GridHeader* __restrict__ hdr, const int* __restrict__ cell_start,
                     const float4* __restrict__ sorted, int32_t* __restrict__ index32, float* __restrict__ weight,
                     int64_t* __restrict__ index64, float* __restrict__ dist) {
  // outputs: index32 + normalised weights (the fused plan) and / or index64 + squared distances (operator surface)
  const int b = blockIdx.y;
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= Nq) return;
  const GridHeader g = hdr[b];
  const int* __restrict__ cs = cell_start + (int64_t)b * (MAX_CELLS + 1);
  const float4* __restrict__ sp = sorted + (int64_t)b * Nk;
  const float* __restrict__ q = qry + (int64_t)b * qst.b;
  const float x1 = q[(int64_t)i * qst.n], y1 = q[(int64_t)i * qst.n + qst.c], z1 = q[(int64_t)i * qst.n + 2 * qst.c];
  const float INF = __int_as_float(0x7f800000);
  float d0 = INF, d1 = INF, d2 = INF;
  int i0 = 0x7fffffff, i1 = 0x7fffffff, i2 = 0x7fffffff;
  // the query may lie outside the key grid: use unclamped cell coordinates for the certified radius
  const float fx = (x1 - g.x0) * g.inv_h, fy = (y1 - g.y0) * g.inv_h;
  const int cx = (int)floorf(fx), cy = (int)floorf(fy);
  bool done = false;
  int prev = -1;  // cells within Chebyshev radius `prev` of (cx,cy) are already scanned
  for (int R = 1; R <= 3 && !done; ++R) {
    for (int ry = cy - R; ry <= cy + R; ++ry) {
      if (ry < 0 || ry >= g.gy) continue;
      const bool edge_row = (ry < cy - prev) || (ry > cy + prev);
      // full row on the new rim, otherwise only the two new end segments
      for (int part = 0; part < 2; ++part) {
        int xa, xb;
        if (edge_row) {
          if (part == 1) break;
          xa = cx - R; xb = cx + R;
        } else {
          xa = part == 0 ? cx - R : cx + prev + 1;
          xb = part == 0 ? cx - prev - 1 : cx + R;
        }
        xa = max(xa, 0); xb = min(xb, g.gx - 1);
        if (xa > xb) continue;
        const int beg = cs[ry * g.gx + xa], end = cs[ry * g.gx + xb + 1];
        for (int t = beg; t < end; ++t) {
          const float4 c = sp[t];
          top3_insert(sqdist_ref(x1, y1, z1, c.x, c.y, c.z), __float_as_int(c.w), d0, d1, d2, i0, i1, i2);
        }
      }
    }
    prev = R;
    // everything not yet scanned is at least `reach` away in x or y
    const float mx = fminf(fx - (float)(cx - R), (float)(cx + R + 1) - fx);
    const float my = fminf(fy - (float)(cy - R), (float)(cy + R + 1) - fy);
    const float reach = fminf(mx, my) * g.h * 0.999f;
    done = reach > 0.f && d2 < reach * reach;
  }
  if (!done) {  // sparse neighbourhood: exact scan of all keys
    const float* __restrict__ kp = key + (int64_t)b * kst.b;
    d0 = d1 = d2 = INF; i0 = i1 = i2 = 0x7fffffff;
    for (int j = 0; j < Nk; ++j)
      top3_insert(sqdist_ref(x1, y1, z1, kp[(int64_t)j * kst.n], kp[(int64_t)j * kst.n + kst.c],
                             kp[(int64_t)j * kst.n + 2 * kst.c]), j, d0, d1, d2, i0, i1, i2);
  }
  const int64_t o = ((int64_t)b * Nq + i) * 3;
  if (index64) { index64[o] = i0; index64[o + 1] = i1; index64[o + 2] = i2; }
  if (dist) { dist[o] = d0; dist[o + 1] = d1; dist[o + 2] = d2; }
  if (!index32) return;
  index32[o] = i0; index32[o + 1] = i1; index32[o + 2] = i2;
  const float v0 = __fdiv_rn(1.0f, fmaxf(d0, 1e-10f));
  const float v1 = __fdiv_rn(1.0f, fmaxf(d1, 1e-10f));
  const float v2 = __fdiv_rn(1.0f, fmaxf(d2, 1e-10f));
  const float s = __fadd_rn(__fadd_rn(v0, v1), v2);
  weight[o] = __fdiv_rn(v0, s);
  weight[o + 1] = __fdiv_rn(v1, s);
  weight[o + 2] = __fdiv_rn(v2, s);
}

}  // namespace

int64_t grid_workspace_bytes(int B, int N) {
  return (int64_t)B * (sizeof(GridHeader) + sizeof(int) * (MAX_CELLS + 1) + sizeof(float4) * (int64_t)N) + 256;
}

GridPtrs grid_carve(void* ws, int B, int N) {
  GridPtrs g;
  g.sorted = reinterpret_cast<float4*>(ws);
  g.hdr = reinterpret_cast<GridHeader*>(g.sorted + (int64_t)B * N);
  g.cell_start = reinterpret_cast<int*>(g.hdr + B);
  return g;
}

int grid_build_launch(const float* pts, Strides3 st, int B, int N, float min_cell, void* ws, cudaStream_t stream,
                      bool stable) {
  const GridPtrs g = grid_carve(ws, B, N);
  if (stable) {
    RN_PREFER_MAX_SMEM(grid_build_kernel<true>);
    grid_build_kernel<true><<<B, 1024, 0, stream>>>(pts, st, N, min_cell, g.hdr, g.cell_start, g.sorted);
    RN_LAUNCH_CHECK("grid_build_kernel");
    return REGNET_OK;
  }
  RN_PREFER_MAX_SMEM(grid_build_kernel<false>);
  grid_build_kernel<false><<<B, 1024, 0, stream>>>(pts, st, N, min_cell, g.hdr, g.cell_start, g.sorted);
  RN_LAUNCH_CHECK("grid_build_kernel");
  return REGNET_OK;
}

int ball_query_grid_launch(const float* pts, Strides3 pst, const float* ctr, Strides3 cst, int B, int N, int M,
                           float radius, const void* ws, int32_t* index32, cudaStream_t stream, int64_t* index64,
                           int64_t* count64) {
  RN_CHECK_ARG(N <= 65536, "ball_query_grid: more than 65536 points per cloud");
  const GridPtrs g = grid_carve(const_cast<void*>(ws), B, N);
  dim3 grid(ceil_div(M, BQG_WARPS), B);
  RN_PREFER_MAX_SMEM(ball_query_grid_kernel);
  ball_query_grid_kernel<<<grid, BQG_WARPS * 32, 0, stream>>>(pts, pst, ctr, cst, N, M, radius, g.hdr, g.cell_start,
                                                             g.sorted, index32, index64, count64);
  RN_LAUNCH_CHECK("ball_query_grid_kernel");
  return REGNET_OK;
}

int three_nn_grid_launch(const float* qry, Strides3 qst, const float* key, Strides3 kst, int B, int Nq, int Nk,
                         const void* ws, int32_t* index32, float* weight, cudaStream_t stream, int64_t* index64,
                         float* dist) {
  const GridPtrs g = grid_carve(const_cast<void*>(ws), B, Nk);
  dim3 grid(ceil_div(Nq, 128), B);
  RN_PREFER_MAX_SMEM(three_nn_grid_kernel);
  three_nn_grid_kernel<<<grid, 128, 0, stream>>>(qry, qst, key, kst, Nq, Nk, g.hdr, g.cell_start, g.sorted, index32, weight,
                                                 index64, dist);
  RN_LAUNCH_CHECK("three_nn_grid_kernel");
  return REGNET_OK;
}

}  // namespace regnet
