/*
 * oracle/pn2_oracle.c -- CPU restatement of the reference's pn2_ext operators.
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (regnet_for_3d_grasping_b200) never does.
 *
 * Each function follows one reference kernel line by line (paths relative to
 * /root/reference/multi_model/utils/pn2_utils/csrc/):
 *
 *   oracle_fps            sampling_kernel.cu:47-117   (+ block-size rule :30-40, :146-165)
 *   oracle_ball_query     ball_query_kernel.cu:31-74  (+ r*r in fp32 :47)
 *   oracle_three_nn       interpolate_kernel.cu:28-77 (+ partial initialiser quirk :49-50)
 *   oracle_interpolate    interpolate_kernel.cu:134-177
 *   oracle_interpolate_bw interpolate_kernel.cu:239-282 (deterministic order instead of atomics)
 *   oracle_group          grouping_kernel.cu:45-48
 *   oracle_group_bw       grouping_kernel.cu:54-93    (deterministic order instead of atomics)
 *
 * Floating point: the reference is compiled by nvcc -O2 with fmad on; the SASS of all three search kernels
 * (see oracle/_ref/pn2_ext_ref.sass.txt, produced by oracle/build_ref.py) evaluates
 *     (x2-x1)*(x2-x1) + (y2-y1)*(y2-y1) + (z2-z1)*(z2-z1)
 * as  t = RN(dy*dy); t = fma(dx,dx,t); t = fma(dz,dz,t).  That order is pinned here with fmaf() and this file
 * must be compiled with -ffp-contract=off so that gcc adds no contraction of its own.
 *
 * Parity pin: tests/golden/ref_cuda_*.npz hold outputs of the reference kernels themselves (oracle/_ref run
 * on a B200 by oracle/gen_golden_gpu.py); tests/test_oracle_golden.py checks this file against them.
 *
 * Layouts: points are AoS (B, N, 3) contiguous, exactly what the reference kernels see after the host-side
 * transpose(1,2).contiguous() (sampling_kernel.cu:139, ball_query_kernel.cu:103-104, interpolate_kernel.cu:107-108).
 * Index outputs are int64 as in the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline float sqdist(float x1, float y1, float z1, float x2, float y2, float z2) {
  float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
  float t = dy * dy;            /* FMUL */
  t = fmaf(dx, dx, t);          /* FFMA */
  t = fmaf(dz, dz, t);          /* FFMA */
  return t;
}

/* sampling_kernel.cu:32-40 get_block(), then the switch at :148-164 (anything below 16 runs as 16) */
int oracle_fps_block_size(int64_t n) {
  int cnt = 0;
  int64_t x = n - 1;
  while (x > 0) { x >>= 1; cnt += 1; }
  int64_t b = (int64_t)1 << cnt;
  if (b > 512) b = 512;
  if (b < 16) b = 16;
  return (int)b;
}

/*
 * Farthest point sampling, emulating the CUDA block: `block` threads, thread t scans j = t, t+block, ...
 * keeping a strict-'>' running best starting from (0, cur); then the shared-memory tree reduction with a
 * strict '<' test (ties keep the lower slot).  temp[] starts at -1 (sampling_kernel.cu:142).
 * points: (B,N,3); index out: (B,M) int64.  Returns 0, or -1 on bad arguments (the reference's TORCH_CHECKs).
 */
int oracle_fps(const float* points, int64_t B, int64_t N, int64_t M, int64_t* index) {
  if (M <= 0 || N < M) return -1;
  const int block = oracle_fps_block_size(N);
  int err = 0;
  /* clouds run in parallel (FPS itself is sequential; splitting one cloud's scan across threads was measured
   * slower than serial because of the 5119 fork/joins) */
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < B; ++b) {
    const float* p = points + b * N * 3;
    int64_t* out = index + b * M;
    float* temp = (float*)malloc(sizeof(float) * (size_t)N);
    float sd[512];
    int32_t si[512];
    if (!temp) { err = -2; continue; }
    for (int64_t j = 0; j < N; ++j) temp[j] = -1.0f;
    int32_t cur = 0;
    out[0] = 0;
    for (int64_t i = 1; i < M; ++i) {
      const float x1 = p[cur * 3 + 0], y1 = p[cur * 3 + 1], z1 = p[cur * 3 + 2];
      for (int t = 0; t < block; ++t) {
        float max_dist = 0.0f;
        int32_t max_ind = cur;
        for (int64_t j = t; j < N; j += block) {
          float dist = sqdist(x1, y1, z1, p[j * 3 + 0], p[j * 3 + 1], p[j * 3 + 2]);
          float last = temp[j];
          if (last > dist || last < 0) temp[j] = dist; else dist = last;
          if (dist > max_dist) { max_dist = dist; max_ind = (int32_t)j; }
        }
        sd[t] = max_dist;
        si[t] = max_ind;
      }
      for (int offset = block / 2; offset > 0; offset /= 2) {
        for (int t = 0; t < offset; ++t) {
          if (sd[t] < sd[t + offset]) { sd[t] = sd[t + offset]; si[t] = si[t + offset]; }
        }
      }
      cur = si[0];
      out[i] = cur;
    }
    free(temp);
  }
  return err;
}

/*
 * Ball query (ball_query_kernel.cu:31-74).  points (B,N,3), centroids (B,M,3) -> index (B,M,K), count (B,M).
 * radius arrives as a C float (ball_query.h:10) and is squared in fp32.  Outputs are pre-zeroed by the
 * reference host code (:107-109): a centroid with no hit keeps zeros and count 0.
 */
int oracle_ball_query(const float* points, const float* centroids, int64_t B, int64_t N, int64_t M,
                      float radius, int64_t K, int64_t* index, int64_t* count) {
  const float r2 = radius * radius;
  memset(index, 0, sizeof(int64_t) * (size_t)(B * M * K));
  memset(count, 0, sizeof(int64_t) * (size_t)(B * M));
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    for (int64_t i = 0; i < M; ++i) {
      const float* p = points + b * N * 3;
      const float* c = centroids + (b * M + i) * 3;
      int64_t* idx = index + (b * M + i) * K;
      const float x1 = c[0], y1 = c[1], z1 = c[2];
      int64_t cnt = 0;
      for (int64_t j = 0; j < N && cnt < K; ++j) {
        float d = sqdist(x1, y1, z1, p[j * 3 + 0], p[j * 3 + 1], p[j * 3 + 2]);
        if (d < r2) {
          if (cnt == 0) { for (int64_t k = 0; k < K; ++k) idx[k] = j; }
          else idx[cnt] = j;
          ++cnt;
        }
      }
      count[b * M + i] = cnt;
    }
  }
  return 0;
}

/*
 * 3-NN search (interpolate_kernel.cu:28-77).  query (B,Nq,3), key (B,Nk,3) -> index (B,Nq,3) int64,
 * distance (B,Nq,3) fp32 (SQUARED).  `scalar_t min_dist[K] = {1e40}` is {+inf, 0, 0} in fp32 and
 * `int min_ind[K] = {-1}` is {-1, 0, 0}; with Nk >= 3 (enforced at :102) the zeros are shifted out.
 * The distance here is written (x1-x2)..., the square is sign-independent so sqdist() applies unchanged.
 */
int oracle_three_nn(const float* query, const float* key, int64_t B, int64_t Nq, int64_t Nk,
                    int64_t* index, float* distance) {
  if (Nk < 3) return -1;
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    for (int64_t i = 0; i < Nq; ++i) {
      const float* q = query + (b * Nq + i) * 3;
      const float* kp = key + b * Nk * 3;
      const float x1 = q[0], y1 = q[1], z1 = q[2];
      float md[3] = {(float)INFINITY, 0.0f, 0.0f};
      int mi[3] = {-1, 0, 0};
      for (int64_t j = 0; j < Nk; ++j) {
        float d = sqdist(x1, y1, z1, kp[j * 3 + 0], kp[j * 3 + 1], kp[j * 3 + 2]);
        for (int k = 0; k < 3; ++k) {
          if (d < md[k]) {
            for (int l = 2; l > k; --l) { md[l] = md[l - 1]; mi[l] = mi[l - 1]; }
            md[k] = d; mi[k] = (int)j;
            break;
          }
        }
      }
      for (int k = 0; k < 3; ++k) {
        index[(b * Nq + i) * 3 + k] = mi[k];
        distance[(b * Nq + i) * 3 + k] = md[k];
      }
    }
  }
  return 0;
}

/* interpolate_kernel.cu:134-177: out[b,c,n] = sum_k in[b,c,idx[b,n,k]] * w[b,n,k], k = 0,1,2 in order,
 * `outputValue += a*b` contracted to fma by nvcc.  input (B,C,Ns), out (B,C,Nd). Returns -1 on OOB index. */
int oracle_interpolate(const float* input, const int64_t* index, const float* weight, int64_t B, int64_t C,
                       int64_t Ns, int64_t Nd, float* out) {
  int err = 0;
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    for (int64_t c = 0; c < C; ++c) {
      const float* in = input + (b * C + c) * Ns;
      float* o = out + (b * C + c) * Nd;
      for (int64_t n = 0; n < Nd; ++n) {
        float acc = 0.0f;
        for (int k = 0; k < 3; ++k) {
          int64_t j = index[(b * Nd + n) * 3 + k];
          if (j < 0 || j >= Ns) { err = -1; continue; }
          acc = fmaf(in[j], weight[(b * Nd + n) * 3 + k], acc);
        }
        o[n] = acc;
      }
    }
  }
  return err;
}

/* interpolate_kernel.cu:239-282: grad_in[b,c,idx[b,n,k]] += grad_out[b,c,n] * w[b,n,k] (serial order here) */
int oracle_interpolate_bw(const float* grad_out, const int64_t* index, const float* weight, int64_t B,
                          int64_t C, int64_t Ns, int64_t Nd, float* grad_in) {
  memset(grad_in, 0, sizeof(float) * (size_t)(B * C * Ns));
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    for (int64_t c = 0; c < C; ++c) {
      const float* go = grad_out + (b * C + c) * Nd;
      float* gi = grad_in + (b * C + c) * Ns;
      for (int64_t n = 0; n < Nd; ++n)
        for (int k = 0; k < 3; ++k)
          gi[index[(b * Nd + n) * 3 + k]] += go[n] * weight[(b * Nd + n) * 3 + k];
    }
  }
  return 0;
}

/* grouping_kernel.cu:45-48: out[b,c,m,k] = in[b,c,idx[b,m,k]] */
int oracle_group(const float* input, const int64_t* index, int64_t B, int64_t C, int64_t N, int64_t M,
                 int64_t K, float* out) {
  int err = 0;
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    for (int64_t c = 0; c < C; ++c) {
      const float* in = input + (b * C + c) * N;
      float* o = out + (b * C + c) * M * K;
      const int64_t* idx = index + b * M * K;
      for (int64_t e = 0; e < M * K; ++e) {
        int64_t j = idx[e];
        if (j < 0 || j >= N) { err = -1; o[e] = 0.0f; } else o[e] = in[j];
      }
    }
  }
  return err;
}

/* grouping_kernel.cu:54-93: grad_in[b,c,idx[b,m,k]] += grad_out[b,c,m,k] (serial order here) */
int oracle_group_bw(const float* grad_out, const int64_t* index, int64_t B, int64_t C, int64_t N, int64_t M,
                    int64_t K, float* grad_in) {
  memset(grad_in, 0, sizeof(float) * (size_t)(B * C * N));
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    for (int64_t c = 0; c < C; ++c) {
      const float* go = grad_out + (b * C + c) * M * K;
      float* gi = grad_in + (b * C + c) * N;
      const int64_t* idx = index + b * M * K;
      for (int64_t e = 0; e < M * K; ++e) gi[idx[e]] += go[e];
    }
  }
  return 0;
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
