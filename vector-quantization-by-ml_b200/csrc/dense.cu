// dense.cu -- consumers of the dense N x K similarity matrix WITHOUT materialising it.
//
// Reference call sites (vector_quantization/vector_quantize_pytorch.py):
//   :284-299  calculate_ce_loss: F.cross_entropy(similarities, codes, ignore_index=-1) -- cross-entropy to given
//             indices (`forward(x, indices=...)`) and the cross-entropy commitment loss (:338-346)
//   :324-333  codebook diversity loss: softmax(-similarities * temperature) averaged over heads and batch
// and torch autograd through codebooks.py:386 (`-cdist` -> ATen _euclidean_dist_backward, einsum -> bmm).
//
// These losses need softmax statistics over ALL K scores of a row with fp32-exact scores (1e-5 relative on the loss),
// which the fp16 tensor-core scores of the search cannot give; they run on the CUDA cores as a tiled fp32 contraction
// (64 x 64 score tiles, 4 x 4 per thread, operands staged through shared memory in 16-wide k-chunks with a register
// prefetch of the next chunk) whose epilogue is an online softmax.  Nothing of size N x K is ever written:
//   pass "rowstats": per row  lse = log sum_k exp(alpha s_k)  and the score of the target code;
//   pass "avgprob" : avg[pos, k] = mean over heads and batch of exp(alpha s_k - lse)         (diversity forward);
//   pass "rowdot"  : per row  r = sum_k p_k G[pos, k]                                       (diversity backward, 1/2);
//   pass "backward": grad_x = sum_k dL/ds_k * ds_k/dx  as a second tiled contraction over the codes, the weights
//                    dL/ds_k recomputed tile by tile from lse (flash-attention style recomputation).
// Scores follow the reference's recipe: euclid  s = -sqrtf(max(|x|^2 + |c|^2 - 2 x.c, 0)),  dot  s = x.c.
// The backward pass takes TWO codebooks: the one the distances were computed with and the one whose rows are combined,
// because the reference's saved `embeddings.detach()` aliases the buffer that the EMA step overwrites before backward
// runs (codebooks.py:425 writes through `.data`): pre-update distances, post-update code vectors.  Pinned by the
// gradients recorded from the live reference (tests/golden/dense/).
#include <stdlib.h>

#include "common.cuh"

namespace vqb {
namespace {

constexpr int BM = 64;    // rows per tile
constexpr int BN = 64;    // codes per tile
// k-chunk staged per step: a template parameter (16 or 32, see dense_bk())
constexpr int LD = 68;    // padded leading dimension of the shared tiles (16-byte aligned rows)
constexpr int kThreads = 256;

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

// 4 consecutive elements of row `row` starting at dim `k`; zero outside [0, row_end) x [0, d)
template <typename T>
__device__ __forceinline__ float4 fetch4(const T* __restrict__ base, int64_t row, int64_t row_end, int d, int k, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < row_end && k < d) {
    const T* p = base + row * (int64_t)d + k;
    if (vec) {
      v = load4<T>(p);
    } else {
      v.x = to_f32<T>(p[0]);
      if (k + 1 < d) v.y = to_f32<T>(p[1]);
      if (k + 2 < d) v.z = to_f32<T>(p[2]);
      if (k + 3 < d) v.w = to_f32<T>(p[3]);
    }
  }
  return v;
}

__device__ __forceinline__ void stage4(float (*S)[LD], int r, int kq, const float4& v) {
  S[kq + 0][r] = v.x; S[kq + 1][r] = v.y; S[kq + 2][r] = v.z; S[kq + 3][r] = v.w;
}

// acc[i][j] = sum_k A[row0 + ty*4 + i][k] * B[col0 + tx*4 + j][k]   (zero-padded outside the bounds)
// Ends with a __syncthreads(): As / Bs may be re-used and shared scalars written before the call are visible.
template <int BKT, typename T, typename TB>
__device__ __forceinline__ void score_tile(const T* __restrict__ a_base, int64_t row0, int64_t row_end,
                                           const TB* __restrict__ b_base, int64_t col0, int64_t col_end, int d,
                                           bool vec, float (*As)[LD], float (*Bs)[LD], float acc[4][4]) {
  constexpr int U = BKT / 16;     // quads per thread and operand in one chunk
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int lr = tid >> 2, kq = (tid & 3) * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float4 pa[U], pb[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    pa[u] = fetch4<T>(a_base, row0 + lr, row_end, d, kq + 16 * u, vec);
    pb[u] = fetch4<TB>(b_base, col0 + lr, col_end, d, kq + 16 * u, vec);
  }
  for (int k0 = 0; k0 < d; k0 += BKT) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      stage4(As, lr, kq + 16 * u, pa[u]);
      stage4(Bs, lr, kq + 16 * u, pb[u]);
    }
    __syncthreads();
    if (k0 + BKT < d) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        pa[u] = fetch4<T>(a_base, row0 + lr, row_end, d, k0 + BKT + kq + 16 * u, vec);
        pb[u] = fetch4<TB>(b_base, col0 + lr, col_end, d, k0 + BKT + kq + 16 * u, vec);
      }
    }
#pragma unroll
    for (int kk = 0; kk < BKT; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
}

// the reference's score (codebooks.py:122-123 / :128-129): larger is more similar
__device__ __forceinline__ float sim_of(float dot, float xn2, float cn2, int metric) {
  if (metric == VQB_DOT) return dot;
  const float d2 = fmaxf(fmaf(-2.f, dot, xn2 + cn2), 0.f);
  return -sqrtf(d2);
}

// ---------------------------------------------------------------------------------------------------------------
// |row|^2, one warp per row (x1.pow(2).sum(-1) of _euclidean_dist)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) row_norm2_kernel(const T* __restrict__ x, int64_t rows, int d, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* p = x + row * (int64_t)d;
  float s = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float v = to_f32<T>(p[k]);
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// rowstats (MODE 0): lse[row] = log sum_k exp(alpha s_k), st[row] = s_target        grid (ceil(N/64), H)
// rowdot   (MODE 1): r[row]   = sum_k exp(alpha s_k - lse[row]) * table[row % n_pos][k]
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int MODE, int BKT>
__global__ void __launch_bounds__(kThreads) dense_rowstats_kernel(
    const T* __restrict__ x, const float* __restrict__ xn2, const float* __restrict__ cb, const float* __restrict__ cn2,
    int metric, float alpha, const int64_t* __restrict__ target, const float* __restrict__ lse_in,
    const float* __restrict__ table, int64_t n_pos, float* __restrict__ out0, float* __restrict__ st_out,
    int64_t N, int K, int d, int vec) {
  __shared__ __align__(16) float pool[2 * BKT * LD];
  float (*As)[LD] = reinterpret_cast<float (*)[LD]>(pool);
  float (*Bs)[LD] = reinterpret_cast<float (*)[LD]>(pool + BKT * LD);
  __shared__ float s_xn2[BM], s_cn2[BN], s_lse[BM];
  __shared__ long long s_t[BM];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t h = blockIdx.y, row0 = (int64_t)blockIdx.x * BM;
  const T* x_h = x + h * N * d;
  const float* cb_h = cb + h * (int64_t)K * d;
  if (tid < BM) {
    const int64_t row = row0 + tid;
    const bool ok = row < N;
    s_xn2[tid] = (ok && xn2 != nullptr) ? xn2[h * N + row] : 0.f;
    s_lse[tid] = (MODE == 1 && ok) ? lse_in[h * N + row] : 0.f;
    s_t[tid] = (MODE == 0 && ok && target != nullptr) ? (long long)target[h * N + row] : -1ll;
  }
  float m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = neg_inf(); l[i] = 0.f; }

  for (int64_t col0 = 0; col0 < K; col0 += BN) {
    __syncthreads();
    if (tid < BN) s_cn2[tid] = (cn2 != nullptr && col0 + tid < K) ? cn2[h * K + col0 + tid] : 0.f;
    float acc[4][4];
    score_tile<BKT, T, float>(x_h, row0, N, cb_h, col0, K, d, vec != 0, As, Bs, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const int64_t row = row0 + r;
      if (row >= N) continue;
      if (MODE == 0) {
        float zs[4];
        float tm = neg_inf();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t col = col0 + tx * 4 + j;
          zs[j] = neg_inf();
          if (col < K) {
            const float s = sim_of(acc[i][j], s_xn2[r], s_cn2[tx * 4 + j], metric);
            if (st_out != nullptr && (long long)col == s_t[r]) st_out[h * N + row] = s;
            zs[j] = alpha * s;
            tm = fmaxf(tm, zs[j]);
          }
        }
        if (tm > neg_inf()) {
          if (tm > m[i]) { l[i] *= expf(m[i] - tm); m[i] = tm; }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (zs[j] > neg_inf()) l[i] += expf(zs[j] - m[i]);
        }
      } else {
        const int64_t pos = row % n_pos;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t col = col0 + tx * 4 + j;
          if (col < K) {
            const float s = sim_of(acc[i][j], s_xn2[r], s_cn2[tx * 4 + j], metric);
            const float p = expf(alpha * s - s_lse[r]);
            l[i] = fmaf(p, __ldg(table + pos * K + col), l[i]);
          }
        }
      }
    }
  }
  // fixed-order merge over the 16 threads that share a row (lanes tx of one half-warp)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float mi = m[i], li = l[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, mi, o);
      const float lo = __shfl_xor_sync(0xffffffffu, li, o);
      if (MODE == 0) {
        const float mn = fmaxf(mi, mo);
        if (mn > neg_inf()) li = li * expf(mi - mn) + lo * expf(mo - mn);
        mi = mn;
      } else {
        li += lo;
      }
    }
    const int64_t row = row0 + ty * 4 + i;
    if (tx == 0 && row < N) out0[h * N + row] = (MODE == 0) ? mi + logf(li) : li;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// avgprob: avg[pos][k] = (1 / (H Bp)) sum_{h, b} exp(alpha s(h, b n_pos + pos, k) - lse)      grid (ceil(n_pos/64), ceil(K/64))
// A block owns a (64 positions x 64 codes) tile of the output and walks over heads and batch: no atomics, fixed order.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int BKT>
__global__ void __launch_bounds__(kThreads) dense_avgprob_kernel(
    const T* __restrict__ x, const float* __restrict__ xn2, const float* __restrict__ cb, const float* __restrict__ cn2,
    int metric, float alpha, const float* __restrict__ lse, float* __restrict__ avg, int64_t n_pos, int64_t H,
    int64_t N, int K, int d, int vec) {
  __shared__ __align__(16) float pool[2 * BKT * LD];
  float (*As)[LD] = reinterpret_cast<float (*)[LD]>(pool);
  float (*Bs)[LD] = reinterpret_cast<float (*)[LD]>(pool + BKT * LD);
  __shared__ float s_xn2[BM], s_cn2[BN], s_lse[BM];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t pos0 = (int64_t)blockIdx.x * BM, col0 = (int64_t)blockIdx.y * BN;
  const int64_t Bp = N / n_pos;
  float accp[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) accp[i][j] = 0.f;

  for (int64_t h = 0; h < H; ++h) {
    const T* x_h = x + h * N * d;
    const float* cb_h = cb + h * (int64_t)K * d;
    for (int64_t b = 0; b < Bp; ++b) {
      const int64_t row0 = b * n_pos + pos0, row_end = (b + 1) * n_pos;
      __syncthreads();
      if (tid < BM) {
        const int64_t row = row0 + tid;
        const bool ok = row < row_end;
        s_xn2[tid] = (ok && xn2 != nullptr) ? xn2[h * N + row] : 0.f;
        s_lse[tid] = ok ? lse[h * N + row] : 0.f;
        s_cn2[tid] = (cn2 != nullptr && col0 + tid < K) ? cn2[h * K + col0 + tid] : 0.f;
      }
      float acc[4][4];
      score_tile<BKT, T, float>(x_h, row0, row_end, cb_h, col0, K, d, vec != 0, As, Bs, acc);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = ty * 4 + i;
        if (row0 + r >= row_end) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (col0 + tx * 4 + j < K) {
            const float s = sim_of(acc[i][j], s_xn2[r], s_cn2[tx * 4 + j], metric);
            accp[i][j] += expf(alpha * s - s_lse[r]);
          }
        }
      }
    }
  }
  const float inv = 1.f / (float)(H * Bp);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t pos = pos0 + ty * 4 + i;
    if (pos >= n_pos) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t col = col0 + tx * 4 + j;
      if (col < K) avg[pos * K + col] = accp[i][j] * inv;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward: grad_x[row] = xcoef x[row] - sum_k rho_k c_comb[k]                 grid (ceil(N/64), H, ceil(d / (64 NSUB)))
//   w_k   = dL/ds_k = coef[row] (p_k - [k == target])                         (cross-entropy; table == NULL)
//         = coef[row] p_k (table[pos][k] - rdot[row])                         (diversity; coef = alpha / (H Bp))
//   euclid: s = -D,  ds/dx = -(x - c)/D:  rho_k = D_k > 0 ? -w_k / D_k : 0,  xcoef = sum_k rho_k
//           (ATen _euclidean_dist_backward: ratio = grad / D, masked where D == 0)
//   dot   : ds/dx = c:  rho_k = -w_k,  xcoef = 0
// The weights of a 64 x 64 tile go to shared memory and feed a second tiled contraction with the code vectors.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int NSUB, int BKT>
__global__ void __launch_bounds__(kThreads) dense_backward_kernel(
    const T* __restrict__ x, const float* __restrict__ xn2, const float* __restrict__ cb_dist,
    const float* __restrict__ cn2, const float* __restrict__ cb_comb, int metric, float alpha,
    const float* __restrict__ lse, const float* __restrict__ coef, const int64_t* __restrict__ target,
    const float* __restrict__ table, const float* __restrict__ rdot, int64_t n_pos, float* __restrict__ grad_x,
    int64_t N, int K, int d, int vec) {
  // the score operands (first contraction) and the code tile (second contraction) are live at different times and
  // share one pool
  constexpr int kPool = (2 * BKT > BN ? 2 * BKT : BN) * LD;
  __shared__ __align__(16) float pool[kPool];
  float (*As)[LD] = reinterpret_cast<float (*)[LD]>(pool);
  float (*Bs)[LD] = reinterpret_cast<float (*)[LD]>(pool + BKT * LD);
  float (*Cs)[LD] = reinterpret_cast<float (*)[LD]>(pool);   // code tile [code][dim]
  __shared__ __align__(16) float Rs[BM][LD];   // rho tile  [row][code]
  __shared__ float s_xn2[BM], s_cn2[BN], s_lse[BM], s_coef[BM], s_r[BM];
  __shared__ long long s_t[BM];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t h = blockIdx.y, row0 = (int64_t)blockIdx.x * BM;
  const int dim_base = blockIdx.z * 64 * NSUB;
  const T* x_h = x + h * N * d;
  const float* cd_h = cb_dist + h * (int64_t)K * d;
  const float* cc_h = cb_comb + h * (int64_t)K * d;
  if (tid < BM) {
    const int64_t row = row0 + tid;
    const bool ok = row < N;
    s_xn2[tid] = (ok && xn2 != nullptr) ? xn2[h * N + row] : 0.f;
    s_lse[tid] = ok ? lse[h * N + row] : 0.f;
    s_coef[tid] = ok ? coef[h * N + row] : 0.f;
    s_r[tid] = (ok && rdot != nullptr) ? rdot[h * N + row] : 0.f;
    s_t[tid] = (ok && target != nullptr) ? (long long)target[h * N + row] : -1ll;
  }
  float acc2[4][NSUB * 4];
  float rs[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    rs[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NSUB * 4; ++j) acc2[i][j] = 0.f;
  }

  for (int64_t col0 = 0; col0 < K; col0 += BN) {
    __syncthreads();
    if (tid < BN) s_cn2[tid] = (cn2 != nullptr && col0 + tid < K) ? cn2[h * K + col0 + tid] : 0.f;
    float acc[4][4];
    score_tile<BKT, T, float>(x_h, row0, N, cd_h, col0, K, d, vec != 0, As, Bs, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const int64_t row = row0 + r;
      float rho[4] = {0.f, 0.f, 0.f, 0.f};
      if (row < N) {
        const int64_t pos = (table != nullptr) ? row % n_pos : 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t col = col0 + tx * 4 + j;
          if (col < K) {
            const float s = sim_of(acc[i][j], s_xn2[r], s_cn2[tx * 4 + j], metric);
            const float p = expf(alpha * s - s_lse[r]);
            float w;
            if (table != nullptr) {
              w = s_coef[r] * p * (__ldg(table + pos * K + col) - s_r[r]);
            } else {
              w = s_coef[r] * p;
              if ((long long)col == s_t[r]) w -= s_coef[r];
            }
            if (metric == VQB_DOT) {
              rho[j] = -w;
            } else {
              const float D = -s;
              rho[j] = D > 0.f ? -w / D : 0.f;
            }
            rs[i] += rho[j];
          }
        }
      }
      *reinterpret_cast<float4*>(&Rs[r][tx * 4]) = make_float4(rho[0], rho[1], rho[2], rho[3]);
    }
    __syncthreads();
#pragma unroll
    for (int sub = 0; sub < NSUB; ++sub) {
      const int dim0 = dim_base + sub * 64;
      if (dim0 < d) {     // block-uniform
        // Cs[code][dim] = c_comb[col0 + code][dim0 + dim]: 64 x 64 floats, 4 quads per thread
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int idx = tid + kThreads * q;
          const int code = idx >> 4, dq = (idx & 15) * 4;
          const float4 v = fetch4<float>(cc_h, col0 + code, K, d, dim0 + dq, vec != 0);
          *reinterpret_cast<float4*>(&Cs[code][dq]) = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < BN; kk += 4) {
          float a[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(&Rs[ty * 4 + i][kk]);
            a[i][0] = t.x; a[i][1] = t.y; a[i][2] = t.z; a[i][3] = t.w;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Cs[kk + e][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc2[i][sub * 4 + j] = fmaf(a[i][e], b[j], acc2[i][sub * 4 + j]);
          }
        }
        __syncthreads();
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s = rs[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int64_t row = row0 + ty * 4 + i;
    if (row >= N) continue;
    const float xcoef = (metric == VQB_DOT) ? 0.f : s;
#pragma unroll
    for (int sub = 0; sub < NSUB; ++sub)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int dim = dim_base + sub * 64 + tx + 16 * j;
        if (dim < d) {
          const float xv = to_f32<T>(x_h[row * (int64_t)d + dim]);
          grad_x[(h * N + row) * (int64_t)d + dim] = fmaf(xcoef, xv, -acc2[i][sub * 4 + j]);
        }
      }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward with respect to the CODEBOOK (learnable codebook): the transposed contraction of dense_backward_kernel.
//   grad_c[k] = c_k sum_n rho_nk - sum_n rho_nk x_n        (euclid; ATen _euclidean_dist_backward, x2 side)
//             =                  - sum_n rho_nk x_n        (dot, rho = -w)
// Rows of the tile are CODES, columns are latents: the per-latent statistics (lse, coef, target, rdot, |x|^2) are
// re-staged for every latent tile.  grid (ceil(K/64), H * n_splits, ceil(d / (64 NSUB))): split s of a codebook
// walks the latent tiles s, s + n_splits, ... and writes its own partial (n_splits, H, K, d), summed by the caller
// in a fixed order (no atomics).
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int NSUB, int BKT>
__global__ void __launch_bounds__(kThreads) dense_backward_codes_kernel(
    const T* __restrict__ x, const float* __restrict__ xn2, const float* __restrict__ cb, const float* __restrict__ cn2,
    int metric, float alpha, const float* __restrict__ lse, const float* __restrict__ coef,
    const int64_t* __restrict__ target, const float* __restrict__ table, const float* __restrict__ rdot,
    int64_t n_pos, float* __restrict__ grad_c, int n_splits, int64_t H, int64_t N, int K, int d, int vec) {
  constexpr int kPool = (2 * BKT > BN ? 2 * BKT : BN) * LD;
  __shared__ __align__(16) float pool[kPool];
  float (*As)[LD] = reinterpret_cast<float (*)[LD]>(pool);
  float (*Bs)[LD] = reinterpret_cast<float (*)[LD]>(pool + BKT * LD);
  float (*Xs)[LD] = reinterpret_cast<float (*)[LD]>(pool);   // latent tile [latent][dim], after the scores are done
  __shared__ __align__(16) float Rs[BM][LD];   // rho tile  [code][latent]
  __shared__ float s_xn2[BN], s_cn2[BM], s_lse[BN], s_coef[BN], s_r[BN];
  __shared__ long long s_t[BN], s_pos[BN];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t h = blockIdx.y / n_splits;
  const int split = blockIdx.y % n_splits;
  const int64_t code0 = (int64_t)blockIdx.x * BM;
  const int dim_base = blockIdx.z * 64 * NSUB;
  const T* x_h = x + h * N * d;
  const float* cb_h = cb + h * (int64_t)K * d;
  if (tid < BM) s_cn2[tid] = (cn2 != nullptr && code0 + tid < K) ? cn2[h * K + code0 + tid] : 0.f;
  float acc2[4][NSUB * 4];
  float rs[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    rs[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NSUB * 4; ++j) acc2[i][j] = 0.f;
  }

  for (int64_t n0 = (int64_t)split * BN; n0 < N; n0 += (int64_t)n_splits * BN) {
    __syncthreads();
    if (tid < BN) {
      const int64_t row = n0 + tid;
      const bool ok = row < N;
      s_xn2[tid] = (ok && xn2 != nullptr) ? xn2[h * N + row] : 0.f;
      s_lse[tid] = ok ? lse[h * N + row] : 0.f;
      s_coef[tid] = ok ? coef[h * N + row] : 0.f;
      s_r[tid] = (ok && rdot != nullptr) ? rdot[h * N + row] : 0.f;
      s_t[tid] = (ok && target != nullptr) ? (long long)target[h * N + row] : -1ll;
      s_pos[tid] = (ok && table != nullptr) ? (long long)(row % n_pos) : 0ll;
    }
    float acc[4][4];
    score_tile<BKT, float, T>(cb_h, code0, K, x_h, n0, N, d, vec != 0, As, Bs, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const int64_t code = code0 + r;
      float rho[4] = {0.f, 0.f, 0.f, 0.f};
      if (code < K) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = tx * 4 + j;
          if (n0 + c < N) {
            const float sc = sim_of(acc[i][j], s_xn2[c], s_cn2[r], metric);
            const float p = expf(alpha * sc - s_lse[c]);
            float w;
            if (table != nullptr) {
              w = s_coef[c] * p * (__ldg(table + s_pos[c] * K + code) - s_r[c]);
            } else {
              w = s_coef[c] * p;
              if ((long long)code == s_t[c]) w -= s_coef[c];
            }
            if (metric == VQB_DOT) {
              rho[j] = -w;
            } else {
              const float D = -sc;
              rho[j] = D > 0.f ? -w / D : 0.f;
            }
            rs[i] += rho[j];
          }
        }
      }
      *reinterpret_cast<float4*>(&Rs[r][tx * 4]) = make_float4(rho[0], rho[1], rho[2], rho[3]);
    }
    __syncthreads();
#pragma unroll
    for (int sub = 0; sub < NSUB; ++sub) {
      const int dim0 = dim_base + sub * 64;
      if (dim0 < d) {     // block-uniform
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int idx = tid + kThreads * q;
          const int lat = idx >> 4, dq = (idx & 15) * 4;
          const float4 v = fetch4<T>(x_h, n0 + lat, N, d, dim0 + dq, vec != 0);
          *reinterpret_cast<float4*>(&Xs[lat][dq]) = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < BN; kk += 4) {
          float a[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(&Rs[ty * 4 + i][kk]);
            a[i][0] = t.x; a[i][1] = t.y; a[i][2] = t.z; a[i][3] = t.w;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Xs[kk + e][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc2[i][sub * 4 + j] = fmaf(a[i][e], b[j], acc2[i][sub * 4 + j]);
          }
        }
        __syncthreads();
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float sum = rs[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const int64_t code = code0 + ty * 4 + i;
    if (code >= K) continue;
    const float ccoef = (metric == VQB_DOT) ? 0.f : sum;
    float* out = grad_c + (((int64_t)split * H + h) * K + code) * (int64_t)d;
#pragma unroll
    for (int sub = 0; sub < NSUB; ++sub)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int dim = dim_base + sub * 64 + tx + 16 * j;
        if (dim < d) out[dim] = fmaf(ccoef, cb_h[code * (int64_t)d + dim], -acc2[i][sub * 4 + j]);
      }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Stochastic (gumbel) sampling of the code, reference utils/general.py:107-129 called at codebooks.py:388:
//   ind = argmax_k ( fl(s_k / temperature) + g_k ),   g = -log(-log(u)),  log(t) = ln(max(t, 1e-5)),  u ~ U[0,1)
// as one more epilogue of the tiled score pass: neither the N x K similarities nor the N x K noise exist in HBM.
// The uniforms are either read from a caller-provided (H,N,K) tensor (tests: the draw a fixture was recorded with)
// or generated in place with Philox4x32-10 in the layout of ATen's `uniform_` CUDA kernel for a tensor of H*N*K
// elements (distribution_elementwise_grid_stride_kernel, unroll 4, T = 256 * grid threads): element `li` is output
// (li / T) % 4 of the ((li / T) / 4)-th curand_uniform4 call of thread li % T, i.e. of the Philox block
// ctr = {offset/4 + (li / T) / 4, subsequence = li % T}, key = seed.  With the generator's (seed, offset) and the same T
// the kernel draws exactly what `torch.zeros_like(similarities).uniform_(0, 1)` would have drawn.
// ---------------------------------------------------------------------------------------------------------------
struct Philox {
  unsigned long long seed, offset4;   // generator seed, philox offset / 4
  unsigned int T;                     // threads of the equivalent ATen launch (0: read `u` instead)
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr unsigned int kA = 0xD2511F53u, kB = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(kA, c.x), lo0 = kA * c.x;
    const unsigned int hi1 = __umulhi(kB, c.z), lo1 = kB * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += kW0; k.y += kW1;
  }
  return c;
}

__device__ __forceinline__ float philox_uniform(const Philox& P, unsigned long long li) {
  const unsigned long long q = li / P.T;
  const unsigned int sub = (unsigned int)(li - q * P.T);
  const unsigned long long blk = P.offset4 + (q >> 2);
  const uint4 r = philox4x32_10(make_uint4((unsigned int)blk, (unsigned int)(blk >> 32), sub, 0u),
                                make_uint2((unsigned int)P.seed, (unsigned int)(P.seed >> 32)));
  const unsigned int w = (q & 3) == 0 ? r.x : ((q & 3) == 1 ? r.y : ((q & 3) == 2 ? r.z : r.w));
  // curand_uniform: (0, 1];  ATen's uniform_(0, 1) maps 1.0 to 0.0
  const float v = fmaf((float)w, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
  return v == 1.0f ? 0.0f : v;
}

__device__ __forceinline__ float gumbel_of(float u) {
  const float a = logf(fmaxf(u, 1e-5f));
  return -logf(fmaxf(-a, 1e-5f));
}

// MODE 0: sample (idx_out);  MODE 1: write the similarities (scores_out (H,N,K) -- the opt-in third return value of
// Codebook.forward, codebooks.py:433-435)                                              grid (ceil(N/64), H)
template <typename T, int MODE, int BKT>
__global__ void __launch_bounds__(kThreads) dense_sample_kernel(
    const T* __restrict__ x, const float* __restrict__ xn2, const float* __restrict__ cb, const float* __restrict__ cn2,
    int metric, float temperature, const float* __restrict__ u, const Philox P, int64_t* __restrict__ idx_out,
    float* __restrict__ scores_out, int64_t N, int K, int d, int vec) {
  __shared__ __align__(16) float pool[2 * BKT * LD];
  float (*As)[LD] = reinterpret_cast<float (*)[LD]>(pool);
  float (*Bs)[LD] = reinterpret_cast<float (*)[LD]>(pool + BKT * LD);
  __shared__ float s_xn2[BM], s_cn2[BN];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t h = blockIdx.y, row0 = (int64_t)blockIdx.x * BM;
  const T* x_h = x + h * N * d;
  const float* cb_h = cb + h * (int64_t)K * d;
  if (tid < BM) {
    const int64_t row = row0 + tid;
    s_xn2[tid] = (row < N && xn2 != nullptr) ? xn2[h * N + row] : 0.f;
  }
  float best[4];
  int bidx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { best[i] = neg_inf(); bidx[i] = 0x7fffffff; }

  for (int64_t col0 = 0; col0 < K; col0 += BN) {
    __syncthreads();
    if (tid < BN) s_cn2[tid] = (cn2 != nullptr && col0 + tid < K) ? cn2[h * K + col0 + tid] : 0.f;
    float acc[4][4];
    score_tile<BKT, T, float>(x_h, row0, N, cb_h, col0, K, d, vec != 0, As, Bs, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const int64_t row = row0 + r;
      if (row >= N) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t col = col0 + tx * 4 + j;
        if (col >= K) continue;
        const float sc = sim_of(acc[i][j], s_xn2[r], s_cn2[tx * 4 + j], metric);
        const unsigned long long li = (unsigned long long)((h * N + row) * (int64_t)K + col);
        if (MODE == 1) {
          scores_out[li] = sc;
        } else {
          const float uu = P.T ? philox_uniform(P, li) : __ldg(u + li);
          const float v = __fadd_rn(__fdiv_rn(sc, temperature), gumbel_of(uu));
          // first maximum wins (torch.argmax); columns are visited in increasing order per thread
          if (v > best[i] || (v == best[i] && (int)col < bidx[i]) || bidx[i] == 0x7fffffff) { best[i] = v; bidx[i] = (int)col; }
        }
      }
    }
  }
  if (MODE == 1) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float bv = best[i];
    int bi = bidx[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    const int64_t row = row0 + ty * 4 + i;
    if (tx == 0 && row < N) idx_out[h * N + row] = (int64_t)bi;
  }
}

int check_common(const void* x, const void* cb, int64_t H, int64_t N, int K, int d, int metric) {
  VQB_REQUIRE((x != nullptr || N == 0) && cb != nullptr, VQB_ERR_INVALID, "dense: null pointer");
  VQB_REQUIRE(H >= 1 && H <= 65535 && N >= 0 && K >= 1 && d >= 1, VQB_ERR_INVALID,
              "dense: bad shape H=%lld N=%lld K=%d d=%d", (long long)H, (long long)N, K, d);
  VQB_REQUIRE((N + BM - 1) / BM < (1ll << 31), VQB_ERR_UNSUPPORTED, "dense: N=%lld too large", (long long)N);
  VQB_REQUIRE(metric == VQB_EUCLID || metric == VQB_DOT, VQB_ERR_INVALID, "dense: unknown metric %d", metric);
  return VQB_OK;
}

// k-chunk of the score contraction: 16 or 32 floats per operand row and step (VQB_DENSE_BK overrides the default).
// With 32 a chunk is ~1100 issue cycles per block, longer than the global-load latency of the prefetched next chunk.
inline int dense_bk() {
  static const int bk = [] {
    const char* e = getenv("VQB_DENSE_BK");
    const int v = e ? atoi(e) : 16;
    return v == 32 ? 32 : 16;
  }();
  return bk;
}

inline int vec_ok(const void* x, const void* a, const void* b, int d) {
  return (d % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0);
}

}  // namespace
}  // namespace vqb

using namespace vqb;

extern "C" int vqb_dense_row_norms(const void* x, int x_dtype, int64_t rows, int d, float* out, void* stream) {
  VQB_REQUIRE(rows >= 0 && d >= 1, VQB_ERR_INVALID, "vqb_dense_row_norms: bad shape");
  if (rows == 0) return VQB_OK;
  VQB_REQUIRE(x != nullptr && out != nullptr, VQB_ERR_INVALID, "vqb_dense_row_norms: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((rows + 7) / 8);
  VQB_DISPATCH_DTYPE(x_dtype, T, row_norm2_kernel<T><<<blocks, 256, 0, st>>>((const T*)x, rows, d, out));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

extern "C" int vqb_dense_rowstats(const void* x, int x_dtype, const float* xn2, const float* codebook,
                                  const float* cn2, int metric, float alpha, const int64_t* target,
                                  float* lse_out, float* target_score_out, int64_t H, int64_t N, int K, int d,
                                  void* stream) {
  if (int rc = check_common(x, codebook, H, N, K, d, metric)) return rc;
  VQB_REQUIRE(lse_out != nullptr || N == 0, VQB_ERR_INVALID, "vqb_dense_rowstats: lse_out is null");
  VQB_REQUIRE(metric == VQB_DOT || N == 0 || (xn2 != nullptr && cn2 != nullptr), VQB_ERR_INVALID,
              "vqb_dense_rowstats: the Euclidean metric needs the row norms of x and of the codebook");
  if (N == 0) return VQB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((unsigned)((N + BM - 1) / BM), (unsigned)H);
  const int vec = vec_ok(x, codebook, codebook, d);
#define VQB_DENSE_RS0(BKT)                                                                                  \
  VQB_DISPATCH_DTYPE(x_dtype, T, dense_rowstats_kernel<T, 0, BKT><<<grid, kThreads, 0, st>>>(                 \
      (const T*)x, xn2, codebook, cn2, metric, alpha, target, nullptr, nullptr, 1, lse_out, target_score_out, \
      N, K, d, vec))
  if (dense_bk() == 32) { VQB_DENSE_RS0(32); } else { VQB_DENSE_RS0(16); }
#undef VQB_DENSE_RS0
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

extern "C" int vqb_dense_rowdot(const void* x, int x_dtype, const float* xn2, const float* codebook, const float* cn2,
                                int metric, float alpha, const float* lse, const float* table, int64_t n_pos,
                                float* rdot_out, int64_t H, int64_t N, int K, int d, void* stream) {
  if (int rc = check_common(x, codebook, H, N, K, d, metric)) return rc;
  VQB_REQUIRE(N == 0 || (lse != nullptr && table != nullptr && rdot_out != nullptr), VQB_ERR_INVALID,
              "vqb_dense_rowdot: null pointer");
  VQB_REQUIRE(n_pos >= 1 && N % n_pos == 0, VQB_ERR_INVALID, "vqb_dense_rowdot: N=%lld is not a multiple of n_pos=%lld",
              (long long)N, (long long)n_pos);
  VQB_REQUIRE(metric == VQB_DOT || N == 0 || (xn2 != nullptr && cn2 != nullptr), VQB_ERR_INVALID,
              "vqb_dense_rowdot: the Euclidean metric needs the row norms");
  if (N == 0) return VQB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((unsigned)((N + BM - 1) / BM), (unsigned)H);
  const int vec = vec_ok(x, codebook, codebook, d);
#define VQB_DENSE_RS1(BKT)                                                                                       \
  VQB_DISPATCH_DTYPE(x_dtype, T, dense_rowstats_kernel<T, 1, BKT><<<grid, kThreads, 0, st>>>(                      \
      (const T*)x, xn2, codebook, cn2, metric, alpha, nullptr, lse, table, n_pos, rdot_out, nullptr, N, K, d, vec))
  if (dense_bk() == 32) { VQB_DENSE_RS1(32); } else { VQB_DENSE_RS1(16); }
#undef VQB_DENSE_RS1
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

extern "C" int vqb_dense_avgprob(const void* x, int x_dtype, const float* xn2, const float* codebook, const float* cn2,
                                 int metric, float alpha, const float* lse, float* avg_out, int64_t n_pos,
                                 int64_t H, int64_t N, int K, int d, void* stream) {
  if (int rc = check_common(x, codebook, H, N, K, d, metric)) return rc;
  VQB_REQUIRE(lse != nullptr && avg_out != nullptr, VQB_ERR_INVALID, "vqb_dense_avgprob: null pointer");
  VQB_REQUIRE(n_pos >= 1 && N >= 1 && N % n_pos == 0, VQB_ERR_INVALID,
              "vqb_dense_avgprob: N=%lld is not a positive multiple of n_pos=%lld", (long long)N, (long long)n_pos);
  VQB_REQUIRE((K + BN - 1) / BN <= 65535, VQB_ERR_UNSUPPORTED, "vqb_dense_avgprob: K=%d too large", K);
  VQB_REQUIRE(metric == VQB_DOT || (xn2 != nullptr && cn2 != nullptr), VQB_ERR_INVALID,
              "vqb_dense_avgprob: the Euclidean metric needs the row norms");
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((unsigned)((n_pos + BM - 1) / BM), (unsigned)((K + BN - 1) / BN));
  const int vec = vec_ok(x, codebook, codebook, d);
#define VQB_DENSE_AP(BKT)                                                                        \
  VQB_DISPATCH_DTYPE(x_dtype, T, dense_avgprob_kernel<T, BKT><<<grid, kThreads, 0, st>>>(         \
      (const T*)x, xn2, codebook, cn2, metric, alpha, lse, avg_out, n_pos, H, N, K, d, vec))
  if (dense_bk() == 32) { VQB_DENSE_AP(32); } else { VQB_DENSE_AP(16); }
#undef VQB_DENSE_AP
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

extern "C" int vqb_dense_backward(const void* x, int x_dtype, const float* xn2, const float* codebook_dist,
                                  const float* cn2, const float* codebook_comb, int metric, float alpha,
                                  const float* lse, const float* coef, const int64_t* target, const float* table,
                                  const float* rdot, int64_t n_pos, float* grad_x, int64_t H, int64_t N, int K, int d,
                                  void* stream) {
  if (int rc = check_common(x, codebook_dist, H, N, K, d, metric)) return rc;
  VQB_REQUIRE(N == 0 || (codebook_comb != nullptr && lse != nullptr && coef != nullptr && grad_x != nullptr),
              VQB_ERR_INVALID, "vqb_dense_backward: null pointer");
  VQB_REQUIRE((target != nullptr) != (table != nullptr), VQB_ERR_INVALID,
              "vqb_dense_backward: exactly one of target (cross-entropy) and table (diversity) must be given");
  VQB_REQUIRE(table == nullptr || (rdot != nullptr && n_pos >= 1 && N % n_pos == 0), VQB_ERR_INVALID,
              "vqb_dense_backward: the diversity form needs rdot and N a multiple of n_pos");
  VQB_REQUIRE(metric == VQB_DOT || N == 0 || (xn2 != nullptr && cn2 != nullptr), VQB_ERR_INVALID,
              "vqb_dense_backward: the Euclidean metric needs the row norms");
  if (N == 0) return VQB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = vec_ok(x, codebook_dist, codebook_comb, d);
  const int nsub = d <= 64 ? 1 : (d <= 128 ? 2 : 4);
  const dim3 grid((unsigned)((N + BM - 1) / BM), (unsigned)H, (unsigned)((d + 64 * nsub - 1) / (64 * nsub)));
  VQB_REQUIRE(grid.z <= 65535, VQB_ERR_UNSUPPORTED, "vqb_dense_backward: d=%d too large", d);
  if (n_pos < 1) n_pos = 1;
#define VQB_DENSE_BWD(NS, BKT)                                                                                   \
  VQB_DISPATCH_DTYPE(x_dtype, T, dense_backward_kernel<T, NS, BKT><<<grid, kThreads, 0, st>>>(                     \
      (const T*)x, xn2, codebook_dist, cn2, codebook_comb, metric, alpha, lse, coef, target, table, rdot, n_pos, \
      grad_x, N, K, d, vec))
  if (dense_bk() == 32) {
    if (nsub == 1) { VQB_DENSE_BWD(1, 32); } else if (nsub == 2) { VQB_DENSE_BWD(2, 32); } else { VQB_DENSE_BWD(4, 32); }
  } else {
    if (nsub == 1) { VQB_DENSE_BWD(1, 16); } else if (nsub == 2) { VQB_DENSE_BWD(2, 16); } else { VQB_DENSE_BWD(4, 16); }
  }
#undef VQB_DENSE_BWD
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

// number of partial sums vqb_dense_backward_codes writes (first dimension of grad_c_partial): enough blocks to fill
// the GPU when K is small, 1 when the code tiles alone do
extern "C" int vqb_dense_backward_codes_splits(int64_t H, int64_t N, int K, int d) {
  if (H < 1 || N < 1 || K < 1 || d < 1) return 1;
  const int nsub = d <= 64 ? 1 : (d <= 128 ? 2 : 4);
  const int64_t blocks = (int64_t)((K + BM - 1) / BM) * H * ((d + 64 * nsub - 1) / (64 * nsub));
  int64_t s = (2 * 148 + blocks - 1) / blocks;
  const int64_t tiles = (N + BN - 1) / BN;
  if (s > tiles) s = tiles;
  if (s > 32) s = 32;
  if (s < 1) s = 1;
  return (int)s;
}

extern "C" int vqb_dense_backward_codes(const void* x, int x_dtype, const float* xn2, const float* codebook,
                                        const float* cn2, int metric, float alpha, const float* lse,
                                        const float* coef, const int64_t* target, const float* table,
                                        const float* rdot, int64_t n_pos, float* grad_c_partial, int n_splits,
                                        int64_t H, int64_t N, int K, int d, void* stream) {
  if (int rc = check_common(x, codebook, H, N, K, d, metric)) return rc;
  VQB_REQUIRE(N >= 1 && lse != nullptr && coef != nullptr && grad_c_partial != nullptr, VQB_ERR_INVALID,
              "vqb_dense_backward_codes: null pointer or empty input");
  VQB_REQUIRE((target != nullptr) != (table != nullptr), VQB_ERR_INVALID,
              "vqb_dense_backward_codes: exactly one of target (cross-entropy) and table (diversity) must be given");
  VQB_REQUIRE(table == nullptr || (rdot != nullptr && n_pos >= 1 && N % n_pos == 0), VQB_ERR_INVALID,
              "vqb_dense_backward_codes: the diversity form needs rdot and N a multiple of n_pos");
  VQB_REQUIRE(metric == VQB_DOT || (xn2 != nullptr && cn2 != nullptr), VQB_ERR_INVALID,
              "vqb_dense_backward_codes: the Euclidean metric needs the row norms");
  VQB_REQUIRE(n_splits >= 1 && n_splits == vqb_dense_backward_codes_splits(H, N, K, d), VQB_ERR_INVALID,
              "vqb_dense_backward_codes: n_splits=%d, expected vqb_dense_backward_codes_splits() = %d", n_splits,
              vqb_dense_backward_codes_splits(H, N, K, d));
  VQB_REQUIRE(H * n_splits <= 65535, VQB_ERR_UNSUPPORTED, "vqb_dense_backward_codes: H * n_splits too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = vec_ok(x, codebook, codebook, d);
  const int nsub = d <= 64 ? 1 : (d <= 128 ? 2 : 4);
  const dim3 grid((unsigned)((K + BM - 1) / BM), (unsigned)(H * n_splits), (unsigned)((d + 64 * nsub - 1) / (64 * nsub)));
  VQB_REQUIRE(grid.z <= 65535, VQB_ERR_UNSUPPORTED, "vqb_dense_backward_codes: d=%d too large", d);
  if (n_pos < 1) n_pos = 1;
#define VQB_DENSE_BWDC(NS, BKT)                                                                                     \
  VQB_DISPATCH_DTYPE(x_dtype, T, dense_backward_codes_kernel<T, NS, BKT><<<grid, kThreads, 0, st>>>(                 \
      (const T*)x, xn2, codebook, cn2, metric, alpha, lse, coef, target, table, rdot, n_pos, grad_c_partial,        \
      n_splits, H, N, K, d, vec))
  if (dense_bk() == 32) {
    if (nsub == 1) { VQB_DENSE_BWDC(1, 32); } else if (nsub == 2) { VQB_DENSE_BWDC(2, 32); } else { VQB_DENSE_BWDC(4, 32); }
  } else {
    if (nsub == 1) { VQB_DENSE_BWDC(1, 16); } else if (nsub == 2) { VQB_DENSE_BWDC(2, 16); } else { VQB_DENSE_BWDC(4, 16); }
  }
#undef VQB_DENSE_BWDC
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}


extern "C" int vqb_dense_gumbel_sample(const void* x, int x_dtype, const float* xn2, const float* codebook,
                                       const float* cn2, int metric, float temperature, const float* uniforms,
                                       uint64_t philox_seed, uint64_t philox_offset, uint32_t philox_threads,
                                       int64_t* idx_out, int64_t H, int64_t N, int K, int d, void* stream) {
  if (int rc = check_common(x, codebook, H, N, K, d, metric)) return rc;
  VQB_REQUIRE(idx_out != nullptr || N == 0, VQB_ERR_INVALID, "vqb_dense_gumbel_sample: idx_out is null");
  VQB_REQUIRE(temperature > 0.f, VQB_ERR_INVALID, "vqb_dense_gumbel_sample: temperature must be > 0");
  VQB_REQUIRE(uniforms != nullptr || philox_threads > 0, VQB_ERR_INVALID,
              "vqb_dense_gumbel_sample: pass the uniforms or the Philox stream (seed, offset, threads)");
  VQB_REQUIRE(philox_offset % 4 == 0, VQB_ERR_INVALID, "vqb_dense_gumbel_sample: philox offset must be a multiple of 4");
  VQB_REQUIRE(metric == VQB_DOT || N == 0 || (xn2 != nullptr && cn2 != nullptr), VQB_ERR_INVALID,
              "vqb_dense_gumbel_sample: the Euclidean metric needs the row norms");
  if (N == 0) return VQB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((unsigned)((N + BM - 1) / BM), (unsigned)H);
  const int vec = vec_ok(x, codebook, codebook, d);
  Philox P;
  P.seed = philox_seed; P.offset4 = philox_offset / 4; P.T = uniforms ? 0u : philox_threads;
  VQB_DISPATCH_DTYPE(x_dtype, T, (dense_sample_kernel<T, 0, 16><<<grid, kThreads, 0, st>>>(
      (const T*)x, xn2, codebook, cn2, metric, temperature, uniforms, P, idx_out, nullptr, N, K, d, vec)));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

extern "C" int vqb_dense_scores(const void* x, int x_dtype, const float* xn2, const float* codebook, const float* cn2,
                                int metric, float* scores_out, int64_t H, int64_t N, int K, int d, void* stream) {
  if (int rc = check_common(x, codebook, H, N, K, d, metric)) return rc;
  VQB_REQUIRE(scores_out != nullptr || N == 0, VQB_ERR_INVALID, "vqb_dense_scores: scores_out is null");
  VQB_REQUIRE(metric == VQB_DOT || N == 0 || (xn2 != nullptr && cn2 != nullptr), VQB_ERR_INVALID,
              "vqb_dense_scores: the Euclidean metric needs the row norms");
  if (N == 0) return VQB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((unsigned)((N + BM - 1) / BM), (unsigned)H);
  const int vec = vec_ok(x, codebook, codebook, d);
  Philox P = {};
  VQB_DISPATCH_DTYPE(x_dtype, T, (dense_sample_kernel<T, 1, 16><<<grid, kThreads, 0, st>>>(
      (const T*)x, xn2, codebook, cn2, metric, 1.f, nullptr, P, nullptr, scores_out, N, K, d, vec)));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}
