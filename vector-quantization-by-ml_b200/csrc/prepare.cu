// prepare.cu -- operand preparation for the tensor-core search:
//   * codebook cache: bf16(-c) padded copy + per-code norms / rounding residual norms
//   * latents: bf16 padded copy + global max row norm / max rounding residual norm
//   * per-search lower-bound bias  L_k = |c_k|^2/2 - E_k  (see search_resolve.cu for the proof sketch)
//   * l2norm of rows (reference utils/losses.py:19)
#include "common.cuh"

namespace vqb {

// one warp per code row (incl. padded rows k in [K, Kp))
__global__ void prepare_codebook_kernel(const float* __restrict__ cb, int64_t H, int K, int Kp, int d, int dp,
                                        int metric, __nv_bfloat16* __restrict__ out, float* __restrict__ cn2h,
                                        float* __restrict__ cn, float* __restrict__ dcn) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= H * (int64_t)Kp) return;
  const int64_t h = row / Kp;
  const int k = (int)(row - h * Kp);
  __nv_bfloat16* o = out + row * dp;
  if (k >= K) {
    for (int j = lane; j < dp; j += 32) o[j] = __float2bfloat16(0.f);
    if (lane == 0) { cn2h[row] = kPadBias; cn[row] = 0.f; dcn[row] = 0.f; }
    return;
  }
  const float* c = cb + (h * K + k) * (int64_t)d;
  double n2 = 0.0, r2 = 0.0;
  for (int j = lane; j < dp; j += 32) {
    float v = j < d ? c[j] : 0.f;
    __nv_bfloat16 b = __float2bfloat16(v);
    float back = __bfloat162float(b);
    o[j] = __float2bfloat16(-back);                  // exact negation: the MMA then yields -x.c
    n2 = fma((double)v, (double)v, n2);
    double r = (double)v - (double)back;
    r2 = fma(r, r, r2);
  }
  n2 = warp_sum(n2);
  r2 = warp_sum(r2);
  if (lane == 0) {
    cn2h[row] = metric == VQB_EUCLID ? (float)(0.5 * n2) : 0.f;
    cn[row] = __double2float_ru(sqrt(n2)) * 1.000001f;
    dcn[row] = __double2float_ru(sqrt(r2)) * 1.000001f;
  }
}

// one warp per latent row: bf16 copy (zero padded to dp) + atomicMax of |x_b| and |x - x_b|
template <typename T>
__global__ void prepare_latents_kernel(const T* __restrict__ x, int64_t rows, int d, int dp,
                                       __nv_bfloat16* __restrict__ xb, uint32_t* __restrict__ scal) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float n2 = 0.f, r2 = 0.f;
  if (row < rows) {
    const T* xr = x + row * (int64_t)d;
    __nv_bfloat16* o = xb + row * (int64_t)dp;
    if ((d & 3) == 0) {
      for (int j = lane * 4; j < dp; j += 128) {
        float4 v = j < d ? load4<T>(xr + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        __nv_bfloat162 b0 = __floats2bfloat162_rn(v.x, v.y), b1 = __floats2bfloat162_rn(v.z, v.w);
        float2 f0 = __bfloat1622float2(b0), f1 = __bfloat1622float2(b1);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&b0);
        pk.y = *reinterpret_cast<uint32_t*>(&b1);
        *reinterpret_cast<uint2*>(o + j) = pk;
        n2 += f0.x * f0.x + f0.y * f0.y + f1.x * f1.x + f1.y * f1.y;
        float e0 = v.x - f0.x, e1 = v.y - f0.y, e2 = v.z - f1.x, e3 = v.w - f1.y;
        r2 += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
      }
    } else {
      for (int j = lane; j < dp; j += 32) {
        float v = j < d ? to_f32<T>(xr[j]) : 0.f;
        __nv_bfloat16 b = __float2bfloat16(v);
        float back = __bfloat162float(b);
        o[j] = b;
        n2 += back * back;
        r2 += (v - back) * (v - back);
      }
    }
  }
  n2 = warp_sum(n2);
  r2 = warp_sum(r2);
  // block-level max first, then one atomic per block
  __shared__ float s_n[32], s_r[32];
  const int w = threadIdx.x >> 5;
  if (lane == 0) { s_n[w] = n2; s_r[w] = r2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mn = 0.f, mr = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { mn = fmaxf(mn, s_n[i]); mr = fmaxf(mr, s_r[i]); }
    // fp32 accumulation of a sum of squares: inflate by (1 + dp * 2^-23) before the square root
    float infl = 1.f + (float)dp * 2.4e-7f;
    float a = sqrtf(mn * infl) * 1.00001f, b = sqrtf(mr * infl) * 1.00001f;
    atomicMax(scal + 0, __float_as_uint(a));   // non-negative floats order like their bit patterns
    atomicMax(scal + 1, __float_as_uint(b));
  }
}

int launch_prepare_latents(const void* x, int x_dtype, int64_t rows, int d, int dp,
                           __nv_bfloat16* xb, uint32_t* scal, cudaStream_t st) {
  const int warps = 8;
  const int64_t blocks = (rows + warps - 1) / warps;
  VQB_REQUIRE(blocks < (1ll << 31), VQB_ERR_UNSUPPORTED, "too many latent rows: %lld", (long long)rows);
  VQB_DISPATCH_DTYPE(x_dtype, T,
    prepare_latents_kernel<T><<<(unsigned)blocks, warps * 32, 0, st>>>((const T*)x, rows, d, dp, xb, scal));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

// E_k = Xmax*|c_k - bf16(c_k)| + DXmax*|c_k| + accumulation slack;  bias_k = |c_k|^2/2 - E_k
__global__ void make_bias_kernel(const float* __restrict__ cn2h, const float* __restrict__ cn,
                                 const float* __restrict__ dcn, int64_t total, int Kp, int K,
                                 const uint32_t* __restrict__ scal, float* __restrict__ bias,
                                 float* __restrict__ err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int k = (int)(i % Kp);
  if (k >= K) { bias[i] = kPadBias; err[i] = 0.f; return; }
  float xmax = __uint_as_float(scal[0]), dxmax = __uint_as_float(scal[1]);
  float c = cn[i], h = cn2h[i];
  // rounding of products is exact in the tensor core (bf16 x bf16 fits fp32); accumulation is fp32-ish:
  // allow 2^-16 of the largest possible |sum| plus the fp32 rounding of |c|^2/2.
  float e = xmax * dcn[i] + dxmax * c + 1.6e-5f * (xmax * c) + 2.4e-7f * h;
  e = e * 1.001f + 1e-30f;
  err[i] = e;
  bias[i] = h - e;
}

int launch_make_bias(const void* cache, const CacheLayout& CL, int64_t H, int K, int metric,
                     const uint32_t* scal, float* bias, float* err, cudaStream_t st) {
  (void)metric;
  const char* base = (const char*)cache;
  int64_t total = H * (int64_t)CL.Kp;
  make_bias_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      (const float*)(base + CL.off_cn2h), (const float*)(base + CL.off_cn), (const float*)(base + CL.off_dcn),
      total, CL.Kp, K, scal, bias, err);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

// reference utils/losses.py:19: x / max(|x|_2, 1e-12); one warp per row
template <typename T>
__global__ void l2norm_rows_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t rows, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * (int64_t)d;
  float* o = out + row * (int64_t)d;
  float s = 0.f;
  if ((d & 3) == 0) {
    for (int j = lane * 4; j < d; j += 128) {
      float4 v = load4<T>(xr + j);
      s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
  } else {
    for (int j = lane; j < d; j += 32) { float v = to_f32<T>(xr[j]); s = fmaf(v, v, s); }
  }
  s = warp_sum(s);
  float nrm = fmaxf(sqrtf(s), 1e-12f);
  if ((d & 3) == 0) {
    for (int j = lane * 4; j < d; j += 128) {
      float4 v = load4<T>(xr + j);
      float4 r = make_float4(__fdiv_rn(v.x, nrm), __fdiv_rn(v.y, nrm), __fdiv_rn(v.z, nrm), __fdiv_rn(v.w, nrm));
      *reinterpret_cast<float4*>(o + j) = r;
    }
  } else {
    for (int j = lane; j < d; j += 32) o[j] = __fdiv_rn(to_f32<T>(xr[j]), nrm);
  }
}

}  // namespace vqb

using namespace vqb;

extern "C" size_t vqb_codebook_cache_bytes(int64_t H, int K, int d) {
  if (H <= 0 || K <= 0 || d <= 0) return 0;
  return cache_layout(H, K, d).total;
}

extern "C" int vqb_prepare_codebook(const float* codebook, int64_t H, int K, int d, int metric,
                                    void* cache, size_t cache_bytes, void* stream) {
  VQB_REQUIRE(codebook && cache && H > 0 && K > 0 && d > 0, VQB_ERR_INVALID, "vqb_prepare_codebook: bad argument");
  VQB_REQUIRE(metric == VQB_EUCLID || metric == VQB_DOT, VQB_ERR_INVALID, "unknown metric %d", metric);
  CacheLayout CL = cache_layout(H, K, d);
  VQB_REQUIRE(cache_bytes >= CL.total, VQB_ERR_WORKSPACE, "codebook cache too small: %zu < %zu", cache_bytes, CL.total);
  char* base = (char*)cache;
  const int warps = 8;
  int64_t rows = H * (int64_t)CL.Kp;
  prepare_codebook_kernel<<<(unsigned)((rows + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(
      codebook, H, K, CL.Kp, d, CL.dp, metric, (__nv_bfloat16*)(base + CL.off_cb), (float*)(base + CL.off_cn2h),
      (float*)(base + CL.off_cn), (float*)(base + CL.off_dcn));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

extern "C" int vqb_l2norm_rows(const void* x, int x_dtype, float* out, int64_t rows, int d, void* stream) {
  VQB_REQUIRE(x && out && rows >= 0 && d > 0, VQB_ERR_INVALID, "vqb_l2norm_rows: bad argument");
  if (rows == 0) return VQB_OK;
  const int warps = 8;
  VQB_DISPATCH_DTYPE(x_dtype, T,
    l2norm_rows_kernel<T><<<(unsigned)((rows + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(
        (const T*)x, out, rows, d));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}
