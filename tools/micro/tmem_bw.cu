// tmem_bw.cu -- microbenchmark: tcgen05.ld latency and throughput on sm_100a (evidence for DESIGN.md section 4.1).
// One CTA per SM allocates all 512 TMEM columns; W warps (W = 4, 8, 16) read their lane quadrant with
// 32x32b.xC loads, D loads in flight per tcgen05.wait::ld.  Reports cycles per load round and bytes/clk/SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu ; run: ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int C> struct Ld;
template <> struct Ld<8> {
  static __device__ __forceinline__ void go(uint32_t addr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(addr));
  }
};
template <> struct Ld<16> {
  static __device__ __forceinline__ void go(uint32_t addr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr));
  }
};
template <> struct Ld<32> {
  static __device__ __forceinline__ void go(uint32_t addr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(addr));
  }
};

// C columns per load, D loads in flight per wait
template <int C, int D>
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s;
  const int q = warp & 3;
  const int nw = blockDim.x >> 5;
  const int cols_per_warp = 512 / (nw / 4);       // each quadrant group of warps splits the 512 columns
  const uint32_t lane_addr = base + ((uint32_t)(q * 32) << 16) + (uint32_t)((warp >> 2) * cols_per_warp);
  uint32_t acc = 0;
  uint32_t r[D][C];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)((it * D * C) % cols_per_warp);
#pragma unroll
    for (int j = 0; j < D; ++j) Ld<C>::go(lane_addr + ((col + j * C) % cols_per_warp), r[j]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
      for (int i = 0; i < C; i += 8) acc ^= r[j][i];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

template <int C, int D>
static void run(int warps, int grid) {
  unsigned long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, grid * 8);
  cudaMalloc(&sink, 4);
  const int iters = 4096;
  tmem_read_kernel<C, D><<<grid, warps * 32>>>(16, cyc, sink);
  tmem_read_kernel<C, D><<<grid, warps * 32>>>(iters, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("C=%d D=%d warps=%d: %s\n", C, D, warps, cudaGetErrorString(e)); return; }
  unsigned long long h[256];
  cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += (double)h[i];
  mean /= grid;
  const double per_round = mean / iters;
  const double bytes = (double)warps * 32 * C * D * 4;        // per CTA per round
  printf("{\"cols_per_ld\": %d, \"lds_in_flight\": %d, \"warps\": %d, \"grid\": %d, \"cycles_per_round\": %.1f, "
         "\"bytes_per_clk_per_sm\": %.1f, \"cycles_per_128x256_f32_tile\": %.0f}\n",
         C, D, warps, grid, per_round, bytes / per_round, 131072.0 / (bytes / per_round));
  cudaFree(cyc);
  cudaFree(sink);
}


// ---- reads under tensor-core load: warps 0..15 read accumulator columns [0,256) while warp 16 keeps
// tcgen05.mma (M=128, N=256, K=16, fp16 -> fp32, garbage operands) accumulating into columns [256,512).
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
template <int C, bool SAME_COLS>
__global__ void __launch_bounds__(544, 1) tmem_read_under_mma_kernel(int iters, int mma_on, unsigned long long* out,
                                                                     uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];     // 16 KB A + 32 KB B, contents irrelevant
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t bar[2];
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    done = 0;
    for (int i = 0; i < 2; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s;
  if (warp < 16) {
    const int q = warp & 3;
    const uint32_t lane_addr = base + ((uint32_t)(q * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    uint32_t acc = 0;
    uint32_t r[C];
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      Ld<C>::go(lane_addr + (uint32_t)((it * C) % 64), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < C; i += 8) acc ^= r[i];
    }
    const long long t1 = clock64();
    if (lane == 0) atomicAdd((int*)&done, 1);
    if (threadIdx.x == 0) out[blockIdx.x * 4 + 0] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
  } else if (mma_on) {
    // MMA warp: batches of 8 MMAs, two barriers so that a batch is always queued behind the running one
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem), b = a + 16384;
    const uint32_t idesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t d_tmem = base + (SAME_COLS ? 0u : 256u);
    unsigned long long n = 0;
    uint32_t ph[2] = {0, 0};
    const long long t0 = clock64();
    int i = 0;
    while (done < 16) {
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t ad = sw128_desc(a) + (uint64_t)((k & 3) * 2), bd = sw128_desc(b) + (uint64_t)((k & 3) * 2);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d_tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(&bar[i & 1])) : "memory");
      }
      __syncwarp();
      n += 8;
      if (i > 0) {   // wait for the PREVIOUS batch
        const int j = (i - 1) & 1;
        uint32_t ok = 0;
        while (!ok) {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                       "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok)
                       : "r"((uint32_t)__cvta_generic_to_shared(&bar[j])), "r"(ph[j]) : "memory");
        }
        ph[j] ^= 1u;
      }
      ++i;
    }
    {   // drain the last batch
      const int j = (i - 1) & 1;
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok)
                     : "r"((uint32_t)__cvta_generic_to_shared(&bar[j])), "r"(ph[j]) : "memory");
      }
    }
    const long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x * 4 + 1] = (unsigned long long)(t1 - t0); out[blockIdx.x * 4 + 2] = n; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

template <int C, bool SAME>
static void run_mma(int mma_on, int grid) {
  unsigned long long* out;
  uint32_t* sink;
  cudaMalloc(&out, grid * 32);
  cudaMemset(out, 0, grid * 32);
  cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(tmem_read_under_mma_kernel<C, SAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024 + 1024);
  const int iters = 8192;
  tmem_read_under_mma_kernel<C, SAME><<<grid, 544, 48 * 1024 + 1024>>>(64, mma_on, out, sink);
  tmem_read_under_mma_kernel<C, SAME><<<grid, 544, 48 * 1024 + 1024>>>(iters, mma_on, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("under-mma C=%d: %s\n", C, cudaGetErrorString(e)); return; }
  unsigned long long h[4 * 256];
  cudaMemcpy(h, out, grid * 32, cudaMemcpyDeviceToHost);
  double rd = 0, mm = 0, n = 0;
  for (int i = 0; i < grid; ++i) { rd += (double)h[i * 4]; mm += (double)h[i * 4 + 1]; n += (double)h[i * 4 + 2]; }
  rd /= grid; mm /= grid; n /= grid;
  const double bytes = 16.0 * 32 * C * 4;
  printf("{\"test\": \"read_under_mma\", \"mma_on\": %d, \"mma_same_columns\": %d, \"cols_per_ld\": %d, \"warps\": 16, "
         "\"cycles_per_round\": %.1f, \"bytes_per_clk_per_sm\": %.1f, \"cycles_per_128x256_f32_tile\": %.0f, "
         "\"mma_cycles_each\": %.1f}\n",
         mma_on, (int)SAME, C, rd / iters, bytes / (rd / iters), 131072.0 / (bytes / (rd / iters)), n > 0 ? mm / n : 0.0);
  cudaFree(out);
  cudaFree(sink);
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  // latency: one warp, one load per wait
  run<8, 1>(4, 1);  run<16, 1>(4, 1);  run<32, 1>(4, 1);
  for (int w : {4, 8, 16}) {
    run<8, 1>(w, sms);  run<16, 1>(w, sms);  run<32, 1>(w, sms);
    run<8, 2>(w, sms);  run<16, 2>(w, sms);  run<32, 2>(w, sms);
    run<16, 4>(w, sms); run<32, 4>(w, sms);
  }
  run_mma<16, false>(0, sms); run_mma<16, false>(1, sms); run_mma<32, false>(0, sms); run_mma<32, false>(1, sms);
  run_mma<16, true>(1, sms);
  return 0;
}
