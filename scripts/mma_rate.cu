// Micro-benchmark: issue rate of tcgen05.mma kind::tf32 (M = 128, K = 8) as a function of N, operands resident in shared
// memory (SWIZZLE_128B K-major tiles, contents irrelevant), accumulator in TMEM.  One CTA per SM, one issuing thread,
// `iters` back-to-back MMAs, one commit; cycles from clock64 around issue + completion.  Answers "what does a narrow-N
// MMA cost" (DESIGN.md section 4).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/mma_rate
// scripts/mma_rate.cu -I centerpose_b200/csrc ; run on the GPU box.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "umma_common.cuh"

using namespace cp::umma;

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, int same_a, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t a0 = base;                 // 4 A tiles of 128 rows x 128 B = 64 KB
  const uint32_t b0 = base + 65536u;        // 4 B tiles of up to 256 rows x 128 B = 128 KB
  for (uint32_t i = threadIdx.x; i < (65536u + 131072u) / 16u; i += blockDim.x) st_shared_v4(base + i * 16u, 0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&tmem_base_s), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_tf32(N);
    const uint64_t dt = make_desc(0, 0, 32);
    const uint32_t dhi = (uint32_t)(dt >> 32), dlo = (uint32_t)dt;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t ta = same_a ? 0u : (uint32_t)(it & 3);
      const uint32_t da = dlo + ((a0 + ta * 16384u) >> 4), db = dlo + ((b0 + (uint32_t)(it & 3) * 32768u) >> 4);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_tf32_lohi(tmem, da + 2u * ks, db + 2u * ks, dhi, idesc, (it | ks) ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0u);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long* out;
  cudaMalloc(&out, sizeof(long long) * sms);
  const size_t smem = 65536 + 131072 + 2048;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2000;
  printf("tcgen05.mma kind::tf32, M=128, K=8, SWIZZLE_128B operands in shared memory; %d SMs, %d MMAs per CTA\n", sms, iters * 4);
  printf("%6s %12s %12s %16s %22s\n", "N", "clk/MMA", "floor N/2", "smem 32+N/4", "clk/MMA one SM only");
  for (int N : {16, 32, 64, 96, 128, 192, 256}) {
    double res[2];
    for (int mode = 0; mode < 2; ++mode) {
      const int grid = mode == 0 ? sms : 1;
      mma_rate_kernel<<<grid, 128, smem>>>(N, iters, 0, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("N=%d: %s\n", N, cudaGetErrorString(e));
        return 1;
      }
      long long h[256];
      cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      res[mode] = (double)mx / (iters * 4.0);
    }
    printf("%6d %12.1f %12.1f %16.1f %22.1f\n", N, res[0], N / 2.0, 32.0 + N / 4.0, res[1]);
  }
  return 0;
}
