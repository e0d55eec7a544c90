// Micro-benchmark (not part of the library): cycles per tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, K = 16, kind::f16, SS
// mode, 64-byte swizzled K-major rows) as a function of N, next to the single-CTA figures of tools/mma_rate.cu
// (max(N/2, 32 + N/4) cycles).  Behind the decision to run only convT6 (N = 128) on the CTA-pair halo kernel (csrc/tc_halo2.cu).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I debvader_b200/csrc -I include tools/mma2_rate.cu -o tools/build/mma2_rate
#include "tc_ptx.cuh"
#include "tc_pair_ptx.cuh"
#include <cstdio>

using namespace dbv;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma2_rate_kernel(int N, int nmma, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 64 * 1024, sBar = base + 128 * 1024, s_tmem = sBar + 64;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (s_tmem - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  for (uint32_t i = threadIdx.x * 16; i < 128 * 1024; i += blockDim.x * 16) asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i), "r"(0));
  if (threadIdx.x == 0) {
    mbar_init(sBar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(s_tmem, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1 && rank == 0) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);
      constexpr uint32_t HI = smem_desc_hi<64>();
      uint32_t phase = 0;
      for (int rep = 0; rep < 3; ++rep) {
        const long long t0 = clock64();
        for (int i = 0; i < nmma; ++i)
          umma2_f16(tmem_base, desc64(HI, kSmemDescLoConst | ((sA + (i & 7) * 8192) >> 4)), desc64(HI, kSmemDescLoConst | (sB >> 4)), idesc, 1u);
        umma2_commit_mc(sBar);
        mbar_wait_cluster(sBar, phase);
        phase ^= 1u;
        if (rep == 2 && blockIdx.x == 0) out[0] = clock64() - t0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(mma2_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  printf("tcgen05.mma.cta_group::2 kind::f16, M = 256 (two CTAs x 128 rows), K = 16, each CTA supplies N/2 rows of B\n");
  printf("  N   clusters  cycles/MMA   MAC/cycle/SM   (single CTA, M = 128: max(N/2, 32 + N/4) cycles)\n");
  for (int clusters : {1, 74})
    for (int N : {32, 64, 128, 256}) {
      cudaMemset(d, 0, 64);
      mma2_rate_kernel<<<2 * clusters, 128, 140 * 1024>>>(N, 2048, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h;
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double c = (double)h / 2048;
      printf("%4d   %3d      %8.1f      %8.0f        (%d)\n", N, clusters, c, 128.0 * N * 16 / c, N / 2 > 32 + N / 4 ? N / 2 : 32 + N / 4);
    }
  return 0;
}
