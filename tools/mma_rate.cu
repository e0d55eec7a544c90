// Micro-benchmark (not part of the library): cycles per tcgen05.mma (M=128, K=16, kind::f16, SS mode) as a
// function of N and of the shared-memory operand layout, alone and with 8 warps issuing tcgen05.ld on
// other TMEM columns.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I debvader_b200/csrc
//   -I include tools/mma_rate.cu -o tools/build/mma_rate     Run on the GPU box: tools/build/mma_rate
#include "tc_ptx.cuh"
#include <cstdio>
#include <vector>

using namespace dbv;

struct Args {
  int N;        // MMA N
  int rowb;     // 128 / 64 / 32 (swizzled K-major rows) or 16 (no swizzle, core matrices: SBO 128, LBO 16 for A)
  int nmma;     // MMAs per measurement
  int a_step;   // bytes added to the A start address between successive MMAs (0: same operand; else distinct tiles/taps)
  int ld_warps; // 0 or 8: epilogue-like warps hammering tcgen05.ld meanwhile
  int ctas;
};

__global__ void __launch_bounds__(320, 1) mma_rate_kernel(Args a, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 128 * 1024, sBar = base + 200 * 1024, s_tmem = sBar + 64;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (s_tmem - smem_u32(smem_raw)));
  volatile int* stop = reinterpret_cast<volatile int*>(smem_raw + (sBar + 128 - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // zero the operands (NaN patterns would not change timing, but keep it clean)
  for (uint32_t i = threadIdx.x * 16; i < 200 * 1024; i += blockDim.x * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i), "r"(0));
  if (threadIdx.x == 0) {
    mbar_init(sBar, 1);
    *stop = 0;
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.N >> 3) << 17) | ((128u >> 4) << 24);
      uint32_t hiA, hiB, loA, loB;
      if (a.rowb == 16) {
        hiA = (128u >> 4) | (1u << 14); hiB = (256u >> 4) | (1u << 14);
        loA = kSmemDescLoConst; loB = ((128u >> 4) << 16);
      } else {
        const uint32_t lay = a.rowb == 128 ? 2u : (a.rowb == 64 ? 4u : 6u);
        hiA = hiB = (uint32_t)((8 * a.rowb) >> 4) | (1u << 14) | (lay << 29);
        loA = loB = kSmemDescLoConst;
      }
      uint32_t phase = 0;
      for (int rep = 0; rep < 3; ++rep) {
        const long long t0 = clock64();
        uint32_t alo = loA | (sA >> 4);
        const uint32_t blo = loB | (sB >> 4);
        for (int i = 0; i < a.nmma; ++i) {
          umma_f16(tmem_base, desc64(hiA, alo), desc64(hiB, blo), idesc, 1u);
          alo += (uint32_t)(a.a_step >> 4);
          if ((i & 7) == 7) alo = loA | (sA >> 4);
        }
        umma_commit(sBar);
        mbar_wait(sBar, phase);
        phase ^= 1u;
        const long long t1 = clock64();
        if (rep == 2 && blockIdx.x == 0) out[0] = t1 - t0;
      }
      *stop = 1;
    }
  } else if (warp >= 2 && warp < 2 + a.ld_warps) {
    const int quad = warp & 3;
    float acc = 0.f;
    long long n = 0;
    const long long t0 = clock64();
    while (!*stop) {
      float v[32];
      tmem_ld_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + 256 + 32 * ((warp >> 2) & 1), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += v[j];
      ++n;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && warp == 2 && lane == 0) { out[1] = n; out[2] = t1 - t0; }
    if (acc == 123.456f) out[3] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int Ns[] = {16, 32, 64, 128, 256};
  const int rowbs[] = {128, 64, 16};
  printf("rowb  N  a_step ld_warps ctas  cycles/MMA   (tcgen05.ld x32 per 1k cycles per warp)\n");
  for (int ctas : {1, 148})
    for (int ldw : {0, 8})
      for (int rowb : rowbs)
        for (int N : Ns)
          for (int a_step : {0, rowb * 128}) {
            if (rowb == 16 && N > 64) continue;
            Args a{N, rowb, 2048, a_step, ldw, ctas};
            cudaMemset(d, 0, 64);
            mma_rate_kernel<<<ctas, 320, 227 * 1024>>>(a, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            long long h[4];
            cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
            printf("%4d %3d %6d %d %3d   %8.1f   %s", rowb, N, a_step, ldw, ctas, (double)h[0] / a.nmma, "");
            if (ldw) printf("%.2f", h[2] ? 1000.0 * h[1] / h[2] : 0.0);
            printf("\n");
          }
  return 0;
}
