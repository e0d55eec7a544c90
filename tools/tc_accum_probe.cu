// Micro-experiment (not part of the library): how much of the error of the fp16 hi/lo split ("fp16x3") tensor-core path
// comes from tcgen05's fp32 ACCUMULATION, as a function of how many K-steps are chained into one TMEM accumulator.
// "accumulate at most one k-block per TMEM slot and promote the partial sums into fp32 registers
// (round-to-nearest adds) in the epilogue ... measure error vs K-chunk first".
//
// One CTA computes D[128 x 64] = A[128 x K] . B[64 x K]^T for conv-like K (288 ... 2304) from fp32 inputs split into fp16
// hi + lo (a = a_hi + a_lo to ~22 bits), three products per k-step (hi.hi, hi.lo, lo.hi) as the library does, and
//   chain      all K-steps of a product family chained in TMEM (what tc_*.cu do today: hi.hi + lo.hi in one accumulator,
//              hi.lo in a second one, summed once in the epilogue)
//   chunk KC   the accumulator is read back every KC values of K and added into fp32 registers (add.rn), then restarted
// against an fp64 product of the ORIGINAL fp32 inputs.  Also printed: the error of the split itself (fp64 product of the
// split operands) and of a sequential fp32 FMA chain (what the SIMT fp32 tier does).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I debvader_b200/csrc -I include tools/tc_accum_probe.cu -o tools/build/tc_accum_probe
#include "tc_ptx.cuh"
#include <cmath>
#include <cstdio>
#include <cuda_fp16.h>
#include <random>
#include <vector>

using namespace dbv;

constexpr int M = 128, N = 64, KSTAGE = 256;            // K values staged in shared memory at a time
constexpr int A_STEP = M * 16 * 2, B_STEP = N * 16 * 2;  // bytes of one k-step operand (16 values of K)
constexpr int STAGE_BYTES = (KSTAGE / 16) * (2 * A_STEP + 2 * B_STEP);  // A_hi, A_lo, B_hi, B_lo

// Shared-memory image of one stage: for every k-step s: [A_hi | A_lo | B_hi | B_lo], each operand as un-swizzled K-major
// core matrices [k half (2)][row group (rows/8)][8 rows][8 values] -> LBO = rows * 16 bytes, SBO = 128 bytes.
__global__ void __launch_bounds__(128, 1) accum_probe_kernel(const uint4* __restrict__ img, int K, int chunk, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sBar = base + STAGE_BYTES, s_tmem = sBar + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + STAGE_BYTES + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(sBar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(s_tmem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // fp16 x fp16 -> fp32
  const uint32_t hi = (128u >> 4) | (1u << 14);                                               // SBO = 128 B, version 1, no swizzle
  const uint32_t loA = ((uint32_t)(M * 16) >> 4) << 16, loB = ((uint32_t)(N * 16) >> 4) << 16;  // LBO
  float sum[N];  // promoted partial sums of this thread's row
  float fin[N];
#pragma unroll
  for (int j = 0; j < N; ++j) sum[j] = 0.f;
  uint32_t phase = 0;
  bool fresh = true;  // next MMA of each family starts a new accumulation
  int since = 0;      // K values accumulated since the last promotion
  auto promote = [&](bool last) {
    // cols [0,64): hi.hi + lo.hi   cols [64,128): hi.lo
    float v[32], w[32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      tmem_ld_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32 * h, v);
      tmem_ld_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 64 + 32 * h, w);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (last && chunk <= 0) fin[32 * h + j] = __fadd_rn(v[j], w[j]);  // the library's epilogue: one sum of the two accumulators
        else sum[32 * h + j] = __fadd_rn(sum[32 * h + j], __fadd_rn(v[j], w[j]));
      }
    }
    tc_fence_before();
  };
  for (int k0 = 0; k0 < K; k0 += KSTAGE) {
    const int kn = (K - k0 < KSTAGE) ? (K - k0) : KSTAGE;
    const int nvec = (kn / 16) * (2 * A_STEP + 2 * B_STEP) / 16;
    const uint4* src = img + (size_t)(k0 / 16) * ((2 * A_STEP + 2 * B_STEP) / 16);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) reinterpret_cast<uint4*>(gen)[i] = src[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    bool pending = false;  // MMAs issued and not yet waited for (uniform across the CTA)
    for (int s = 0; s < kn / 16; ++s) {
      if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t a_hi = base + s * (2 * A_STEP + 2 * B_STEP), a_lo = a_hi + A_STEP, b_hi = a_lo + A_STEP, b_lo = b_hi + B_STEP;
        umma_f16(tmem_base, desc64(hi, loA | (a_hi >> 4)), desc64(hi, loB | (b_hi >> 4)), idesc, fresh ? 0u : 1u);       // hi.hi
        umma_f16(tmem_base, desc64(hi, loA | (a_lo >> 4)), desc64(hi, loB | (b_hi >> 4)), idesc, 1u);                    // + lo.hi
        umma_f16(tmem_base + 64, desc64(hi, loA | (a_hi >> 4)), desc64(hi, loB | (b_lo >> 4)), idesc, fresh ? 0u : 1u);  // hi.lo
      }
      fresh = false;
      pending = true;
      since += 16;
      const bool end = (k0 + 16 * (s + 1) == K);
      if ((chunk > 0 && since >= chunk) || end) {
        if (threadIdx.x == 0) umma_commit(sBar);
        mbar_wait(sBar, phase);
        phase ^= 1u;
        tc_fence_after();
        promote(end);
        __syncthreads();
        fresh = true;
        since = 0;
        pending = false;
      }
    }
    if (pending) {  // the stage buffer is about to be overwritten: its MMAs must have read it
      if (threadIdx.x == 0) umma_commit(sBar);
      mbar_wait(sBar, phase);
      phase ^= 1u;
      tc_fence_after();
    }
    __syncthreads();
  }
  const int row = warp * 32 + lane;
#pragma unroll
  for (int j = 0; j < N; ++j) out[row * N + j] = (chunk <= 0) ? fin[j] : sum[j];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static void split16(float v, __half& h, __half& l) {
  h = __float2half_rn(v);
  l = __float2half_rn(v - __half2float(h));
}

int main() {
  std::mt19937_64 rng(7);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::uniform_real_distribution<float> ud(-1.f, 1.f);
  cudaFuncSetAttribute(accum_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE_BYTES + 2048);
  printf("K     variant        max|err|/max|D|   rms err/max|D|\n");
  for (int K : {288, 576, 1152, 2304}) {
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    for (auto& a : A) { float t = nd(rng); a = t > 0 ? t : 0.15f * t; }  // post-PReLU activations: mostly positive
    const float lim = 1.6f * std::sqrt(6.f / (float)(K + 9 * N));
    for (auto& b : B) b = lim * ud(rng);
    std::vector<double> ref((size_t)M * N), refsplit((size_t)M * N);
    std::vector<float> fma32((size_t)M * N);
    std::vector<__half> Ah(A.size()), Al(A.size()), Bh(B.size()), Bl(B.size());
    for (size_t i = 0; i < A.size(); ++i) split16(A[i], Ah[i], Al[i]);
    for (size_t i = 0; i < B.size(); ++i) split16(B[i], Bh[i], Bl[i]);
    double dmax = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0, t = 0;
        float f = 0.f;
        for (int k = 0; k < K; ++k) {
          s += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
          const double ah = __half2float(Ah[(size_t)m * K + k]), al = __half2float(Al[(size_t)m * K + k]);
          const double bh = __half2float(Bh[(size_t)n * K + k]), bl = __half2float(Bl[(size_t)n * K + k]);
          t += ah * bh + ah * bl + al * bh;
          f = std::fmaf(A[(size_t)m * K + k], B[(size_t)n * K + k], f);
        }
        ref[(size_t)m * N + n] = s;
        refsplit[(size_t)m * N + n] = t;
        fma32[(size_t)m * N + n] = f;
        dmax = std::fmax(dmax, std::fabs(s));
      }
    auto report = [&](const char* name, auto get) {
      double mx = 0, ss = 0;
      for (size_t i = 0; i < ref.size(); ++i) {
        const double e = (double)get(i) - ref[i];
        mx = std::fmax(mx, std::fabs(e));
        ss += e * e;
      }
      printf("%-5d %-14s %.3e         %.3e\n", K, name, mx / dmax, std::sqrt(ss / ref.size()) / dmax);
    };
    report("split (fp64)", [&](size_t i) { return refsplit[i]; });
    report("fp32 FMA chain", [&](size_t i) { return (double)fma32[i]; });
    // shared-memory image
    const int steps = K / 16;
    std::vector<__half> img((size_t)steps * (2 * A_STEP + 2 * B_STEP) / 2);
    auto put = [&](size_t byte_off, const std::vector<__half>& src, int rows, int s) {
      for (int kh = 0; kh < 2; ++kh)
        for (int r = 0; r < rows; ++r)
          for (int e = 0; e < 8; ++e)
            img[byte_off / 2 + ((size_t)kh * (rows / 8) + r / 8) * 64 + (r % 8) * 8 + e] = src[(size_t)r * K + 16 * s + 8 * kh + e];
    };
    for (int s = 0; s < steps; ++s) {
      const size_t o = (size_t)s * (2 * A_STEP + 2 * B_STEP);
      put(o, Ah, M, s);
      put(o + A_STEP, Al, M, s);
      put(o + 2 * A_STEP, Bh, N, s);
      put(o + 2 * A_STEP + B_STEP, Bl, N, s);
    }
    uint4* dimg;
    float* dout;
    cudaMalloc(&dimg, img.size() * 2);
    cudaMalloc(&dout, (size_t)M * N * 4);
    cudaMemcpy(dimg, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
    std::vector<float> got((size_t)M * N);
    for (int chunk : {0, 1024, 512, 256, 128, 64, 32, 16}) {
      if (chunk > K) continue;
      cudaMemset(dout, 0, (size_t)M * N * 4);
      accum_probe_kernel<<<1, 128, STAGE_BYTES + 2048>>>(dimg, K, chunk, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
      char name[32];
      if (chunk) snprintf(name, sizeof name, "chunk %d", chunk);
      else snprintf(name, sizeof name, "chain (today)");
      report(name, [&](size_t i) { return (double)got[i]; });
    }
    cudaFree(dimg);
    cudaFree(dout);
  }
  return 0;
}
