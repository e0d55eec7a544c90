// Micro-experiment (not part of the library): tcgen05.mma.kind::tf32 on the conv-like GEMM of tools/tc_accum_probe.cu.
// north_star names "TF32/bf16 path <= 1e-3"; VERDICT r1 weak #6: "no kind::tf32 measurement exists anywhere".
//   rate      cycles per tcgen05.mma (M = 128, K = 8 tf32 values = the same 32 bytes per operand row as K = 16 of fp16) vs N
//   accuracy  D[128 x 64] = A[128 x K] . B[64 x K]^T, fp32 data, against an fp64 product:
//               tf32      operands rounded to tf32 (10-bit mantissa, cvt.rna), one MMA per product
//               tf32x3    a = a_hi + a_lo (two tf32 values, ~21 bits): hi.hi + lo.hi chained in one accumulator, hi.lo in a second
//             next to the fp16x3 figures of tc_accum_probe (same data generator, same seed).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I debvader_b200/csrc -I include tools/tf32_probe.cu -o tools/build/tf32_probe
#include "tc_ptx.cuh"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

using namespace dbv;

constexpr int M = 128, N = 64, KSTAGE = 128;            // K values staged in shared memory at a time
constexpr int A_STEP = M * 8 * 4, B_STEP = N * 8 * 4;    // bytes of one k-step operand (8 tf32 values of K)
constexpr int STAGE_BYTES = (KSTAGE / 8) * (2 * A_STEP + 2 * B_STEP);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

// mode 0: single pass (hi planes only), 1: x3.  Shared-memory image per k-step: [A_hi | A_lo | B_hi | B_lo], each operand as
// un-swizzled K-major core matrices [k half (2)][row group][8 rows][4 values] -> LBO = rows * 16 bytes, SBO = 128 bytes.
__global__ void __launch_bounds__(128, 1) tf32_probe_kernel(const uint4* __restrict__ img, int K, int mode, float* __restrict__ out,
                                                             long long* __restrict__ cycles, int rate_n, int rate_count) {
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
  if (warp == 0) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t hi = (128u >> 4) | (1u << 14);
  const uint32_t loA = ((uint32_t)(M * 16) >> 4) << 16, loB = ((uint32_t)(N * 16) >> 4) << 16;
  uint32_t phase = 0;
  if (rate_count > 0) {  // ---- rate: rate_count MMAs of N = rate_n on the (zeroed) stage buffer --------------------------------
    for (int i = threadIdx.x; i < STAGE_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(gen)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(rate_n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      const uint32_t loBn = ((uint32_t)(rate_n * 16) >> 4) << 16;
      for (int rep = 0; rep < 3; ++rep) {
        const long long t0 = clock64();
        for (int i = 0; i < rate_count; ++i)
          umma_tf32(tmem_base, desc64(hi, loA | (base >> 4)), desc64(hi, loBn | ((base + 65536) >> 4)), idesc, 1u);
        umma_commit(sBar);
        mbar_wait(sBar, phase);
        phase ^= 1u;
        if (rep == 2) cycles[0] = clock64() - t0;
      }
    }
    __syncthreads();
  } else {  // ---- accuracy ---------------------------------------------------------------------------------------------------------
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // tf32 x tf32 -> fp32
    bool fresh = true;
    for (int k0 = 0; k0 < K; k0 += KSTAGE) {
      const int kn = (K - k0 < KSTAGE) ? (K - k0) : KSTAGE;
      const int nvec = (kn / 8) * (2 * A_STEP + 2 * B_STEP) / 16;
      const uint4* src = img + (size_t)(k0 / 8) * ((2 * A_STEP + 2 * B_STEP) / 16);
      for (int i = threadIdx.x; i < nvec; i += blockDim.x) reinterpret_cast<uint4*>(gen)[i] = src[i];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (threadIdx.x == 0) {
        tc_fence_after();
        for (int s = 0; s < kn / 8; ++s) {
          const uint32_t a_hi = base + s * (2 * A_STEP + 2 * B_STEP), a_lo = a_hi + A_STEP, b_hi = a_lo + A_STEP, b_lo = b_hi + B_STEP;
          umma_tf32(tmem_base, desc64(hi, loA | (a_hi >> 4)), desc64(hi, loB | (b_hi >> 4)), idesc, fresh ? 0u : 1u);
          if (mode == 1) {
            umma_tf32(tmem_base, desc64(hi, loA | (a_lo >> 4)), desc64(hi, loB | (b_hi >> 4)), idesc, 1u);
            umma_tf32(tmem_base + 64, desc64(hi, loA | (a_hi >> 4)), desc64(hi, loB | (b_lo >> 4)), idesc, fresh ? 0u : 1u);
          }
          fresh = false;
        }
        umma_commit(sBar);
      }
      mbar_wait(sBar, phase);
      phase ^= 1u;
      tc_fence_after();
      __syncthreads();
    }
    float v[32], w[32];
    const int row = warp * 32 + lane;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      tmem_ld_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32 * h, v);
      if (mode == 1) tmem_ld_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 64 + 32 * h, w);
#pragma unroll
      for (int j = 0; j < 32; ++j) out[row * N + 32 * h + j] = mode == 1 ? v[j] + w[j] : v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static float to_tf32(float v) {  // cvt.rna.tf32.f32: round to nearest (ties away) to 10 explicit mantissa bits
  uint32_t u;
  memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}

int main() {
  cudaFuncSetAttribute(tf32_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long* dcyc;
  cudaMalloc(&dcyc, 8);
  printf("rate: cycles per tcgen05.mma.kind::tf32, M = 128, K = 8 (32 bytes per operand row, as K = 16 of kind::f16), one CTA\n");
  for (int n : {32, 64, 128, 256}) {
    tf32_probe_kernel<<<1, 128, 200 * 1024>>>(nullptr, 0, 0, nullptr, dcyc, n, 2048);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    long long c;
    cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
    printf("  N = %3d: %.1f cycles per MMA  (%.0f MAC/cycle; kind::f16 at the same N: K = 16 in max(N/2, 32 + N/4) cycles)\n", n, (double)c / 2048,
           128.0 * n * 8 / ((double)c / 2048));
  }
  std::mt19937_64 rng(7);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::uniform_real_distribution<float> ud(-1.f, 1.f);
  printf("K     variant        max|err|/max|D|   rms err/max|D|\n");
  for (int K : {288, 576, 1152, 2304}) {
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    for (auto& a : A) { float t = nd(rng); a = t > 0 ? t : 0.15f * t; }
    const float lim = 1.6f * std::sqrt(6.f / (float)(K + 9 * N));
    for (auto& b : B) b = lim * ud(rng);
    std::vector<double> ref((size_t)M * N);
    double dmax = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
        ref[(size_t)m * N + n] = s;
        dmax = std::fmax(dmax, std::fabs(s));
      }
    std::vector<float> Ah(A.size()), Al(A.size()), Bh(B.size()), Bl(B.size());
    for (size_t i = 0; i < A.size(); ++i) { Ah[i] = to_tf32(A[i]); Al[i] = to_tf32(A[i] - Ah[i]); }
    for (size_t i = 0; i < B.size(); ++i) { Bh[i] = to_tf32(B[i]); Bl[i] = to_tf32(B[i] - Bh[i]); }
    const int steps = K / 8;
    std::vector<float> img((size_t)steps * (2 * A_STEP + 2 * B_STEP) / 4);
    auto put = [&](size_t byte_off, const std::vector<float>& src, int rows, int s) {
      for (int kh = 0; kh < 2; ++kh)
        for (int r = 0; r < rows; ++r)
          for (int e = 0; e < 4; ++e)
            img[byte_off / 4 + ((size_t)kh * (rows / 8) + r / 8) * 32 + (r % 8) * 4 + e] = src[(size_t)r * K + 8 * s + 4 * kh + e];
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
    cudaMalloc(&dimg, img.size() * 4);
    cudaMalloc(&dout, (size_t)M * N * 4);
    cudaMemcpy(dimg, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
    std::vector<float> got((size_t)M * N);
    for (int mode : {0, 1}) {
      cudaMemset(dout, 0, (size_t)M * N * 4);
      tf32_probe_kernel<<<1, 128, 200 * 1024>>>(dimg, K, mode, dout, dcyc, 0, 0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
      double mx = 0, ss = 0;
      for (size_t i = 0; i < ref.size(); ++i) {
        const double er = (double)got[i] - ref[i];
        mx = std::fmax(mx, std::fabs(er));
        ss += er * er;
      }
      printf("%-5d %-14s %.3e         %.3e\n", K, mode ? "tf32x3" : "tf32", mx / dmax, std::sqrt(ss / ref.size()) / dmax);
    }
    cudaFree(dimg);
    cudaFree(dout);
  }
  return 0;
}
