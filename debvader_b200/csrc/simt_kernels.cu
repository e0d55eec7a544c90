// fp32 SIMT kernels.
//  * simt_conv_kernel: one gather-form convolution kernel that runs every layer of the network in
//    fp32 (DBV_PREC_FP32: the <=1e-5 parity tier and the on-device cross-check of the tensor-core
//    path), and, in the tensor-core modes, the two layers that are not GEMM-shaped enough for
//    tcgen05: encoder conv1 (Cin=6, with the BatchNorm fused on operand load because TF applies
//    SAME zero padding *after* BN — model/model.py:79-82) and decoder Dense(32->560)
//    (model/model.py:114).
//  * latent_kernel: MultivariateNormalTriL / MvNormal (model/model.py:43-58, 211-214), one warp
//    per stamp, plus the decoder's leading PReLU (model/model.py:113).
#include "epilogue.cuh"
#include "kernels.h"
#include <curand_kernel.h>
#include <cstdlib>

namespace dbv {

// gather form: out[b,y,x,co] = sum_{t,ci} in[b, iy(t,y), ix(t,x), ci] * w[t][ci][co]
//   mode 0 (Conv2D, and stride-1 Conv2DTranspose with a flipped kernel): iy = stride*y + ky - pb
//   mode 1 (stride-2 Conv2DTranspose, pb=0):  iy = (y - ky)/2 when y-ky is even and >= 0
__global__ void __launch_bounds__(256) simt_conv_kernel(SimtConv p) {
  const int cgs = p.CoutP >> 2;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = (int)(gid % cgs);
  const long long pix = gid / cgs;
  const long long npix = (long long)p.B * p.Hout * p.Wout;
  if (pix >= npix) return;
  const int x = (int)(pix % p.Wout);
  const int y = (int)((pix / p.Wout) % p.Hout);
  const long long b = pix / ((long long)p.Wout * p.Hout);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int c0 = cg * 4;
  for (int ky = 0; ky < p.ksz; ++ky) {
    int iy;
    if (p.mode == 0) {
      iy = p.stride * y + ky - p.pb;
    } else {
      const int t = y - ky;
      if (t < 0 || (t & 1)) continue;
      iy = t >> 1;
    }
    if (iy < 0 || iy >= p.Hin) continue;
    for (int kx = 0; kx < p.ksz; ++kx) {
      int ix;
      if (p.mode == 0) {
        ix = p.stride * x + kx - p.pb;
      } else {
        const int t = x - kx;
        if (t < 0 || (t & 1)) continue;
        ix = t >> 1;
      }
      if (ix < 0 || ix >= p.Win) continue;
      const float* __restrict__ ip = p.in + ((b * p.Hin + iy) * p.Win + ix) * (long long)p.Cin;
      const float* __restrict__ wp = p.w + (long long)(ky * p.ksz + kx) * p.Cin * p.CoutP + c0;
      if ((p.Cin & 3) == 0 && !p.in_scale) {
        for (int ci = 0; ci < p.Cin; ci += 4) {
          const float4 a = *reinterpret_cast<const float4*>(ip + ci);
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + (long long)(ci + 0) * p.CoutP));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + (long long)(ci + 1) * p.CoutP));
          const float4 w2 = __ldg(reinterpret_cast<const float4*>(wp + (long long)(ci + 2) * p.CoutP));
          const float4 w3 = __ldg(reinterpret_cast<const float4*>(wp + (long long)(ci + 3) * p.CoutP));
          acc[0] = fmaf(a.x, w0.x, acc[0]); acc[1] = fmaf(a.x, w0.y, acc[1]); acc[2] = fmaf(a.x, w0.z, acc[2]); acc[3] = fmaf(a.x, w0.w, acc[3]);
          acc[0] = fmaf(a.y, w1.x, acc[0]); acc[1] = fmaf(a.y, w1.y, acc[1]); acc[2] = fmaf(a.y, w1.z, acc[2]); acc[3] = fmaf(a.y, w1.w, acc[3]);
          acc[0] = fmaf(a.z, w2.x, acc[0]); acc[1] = fmaf(a.z, w2.y, acc[1]); acc[2] = fmaf(a.z, w2.z, acc[2]); acc[3] = fmaf(a.z, w2.w, acc[3]);
          acc[0] = fmaf(a.w, w3.x, acc[0]); acc[1] = fmaf(a.w, w3.y, acc[1]); acc[2] = fmaf(a.w, w3.z, acc[2]); acc[3] = fmaf(a.w, w3.w, acc[3]);
        }
      } else {
        for (int ci = 0; ci < p.Cin; ++ci) {
          float a = ip[ci];
          if (p.in_scale) a = fmaf(a, __ldg(p.in_scale + ci), __ldg(p.in_shift + ci));
          const float4 w = __ldg(reinterpret_cast<const float4*>(wp + (long long)ci * p.CoutP));
          acc[0] = fmaf(a, w.x, acc[0]); acc[1] = fmaf(a, w.y, acc[1]); acc[2] = fmaf(a, w.z, acc[2]); acc[3] = fmaf(a, w.w, acc[3]);
        }
      }
    }
  }
  apply_act<4>(p.o, y, x, c0, acc);
  store_act<4>(p.o, b, y, x, c0, acc);
}

// ---------------------------------------------------------------------------------------------
// simt_tile_kernel: the same gather-form convolution as an implicit GEMM tiled through shared memory
// (the fp32 tier's hot kernel).  A CTA of 256 threads owns BM output pixels x BN output channels; a
// thread owns 8 pixels x 4 channels (32 accumulators as 16 packed pairs: 16 FFMA2 + 3 LDS.128 per k).  K runs over (tap,
// 16-channel chunk) in exactly the order of simt_conv_kernel — ky, kx, ci ascending, one fmaf per
// product, out-of-bounds taps contribute fmaf(0, w, acc) = acc — so the two kernels are bit-identical
// and the naive one stays as the cross-check (DBV_SIMT_TILED=0).  Stride-2 transposed convolutions are
// tiled per output-parity class so that every pixel of a tile uses the same (1, 2 or 4) taps.  The next
// chunk's global loads are issued before the current chunk's FMAs (register prefetch).
// ---------------------------------------------------------------------------------------------
constexpr int ST_TM = 8, ST_TN = 4;

// packed fp32 FMA (sm_100 FFMA2): two independent IEEE fma.rn per instruction — the inner loop is issue bound
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long dup2(float x) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

template <int BN, int ST_KC>
__global__ void __launch_bounds__(256) simt_tile_kernel(SimtConv p, int tiles_m) {
  constexpr int NCG = BN / ST_TN;    // threads along the channels
  constexpr int NPG = 256 / NCG;     // threads along the pixels
  constexpr int BM = NPG * ST_TM;    // pixels per tile: 128 (BN=64), 256 (32), 512 (16)
  constexpr int A_F4 = BM * ST_KC / 4 / 256;
  constexpr int B_TOTAL = ST_KC * BN / 4;             // float4 of the weight tile
  constexpr int B_F4 = (B_TOTAL + 255) / 256;         // per thread
  __shared__ __align__(16) float As[ST_KC][BM];
  __shared__ __align__(16) float Bs[ST_KC][BN];

  const int t = threadIdx.x;
  const int tn = t % NCG, tm = t / NCG;
  const int n0 = blockIdx.y * BN;
  // output-pixel enumeration of this tile: all pixels (mode 0) or one parity class (mode 1)
  int cls_y = 0, cls_x = 0, Hc = p.Hout, Wc = p.Wout;
  long long m0 = (long long)blockIdx.x * BM;
  if (p.mode == 1) {
    const int cls = blockIdx.x / tiles_m;
    cls_y = cls >> 1;
    cls_x = cls & 1;
    Hc = p.Hout >> 1;
    Wc = p.Wout >> 1;
    m0 = (long long)(blockIdx.x - cls * tiles_m) * BM;
  }
  const long long npix = (long long)p.B * Hc * Wc;
  auto decode = [&](long long pp, long long& b, int& y, int& x) -> bool {
    if (pp >= npix) return false;
    const int tx = (int)(pp % Wc);
    const int ty = (int)((pp / Wc) % Hc);
    b = pp / ((long long)Wc * Hc);
    if (p.mode == 1) { y = 2 * ty + cls_y; x = 2 * tx + cls_x; }
    else { y = ty; x = tx; }
    return true;
  };
  // taps of this tile, in simt_conv_kernel's order
  int ntap_y, ntap_x, ky0, kx0, kstep;
  if (p.mode == 1) {
    ky0 = cls_y; kx0 = cls_x; kstep = 2;
    ntap_y = cls_y ? 1 : 2;
    ntap_x = cls_x ? 1 : 2;
  } else {
    ky0 = kx0 = 0; kstep = 1;
    ntap_y = ntap_x = p.ksz;
  }
  const int nchunk = (p.Cin + ST_KC - 1) / ST_KC;
  const int nit = ntap_y * ntap_x * nchunk;

  // the pixels this thread gathers for the A tile (fixed over the K loop)
  long long lb[A_F4];
  int ly[A_F4], lx[A_F4];
#pragma unroll
  for (int l = 0; l < A_F4; ++l) {
    const int idx = t + l * 256;
    long long b = 0;
    int y = 0, x = 0;
    if (!decode(m0 + idx % BM, b, y, x)) y = -(1 << 28);  // never in bounds
    lb[l] = b;
    ly[l] = y;
    lx[l] = x;
  }

  float4 ra[A_F4], rb[B_F4];
  auto fetch = [&](int it) {
    const int tap_i = it / nchunk;
    const int c0 = (it - tap_i * nchunk) * ST_KC;
    const int ky = ky0 + (tap_i / ntap_x) * kstep, kx = kx0 + (tap_i % ntap_x) * kstep;
#pragma unroll
    for (int l = 0; l < A_F4; ++l) {
      const int quad = (t + l * 256) / BM;
      int iy, ix;
      if (p.mode == 0) { iy = p.stride * ly[l] + ky - p.pb; ix = p.stride * lx[l] + kx - p.pb; }
      else { iy = (ly[l] - ky) >> 1; ix = (lx[l] - kx) >> 1; }  // same parity by construction (or far negative)
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      const int ci = c0 + quad * 4;
      if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win && ci < p.Cin) {
        const float* __restrict__ ip = p.in + ((lb[l] * p.Hin + iy) * p.Win + ix) * (long long)p.Cin + ci;
        if ((p.Cin & 3) == 0) {
          a = *reinterpret_cast<const float4*>(ip);
        } else {
          a.x = ip[0];
          if (ci + 1 < p.Cin) a.y = ip[1];
          if (ci + 2 < p.Cin) a.z = ip[2];
          if (ci + 3 < p.Cin) a.w = ip[3];
        }
        if (p.in_scale) {  // BatchNorm on in-bounds pixels only (TF pads after the BN)
          a.x = fmaf(a.x, __ldg(p.in_scale + ci), __ldg(p.in_shift + ci));
          if (ci + 1 < p.Cin) a.y = fmaf(a.y, __ldg(p.in_scale + ci + 1), __ldg(p.in_shift + ci + 1));
          if (ci + 2 < p.Cin) a.z = fmaf(a.z, __ldg(p.in_scale + ci + 2), __ldg(p.in_shift + ci + 2));
          if (ci + 3 < p.Cin) a.w = fmaf(a.w, __ldg(p.in_scale + ci + 3), __ldg(p.in_shift + ci + 3));
        }
      }
      ra[l] = a;
    }
#pragma unroll
    for (int l = 0; l < B_F4; ++l) {
      rb[l] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int idx = t + l * 256;
      if (idx < B_TOTAL) {
        const int k = idx / (BN / 4), n4 = idx % (BN / 4);
        const int ci = c0 + k, n = n0 + n4 * 4;
        if (ci < p.Cin && n < p.CoutP)
          rb[l] = __ldg(reinterpret_cast<const float4*>(p.w + ((long long)(ky * p.ksz + kx) * p.Cin + ci) * p.CoutP + n));
      }
    }
  };

  // accumulators packed two pixels per 64-bit register: acc2[ip][j] = (pixel 2ip, pixel 2ip+1) of channel j
  unsigned long long acc2[ST_TM / 2][ST_TN];
#pragma unroll
  for (int i = 0; i < ST_TM / 2; ++i)
#pragma unroll
    for (int j = 0; j < ST_TN; ++j) acc2[i][j] = 0ull;

  fetch(0);
  for (int it = 0; it < nit; ++it) {
    __syncthreads();  // the previous chunk's FMAs are done with the tiles
#pragma unroll
    for (int l = 0; l < A_F4; ++l) {
      const int idx = t + l * 256;
      const int pl = idx % BM, k = (idx / BM) * 4;
      As[k + 0][pl] = ra[l].x;
      As[k + 1][pl] = ra[l].y;
      As[k + 2][pl] = ra[l].z;
      As[k + 3][pl] = ra[l].w;
    }
#pragma unroll
    for (int l = 0; l < B_F4; ++l) {
      const int idx = t + l * 256;
      if (idx < B_TOTAL) *reinterpret_cast<float4*>(&Bs[idx / (BN / 4)][(idx % (BN / 4)) * 4]) = rb[l];
    }
    __syncthreads();
    if (it + 1 < nit) fetch(it + 1);
#pragma unroll
    for (int k = 0; k < ST_KC; ++k) {
      const ulonglong2 a0 = *reinterpret_cast<const ulonglong2*>(&As[k][tm * ST_TM]);      // pixel pairs (0,1), (2,3)
      const ulonglong2 a1 = *reinterpret_cast<const ulonglong2*>(&As[k][tm * ST_TM + 4]);  // (4,5), (6,7)
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn * ST_TN]);
      const unsigned long long ap[4] = {a0.x, a0.y, a1.x, a1.y};
      const unsigned long long bd[4] = {dup2(b.x), dup2(b.y), dup2(b.z), dup2(b.w)};
#pragma unroll
      for (int i = 0; i < ST_TM / 2; ++i)
#pragma unroll
        for (int j = 0; j < ST_TN; ++j) ffma2(acc2[i][j], ap[i], bd[j]);
    }
  }
  float acc[ST_TM][ST_TN];
#pragma unroll
  for (int i = 0; i < ST_TM / 2; ++i)
#pragma unroll
    for (int j = 0; j < ST_TN; ++j) unpack2(acc2[i][j], acc[2 * i][j], acc[2 * i + 1][j]);

  const int c0 = n0 + tn * ST_TN;
  if (c0 >= p.CoutP) return;
#pragma unroll
  for (int i = 0; i < ST_TM; ++i) {
    long long b = 0;
    int y = 0, x = 0;
    if (!decode(m0 + tm * ST_TM + i, b, y, x)) continue;
    apply_act<4>(p.o, y, x, c0, acc[i]);
    store_act<4>(p.o, b, y, x, c0, acc[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// latent: one warp per stamp.  lane i owns row i of the 32x32 scale_tril.
//   loc = t[0:32]; L = fill_triangular(t[32:560]) -> rows 0-15: L[i][j] = t[64+32i+j],
//   rows 16-31: L[i][j] = t[1055-32i-j] (j<=i); L[i][i] = softplus(L[i][i]) + 1e-5;
//   z = loc + L eps;  stddev_i = sqrt(sum_j L_ij^2);  zp = PReLU(z, alpha0) feeds the decoder.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) latent_kernel(const float* __restrict__ params, const float* __restrict__ eps,
                                                     unsigned long long seed, int sample, long long first_stamp, long long B,
                                                     float* __restrict__ z, float* __restrict__ loc_out,
                                                     float* __restrict__ std_out, float* __restrict__ zp,
                                                     const float* __restrict__ alpha0) {
  __shared__ float s_t[8][NPAR];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long k = (long long)blockIdx.x * 8 + w;
  if (k >= B) return;
  const float* __restrict__ t = params + k * NPAR;
  for (int i = lane; i < NPAR; i += 32) s_t[w][i] = t[i];
  __syncwarp();
  float e;
  if (eps) {
    e = eps[k * LAT + lane];
  } else if (sample) {
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)(first_stamp + k) * LAT + lane, 0ull, &st);
    e = curand_normal(&st);
  } else {
    e = 0.f;
  }
  const int i = lane;
  float acc = s_t[w][i];
  float ss = 0.f;
  for (int j = 0; j < LAT; ++j) {  // uniform trip count: every lane takes part in every shuffle
    const float ej = __shfl_sync(0xffffffffu, e, j);
    if (j <= i) {
      float l = (i < 16) ? s_t[w][64 + 32 * i + j] : s_t[w][1055 - 32 * i - j];
      if (j == i) l = (l > 20.f ? l : log1pf(expf(l))) + 1e-5f;
      acc = fmaf(l, ej, acc);
      ss = fmaf(l, l, ss);
    }
  }
  z[k * LAT + i] = acc;
  if (loc_out) loc_out[k * LAT + i] = s_t[w][i];
  if (std_out) std_out[k * LAT + i] = sqrtf(ss);
  if (zp) zp[k * LAT + i] = prelu_f(acc, alpha0[i]);
}

__global__ void __launch_bounds__(256) prelu_vec_kernel(const float* __restrict__ z, const float* __restrict__ alpha,
                                                        long long n, int C, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = prelu_f(z[i], alpha[i % C]);
}

// tf.cast(images, tf.float32) for float64 host input (deblend_cutout/deblender.py:18)
__global__ void __launch_bounds__(256) cast_f64_f32_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    const double2 v = *reinterpret_cast<const double2*>(in + i);
    *reinterpret_cast<float2*>(out + i) = make_float2(__double2float_rn(v.x), __double2float_rn(v.y));
  } else if (i < n) {
    out[i] = __double2float_rn(in[i]);
  }
}

// debug: bf16 activation buffer (plain or parity, 1 or 2 planes) -> fp32 NHWC
__global__ void __launch_bounds__(256) act_to_f32_kernel(OutSpec o, long long B, float* __restrict__ out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = B * o.OH * o.OW * o.Cout;
  if (gid >= n) return;
  const int c = (int)(gid % o.Cout);
  const long long pix = gid / o.Cout;
  const int x = (int)(pix % o.OW), y = (int)((pix / o.OW) % o.OH);
  const long long b = pix / ((long long)o.OW * o.OH);
  if (o.mode == OUT_BF16_CG8) {
    const long long gstride = (long long)o.OH * o.OW * 8;
    const long long e = pixel_offset(o, b, y, x) + (long long)(c >> 3) * gstride + (c & 7);
    const long long pl = (long long)(o.Cpad >> 3) * gstride;
    float v;
    if (o.f16) {
      const __half* p = reinterpret_cast<const __half*>(o.out);
      v = __half2float(p[e]) + (o.planes == 2 ? __half2float(p[e + pl]) : 0.f);
    } else {
      const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(o.out);
      v = __bfloat162float(p[e]) + (o.planes == 2 ? __bfloat162float(p[e + pl]) : 0.f);
    }
    out[gid] = v;
    return;
  }
  const long long off = pixel_offset(o, b, y, x) + c;
  if (o.mode == OUT_F32_NHWC) {
    out[gid] = reinterpret_cast<const float*>(o.out)[off];
  } else {
    float v;
    if (o.f16) {
      const __half* p = reinterpret_cast<const __half*>(o.out);
      v = __half2float(p[off]);
      if (o.planes == 2) v += __half2float(p[off + o.Cpad]);
    } else {
      const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(o.out);
      v = __bfloat162float(p[off]);
      if (o.planes == 2) v += __bfloat162float(p[off + o.Cpad]);
    }
    out[gid] = v;
  }
}

// BatchNorm of the input stamp (model/model.py:79; TF pads AFTER the BN, so it cannot be folded into conv1's
// bias) and conversion to the tensor-core operand of conv1: bf16 hi[/lo], 6 bands padded to 8 channels =
// one 16-byte row per pixel and plane.  One thread per pixel.
__global__ void __launch_bounds__(256) bn_pack8_kernel(const float* __restrict__ x, const float* __restrict__ sc,
                                                       const float* __restrict__ sh, long long npix, OutSpec o) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const float2* xp = reinterpret_cast<const float2*>(x + pix * CB_);  // 24-byte pixels: 8-byte aligned
  const float2 a = __ldg(xp), b = __ldg(xp + 1), c = __ldg(xp + 2);
  float v[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
#pragma unroll
  for (int ch = 0; ch < 6; ++ch) v[ch] = fmaf(v[ch], __ldg(sc + ch), __ldg(sh + ch));
  uint4 q, l;
  split16x2(o.f16, v[0], v[1], q.x, l.x);
  split16x2(o.f16, v[2], v[3], q.y, l.y);
  split16x2(o.f16, v[4], v[5], q.z, l.z);
  q.w = l.w = 0u;
  __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(o.out) + pix * (long long)(o.planes * o.Cpad);
  *reinterpret_cast<uint4*>(p) = q;
  if (o.planes == 2) *reinterpret_cast<uint4*>(p + o.Cpad) = l;
}

int launch_bn_pack8(const float* x, const float* bn_scale, const float* bn_shift, long long B, const OutSpec& o, cudaStream_t st) {
  const long long npix = B * S_ * S_;
  if (npix == 0) return DBV_OK;
  bn_pack8_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(x, bn_scale, bn_shift, npix, o);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

template <int BN, int KC>
static int launch_simt_tile(const SimtConv& p, cudaStream_t st) {
  constexpr int BM = (256 / (BN / ST_TN)) * ST_TM;
  const int ncls = p.mode == 1 ? 4 : 1;
  const long long npix = p.mode == 1 ? (long long)p.B * (p.Hout >> 1) * (p.Wout >> 1) : (long long)p.B * p.Hout * p.Wout;
  const long long tiles_m = (npix + BM - 1) / BM;
  if (tiles_m * ncls >= (1ll << 31)) return fail(DBV_ERR_INVALID, "simt conv: batch too large");
  dim3 grid((unsigned)(tiles_m * ncls), (unsigned)((p.CoutP + BN - 1) / BN));
  simt_tile_kernel<BN, KC><<<grid, 256, 0, st>>>(p, (int)tiles_m);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

int launch_simt_conv(const SimtConv& p, cudaStream_t st) {
  const long long threads = (long long)p.B * p.Hout * p.Wout * (p.CoutP >> 2);
  if (threads == 0) return DBV_OK;
  bool tiled = (p.CoutP & 3) == 0 && (p.mode == 0 || ((p.Hout & 1) == 0 && (p.Wout & 1) == 0 && p.ksz == 3));
  if (const char* e = dbv_env("DBV_SIMT_TILED")) tiled = tiled && atoi(e) != 0;  // 0: the naive gather kernel (cross-check)
  if (tiled) {
    if (p.CoutP >= 64) {
      static const int kc = dbv_env("DBV_SIMT_KC") ? atoi(dbv_env("DBV_SIMT_KC")) : 32;  // tuning knob (measured: 32 is 6-8 % faster than 16)
      return (kc == 32 && p.Cin >= 32) ? launch_simt_tile<64, 32>(p, st) : launch_simt_tile<64, 16>(p, st);
    }
    static const int kcs = dbv_env("DBV_SIMT_KC_SMALL") ? atoi(dbv_env("DBV_SIMT_KC_SMALL")) : 16;  // tuning knob (8 measured 3-4 % slower)
    if (p.CoutP >= 32) return (p.Cin <= 8 || kcs == 8) ? launch_simt_tile<32, 8>(p, st) : launch_simt_tile<32, 16>(p, st);  // conv1: Cin = 6
    return kcs == 8 ? launch_simt_tile<16, 8>(p, st) : launch_simt_tile<16, 16>(p, st);
  }
  const long long blocks = (threads + 255) / 256;
  simt_conv_kernel<<<(unsigned)blocks, 256, 0, st>>>(p);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

int launch_latent(const float* params, const float* eps, unsigned long long seed, int sample, long long first_stamp,
                  long long B, float* z, float* loc, float* std_out, float* zp, const float* alpha0, cudaStream_t st) {
  if (B == 0) return DBV_OK;
  latent_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(params, eps, seed, sample, first_stamp, B, z, loc, std_out, zp, alpha0);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

int launch_prelu_vec(const float* z, const float* alpha, long long n, int C, float* out, cudaStream_t st) {
  if (n == 0) return DBV_OK;
  prelu_vec_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(z, alpha, n, C, out);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

int launch_cast_f64_f32(const double* in, float* out, long long n, cudaStream_t st) {
  if (n == 0) return DBV_OK;
  const long long thr = (n + 1) / 2;
  cast_f64_f32_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, st>>>(in, out, n);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

int launch_act_to_f32(const OutSpec& o, long long B, float* out, cudaStream_t st) {
  const long long n = B * o.OH * o.OW * o.Cout;
  if (n == 0) return DBV_OK;
  act_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(o, B, out);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

}  // namespace dbv
