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

int launch_simt_conv(const SimtConv& p, cudaStream_t st) {
  const long long threads = (long long)p.B * p.Hout * p.Wout * (p.CoutP >> 2);
  if (threads == 0) return DBV_OK;
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
