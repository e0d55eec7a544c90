// Output addressing + bias / PReLU / ReLU epilogue shared by the SIMT and the tcgen05 kernels.
//
// Activation tensors between layers are NHWC.  Three storage forms:
//   OUT_F32_NHWC     fp32 [B][OH][OW][Cstore]                       (fp32 path, encoder params)
//   OUT_BF16_NHWC    bf16 [B][OH][OW][planes*Cpad]                  (consumer is a stride-1 layer)
//   OUT_BF16_PARITY  bf16 [B][4][PH][PW][planes*Cpad], plane = (y&1)*2+(x&1), row y>>1, col x>>1
//                    (consumer is a stride-2 Conv2D: each of its taps then reads ONE plane with
//                    unit stride, so a plain tiled TMA box + out-of-bounds zero fill implements
//                    TF "SAME" padding; rows/cols a plane lacks stay zero from allocation)
//   OUT_HEAD         decoder head (model/model.py:137-159): ReLU, crop [2:61]^2 of the 64x64 map,
//                    channels 0-5 -> mean, 6-11 -> 1e-4 + v -> stddev, both fp32 (B,59,59,6)
//   OUT_BF16_CG8     16-bit [B][planes][Cpad/8][OH][OW][8]: channel groups of 8 are PLANAR (the consumer is a resident-halo
//                    layer).  One pixel of one group is 16 bytes and eight consecutive pixels are one 128-byte tcgen05 core
//                    matrix, so (a) the consumer loads whole image rows of all groups with ONE un-swizzled TMA box and reads
//                    a K=16 operand as two groups one region apart (descriptor LBO), and (b) in the producer's epilogue,
//                    where lane = pixel, a warp's 128-bit store covers 512 CONTIGUOUS bytes: 4 L1 wavefronts instead of the
//                    32 of a pixel-major layout (measured: the pixel-major stores saturate the L1 data pipe, which is shared
//                    with the MMA's operand fetch from shared memory — tools/ + DESIGN.md section 4).
// planes = 2 stores a hi/lo bf16 split (v ~= hi + lo) at channel offsets [0,Cpad) and [Cpad,2Cpad).
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace dbv {

enum OutMode { OUT_F32_NHWC = 0, OUT_BF16_NHWC = 1, OUT_BF16_PARITY = 2, OUT_HEAD = 3, OUT_BF16_CG8 = 4 };

struct OutSpec {
  void* out;
  void* out2;          // OUT_HEAD: stddev
  int mode;
  int planes;          // 16-bit modes: 1 or 2
  int f16;             // 16-bit modes: 0 = bf16, 1 = fp16 storage (DBV_PREC_FP16X3)
  int OH, OW;          // full output image extents (addressing + alpha indexing)
  int Cout;            // real channels
  int Cpad;            // channels per plane as stored (>= Cout)
  int PH, PW;          // OUT_BF16_PARITY plane extents
  const float* bias;   // [Cout]
  const float* alpha;  // PReLU slopes of the (OH,OW,Cout) map in the library's layout [Cout/4][OH*OW][4] (see alpha_index), or null
  const float* alpha2; // second PReLU (encoder Flatten PReLU, model/model.py:95), same layout, or null
  int relu;
  int alpha_le1;       // every PReLU slope of this layer (alpha and alpha2) is <= 1: prelu(v) = max(v, a v) (tc_ptx.cuh:prelu4)
  int* ovf;            // single-plane fp16 outputs (the tail of DBV_PREC_MIXED): host-mapped flag set to 1 when a value saturates
                       // at +-65504, so that leaving the fp16 range fails loudly instead of silently (dbv_fp16_overflow); or null
};

__device__ __forceinline__ float prelu_f(float v, float a) { return v > 0.f ? v : a * v; }

// PReLU slopes are stored channel-group-major: element (pixel, c) at ((c/4) * npix + pixel) * 4 + c%4.  In the
// tensor-core epilogues lane = pixel, so a warp's 16-byte load of four channels covers 512 contiguous bytes
// (4 L1 wavefronts) instead of 32 separate lines; the L1 data path is shared with the MMA operand fetch from
// shared memory, which is what bounds the few-channel layers.  Cout % 4 == 0 everywhere.
__host__ __device__ __forceinline__ long long alpha_index(long long npix, long long pix, int c) {
  return ((long long)(c >> 2) * npix + pix) * 4 + (c & 3);
}

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256); pointers must be 32-byte aligned
__device__ __forceinline__ void ldg256(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
               "r"(b.z), "r"(b.w)
               : "memory");
}

// element offset of channel 0 of pixel (b,y,x); OUT_HEAD handled by the callers
__device__ __forceinline__ long long pixel_offset(const OutSpec& o, long long b, int y, int x) {
  const int cs = o.planes * o.Cpad;
  if (o.mode == OUT_BF16_CG8)  // offset of (plane 0, group 0); + (plane * Cpad/8 + group) * OH*OW*8 for the others
    return b * cs * o.OH * o.OW + ((long long)y * o.OW + x) * 8;
  if (o.mode == OUT_BF16_PARITY) {
    const long long plane = b * 4 + ((y & 1) * 2 + (x & 1));
    return ((plane * o.PH + (y >> 1)) * o.PW + (x >> 1)) * cs;
  }
  return ((b * o.OH + y) * o.OW + x) * (long long)cs;
}

// bias + PReLU(+PReLU) / ReLU on NV consecutive channels starting at c of pixel (y,x).
// bias_off shifts the bias index only (Dense -> Reshape layers whose N tile is a pixel).
template <int NV>
__device__ __forceinline__ void apply_act(const OutSpec& o, int y, int x, int c, float (&v)[NV], int bias_off = 0) {
  const long long pix = (long long)y * o.OW + x, npix = (long long)o.OH * o.OW;
  if (c + NV <= o.Cout && (o.Cout & 3) == 0) {  // fast path: 16-byte loads (c is a multiple of 4)
    const float4* bp = reinterpret_cast<const float4*>(o.bias + bias_off + c);
    const float4* ap = o.alpha ? reinterpret_cast<const float4*>(o.alpha) + (long long)(c >> 2) * npix + pix : nullptr;
    const float4* a2p = o.alpha2 ? reinterpret_cast<const float4*>(o.alpha2) + (long long)(c >> 2) * npix + pix : nullptr;
#pragma unroll
    for (int j = 0; j < NV; j += 4) {
      const float4 bb = __ldg(bp + (j >> 2));
      v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
      if (ap) {
        const float4 a = __ldg(ap + (long long)(j >> 2) * npix);
        v[j] = prelu_f(v[j], a.x); v[j + 1] = prelu_f(v[j + 1], a.y); v[j + 2] = prelu_f(v[j + 2], a.z); v[j + 3] = prelu_f(v[j + 3], a.w);
      }
      if (a2p) {
        const float4 a = __ldg(a2p + (long long)(j >> 2) * npix);
        v[j] = prelu_f(v[j], a.x); v[j + 1] = prelu_f(v[j + 1], a.y); v[j + 2] = prelu_f(v[j + 2], a.z); v[j + 3] = prelu_f(v[j + 3], a.w);
      }
      if (o.relu) {
        v[j] = fmaxf(v[j], 0.f); v[j + 1] = fmaxf(v[j + 1], 0.f); v[j + 2] = fmaxf(v[j + 2], 0.f); v[j + 3] = fmaxf(v[j + 3], 0.f);
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (c + j < o.Cout) {
      float t = v[j] + __ldg(o.bias + bias_off + c + j);
      if (o.alpha) t = prelu_f(t, __ldg(o.alpha + alpha_index(npix, pix, c + j)));
      if (o.alpha2) t = prelu_f(t, __ldg(o.alpha2 + alpha_index(npix, pix, c + j)));
      if (o.relu) t = fmaxf(t, 0.f);
      v[j] = t;
    } else {
      v[j] = 0.f;
    }
  }
}

// ---- the same epilogue with its global loads hoisted: bias / alpha vectors are fetched into registers
// (act_prefetch) BEFORE the accumulator is waited for, so their L2 latency overlaps the MMA / TMEM wait.
template <int NV>
struct ActRegs {
  float4 a[NV / 4];  // PReLU alpha of this thread's pixel (the bias is warp-uniform and L1-resident: loaded late)
  bool fast;
};

template <int NV>
__device__ __forceinline__ void act_prefetch(const OutSpec& o, bool ok, int y, int x, int c, int bias_off, ActRegs<NV>& r) {
  r.fast = ok && (c + NV <= o.Cout) && ((o.Cout & 3) == 0) && o.alpha != nullptr;
  if (r.fast) {
    const long long npix = (long long)o.OH * o.OW;
    const float4* ap = reinterpret_cast<const float4*>(o.alpha) + (long long)(c >> 2) * npix + ((long long)y * o.OW + x);
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) r.a[j] = __ldg(ap + (long long)j * npix);
  }
}

// BIASED: the caller has already added the bias (from the kernel's constant bank) — only valid with r.fast or the head
template <int NV, bool BIASED = false>
__device__ __forceinline__ void act_apply(const OutSpec& o, int y, int x, int c, int bias_off, const ActRegs<NV>& r, float (&v)[NV]) {
  if (!r.fast) {
    if constexpr (BIASED) {
      if (o.mode == OUT_HEAD) {  // decoder head: bias already in, no PReLU, ReLU on every channel (model/model.py:137)
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = fmaxf(v[j], 0.f);
        return;
      }
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (c + j < o.Cout) v[j] -= __ldg(o.bias + bias_off + c + j);  // generic path re-adds it
    }
    apply_act<NV>(o, y, x, c, v, bias_off);
    return;
  }
  const float4* bp = reinterpret_cast<const float4*>(o.bias + bias_off + c);
#pragma unroll
  for (int j = 0; j < NV / 4; ++j) {
    const float4 bb = BIASED ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(bp + j);
    v[4 * j + 0] = prelu_f(v[4 * j + 0] + bb.x, r.a[j].x);
    v[4 * j + 1] = prelu_f(v[4 * j + 1] + bb.y, r.a[j].y);
    v[4 * j + 2] = prelu_f(v[4 * j + 2] + bb.z, r.a[j].z);
    v[4 * j + 3] = prelu_f(v[4 * j + 3] + bb.w, r.a[j].w);
  }
  if (o.alpha2) {
    const long long npix = (long long)o.OH * o.OW;
    const float4* a2p = reinterpret_cast<const float4*>(o.alpha2) + (long long)(c >> 2) * npix + ((long long)y * o.OW + x);
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
      const float4 a = __ldg(a2p + (long long)j * npix);
      v[4 * j + 0] = prelu_f(v[4 * j + 0], a.x);
      v[4 * j + 1] = prelu_f(v[4 * j + 1], a.y);
      v[4 * j + 2] = prelu_f(v[4 * j + 2], a.z);
      v[4 * j + 3] = prelu_f(v[4 * j + 3], a.w);
    }
  }
  if (o.relu) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = fmaxf(v[j], 0.f);
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// fp16 conversions SATURATE at +-65504 (one F2FP.SATFINITE), so that a large activation never becomes inf
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));  // a -> low half, b -> high half
  return r;
}
// fp16 hi/lo split (hi saturates, lo = v - hi stays finite)
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16x2(a, b);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = pack_f16x2(a - hf.x, b - hf.y);
}
// running max |h| of packed fp16 pairs (saturation watch of the single-plane fp16 outputs): one HMNMX2 per pair
__device__ __forceinline__ uint32_t habs2_max(uint32_t m, uint32_t q) {
  uint32_t a;
  asm("abs.f16x2 %0, %1;" : "=r"(a) : "r"(q));
  asm("max.f16x2 %0, %1, %2;" : "=r"(m) : "r"(m), "r"(a));
  return m;
}
__device__ __forceinline__ void ovf_report(int* flag, uint32_t m) {
  if ((m & 0xffffu) >= 0x7bffu || (m >> 16) >= 0x7bffu) *reinterpret_cast<volatile int*>(flag) = 1;  // 0x7bff = 65504, the saturation value
}
// format-dispatching versions (f16 is warp-uniform)
__device__ __forceinline__ uint32_t pack16x2(int f16, float a, float b) { return f16 ? pack_f16x2(a, b) : pack_bf16x2(a, b); }
__device__ __forceinline__ float round16(int f16, float a) {
  if (!f16) return __bfloat162float(__float2bfloat16_rn(a));
  const uint32_t h = pack_f16x2(a, 0.f);
  return __half2float(*reinterpret_cast<const __half*>(&h));
}
__device__ __forceinline__ float bf16_round(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }
// hi = bf16x2(a,b); lo = bf16x2(a - hi.a, b - hi.b): 6 instructions per pair
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(a, b);
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
  lo = pack_bf16x2(a - ha, b - hb);
}

__device__ __forceinline__ void split16x2(int f16, float a, float b, uint32_t& hi, uint32_t& lo) {
  if (f16) split_f16x2(a, b, hi, lo);
  else split_bf16x2(a, b, hi, lo);
}

// store NV (multiple of 4, c multiple of 4) activated channels of pixel (b,y,x)
template <int NV>
__device__ __forceinline__ void store_act(const OutSpec& o, long long b, int y, int x, int c, const float (&v)[NV]) {
  if (o.mode == OUT_HEAD) {
    if (y < 2 || y >= 61 || x < 2 || x >= 61) return;
    const long long base = ((b * 59 + (y - 2)) * 59 + (x - 2)) * 6;
    if constexpr (NV >= 12) {
      if (c == 0) {  // 24-byte pixels: three 8-byte stores per output
        float2* pm = reinterpret_cast<float2*>(reinterpret_cast<float*>(o.out) + base);
        pm[0] = make_float2(v[0], v[1]);
        pm[1] = make_float2(v[2], v[3]);
        pm[2] = make_float2(v[4], v[5]);
        if (o.out2) {
          float2* ps = reinterpret_cast<float2*>(reinterpret_cast<float*>(o.out2) + base);
          ps[0] = make_float2(1e-4f + v[6], 1e-4f + v[7]);
          ps[1] = make_float2(1e-4f + v[8], 1e-4f + v[9]);
          ps[2] = make_float2(1e-4f + v[10], 1e-4f + v[11]);
        }
        return;
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int cc = c + j;
      if (cc < 6) reinterpret_cast<float*>(o.out)[base + cc] = v[j];
      else if (cc < 12 && o.out2) reinterpret_cast<float*>(o.out2)[base + cc - 6] = 1e-4f + v[j];
    }
    return;
  }
  const long long off = pixel_offset(o, b, y, x);
  if (o.mode == OUT_F32_NHWC) {
    float* p = reinterpret_cast<float*>(o.out) + off + c;
    if (c + NV <= o.Cout && (o.Cpad & 3) == 0) {
#pragma unroll
      for (int j = 0; j < NV; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (c + j < o.Cout) p[j] = v[j];
    }
    return;
  }
  if (o.mode == OUT_BF16_CG8) {
    if constexpr (NV % 8 == 0) {
      const long long gstride = (long long)o.OH * o.OW * 8;  // elements between channel groups (and, x Cpad/8, between planes)
      uint16_t* p = reinterpret_cast<uint16_t*>(o.out) + off + (long long)(c >> 3) * gstride;
      const long long pstride = (long long)(o.Cpad >> 3) * gstride;
#pragma unroll
      for (int j = 0; j < NV; j += 8) {
        if (c + j >= o.Cpad) break;
        uint4 q, l;
        if (o.planes == 2) {
          split16x2(o.f16, v[j], v[j + 1], q.x, l.x);
          split16x2(o.f16, v[j + 2], v[j + 3], q.y, l.y);
          split16x2(o.f16, v[j + 4], v[j + 5], q.z, l.z);
          split16x2(o.f16, v[j + 6], v[j + 7], q.w, l.w);
        } else {
          q.x = pack16x2(o.f16, v[j], v[j + 1]);
          q.y = pack16x2(o.f16, v[j + 2], v[j + 3]);
          q.z = pack16x2(o.f16, v[j + 4], v[j + 5]);
          q.w = pack16x2(o.f16, v[j + 6], v[j + 7]);
        }
        *reinterpret_cast<uint4*>(p + (long long)(j >> 3) * gstride) = q;
        if (o.planes == 2) *reinterpret_cast<uint4*>(p + pstride + (long long)(j >> 3) * gstride) = l;
      }
    }
    return;
  }
  // bf16 (hi[/lo]) — Cpad is a multiple of 8 and c of 4, so 8-byte stores are aligned
  __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(o.out) + off + c;
  const bool full = c + NV <= o.Cpad;
  if (full) {
    if constexpr (NV % 16 == 0) {
      // 16 channels = 32 bytes per plane: 256-bit stores when the pixel base is 32-byte aligned (Cpad % 16 == 0)
      const bool a32 = (o.Cpad & 15) == 0 && (c & 15) == 0;
#pragma unroll
      for (int j = 0; j < NV; j += 16) {
        uint4 q0, q1, l0, l1;
        if (o.planes == 2) {
          split16x2(o.f16, v[j], v[j + 1], q0.x, l0.x);
          split16x2(o.f16, v[j + 2], v[j + 3], q0.y, l0.y);
          split16x2(o.f16, v[j + 4], v[j + 5], q0.z, l0.z);
          split16x2(o.f16, v[j + 6], v[j + 7], q0.w, l0.w);
          split16x2(o.f16, v[j + 8], v[j + 9], q1.x, l1.x);
          split16x2(o.f16, v[j + 10], v[j + 11], q1.y, l1.y);
          split16x2(o.f16, v[j + 12], v[j + 13], q1.z, l1.z);
          split16x2(o.f16, v[j + 14], v[j + 15], q1.w, l1.w);
        } else {
          q0.x = pack16x2(o.f16, v[j], v[j + 1]);
          q0.y = pack16x2(o.f16, v[j + 2], v[j + 3]);
          q0.z = pack16x2(o.f16, v[j + 4], v[j + 5]);
          q0.w = pack16x2(o.f16, v[j + 6], v[j + 7]);
          q1.x = pack16x2(o.f16, v[j + 8], v[j + 9]);
          q1.y = pack16x2(o.f16, v[j + 10], v[j + 11]);
          q1.z = pack16x2(o.f16, v[j + 12], v[j + 13]);
          q1.w = pack16x2(o.f16, v[j + 14], v[j + 15]);
          l0 = l1 = make_uint4(0u, 0u, 0u, 0u);
          if (o.ovf) {
            uint32_t m = habs2_max(habs2_max(habs2_max(habs2_max(0u, q0.x), q0.y), q0.z), q0.w);
            m = habs2_max(habs2_max(habs2_max(habs2_max(m, q1.x), q1.y), q1.z), q1.w);
            ovf_report(o.ovf, m);
          }
        }
        if (a32) {
          stg256(p + j, q0, q1);
          if (o.planes == 2) stg256(p + o.Cpad + j, l0, l1);
        } else {
          *reinterpret_cast<uint4*>(p + j) = q0;
          *reinterpret_cast<uint4*>(p + j + 8) = q1;
          if (o.planes == 2) {
            *reinterpret_cast<uint4*>(p + o.Cpad + j) = l0;
            *reinterpret_cast<uint4*>(p + o.Cpad + j + 8) = l1;
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; j += 4) {
        uint2 q;
        q.x = pack16x2(o.f16, v[j], v[j + 1]);
        q.y = pack16x2(o.f16, v[j + 2], v[j + 3]);
        if (o.ovf && o.planes == 1) ovf_report(o.ovf, habs2_max(habs2_max(0u, q.x), q.y));
        *reinterpret_cast<uint2*>(p + j) = q;
        if (o.planes == 2) {
          uint2 r;
          r.x = pack16x2(o.f16, v[j] - round16(o.f16, v[j]), v[j + 1] - round16(o.f16, v[j + 1]));
          r.y = pack16x2(o.f16, v[j + 2] - round16(o.f16, v[j + 2]), v[j + 3] - round16(o.f16, v[j + 3]));
          *reinterpret_cast<uint2*>(p + o.Cpad + j) = r;
        }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (c + j < o.Cpad) {
        const float hv = round16(o.f16, v[j]);
        const uint32_t hb = pack16x2(o.f16, hv, 0.f), lb = pack16x2(o.f16, v[j] - hv, 0.f);
        reinterpret_cast<uint16_t*>(p)[j] = (uint16_t)hb;
        if (o.planes == 2) reinterpret_cast<uint16_t*>(p)[o.Cpad + j] = (uint16_t)lb;
      }
    }
  }
}

}  // namespace dbv
