// Field-level operators: stamp extraction (gather), deterministic windowed subtract / scatter-add,
// centre-window MSE and field MSE.  All are HBM-bound copy / read-modify-write kernels:
// coalesced, vectorised where alignment allows, no tensor cores.
//
//   dbv_extract      <- extract/extraction.py:21-36
//   dbv_window_axpy  <- deblend/field_deblender.py:46-97 (residual), :99-189 (predicted fields)
//   dbv_center_mse   <- deblend/field_deblender.py:323-332
//   dbv_mse          <- training/metrics.py:4-12
#include "common.cuh"
#include <cstdlib>

namespace dbv {

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

template <typename Tout, typename Tin> __device__ __forceinline__ Tout cvt(Tin v);
template <> __device__ __forceinline__ double cvt<double, double>(double v) { return v; }
template <> __device__ __forceinline__ float cvt<float, float>(float v) { return v; }
template <> __device__ __forceinline__ float cvt<float, double>(double v) { return __double2float_rn(v); }
template <> __device__ __forceinline__ double cvt<double, float>(float v) { return (double)v; }

// ---------------------------------------------------------------------------------------------
// extraction: one stamp per blockIdx.x, gridDim.y CTAs cooperate on it.  A stamp row is S*C
// contiguous elements in the field (S*C*8 = 2832 B for DC2/f64) and the whole stamp is contiguous
// in the output, so both sides are streamed with 16-byte (f64) / 8-byte (f32) vectors, 4 in flight
// per thread.
// ---------------------------------------------------------------------------------------------
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256) extract_kernel(const Tin* __restrict__ field, long long F, int C,
                                                      const int32_t* __restrict__ sx, const int32_t* __restrict__ sy,
                                                      const uint8_t* __restrict__ flags, const int64_t* __restrict__ slot,
                                                      int S, Tout* __restrict__ out) {
  using VI = typename Vec2<Tin>::type;
  using VO = typename Vec2<Tout>::type;
  const long long k = blockIdx.x;
  const int x0 = sx[k], y0 = sy[k];
  const int fl = flags ? flags[k] : 0;
  const long long dst = slot ? slot[k] : k;
  const int L = S * C;
  Tout* __restrict__ o = out + dst * (long long)S * L;
  const int tid = blockIdx.y * blockDim.x + threadIdx.x;
  const int nthr = gridDim.y * blockDim.x;
  if (fl == 0 && (C & 1) == 0) {
    const int LV = L >> 1;
    const int total = S * LV;
    const Tin* __restrict__ base = field + ((long long)x0 * F + y0) * C;
    const long long rstride = F * C;
    int idx = tid;
    for (; idx + 3 * nthr < total; idx += 4 * nthr) {
      VI v[4];
      int r[4], c[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int id = idx + j * nthr;
        r[j] = id / LV;
        c[j] = id - r[j] * LV;
        v[j] = __ldg(reinterpret_cast<const VI*>(base + r[j] * rstride) + c[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        VO w;
        w.x = cvt<Tout, Tin>(v[j].x);
        w.y = cvt<Tout, Tin>(v[j].y);
        __stcs(reinterpret_cast<VO*>(o + (long long)r[j] * L) + c[j], w);
      }
    }
    for (; idx < total; idx += nthr) {
      const int r = idx / LV, c = idx - r * LV;
      VI v = __ldg(reinterpret_cast<const VI*>(base + r * rstride) + c);
      VO w;
      w.x = cvt<Tout, Tin>(v.x);
      w.y = cvt<Tout, Tin>(v.y);
      __stcs(reinterpret_cast<VO*>(o + (long long)r * L) + c, w);
    }
  } else {
    // generic path: odd C, or a length-1 source axis broadcast over the stamp (numpy assignment)
    const int bx = fl & 1, by = (fl >> 1) & 1;
    const int total = S * L;
    for (int idx = tid; idx < total; idx += nthr) {
      const int r = idx / L;
      const int rem = idx - r * L;
      const int c = rem / C, ch = rem - c * C;
      const long long xr = x0 + (bx ? 0 : r), yc = y0 + (by ? 0 : c);
      o[idx] = cvt<Tout, Tin>(__ldg(field + (xr * F + yc) * C + ch));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// windowed axpy, owner-computes: a CTA owns a TR x TC pixel tile of the field, collects (in
// ascending stamp index) the stamps that overlap it, and applies them to each of its pixels in
// that order.  No atomics; one rounding per addition; bit-identical to the sequential host loop.
// ---------------------------------------------------------------------------------------------
constexpr int AX_TR = 32, AX_TC = 64, AX_CAP = 768, AX_THREADS = 256;

template <typename T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }

template <typename T>
__global__ void __launch_bounds__(AX_THREADS) window_axpy_kernel(const T* in, T* out, long long F, int C,
                                                                  const float* __restrict__ stamps,
                                                                  const int32_t* __restrict__ x0, const int32_t* __restrict__ y0,
                                                                  int N, int S, double alpha_d, int tiles_c) {
  __shared__ int s_id[AX_CAP], s_x[AX_CAP], s_y[AX_CAP];
  __shared__ int s_wcnt[AX_THREADS / 32];
  __shared__ int s_count, s_next;
  const int tr0 = (blockIdx.x / tiles_c) * AX_TR;
  const int tc0 = (blockIdx.x % tiles_c) * AX_TC;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const T alpha = (T)alpha_d;
  const int nelt = AX_TR * AX_TC * C;
  const long long stamp_sz = (long long)S * S * C;
  int start = 0;
  bool first = true;
  do {
    if (threadIdx.x == 0) { s_count = 0; s_next = N; }
    __syncthreads();
    // ordered compaction of the overlapping stamps of [start, N)
    for (int i0 = start; i0 < N; i0 += AX_THREADS) {
      const int i = i0 + threadIdx.x;
      int ov = 0, xi = 0, yi = 0;
      if (i < N) {
        xi = x0[i];
        yi = y0[i];
        ov = (xi < tr0 + AX_TR) && (xi + S > tr0) && (yi < tc0 + AX_TC) && (yi + S > tc0);
      }
      const unsigned m = __ballot_sync(0xffffffffu, ov);
      if (lane == 0) s_wcnt[wid] = __popc(m);
      __syncthreads();
      int before = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < AX_THREADS / 32; ++w) {
        const int c = s_wcnt[w];
        if (w < wid) before += c;
        tot += c;
      }
      const int base = s_count;
      const bool fits = base + tot <= AX_CAP;
      if (fits && ov) {
        const int p = base + before + __popc(m & ((1u << lane) - 1u));
        s_id[p] = i;
        s_x[p] = xi;
        s_y[p] = yi;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        if (fits) s_count = base + tot;
        else s_next = i0;
      }
      __syncthreads();
      if (!fits) break;
    }
    const int cnt = s_count;
    if ((cnt > 0 || first) && (C & 1) == 0) {
      // vector path: 2 bands per thread (16-byte field accesses for f64, 8-byte stamp reads); a pixel's C
      // values start at an even element index, so the pairs never straddle pixels
      using V2 = typename Vec2<T>::type;
      const int rowv = AX_TC * C / 2;  // vectors per tile row
      const int nvec = AX_TR * rowv;
      for (int e = threadIdx.x; e < nvec; e += AX_THREADS) {
        const int pr = e / rowv;
        const int rem = e - pr * rowv;
        const int X = tr0 + pr;
        const int Y = tc0 + (2 * rem) / C;
        const int ch = 2 * rem - ((2 * rem) / C) * C;
        if (X >= F || Y >= F) continue;
        const long long idx = ((long long)X * F + Y) * C + ch;
        V2 acc;
        if (first) {
          if (in) acc = *reinterpret_cast<const V2*>(in + idx);
          else { acc.x = (T)0; acc.y = (T)0; }
        } else {
          acc = *reinterpret_cast<const V2*>(out + idx);
        }
        for (int k = 0; k < cnt; ++k) {
          const int dx = X - s_x[k], dy = Y - s_y[k];
          if ((unsigned)dx < (unsigned)S && (unsigned)dy < (unsigned)S) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(stamps + s_id[k] * stamp_sz + ((long long)dx * S + dy) * C + ch));
            acc.x = add_rn<T>(acc.x, mul_rn<T>(alpha, (T)v.x));
            acc.y = add_rn<T>(acc.y, mul_rn<T>(alpha, (T)v.y));
          }
        }
        *reinterpret_cast<V2*>(out + idx) = acc;
      }
    } else if (cnt > 0 || first) {
      for (int e = threadIdx.x; e < nelt; e += AX_THREADS) {
        const int ch = e % C;
        const int pc = (e / C) % AX_TC;
        const int pr = e / (C * AX_TC);
        const int X = tr0 + pr, Y = tc0 + pc;
        if (X >= F || Y >= F) continue;
        const long long idx = ((long long)X * F + Y) * C + ch;
        T acc = first ? (in ? in[idx] : (T)0) : out[idx];
        for (int k = 0; k < cnt; ++k) {
          const int dx = X - s_x[k], dy = Y - s_y[k];
          if ((unsigned)dx < (unsigned)S && (unsigned)dy < (unsigned)S) {
            const float v = __ldg(stamps + s_id[k] * stamp_sz + ((long long)dx * S + dy) * C + ch);
            acc = add_rn<T>(acc, mul_rn<T>(alpha, (T)v));
          }
        }
        out[idx] = acc;
      }
    }
    first = false;
    start = s_next;
    __syncthreads();
  } while (start < N);
}

// ---------------------------------------------------------------------------------------------
// centre-window MSE: one warp per stamp, fp64, fixed shuffle tree.
// ---------------------------------------------------------------------------------------------
template <typename Tc>
__global__ void __launch_bounds__(256) center_mse_kernel(const Tc* __restrict__ cut, const float* __restrict__ mean,
                                                         long long N, int S, int C, int lo, int hi, double* __restrict__ out) {
  const long long k = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= N) return;
  const int lane = threadIdx.x & 31;
  const int W = hi - lo;
  const int row = W * C;
  const int total = W * row;
  const long long base = k * (long long)S * S * C;
  double acc = 0.0;
  for (int e = lane; e < total; e += 32) {
    const int r = e / row, rem = e - r * row;
    const long long idx = base + ((long long)(lo + r) * S + lo) * C + rem;
    const double d = (double)cut[idx] - (double)mean[idx];
    acc += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[k] = acc / (double)total;
}

// ---------------------------------------------------------------------------------------------
// field MSE: two deterministic passes (fixed grid, fixed tree).
// ---------------------------------------------------------------------------------------------
constexpr int MSE_BLOCKS = 1184;  // 8 CTAs per SM
template <typename T>
__global__ void __launch_bounds__(256) sqdiff_partial_kernel(const T* __restrict__ a, const T* __restrict__ b, long long n,
                                                             double* __restrict__ partial) {
  __shared__ double s[8];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double d = (double)__ldg(a + i) - (double)__ldg(b + i);
    acc += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) sqdiff_final_kernel(const double* __restrict__ partial, int nb, long long n,
                                                           double* __restrict__ out) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = s[0] / (double)n;
}

}  // namespace dbv

using namespace dbv;

extern "C" int dbv_extract(const void* field, int field_dtype, int64_t F, int C, const int32_t* sx, const int32_t* sy,
                           const uint8_t* flags, const int64_t* slot, int64_t N, int S, void* out, int out_dtype,
                           void* stream) {
  DBV_REQUIRE(field && sx && sy && out, "dbv_extract: null pointer");
  DBV_REQUIRE(F > 0 && C > 0 && S > 0 && N >= 0, "dbv_extract: bad sizes F=%lld C=%d S=%d N=%lld", (long long)F, C, S, (long long)N);
  DBV_REQUIRE(N < (1ll << 31), "dbv_extract: N too large");
  if (N == 0) return DBV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // enough CTAs per stamp to keep >= 4 waves of 148 SMs busy for small N, 1-2 for large N
  int per = 2;
  if (N < 2048) per = 4;
  if (N < 256) per = 8;
  if (const char* e = getenv("DBV_EXTRACT_PER")) per = atoi(e) > 0 ? atoi(e) : per;  // tuning knob
  dim3 grid((unsigned)N, per), block(256);
  if (field_dtype == DBV_F64 && out_dtype == DBV_F64)
    extract_kernel<double, double><<<grid, block, 0, st>>>((const double*)field, F, C, sx, sy, flags, slot, S, (double*)out);
  else if (field_dtype == DBV_F64 && out_dtype == DBV_F32)
    extract_kernel<double, float><<<grid, block, 0, st>>>((const double*)field, F, C, sx, sy, flags, slot, S, (float*)out);
  else if (field_dtype == DBV_F32 && out_dtype == DBV_F32)
    extract_kernel<float, float><<<grid, block, 0, st>>>((const float*)field, F, C, sx, sy, flags, slot, S, (float*)out);
  else if (field_dtype == DBV_F32 && out_dtype == DBV_F64)
    extract_kernel<float, double><<<grid, block, 0, st>>>((const float*)field, F, C, sx, sy, flags, slot, S, (double*)out);
  else
    return fail(DBV_ERR_INVALID, "dbv_extract: bad dtype %d -> %d", field_dtype, out_dtype);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

extern "C" int dbv_window_axpy(const void* in, void* out, int dtype, int64_t F, int C, const float* stamps,
                               const int32_t* x0, const int32_t* y0, int64_t N, int S, double alpha, void* stream) {
  DBV_REQUIRE(out, "dbv_window_axpy: null out");
  DBV_REQUIRE(N == 0 || (stamps && x0 && y0), "dbv_window_axpy: null stamp arrays");
  DBV_REQUIRE(F > 0 && C > 0 && S > 0 && N >= 0 && N < (1ll << 31), "dbv_window_axpy: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_r = (int)((F + AX_TR - 1) / AX_TR), tiles_c = (int)((F + AX_TC - 1) / AX_TC);
  dim3 grid((unsigned)(tiles_r * tiles_c)), block(AX_THREADS);
  if (dtype == DBV_F64)
    window_axpy_kernel<double><<<grid, block, 0, st>>>((const double*)in, (double*)out, F, C, stamps, x0, y0, (int)N, S, alpha, tiles_c);
  else if (dtype == DBV_F32)
    window_axpy_kernel<float><<<grid, block, 0, st>>>((const float*)in, (float*)out, F, C, stamps, x0, y0, (int)N, S, alpha, tiles_c);
  else
    return fail(DBV_ERR_INVALID, "dbv_window_axpy: bad dtype %d", dtype);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

extern "C" int dbv_center_mse(const void* cut, int cut_dtype, const float* mean, int64_t N, int S, int C, int lo, int hi,
                              double* out, void* stream) {
  DBV_REQUIRE(out && (N == 0 || (cut && mean)), "dbv_center_mse: null pointer");
  DBV_REQUIRE(0 <= lo && lo < hi && hi <= S, "dbv_center_mse: bad window [%d,%d) for S=%d", lo, hi, S);
  if (N == 0) return DBV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((N + 7) / 8)), block(256);
  if (cut_dtype == DBV_F64)
    center_mse_kernel<double><<<grid, block, 0, st>>>((const double*)cut, mean, N, S, C, lo, hi, out);
  else if (cut_dtype == DBV_F32)
    center_mse_kernel<float><<<grid, block, 0, st>>>((const float*)cut, mean, N, S, C, lo, hi, out);
  else
    return fail(DBV_ERR_INVALID, "dbv_center_mse: bad dtype %d", cut_dtype);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

extern "C" int64_t dbv_mse_scratch_bytes(void) { return (int64_t)MSE_BLOCKS * sizeof(double); }

extern "C" int dbv_mse(const void* a, const void* b, int dtype, int64_t n, double* out, void* scratch, int64_t scratch_bytes,
                       void* stream) {
  DBV_REQUIRE(a && b && out && scratch, "dbv_mse: null pointer");
  DBV_REQUIRE(n > 0, "dbv_mse: n must be positive");
  DBV_REQUIRE(scratch_bytes >= dbv_mse_scratch_bytes(), "dbv_mse: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  long long want = (n + 1023) / 1024;
  const int nb = (int)(want < MSE_BLOCKS ? (want < 1 ? 1 : want) : MSE_BLOCKS);
  if (dtype == DBV_F64)
    sqdiff_partial_kernel<double><<<nb, 256, 0, st>>>((const double*)a, (const double*)b, n, (double*)scratch);
  else if (dtype == DBV_F32)
    sqdiff_partial_kernel<float><<<nb, 256, 0, st>>>((const float*)a, (const float*)b, n, (double*)scratch);
  else
    return fail(DBV_ERR_INVALID, "dbv_mse: bad dtype %d", dtype);
  DBV_LAUNCH_CHECK();
  sqdiff_final_kernel<<<1, 256, 0, st>>>((const double*)scratch, nb, n, out);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}
