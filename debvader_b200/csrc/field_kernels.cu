// Field-level operators: stamp extraction (gather), deterministic windowed subtract / scatter-add,
// centre-window MSE and field MSE.  All are HBM-bound copy / read-modify-write kernels:
// coalesced, vectorised where alignment allows, no tensor cores.
//
//   dbv_extract      <- extract/extraction.py:21-36
//   dbv_window_axpy  <- deblend/field_deblender.py:46-97 (residual), :99-189 (predicted fields)
//   dbv_center_mse   <- deblend/field_deblender.py:323-332
//   dbv_mse          <- training/metrics.py:4-12
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cstdlib>
#include <mutex>

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
// extraction through shared memory with bulk asynchronous copies (the default when alignment allows):
// a CTA takes `rows_per_cta` rows of one stamp; every row is one cp.async.bulk global -> shared
// (row = S*C contiguous elements of the field, 2832 B for DC2/f64, 16-byte aligned because a pixel is
// 48 B), completion on one mbarrier.  Same-type output leaves as ONE bulk copy shared -> global (the
// rows are contiguous in the stamp); the fused f64 -> f32 cast reads the staged rows and stores
// 8-byte pairs.  No per-element address arithmetic, ~42 KB in flight per CTA, 5 CTAs per SM.
// ---------------------------------------------------------------------------------------------
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(128) extract_bulk_kernel(const Tin* __restrict__ field, long long F, int C,
                                                           const int32_t* __restrict__ sx, const int32_t* __restrict__ sy,
                                                           const uint8_t* __restrict__ flags, const int64_t* __restrict__ slot,
                                                           int S, Tout* __restrict__ out, int rows_per_cta) {
  extern __shared__ __align__(128) unsigned char ex_smem[];
  __shared__ __align__(8) uint64_t ex_bar;
  const long long k = blockIdx.x;
  const int r0 = blockIdx.y * rows_per_cta;
  const int nr = min(rows_per_cta, S - r0);
  const int fl = flags ? flags[k] : 0;
  const int x0 = sx[k], y0 = sy[k];
  const long long dst = slot ? slot[k] : k;
  const int L = S * C;
  Tout* __restrict__ o = out + (dst * S + r0) * (long long)L;
  if (fl != 0) {
    // a length-1 source axis broadcast over the stamp (numpy assignment): rare, element-wise
    const int bx = fl & 1, by = (fl >> 1) & 1;
    for (int idx = threadIdx.x; idx < nr * L; idx += blockDim.x) {
      const int r = idx / L;
      const int rem = idx - r * L;
      const int c = rem / C, ch = rem - c * C;
      const long long xr = x0 + (bx ? 0 : r0 + r), yc = y0 + (by ? 0 : c);
      o[idx] = cvt<Tout, Tin>(__ldg(field + (xr * F + yc) * C + ch));
    }
    return;
  }
  const uint32_t bar = smem_u32(&ex_bar);
  const uint32_t row_bytes = (uint32_t)L * sizeof(Tin);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) mbar_expect_tx(bar, (uint32_t)nr * row_bytes);
    __syncwarp();
    const Tin* base = field + ((long long)(x0 + r0) * F + y0) * C;
    for (int r = threadIdx.x; r < nr; r += 32)
      bulk_load(smem_u32(ex_smem) + (uint32_t)r * row_bytes, base + (long long)r * F * C, row_bytes, bar);
  }
  mbar_wait(bar, 0);
  if constexpr (sizeof(Tin) == sizeof(Tout)) {
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(o), "r"(smem_u32(ex_smem)),
                   "r"((uint32_t)nr * row_bytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the read
    }
  } else {
    using VI = typename Vec2<Tin>::type;
    using VO = typename Vec2<Tout>::type;
    const VI* __restrict__ sv = reinterpret_cast<const VI*>(ex_smem);
    VO* __restrict__ ov = reinterpret_cast<VO*>(o);
    const int total = nr * (L >> 1);
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const VI v = sv[i];
      VO w;
      w.x = cvt<Tout, Tin>(v.x);
      w.y = cvt<Tout, Tin>(v.y);
      __stcs(ov + i, w);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// windowed axpy, owner-computes: a CTA owns a TR x TC pixel tile of the field, collects (in
// ascending stamp index) the stamps that overlap it, and applies them to each of its pixels in
// that order.  No atomics on the data; one rounding per addition; bit-identical to the sequential
// host loop.  The per-tile stamp lists come from a binning pass (one thread per stamp appends its
// index to the <= AX_LCAP-entry list of every tile it touches; the tile CTA sorts its list, so the
// order of the appends does not matter); a tile whose list overflowed scans all N stamps instead.
// Field accesses are 16-byte vectors, 4 independent loads in flight per thread.
// ---------------------------------------------------------------------------------------------
constexpr int AX_TR = 32, AX_TC = 64, AX_CAP = 768, AX_THREADS = 256, AX_LCAP = 32, AX_U = 4;

template <typename T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }

__global__ void __launch_bounds__(256) axpy_bin_kernel(const int32_t* __restrict__ x0, const int32_t* __restrict__ y0, int N, int S,
                                                       long long FH, long long FW, int tiles_c, int* __restrict__ cnt, int4* __restrict__ list) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const long long xi = x0[i], yi = y0[i];
  if (xi + S <= 0 || xi >= FH || yi + S <= 0 || yi >= FW) return;
  const int r_lo = (int)((xi > 0 ? xi : 0) / AX_TR), r_hi = (int)((xi + S - 1 < FH - 1 ? xi + S - 1 : FH - 1) / AX_TR);
  const int c_lo = (int)((yi > 0 ? yi : 0) / AX_TC), c_hi = (int)((yi + S - 1 < FW - 1 ? yi + S - 1 : FW - 1) / AX_TC);
  for (int r = r_lo; r <= r_hi; ++r)
    for (int c = c_lo; c <= c_hi; ++c) {
      const int t = r * tiles_c + c;
      const int p = atomicAdd(cnt + t, 1);
      if (p < AX_LCAP) list[(long long)t * AX_LCAP + p] = make_int4(i, (int)xi, (int)yi, 0);  // the tile needs no second look-up
    }
}

// Sorted copy of a tile's binned list (<= AX_LCAP = 32 records (stamp, x, y), appended in arbitrary order) into the shared
// lists: warp 0 alone, one record per lane, rank by shuffles (the indices are distinct) — no barrier, no second global
// look-up (the records carry the positions).  The caller's next __syncthreads() publishes it.
static_assert(AX_LCAP == 32, "one binned record per lane of warp 0");
__device__ __forceinline__ int4 ax_load_bin_record(const int4* __restrict__ bin_list) {
  // issued BEFORE the tile's count is known (one global round trip instead of two); slots past the count hold stale bytes
  // of the caller's scratch and are masked out by ax_sorted_bin
  return (bin_list && threadIdx.x < 32) ? __ldg(bin_list + (long long)blockIdx.x * AX_LCAP + threadIdx.x) : make_int4(0, 0, 0, 0);
}
__device__ __forceinline__ void ax_sorted_bin(int binned, int4 e, int* s_id, int* s_x, int* s_y) {
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  if (lane >= binned) e.x = 0x7fffffff;
  int rank = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) rank += __shfl_sync(0xffffffffu, e.x, j) < e.x;
  if (lane < binned) {
    s_id[rank] = e.x;
    s_x[rank] = e.y;
    s_y[rank] = e.z;
  }
}

template <typename T, typename TS>
__global__ void __launch_bounds__(AX_THREADS) window_axpy_kernel(const T* in, T* out, long long FH, long long F, int C,
                                                                  const TS* __restrict__ stamps,
                                                                  const int32_t* __restrict__ x0, const int32_t* __restrict__ y0,
                                                                  int N, int S, double alpha_d, int tiles_c,
                                                                  const int* __restrict__ bin_cnt, const int4* __restrict__ bin_list,
                                                                  int planar, int inplace) {
  // F = columns of the (FH, F, C) field = its row pitch in pixels.  inplace (in == out): only the elements a stamp
  // actually covers are read and written back, so the traffic is the algorithmic one (window read-modify-write +
  // stamp), not a pass over the whole field; a tile no stamp touches returns at once.
  __shared__ int s_id[AX_CAP], s_x[AX_CAP], s_y[AX_CAP];
  __shared__ int s_wcnt[AX_THREADS / 32];
  __shared__ int s_count, s_next;
  const int tr0 = (blockIdx.x / tiles_c) * AX_TR;
  const int tc0 = (blockIdx.x % tiles_c) * AX_TC;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const T alpha = (T)alpha_d;
  const int nelt = AX_TR * AX_TC * C;
  const long long stamp_sz = (long long)S * S * C;
  const int4 rec = ax_load_bin_record(bin_list);
  const int binned = bin_cnt ? bin_cnt[blockIdx.x] : -1;
  if (inplace && binned == 0) return;
  int start = 0;
  bool first = true;
  do {
    if (binned >= 0 && binned <= AX_LCAP) {
      // sorted copy of the binned list (rank sort: the indices are distinct)
      ax_sorted_bin(binned, rec, s_id, s_x, s_y);
      if (threadIdx.x == 0) { s_count = binned; s_next = N; }
      __syncthreads();
    } else {
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
    }
    const int cnt = s_count;
    const bool work = cnt > 0 || (first && !inplace);
    if (work && (C & 1) == 0) {
      // vector path: 2 bands per thread (16-byte field accesses for f64); a pixel's C values start at an even
      // element index, so the pairs never straddle pixels
      using V2 = typename Vec2<T>::type;
      using VS = typename Vec2<TS>::type;
      const int rowv = AX_TC * C / 2;  // vectors per tile row
      const int nvec = AX_TR * rowv;
      for (int e0 = threadIdx.x; e0 < nvec; e0 += AX_U * AX_THREADS) {
        V2 acc[AX_U];
        long long idx[AX_U];
        int X[AX_U], Y[AX_U], ch[AX_U];
        bool ok[AX_U];
#pragma unroll
        for (int j = 0; j < AX_U; ++j) {
          const int e = e0 + j * AX_THREADS;
          const int pr = e / rowv;
          const int rem = e - pr * rowv;
          X[j] = tr0 + pr;
          Y[j] = tc0 + (2 * rem) / C;
          ch[j] = 2 * rem - ((2 * rem) / C) * C;
          ok[j] = e < nvec && X[j] < FH && Y[j] < F;
          idx[j] = ((long long)X[j] * F + Y[j]) * C + ch[j];
          acc[j].x = (T)0;
          acc[j].y = (T)0;
        }
        if (inplace) {  // untouched elements are neither read nor written (kept out of the load loop below, so that its AX_U
                        // independent loads are still issued back to back)
          bool any[AX_U];
#pragma unroll
          for (int j = 0; j < AX_U; ++j) any[j] = false;
          for (int k = 0; k < cnt; ++k) {
            const int sx = s_x[k], sy = s_y[k];
#pragma unroll
            for (int j = 0; j < AX_U; ++j) any[j] = any[j] || ((unsigned)(X[j] - sx) < (unsigned)S && (unsigned)(Y[j] - sy) < (unsigned)S);
          }
#pragma unroll
          for (int j = 0; j < AX_U; ++j) ok[j] = ok[j] && any[j];
        }
#pragma unroll
        for (int j = 0; j < AX_U; ++j) {
          if (ok[j]) {
            if (!first) acc[j] = *reinterpret_cast<const V2*>(out + idx[j]);
            else if (in) acc[j] = *reinterpret_cast<const V2*>(in + idx[j]);
          }
        }
        // stamps in groups of KU: all the (independent) stamp loads of a group are issued before its additions, which
        // run in ascending k for every accumulator (the order of the sequential host loop)
        constexpr int KU = 2;  // 4 costs 124 registers (2 CTAs per SM): measured slower, also for the plain copy tiles
        for (int k0 = 0; k0 < cnt; k0 += KU) {
          VS v[KU][AX_U];
          bool hit[KU][AX_U];
#pragma unroll
          for (int kk = 0; kk < KU; ++kk) {
            const int k = k0 + kk;
            const bool live = k < cnt;
            const int sx = live ? s_x[k] : 0, sy = live ? s_y[k] : 0;
            const long long sbase = live ? s_id[k] * stamp_sz : 0;
#pragma unroll
            for (int j = 0; j < AX_U; ++j) {
              const int dx = X[j] - sx, dy = Y[j] - sy;
              hit[kk][j] = live && ok[j] && (unsigned)dx < (unsigned)S && (unsigned)dy < (unsigned)S;
              v[kk][j].x = (TS)0;
              v[kk][j].y = (TS)0;
              if (hit[kk][j]) {
                if (planar) {  // (N, C, S, S): the two bands of the pair live in different planes
                  const TS* sp = stamps + sbase + ((long long)ch[j] * S + dx) * S + dy;
                  v[kk][j].x = __ldg(sp);
                  v[kk][j].y = __ldg(sp + (long long)S * S);
                } else {
                  v[kk][j] = __ldg(reinterpret_cast<const VS*>(stamps + sbase + ((long long)dx * S + dy) * C + ch[j]));
                }
              }
            }
          }
#pragma unroll
          for (int kk = 0; kk < KU; ++kk)
#pragma unroll
            for (int j = 0; j < AX_U; ++j)
              if (hit[kk][j]) {
                acc[j].x = add_rn<T>(acc[j].x, mul_rn<T>(alpha, (T)v[kk][j].x));
                acc[j].y = add_rn<T>(acc[j].y, mul_rn<T>(alpha, (T)v[kk][j].y));
              }
        }
#pragma unroll
        for (int j = 0; j < AX_U; ++j)
          if (ok[j]) *reinterpret_cast<V2*>(out + idx[j]) = acc[j];
      }
    } else if (work) {
      for (int e = threadIdx.x; e < nelt; e += AX_THREADS) {
        const int ch = e % C;
        const int pc = (e / C) % AX_TC;
        const int pr = e / (C * AX_TC);
        const int X = tr0 + pr, Y = tc0 + pc;
        if (X >= FH || Y >= F) continue;
        const long long idx = ((long long)X * F + Y) * C + ch;
        T acc = first ? (in ? in[idx] : (T)0) : out[idx];
        for (int k = 0; k < cnt; ++k) {
          const int dx = X - s_x[k], dy = Y - s_y[k];
          if ((unsigned)dx < (unsigned)S && (unsigned)dy < (unsigned)S) {
            const TS v = planar ? __ldg(stamps + s_id[k] * stamp_sz + ((long long)ch * S + dx) * S + dy)
                                : __ldg(stamps + s_id[k] * stamp_sz + ((long long)dx * S + dy) * C + ch);
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

// The same operator for pixel-interleaved stamps and a compile-time band count (C even), laid out so that nothing is
// computed per element: a WARP owns a tile row (X is warp-uniform, so "does stamp k reach this row" is a uniform branch
// and a stamp that misses the row costs one comparison per warp), and lane v of it owns the 16-byte vectors v, v + 32, ...
// of the row's AX_TC * C contiguous values — the field offset (row base + 2v) and the stamp offset
// ((X - sx) * S + tc0 - sy) * C + 2v are both affine in v, every access is a fully coalesced 512-byte warp transaction,
// and all NI field loads of a row are in flight together.  ~64 registers: 4 CTAs per SM (the generic kernel: 104 -> 2).
// INPLACE: only the column span of the stamps reaching a row is read and written (a first, warp-uniform walk over the list).
// Stamps are applied in ascending index with one rounding per addition: bit-identical to the sequential host loop.
template <typename T, typename TS, int C, bool INPLACE>
__global__ void __launch_bounds__(AX_THREADS, INPLACE ? 6 : 4) window_axpy_rows_kernel(const T* in, T* out, long long FH, long long F,
                                                                         const TS* __restrict__ stamps, const int32_t* __restrict__ x0,
                                                                         const int32_t* __restrict__ y0, int N, int S, double alpha_d, int tiles_c,
                                                                         const int* __restrict__ bin_cnt, const int4* __restrict__ bin_list) {
  static_assert((C & 1) == 0 && (AX_TC * C / 2) % 32 == 0, "a tile row is a whole number of 32-vector groups");
  constexpr int RV = AX_TC * C / 2;  // 16-byte (f64) / 8-byte (f32) vectors per tile row
  // INPLACE is latency bound (a row = one round trip to DRAM) and its throughput is the number of rows in flight = warps per
  // SM: there a warp takes HALF a row (3 vectors per lane instead of 6: ~40 registers, 6 CTAs = 48 warps per SM instead of 32)
  constexpr int CS = INPLACE ? 2 : 1;  // warps per tile row
  static_assert(RV % (32 * CS) == 0, "a warp's share of a row is a whole number of 32-vector groups");
  constexpr int NI = RV / CS / 32;     // vectors per lane and row
  using V2 = typename Vec2<T>::type;
  using VS = typename Vec2<TS>::type;
  __shared__ int s_id[AX_CAP], s_x[AX_CAP], s_y[AX_CAP];
  __shared__ int s_wcnt[AX_THREADS / 32];
  __shared__ int s_count, s_next;
  const int tr0 = (blockIdx.x / tiles_c) * AX_TR;
  const int tc0 = (blockIdx.x % tiles_c) * AX_TC;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const T alpha = (T)alpha_d;
  const long long stamp_sz = (long long)S * S * C;
  const int4 rec = ax_load_bin_record(bin_list);
  const int binned = bin_cnt ? bin_cnt[blockIdx.x] : -1;
  if (INPLACE && binned == 0) return;
  // this lane's columns: vector v = lane + 32 i covers values 2v, 2v + 1 of the row = band pair (2v % C) of pixel tc0 + 2v / C
  const int half = wid % CS;                    // which share of its rows this warp takes
  const int vbase = half * (RV / CS) + lane;    // this lane's first vector of a row
  int Ycol[NI];
  bool okc[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    Ycol[i] = tc0 + (2 * (vbase + 32 * i)) / C;
    okc[i] = Ycol[i] < F;
  }
  int start = 0;
  bool first = true;
  do {
    if (binned >= 0 && binned <= AX_LCAP) {
      ax_sorted_bin(binned, rec, s_id, s_x, s_y);
      if (threadIdx.x == 0) { s_count = binned; s_next = N; }
      __syncthreads();
    } else {
      if (threadIdx.x == 0) { s_count = 0; s_next = N; }
      __syncthreads();
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
    }
    const int cnt = s_count;
    if (cnt > 0 || (first && !INPLACE)) {
      if (INPLACE) {
        // The in-place form is latency bound (each row of a warp waits for its field and stamp segments in turn, and the
        // registers that hold loaded values limit how many rows can be in flight): pull every segment this warp is going
        // to touch into L2 first — prefetches hold no registers, one 128-byte line per lane.
        for (int pr = wid / CS; pr < AX_TR && half == 0; pr += AX_THREADS / 32 / CS) {  // one warp per row issues the prefetches
          const int X = tr0 + pr;
          if (X >= FH) break;
          for (int k = 0; k < cnt; ++k) {
            const int dx = X - s_x[k];
            if ((unsigned)dx >= (unsigned)S) continue;  // warp-uniform
            const int sy = s_y[k];
            const int y_lo = max(tc0, sy), y_hi = min(min(tc0 + AX_TC, sy + S), (int)F);
            if (y_hi <= y_lo) continue;
            const char* fp = reinterpret_cast<const char*>(out + ((long long)X * F + y_lo) * C);
            const char* sp = reinterpret_cast<const char*>(stamps + s_id[k] * stamp_sz + ((long long)dx * S + (y_lo - sy)) * C);
            const int fb = (y_hi - y_lo) * C * (int)sizeof(T), sb = (y_hi - y_lo) * C * (int)sizeof(TS);
            const int fo = (int)(reinterpret_cast<uintptr_t>(fp) & 127), so = (int)(reinterpret_cast<uintptr_t>(sp) & 127);
            for (int o = lane * 128; o < fb + fo; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(fp - fo + o));
            for (int o = lane * 128; o < sb + so; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp - so + o));
          }
        }
      }
      for (int pr = wid / CS; pr < AX_TR; pr += AX_THREADS / 32 / CS) {
        const int X = tr0 + pr;
        if (X >= FH) break;
        const long long rowbase = ((long long)X * F + tc0) * C + 2 * vbase;
        unsigned touched = ~0u;
        if (INPLACE) {
          // column span [lo, hi) of the stamps reaching this row (warp-uniform, two operations per stamp): elements inside it
          // are read and written back — a gap between two stamps of the same row is rewritten with the value just read,
          // harmless because every element has exactly one owner thread
          int lo = 0x7fffffff, hi = -0x7fffffff;
          for (int k = 0; k < cnt; ++k) {
            if ((unsigned)(X - s_x[k]) >= (unsigned)S) continue;
            const int sy = s_y[k];
            lo = min(lo, sy);
            hi = max(hi, sy + S);
          }
          if (hi <= lo) continue;  // no stamp reaches this row
          touched = 0u;
#pragma unroll
          for (int i = 0; i < NI; ++i) touched |= ((Ycol[i] >= lo && Ycol[i] < hi) ? 1u : 0u) << i;
        }
        V2 acc[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          acc[i].x = (T)0;
          acc[i].y = (T)0;
          if (okc[i] && ((touched >> i) & 1u)) {
            if (!first || INPLACE) acc[i] = *reinterpret_cast<const V2*>(out + rowbase + 64 * i);
            else if (in) acc[i] = *reinterpret_cast<const V2*>(in + rowbase + 64 * i);
          }
        }
        for (int k = 0; k < cnt; ++k) {
          const int dx = X - s_x[k];
          if ((unsigned)dx >= (unsigned)S) continue;  // warp-uniform
          const int sy = s_y[k];
          const TS* sp = stamps + s_id[k] * stamp_sz + ((long long)dx * S + (tc0 - sy)) * C + 2 * vbase;
          VS v[NI];
          bool hit[NI];
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            hit[i] = okc[i] && (unsigned)(Ycol[i] - sy) < (unsigned)S;
            v[i].x = (TS)0;
            v[i].y = (TS)0;
            if (hit[i]) v[i] = __ldg(reinterpret_cast<const VS*>(sp + 64 * i));
          }
#pragma unroll
          for (int i = 0; i < NI; ++i)
            if (hit[i]) {
              acc[i].x = add_rn<T>(acc[i].x, mul_rn<T>(alpha, (T)v[i].x));
              acc[i].y = add_rn<T>(acc[i].y, mul_rn<T>(alpha, (T)v[i].y));
            }
        }
#pragma unroll
        for (int i = 0; i < NI; ++i)
          if (okc[i] && ((touched >> i) & 1u)) *reinterpret_cast<V2*>(out + rowbase + 64 * i) = acc[i];
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
template <typename T>
__global__ void __launch_bounds__(256) sqdiff_rect_kernel(const T* __restrict__ a, const T* __restrict__ b, long long rows, long long cols,
                                                          long long pitch_a, long long pitch_b, double* __restrict__ partial) {
  // one row per CTA iteration (rows in grid stride, columns in thread stride): sub-rectangles of two pitched arrays
  __shared__ double s[8];
  double acc = 0.0;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const T* pa = a + r * pitch_a;
    const T* pb = b + r * pitch_b;
    for (long long c = threadIdx.x; c < cols; c += blockDim.x) {
      const double d = (double)__ldg(pa + c) - (double)__ldg(pb + c);
      acc += d * d;
    }
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
  const size_t isz = field_dtype == DBV_F64 ? 8 : 4, osz = out_dtype == DBV_F64 ? 8 : 4;
  const size_t row_bytes = (size_t)S * C * isz;
  // bulk-copy path: every source row and (same-type) destination block must be 16-byte aligned
  bool bulk = ((size_t)C * isz) % 16 == 0 && ((uintptr_t)field % 16) == 0 && row_bytes <= 48 * 1024 && ((S * C) & 1) == 0;
  if (isz == osz) bulk = bulk && ((uintptr_t)out % 16) == 0;  // S*S*C*size and r0*row_bytes are multiples of 16 given the row is
  else bulk = bulk && ((uintptr_t)out % 8) == 0;
  if (const char* e = dbv_env("DBV_EXTRACT_BULK")) bulk = bulk && atoi(e) != 0;  // tuning knob (0 = vector kernel)
  if (bulk) {
    int rows = (int)((44 * 1024) / row_bytes);  // <= 44 KB of shared memory per CTA: 5 CTAs per SM
    if (rows > S) rows = S;
    if (rows < 1) rows = 1;
    if (const char* e = dbv_env("DBV_EXTRACT_ROWS")) rows = atoi(e) > 0 && (size_t)atoi(e) * row_bytes <= 48 * 1024 ? atoi(e) : rows;
    dim3 grid((unsigned)N, (unsigned)((S + rows - 1) / rows)), block(128);
    const size_t smem = (size_t)rows * row_bytes;
    if (field_dtype == DBV_F64 && out_dtype == DBV_F64)
      extract_bulk_kernel<double, double><<<grid, block, smem, st>>>((const double*)field, F, C, sx, sy, flags, slot, S, (double*)out, rows);
    else if (field_dtype == DBV_F64 && out_dtype == DBV_F32)
      extract_bulk_kernel<double, float><<<grid, block, smem, st>>>((const double*)field, F, C, sx, sy, flags, slot, S, (float*)out, rows);
    else if (field_dtype == DBV_F32 && out_dtype == DBV_F32)
      extract_bulk_kernel<float, float><<<grid, block, smem, st>>>((const float*)field, F, C, sx, sy, flags, slot, S, (float*)out, rows);
    else if (field_dtype == DBV_F32 && out_dtype == DBV_F64)
      extract_bulk_kernel<float, double><<<grid, block, smem, st>>>((const float*)field, F, C, sx, sy, flags, slot, S, (double*)out, rows);
    else
      return fail(DBV_ERR_INVALID, "dbv_extract: bad dtype %d -> %d", field_dtype, out_dtype);
    DBV_LAUNCH_CHECK();
    return DBV_OK;
  }
  // enough CTAs per stamp to keep >= 4 waves of 148 SMs busy for small N, 1-2 for large N
  int per = 2;
  if (N < 2048) per = 4;
  if (N < 256) per = 8;
  if (const char* e = dbv_env("DBV_EXTRACT_PER")) per = atoi(e) > 0 ? atoi(e) : per;  // tuning knob
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

static inline int64_t ax_list_offset(int64_t ntiles) { return (ntiles + 3) / 4 * 4; }  // ints before the 16-byte-aligned record lists

// out = in + alpha * sum_k paste(stamps[k]) on a (FH, FW, C) field; `bins` = caller scratch of
// dbv_window_axpy_scratch_bytes(FH, FW) bytes (tile counts + fixed-capacity lists), or NULL (every tile scans all stamps)
static int window_axpy_impl(const void* in, void* out, int dtype, int64_t FH, int64_t FW, int C, const void* stamps, int stamp_dtype,
                            int stamp_planar, const int32_t* x0, const int32_t* y0, int64_t N, int S, double alpha, int* bins,
                            cudaStream_t st) {
  const int tiles_r = (int)((FH + AX_TR - 1) / AX_TR), tiles_c = (int)((FW + AX_TC - 1) / AX_TC);
  const int ntiles = tiles_r * tiles_c;
  const int inplace = (in == out) ? 1 : 0;
  if (N == 0 && inplace) return DBV_OK;
  if (N == 0) bins = nullptr;
  if (bins && (reinterpret_cast<uintptr_t>(bins) & 15)) return fail(DBV_ERR_INVALID, "window_axpy: the binning scratch must be 16-byte aligned");
  if (bins) {
    DBV_CUDA(cudaMemsetAsync(bins, 0, (size_t)ntiles * sizeof(int), st));
    axpy_bin_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(x0, y0, (int)N, S, FH, FW, tiles_c, bins, reinterpret_cast<int4*>(bins + ax_list_offset(ntiles)));
    DBV_LAUNCH_CHECK();
  }
  const int* bc = bins;
  const int4* bl = bins ? reinterpret_cast<const int4*>(bins + ax_list_offset(ntiles)) : nullptr;
  dim3 grid((unsigned)ntiles), block(AX_THREADS);
  if (C == 6 && !stamp_planar && !dbv_env("DBV_AXPY_GENERIC")) {  // the DC2 band count, pixel-interleaved stamps: warp-per-row kernel
#define DBV_AXPY_ROWS(T, TS)                                                                                                              \
  do {                                                                                                                                     \
    if (inplace) window_axpy_rows_kernel<T, TS, 6, true><<<grid, block, 0, st>>>((const T*)in, (T*)out, FH, FW, (const TS*)stamps, x0, y0, (int)N, S, alpha, tiles_c, bc, bl); \
    else window_axpy_rows_kernel<T, TS, 6, false><<<grid, block, 0, st>>>((const T*)in, (T*)out, FH, FW, (const TS*)stamps, x0, y0, (int)N, S, alpha, tiles_c, bc, bl);       \
  } while (0)
    if (dtype == DBV_F64 && stamp_dtype == DBV_F32) DBV_AXPY_ROWS(double, float);
    else if (dtype == DBV_F32 && stamp_dtype == DBV_F32) DBV_AXPY_ROWS(float, float);
    else if (dtype == DBV_F64 && stamp_dtype == DBV_F64) DBV_AXPY_ROWS(double, double);
    else DBV_AXPY_ROWS(float, double);
#undef DBV_AXPY_ROWS
    DBV_LAUNCH_CHECK();
    return DBV_OK;
  }
  if (dtype == DBV_F64 && stamp_dtype == DBV_F32)
    window_axpy_kernel<double, float><<<grid, block, 0, st>>>((const double*)in, (double*)out, FH, FW, C, (const float*)stamps, x0, y0, (int)N, S, alpha, tiles_c, bc, bl, stamp_planar != 0, inplace);
  else if (dtype == DBV_F32 && stamp_dtype == DBV_F32)
    window_axpy_kernel<float, float><<<grid, block, 0, st>>>((const float*)in, (float*)out, FH, FW, C, (const float*)stamps, x0, y0, (int)N, S, alpha, tiles_c, bc, bl, stamp_planar != 0, inplace);
  else if (dtype == DBV_F64 && stamp_dtype == DBV_F64)
    window_axpy_kernel<double, double><<<grid, block, 0, st>>>((const double*)in, (double*)out, FH, FW, C, (const double*)stamps, x0, y0, (int)N, S, alpha, tiles_c, bc, bl, stamp_planar != 0, inplace);
  else
    window_axpy_kernel<float, double><<<grid, block, 0, st>>>((const float*)in, (float*)out, FH, FW, C, (const double*)stamps, x0, y0, (int)N, S, alpha, tiles_c, bc, bl, stamp_planar != 0, inplace);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

extern "C" int64_t dbv_window_axpy_scratch_bytes(int64_t FH, int64_t FW) {
  if (FH <= 0 || FW <= 0) return DBV_ERR_INVALID;
  const int64_t ntiles = ((FH + AX_TR - 1) / AX_TR) * ((FW + AX_TC - 1) / AX_TC);
  return (ax_list_offset(ntiles) + ntiles * AX_LCAP * 4) * (int64_t)sizeof(int);  // counts (padded to 16 bytes) + 16-byte records
}

extern "C" int dbv_window_axpy_rect(const void* in, void* out, int dtype, int64_t FH, int64_t FW, int C, const void* stamps,
                                    int stamp_dtype, int stamp_planar, const int32_t* x0, const int32_t* y0, int64_t N, int S,
                                    double alpha, void* scratch, int64_t scratch_bytes, void* stream) {
  DBV_REQUIRE(out, "dbv_window_axpy_rect: null out");
  DBV_REQUIRE(N == 0 || (stamps && x0 && y0), "dbv_window_axpy_rect: null stamp arrays");
  DBV_REQUIRE(FH > 0 && FW > 0 && C > 0 && S > 0 && N >= 0 && N < (1ll << 31), "dbv_window_axpy_rect: bad sizes");
  DBV_REQUIRE(FH * FW * C < (1ll << 40), "dbv_window_axpy_rect: field too large");
  DBV_REQUIRE(dtype == DBV_F64 || dtype == DBV_F32, "dbv_window_axpy_rect: bad dtype %d", dtype);
  DBV_REQUIRE(stamp_dtype == DBV_F64 || stamp_dtype == DBV_F32, "dbv_window_axpy_rect: bad stamp dtype %d", stamp_dtype);
  DBV_REQUIRE(scratch == nullptr || scratch_bytes >= dbv_window_axpy_scratch_bytes(FH, FW), "dbv_window_axpy_rect: scratch too small");
  return window_axpy_impl(in, out, dtype, FH, FW, C, stamps, stamp_dtype, stamp_planar, x0, y0, N, S, alpha, (int*)scratch,
                          (cudaStream_t)stream);
}

// square field, library-allocated scratch: taken from the device's stream-ordered pool on the caller's stream and
// returned to it right after the launch, so concurrent calls on different streams never share it
extern "C" int dbv_window_axpy_ex(const void* in, void* out, int dtype, int64_t F, int C, const void* stamps, int stamp_dtype,
                                  int stamp_planar, const int32_t* x0, const int32_t* y0, int64_t N, int S, double alpha, void* stream) {
  DBV_REQUIRE(out, "dbv_window_axpy: null out");
  DBV_REQUIRE(N == 0 || (stamps && x0 && y0), "dbv_window_axpy: null stamp arrays");
  DBV_REQUIRE(F > 0 && C > 0 && S > 0 && N >= 0 && N < (1ll << 31), "dbv_window_axpy: bad sizes");
  DBV_REQUIRE(dtype == DBV_F64 || dtype == DBV_F32, "dbv_window_axpy: bad dtype %d", dtype);
  DBV_REQUIRE(stamp_dtype == DBV_F64 || stamp_dtype == DBV_F32, "dbv_window_axpy: bad stamp dtype %d", stamp_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  void* bins = nullptr;
  if (N > 0) DBV_CUDA(cudaMallocAsync(&bins, (size_t)dbv_window_axpy_scratch_bytes(F, F), st));
  const int r = window_axpy_impl(in, out, dtype, F, F, C, stamps, stamp_dtype, stamp_planar, x0, y0, N, S, alpha, (int*)bins, st);
  if (bins) cudaFreeAsync(bins, st);
  return r;
}

extern "C" int dbv_window_axpy(const void* in, void* out, int dtype, int64_t F, int C, const float* stamps,
                               const int32_t* x0, const int32_t* y0, int64_t N, int S, double alpha, void* stream) {
  return dbv_window_axpy_ex(in, out, dtype, F, C, stamps, DBV_F32, 0, x0, y0, N, S, alpha, stream);
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

// sum of squared differences over a rows x cols sub-rectangle of two pitched arrays (the owner tile of a rank's local
// region): the partial sum a tiled field MSE all-reduces.  out[0] = sum, fixed reduction order.
extern "C" int dbv_sqdiff_sum_rect(const void* a, const void* b, int dtype, int64_t rows, int64_t cols, int64_t pitch_a,
                                   int64_t pitch_b, double* out, void* scratch, int64_t scratch_bytes, void* stream) {
  DBV_REQUIRE(a && b && out && scratch, "dbv_sqdiff_sum_rect: null pointer");
  DBV_REQUIRE(rows > 0 && cols > 0 && pitch_a >= cols && pitch_b >= cols, "dbv_sqdiff_sum_rect: bad sizes");
  DBV_REQUIRE(scratch_bytes >= dbv_mse_scratch_bytes(), "dbv_sqdiff_sum_rect: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (int)(rows < MSE_BLOCKS ? rows : MSE_BLOCKS);
  if (dtype == DBV_F64)
    sqdiff_rect_kernel<double><<<nb, 256, 0, st>>>((const double*)a, (const double*)b, rows, cols, pitch_a, pitch_b, (double*)scratch);
  else if (dtype == DBV_F32)
    sqdiff_rect_kernel<float><<<nb, 256, 0, st>>>((const float*)a, (const float*)b, rows, cols, pitch_a, pitch_b, (double*)scratch);
  else
    return fail(DBV_ERR_INVALID, "dbv_sqdiff_sum_rect: bad dtype %d", dtype);
  DBV_LAUNCH_CHECK();
  sqdiff_final_kernel<<<1, 256, 0, st>>>((const double*)scratch, nb, 1, out);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

namespace dbv {

// ---------------------------------------------------------------------------------------------
// Sub-pixel placement: scipy.ndimage.shift(canvas, (x_pos, y_pos)) of deblend/field_deblender.py:92-95
// (order 3, mode 'constant', prefilter) restricted to the window where the result is not negligible.
// The canvas is zero but for one block of data (the stamp), the cubic B-spline prefilter has one pole
// z = sqrt(3)-2 and its response decays as |z|^d, so per axis only the line segment
// [origin - P, origin + S + P) clipped to the canvas is filtered (P = 28: |z|^28 = 1e-16): zero causal
// state before it, exact geometric tail after it, and scipy's mirror initialisation at an end that IS
// the canvas edge — a canvas not larger than the segment is therefore filtered whole, exactly as scipy
// does.  The 2-D operation is separable: pass X filters and interpolates every data column along the
// rows, pass Y does the same along the columns.  One thread owns one line, which lives in its local
// memory (all threads of a warp touch the same element index, so the accesses coalesce); fp64.
// ---------------------------------------------------------------------------------------------
struct SplineGeom {
  long long F;  // canvas (field) size
  int S;        // data samples per axis
  int P;        // margin kept around the data
  int origin;   // canvas coordinate of data sample 0 (when no per-stamp origin array is given)
  int n_out;    // outputs per axis = S + 2P + 2, output a of stamp k <-> canvas index anchor[k] + a
  // by-value parameters of a single item (used when the per-item arrays are NULL): the position fit
  double pos_x1, pos_y1;
  int origin_x1, origin_y1, ax1, ay1;
};

// the line segment of one (stamp, axis): canvas range [lo, hi), filtered in place
struct SplineLine {
  int lo, hi;
};

// line element i of this thread lives at line[i * LS] (LS = threads per CTA: consecutive threads, consecutive words)
#define SPL_AT(i) line[(size_t)(i) * LS]
constexpr int SPL_THREADS = 64, SPL_LD = 16;

__device__ __forceinline__ void spline_prefilter_line(double* line, const int LS, const SplineLine& sl, long long F) {
  const double z = -0.26794919243112270647;  // sqrt(3) - 2
  const int n = sl.hi - sl.lo;
  if (n < 2) return;  // scipy leaves lines shorter than 2 untouched (n == F == 1)
  if (sl.lo == 0) {
    // ni_splines.c:_init_causal_mirror over the canvas line of length F (zero outside the segment)
    const double z_n_1 = pow(z, (double)(F - 1));
    const bool tail = sl.hi == F;  // canvas index F-1-i lies inside the segment only then
    double c0 = SPL_AT(0) + (tail ? z_n_1 * SPL_AT(n - 1) : 0.0);
    double z_i = z;
    const int last = (long long)n - 1 < F - 2 ? n - 1 : (int)(F - 2);
    for (int i = 1; i <= last; ++i) {
      const long long j = F - 1 - i - sl.lo;
      c0 = c0 + z_i * (SPL_AT(i) + ((tail && j >= 0 && j < n) ? z_n_1 * SPL_AT(j) : 0.0));
      z_i *= z;
    }
    SPL_AT(0) = c0 / (1.0 - z_n_1 * z_n_1);
  }
  double prev = SPL_AT(0);
  for (int i = 1; i < n; ++i) {
    prev = SPL_AT(i) + z * prev;
    SPL_AT(i) = prev;
  }
  if (sl.hi == F) prev = (z * SPL_AT(n - 2) + prev) * z / (z * z - 1.0);  // _init_anticausal_mirror
  else prev = prev * (z / (z * z - 1.0));                                  // infinite geometric tail
  SPL_AT(n - 1) = prev;
  for (int i = n - 2; i >= 0; --i) {
    prev = z * (prev - SPL_AT(i));
    SPL_AT(i) = prev;
  }
}

__device__ __forceinline__ long long spline_mirror_index(long long idx, long long n) {
  if (n <= 1) return 0;
  const long long s2 = 2 * n - 2;
  if (idx < 0) {
    idx = s2 * (-idx / s2) + idx;
    idx = idx <= 1 - n ? idx + s2 : -idx;
  } else if (idx >= n) {
    idx -= s2 * (idx / s2);
    if (idx >= n) idx = s2 - idx;
  }
  return idx;
}

// Interpolation weights of one output index (NI_ZoomShift + get_spline_interpolation_weights, order 3): they depend on
// (item, axis, output index) only, so a small kernel tabulates them once instead of every line recomputing them (three
// fp64 divisions per output).  Entry = {first tap (canvas index; < -2^40: output is cval), w0, w1, w2, w3}, stored
// component-major per (item, axis): component c of output a at [c * n_out + a].
constexpr int SPL_WSTRIDE = 5;
constexpr double SPL_INVALID = -4.0e15;

__global__ void __launch_bounds__(128) spline_weights_kernel(long long N, SplineGeom g, const double* __restrict__ pos_x,
                                                             const double* __restrict__ pos_y, const int32_t* __restrict__ ax,
                                                             const int32_t* __restrict__ ay, double* __restrict__ W) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * 2 * g.n_out) return;
  const long long k = t / (2 * g.n_out);
  const int rem = (int)(t - k * 2 * g.n_out);
  const int axis = rem / g.n_out, a = rem - axis * g.n_out;
  const double pos = axis == 0 ? (pos_x ? pos_x[k] : g.pos_x1) : (pos_y ? pos_y[k] : g.pos_y1);
  const long long anchor = axis == 0 ? (pos_x ? ax[k] : g.ax1) : (pos_y ? ay[k] : g.ay1);
  double* __restrict__ w = W + (k * 2 + axis) * (long long)g.n_out * SPL_WSTRIDE + a;  // component c at w[c * n_out]
  const int E = g.n_out;
  const double cc = (double)(anchor + a) + (-pos);  // source coordinate of canvas index anchor + a
  if (cc < 0.0 || cc > (double)(g.F - 1)) {         // mode='constant': outside the canvas -> cval
    w[0] = SPL_INVALID;
    w[E] = w[2 * E] = w[3 * E] = w[4 * E] = 0.0;
    return;
  }
  const double fl = floor(cc);
  const double y = cc - fl, zc = 1.0 - y;
  const double w1 = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
  const double w2 = (zc * zc * (zc - 2.0) * 3.0 + 4.0) / 6.0;
  const double w0 = zc * zc * zc / 6.0;
  w[0] = fl - 1.0;
  w[E] = w0;
  w[2 * E] = w1;
  w[3 * E] = w2;
  w[4 * E] = 1.0 - w0 - w1 - w2;
}

// value of the shifted line for one tabulated output
__device__ __forceinline__ double spline_eval(const double* line, const int LS, const SplineLine& sl, long long F,
                                              const double* __restrict__ w, const int E) {
  const double s0 = __ldg(w);
  if (s0 < -1.0e15) return 0.0;
  const long long start = (long long)s0;
  double t = 0.0;
  if (start >= sl.lo && start + 3 < sl.hi) {  // the common case: all four taps inside the segment (and the canvas)
    const int u = (int)(start - sl.lo);
#pragma unroll
    for (int l = 0; l < 4; ++l) t += SPL_AT(u + l) * __ldg(w + (1 + l) * E);
    return t;
  }
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const long long idx = spline_mirror_index(start + l, F);
    const double c = (idx >= sl.lo && idx < sl.hi) ? SPL_AT(idx - sl.lo) : 0.0;  // beyond the segment: < |z|^P
    t += c * __ldg(w + (1 + l) * E);
  }
  return t;
}

__device__ __forceinline__ SplineLine spline_segment(const SplineGeom& g, long long origin) {
  SplineLine sl;
  const long long lo = origin - g.P, hi = origin + g.S + g.P;
  sl.lo = (int)(lo < 0 ? 0 : (lo > g.F ? g.F : lo));
  sl.hi = (int)(hi > g.F ? g.F : (hi < 0 ? 0 : hi));
  if (sl.hi < sl.lo) sl.hi = sl.lo;
  return sl;
}

// pass X: one thread per (stamp, column s, band): data (N,S,S,C) -> U (N, n_out, S, C)
template <typename TS>
__global__ void __launch_bounds__(SPL_THREADS) spline_pass_x_kernel(const TS* __restrict__ data, long long N, int C, SplineGeom g,
                                                                    const int32_t* __restrict__ origin_x,
                                                                    const double* __restrict__ pos_x, const double* __restrict__ W,
                                                                    double* __restrict__ U) {
  extern __shared__ double spl_lines[];
  const int LS = SPL_THREADS;
  double* line = spl_lines + threadIdx.x;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = g.S * C;
  if (t >= N * per) return;
  const long long k = t / per;
  const int sc = (int)(t - k * per);  // s * C + ch
  const long long origin = origin_x ? origin_x[k] : (pos_x ? g.origin : g.origin_x1);
  const SplineLine sl = spline_segment(g, origin);
  const int n = sl.hi - sl.lo;
  const double gain = 6.0;  // (1 - z)(1 - 1/z)
  for (int i = 0; i < n; ++i) SPL_AT(i) = 0.0;
  const TS* __restrict__ src = data + k * (long long)g.S * per + sc;
  for (int r0 = 0; r0 < g.S; r0 += SPL_LD) {  // SPL_LD independent loads in flight per thread
    TS v[SPL_LD];
#pragma unroll
    for (int j = 0; j < SPL_LD; ++j)
      if (r0 + j < g.S) v[j] = __ldg(src + (long long)(r0 + j) * per);
#pragma unroll
    for (int j = 0; j < SPL_LD; ++j) {
      const long long u = origin + r0 + j - sl.lo;  // data outside the canvas is dropped
      if (r0 + j < g.S && u >= 0 && u < n) SPL_AT(u) = gain * (double)v[j];
    }
  }
  spline_prefilter_line(line, LS, sl, g.F);
  const double* __restrict__ w = W + (k * 2 + 0) * (long long)g.n_out * SPL_WSTRIDE;
  double* __restrict__ dst = U + k * (long long)g.n_out * per + sc;
#pragma unroll 4
  for (int a = 0; a < g.n_out; ++a) dst[(long long)a * per] = spline_eval(line, LS, sl, g.F, w + a, g.n_out);
}

// pass Y: one thread per (stamp, output row a, band): U (N, n_out, S, C) -> T (N, C, n_out, n_out)
__global__ void __launch_bounds__(SPL_THREADS) spline_pass_y_kernel(const double* __restrict__ U, long long N, int C, SplineGeom g,
                                                                    const int32_t* __restrict__ origin_y,
                                                                    const double* __restrict__ pos_y, const double* __restrict__ W,
                                                                    double* __restrict__ T) {
  extern __shared__ double spl_lines[];
  const int LS = SPL_THREADS;
  double* line = spl_lines + threadIdx.x;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = g.n_out * C;
  if (t >= N * per) return;
  const long long k = t / per;
  const int ac = (int)(t - k * per);
  const int a = ac / C, ch = ac - a * C;
  const long long origin = origin_y ? origin_y[k] : (pos_y ? g.origin : g.origin_y1);
  const SplineLine sl = spline_segment(g, origin);
  const int n = sl.hi - sl.lo;
  const double gain = 6.0;
  for (int i = 0; i < n; ++i) SPL_AT(i) = 0.0;
  const double* __restrict__ src = U + (k * g.n_out + a) * (long long)g.S * C + ch;
  for (int s0 = 0; s0 < g.S; s0 += SPL_LD) {
    double v[SPL_LD];
#pragma unroll
    for (int j = 0; j < SPL_LD; ++j)
      if (s0 + j < g.S) v[j] = __ldg(src + (long long)(s0 + j) * C);
#pragma unroll
    for (int j = 0; j < SPL_LD; ++j) {
      const long long u = origin + s0 + j - sl.lo;
      if (s0 + j < g.S && u >= 0 && u < n) SPL_AT(u) = gain * v[j];
    }
  }
  spline_prefilter_line(line, LS, sl, g.F);
  const double* __restrict__ w = W + (k * 2 + 1) * (long long)g.n_out * SPL_WSTRIDE;
  double* __restrict__ dst = T + ((k * C + ch) * g.n_out + a) * (long long)g.n_out;  // planar windows (N, C, E, E)
#pragma unroll 4
  for (int b = 0; b < g.n_out; ++b) dst[b] = spline_eval(line, LS, sl, g.F, w + b, g.n_out);
}

// ---------------------------------------------------------------------------------------------
// The common case of the placement — the data block and its +-P margin lie strictly inside the canvas,
// which is always so for the padded stamp of field_deblender.py:66-81 once F >= S + 2P + 2 — as ONE
// kernel with a warp per line.  A CTA owns one (stamp, band): pass X over the S columns leaves its
// (E x S) result in shared memory, pass Y over the E rows writes the window.  A line's recursions are
// warp scans (two samples per lane, Kogge-Stone over the lanes with multiplier z^2), head and tail
// coefficients are closed forms (c[-m] = z^m c[0]; c[S-1+m] = kappa z^m c+[S-1]), so only S coefficients
// are stored per line, and the E outputs of a line are evaluated 32 at a time.  Same truncation as the
// thread-per-line kernels (taps beyond +-P are dropped), same weight table; the association order of
// the recursions differs (scan vs. sequential): ~1e-16 relative.
// ---------------------------------------------------------------------------------------------
constexpr int SPW_WARPS = 8, SPW_ZPOW = 80;
__constant__ double c_zpow[SPW_ZPOW];  // z^m by repeated multiplication, like the sequential recursion

struct SpwCoef {
  double c0, c1;     // final coefficients of this lane's two samples (2*lane, 2*lane+1)
  double cfirst;     // c[0]
  double cplast;     // c+[S-1]
};

__device__ __forceinline__ SpwCoef spw_prefilter(double x0, double x1, int S, int lane) {
  const double z = -0.26794919243112270647, kappa = z / (z * z - 1.0);
  const double q1 = z * z, q2 = q1 * q1, q4 = q2 * q2, q8 = q4 * q4, q16 = q8 * q8;
  const unsigned full = 0xffffffffu;
  // causal: c+[i] = x[i] + z c+[i-1]
  const double t0 = x0, t1 = x1 + z * t0;
  double v = t1, u;
  u = __shfl_up_sync(full, v, 1);  if (lane >= 1) v += q1 * u;
  u = __shfl_up_sync(full, v, 2);  if (lane >= 2) v += q2 * u;
  u = __shfl_up_sync(full, v, 4);  if (lane >= 4) v += q4 * u;
  u = __shfl_up_sync(full, v, 8);  if (lane >= 8) v += q8 * u;
  u = __shfl_up_sync(full, v, 16); if (lane >= 16) v += q16 * u;
  double carry = __shfl_up_sync(full, v, 1);
  if (lane == 0) carry = 0.0;
  const double cp0 = t0 + z * carry, cp1 = t1 + q1 * carry;
  SpwCoef r;
  const double last_pair = (S - 1) & 1 ? cp1 : cp0;
  r.cplast = __shfl_sync(full, last_pair, (S - 1) >> 1);
  // anti-causal: c[k] = z (c[k+1] - c+[k]), beyond the 64 samples c[64] = kappa c+[64] = kappa z c+[63]
  double u0 = -z * cp0, u1 = -z * cp1;
  if (lane == 31) u1 += z * (kappa * z * cp1);
  const double r1 = u1, r0 = u0 + z * r1;
  v = r0;
  u = __shfl_down_sync(full, v, 1);  if (lane + 1 < 32) v += q1 * u;
  u = __shfl_down_sync(full, v, 2);  if (lane + 2 < 32) v += q2 * u;
  u = __shfl_down_sync(full, v, 4);  if (lane + 4 < 32) v += q4 * u;
  u = __shfl_down_sync(full, v, 8);  if (lane + 8 < 32) v += q8 * u;
  u = __shfl_down_sync(full, v, 16); if (lane + 16 < 32) v += q16 * u;
  carry = __shfl_down_sync(full, v, 1);
  if (lane == 31) carry = 0.0;
  r.c1 = r1 + z * carry;
  r.c0 = r0 + q1 * carry;
  r.cfirst = __shfl_sync(full, r.c0, 0);
  return r;
}

// the line of one warp, extended: ext[3 + P + i] = c[i] for i in [-P, S+P), three zeros on either side, so that
// every output reads its four taps without a branch (taps beyond +-P are the zeros)
__device__ __forceinline__ void spw_extend(double* ext, const double* zp, const SpwCoef& cf, int S, int P, int lane) {
  const double kappa = -0.26794919243112270647 / (0.26794919243112270647 * 0.26794919243112270647 - 1.0);
  const int i0 = 2 * lane, i1 = 2 * lane + 1;
  if (i0 < S) ext[3 + P + i0] = cf.c0;
  if (i1 < S) ext[3 + P + i1] = cf.c1;
  for (int m = 1 + lane; m <= P; m += 32) {
    ext[3 + P - m] = zp[m] * cf.cfirst;               // head: c[-m] = z^m c[0]
    ext[3 + P + S - 1 + m] = kappa * zp[m] * cf.cplast;  // tail: c[S-1+m] = kappa z^m c+[S-1]
  }
}

// one output of a line: four taps of the extended line, tabulated weights (component-major, E apart)
__device__ __forceinline__ double spw_eval(const double* ext, int S, int P, long long origin, const double* __restrict__ w, int E) {
  const double s0 = __ldg(w);
  if (s0 < -1.0e15) return 0.0;
  const long long rr = (long long)s0 - origin;
  if (rr < -(long long)P - 3 || rr >= (long long)S + P) return 0.0;
  const double* e = ext + (int)rr + P + 3;
  return e[0] * __ldg(w + E) + e[1] * __ldg(w + 2 * E) + e[2] * __ldg(w + 3 * E) + e[3] * __ldg(w + 4 * E);
}

template <typename TS>
__global__ void __launch_bounds__(256) spline_place_warp_kernel(const TS* __restrict__ data, int C, SplineGeom g,
                                                                const double* __restrict__ W, double* __restrict__ T) {
  extern __shared__ double spw_smem[];
  const int S = g.S, E = g.n_out, P = g.P;
  const int pitch = S | 1, next = S + 2 * P + 6;
  double* U = spw_smem;                                   // [E][pitch]: pass-X result of this (stamp, band)
  double* zp = U + (size_t)E * pitch;                     // z^m
  double* extall = zp + SPW_ZPOW;                         // [warps][next] extended line in flight
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* ext = extall + warp * next;
  const long long k = blockIdx.x / C;
  const int ch = blockIdx.x - (int)k * C;
  for (int i = threadIdx.x; i < SPW_ZPOW; i += blockDim.x) zp[i] = c_zpow[i];
  for (int i = threadIdx.x; i < SPW_WARPS * next; i += blockDim.x) extall[i] = 0.0;  // the zero borders stay zero
  __syncthreads();
  const long long origin = g.origin;
  const double* __restrict__ wx = W + (k * 2 + 0) * (long long)E * SPL_WSTRIDE;
  const double* __restrict__ wy = W + (k * 2 + 1) * (long long)E * SPL_WSTRIDE;
  const int i0 = 2 * lane, i1 = 2 * lane + 1;
  const TS* __restrict__ src = data + k * (long long)S * S * C + ch;
  for (int s = warp; s < S; s += SPW_WARPS) {
    const double x0 = i0 < S ? 6.0 * (double)__ldg(src + ((long long)i0 * S + s) * C) : 0.0;
    const double x1 = i1 < S ? 6.0 * (double)__ldg(src + ((long long)i1 * S + s) * C) : 0.0;
    const SpwCoef cf = spw_prefilter(x0, x1, S, lane);
    spw_extend(ext, zp, cf, S, P, lane);
    __syncwarp();
    for (int a = lane; a < E; a += 32) U[(size_t)a * pitch + s] = spw_eval(ext, S, P, origin, wx + a, E);
    __syncwarp();
  }
  __syncthreads();
  double* __restrict__ dst = T + (k * C + ch) * (long long)E * E;
  for (int a = warp; a < E; a += SPW_WARPS) {
    const double x0 = i0 < S ? 6.0 * U[(size_t)a * pitch + i0] : 0.0;
    const double x1 = i1 < S ? 6.0 * U[(size_t)a * pitch + i1] : 0.0;
    const SpwCoef cf = spw_prefilter(x0, x1, S, lane);
    spw_extend(ext, zp, cf, S, P, lane);
    __syncwarp();
    for (int b = lane; b < E; b += 32) dst[(long long)a * E + b] = spw_eval(ext, S, P, origin, wy + b, E);
    __syncwarp();
  }
}

// position fit objective (deblend_cutout/optimization.py:21-33): sum over a placed window T (E,E) of
// T^2 - 2*img*T against one band of the field; with the field's own sum of squares this gives
// mean((img - shifted)^2) over the whole canvas without touching the rest of it.  One CTA, fixed order.
__global__ void __launch_bounds__(256) shift_objective_kernel(const double* __restrict__ field, long long F, int C, int band,
                                                              const double* __restrict__ T, int E, int ax, int ay,
                                                              double sumsq_field, double* __restrict__ out) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int e = threadIdx.x; e < E * E; e += 256) {
    const int a = e / E, b = e - a * E;
    const long long X = (long long)ax + a, Y = (long long)ay + b;
    if (X < 0 || X >= F || Y < 0 || Y >= F) continue;
    const double t = T[e], v = field[(X * F + Y) * C + band];
    acc += t * t - 2.0 * v * t;
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (sumsq_field + s[0]) / ((double)F * (double)F);
}

// The same for a batch: item i = placed window T[i] (E x E) at (ax[i], ay[i]); one CTA per item, fixed order.
__global__ void __launch_bounds__(256) shift_objective_batch_kernel(const double* __restrict__ field, long long F, int C, int band,
                                                                    const double* __restrict__ T, int E, const int32_t* __restrict__ ax,
                                                                    const int32_t* __restrict__ ay, double sumsq_field,
                                                                    double* __restrict__ out) {
  __shared__ double s[256];
  const long long i = blockIdx.x;
  const double* Ti = T + i * (long long)E * E;
  const long long x0 = ax[i], y0 = ay[i];
  double acc = 0.0;
  for (int e = threadIdx.x; e < E * E; e += 256) {
    const int a = e / E, b = e - a * E;
    const long long X = x0 + a, Y = y0 + b;
    if (X < 0 || X >= F || Y < 0 || Y >= F) continue;
    const double t = Ti[e], v = field[(X * F + Y) * C + band];
    acc += t * t - 2.0 * v * t;
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[i] = (sumsq_field + s[0]) / ((double)F * (double)F);
}

// sum of squares of one band of the field (two deterministic passes through `partial`)
__global__ void __launch_bounds__(256) band_sumsq_kernel(const double* __restrict__ field, long long npix, int C, int band,
                                                         double* __restrict__ partial) {
  __shared__ double s[8];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const double v = field[i * C + band];
    acc += v * v;
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
__global__ void __launch_bounds__(256) sum_final_kernel(const double* __restrict__ partial, int nb, double* __restrict__ out) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = s[0];
}

constexpr int SPL_LMAX = 192;  // samples per line: SPL_LMAX * SPL_THREADS doubles of shared memory at most (96 KB)

// per-device one-time set-up of the spline kernels (function attributes and the z^m table are per device)
static int spline_device_setup() {
  static bool done[64] = {};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  std::lock_guard<std::mutex> lk(mu);
  if (done[dev]) return 0;
  const int bytes = SPL_LMAX * SPL_THREADS * (int)sizeof(double);
  double zp[SPW_ZPOW];
  zp[0] = 1.0;
  for (int i = 1; i < SPW_ZPOW; ++i) zp[i] = zp[i - 1] * -0.26794919243112270647;
  if (cudaFuncSetAttribute(spline_pass_x_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess ||
      cudaFuncSetAttribute(spline_pass_x_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess ||
      cudaFuncSetAttribute(spline_pass_y_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess ||
      cudaFuncSetAttribute(spline_place_warp_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(spline_place_warp_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
      cudaMemcpyToSymbol(c_zpow, zp, sizeof zp) != cudaSuccess)
    return -1;
  done[dev] = true;
  return 0;
}

}  // namespace dbv
using namespace dbv;

extern "C" int64_t dbv_spline_scratch_doubles(int64_t N, int S, int C, int P) {
  return N * (int64_t)(S + 2 * P + 2) * ((int64_t)S * C + 2 * SPL_WSTRIDE);  // pass-X output + weight table
}

extern "C" int dbv_spline_extent(int S, int P) {
  if (S < 1 || P < 0 || S + 2 * P > SPL_LMAX) return fail(DBV_ERR_UNSUPPORTED, "dbv_spline_extent: S + 2P = %d exceeds the %d-sample line buffer", S + 2 * P, SPL_LMAX);
  return S + 2 * P + 2;
}

extern "C" int dbv_spline_place(const void* data, int data_dtype, int64_t N, int S, int C, int64_t F, int origin,
                                const int32_t* origin_x, const int32_t* origin_y, const double* pos_x, const double* pos_y,
                                const int32_t* ax, const int32_t* ay, int P, double* scratch, double* placed, void* stream) {
  DBV_REQUIRE(N >= 0 && S > 0 && C > 0 && F > 0 && P >= 0, "dbv_spline_place: bad sizes");
  if (N == 0) return DBV_OK;
  DBV_REQUIRE(data && pos_x && pos_y && ax && ay && scratch && placed, "dbv_spline_place: null pointer");
  DBV_REQUIRE((origin_x == nullptr) == (origin_y == nullptr), "dbv_spline_place: give both origin arrays or neither");
  DBV_REQUIRE(data_dtype == DBV_F32 || data_dtype == DBV_F64, "dbv_spline_place: bad data dtype %d", data_dtype);
  const int n_out = dbv_spline_extent(S, P);
  if (n_out < 0) return n_out;
  SplineGeom g = {};
  g.F = F; g.S = S; g.P = P; g.origin = origin; g.n_out = n_out;
  cudaStream_t st = (cudaStream_t)stream;
  if (spline_device_setup()) return fail(DBV_ERR_CUDA, "dbv_spline_place: per-device set-up of the spline kernels failed");
  const long long tx = N * (long long)S * C, ty = N * (long long)n_out * C;
  const unsigned gx = (unsigned)((tx + SPL_THREADS - 1) / SPL_THREADS), gy = (unsigned)((ty + SPL_THREADS - 1) / SPL_THREADS);
  const size_t smem = (size_t)(S + 2 * P) * SPL_THREADS * sizeof(double);
  double* W = scratch + N * (long long)n_out * S * C;  // weight table behind the pass-X output
  spline_weights_kernel<<<(unsigned)((N * 2 * n_out + 127) / 128), 128, 0, st>>>(N, g, pos_x, pos_y, ax, ay, W);
  DBV_LAUNCH_CHECK();
  // fast path: every item's data sits at `origin` with its +-P margin strictly inside the canvas
  bool fast = !origin_x && S <= 64 && P + 1 < SPW_ZPOW && (long long)origin - P > 0 && (long long)origin + S + P < F && N * (long long)C < (1ll << 31);
  if (const char* e = dbv_env("DBV_SPLINE_WARP")) fast = fast && atoi(e) != 0;  // 0: thread-per-line kernels (cross-check)
  if (fast) {
    const size_t smem_w = ((size_t)n_out * (S | 1) + SPW_ZPOW + (size_t)SPW_WARPS * (S + 2 * P + 6)) * sizeof(double);
    if (smem_w <= 200 * 1024) {
      if (data_dtype == DBV_F32)
        spline_place_warp_kernel<float><<<(unsigned)(N * C), 32 * SPW_WARPS, smem_w, st>>>((const float*)data, C, g, W, placed);
      else
        spline_place_warp_kernel<double><<<(unsigned)(N * C), 32 * SPW_WARPS, smem_w, st>>>((const double*)data, C, g, W, placed);
      DBV_LAUNCH_CHECK();
      return DBV_OK;
    }
  }
  if (data_dtype == DBV_F32)
    spline_pass_x_kernel<float><<<gx, SPL_THREADS, smem, st>>>((const float*)data, N, C, g, origin_x, pos_x, W, scratch);
  else
    spline_pass_x_kernel<double><<<gx, SPL_THREADS, smem, st>>>((const double*)data, N, C, g, origin_x, pos_x, W, scratch);
  DBV_LAUNCH_CHECK();
  spline_pass_y_kernel<<<gy, SPL_THREADS, smem, st>>>(scratch, N, C, g, origin_y, pos_y, W, placed);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

extern "C" int dbv_band_sumsq(const double* field, int64_t F, int C, int band, double* out, void* scratch, int64_t scratch_bytes,
                              void* stream) {
  DBV_REQUIRE(field && out && scratch, "dbv_band_sumsq: null pointer");
  DBV_REQUIRE(F > 0 && C > 0 && band >= 0 && band < C, "dbv_band_sumsq: bad sizes");
  DBV_REQUIRE(scratch_bytes >= dbv_mse_scratch_bytes(), "dbv_band_sumsq: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  const long long npix = (long long)F * F;
  long long want = (npix + 1023) / 1024;
  const int nb = (int)(want < MSE_BLOCKS ? (want < 1 ? 1 : want) : MSE_BLOCKS);
  band_sumsq_kernel<<<nb, 256, 0, st>>>(field, npix, C, band, (double*)scratch);
  DBV_LAUNCH_CHECK();
  sum_final_kernel<<<1, 256, 0, st>>>((const double*)scratch, nb, out);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

extern "C" int dbv_shift_objective(const double* field, int64_t F, int C, int band, const double* placed, int E, int ax, int ay,
                                   double sumsq_field, double* out, void* stream) {
  DBV_REQUIRE(field && placed && out, "dbv_shift_objective: null pointer");
  DBV_REQUIRE(F > 0 && C > 0 && band >= 0 && band < C && E > 0, "dbv_shift_objective: bad sizes");
  shift_objective_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(field, F, C, band, placed, E, ax, ay, sumsq_field, out);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

extern "C" int dbv_shift_objective_batch(const double* field, int64_t F, int C, int band, const double* placed, int E,
                                         const int32_t* ax, const int32_t* ay, int64_t M, double sumsq_field, double* out, void* stream) {
  DBV_REQUIRE(M >= 0 && M < (1ll << 31), "dbv_shift_objective_batch: bad batch size");
  if (M == 0) return DBV_OK;
  DBV_REQUIRE(field && placed && ax && ay && out, "dbv_shift_objective_batch: null pointer");
  DBV_REQUIRE(F > 0 && C > 0 && band >= 0 && band < C && E > 0, "dbv_shift_objective_batch: bad sizes");
  shift_objective_batch_kernel<<<(unsigned)M, 256, 0, (cudaStream_t)stream>>>(field, F, C, band, placed, E, ax, ay, sumsq_field, out);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

extern "C" int dbv_position_objective(const double* field, int64_t F, int C, int band, const double* placed1, int E1, int a1x, int a1y,
                                      double x0, double x1, int P, double sumsq_field, double* scratch, double* placed2,
                                      double* out_dev, double* out_host, void* stream) {
  DBV_REQUIRE(field && placed1 && scratch && placed2 && out_dev && out_host, "dbv_position_objective: null pointer");
  DBV_REQUIRE(F > 0 && C > 0 && band >= 0 && band < C && E1 > 0 && P >= 0, "dbv_position_objective: bad sizes");
  DBV_REQUIRE(x0 == x0 && x1 == x1 && fabs(x0) < 1e9 && fabs(x1) < 1e9, "dbv_position_objective: bad shift");
  const int E2 = dbv_spline_extent(E1, P);
  if (E2 < 0) return E2;
  SplineGeom g = {};
  g.F = F; g.S = E1; g.P = P; g.n_out = E2;
  g.pos_x1 = x0; g.pos_y1 = x1;
  g.origin_x1 = a1x; g.origin_y1 = a1y;
  g.ax1 = a1x - P - 1 + (int)floor(x0);
  g.ay1 = a1y - P - 1 + (int)floor(x1);
  cudaStream_t st = (cudaStream_t)stream;
  if (spline_device_setup()) return fail(DBV_ERR_CUDA, "dbv_position_objective: per-device set-up of the spline kernels failed");
  const unsigned gx = (unsigned)((E1 + SPL_THREADS - 1) / SPL_THREADS), gy = (unsigned)((E2 + SPL_THREADS - 1) / SPL_THREADS);
  const size_t smem = (size_t)(E1 + 2 * P) * SPL_THREADS * sizeof(double);
  double* W = scratch + (long long)E2 * E1;
  spline_weights_kernel<<<(unsigned)((2 * E2 + 127) / 128), 128, 0, st>>>(1, g, nullptr, nullptr, nullptr, nullptr, W);
  DBV_LAUNCH_CHECK();
  spline_pass_x_kernel<double><<<gx, SPL_THREADS, smem, st>>>(placed1, 1, 1, g, nullptr, nullptr, W, scratch);
  DBV_LAUNCH_CHECK();
  spline_pass_y_kernel<<<gy, SPL_THREADS, smem, st>>>(scratch, 1, 1, g, nullptr, nullptr, W, placed2);
  DBV_LAUNCH_CHECK();
  shift_objective_kernel<<<1, 256, 0, st>>>(field, F, C, band, placed2, E2, g.ax1, g.ay1, sumsq_field, out_dev);
  DBV_LAUNCH_CHECK();
  DBV_CUDA(cudaMemcpyAsync(out_host, out_dev, sizeof(double), cudaMemcpyDeviceToHost, st));
  DBV_CUDA(cudaStreamSynchronize(st));
  return DBV_OK;
}
