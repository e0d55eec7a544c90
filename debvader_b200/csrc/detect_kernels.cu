// Object detection on the device: the step reference detect/detection.py:5-56 delegates to the CPU library
// `sep` (sep.Background + sep.extract on the r band), restated stage by stage from the published algorithm
// (Bertin & Arnouts 1996; Barbary 2016) — see oracle/detect_numpy.py, whose arithmetic this file reproduces BIT FOR BIT:
//
//   B1  det_mesh_kernel       per 64x64 mesh: mean / sigma, one +-2 sigma clip, level histogram, iterated 3-sigma clipping
//   B2  det_mesh_{fill,median,rank,final}_kernel  bad-mesh fill, 3x3 median filter, global background / rms (medians by rank
//                             selection), threshold, y-spline of the mesh map
//   B3  det_nodes_kernel + det_foreground_kernel   bicubic-spline background map, foreground = band - background
//   E1  det_filter_tiled_kernel<7> (det_filter_kernel for other mask sizes)   normalised 7x7 matched filter (zero outside the image), threshold test -> initial labels
//   E2  det_ccl_merge_kernel  8-connected components by union-find on the label image (root = smallest raster index)
//   E3  det_ccl_flatten_stats / mark / count / scan / scatter   pixel count, bounding columns and LARGEST raster index per object; the
//                             objects with >= minarea pixels listed by ascending largest raster index = the order in which
//                             Lutz's one-pass scan completes them = sep's output order
//   E4  det_moments_kernel    barycentre of the unfiltered foreground (double, raster order), centres rounded half-to-even
//
// Not restated: multi-threshold deblending and the `clean` pass (a connected footprint is ONE detection).
// Parity with `sep` itself is UNPINNED (sep is not installable here, the reference holds no golden detections); parity with
// the oracle is exact.  The file is compiled with -fmad=false: every product and sum is rounded on its own, as numpy does.
// All HBM-bound integer / float work: coalesced loads, shared-memory tiles, integer atomics only (deterministic).
#ifdef DBV_EMULATE  // host build of the SAME kernels for the container's logic check (tools/detect_emul, no GPU there); never in the library
#include "detect_emul.h"
#else
#include "common.cuh"
#include <limits.h>
#define DET_LAUNCH(kernel, grid, block, smem, ...)     \
  do {                                                 \
    kernel<<<grid, block, smem, st>>>(__VA_ARGS__);    \
    DBV_LAUNCH_CHECK();                                \
  } while (0)
#define DET_LAUNCH_NOSYNC DET_LAUNCH  // (the emulation runs barrier-free kernels without OS threads)
#define DET_MEMSET(p, v, n) DBV_CUDA(cudaMemsetAsync(p, v, n, st))
#define DET_DYN_SMEM(T, name) extern __shared__ T name[]
#endif

namespace dbv {

constexpr int DET_BW = 64;            // mesh size (sep.Background defaults)
constexpr int DET_MAXLEVELS = 4096;   // histogram levels
constexpr float DET_BIG = 1e30f;
constexpr int DET_CHUNK = 4096;       // pixels per CTA of the compaction kernels (256 threads x 16)
constexpr int DET_MAXK = 15;          // largest filter mask side

struct DetTaps { float v[DET_MAXK * DET_MAXK]; };
// A call works on a REGION of the field (the whole field on one GPU; a rank's owner tile + halo when the field is tiled over GPUs):
// the mesh grid, the spline positions and the output coordinates are those of the WHOLE field, indices inside kernels are region-local.
struct DetGeom {
  int H, W;                // the whole field
  int RH, RW;              // the region
  int gy0, gx0;            // origin of the region in the field
  int ny, nx;              // mesh grid of the whole field
  int vy0, vy1, vx0, vx1;  // region-local area where the filtered image is exact (region minus kh/2, kw/2 on edges inside the field)
  int oy0, oy1, ox0, ox1;  // owner tile in field coordinates: an object belongs to the call whose tile holds its last pixel
};  // passed by value: a launch parameter (constant bank), no shared state between calls

// ---- band -> compact f64 plane ---------------------------------------------------------------------------------------
template <typename T>
__global__ void det_band_kernel(const T* __restrict__ field, long long H, long long W, long long pitch, int C, int band, double* __restrict__ out) {
  const long long n = H * W;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    const long long y = p / W, x = p - y * W;
    out[p] = (double)field[(y * pitch + x) * C + band];
  }
}

// ---- B1 --------------------------------------------------------------------------------------------------------------
// One CTA of 64 threads per mesh: thread c owns mesh column c (sequential double sums down the column, the columns then combined in
// order by thread 0: the oracle's summation order).  The level histogram is counted with shared-memory integer atomics and turned into an exclusive
// prefix sum in place, so that the iterated clipping (oracle: histogram_guess) costs O(range / 64 + log^2) per iteration
// instead of a serial pass over <= 4096 levels:
//   * the three sums over [lcut, hcut] are sums of integers (< 2^53): any order gives the oracle's doubles exactly — strided over the CTA;
//   * the two-pointer walk "advance the side with the smaller running count" is the merge of the two prefix-sum sequences
//     P(j) = count of the first j levels, Q(i) = count of the last i levels, ties to the high side: after n = hcut - lcut + 1 steps
//     it has taken k = #{ j < n : j + #{ i < n : Q(i) <= P(j) } < n } = #{ j < n : Q(n - 1 - j) > P(j) } low steps, and the predicate is
//     monotone in j — one binary search on the prefix sums.
__device__ __forceinline__ int det_cnt(const int* pre, int i) { return pre[i + 1] - pre[i]; }

__global__ void __launch_bounds__(64) det_mesh_kernel(const double* __restrict__ band, const DetGeom G, float* __restrict__ back0,
                                                      float* __restrict__ sig0) {
  __shared__ int pre[DET_MAXLEVELS + 1];  // level counts, then their exclusive prefix sums (pre[nlevels .. 4096] = total)
  __shared__ double rs[DET_BW], rq[DET_BW], rn[DET_BW];
  __shared__ long long p0[DET_BW], p1[DET_BW], p2[DET_BW];
  __shared__ int part[DET_BW];
  __shared__ float s_lcut, s_hcut, s_qscale, s_cste, s_qzero;
  __shared__ int s_nlevels, s_bad, s_lo, s_hi, s_go;
  __shared__ double s_mean2;
  const int mx = blockIdx.x, my = blockIdx.y, t = threadIdx.x;
  const int nx = G.nx;
  const int x0 = mx * DET_BW, y0 = my * DET_BW;  // field coordinates
  const int mw = min(DET_BW, G.W - x0), mh = min(DET_BW, G.H - y0);
  // a mesh is computed by the call whose region holds all of it (tiled fields: the other meshes come from the other ranks)
  if (y0 < G.gy0 || y0 + mh > G.gy0 + G.RH || x0 < G.gx0 || x0 + mw > G.gx0 + G.RW) return;
  // thread t owns mesh COLUMN t: every pass reads the mesh row by row, the 64 threads one 512-byte segment at a time (coalesced; the
  // 32 KB of the mesh stay in L1 / L2 between the passes), so the CTA needs no shared-memory copy of the mesh and twice as many CTAs fit
  const double* col = band + (long long)(y0 - G.gy0) * G.RW + (x0 - G.gx0) + t;
  const long long pitch = G.RW;
  // pass 1: all pixels (sequential double sums down the column, the columns then combined in order by thread 0: the oracle's order)
  if (t < mw) {
    double s = 0.0, q = 0.0;
    for (int r0 = 0; r0 < mh; r0 += 8) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = r0 + i < mh ? (float)col[(r0 + i) * pitch] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (r0 + i < mh) {
          const double v = (double)f[i];
          s += v;
          q += v * v;
        }
    }
    rs[t] = s;
    rq[t] = q;
  }
  __syncthreads();
  if (t == 0) {
    double s = 0.0, q = 0.0;
    for (int c = 0; c < mw; ++c) { s += rs[c]; q += rq[c]; }
    const double n = (double)(mh * mw);
    const double mean = s / n;
    const double var = q / n - mean * mean;
    const double sigma = var > 0.0 ? sqrt(var) : 0.0;
    s_lcut = (float)(mean - 2.0 * sigma);
    s_hcut = (float)(mean + 2.0 * sigma);
  }
  __syncthreads();
  // pass 2: pixels inside the cuts
  if (t < mw) {
    const float lc = s_lcut, hc = s_hcut;
    double s = 0.0, q = 0.0, n = 0.0;
    for (int r0 = 0; r0 < mh; r0 += 8) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = r0 + i < mh ? (float)col[(r0 + i) * pitch] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (r0 + i < mh && f[i] >= lc && f[i] <= hc) {
          const double v = (double)f[i];
          n += 1.0;
          s += v;
          q += v * v;
        }
    }
    rs[t] = s;
    rq[t] = q;
    rn[t] = n;
  }
  __syncthreads();
  if (t == 0) {
    double s = 0.0, q = 0.0, n = 0.0;
    for (int c = 0; c < mw; ++c) { n += rn[c]; s += rs[c]; q += rq[c]; }
    const double nall = (double)(mh * mw);
    if (n < nall * 0.5 || n < 1.0) {
      s_bad = 1;
    } else {
      s_bad = 0;
      const double mean2 = s / n;
      const double var2 = q / n - mean2 * mean2;
      const double sigma2 = var2 > 0.0 ? sqrt(var2) : 0.0;
      int nl = (int)(0.9973557010035817 /* sqrt(2/pi) * 5 / 4 */ * n + 1.0);
      if (nl > DET_MAXLEVELS) nl = DET_MAXLEVELS;
      const float qscale = sigma2 > 0.0 ? (float)(10.0 * sigma2 / (double)nl) : 1.0f;
      const float qzero = (float)(mean2 - 5.0 * sigma2);
      s_nlevels = nl;
      s_qscale = qscale;
      s_qzero = qzero;
      s_cste = (float)(0.499999 - (double)qzero / (double)qscale);
      s_mean2 = mean2;
    }
  }
  for (int i = t; i <= DET_MAXLEVELS; i += DET_BW) pre[i] = 0;
  __syncthreads();
  if (s_bad) {
    if (t == 0) {
      back0[my * nx + mx] = -DET_BIG;
      sig0[my * nx + mx] = -DET_BIG;
    }
    return;
  }
  const int nl = s_nlevels, nm1 = nl - 1;
  if (t < mw) {
    const float qs = s_qscale, cs = s_cste;
    for (int r0 = 0; r0 < mh; r0 += 8) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = r0 + i < mh ? (float)col[(r0 + i) * pitch] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (r0 + i >= mh) continue;
        const float lev = f[i] / qs + cs;
        if (lev > -1.0f && lev < (float)nl) {
          const int b = (int)lev;
          if (b >= 0 && b < nl) atomicAdd(&pre[b], 1);
        }
      }
    }
  }
  __syncthreads();
  // counts -> exclusive prefix sums, in place: thread t scans levels [64 t, 64 t + 64)
  {
    int s = 0;
    for (int i = 0; i < DET_BW; ++i) s += pre[t * DET_BW + ((i + t) & (DET_BW - 1))];  // skewed: the chunks are 64 words apart (one bank)
    part[t] = s;
    __syncthreads();
    if (t == 0) {
      int run = 0;
      for (int i = 0; i < DET_BW; ++i) {
        const int v = part[i];
        part[i] = run;
        run += v;
      }
      pre[DET_MAXLEVELS] = run;
    }
    __syncthreads();
    int run = part[t];
    for (int i = 0; i < DET_BW; ++i) {
      const int v = pre[t * DET_BW + i];
      pre[t * DET_BW + i] = run;
      run += v;
    }
  }
  // iterated clipping around the histogram median (doubles; thread 0 keeps the state, the CTA sums the levels in range)
  double sig = 10.0 * nm1, sig1 = 1.0, mea = s_mean2, med = s_mean2;
  int lcut = 0, hcut = nm1, iters = 100;
  if (t == 0) {
    s_lo = lcut;
    s_hi = hcut;
    s_go = (iters > 0 && sig >= 0.1 && fabs(sig / sig1 - 1.0) > 1e-4);
  }
  __syncthreads();
  while (s_go) {
    const int lc = s_lo, hc = s_hi;
    long long a0 = 0, a1 = 0, a2 = 0;
    for (int i = lc + t; i <= hc; i += DET_BW) {
      const long long c = det_cnt(pre, i);
      a0 += c;
      a1 += c * i;
      a2 += c * i * (long long)i;
    }
    p0[t] = a0;
    p1[t] = a1;
    p2[t] = a2;
    __syncthreads();
    if (t == 0) {
      long long tot = 0, m1 = 0, m2 = 0;
      for (int i = 0; i < DET_BW; ++i) { tot += p0[i]; m1 += p1[i]; m2 += p2[i]; }
      --iters;
      sig1 = sig;
      mea = (double)m1;
      sig = (double)m2;
      // the two-pointer walk over [lcut, hcut] by merge path (see the header of this section)
      const int n = hcut - lcut + 1;
      const int base_lo = pre[lcut], base_hi = pre[hcut + 1];
      // low step j is taken within the first n steps  <=>  fewer than n - j high counts are <= P(j)  <=>  Q(n - 1 - j) > P(j)  (Q monotone)
      int ka = 0, kb = n;
      while (ka < kb) {
        const int j = (ka + kb) >> 1;
        if (base_hi - pre[hcut + 2 - n + j] > pre[lcut + j] - base_lo) ka = j + 1;
        else kb = j;
      }
      const int lo = lcut + ka, hi = hcut - (n - ka);
      const long long lowsum = pre[lo] - base_lo, highsum = base_hi - pre[hi + 1];
      if (hi >= 0) {
        const int ca = det_cnt(pre, lo < nm1 ? lo : nm1), cb = det_cnt(pre, hi);
        const int big = ca > cb ? ca : cb;
        med = (double)hi + 0.5 + (big > 0 ? (double)(highsum - lowsum) / (2.0 * (double)big) : 0.0);
      } else {
        med = 0.0;
      }
      if (tot) {
        mea = mea / (double)tot;
        sig = sig / (double)tot - mea * mea;
      }
      sig = sig > 0.0 ? sqrt(sig) : 0.0;
      double v = med - 3.0 * sig;
      lcut = v > 0.0 ? (int)(v + 0.5) : 0;
      v = med + 3.0 * sig;
      hcut = v < (double)nm1 ? (v > 0.0 ? (int)(v + 0.5) : (int)(v - 0.5)) : nm1;
      if (hcut < lcut) hcut = lcut;
      s_lo = lcut;
      s_hi = hcut;
      s_go = (iters > 0 && sig >= 0.1 && fabs(sig / sig1 - 1.0) > 1e-4);
    }
    __syncthreads();
  }
  if (t == 0) {
    const double qzero = (double)s_qzero, qscale = (double)s_qscale;
    double bk;
    if (sig > 0.0) {
      if (fabs((mea - med) / sig) < 0.3) bk = qzero + (2.5 * med - 1.5 * mea) * qscale;
      else bk = qzero + med * qscale;
    } else {
      bk = qzero + mea * qscale;
    }
    back0[my * nx + mx] = (float)bk;
    sig0[my * nx + mx] = (float)(sig * qscale);
  }
}

// ---- B2 --------------------------------------------------------------------------------------------------------------
__device__ float det_median_small(float* a, int n) {  // insertion sort of <= 9 values
  for (int i = 1; i < n; ++i) {
    const float v = a[i];
    int j = i - 1;
    for (; j >= 0 && a[j] > v; --j) a[j + 1] = a[j];
    a[j + 1] = v;
  }
  return (n & 1) ? a[n / 2] : (a[n / 2 - 1] + a[n / 2]) * 0.5f;
}

// second derivatives / 6 of the natural cubic spline through n samples `v` (stride sv), Thomas algorithm in float32
// (oracle: _spline_d2); cp / u are n-element scratch rows of stride ss (shared memory, [k][thread]: the recurrence is a chain of
// dependent divisions, so the scratch must not add a global-memory round trip to every link)
__device__ void det_spline_d2(const float* v, int sv, int n, float* d, int sd, float* cp, float* u, int ss) {
  for (int k = 0; k < n; ++k) d[k * sd] = 0.f;
  if (n < 3) return;
  float cpk = 0.f, uk = 0.f, vm = v[0], vc = v[sv];
  for (int k = 1; k < n - 1; ++k) {
    const float vp = v[(k + 1) * sv];
    const float rhs = 6.0f * ((vp + vm) - 2.0f * vc);
    const float den = 4.0f - cpk;
    cpk = 1.0f / den;
    uk = (rhs - uk) / den;
    cp[k * ss] = cpk;
    u[k * ss] = uk;
    vm = vc;
    vc = vp;
  }
  float m = 0.f;  // M[n-1]
  for (int k = n - 2; k >= 1; --k) {
    m = u[k * ss] - cp[k * ss] * m;
    d[k * sd] = m / 6.0f;
  }
}

// bad meshes: float32 mean of the nearest good ones (raster order); thread per mesh
__global__ void det_mesh_fill_kernel(const float* __restrict__ back0, const float* __restrict__ sig0, int ny, int nx, float* __restrict__ back1,
                                     float* __restrict__ sig1) {
  const int n = ny * nx, m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  float b = back0[m], s = sig0[m];
  if (!(b > -DET_BIG)) {
    const int y = m / nx, x = m - y * nx;
    long long best = LLONG_MAX;
    for (int g = 0; g < n; ++g)
      if (back0[g] > -DET_BIG) {
        const long long dy = g / nx - y, dx = g % nx - x, d2 = dy * dy + dx * dx;
        if (d2 < best) best = d2;
      }
    float sb = 0.f, ss = 0.f;
    int k = 0;
    for (int g = 0; g < n; ++g)
      if (back0[g] > -DET_BIG) {
        const long long dy = g / nx - y, dx = g % nx - x;
        if (dy * dy + dx * dx == best) { sb += back0[g]; ss += sig0[g]; ++k; }
      }
    if (k) {
      b = sb / (float)k;
      s = ss / (float)k;
    }
  }
  back1[m] = b;
  sig1[m] = s;
}

// 3x3 median filter of both mesh maps (clipped at the rims); thread per mesh
__global__ void det_mesh_median_kernel(const float* __restrict__ back1, const float* __restrict__ sig1, int ny, int nx, float* __restrict__ back,
                                       float* __restrict__ sig) {
  const int n = ny * nx, m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const int y = m / nx, x = m - y * nx;
  float a[9], c[9];
  int k = 0;
  for (int yy = max(y - 1, 0); yy < min(y + 2, ny); ++yy)
    for (int xx = max(x - 1, 0); xx < min(x + 2, nx); ++xx) {
      a[k] = back1[yy * nx + xx];
      c[k] = sig1[yy * nx + xx];
      ++k;
    }
  back[m] = det_median_small(a, k);
  sig[m] = det_median_small(c, k);
}

// global medians by rank selection: thread i ranks element i among all n (ties broken by index) and, if it is one of the two
// middle ranks, writes itself into pick[2 * map + {0, 1}]; blockIdx.y = 0 background map, 1 sigma map
__global__ void det_mesh_rank_kernel(const float* __restrict__ back, const float* __restrict__ sig, int n, int staged, float* __restrict__ pick) {
  DET_DYN_SMEM(float, sa);
  const float* a = blockIdx.y ? sig : back;
  if (staged) {  // the whole map fits in shared memory (a 4096^2 field: 16 KB): every thread then scans it from there
    for (int j = threadIdx.x; j < n; j += blockDim.x) sa[j] = a[j];
    __syncthreads();
    a = sa;
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k1 = n / 2, k0 = (n & 1) ? k1 : k1 - 1;
  const float v = a[i];
  int rank = 0;
  for (int j = 0; j < n; ++j) {
    const float w = a[j];
    rank += (w < v) || (w == v && j < i);
  }
  if (rank == k0) pick[2 * blockIdx.y] = v;
  if (rank == k1) pick[2 * blockIdx.y + 1] = v;
}

// global background / rms / threshold, and the y-spline of the background map (thread per mesh column)
__global__ void det_mesh_final_kernel(const float* __restrict__ pick, int ny, int nx, const float* __restrict__ back, float* dback, float* gcp, float* gu,
                                      int staged, double thresh_sigma, float* stats) {
  DET_DYN_SMEM(float, sm);
  const int n = ny * nx, x = blockIdx.x * blockDim.x + threadIdx.x;
  float* cp = staged ? sm + threadIdx.x : gcp + x;
  float* u = staged ? sm + blockDim.x * ny + threadIdx.x : gu + x;
  const int ss = staged ? (int)blockDim.x : nx;
  if (x == 0) {
    const float gback = (n & 1) ? pick[1] : (pick[0] + pick[1]) * 0.5f;
    const float grms = (n & 1) ? pick[3] : (pick[2] + pick[3]) * 0.5f;
    stats[0] = gback;
    stats[1] = grms;
    stats[2] = (float)(thresh_sigma * (double)grms);
  }
  if (x < nx) det_spline_d2(back + x, nx, ny, dback + x, nx, cp, u, ss);
}

// ---- B3 --------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float det_spline_eval(float lo, float hi, float dlo, float dhi, float t) {
  const float ct = 1.0f - t;
  return ((ct * lo + t * hi) + ((ct * ct) * ct - ct) * dlo) + ((t * t) * t - t) * dhi;
}
__device__ __forceinline__ void det_spline_pos(int i, int n, int* lo, float* t) {
  const float u = ((float)i + 0.5f) / (float)DET_BW - 0.5f;
  int l = (int)floorf(u);
  l = l < 0 ? 0 : (l > n - 2 ? n - 2 : l);
  *lo = l;
  *t = u - (float)l;
}

// one thread per image row: the mesh map interpolated along y at this row (node), then its x-spline (dnode)
__global__ void det_nodes_kernel(const float* __restrict__ back, const float* __restrict__ dback, const DetGeom G, float* node,
                                 float* dnode, float* gcp, float* gu, int staged) {
  DET_DYN_SMEM(float, sm);
  const int y = blockIdx.x * blockDim.x + threadIdx.x;  // region row
  if (y >= G.RH) return;
  const int ny = G.ny, nx = G.nx;
  float* cp = staged ? sm + threadIdx.x : gcp + (long long)y * nx;
  float* u = staged ? sm + blockDim.x * nx + threadIdx.x : gu + (long long)y * nx;
  const int ss = staged ? (int)blockDim.x : 1;
  float* nd = node + (long long)y * nx;
  if (ny > 1) {
    int yl;
    float t;
    det_spline_pos(G.gy0 + y, ny, &yl, &t);
    for (int x = 0; x < nx; ++x)
      nd[x] = det_spline_eval(back[yl * nx + x], back[(yl + 1) * nx + x], dback[yl * nx + x], dback[(yl + 1) * nx + x], t);
  } else {
    for (int x = 0; x < nx; ++x) nd[x] = back[x];
  }
  det_spline_d2(nd, 1, nx, dnode + (long long)y * nx, 1, cp, u, ss);
}

__global__ void det_foreground_kernel(const double* __restrict__ band, const float* __restrict__ node, const float* __restrict__ dnode,
                                      const DetGeom G, float* __restrict__ fg) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;  // region pixel
  const int W = G.RW, nx = G.nx;
  if (x >= W) return;
  const float* nd = node + (long long)y * nx;
  float b;
  if (nx > 1) {
    const float* dn = dnode + (long long)y * nx;
    int xl;
    float t;
    det_spline_pos(G.gx0 + x, nx, &xl, &t);
    b = det_spline_eval(nd[xl], nd[xl + 1], dn[xl], dn[xl + 1], t);
  } else {
    b = nd[0];
  }
  const long long p = (long long)y * W + x;
  fg[p] = (float)(band[p] - (double)b);
}

// ---- E1 --------------------------------------------------------------------------------------------------------------
// 32 x 32 outputs per CTA from a zero-padded shared tile; taps accumulated in raster order (float32, one rounding per product
// and per sum).  Writes the filtered image and the initial label image: own raster index above the threshold, -1 below.
__global__ void __launch_bounds__(256) det_filter_kernel(const float* __restrict__ fg, const DetGeom G, int kh, int kw, const DetTaps taps,
                                                         const float* __restrict__ stats, float* __restrict__ conv, int* __restrict__ label) {
  DET_DYN_SMEM(float, sm);
  const int H = G.RH, W = G.RW;
  const int tw = 32 + kw - 1, th = 32 + kh - 1;
  const int x0 = blockIdx.x * 32 - kw / 2, y0 = blockIdx.y * 32 - kh / 2;
  for (int i = threadIdx.x; i < tw * th; i += blockDim.x) {
    const int ty = i / tw, tx = i - ty * tw;
    const int y = y0 + ty, x = x0 + tx;
    sm[i] = (y >= 0 && y < H && x >= 0 && x < W) ? fg[(long long)y * W + x] : 0.f;
  }
  __syncthreads();
  const float thresh = stats[2];
  const int lx = threadIdx.x & 31;
  for (int ly = threadIdx.x >> 5; ly < 32; ly += 8) {
    const int x = blockIdx.x * 32 + lx, y = blockIdx.y * 32 + ly;
    if (x >= W || y >= H) continue;
    float acc = 0.f;
    for (int ky = 0; ky < kh; ++ky)
      for (int kx = 0; kx < kw; ++kx) acc = acc + taps.v[ky * kw + kx] * sm[(ly + ky) * tw + lx + kx];
    const long long p = (long long)y * W + x;
    conv[p] = acc;
    // on a region edge inside the field the taps beyond the region are missing: those pixels carry no label (their objects belong to a neighbour)
    label[p] = (acc > thresh && y >= G.vy0 && y < G.vy1 && x >= G.vx0 && x < G.vx1) ? (int)p : -1;
  }
}

// the 7x7 mask of the reference (and any other K x K one instantiated below): a thread keeps FOUR vertically adjacent outputs and
// streams the K + 3 tile rows they share through registers (K shared-memory loads per row instead of 4 K); every output still adds
// its taps in raster order, one rounding per product and per sum, so the result is bit-identical to det_filter_kernel
template <int K>
__global__ void __launch_bounds__(256) det_filter_tiled_kernel(const float* __restrict__ fg, const DetGeom G, const DetTaps taps,
                                                               const float* __restrict__ stats, float* __restrict__ conv, int* __restrict__ label) {
  DET_DYN_SMEM(float, sm);
  const int H = G.RH, W = G.RW;
  constexpr int tw = 32 + K - 1, th = 32 + K - 1;
  const int x0 = blockIdx.x * 32 - K / 2, y0 = blockIdx.y * 32 - K / 2;
  for (int i = threadIdx.x; i < tw * th; i += blockDim.x) {
    const int ty = i / tw, tx = i - ty * tw;
    const int y = y0 + ty, x = x0 + tx;
    sm[i] = (y >= 0 && y < H && x >= 0 && x < W) ? fg[(long long)y * W + x] : 0.f;
  }
  __syncthreads();
  const float thresh = stats[2];
  const int lx = threadIdx.x & 31, ly0 = (threadIdx.x >> 5) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < K + 3; ++j) {
    float v[K];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) v[kx] = sm[(ly0 + j) * tw + lx + kx];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ky = j - r;
      if (ky >= 0 && ky < K) {
#pragma unroll
        for (int kx = 0; kx < K; ++kx) acc[r] = acc[r] + taps.v[ky * K + kx] * v[kx];
      }
    }
  }
  const int x = blockIdx.x * 32 + lx;
  if (x >= W) return;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int y = blockIdx.y * 32 + ly0 + r;
    if (y >= H) break;
    const long long p = (long long)y * W + x;
    conv[p] = acc[r];
    label[p] = (acc[r] > thresh && y >= G.vy0 && y < G.vy1 && x >= G.vx0 && x < G.vx1) ? (int)p : -1;
  }
}

// ---- E2: union-find on the label image ------------------------------------------------------------------------------------
__device__ __forceinline__ int det_find(int* L, int i) {
  volatile int* V = L;
  int p = V[i];
  while (p != i) {
    i = p;
    p = V[i];
  }
  return i;
}
__device__ void det_unite(int* L, int a, int b) {
  bool done;
  do {
    a = det_find(L, a);
    b = det_find(L, b);
    if (a < b) {
      const int old = atomicMin(&L[b], a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(&L[a], b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// links to the already-scanned neighbours W, NW, N, NE; NW is implied when W or N is set, NE when N is set
__global__ void det_ccl_merge_kernel(int* L, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int p = y * W + x;
  if (L[p] < 0) return;
  const bool w = x > 0 && L[p - 1] >= 0;
  const bool n = y > 0 && L[p - W] >= 0;
  const bool nw = x > 0 && y > 0 && L[p - W - 1] >= 0;
  const bool ne = x + 1 < W && y > 0 && L[p - W + 1] >= 0;
  if (w) det_unite(L, p, p - 1);
  if (n) det_unite(L, p, p - W);
  if (nw && !w && !n) det_unite(L, p, p - W - 1);
  if (ne && !n) det_unite(L, p, p - W + 1);
}

// every pixel takes its root as label (the trees are final once the merge kernel has finished) and adds itself to the root's
// statistics: pixel count, largest raster index, bounding columns — integer atomics, order-independent
__global__ void det_ccl_flatten_stats_kernel(int* L, const DetGeom G, int* npix, int* last, int* xmin, int* xmax, int* touch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const int W = G.RW;
  if (x >= W) return;
  const int p = y * W + x;
  if (L[p] < 0) return;
  const int r = det_find(L, p);
  L[p] = r;
  atomicAdd(&npix[r], 1);
  atomicMax(&last[r], p);
  atomicMin(&xmin[r], x);
  atomicMax(&xmax[r], x);
  // a pixel on the rim of the exact area, on a side that is not the field's edge: the object may continue beyond what this call sees
  if ((y == G.vy0 && G.vy0 > 0) || (y == G.vy1 - 1 && G.vy1 < G.RH) || (x == G.vx0 && G.vx0 > 0) || (x == G.vx1 - 1 && G.vx1 < G.RW)) touch[r] = 1;
}

// ---- E3 --------------------------------------------------------------------------------------------------------------
__global__ void det_mark_kernel(const int* __restrict__ L, long long n, const int* __restrict__ npix, const int* __restrict__ last, int minarea,
                                const DetGeom G, unsigned char* flag) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (L[p] == (int)p && npix[p] >= minarea) {
    const int e = last[p], ey = G.gy0 + e / G.RW, ex = G.gx0 + e % G.RW;
    if (ey >= G.oy0 && ey < G.oy1 && ex >= G.ox0 && ex < G.ox1) flag[e] = 1;
  }
}

// flags are compacted in raster order: per-CTA counts, one-CTA exclusive scan, scatter
__global__ void __launch_bounds__(256) det_count_kernel(const unsigned char* __restrict__ flag, long long n, int* cnt) {
  const long long base = (long long)blockIdx.x * DET_CHUNK + threadIdx.x * 16;
  int c = 0;
  for (int j = 0; j < 16; ++j)
    if (base + j < n) c += flag[base + j];
  __shared__ int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  if (c) atomicAdd(&s, c);
  __syncthreads();
  if (threadIdx.x == 0) cnt[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024) det_scan_kernel(const int* __restrict__ cnt, int nblk, int* off, int* n_out) {
  __shared__ int buf[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblk ? cnt[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      const int a = threadIdx.x >= d ? buf[threadIdx.x - d] : 0;
      __syncthreads();
      buf[threadIdx.x] += a;
      __syncthreads();
    }
    if (i < nblk) off[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) n_out[0] = carry;
}

__global__ void __launch_bounds__(256) det_scatter_kernel(const unsigned char* __restrict__ flag, long long n, const int* __restrict__ off,
                                                          long long max_objects, int* endpos) {
  __shared__ int buf[256];
  const long long base = (long long)blockIdx.x * DET_CHUNK + threadIdx.x * 16;
  int c = 0;
  for (int j = 0; j < 16; ++j)
    if (base + j < n) c += flag[base + j];
  buf[threadIdx.x] = c;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    const int a = threadIdx.x >= d ? buf[threadIdx.x - d] : 0;
    __syncthreads();
    buf[threadIdx.x] += a;
    __syncthreads();
  }
  long long k = (long long)off[blockIdx.x] + buf[threadIdx.x] - c;
  for (int j = 0; j < 16; ++j)
    if (base + j < n && flag[base + j]) {
      if (k < max_objects) endpos[k] = (int)(base + j);
      ++k;
    }
}

// ---- E4 --------------------------------------------------------------------------------------------------------------
// one 32-thread CTA per object (grid-stride over the objects): the lanes fetch 32 pixels of a bounding-box row at a time (coalesced,
// all loads in flight together) into shared memory, lane 0 adds the member pixels in raster order — the oracle's sequential double sums
// (np.cumsum(...)[-1]) without a memory round trip per pixel
__global__ void __launch_bounds__(32) det_moments_kernel(const int* __restrict__ L, const float* __restrict__ fg, const float* __restrict__ conv,
                                                         const DetGeom G, const int* __restrict__ endpos, const int* __restrict__ n_found,
                                                         long long max_objects, const int* __restrict__ npix, const int* __restrict__ xmin,
                                                         const int* __restrict__ xmax, const int* __restrict__ touch, int cy, int cx, double* xy,
                                                         double* centres, int* npix_out, long long* last_out, int* flags) {
  __shared__ float s_v[32], s_c[32];
  __shared__ int s_in[32];
  const long long n = n_found[0] < max_objects ? n_found[0] : max_objects;
  const int W = G.RW, lane = threadIdx.x;
  for (long long k = blockIdx.x; k < n; k += gridDim.x) {
    const int e = endpos[k], r = L[e];
    const int y0 = r / W, y1 = e / W, x0 = xmin[r], x1 = xmax[r];
    double tv = 0.0, mx = 0.0, my = 0.0, tc = 0.0, cmx = 0.0, cmy = 0.0;
    for (int y = y0; y <= y1; ++y)
      for (int xb = x0; xb <= x1; xb += 32) {
        const int x = xb + lane;
        const int p = y * W + x;
        const int in = (x <= x1) && (L[p] == r);
        s_in[lane] = in;
        s_v[lane] = in ? fg[p] : 0.f;
        s_c[lane] = in ? conv[p] : 0.f;
        __syncthreads();
        if (lane == 0) {
          const double dy = (double)(y - y0);
          for (int l = 0; l < 32; ++l) {
            if (!s_in[l]) continue;
            const double v = (double)s_v[l], c = (double)s_c[l];
            const double dx = (double)(xb + l - x0);
            tv += v;
            mx += v * dx;
            my += v * dy;
            tc += c;
            cmx += c * dx;
            cmy += c * dy;
          }
        }
        __syncthreads();
      }
    if (lane == 0) {
      if (!(tv > 0.0)) { tv = tc; mx = cmx; my = cmy; }  // faint detections: weight with the filtered values (> thresh > 0)
      const double X = mx / tv + (double)(x0 + G.gx0), Y = my / tv + (double)(y0 + G.gy0);  // field coordinates
      xy[2 * k] = X;
      xy[2 * k + 1] = Y;
      centres[2 * k] = rint(Y - (double)cy);      // (row, col) offsets from the field centre, detection.py:48-54
      centres[2 * k + 1] = rint(X - (double)cx);
      npix_out[k] = npix[r];
      last_out[k] = (long long)(G.gy0 + e / W) * G.W + (G.gx0 + e % W);  // the order key, in the raster of the whole field
      if (touch[r]) flags[0] = 1;  // an object of this call's tile leaves the area it can see: the caller must detect on the assembled field
    }
  }
}

struct DetLayout {
  size_t band, fg, conv, label, npix, last, xmin, xmax, touch, flag, mesh, rows, cnt, off, endpos, last64, flags, total;
  int ny, nx, nblk;
};
static DetLayout det_layout(long long H, long long W, long long RH, long long RW, long long max_objects) {
  DetLayout L{};
  const size_t n = (size_t)RH * RW;
  L.ny = (int)((H - 1) / DET_BW + 1);
  L.nx = (int)((W - 1) / DET_BW + 1);
  L.nblk = (int)((n + DET_CHUNK - 1) / DET_CHUNK);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
  L.band = take(n * 8);
  L.fg = take(n * 4);
  L.conv = take(n * 4);
  L.label = take(n * 4);
  L.npix = take(n * 4);
  L.last = take(n * 4);
  L.xmin = take(n * 4);
  L.xmax = take(n * 4);
  L.touch = take(n * 4);
  L.flag = take(n);
  L.mesh = take((size_t)L.ny * L.nx * 4 * 9 + 16);  // back0 sig0 back1 sig1 back sig dback cp u, then the 4 median picks
  L.rows = take((size_t)RH * L.nx * 4 * 4);        // node dnode cp u
  L.cnt = take((size_t)L.nblk * 4);
  L.off = take((size_t)L.nblk * 4);
  L.endpos = take((size_t)max_objects * 4);
  L.last64 = take((size_t)max_objects * 8);
  L.flags = take(16);
  L.total = o;
  return L;
}

static DetGeom det_geom(long long H, long long W, long long RH, long long RW, long long gy0, long long gx0, int kh, int kw, const DetLayout& Y) {
  DetGeom G{};
  G.H = (int)H; G.W = (int)W; G.RH = (int)RH; G.RW = (int)RW; G.gy0 = (int)gy0; G.gx0 = (int)gx0;
  G.ny = Y.ny; G.nx = Y.nx;
  G.vy0 = gy0 > 0 ? kh / 2 : 0;
  G.vy1 = (int)RH - (gy0 + RH < H ? kh / 2 : 0);
  G.vx0 = gx0 > 0 ? kw / 2 : 0;
  G.vx1 = (int)RW - (gx0 + RW < W ? kw / 2 : 0);
  G.oy0 = 0; G.oy1 = (int)H; G.ox0 = 0; G.ox1 = (int)W;
  return G;
}

}  // namespace dbv

using namespace dbv;

extern "C" int64_t dbv_detect_scratch_bytes_region(int64_t H, int64_t W, int64_t RH, int64_t RW, int64_t max_objects) {
  if (H <= 0 || W <= 0 || RH <= 0 || RW <= 0 || RH > H || RW > W || max_objects <= 0) return 0;
  return (int64_t)det_layout(H, W, RH, RW, max_objects).total;
}
extern "C" int64_t dbv_detect_scratch_bytes(int64_t H, int64_t W, int64_t max_objects) {
  return dbv_detect_scratch_bytes_region(H, W, H, W, max_objects);
}

// phase A: the band plane of the region and the statistics of every mesh that lies wholly inside it, written at the mesh's place
// in the caller's field-sized (ny, nx) maps (other entries untouched: a tiled caller presets -inf and max-reduces over the ranks)
extern "C" int dbv_detect_meshes(const void* region, int dtype, int64_t RH, int64_t RW, int64_t pitch, int C, int band, int64_t gy0,
                                 int64_t gx0, int64_t H, int64_t W, int64_t max_objects, void* scratch, int64_t scratch_bytes,
                                 float* back0, float* sig0, void* stream) {
  DBV_REQUIRE(region && scratch && back0 && sig0, "dbv_detect_meshes: null pointer");
  DBV_REQUIRE(dtype == DBV_F64 || dtype == DBV_F32, "dbv_detect_meshes: bad dtype %d", dtype);
  DBV_REQUIRE(H > 0 && W > 0 && RH > 0 && RW > 0 && gy0 >= 0 && gx0 >= 0 && gy0 + RH <= H && gx0 + RW <= W && H * W < (1ll << 31) && pitch >= RW &&
                  C > 0 && band >= 0 && band < C && max_objects > 0,
              "dbv_detect_meshes: bad geometry");
  const DetLayout Y = det_layout(H, W, RH, RW, max_objects);
  DBV_REQUIRE(scratch_bytes >= (int64_t)Y.total, "dbv_detect_meshes: scratch too small (%lld < %lld bytes)", (long long)scratch_bytes, (long long)Y.total);
  DBV_REQUIRE(((uintptr_t)scratch & 255) == 0, "dbv_detect_meshes: scratch must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  (void)st;
  double* d_band = (double*)((char*)scratch + Y.band);
  const DetGeom G = det_geom(H, W, RH, RW, gy0, gx0, 1, 1, Y);
  const long long n = RH * RW;
  const int gb = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  if (dtype == DBV_F64) DET_LAUNCH_NOSYNC(det_band_kernel<double>, gb, 256, 0, (const double*)region, RH, RW, pitch, C, band, d_band);
  else DET_LAUNCH_NOSYNC(det_band_kernel<float>, gb, 256, 0, (const float*)region, RH, RW, pitch, C, band, d_band);
#ifndef DBV_EMULATE
  // 20 KB of static shared memory per 64-thread CTA of mostly serial work: ask for the largest shared-memory carve-out (11 CTAs per SM)
  DBV_CUDA(cudaFuncSetAttribute(det_mesh_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
#endif
  DET_LAUNCH(det_mesh_kernel, dim3(Y.nx, Y.ny), DET_BW, 0, d_band, G, back0, sig0);
  return DBV_OK;
}

// phase B: everything after the mesh statistics, on the region whose band plane phase A left in `scratch`; back0 / sig0 = the
// COMPLETE field-sized mesh maps.  Objects are reported by the call whose owner tile [oy0, oy1) x [ox0, ox1) holds their last pixel;
// flags[0] is set when such an object reaches the rim of what the region can see (the caller then detects on the assembled field).
extern "C" int dbv_detect_objects(int64_t RH, int64_t RW, int64_t gy0, int64_t gx0, int64_t H, int64_t W, const float* back0, const float* sig0,
                                  const float* taps, int kh, int kw, double thresh_sigma, int minarea, int cy, int cx, int64_t oy0,
                                  int64_t oy1, int64_t ox0, int64_t ox1, int64_t max_objects, void* scratch, int64_t scratch_bytes,
                                  int32_t* n_found, double* xy, double* centres, int32_t* npix_out, int64_t* last_out, int32_t* flags,
                                  float* stats, void* stream) {
  DBV_REQUIRE(back0 && sig0 && taps && scratch && n_found && xy && centres && npix_out && last_out && flags && stats, "dbv_detect_objects: null pointer");
  DBV_REQUIRE(H > 0 && W > 0 && RH > 0 && RW > 0 && gy0 >= 0 && gx0 >= 0 && gy0 + RH <= H && gx0 + RW <= W && H * W < (1ll << 31),
              "dbv_detect_objects: bad geometry");
  DBV_REQUIRE(kh > 0 && kw > 0 && (kh & 1) && (kw & 1) && kh <= DET_MAXK && kw <= DET_MAXK, "dbv_detect_objects: the filter mask must be odd-sized, at most %d x %d", DET_MAXK, DET_MAXK);
  DBV_REQUIRE(minarea >= 1 && max_objects > 0 && thresh_sigma > 0.0, "dbv_detect_objects: bad minarea / max_objects / thresh");
  const DetLayout Y = det_layout(H, W, RH, RW, max_objects);
  DBV_REQUIRE(scratch_bytes >= (int64_t)Y.total, "dbv_detect_objects: scratch too small (%lld < %lld bytes)", (long long)scratch_bytes, (long long)Y.total);
  DBV_REQUIRE(((uintptr_t)scratch & 255) == 0, "dbv_detect_objects: scratch must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  (void)st;
  char* base = (char*)scratch;
  const long long n = RH * RW;
  const int ny = Y.ny, nx = Y.nx, nm = ny * nx;
  DetGeom G = det_geom(H, W, RH, RW, gy0, gx0, kh, kw, Y);
  G.oy0 = (int)oy0; G.oy1 = (int)oy1; G.ox0 = (int)ox0; G.ox1 = (int)ox1;
  double* d_band = (double*)(base + Y.band);
  float* d_fg = (float*)(base + Y.fg);
  float* d_conv = (float*)(base + Y.conv);
  int* d_label = (int*)(base + Y.label);
  int* d_npix = (int*)(base + Y.npix);
  int* d_last = (int*)(base + Y.last);
  int* d_xmin = (int*)(base + Y.xmin);
  int* d_xmax = (int*)(base + Y.xmax);
  int* d_touch = (int*)(base + Y.touch);
  unsigned char* d_flag = (unsigned char*)(base + Y.flag);
  float* m = (float*)(base + Y.mesh);
  float *back1 = m + 2 * nm, *sig1 = m + 3 * nm, *back = m + 4 * nm, *sig = m + 5 * nm, *dback = m + 6 * nm, *mcp = m + 7 * nm, *mu = m + 8 * nm,
        *pick = m + 9 * nm;
  float* rw = (float*)(base + Y.rows);
  const size_t hn = (size_t)RH * nx;
  float *node = rw, *dnode = rw + hn, *rcp = rw + 2 * hn, *ru = rw + 3 * hn;
  int* d_cnt = (int*)(base + Y.cnt);
  int* d_off = (int*)(base + Y.off);
  int* d_end = (int*)(base + Y.endpos);
  DetTaps tp{};
  for (int i = 0; i < kh * kw; ++i) tp.v[i] = taps[i];

  const unsigned gm = (unsigned)((nm + 127) / 128);
  DET_LAUNCH_NOSYNC(det_mesh_fill_kernel, gm, 128, 0, back0, sig0, ny, nx, back1, sig1);
  DET_LAUNCH_NOSYNC(det_mesh_median_kernel, gm, 128, 0, back1, sig1, ny, nx, back, sig);
  const int staged = nm <= 10240;
  DET_LAUNCH(det_mesh_rank_kernel, dim3(gm, 2), 128, staged ? sizeof(float) * nm : 0, back, sig, nm, staged, pick);
  // the spline recurrences keep their scratch in shared memory ([k][thread], 32 threads per CTA) whenever it fits in 48 KB
  const int st_y = ny <= 192, st_x = nx <= 192;
  DET_LAUNCH_NOSYNC(det_mesh_final_kernel, (unsigned)((nx + 31) / 32), 32, st_y ? sizeof(float) * 2 * 32 * ny : 0, pick, ny, nx, back, dback, mcp, mu, st_y, thresh_sigma,
             stats);
  DET_LAUNCH_NOSYNC(det_nodes_kernel, (unsigned)((RH + 31) / 32), 32, st_x ? sizeof(float) * 2 * 32 * nx : 0, back, dback, G, node, dnode, rcp, ru, st_x);
  const dim3 grow((unsigned)((RW + 255) / 256), (unsigned)RH);
  DET_LAUNCH_NOSYNC(det_foreground_kernel, grow, 256, 0, d_band, node, dnode, G, d_fg);
  const size_t fsm = sizeof(float) * (32 + kw - 1) * (32 + kh - 1);
  const dim3 gf((unsigned)((RW + 31) / 32), (unsigned)((RH + 31) / 32));
  if (kh == 7 && kw == 7) DET_LAUNCH(det_filter_tiled_kernel<7>, gf, 256, fsm, d_fg, G, tp, stats, d_conv, d_label);
  else DET_LAUNCH(det_filter_kernel, gf, 256, fsm, d_fg, G, kh, kw, tp, stats, d_conv, d_label);
  DET_LAUNCH_NOSYNC(det_ccl_merge_kernel, grow, 256, 0, d_label, (int)RH, (int)RW);
  const unsigned gn = (unsigned)((n + 255) / 256);
  DET_MEMSET(d_npix, 0, n * 4);
  DET_MEMSET(d_last, 0xFF, n * 4);
  DET_MEMSET(d_xmax, 0xFF, n * 4);
  DET_MEMSET(d_xmin, 0x7F, n * 4);
  DET_MEMSET(d_touch, 0, n * 4);
  DET_MEMSET(d_flag, 0, n);
  DET_MEMSET(flags, 0, 4);
  DET_LAUNCH_NOSYNC(det_ccl_flatten_stats_kernel, grow, 256, 0, d_label, G, d_npix, d_last, d_xmin, d_xmax, d_touch);
  DET_LAUNCH_NOSYNC(det_mark_kernel, gn, 256, 0, d_label, n, d_npix, d_last, minarea, G, d_flag);
  DET_LAUNCH(det_count_kernel, Y.nblk, 256, 0, d_flag, n, d_cnt);
  DET_LAUNCH(det_scan_kernel, 1, 1024, 0, d_cnt, Y.nblk, d_off, n_found);
  DET_LAUNCH(det_scatter_kernel, Y.nblk, 256, 0, d_flag, n, d_off, max_objects, d_end);
  DET_LAUNCH(det_moments_kernel, (unsigned)(max_objects < 148 * 32 ? max_objects : 148 * 32), 32, 0, d_label, d_fg, d_conv, G, d_end, n_found, max_objects, d_npix, d_xmin,
             d_xmax, d_touch, cy, cx, xy, centres, npix_out, (long long*)last_out, flags);
  return DBV_OK;
}

// the whole field in one call (one GPU): phase A into the scratch's own mesh maps, then phase B with the field as its own owner tile
extern "C" int dbv_detect(const void* field, int dtype, int64_t H, int64_t W, int64_t pitch, int C, int band, const float* taps, int kh,
                          int kw, double thresh_sigma, int minarea, int cy, int cx, int64_t max_objects, void* scratch,
                          int64_t scratch_bytes, int32_t* n_found, double* xy, double* centres, int32_t* npix_out, float* stats,
                          void* stream) {
  DBV_REQUIRE(field && scratch, "dbv_detect: null pointer");
  DBV_REQUIRE(H > 0 && W > 0 && max_objects > 0, "dbv_detect: bad field geometry");
  const DetLayout Y = det_layout(H, W, H, W, max_objects);
  DBV_REQUIRE(scratch_bytes >= (int64_t)Y.total, "dbv_detect: scratch too small (%lld < %lld bytes)", (long long)scratch_bytes, (long long)Y.total);
  char* base = (char*)scratch;
  float* m = (float*)(base + Y.mesh);
  const int nm = Y.ny * Y.nx;
  int r = dbv_detect_meshes(field, dtype, H, W, pitch, C, band, 0, 0, H, W, max_objects, scratch, scratch_bytes, m, m + nm, stream);
  if (r) return r;
  return dbv_detect_objects(H, W, 0, 0, H, W, m, m + nm, taps, kh, kw, thresh_sigma, minarea, cy, cx, 0, H, 0, W, max_objects, scratch,
                            scratch_bytes, n_found, xy, centres, npix_out, (int64_t*)(base + Y.last64), (int32_t*)(base + Y.flags), stats, stream);
}

// debug / parity access to the intermediate planes of the last call that used `scratch` (region-sized planes):
// what = 0 foreground f32 (RH,RW), 1 filtered f32 (RH,RW), 2 labels i32 (RH,RW), 3 mesh background f32 (ny,nx), 4 mesh sigma f32 (ny,nx),
// 5 raw mesh background, 6 raw mesh sigma (5 / 6: the scratch's own maps, filled by dbv_detect only).  Returns a device pointer inside scratch.
extern "C" const void* dbv_detect_plane_region(void* scratch, int64_t H, int64_t W, int64_t RH, int64_t RW, int64_t max_objects, int what) {
  if (!scratch || H <= 0 || W <= 0 || RH <= 0 || RW <= 0 || max_objects <= 0) return nullptr;
  const DetLayout Y = det_layout(H, W, RH, RW, max_objects);
  char* base = (char*)scratch;
  const size_t nm = (size_t)Y.ny * Y.nx;
  switch (what) {
    case 0: return base + Y.fg;
    case 1: return base + Y.conv;
    case 2: return base + Y.label;
    case 3: return base + Y.mesh + 4 * nm * 4;
    case 4: return base + Y.mesh + 5 * nm * 4;
    case 5: return base + Y.mesh;
    case 6: return base + Y.mesh + 1 * nm * 4;
  }
  return nullptr;
}
extern "C" const void* dbv_detect_plane(void* scratch, int64_t H, int64_t W, int64_t max_objects, int what) {
  return dbv_detect_plane_region(scratch, H, W, H, W, max_objects, what);
}
