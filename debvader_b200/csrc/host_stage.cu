// see host_stage.h
#include "host_stage.h"

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace dbv {

struct HostStagePool::Impl {
  std::vector<std::thread> workers;
  std::mutex m;
  std::condition_variable cv_go, cv_done;
  unsigned long long generation = 0;
  int pending = 0;
  bool stop = false;
  // the current job
  const void* src = nullptr;
  float* dst = nullptr;
  size_t n = 0;
  bool is_f64 = false;
  int parts = 1;
};

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

// Streaming (non-temporal) stores: the staging buffer is written once and read next by the DMA engine, so pulling its
// lines into the cache first (read-for-ownership) would only add a third of memory traffic.  i0 is a multiple of 1024
// elements and dst comes from cudaHostAlloc (page aligned): dst + i0 is 16-byte aligned.
static void convert_range(const void* src, bool is_f64, float* dst, size_t i0, size_t i1) {
  size_t i = i0;
#if defined(__SSE2__)
  if (is_f64) {
    const double* s = static_cast<const double*>(src);
    for (; i + 8 <= i1; i += 8) {
      const __m128 a = _mm_movelh_ps(_mm_cvtpd_ps(_mm_loadu_pd(s + i)), _mm_cvtpd_ps(_mm_loadu_pd(s + i + 2)));      // cvtpd2ps: MXCSR rounding
      const __m128 b = _mm_movelh_ps(_mm_cvtpd_ps(_mm_loadu_pd(s + i + 4)), _mm_cvtpd_ps(_mm_loadu_pd(s + i + 6)));  // = round to nearest even
      _mm_stream_ps(dst + i, a);
      _mm_stream_ps(dst + i + 4, b);
    }
    for (; i < i1; ++i) dst[i] = (float)s[i];
  } else {
    const float* s = static_cast<const float*>(src);
    for (; i + 16 <= i1; i += 16) {
      const __m128 a = _mm_loadu_ps(s + i), b = _mm_loadu_ps(s + i + 4), c = _mm_loadu_ps(s + i + 8), d = _mm_loadu_ps(s + i + 12);
      _mm_stream_ps(dst + i, a);
      _mm_stream_ps(dst + i + 4, b);
      _mm_stream_ps(dst + i + 8, c);
      _mm_stream_ps(dst + i + 12, d);
    }
    for (; i < i1; ++i) dst[i] = s[i];
  }
  _mm_sfence();
#else
  if (is_f64) {
    const double* s = static_cast<const double*>(src);
    for (; i < i1; ++i) dst[i] = (float)s[i];
  } else {
    memcpy(dst + i0, static_cast<const float*>(src) + i0, (i1 - i0) * sizeof(float));
  }
#endif
}

static void run_part(HostStagePool::Impl* p, int part) {
  // contiguous shares, multiples of 1024 elements so that no two threads write the same cache line
  const size_t per = ((p->n + p->parts - 1) / p->parts + 1023) / 1024 * 1024;
  const size_t i0 = std::min(p->n, per * (size_t)part), i1 = std::min(p->n, i0 + per);
  if (i1 > i0) convert_range(p->src, p->is_f64, p->dst, i0, i1);
}

HostStagePool::HostStagePool(int threads) : p_(new Impl), nthreads_(std::max(1, threads)) {
  p_->parts = nthreads_;
  for (int w = 1; w < nthreads_; ++w)  // the caller's thread is part 0
    p_->workers.emplace_back([this, w] {
      unsigned long long seen = 0;
      for (;;) {
        {
          std::unique_lock<std::mutex> lk(p_->m);
          p_->cv_go.wait(lk, [&] { return p_->stop || p_->generation != seen; });
          if (p_->stop) return;
          seen = p_->generation;
        }
        run_part(p_, w);
        {
          std::lock_guard<std::mutex> lk(p_->m);
          if (--p_->pending == 0) p_->cv_done.notify_one();
        }
      }
    });
}

HostStagePool::~HostStagePool() {
  {
    std::lock_guard<std::mutex> lk(p_->m);
    p_->stop = true;
  }
  p_->cv_go.notify_all();
  for (auto& t : p_->workers) t.join();
  delete p_;
}

void HostStagePool::convert(const void* src, bool is_f64, float* dst, size_t n) {
  if (nthreads_ == 1 || n < (size_t)1 << 16) {
    convert_range(src, is_f64, dst, 0, n);
    return;
  }
  {
    std::lock_guard<std::mutex> lk(p_->m);
    p_->src = src;
    p_->dst = dst;
    p_->n = n;
    p_->is_f64 = is_f64;
    p_->pending = nthreads_ - 1;
    ++p_->generation;
  }
  p_->cv_go.notify_all();
  run_part(p_, 0);
  std::unique_lock<std::mutex> lk(p_->m);
  p_->cv_done.wait(lk, [&] { return p_->pending == 0; });
}

int host_stage_default_threads() {
  if (const char* e = getenv("DEBVADER_B200_HOST_THREADS")) {  // a resource knob like OMP_NUM_THREADS, not an ablation switch
    const int v = atoi(e);
    if (v >= 1) return std::min(v, 64);
  }
  const unsigned hw = std::thread::hardware_concurrency();
  return (int)std::max(1u, std::min(8u, hw / 2));
}

}  // namespace dbv
