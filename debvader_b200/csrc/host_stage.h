// Host side of dbv_deblend_host for PAGEABLE input (what a numpy array is): a small pool of threads copies — and, for
// float64 input, converts — a piece of the caller's buffer into the library's pinned staging memory, from where the H2D copy
// is a plain DMA transfer.  cudaMemcpyAsync straight from pageable memory is staged by the driver on ONE thread (~10 GB/s:
// 68 ms for the 684 MB of a 4096-stamp float64 batch, seven times the whole deblending pass).
#pragma once
#include <cstddef>
#include <cstdint>

namespace dbv {

class HostStagePool {
 public:
  explicit HostStagePool(int threads);
  ~HostStagePool();
  // dst[i] = (float)src[i] for i in [0, n): src float64 (is_f64) or float32; round-to-nearest-even like the device cast
  void convert(const void* src, bool is_f64, float* dst, size_t n);
  int threads() const { return nthreads_; }
  struct Impl;

 private:
  Impl* p_;
  int nthreads_;
};

int host_stage_default_threads();

}  // namespace dbv
