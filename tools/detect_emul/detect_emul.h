// Host emulation shim for csrc/detect_kernels.cu — TEST INFRASTRUCTURE (tools/detect_emul, tests/test_detect_emul.py), never part of
// libdebvader_b200.so.  With -DDBV_EMULATE the kernels of detect_kernels.cu compile as plain C++: a "launch" runs the CTAs one
// after the other, every CUDA thread of a CTA as a std::thread, __syncthreads() as a barrier, __shared__ as static storage,
// integer atomics as GCC atomics.  It exists to check the kernels' indexing, barriers and bit-exactness against
// oracle/detect_numpy.py in the container, which has no GPU; the GPU suite (tests/test_gpu_detect.py) checks the real thing.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>
#include <limits.h>

#define DBV_OK 0
#define DBV_ERR_INVALID -1
#define DBV_F32 0
#define DBV_F64 1

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct emu_idx { unsigned x, y, z; };
static thread_local emu_idx threadIdx, blockIdx;
static dim3 blockDim, gridDim;
static std::barrier<>* g_emu_bar = nullptr;
static thread_local char* g_emu_dynsmem = nullptr;

#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(x)
#define __shared__ static
#define __syncthreads() g_emu_bar->arrive_and_wait()

static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicMin(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

namespace dbv {
static inline int fail(int code, const char* fmt, ...) { fprintf(stderr, "dbv_detect (emulated): %s\n", fmt); return code; }
}
#define DBV_REQUIRE(cond, ...) do { if (!(cond)) return dbv::fail(DBV_ERR_INVALID, __VA_ARGS__); } while (0)

static inline void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  gridDim = grid;
  blockDim = block;
  const unsigned nt = block.x * block.y * block.z;
  std::vector<char> dyn(smem + 16);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar(nt);
        g_emu_bar = &bar;
        std::vector<std::thread> th;
        th.reserve(nt);
        for (unsigned t = 0; t < nt; ++t)
          th.emplace_back([&, t] {
            threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            blockIdx = {bx, by, bz};
            g_emu_dynsmem = dyn.data();
            body();
            bar.arrive_and_drop();  // a CUDA thread that has exited no longer takes part in __syncthreads()
          });
        for (auto& x : th) x.join();
      }
}
// kernels without __syncthreads(): the CUDA threads of a CTA run one after the other on the calling thread (a std::thread per pixel is slow)
static inline void emu_launch_seq(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  gridDim = grid;
  blockDim = block;
  const unsigned nt = block.x * block.y * block.z;
  std::vector<char> dyn(smem + 16);
  g_emu_dynsmem = dyn.data();
  g_emu_bar = nullptr;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned t = 0; t < nt; ++t) {
          threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
          blockIdx = {bx, by, bz};
          body();
        }
}
typedef void* cudaStream_t;
#define DET_LAUNCH(kernel, grid, block, smem, ...) emu_launch(dim3(grid), dim3(block), smem, [&] { kernel(__VA_ARGS__); })
#define DET_LAUNCH_NOSYNC(kernel, grid, block, smem, ...) emu_launch_seq(dim3(grid), dim3(block), smem, [&] { kernel(__VA_ARGS__); })
#define DET_MEMSET(p, v, n) memset(p, v, n)
#define DET_DYN_SMEM(T, name) T* name = (T*)g_emu_dynsmem
