// TEST INFRASTRUCTURE (never part of the product library).
//
// A small host-side emulation of the CUDA execution model for the *portable SIMT subset* that
// anerf_b200/csrc/train_kernels.cuh is written in (thread/block indices, static __shared__ arrays,
// __syncthreads, float atomicAdd, __ldg; no warp intrinsics, no inline PTX).  It lets the host tests
// run the training (backward) kernels AND their launch sequence (train_path.cuh) on the CPU of the
// build container, where there is no GPU, and compare the gradients with the oracle's autograd:
// indexing, strides, buffer carve-up and the math are checked before a GPU is spent on them.
//
// Execution: the threads of one block are real host threads (so __syncthreads and shared-memory races
// are real); blocks run one after the other (static __shared__ storage is reused between them).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <barrier>
#include <functional>
#include <thread>
#include <vector>

#define ANERF_SIMT_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)

struct uint3_emu { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float4 { float x, y, z, w; };
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }

namespace simt_emu {
inline thread_local uint3_emu t_idx, b_idx;
inline dim3 g_block, g_grid;
inline std::barrier<>* g_bar = nullptr;
}  // namespace simt_emu
#define threadIdx simt_emu::t_idx
#define blockIdx simt_emu::b_idx
#define blockDim simt_emu::g_block
#define gridDim simt_emu::g_grid

inline void __syncthreads() { simt_emu::g_bar->arrive_and_wait(); }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float atomicAdd(float* addr, float v) {
  uint32_t* p = reinterpret_cast<uint32_t*>(addr);
  uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED), neu;
  float f;
  do {
    memcpy(&f, &old, 4);
    f += v;
    memcpy(&neu, &f, 4);
  } while (!__atomic_compare_exchange_n(p, &old, neu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  memcpy(&f, &old, 4);
  return f;
}
inline float __expf(float x) { return expf(x); }
inline float __fdividef(float a, float b) { return a / b; }

namespace simt_emu {
// kernel<<<grid, block>>>(args...)
template <typename K, typename... A>
void launch(dim3 grid, dim3 block, K kernel, A... args) {
  const unsigned nt = block.x * block.y * block.z;
  g_block = block;
  g_grid = grid;
  std::barrier<> bar((std::ptrdiff_t)nt);
  g_bar = &bar;
  std::vector<std::thread> th;
  th.reserve(nt);
  for (unsigned t = 0; t < nt; ++t) {
    th.emplace_back([=]() {
      t_idx.x = t % block.x;
      t_idx.y = (t / block.x) % block.y;
      t_idx.z = t / (block.x * block.y);
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            b_idx.x = bx; b_idx.y = by; b_idx.z = bz;
            kernel(args...);
            g_bar->arrive_and_wait();     // the next block reuses the static shared arrays
          }
    });
  }
  for (auto& x : th) x.join();
  g_bar = nullptr;
}
}  // namespace simt_emu
