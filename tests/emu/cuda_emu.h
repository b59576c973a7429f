// TEST INFRASTRUCTURE ONLY -- a small host emulation of the CUDA execution model, enough to run the device code of
// multimodal-sae_b200/csrc/kernels_*.cuh on the CPU (no GPU in the build container):
//   * every CUDA thread of a block is a host thread; blocks run one after the other;
//   * __syncthreads / __syncwarp and the *_sync warp collectives are std::barrier rendezvous (a thread that leaves the
//     kernel drops out of its barriers, like an exited CUDA thread);
//   * atomics take one global mutex; read-only / no-allocate loads are plain loads;
//   * fp16 / bf16 types and vector types come from the CUDA toolkit headers, which have host implementations.
// The kernel sources are compiled with `extern __shared__` -> `extern` and `__shared__` -> `static` (a text
// substitution done by the test), so shared memory is ordinary static storage shared by the block's host threads.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <vector_functions.h>
#include <vector_types.h>

#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <ctime>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#ifndef __global__
#define __global__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __host__
#define __host__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __grid_constant__
#define __grid_constant__
#endif


namespace emu {

struct Dim {
  unsigned x = 1, y = 1, z = 1;
};

struct Block {
  int nthreads = 0;
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<unsigned long long> slots;   // one exchange slot per thread (largest collective operand: 8 bytes)
};

inline Block* g_block = nullptr;
inline std::mutex g_atomic_mutex;
inline thread_local int t_tid = 0;

inline int warp_of(int tid) { return tid >> 5; }
inline int lanes_in_warp(int warp) {
  const int left = g_block->nthreads - warp * 32;
  return left < 32 ? left : 32;
}
inline void warp_rendezvous() { g_block->warp_bar[warp_of(t_tid)]->arrive_and_wait(); }

// all-to-all exchange inside the calling thread's warp: every lane publishes `v`, then reads through `pick`
template <typename T, typename F>
inline auto warp_exchange(T v, F&& pick) {
  static_assert(sizeof(T) <= sizeof(unsigned long long), "operand too large for the exchange slot");
  unsigned long long raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  g_block->slots[t_tid] = raw;
  warp_rendezvous();
  const int base = warp_of(t_tid) * 32;
  auto get = [&](int lane) {
    T out;
    std::memcpy(&out, &g_block->slots[base + lane], sizeof(T));
    return out;
  };
  auto result = pick(get, t_tid - base, lanes_in_warp(warp_of(t_tid)));
  warp_rendezvous();   // nobody overwrites a slot before every lane has read it
  return result;
}

// run `kernel()` for every thread of every block of the grid
inline void launch(Dim grid, Dim block, const std::function<void()>& kernel);

}  // namespace emu

// ---- built-in variables -------------------------------------------------------------------------------------------
inline thread_local emu::Dim threadIdx, blockIdx;
inline emu::Dim blockDim, gridDim;

inline void emu::launch(Dim grid, Dim block, const std::function<void()>& kernel) {
  gridDim = grid;
  blockDim = block;
  const int n = (int)(block.x * block.y * block.z);
  for (unsigned by = 0; by < grid.y; ++by) {
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      Block blk;
      blk.nthreads = n;
      blk.block_bar = std::make_unique<std::barrier<>>(n);
      for (int w = 0; w * 32 < n; ++w)
        blk.warp_bar.push_back(std::make_unique<std::barrier<>>(n - w * 32 < 32 ? n - w * 32 : 32));
      blk.slots.assign(n, 0);
      g_block = &blk;
      std::vector<std::thread> threads;
      threads.reserve(n);
      for (int t = 0; t < n; ++t) {
        threads.emplace_back([&, t, bx, by] {
          t_tid = t;
          threadIdx.x = (unsigned)t % block.x;
          threadIdx.y = ((unsigned)t / block.x) % block.y;
          threadIdx.z = (unsigned)t / (block.x * block.y);
          blockIdx.x = bx;
          blockIdx.y = by;
          kernel();
          // an exited thread no longer takes part in the collectives of its warp / block
          blk.warp_bar[emu::warp_of(t)]->arrive_and_drop();
          blk.block_bar->arrive_and_drop();
        });
      }
      for (auto& th : threads) th.join();
      g_block = nullptr;
    }
  }
}

// ---- synchronisation and warp collectives ---------------------------------------------------------------------------
inline void __syncthreads() { emu::g_block->block_bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_rendezvous(); }
inline void __threadfence() {}
inline void __threadfence_system() {}

template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  return emu::warp_exchange(v, [&](auto get, int lane, int n) { return (lane ^ lane_mask) < n ? get(lane ^ lane_mask) : v; });
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) {
  return emu::warp_exchange(v, [&](auto get, int, int n) { return get((src & 31) < n ? (src & 31) : 0); });
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
  return emu::warp_exchange(v, [&](auto get, int lane, int n) { return lane + (int)delta < n ? get(lane + (int)delta) : v; });
}
inline unsigned __ballot_sync(unsigned, int pred) {
  return emu::warp_exchange(pred ? 1 : 0, [&](auto get, int, int n) {
    unsigned m = 0;
    for (int l = 0; l < n; ++l) m |= get(l) ? (1u << l) : 0u;
    return m;
  });
}
inline int __reduce_add_sync(unsigned, int v) {
  return emu::warp_exchange(v, [&](auto get, int, int n) {
    int s = 0;
    for (int l = 0; l < n; ++l) s += get(l);
    return s;
  });
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  return emu::warp_exchange(v, [&](auto get, int, int n) {
    unsigned s = 0;
    for (int l = 0; l < n; ++l) s = get(l) > s ? get(l) : s;
    return s;
  });
}
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  return emu::warp_exchange(v, [&](auto get, int, int n) {
    unsigned s = 0xffffffffu;
    for (int l = 0; l < n; ++l) s = get(l) < s ? get(l) : s;
    return s;
  });
}

// ---- atomics ----------------------------------------------------------------------------------------------------------
template <typename T>
inline T atomicAdd(T* p, T v) {
  std::lock_guard<std::mutex> g(emu::g_atomic_mutex);
  const T old = *p;
  *p = old + v;
  return old;
}
inline float4 atomicAdd(float4* p, float4 v) {
  std::lock_guard<std::mutex> g(emu::g_atomic_mutex);
  const float4 old = *p;
  p->x += v.x;
  p->y += v.y;
  p->z += v.z;
  p->w += v.w;
  return old;
}
template <typename T>
inline T atomicExch(T* p, T v) {
  std::lock_guard<std::mutex> g(emu::g_atomic_mutex);
  const T old = *p;
  *p = v;
  return old;
}
template <typename T>
inline T atomicMax(T* p, T v) {
  std::lock_guard<std::mutex> g(emu::g_atomic_mutex);
  const T old = *p;
  if (v > old) *p = v;
  return old;
}
template <typename T>
inline T atomicCAS(T* p, T cmp, T v) {
  std::lock_guard<std::mutex> g(emu::g_atomic_mutex);
  const T old = *p;
  if (old == cmp) *p = v;
  return old;
}

// ---- scalar intrinsics ------------------------------------------------------------------------------------------------
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline void __trap() { std::abort(); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }

// the two streaming-load helpers of common.cuh that the emulated kernels use (inline PTX on the GPU)
namespace saeb {
using std::max;
using std::min;
inline float4 ldg_nc_f4(const float4* p) { return *p; }
inline uint4 ldg_nc_u4(const uint4* p) { return *p; }
}  // namespace saeb
