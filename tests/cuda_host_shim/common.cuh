// Test infrastructure: a host stand-in for csrc/common.cuh + the CUDA execution model, just enough to compile the one-pass /
// CUDA-core kernels of csrc/dropout.cu as plain C++20 and RUN them on the CPU (tests/test_host_logic.py copies dropout.cu and
// dropmask.cuh next to this file and builds a shared library with g++).  One OS thread per CUDA thread of a block, blocks run one
// after another; __syncthreads is a block barrier, __shfl_xor_sync a per-warp exchange, atomicAdd an atomic_ref.  This checks the
// kernels' index arithmetic, shared-memory layouts, shuffles and atomics against the oracle where no GPU is at hand; it says
// nothing about performance, and the tcgen05 / mma.sync kernels cannot be run this way.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <barrier>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define MRB_OK 0
#define MRB_ERR_ARG (-1)
#define MRB_ERR_CUDA (-2)
#define MRB_ERR_UNSUPPORTED (-3)
#define MRB_DT_F16 0
#define MRB_DT_BF16 1
#define MRB_DT_F32 2

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __CUDACC_HOST_SHIM__ 1

struct uint3s { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
inline uint2 make_uint2(uint32_t a, uint32_t b) { return {a, b}; }
inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return {a, b, c, d}; }
inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr int cudaSuccess = 0;

inline thread_local uint3s threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

namespace shim {
struct Block {
  std::unique_ptr<std::barrier<>> all;
  std::vector<std::unique_ptr<std::barrier<>>> warp;
  std::vector<uint32_t> xchg;
};
inline Block* g_block = nullptr;
inline thread_local int t_linear = 0;

template <typename F>
void launch(dim3 grid, dim3 block, F&& body) {
  const int nthr = static_cast<int>(block.x * block.y * block.z);
  const int nwarp = (nthr + 31) / 32;
  Block blk;
  blk.all = std::make_unique<std::barrier<>>(nthr);
  for (int w = 0; w < nwarp; ++w) blk.warp.push_back(std::make_unique<std::barrier<>>(std::min(32, nthr - 32 * w)));
  blk.xchg.assign(nwarp * 32, 0u);
  g_block = &blk;
  std::vector<std::thread> pool;
  for (int t = 0; t < nthr; ++t) {
    pool.emplace_back([&, t] {
      t_linear = t;
      blockDim = block; gridDim = grid;
      threadIdx = {static_cast<unsigned>(t) % block.x, (static_cast<unsigned>(t) / block.x) % block.y, static_cast<unsigned>(t) / (block.x * block.y)};
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            blockIdx = {bx, by, bz};
            body();
            blk.all->arrive_and_wait();          // blocks run one after another (static __shared__ storage is reused)
          }
    });
  }
  for (auto& th : pool) th.join();
  g_block = nullptr;
}
}  // namespace shim

inline void __syncthreads() { shim::g_block->all->arrive_and_wait(); }
inline uint32_t shim_shfl_xor_bits(uint32_t v, int o) {
  shim::Block& b = *shim::g_block;
  const int w = shim::t_linear >> 5, l = shim::t_linear & 31;
  b.xchg[w * 32 + l] = v;
  b.warp[w]->arrive_and_wait();
  const uint32_t r = b.xchg[w * 32 + (l ^ o)];
  b.warp[w]->arrive_and_wait();
  return r;
}
inline float __shfl_xor_sync(unsigned, float v, int o) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u = shim_shfl_xor_bits(u, o);
  memcpy(&v, &u, 4);
  return v;
}
inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
using std::max;
using std::min;

namespace mrb {
inline void pdl_trigger() {}
inline void pdl_wait() {}
inline uint16_t bf16_rn(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fff;
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
inline uint16_t f16_rn(float f) { _Float16 h = static_cast<_Float16>(f); uint16_t u; memcpy(&u, &h, 2); return u; }
inline float f16_to_f(uint16_t u) { _Float16 h; memcpy(&h, &u, 2); return static_cast<float>(h); }
inline uint32_t pack2(float a, float b, int dt) {
  if (dt == MRB_DT_F16) return static_cast<uint32_t>(f16_rn(a)) | (static_cast<uint32_t>(f16_rn(b)) << 16);
  return static_cast<uint32_t>(bf16_rn(a)) | (static_cast<uint32_t>(bf16_rn(b)) << 16);
}
inline float unpack_lo(uint32_t w, int dt) { return dt == MRB_DT_F16 ? f16_to_f(static_cast<uint16_t>(w & 0xffff)) : __uint_as_float(w << 16); }
inline float unpack_hi(uint32_t w, int dt) { return dt == MRB_DT_F16 ? f16_to_f(static_cast<uint16_t>(w >> 16)) : __uint_as_float(w & 0xffff0000u); }
inline float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
inline float gelu_erf_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}
}  // namespace mrb

#define MRB_LAUNCH(kernel, grid, block, smem, stream, ...) shim::launch(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); })
#define MRB_CHECK_LAUNCH() do { } while (0)
inline int mrb_set_error(cudaError_t) { return MRB_ERR_CUDA; }
