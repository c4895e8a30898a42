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
#include <condition_variable>
#include <functional>
#include <mutex>
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
#define MRB_HOST_SHIM 1

struct uint3s { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint2 { uint32_t x, y; };
struct alignas(8) float2 { float x, y; };
inline float2 make_float2(float a, float b) { return {a, b}; }
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
  std::vector<uint32_t> regs;        // [warp][32 lanes][8]: register exchange of the emulated mma.sync / ldmatrix
  std::atomic<int> count{0};         // __syncthreads_count
};
inline Block* g_block = nullptr;
inline thread_local int t_linear = 0;

// worker threads are created once and reused by every launch (thread creation dominated the run time of the engine-level tests)
struct Pool {
  std::vector<std::thread> th;
  std::mutex m;
  std::condition_variable cv, done_cv;
  std::function<void(int)> job;
  int nactive = 0, generation = 0, remaining = 0;
  void worker(int id, int seen) {
    for (;;) {
      bool mine;
      {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return generation != seen; });
        seen = generation;
        mine = id < nactive;
      }
      if (mine) {
        job(id);
        std::lock_guard<std::mutex> lk(m);
        if (--remaining == 0) done_cv.notify_one();
      }
    }
  }
  void run(int n, std::function<void(int)> f) {
    {
      std::lock_guard<std::mutex> lk(m);
      while (static_cast<int>(th.size()) < n) {
        const int id = static_cast<int>(th.size()), seen = generation;
        th.emplace_back([this, id, seen] { worker(id, seen); });
        th.back().detach();
      }
      job = std::move(f); nactive = n; remaining = n; ++generation;
    }
    cv.notify_all();
    std::unique_lock<std::mutex> lk(m);
    done_cv.wait(lk, [&] { return remaining == 0; });
  }
};
inline Pool& pool() { static Pool* p = new Pool(); return *p; }      // never destroyed: its threads are parked for good at exit

template <typename F>
void launch(dim3 grid, dim3 block, F&& body) {
  const int nthr = static_cast<int>(block.x * block.y * block.z);
  const int nwarp = (nthr + 31) / 32;
  Block blk;
  blk.all = std::make_unique<std::barrier<>>(nthr);
  for (int w = 0; w < nwarp; ++w) blk.warp.push_back(std::make_unique<std::barrier<>>(std::min(32, nthr - 32 * w)));
  blk.xchg.assign(nwarp * 32, 0u);
  blk.regs.assign(nwarp * 32 * 8, 0u);
  g_block = &blk;
  pool().run(nthr, [&](int t) {
    {
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
    }
  });
  g_block = nullptr;
}
}  // namespace shim

inline void __syncthreads() { shim::g_block->all->arrive_and_wait(); }
inline int __syncthreads_count(int pred) {
  shim::Block& b = *shim::g_block;
  if (pred) b.count.fetch_add(1);
  b.all->arrive_and_wait();
  const int r = b.count.load();
  b.all->arrive_and_wait();
  if (shim::t_linear == 0) b.count.store(0);
  b.all->arrive_and_wait();
  return r;
}
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline void __syncwarp() { shim::g_block->warp[shim::t_linear >> 5]->arrive_and_wait(); }
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
inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) {
  shim::Block& b = *shim::g_block;
  const int w = shim::t_linear >> 5, l = shim::t_linear & 31;
  b.xchg[w * 32 + l] = v;
  b.warp[w]->arrive_and_wait();
  const uint32_t r = b.xchg[w * 32 + (src & 31)];
  b.warp[w]->arrive_and_wait();
  return r;
}
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float __expf(float x) { return expf(x); }
template <typename K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return 0; }
constexpr int cudaFuncAttributeMaxDynamicSharedMemorySize = 8;

// dynamic shared memory: one buffer for the running block (the test strips `extern __shared__ ... smem_attn[]` declarations)
alignas(1024) inline uint8_t smem_attn[512 * 1024];

// 16-bit storage types of the kernels' templates
struct __half { uint16_t x; };
struct __nv_bfloat16 { uint16_t x; };
struct __half2 { uint16_t x, y; };
inline __half __ushort_as_half(unsigned short u) { return {u}; }
struct __nv_bfloat162 { uint16_t x, y; };

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
inline uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(static_cast<const uint8_t*>(p) - smem_attn); }
inline float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
inline float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <typename T> inline float to_f32(T v);
template <> inline float to_f32<__half>(__half v) { return f16_to_f(v.x); }
template <> inline float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __uint_as_float(static_cast<uint32_t>(v.x) << 16); }
template <> inline float to_f32<float>(float v) { return v; }
template <typename T> inline T from_f32(float v);
template <> inline __half from_f32<__half>(float v) { return {f16_rn(v)}; }
template <> inline __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return {bf16_rn(v)}; }
template <> inline float from_f32<float>(float v) { return v; }
inline float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
inline float gelu_erf_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}
inline void gelu_erf_both(float x, float& g, float& dg) { g = gelu_erf(x); dg = gelu_erf_grad(x); }
}  // namespace mrb

inline float __half2float(__half h) { return mrb::f16_to_f(h.x); }
inline __half2 __floats2half2_rn(float a, float b) { return {mrb::f16_rn(a), mrb::f16_rn(b)}; }
inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return {mrb::bf16_rn(a), mrb::bf16_rn(b)}; }

namespace shim {
inline uint32_t* warp_regs() { return g_block->regs.data() + (t_linear >> 5) * 32 * 8; }
inline void warp_sync() { g_block->warp[t_linear >> 5]->arrive_and_wait(); }
inline float h2f(uint32_t w, int hi, int dt) {
  const uint16_t h = hi ? static_cast<uint16_t>(w >> 16) : static_cast<uint16_t>(w & 0xffff);
  return dt == MRB_DT_F16 ? mrb::f16_to_f(h) : __uint_as_float(static_cast<uint32_t>(h) << 16);
}
// mma.sync.aligned.m16n8k16.row.col (PTX ISA fragment layouts): g = lane / 4, t = lane % 4;
//   A: a0 (g, 2t..2t+1) a1 (g+8, 2t..) a2 (g, 2t+8..) a3 (g+8, 2t+8..);  B: b0 (k 2t..2t+1, n g) b1 (k 2t+8.., n g);
//   C: c0 c1 (g, 2t..2t+1) c2 c3 (g+8, 2t..2t+1)
inline void mma_m16n8k16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1, int dt) {
  uint32_t* R = warp_regs();
  const int l = t_linear & 31, g = l >> 2, t = l & 3;
  for (int i = 0; i < 4; ++i) R[l * 8 + i] = a[i];
  R[l * 8 + 4] = b0; R[l * 8 + 5] = b1;
  warp_sync();
  float acc[4];
  for (int e = 0; e < 4; ++e) {
    const int row = g + (e >> 1) * 8, col = 2 * t + (e & 1);
    float s = 0.f;
    for (int k = 0; k < 16; ++k) {
      const uint32_t wa = R[((row & 7) * 4 + (k & 7) / 2) * 8 + (row >= 8 ? 1 : 0) + (k >= 8 ? 2 : 0)];
      const uint32_t wb = R[(col * 4 + (k & 7) / 2) * 8 + 4 + (k >= 8 ? 1 : 0)];
      s += h2f(wa, k & 1, dt) * h2f(wb, k & 1, dt);
    }
    acc[e] = s;
  }
  warp_sync();
  for (int e = 0; e < 4; ++e) c[e] += acc[e];
}
// ldmatrix.sync.aligned.m8n8.x4[.trans].shared.b16: lane l supplies the address of row l % 8 of matrix l / 8; lane l receives, per
// matrix, the word at (row l / 4, columns 2 (l % 4)..+1) -- with .trans the elements (2 (l % 4), l / 4) and (2 (l % 4) + 1, l / 4)
inline void ldsm_x4(uint32_t* r, uint32_t addr, bool trans) {
  uint32_t* R = warp_regs();
  const int l = t_linear & 31;
  R[l * 8] = addr;
  warp_sync();
  for (int m = 0; m < 4; ++m) {
    if (!trans) {
      memcpy(&r[m], smem_attn + R[(m * 8 + l / 4) * 8] + (l % 4) * 4, 4);
    } else {
      uint16_t lo, hi;
      memcpy(&lo, smem_attn + R[(m * 8 + 2 * (l % 4)) * 8] + (l / 4) * 2, 2);
      memcpy(&hi, smem_attn + R[(m * 8 + 2 * (l % 4) + 1) * 8] + (l / 4) * 2, 2);
      r[m] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
    }
  }
  warp_sync();
}
inline void cp_async16(uint32_t dst, const void* src, int bytes) {
  if (bytes) memcpy(smem_attn + dst, src, 16); else memset(smem_attn + dst, 0, 16);
}
}  // namespace shim

#define MRB_LAUNCH(kernel, grid, block, smem, stream, ...) shim::launch(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); })
#define MRB_CHECK_LAUNCH() do { } while (0)
inline int mrb_set_error(cudaError_t) { return MRB_ERR_CUDA; }
