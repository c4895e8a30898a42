// How fast can the softmax / elementwise warps of the attention kernels read tensor memory?  One CTA per SM allocates all 512
// columns; W warps (1 or 2 per 32-lane quadrant) sweep tcgen05.ld.32x32b over them REPS times; bytes / clk per SM from clock64.
// Decides whether the S / dP reads (64 KB per 128 x 128 fp32 tile) are a hard floor of the attention kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tmem_ld_probe tools/probe/tmem_ld_probe.cu && /tmp/tmem_ld_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int X>
__device__ __forceinline__ uint32_t ld_cols(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld_cols<32>(uint32_t taddr) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s ^= r[i];
  return s;
}
template <>
__device__ __forceinline__ uint32_t ld_cols<16>(uint32_t taddr) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= r[i];
  return s;
}

// INFLIGHT loads are issued before one tcgen05.wait::ld
template <int X, int INFLIGHT>
__global__ void __launch_bounds__(256, 1) probe(int warps, int reps, long long* clocks, uint32_t* sink) {
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = holder + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < warps) {
    for (int r = 0; r < reps; ++r) {
      for (int c = 0; c < 512; c += X * INFLIGHT) {
#pragma unroll
        for (int i = 0; i < INFLIGHT; ++i) acc ^= ld_cols<X>(base + ((c + i * X) & 511));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(holder), "r"(512) : "memory");
}

template <int X, int INFLIGHT>
static void run(int warps, int sms) {
  long long* d_clk; uint32_t* d_sink;
  cudaMalloc(&d_clk, sms * sizeof(long long)); cudaMalloc(&d_sink, 4);
  const int reps = 64;
  probe<X, INFLIGHT><<<sms, 256>>>(warps, 4, d_clk, d_sink);
  probe<X, INFLIGHT><<<sms, 256>>>(warps, reps, d_clk, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
  long long* h = new long long[sms];
  cudaMemcpy(h, d_clk, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += h[i];
  avg /= sms;
  const double bytes = static_cast<double>(warps) * reps * 512 * 32 * 4;     // per CTA: warps x reps x 512 columns x 32 lanes x 4 B
  printf("x%-3d in flight %d  warps %d | %9.0f clk | %7.1f B/clk/SM | %6.1f B/clk per warp\n", X, INFLIGHT, warps, avg, bytes / avg,
         bytes / avg / warps);
  delete[] h; cudaFree(d_clk); cudaFree(d_sink);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d; tcgen05.ld.32x32b of 512 columns x 32 lanes per warp and sweep\n", sms);
  for (int w : {1, 4, 8}) {
    run<32, 1>(w, sms);
    run<32, 2>(w, sms);
    run<32, 4>(w, sms);
    run<16, 4>(w, sms);
  }
  return 0;
}
