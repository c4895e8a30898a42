// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define MRB_OK 0
#define MRB_ERR_ARG (-1)
#define MRB_ERR_CUDA (-2)
#define MRB_ERR_UNSUPPORTED (-3)

#define MRB_DT_F16 0
#define MRB_DT_BF16 1
#define MRB_DT_F32 2

namespace mrb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// MRB_WAIT_HINT_NS (build-time): suspend-time hint of mbarrier.try_wait -- the waiting warp may stay descheduled up to that
// long instead of returning "not yet" at once, so spinning TMA / MMA / epilogue warps stop competing for issue slots.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#ifdef MRB_WAIT_HINT_NS
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(static_cast<uint32_t>(MRB_WAIT_HINT_NS))
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > 8000000000LL) {
      printf("mrb: mbarrier wait timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// True in exactly one lane of the converged warp (elect.sync).  Single-thread tcgen05 / TMA issue goes under `if (elect_one())`
// inside a loop the WHOLE warp runs: the compiler then knows that one thread is active and that the operands are warp-uniform;
// `if (lane == 0)` made it wrap every tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (profiles/ncu_gemm2_issue_r02d.md).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16/bf16 in, fp32 accumulate); single thread issues.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major operand tile in shared memory, 128-byte rows, SWIZZLE_128B (what TMA writes with
// CU_TENSOR_MAP_SWIZZLE_128B and a 64 x rows box of 16-bit elements):
//   start address >>4 | LBO=1 (ignored for swizzled K-major) | SBO = 8 rows * 128 B = 1024 B | version 1 | layout 2
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor: fp32 accumulate, A/B both K-major, fmt 0 = fp16, 1 = bf16.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int fmt, int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// 32 lanes x 32 consecutive fp32 columns: thread t gets row (lane base + t), columns [c, c+32)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- misc math
// Exact (erf) GELU, gelu(x) = x * Phi(x), with ONE MUFU: Phi(-|x|) = 2^q(|x|), q = degree-7 minimax-style fit of
// log2(Phi(-u)) on u in [0, 6] (Chebyshev nodes, fitted offline against scipy.special.log_ndtr; beyond 6 Phi(-u) < 1e-9):
//   gelu(x) = x - x * h  for x >= 0,   x * h  for x < 0,   h = Phi(-|x|)
// Max |error| vs the exact function over [-12, 12] in fp32 arithmetic: 6.3e-7 absolute, 1.2e-5 relative -- two orders of
// magnitude below the fp16 resolution of the fc1 output it feeds.  11 FMA-class ops + 1 MUFU (the previous A&S 7.1.26 erfc
// form needed 2 MUFU; the GELU epilogue of the short-K ViT fc1 GEMM is what bounds that GEMM).
__device__ __forceinline__ float gelu_erf(float x) {
  const float u = fminf(fabsf(x), 6.0f);
  float q = fmaf(-1.889626219e-06f, u, 6.268139987e-05f);
  q = fmaf(q, u, -9.388679173e-04f);
  q = fmaf(q, u, 8.539461531e-03f);
  q = fmaf(q, u, -5.402068794e-02f);
  q = fmaf(q, u, -4.584097862e-01f);
  q = fmaf(q, u, -1.151269197e+00f);
  q = fmaf(q, u, -9.999943376e-01f);
  float h;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(q));
  const float t = x * h;
  return x >= 0.f ? x - t : t;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// gelu(x) AND d gelu / dx = Phi(x) + x phi(x) from the same Phi(-|x|) = 2^q(|x|) as gelu_erf (g is bit-identical to gelu_erf(x)); the
// density is a second ex2: 2 MUFU + ~14 FMA-class ops for both values, where erff + __expf + gelu_erf cost ~70 instructions per
// element and made the gated-GELU backward pass issue-bound (110 us for 420 MB at the QVH encoder shape).
__device__ __forceinline__ void gelu_erf_both(float x, float& g, float& dg) {
  const float u = fminf(fabsf(x), 6.0f);
  float q = fmaf(-1.889626219e-06f, u, 6.268139987e-05f);
  q = fmaf(q, u, -9.388679173e-04f);
  q = fmaf(q, u, 8.539461531e-03f);
  q = fmaf(q, u, -5.402068794e-02f);
  q = fmaf(q, u, -4.584097862e-01f);
  q = fmaf(q, u, -1.151269197e+00f);
  q = fmaf(q, u, -9.999943376e-01f);
  float h, e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(q));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-0.72134752044448170f * x * x));      // exp(-x^2 / 2)
  const float t = x * h;
  g = x >= 0.f ? x - t : t;
  dg = fmaf(x, 0.3989422804014327f * e, x >= 0.f ? 1.0f - h : h);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }

// pack two fp32 into one 32-bit word of fp16 / bf16 (dt: MRB_DT_F16 / MRB_DT_BF16)
__device__ __forceinline__ uint32_t pack2(float a, float b, int dt) {
  if (dt == MRB_DT_F16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float unpack_lo(uint32_t w, int dt) {
  if (dt == MRB_DT_F16) return __half2float(__ushort_as_half(static_cast<unsigned short>(w & 0xffff)));
  return __uint_as_float(w << 16);
}
__device__ __forceinline__ float unpack_hi(uint32_t w, int dt) {
  if (dt == MRB_DT_F16) return __half2float(__ushort_as_half(static_cast<unsigned short>(w >> 16)));
  return __uint_as_float(w & 0xffff0000u);
}

}  // namespace mrb

// ---------------------------------------------------------------- programmatic dependent launch (build flag -DMRB_PDL)
// With MRB_PDL every kernel is launched with the programmatic-stream-serialization attribute: it may become resident while
// its predecessor in the stream is still draining, runs its set-up (barrier init, TMEM allocation, descriptor prefetch) and
// blocks in pdl_wait() until the predecessor has completed and flushed.  Rule that makes this safe by construction: every
// kernel executes pdl_wait() before its FIRST global-memory access (reads and writes alike), so completion is transitive
// along the stream.  pdl_trigger() (all CTAs, at entry) is what allows the successor to be scheduled early.  Without the
// flag both are empty and launches are ordinary <<< >>> launches: the default build is unchanged.
namespace mrb {
#ifdef MRB_PDL
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
#ifdef MRB_PDL_MAX_BLOCKS   // early launch only for small grids (the decoder's latency-bound chain); big kernels launch as usual
  at[0].val.programmaticStreamSerializationAllowed =
      (static_cast<unsigned long long>(grid.x) * grid.y * grid.z <= static_cast<unsigned long long>(MRB_PDL_MAX_BLOCKS)) ? 1 : 0;
#else
  at[0].val.programmaticStreamSerializationAllowed = 1;
#endif
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface through MRB_CHECK_LAUNCH
}
#else
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_trigger() {}
#endif
}  // namespace mrb
#ifdef MRB_PDL
#define MRB_LAUNCH(kernel, grid, block, smem, stream, ...) mrb::launch_pdl(kernel, grid, block, smem, stream, __VA_ARGS__)
#else
#define MRB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

// host-side launch check used by every C-ABI entry point
#define MRB_CHECK_LAUNCH()                          \
  do {                                              \
    cudaError_t e__ = cudaGetLastError();           \
    if (e__ != cudaSuccess) return mrb_set_error(e__); \
  } while (0)

extern "C" int mrb_set_error(cudaError_t e);
