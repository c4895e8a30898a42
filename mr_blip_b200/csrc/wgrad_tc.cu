// Tensor-core "skinny" weight-gradient reduction for the LoRA adapters (peft lora.Linear backward):
//     out[c, r] += sum_m P[m, c] * Q[m, r]         P: [M, C] 16-bit (activations or output grads), Q: [M, 8] 16-bit
// i.e. dB = dY^T (x A^T) and dA = (dY sB)^T x.  The contraction runs over the ROWS of both operands, so both are read
// MN-major straight from their row-major tiles (A: 128 columns of P as the UMMA M dimension, two SWIZZLE_128B atoms;
// B: 16 columns of Q as N, SWIZZLE_32B) -- no transposes in memory.  Grid = (C / 128) x split-K over M; each CTA
// accumulates a 128 x 16 fp32 tile in TMEM and adds its 128 x 8 slice to `out` with fp32 atomics.
#include "common.cuh"

namespace mrb {

struct WgradParams {
  int M, C, rows_per_cta;
  float* out;
  float* out2;              // optional: columns 8..15 of Q (the next LoRA slot of the same group) accumulate here
  int transposed_out, dtype;
};

constexpr int WG_BK = 64;            // rows (contraction) per pipeline stage
constexpr int WG_STAGES = 6;
constexpr int WG_A_BYTES = 2 * WG_BK * 128;    // two 64-column atoms x 64 rows x 128 B
constexpr int WG_B_BYTES = WG_BK * 32;
constexpr int WG_STAGE = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE + (2 * WG_STAGES + 1) * 8 + 16 + 1024;

__device__ __forceinline__ uint64_t wg_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ, const WgradParams p) {
  mrb::pdl_trigger();   // the successor may become resident and run its set-up; it blocks in its own pdl_wait()
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* acc_bar = empty_bar + WG_STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 128;
  const int m0 = blockIdx.y * p.rows_per_cta;
  const int m1 = min(p.M, m0 + p.rows_per_cta);
  const int k_blocks = (m1 - m0 + WG_BK - 1) / WG_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmP);
    tma_prefetch_desc(&tmQ);
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  mrb::pdl_wait();      // set-up done; nothing above touches global memory (MRB_PDL, common.cuh)

  if (warp == 0) {       // TMA producer: converged warp, one elected lane issues (elect_one, common.cuh)
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < k_blocks; ++kb) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sa = smem + stage * WG_STAGE;
        mbar_expect_tx(&full_bar[stage], WG_STAGE);
        const int m = m0 + kb * WG_BK;
        tma_load_2d(sa, &tmP, &full_bar[stage], c0, m);
        tma_load_2d(sa + WG_BK * 128, &tmP, &full_bar[stage], c0 + 64, m);
        tma_load_2d(sa + WG_A_BYTES, &tmQ, &full_bar[stage], 0, m);
      }
      __syncwarp();
      if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {  // MMA issuer: same pattern, descriptors = bases + offsets of the 14-bit address field (16-byte units)
    const uint32_t fmt = p.dtype == MRB_DT_BF16 ? 1u : 0u;
    // fp32 accumulate, A and B both MN-major, M = 128, N = 16
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    // A: 16 rows = two 8-row groups (SBO 1024 B); the second 64-column atom sits LBO = 8192 B further
    const uint64_t ad0 = wg_desc(smem_u32(smem), WG_BK * 128, 1024, 2);
    // B: 16 rows x 32 B, SWIZZLE_32B: 8-row groups of 256 B
    const uint64_t bd0 = wg_desc(smem_u32(smem) + WG_A_BYTES, WG_BK * 32, 256, 6);
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < k_blocks; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t so = static_cast<uint32_t>(stage * (WG_STAGE >> 4));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < WG_BK / 16; ++k)
          umma_f16(tmem_base, ad0 + (so + k * 128), bd0 + (so + k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (kb == k_blocks - 1) umma_commit(acc_bar);
      }
      __syncwarp();
      if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    const int quad = warp & 3;
    const int c = c0 + quad * 32 + lane;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    uint32_t r[16];
    tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16), r);
    tmem_ld_wait();
    if (c < p.C) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = __uint_as_float(r[j]);
        if (p.transposed_out) atomicAdd(p.out + static_cast<long long>(j) * p.C + c, v);
        else atomicAdd(p.out + static_cast<long long>(c) * 8 + j, v);
      }
      if (p.out2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = __uint_as_float(r[8 + j]);
          if (p.transposed_out) atomicAdd(p.out2 + static_cast<long long>(j) * p.C + c, v);
          else atomicAdd(p.out2 + static_cast<long long>(c) * 8 + j, v);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn wg_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
static int wg_tmap(CUtensorMap* map, const void* base, int dtype, long long rows, long long cols, long long ld, int box_cols,
                   CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = wg_encode_fn();
  if (!fn) return MRB_ERR_CUDA;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(WG_BK)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dtype == MRB_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MRB_OK : MRB_ERR_CUDA;
}

}  // namespace mrb

using namespace mrb;

// Q must have at least 16 readable 16-bit columns per row starting at its pointer (the extended [.., +32] buffers do).
extern "C" int mrb_skinny_wgrad_tc2(const void* P, long long ldp, const void* Q, long long ldq, int M, int C, float* out,
                                    float* out2, int transposed_out, int dtype, void* stream);
extern "C" int mrb_skinny_wgrad_tc(const void* P, long long ldp, const void* Q, long long ldq, int M, int C, float* out,
                                   int transposed_out, int dtype, void* stream) {
  return mrb_skinny_wgrad_tc2(P, ldp, Q, ldq, M, C, out, nullptr, transposed_out, dtype, stream);
}
// Two adjacent LoRA slots in one pass over P: out += P^T Q[:, 0:8], out2 += P^T Q[:, 8:16] (out2 may be NULL).
extern "C" int mrb_skinny_wgrad_tc2(const void* P, long long ldp, const void* Q, long long ldq, int M, int C, float* out,
                                    float* out2, int transposed_out, int dtype, void* stream) {
  if (M <= 0 || C <= 0) return MRB_OK;
  if ((ldp & 7) || (ldq & 7) || (C & 7) || ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(Q)) & 15)) return MRB_ERR_ARG;
  static int sms = 0;
  static bool cfg = false;
  if (!cfg) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    if (e != cudaSuccess) return mrb_set_error(e);
    cfg = true;
  }
  CUtensorMap tmP, tmQ;
  int rc = wg_tmap(&tmP, P, dtype, M, C, ldp, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (!rc) rc = wg_tmap(&tmQ, Q, dtype, M, 16, ldq, 16, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc) return rc;
  const int c_tiles = (C + 127) / 128;
  int splits = (2 * sms + c_tiles - 1) / c_tiles;
  const int k_blocks = (M + WG_BK - 1) / WG_BK;
  if (splits > k_blocks) splits = k_blocks;
  if (splits < 1) splits = 1;
  WgradParams p;
  p.M = M; p.C = C;
  p.rows_per_cta = ((k_blocks + splits - 1) / splits) * WG_BK;
  p.out = out; p.out2 = out2; p.transposed_out = transposed_out; p.dtype = dtype;
  dim3 grid(c_tiles, (M + p.rows_per_cta - 1) / p.rows_per_cta);
  MRB_LAUNCH((wgrad_tc_kernel), grid, 192, WG_SMEM, static_cast<cudaStream_t>(stream), tmP, tmQ, p);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
