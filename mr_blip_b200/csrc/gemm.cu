// tcgen05 GEMM for every dense contraction on the BLIP2_MR hot path (SURVEY.md §2c K1,K3,K5,K6,K8,
// K10,K11,K14,K16,K17,K18 and the dgrad GEMMs of K19):
//
//     D[M,N] = epilogue( A[M,K] . B[N,K]^T )        A, B: fp16 or bf16, K-major (nn.Linear layout)
//
// Blackwell-native structure: persistent CTAs (one per SM), warp-specialised
//   warp 0     TMA producer   (cp.async.bulk.tensor 2D, SWIZZLE_128B, 64-element K slabs)
//   warp 1     MMA issuer     (single elected thread, tcgen05.mma cta_group::1 kind::f16, 128 x BN x 16)
//   warps 2-9  epilogue       (tcgen05.ld 32x32b from TMEM -> bias / GELU / fp32 residual -> global; two warps per
//                              TMEM lane quadrant, each taking half of the tile's columns)
// with a STAGES-deep smem ring (full/empty mbarriers) and a double-buffered fp32 accumulator in TMEM
// (tmem_full/tmem_empty mbarriers) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include <stdlib.h>
#include <mutex>

namespace mrb {

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles;
  int dtype;            // A/B: MRB_DT_F16 / MRB_DT_BF16
  // epilogue
  const float* bias;    // [N] or null
  int gelu;             // exact erf GELU after bias
  const float* resid;   // fp32 [*, ldr] added after activation, or null (may alias out when out is fp32)
  long long ldr;
  void* out;
  int out_dtype;        // MRB_DT_*
  long long ldc;
  int epi_direct;       // 16-bit outputs: row-per-thread stores instead of the staged (coalesced) epilogue
  int row_group;        // G > 0: patch-embed row remap  out_row = (m/G)*(G+1)+1+m%G, resid_row = 1+m%G
  // split-K instantiations only: tile index t = split * (m_tiles * n_tiles) + tile; split s contracts the K blocks
  // [s * kb_per_split, min(k_blocks, (s + 1) * kb_per_split)) into fp32 partials at out + s * split_stride
  int splits, kb_per_split;
  long long split_stride;
  // split-K with the reduce folded into the launch (tile_counters != null): the CTA that finishes an output tile LAST (per-tile
  // arrival counter, zero on entry and left zero) sums the partials of all splits in split order and applies this epilogue --
  // the arithmetic of splitk_reduce_kernel, without the second launch
  int* tile_counters;
  const float* fin_bias; int fin_gelu; const float* fin_resid; long long fin_ldr;
  void* fin_out; int fin_dtype; long long fin_ldc;
};

// per-workspace arrival counters of the folded split-K reduce: zero at module load, every launch leaves them zero
constexpr int SPLITK_SLOTS = 16, SPLITK_COUNTERS = 256;
__device__ int g_splitk_counters[SPLITK_SLOTS][SPLITK_COUNTERS];

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GROUP_M = 16;   // m-tiles per L2 super-group

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;                   // 8 warps x [32 rows][32 fp32], XOR-swizzled
  static constexpr int BAR_OFFSET = EPI_OFFSET + 8 * 4096;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;  // + alignment slack
};

__device__ __forceinline__ void tile_coords(int t, int m_tiles, int n_tiles, int& tm, int& tn) {
  const int per_group = GROUP_M * n_tiles;
  const int g = t / per_group;
  const int r = t - g * per_group;
  const int gm = min(GROUP_M, m_tiles - g * GROUP_M);
  tm = g * GROUP_M + r % gm;
  tn = r / gm;
}

// Tile t of a split-K launch -> (output tile, split, K-block range); the identity for ordinary launches.
template <bool SPLIT>
__device__ __forceinline__ void split_coords(const GemmParams& p, int k_blocks, int t, int& tile, int& split, int& kb0, int& kb1) {
  tile = t; split = 0; kb0 = 0; kb1 = k_blocks;
  if (SPLIT) {
    const int base = p.m_tiles * p.n_tiles;
    split = t / base;
    tile = t - split * base;
    kb0 = split * p.kb_per_split;
    kb1 = min(k_blocks, kb0 + p.kb_per_split);
  }
}

template <int BN, int STAGES, bool SPLIT = false>
__global__ void __launch_bounds__(320, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  mrb::pdl_trigger();   // the successor may become resident and run its set-up; it blocks in its own pdl_wait()
  using S = GemmSmem<BN, STAGES>;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  volatile uint32_t* last_flag = tmem_holder + 1;      // [2]: "this CTA arrived last at its output tile" (folded split-K reduce)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = SPLIT ? p.m_tiles * p.n_tiles * p.splits : p.m_tiles * p.n_tiles;
  const int k_blocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  mrb::pdl_wait();      // set-up done; nothing above touches global memory (MRB_PDL, common.cuh)

  if (warp == 0) {
    // ===================== TMA producer: converged warp, one elected lane issues (see gemm2.cu) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tm, tn, tile, split, kb0, kb1;
      split_coords<SPLIT>(p, k_blocks, t, tile, split, kb0, kb1);
      tile_coords(tile, p.m_tiles, p.n_tiles, tm, tn);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * S::STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, tm * BM);
          tma_load_2d(sa + S::A_BYTES, &tmB, &full_bar[stage], kb * BK, tn * BN);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: converged warp, uniform descriptors, one elected lane issues =====================
    const uint32_t idesc = umma_idesc_f16(p.dtype == MRB_DT_BF16 ? 1 : 0, BM, BN);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(smem));
    const uint64_t b_desc0 = umma_desc_sw128(smem_u32(smem) + S::A_BYTES);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      int tile, split, kb0, kb1;
      split_coords<SPLIT>(p, k_blocks, t, tile, split, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t a_desc = a_desc0 + static_cast<uint32_t>(stage * (S::STAGE_BYTES >> 4));
        const uint64_t b_desc = b_desc0 + static_cast<uint32_t>(stage * (S::STAGE_BYTES >> 4));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);                    // frees the smem slot when the MMAs retire
          if (kb == kb1 - 1) umma_commit(&tmem_full[as]);      // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (8): warps 2-5 take the low half of the tile's columns, 6-9 the high half.
    // Each 32 x 32 fp32 chunk goes TMEM -> registers (row per thread: bias, GELU) -> XOR-swizzled shared staging ->
    // registers (4 rows x 128 B per warp instruction) so that the residual reads and the output writes are coalesced.
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int eg = (warp - 2) >> 2;            // column half
    constexpr int HALF = (BN >= 64) ? BN / 2 : BN;          // BN = 32: only the first group has columns
    const int c_begin = eg * HALF, c_end = (BN >= 64) ? c_begin + HALF : (eg == 0 ? BN : 0);
    float4* stage4 = reinterpret_cast<float4*>(smem + S::EPI_OFFSET + (warp - 2) * 4096);
    const int sub_row = lane >> 3, chunk = lane & 7;        // transposed phase: lane -> (row within group of 4, 16-byte chunk)
    const bool direct16 = p.epi_direct && (p.out_dtype != MRB_DT_F32) && (p.resid == nullptr) && (p.row_group == 0);
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      int tm, tn, tile, split, kb0, kb1;
      split_coords<SPLIT>(p, k_blocks, t, tile, split, kb0, kb1);
      tile_coords(tile, p.m_tiles, p.n_tiles, tm, tn);
      void* const outp = SPLIT ? static_cast<void*>(static_cast<float*>(p.out) + split * p.split_stride) : p.out;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m_base = tm * BM + quad * 32;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 32) {
        const int n0 = tn * BN + c;
        if (n0 >= p.N) break;                  // warp-uniform
        const int ncols = min(32, p.N - n0);   // multiple of 8
        if (direct16) {
          // ---- 16-bit output without residual: row-per-thread stores (64 B per thread and chunk); cheapest in instructions,
          //      which is what bounds the bias / GELU epilogues of the short-K ViT GEMMs
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + c, r);
          tmem_ld_wait();
          const int m = m_base + lane;
          if (m < p.M) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (4 * j < ncols) {
                  const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j);
                  v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
                }
              }
            }
            if (p.gelu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
            }
            uint4* op = reinterpret_cast<uint4*>(static_cast<uint16_t*>(outp) + static_cast<long long>(m) * p.ldc + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j * 8 < ncols) {
                uint4 w;
                w.x = pack2(v[8 * j], v[8 * j + 1], p.out_dtype);
                w.y = pack2(v[8 * j + 2], v[8 * j + 3], p.out_dtype);
                w.z = pack2(v[8 * j + 4], v[8 * j + 5], p.out_dtype);
                w.w = pack2(v[8 * j + 6], v[8 * j + 7], p.out_dtype);
                op[j] = w;
              }
            }
          }
          continue;
        }
        // ---- staged path: residual / bias for the transposed phase first (coalesced loads in flight during the TMEM read)
        float4 rr[8];
        const int col = n0 + chunk * 4;
        const bool col_ok = chunk * 4 < ncols;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m_base + i * 4 + sub_row;
          rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.resid && m < p.M && col_ok) {
            const long long rrow = p.row_group > 0 ? 1 + (m % p.row_group) : m;
            rr[i] = *reinterpret_cast<const float4*>(p.resid + rrow * p.ldr + col);
          }
        }
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + c, r);
        tmem_ld_wait();
        __syncwarp();                          // previous chunk's readers are done with the staging buffer
#pragma unroll
        for (int j = 0; j < 8; ++j)
          stage4[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                           __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + sub_row;
          const int m = m_base + rl;
          float4 x = stage4[rl * 8 + (chunk ^ (rl & 7))];
          x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
          if (p.gelu) { x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w); }
          x.x += rr[i].x; x.y += rr[i].y; x.z += rr[i].z; x.w += rr[i].w;
          if (m < p.M && col_ok) {
            const long long orow = p.row_group > 0 ? static_cast<long long>(m / p.row_group) * (p.row_group + 1) + 1 + (m % p.row_group) : m;
            if (p.out_dtype == MRB_DT_F32) {
              *reinterpret_cast<float4*>(static_cast<float*>(outp) + orow * p.ldc + col) = x;
            } else {
              *reinterpret_cast<uint2*>(static_cast<uint16_t*>(outp) + orow * p.ldc + col) =
                  make_uint2(pack2(x.x, x.y, p.out_dtype), pack2(x.z, x.w, p.out_dtype));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if constexpr (SPLIT) {
        if (p.tile_counters) {
          // ---- folded reduce.  Release: every epilogue thread fences its partial stores, the eight warps meet, one thread
          //      bumps the tile's counter.  Acquire: the CTA that sees splits - 1 earlier arrivals fences and reads all partials
          //      (L2 loads: other CTAs wrote them).  Counter reset by the last arriver: the next launch finds zeros.
          __threadfence();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (warp == 2 && lane == 0) {
            const int prev = atomicAdd(p.tile_counters + tile, 1);
            const bool last = prev == p.splits - 1;
            if (last) { p.tile_counters[tile] = 0; __threadfence(); }
            last_flag[it & 1] = last ? 1u : 0u;
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (last_flag[it & 1]) {
            const float* ws = static_cast<const float*>(p.out);
#pragma unroll 1
            for (int c = c_begin; c < c_end; c += 32) {
              const int n0 = tn * BN + c;
              if (n0 >= p.N) break;
              const int col = n0 + chunk * 4;
              if (col >= p.N) continue;             // N is a multiple of 8: a float4 is all in or all out
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.fin_bias) b4 = __ldg(reinterpret_cast<const float4*>(p.fin_bias + col));
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int m = m_base + i * 4 + sub_row;
                if (m >= p.M) continue;
                const float* src = ws + static_cast<long long>(m) * p.N + col;
                float4 x = __ldcg(reinterpret_cast<const float4*>(src));
                for (int sp = 1; sp < p.splits; ++sp) {
                  const float4 y = __ldcg(reinterpret_cast<const float4*>(src + sp * p.split_stride));
                  x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
                }
                x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
                if (p.fin_gelu) { x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w); }
                if (p.fin_resid) {
                  const float4 r4 = *reinterpret_cast<const float4*>(p.fin_resid + m * p.fin_ldr + col);
                  x.x += r4.x; x.y += r4.y; x.z += r4.z; x.w += r4.w;
                }
                if (p.fin_dtype == MRB_DT_F32) {
                  *reinterpret_cast<float4*>(static_cast<float*>(p.fin_out) + m * p.fin_ldc + col) = x;
                } else {
                  *reinterpret_cast<uint2*>(static_cast<uint16_t*>(p.fin_out) + m * p.fin_ldc + col) =
                      make_uint2(pack2(x.x, x.y, p.fin_dtype), pack2(x.z, x.w, p.fin_dtype));
                }
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

// 2D row-major [rows, cols] 16-bit tensor, box = 64 cols x box_rows, 128B swizzle
static int make_tmap(CUtensorMap* map, const void* base, int dtype, long long rows, long long cols, long long ld,
                     int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return MRB_ERR_CUDA;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dtype == MRB_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MRB_OK : MRB_ERR_CUDA;
}

static int g_num_sms = 0;

template <int BN, int STAGES, bool SPLIT = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, cudaStream_t stream) {
  using S = GemmSmem<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, STAGES, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return mrb_set_error(e);
    configured = true;
  }
  p.n_tiles = (p.N + BN - 1) / BN;
  const int tiles = p.m_tiles * p.n_tiles * (SPLIT ? p.splits : 1);
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  MRB_LAUNCH((gemm_tcgen05_kernel<BN, STAGES, SPLIT>), grid, 320, S::TOTAL, stream, tmA, tmB, p);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

// ---------------------------------------------------------------- split-K: second pass
// out = epilogue( sum_s ws[s] ): the partial tiles are summed in split order (deterministic) and the caller's epilogue --
// bias, exact GELU, fp32 residual, output type -- is applied once, in the order of the fused epilogue above.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long split_stride, int M, int N, const float* __restrict__ bias,
                     int gelu, const float* resid, long long ldr, void* out, int out_dtype, long long ldc) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const int n4 = N >> 2;
  const long long total = static_cast<long long>(M) * n4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / n4);
    const int col = static_cast<int>(i - static_cast<long long>(m) * n4) * 4;
    const float* src = ws + static_cast<long long>(m) * N + col;
    float4 x = *reinterpret_cast<const float4*>(src);
    for (int sp = 1; sp < splits; ++sp) {
      const float4 y = *reinterpret_cast<const float4*>(src + sp * split_stride);
      x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    }
    if (bias) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col));
      x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
    }
    if (gelu) { x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w); }
    if (resid) {
      const float4 r4 = *reinterpret_cast<const float4*>(resid + m * ldr + col);
      x.x += r4.x; x.y += r4.y; x.z += r4.z; x.w += r4.w;
    }
    if (out_dtype == MRB_DT_F32) {
      *reinterpret_cast<float4*>(static_cast<float*>(out) + m * ldc + col) = x;
    } else {
      *reinterpret_cast<uint2*>(static_cast<uint16_t*>(out) + m * ldc + col) =
          make_uint2(pack2(x.x, x.y, out_dtype), pack2(x.z, x.w, out_dtype));
    }
  }
}

// Split-K plan.  Launches with few output tiles are bound by what ONE SM can pull through its TMA / L2 port, not by HBM or
// the tensor pipe: measured time ~ 2.1 ns x (K blocks per CTA) x (128 + BN) rows per stage over every tile width
// (profiles/gemm_small_m_r01b.log).  So: pick (BN, splits) that minimises  waves x ceil(k_blocks / splits) x (128 + BN)
// with tiles x splits CTAs, charging a split launch the fixed cost of its reduce pass.
struct SplitPlan { int bn, splits, kb_per; };

static SplitPlan plan_split(int M, int N, int K, int sms, int force_bn, int max_splits) {
  const int k_blocks = (K + BK - 1) / BK;
  const int m_tiles = (M + BM - 1) / BM;
  const int cand[4] = {32, 64, 128, 256};
  SplitPlan best = {0, 1, k_blocks};
  long long best_cost = -1;
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    if (force_bn ? bn != force_bn : ((bn == 32) != (N <= 32))) continue;   // 32-wide tiles only for the 32-column problems
    const long long tiles = static_cast<long long>(m_tiles) * ((N + bn - 1) / bn);
    for (int sp = 1; sp <= max_splits; ++sp) {
      const int kb_per = (k_blocks + sp - 1) / sp;
      if (sp > 1 && (kb_per < 4 || tiles * sp > sms)) break;
      if ((k_blocks + kb_per - 1) / kb_per != sp) continue;                // this count leaves a split empty
      const long long waves = (tiles * sp + sms - 1) / sms;
      const long long cost = waves * kb_per * (128 + bn) + (sp > 1 ? 1500 : 0);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = {bn, sp, kb_per}; }
    }
  }
  return best;
}

}  // namespace mrb

using namespace mrb;

// Small-M note (decoder steps, M = 56): these launches take ~15 us for K = 2080 and ~55 us for K = 10272
// (profiles/gemm_small_m_r01b.log).  Over every tile width measured the time fits  2.1 ns x K blocks x (128 + BN) rows per
// stage: 26.4 / 17.7 us for N = 10240 at BN 64 (two waves) / 128, 13-15 us for N = 2048, 55 us for K = 10272 -- i.e. the
// bytes ONE SM pulls through its TMA / L2 port (~45 B/clk, the same per-SM feed rate that bounds the large GEMMs), with the
// zero-filled rows of the 128-row A box counted, not HBM (floor 1.4 us) and not the tensor pipe.  With 32-96 CTAs most SMs
// idle; mrb_gemm_splitk below spreads the K blocks of such problems over the idle SMs.
// Pick the N tile: minimise (waves over the SMs) x (tile width) x (relative MMA inefficiency of narrow tiles).
// Narrow tiles win only when the grid would otherwise leave most SMs idle (decoder steps, M <= 128).
static int pick_bn(int M, int N, int sms) {
  if (N <= 32) return 32;     // skinny LoRA down-projections: [M,K] x [32,K]^T
  const int cand[4] = {256, 192, 128, 64};
  const double eff[4] = {1.0, 1.05, 1.35, 1.8};
  const long long m_tiles = (M + 127) / 128;
  int best = 256;
  double best_cost = -1.0;
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    const long long tiles = m_tiles * ((N + bn - 1) / bn);
    const double waves = tiles <= sms ? 1.0 : static_cast<double>(tiles) / sms;
    const double cost = waves * bn * eff[i];
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// Arrival counters of the folded split-K reduce for the workspace `ws` (one workspace per issuing stream, so launches that
// share a slot are stream-ordered): a slot of g_splitk_counters per distinct workspace address.  nullptr -> two-pass path
// (the default; more output tiles than counters; more than SPLITK_SLOTS workspaces).
// OFF by default, MRB_SPLITK_FUSED=1 enables: measured on a B200 (call 26) the decoder chain went from 18.8 to 27.3 ms in-graph
// with it -- the last CTA of a tile reads splits x 64 x BN fp32 through ONE SM's L2 port as dependent round trips, where the
// separate reduce launch spreads the same bytes over >= 128 blocks in 4.6 us.  Kept (bit-identical to the two-pass result,
// tests/test_splitk_gpu.py) as the measured negative result.
static int* splitk_counters_for(const void* ws, int tiles) {
  static int fused = -1;
  if (fused < 0) { const char* e = getenv("MRB_SPLITK_FUSED"); fused = (e && e[0] == '1') ? 1 : 0; }
  if (!fused || tiles > SPLITK_COUNTERS) return nullptr;
  static std::mutex mu;
  static const void* owner[SPLITK_SLOTS] = {};
  static int* base = nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!base) {
    void* sym = nullptr;
    if (cudaGetSymbolAddress(&sym, g_splitk_counters) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    base = static_cast<int*>(sym);
  }
  for (int i = 0; i < SPLITK_SLOTS; ++i) {
    if (owner[i] == ws) return base + i * SPLITK_COUNTERS;
    if (owner[i] == nullptr) { owner[i] = ws; return base + i * SPLITK_COUNTERS; }
  }
  return nullptr;
}

// SM cap of the persistent 2-CTA GEMM kernel for the launches the CALLING THREAD issues next (mrb_gemm_sm_limit).  A CTA of that
// kernel owns the whole shared memory of its SM for the length of the launch, so a GEMM on a SIDE stream that takes all SMs stalls
// the dependent chain of small kernels on the main stream for that long (the T5 decoder: 48 encoder-sized cross-attention K/V
// GEMMs next to ~1 500 decoder-sized kernels); capped, it leaves SMs to the chain and the two really overlap.  Per thread and per
// call site, not per stream: stream handles are pooled and reused, a table keyed by them would outlive its owner.
static thread_local int t_sm_limit = 0;
static int sms_for_launch() { return (t_sm_limit > 0 && t_sm_limit < g_num_sms) ? t_sm_limit : g_num_sms; }

extern "C" int mrb_gemm_sm_limit(int sms) {
  if (sms < 0 || (sms & 1)) return MRB_ERR_ARG;
  t_sm_limit = sms;
  return MRB_OK;
}

extern "C" int mrb_gemm2_launch(const CUtensorMap* tmA, const CUtensorMap* tmB, int M, int N, int K, int dtype,
                                const float* bias, int gelu, const float* resid, long long ldr, void* out, int out_dtype,
                                long long ldc, int row_group, int num_sms, void* stream);

static int gemm_impl(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, int dtype,
                     const float* bias, int gelu, const float* resid, long long ldr, void* out, int out_dtype,
                     long long ldc, int row_group, int force_bn, void* ws, long long ws_bytes, int max_splits, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return MRB_OK;
  if ((dtype != MRB_DT_F16 && dtype != MRB_DT_BF16) || (N & 7) || (lda & 7) || (ldb & 7) || (K & 7)) return MRB_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(out)) & 15) return MRB_ERR_ARG;
  if (out_dtype == MRB_DT_F32 ? (ldc & 3) : (ldc & 7)) return MRB_ERR_ARG;
  if (resid && ((ldr & 3) || (reinterpret_cast<uintptr_t>(resid) & 15))) return MRB_ERR_ARG;
  if (resid && gelu) return MRB_ERR_UNSUPPORTED;      // no call site combines them (residual is added after a linear)
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  static int use_2cta = -1;               // MRB_GEMM_2CTA=0 disables the CTA-pair kernel (A/B measurements)
  if (use_2cta < 0) { const char* e = getenv("MRB_GEMM_2CTA"); use_2cta = (e && e[0] == '0') ? 0 : 1; }
  if (use_2cta && force_bn == 0 && M >= 512 && N >= 256 && pick_bn(M, N, g_num_sms) == 256) {
    CUtensorMap tmA2, tmB2;
    int rc2 = make_tmap(&tmA2, A, dtype, M, K, lda, 128);
    if (rc2) return rc2;
    rc2 = make_tmap(&tmB2, B, dtype, N, K, ldb, 128);
    if (rc2) return rc2;
    return mrb_gemm2_launch(&tmA2, &tmB2, M, N, K, dtype, bias, gelu, resid, ldr, out, out_dtype, ldc, row_group,
                            sms_for_launch(), stream);
  }
  int bn = force_bn ? force_bn : pick_bn(M, N, g_num_sms);
  SplitPlan plan = {bn, 1, 0};
  if (ws && max_splits > 1 && row_group == 0 && (M <= BM || N <= 32)) {
    if (reinterpret_cast<uintptr_t>(ws) & 15) return MRB_ERR_ARG;
    max_splits = max_splits < 8 ? max_splits : 8;
    const SplitPlan sp = plan_split(M, N, K, g_num_sms, force_bn, max_splits);
    if (sp.splits > 1 && static_cast<long long>(sp.splits) * M * N * 4 <= ws_bytes) { plan = sp; bn = sp.bn; }
  }
  CUtensorMap tmA, tmB;
  int rc = make_tmap(&tmA, A, dtype, M, K, lda, BM);
  if (rc) return rc;
  rc = make_tmap(&tmB, B, dtype, N, K, ldb, bn);
  if (rc) return rc;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = 0;
  p.dtype = dtype;
  p.bias = bias; p.gelu = gelu; p.resid = resid; p.ldr = ldr;
  p.out = out; p.out_dtype = out_dtype; p.ldc = ldc; p.row_group = row_group;
  p.splits = 1; p.kb_per_split = 0; p.split_stride = 0;
  p.tile_counters = nullptr; p.fin_bias = nullptr; p.fin_gelu = 0; p.fin_resid = nullptr; p.fin_ldr = 0;
  p.fin_out = nullptr; p.fin_dtype = 0; p.fin_ldc = 0;
  {
    static int mode = -1;                 // MRB_GEMM_EPI=direct selects the row-per-thread 16-bit epilogue (A/B measurements)
    if (mode < 0) { const char* e = getenv("MRB_GEMM_EPI"); mode = e ? (e[0] == 'd' ? 1 : 2) : 0; }
    p.epi_direct = mode == 1 ? 1 : 0;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (plan.splits > 1) {
    // partial sums: fp32 [splits][M][N] in the caller's workspace, no epilogue; the reduce pass applies it
    p.bias = nullptr; p.gelu = 0; p.resid = nullptr; p.ldr = 0; p.epi_direct = 0;
    p.out = ws; p.out_dtype = MRB_DT_F32; p.ldc = N;
    p.splits = plan.splits; p.kb_per_split = plan.kb_per; p.split_stride = static_cast<long long>(M) * N;
    int* counters = splitk_counters_for(ws, p.m_tiles * ((N + bn - 1) / bn));
    if (counters) {                       // reduce folded into the launch: the last CTA of every output tile applies the epilogue
      p.tile_counters = counters;
      p.fin_bias = bias; p.fin_gelu = gelu; p.fin_resid = resid; p.fin_ldr = ldr;
      p.fin_out = out; p.fin_dtype = out_dtype; p.fin_ldc = ldc;
    }
    switch (bn) {
      case 256: rc = launch_gemm<256, 4, true>(tmA, tmB, p, s); break;
      case 128: rc = launch_gemm<128, 6, true>(tmA, tmB, p, s); break;
      case 64: rc = launch_gemm<64, 8, true>(tmA, tmB, p, s); break;
      case 32: rc = launch_gemm<32, 8, true>(tmA, tmB, p, s); break;
      default: return MRB_ERR_ARG;
    }
    if (rc) return rc;
    if (counters) return MRB_OK;
    const long long quads = static_cast<long long>(M) * (N >> 2);
    long long blocks = (quads + 255) / 256;
    if (blocks > 4LL * g_num_sms) blocks = 4LL * g_num_sms;
    MRB_LAUNCH((splitk_reduce_kernel), static_cast<int>(blocks), 256, 0, s, static_cast<const float*>(ws), plan.splits, p.split_stride, M, N,
                                                                  bias, gelu, resid, ldr, out, out_dtype, ldc);
    MRB_CHECK_LAUNCH();
    return MRB_OK;
  }
  switch (bn) {
    case 256: return launch_gemm<256, 4>(tmA, tmB, p, s);
    case 192: return launch_gemm<192, 4>(tmA, tmB, p, s);
    case 128: return launch_gemm<128, 6>(tmA, tmB, p, s);
    case 64: return launch_gemm<64, 8>(tmA, tmB, p, s);
    case 32: return launch_gemm<32, 8>(tmA, tmB, p, s);
    default: return MRB_ERR_ARG;
  }
}

extern "C" int mrb_gemm(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, int dtype,
                        const float* bias, int gelu, const float* resid, long long ldr, void* out, int out_dtype,
                        long long ldc, int row_group, int force_bn, void* stream) {
  return gemm_impl(A, lda, B, ldb, M, N, K, dtype, bias, gelu, resid, ldr, out, out_dtype, ldc, row_group, force_bn, nullptr, 0, 1,
                   stream);
}

// The split-K decision of mrb_gemm_splitk for a problem on a device with `sms` SMs (host arithmetic only; lets callers size
// workspaces and lets the CPU tests check the plan): *splits == 1 means the unsplit path.
extern "C" int mrb_gemm_splitk_plan(int M, int N, int K, int sms, int force_bn, int max_splits, int* bn, int* splits,
                                    int* kb_per_split) {
  if (M <= 0 || N <= 0 || K <= 0 || sms <= 0 || !bn || !splits || !kb_per_split) return MRB_ERR_ARG;
  SplitPlan sp = {force_bn ? force_bn : pick_bn(M, N, sms), 1, (K + BK - 1) / BK};
  if (max_splits > 1 && (M <= BM || N <= 32)) {
    const SplitPlan cand = plan_split(M, N, K, sms, force_bn, max_splits < 8 ? max_splits : 8);
    if (cand.splits > 1) sp = cand;
  }
  *bn = sp.bn; *splits = sp.splits; *kb_per_split = sp.kb_per;
  return MRB_OK;
}

// mrb_gemm with a caller-owned fp32 workspace: problems with one row tile (decoder steps, M <= 128) or 32 columns (LoRA
// down-projections) whose grid would leave most SMs idle are split along K over up to max_splits CTAs per output tile
// (partials in `ws`, >= splits * M * N * 4 bytes, 16-byte aligned; summed in split order by a second launch that applies the
// epilogue).  Everything else, and any problem the workspace is too small for, runs exactly as mrb_gemm.  The workspace is in
// use until the call's work has finished on `stream`: one workspace per stream that issues such calls.
extern "C" int mrb_gemm_splitk(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, int dtype,
                               const float* bias, int gelu, const float* resid, long long ldr, void* out, int out_dtype,
                               long long ldc, int row_group, int force_bn, void* ws, long long ws_bytes, int max_splits,
                               void* stream) {
  return gemm_impl(A, lda, B, ldb, M, N, K, dtype, bias, gelu, resid, ldr, out, out_dtype, ldc, row_group, force_bn, ws, ws_bytes,
                   max_splits, stream);
}
