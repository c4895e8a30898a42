// How fast can ONE SM pull GEMM operand tiles through TMA, and what counts against that rate?
//
// Background: the decoder-sized GEMMs (M = 56) take 2.1 ns x K blocks x (128 + BN) box rows whatever the tile
// width, and the large 2-CTA GEMMs stall at ~43 B/clk/SM of operand feed.  This probe replays just the producer side of
// csrc/gemm.cu -- per stage one A box (64 x rowsA 16-bit elements, SWIZZLE_128B) and one B box (64 x rowsB) into a ring,
// a consumer warp that only releases the slot -- and times it for
//   * 1 ... 148 CTAs (is the cap per SM or chip-wide?)
//   * A boxes of 128 rows over a 56-row tensor (are zero-filled out-of-bounds rows charged?) vs 64-row boxes vs 128 valid rows
//   * B boxes of 32 / 64 / 128 / 256 rows, ring depth 4 / 8
//   * operands streamed from HBM (distinct bytes per CTA, larger than L2) vs re-read from L2
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_feed_probe tools/probe/tma_feed_probe.cu -lcuda && /tmp/tma_feed_probe
#include "../../mr_blip_b200/csrc/common.cuh"
#include <cstdlib>
#include <vector>

using namespace mrb;

struct ProbeParams {
  int k_blocks;     // stages streamed per CTA
  int rows_a, rows_b;
  int stages;
  int b_slab_rows;  // B row offset per CTA (0: every CTA reads the same rows -> L2 hits after the first)
  int repeat;       // passes over the K range
};

__global__ void __launch_bounds__(64, 1)
feed_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = (p.rows_a + p.rows_b) * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int r = 0; r < p.repeat; ++r)
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * 64, 0);
          tma_load_2d(sa + p.rows_a * 128, &tmB, &full_bar[stage], kb * 64, blockIdx.x * p.b_slab_rows);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    }
  } else {
    int stage = 0;
    uint32_t phase = 0;
    for (int r = 0; r < p.repeat; ++r)
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    fprintf(stderr, "no cuTensorMapEncodeTiled\n");
    exit(1);
  }
  return reinterpret_cast<EncodeTiledFn>(ptr);
}

static CUtensorMap make_map(void* base, long long rows, long long cols, int box_rows) {
  static EncodeTiledFn fn = encode_fn();
  CUtensorMap m;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "encode failed %d\n", static_cast<int>(r)); exit(1); }
  return m;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const long long K = 16384;                        // 256 K blocks of 64
  const long long b_rows_total = 148LL * 256;       // a private 256-row slab per CTA: 1.24 GB, far beyond L2
  void *dA = nullptr, *dB = nullptr;
  cudaMalloc(&dA, 128 * K * 2);
  cudaMalloc(&dB, b_rows_total * K * 2);
  cudaMemset(dA, 0, 128 * K * 2);
  cudaMemset(dB, 0, b_rows_total * K * 2);
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("SMs %d, nominal clock %.2f GHz; B/clk figures use the nominal clock\n", sms, khz / 1e6);
  printf("%5s %6s %6s %6s %6s %5s %6s | %9s %12s %12s %10s\n", "ctas", "boxA", "validA", "boxB", "stages", "src", "", "us", "GB/s valid", "GB/s boxes", "B/clk/SM");
  struct Case { int ctas, box_a, valid_a, box_b, stages, l2; };
  std::vector<Case> cases;
  for (int ctas : {1, 8, 32, 74, 148})
    for (int box_b : {32, 64, 128, 256}) cases.push_back({ctas, 128, 56, box_b, box_b >= 256 ? 4 : 8, 0});
  for (int ctas : {1, 32, 148}) {
    cases.push_back({ctas, 128, 128, 64, 8, 0});    // all A rows valid
    cases.push_back({ctas, 64, 56, 64, 8, 0});      // 64-row A box
    cases.push_back({ctas, 64, 56, 64, 12, 0});     // ... with a deeper ring
    cases.push_back({ctas, 128, 56, 64, 4, 0});     // shallow ring
    cases.push_back({ctas, 128, 56, 64, 8, 1});     // B re-read from L2 (every CTA the same slab)
    cases.push_back({ctas, 128, 128, 128, 6, 1});   // the large-GEMM stage shape from L2
    cases.push_back({ctas, 128, 128, 256, 4, 1});
  }
  for (const Case& c : cases) {
    CUtensorMap tmA = make_map(dA, c.valid_a, K, c.box_a);
    CUtensorMap tmB = make_map(dB, b_rows_total, K, c.box_b);
    ProbeParams p;
    p.k_blocks = static_cast<int>(K / 64);
    p.rows_a = c.box_a; p.rows_b = c.box_b; p.stages = c.stages;
    p.b_slab_rows = c.l2 ? 0 : 256;
    p.repeat = c.l2 ? 4 : 1;
    const int smem = c.stages * (c.box_a + c.box_b) * 128 + 2 * c.stages * 8 + 1024 + 64;
    if (smem > 220 * 1024) continue;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
      cudaEventRecord(e0);
      feed_kernel<<<c.ctas, 64, smem>>>(tmA, tmB, p);
      cudaEventRecord(e1);
      cudaError_t err = cudaEventSynchronize(e1);
      if (err != cudaSuccess) { fprintf(stderr, "launch failed: %s\n", cudaGetErrorString(err)); return 1; }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (it > 0 && ms < best) best = ms;          // first pass warms the instruction cache / descriptors
    }
    const double stages = static_cast<double>(p.k_blocks) * p.repeat * c.ctas;
    const double valid = stages * (c.valid_a < c.box_a ? c.valid_a : c.box_a) * 128 + stages * c.box_b * 128;
    const double boxes = stages * (c.box_a + c.box_b) * 128;
    const double s = best * 1e-3;
    printf("%5d %6d %6d %6d %6d %5s %6s | %9.1f %12.1f %12.1f %10.1f\n", c.ctas, c.box_a, c.valid_a, c.box_b, c.stages,
           c.l2 ? "L2" : "HBM", "", best * 1e3, valid / s / 1e9, boxes / s / 1e9, boxes / c.ctas / s / (khz * 1e3));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  cudaFree(dA);
  cudaFree(dB);
  return 0;
}
