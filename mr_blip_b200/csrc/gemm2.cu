// 2-CTA (cta_group::2) variant of the tcgen05 GEMM for the large contractions (M >= 512):  a CTA pair on one TPC
// computes a 256 x 256 output tile.  Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256
// weight rows); one thread of the leader CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 256) which reads both
// CTAs' shared memory, so every SM moves 8 KB instead of 12 KB per 16-wide K step and the pipeline holds 6 stages.
// Each CTA's TMEM receives its own 128 accumulator rows (2 x 256 columns, double buffered); both CTAs run the same
// 8-warp staged epilogue as gemm.cu.  Synchronisation: TMA completions of both CTAs land on the leader's full
// barrier; tcgen05.commit multicasts to the empty / tmem_full barriers of both CTAs; the peer's epilogue warps
// arrive remotely on the leader's tmem_empty barrier.
#include "common.cuh"

namespace mrb {

struct Gemm2Params {
  int M, N, K;
  int m_tiles, n_tiles;      // 256-row, 256-column tiles
  int dtype;
  const float* bias; int gelu;
  const float* resid; long long ldr;
  void* out; int out_dtype; long long ldc;
  int row_group;
  int sem_cluster;          // 1: cluster-scope acquire / release on the mbarrier waits / remote arrives (old behaviour; costs a
                            //    CCTL.IVALL per wait and a MEMBAR + ERRBAR per arrive: profiles/ncu_gemm2_fc1_r01b); 0: CTA scope
  int tail;                 // 1: the last column tile runs a narrower MMA (MRB_GEMM2_TAIL=0 disables, for A/B runs)
};

// Diagnostic builds (python -m mr_blip_b200.build --variant _x -DMRB_G2_...; never the default library): MRB_G2_STAGES=n ring depth,
// MRB_G2_NOEPI the epilogue warps release the accumulator without reading it (main loop alone), MRB_G2_NOTMA the producer
// signals "full" without loading anything (MMA issue rate alone, operands are whatever shared memory holds).
#ifndef MRB_G2_STAGES
#define MRB_G2_STAGES 6
#endif
constexpr int G2_BM = 128, G2_BN = 256, G2_BK = 64, G2_STAGES = MRB_G2_STAGES, G2_GROUP_M = 8;
constexpr int G2_A_BYTES = G2_BM * G2_BK * 2;            // 16 KB: this CTA's 128 rows
constexpr int G2_B_BYTES = (G2_BN / 2) * G2_BK * 2;      // 16 KB: this CTA's half of the B tile
constexpr int G2_STAGE = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_EPI_OFFSET = G2_STAGES * G2_STAGE;
constexpr int G2_BAR_OFFSET = G2_EPI_OFFSET + 8 * 4096;
constexpr int G2_SMEM = G2_BAR_OFFSET + (2 * G2_STAGES + 4) * 8 + 16 + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {   // arrive on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr, int sem_cluster) {
  if (sem_cluster) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
  else asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// The pair's barriers carry no generic-memory data: operands arrive through the async proxy (TMA complete_tx), accumulators
// through tcgen05.commit / tcgen05.fence.  CTA-scope acquire (the PTX default, what CUTLASS's ClusterBarrier uses) is enough;
// cluster scope makes every successful wait invalidate L1 (CCTL.IVALL) in the single-thread TMA / MMA issue loops.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int sem_cluster) {
  if (!sem_cluster) { mbar_wait(bar, parity); return; }
  if (mbar_try_wait_cluster(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > 8000000000LL) {
      printf("mrb: cluster mbarrier wait timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void tile_coords2(int t, int m_tiles, int n_tiles, int& tm, int& tn) {
  const int per_group = G2_GROUP_M * n_tiles;
  const int g = t / per_group;
  const int r = t - g * per_group;
  const int gm = min(G2_GROUP_M, m_tiles - g * G2_GROUP_M);
  tm = g * G2_GROUP_M + r % gm;
  tn = r / gm;
}

__device__ __forceinline__ int tile_col2(int t, int m_tiles, int n_tiles) {
  int tm, tn;
  tile_coords2(t, m_tiles, n_tiles, tm, tn);
  return tn;
}

// Width of the MMA for column tile tn: 256, or the remaining columns rounded up to 64 for the last tile (N = 1408 = 5 x 256
// + 128 in the ViT proj / fc2 GEMMs: the tail tile then costs half the tensor time).  Each CTA supplies n_mma / 2 rows of B.
__device__ __forceinline__ int tile_n_mma(const Gemm2Params& p, int tn) {
  const int rem = p.N - tn * G2_BN;
  return (rem >= G2_BN || !p.tail) ? G2_BN : ((rem + 63) & ~63);
}

// Epilogue kinds with a specialised fast path (everything else, and every edge tile, takes the generic path)
enum { EPI_GENERIC = 0, EPI_F16 = 1, EPI_GELU_F16 = 2, EPI_BF16 = 3, EPI_RESID_F32 = 4, EPI_F32 = 5 };

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Gemm2Params p) {
  mrb::pdl_trigger();   // the successor may become resident and run its set-up; it blocks in its own pdl_wait()
  constexpr uint32_t TMEM_COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G2_BAR_OFFSET);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tmem_full = empty_bar + G2_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int k_blocks = (p.K + G2_BK - 1) / G2_BK;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);        // leader's producer arrives (expect_tx covers both CTAs' bytes)
      mbar_init(&empty_bar[s], 1);       // multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);       // multicast tcgen05.commit
      mbar_init(&tmem_empty[s], 16);     // 8 epilogue warps x 2 CTAs (only the leader's copy is used)
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_holder, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  mrb::pdl_wait();      // set-up done; nothing above touches global memory (MRB_PDL, common.cuh)

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): converged warp, one elected lane issues =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += n_pairs) {
        int tm, tn;
        tile_coords2(t, p.m_tiles, p.n_tiles, tm, tn);
        const int a_row = tm * 256 + static_cast<int>(rank) * G2_BM;
        const int b_row = tn * G2_BN + static_cast<int>(rank) * (tile_n_mma(p, tn) / 2);   // this CTA's half of the (tail) tile
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * G2_STAGE;
          if (elect_one()) {
#ifdef MRB_G2_NOTMA
            (void)sa; (void)a_row; (void)b_row;
            if (leader) mbar_arrive(&full_bar[stage]);
#else
            const uint32_t lbar = map_to_cta(smem_u32(&full_bar[stage]), 0);
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * G2_STAGE);
            tma2_load_2d(sa, &tmA, lbar, kb * G2_BK, a_row);
            tma2_load_2d(sa + G2_A_BYTES, &tmB, lbar, kb * G2_BK, b_row);
#endif
          }
          __syncwarp();
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // The whole warp runs the loop converged (every quantity below is warp-uniform, so it lives in uniform registers) and ONE
    // elected lane issues: with `if (lane == 0)` the compiler had to move every tcgen05.mma operand through an
    // ELECT / R2UR.BROADCAST loop -- ~200 instructions, ~700 clk per 64-wide K block against 512 clk of tensor work
    // (profiles/ncu_gemm2_issue_r02d.md): the issue loop, not the operand feed, was what held the tensor pipe at 58-70 %.
    if (leader) {
      const int fmt = p.dtype == MRB_DT_BF16 ? 1 : 0;
      const uint32_t idesc_full = umma_idesc_f16(fmt, 256, G2_BN);
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(smem));               // stage 0, k step 0; stage s adds s * G2_STAGE / 16,
      const uint64_t b_desc0 = umma_desc_sw128(smem_u32(smem) + G2_A_BYTES);  // k step k adds 2 k to the (14-bit) address field
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair; t < num_tiles; t += n_pairs, ++it) {
        const int tn = tile_col2(t, p.m_tiles, p.n_tiles);
        const int n_mma = tile_n_mma(p, tn);
        const uint32_t idesc = n_mma == G2_BN ? idesc_full : umma_idesc_f16(fmt, 256, n_mma);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * G2_BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + static_cast<uint32_t>(stage * (G2_STAGE >> 4));
          const uint64_t b_desc = b_desc0 + static_cast<uint32_t>(stage * (G2_STAGE >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < G2_BK / 16; ++k)
              umma2_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma2_commit_mc(&empty_bar[stage]);
            if (kb == k_blocks - 1) umma2_commit_mc(&tmem_full[as]);
          }
          __syncwarp();
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps (8 per CTA): same staged epilogue as gemm.cu =====================
    const int quad = warp & 3;
    const int eg = (warp - 2) >> 2;
    float4* stage4 = reinterpret_cast<float4*>(smem + G2_EPI_OFFSET + (warp - 2) * 4096);
    const int sub_row = lane >> 3, chunk = lane & 7;
    const uint32_t empty_remote0 = map_to_cta(smem_u32(&tmem_empty[0]), 0);
    const uint32_t empty_remote1 = map_to_cta(smem_u32(&tmem_empty[1]), 0);
    int it = 0;
    for (int t = pair; t < num_tiles; t += n_pairs, ++it) {
      int tm, tn;
      tile_coords2(t, p.m_tiles, p.n_tiles, tm, tn);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int half = tile_n_mma(p, tn) / 2;             // multiple of 32: each warp group takes half of the tile's columns
      const int c_begin = eg * half, c_end = c_begin + half;
      const int m_base = tm * 256 + static_cast<int>(rank) * G2_BM + quad * 32;
      mbar_wait_cluster(&tmem_full[as], aphase, p.sem_cluster);
      tc_fence_after();
#ifdef MRB_G2_NOEPI
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(as ? empty_remote1 : empty_remote0, p.sem_cluster);
      (void)half; (void)c_begin; (void)c_end; (void)m_base; (void)stage4; (void)sub_row; (void)chunk;
      continue;
#endif
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * G2_BN;
      // One 32 x 32 fp32 chunk: TMEM registers (row per thread) -> XOR-swizzled staging -> 4 rows x 128 B per warp instruction
      // (coalesced residual reads / output writes), bias + GELU + residual in between.
      auto epi_chunk = [&](const uint32_t (&r)[32], const int c, const float4 (&rr)[8], const float4 b4) {
        const int n0 = tn * G2_BN + c;
        const int ncols = min(32, p.N - n0);
        const int col = n0 + chunk * 4;
        const bool col_ok = chunk * 4 < ncols;
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
              *reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.ldc + col) = x;
            } else {
              *reinterpret_cast<uint2*>(static_cast<uint16_t*>(p.out) + orow * p.ldc + col) =
                  make_uint2(pack2(x.x, x.y, p.out_dtype), pack2(x.z, x.w, p.out_dtype));
            }
          }
        }
      };
      // residual / bias of a chunk (issued before the TMEM wait so that the loads are in flight during it)
      auto epi_prefetch = [&](const int c, float4 (&rr)[8], float4& b4) {
        const int n0 = tn * G2_BN + c;
        const int col = n0 + chunk * 4;
        const bool col_ok = n0 + chunk * 4 < p.N;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m_base + i * 4 + sub_row;
          rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.resid && m < p.M && col_ok) {
            const long long rrow = p.row_group > 0 ? 1 + (m % p.row_group) : m;
            rr[i] = *reinterpret_cast<const float4*>(p.resid + rrow * p.ldr + col);
          }
        }
        b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
      };
      // ---- fast path (tile rows and columns all inside the matrix, no row remap): the epilogue kind is a template
      //      parameter, so the per-element code is pure math + one store (no flag tests, no 64-bit index arithmetic, no
      //      bounds predicates); bias of all four chunks is fetched before the accumulator wait.
      if (EPI != EPI_GENERIC && p.row_group == 0 && m_base + 32 <= p.M && tn * G2_BN + c_end <= p.N) {
        constexpr bool GELU = (EPI == EPI_GELU_F16);
        constexpr bool RESID = (EPI == EPI_RESID_F32);
        constexpr bool OUT32 = (EPI == EPI_RESID_F32 || EPI == EPI_F32);
        constexpr int ODT = (EPI == EPI_BF16) ? MRB_DT_BF16 : MRB_DT_F16;
        const int col0 = tn * G2_BN + c_begin + chunk * 4;
        const int nch = half >> 5;
        float4 bz[4];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
          bz[ci] = (p.bias && ci < nch) ? __ldg(reinterpret_cast<const float4*>(p.bias + col0 + ci * 32)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const long long row0 = m_base + sub_row;
        const float* rptr = RESID ? p.resid + row0 * p.ldr + col0 : nullptr;
        char* optr = static_cast<char*>(p.out) + (row0 * p.ldc + col0) * (OUT32 ? 4 : 2);
        const long long rstep = 4 * p.ldr, ostep = 4 * p.ldc * (OUT32 ? 4 : 2);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          if (ci < nch) {
            float4 rr[8];
            if (RESID) {
#pragma unroll
              for (int i = 0; i < 8; ++i) rr[i] = *reinterpret_cast<const float4*>(rptr + i * rstep + ci * 32);
            }
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_row + c_begin + ci * 32, r);
            tmem_ld_wait();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              stage4[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                               __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            __syncwarp();
            const float4 b4 = bz[ci];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rl = i * 4 + sub_row;
              float4 x = stage4[rl * 8 + (chunk ^ (rl & 7))];
              x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
              if (GELU) { x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w); }
              if (RESID) { x.x += rr[i].x; x.y += rr[i].y; x.z += rr[i].z; x.w += rr[i].w; }
              char* o = optr + i * ostep + ci * (OUT32 ? 128 : 64);
              if (OUT32) *reinterpret_cast<float4*>(o) = x;
              else *reinterpret_cast<uint2*>(o) = make_uint2(pack2(x.x, x.y, ODT), pack2(x.z, x.w, ODT));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(as ? empty_remote1 : empty_remote0, p.sem_cluster);
        continue;
      }
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 32) {
        if (tn * G2_BN + c >= p.N) break;
        float4 rr[8], b4;
        epi_prefetch(c, rr, b4);
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + c, r);
        tmem_ld_wait();
        epi_chunk(r, c, rr, b4);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(as ? empty_remote1 : empty_remote0, p.sem_cluster);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
  }
}

template <int EPI>
static int launch2(const CUtensorMap* tmA, const CUtensorMap* tmB, const Gemm2Params& p, int pairs, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_tcgen05_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM);
    if (e != cudaSuccess) return mrb_set_error(e);
    configured = true;
  }
  MRB_LAUNCH((gemm2_tcgen05_kernel<EPI>), 2 * pairs, 320, G2_SMEM, st, *tmA, *tmB, p);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

}  // namespace mrb

using namespace mrb;

// Called by mrb_gemm (gemm.cu) with tensor maps already built: A box 64 x 128, B box 64 x 128.
extern "C" int mrb_gemm2_launch(const CUtensorMap* tmA, const CUtensorMap* tmB, int M, int N, int K, int dtype,
                                const float* bias, int gelu, const float* resid, long long ldr, void* out, int out_dtype,
                                long long ldc, int row_group, int num_sms, void* stream) {
  static int semc = -1;                   // MRB_GEMM2_SEM=cluster restores cluster-scope barrier semantics (A/B measurements)
  if (semc < 0) { const char* e = getenv("MRB_GEMM2_SEM"); semc = (e && e[0] == 'c') ? 1 : 0; }
  static int spec = -1;                   // MRB_GEMM2_EPI=generic disables the specialised epilogues (A/B measurements)
  if (spec < 0) { const char* e = getenv("MRB_GEMM2_EPI"); spec = (e && e[0] == 'g') ? 0 : 1; }
  Gemm2Params p;
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + 255) / 256;
  p.n_tiles = (N + G2_BN - 1) / G2_BN;
  p.dtype = dtype; p.bias = bias; p.gelu = gelu; p.resid = resid; p.ldr = ldr;
  p.out = out; p.out_dtype = out_dtype; p.ldc = ldc; p.row_group = row_group;
  static int tail = -1;
  if (tail < 0) { const char* e = getenv("MRB_GEMM2_TAIL"); tail = (e && e[0] == '0') ? 0 : 1; }
  p.tail = tail;
  p.sem_cluster = semc;
  const int tiles = p.m_tiles * p.n_tiles;
  int pairs = num_sms / 2;
  if (tiles < pairs) pairs = tiles;
  int epi = EPI_GENERIC;
  if (spec && row_group == 0) {
    if (resid && !gelu && out_dtype == MRB_DT_F32) epi = EPI_RESID_F32;
    else if (!resid && !gelu && out_dtype == MRB_DT_F32) epi = EPI_F32;
    else if (!resid && gelu && out_dtype == MRB_DT_F16) epi = EPI_GELU_F16;
    else if (!resid && !gelu && out_dtype == MRB_DT_F16) epi = EPI_F16;
    else if (!resid && !gelu && out_dtype == MRB_DT_BF16) epi = EPI_BF16;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (epi) {
    case EPI_F16: return launch2<EPI_F16>(tmA, tmB, p, pairs, st);
    case EPI_GELU_F16: return launch2<EPI_GELU_F16>(tmA, tmB, p, pairs, st);
    case EPI_BF16: return launch2<EPI_BF16>(tmA, tmB, p, pairs, st);
    case EPI_RESID_F32: return launch2<EPI_RESID_F32>(tmA, tmB, p, pairs, st);
    case EPI_F32: return launch2<EPI_F32>(tmA, tmB, p, pairs, st);
    default: return launch2<EPI_GENERIC>(tmA, tmB, p, pairs, st);
  }
}
