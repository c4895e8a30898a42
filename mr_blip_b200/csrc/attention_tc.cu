// tcgen05 flash-attention forward for the two big attention shapes of the path:
//   * T5 encoder self-attention (modeling_t5.py:561-610): hd 64, L ~ 2037, additive bucketed bias + padding mask
//   * EVA ViT attention (eva_vit.py:128-145): hd 88 (handled as 64 + 32 with TMA zero fill of d >= 88), 257 tokens
// One CTA per (G x 128 query rows, head, batch), G softmax groups (template parameter, see TcSmem):
//   warp 0        TMA producer: Q once, then the K/V tiles (4-D tensor maps: d, head, token, batch)
//   warp 1        single-thread tcgen05.mma issuer:  S_j = Q K_j^T  (128 x 128 x hd per group, fp32 in TMEM)
//                                                    O += P_j V_j   (accumulated in TMEM, V read MN-major from its [key][d] tile)
//   warps 2..     softmax, 4 warps per group, one thread per query row (TMEM lane): tcgen05.ld S, one FFMA per element for
//                 the exp2 argument (scale, bias, reference maximum folded), exp2, row sum, P_j -> bf16/fp16 -> SWIZZLE_128B
//                 smem (A operand of the PV MMA).  Single pass against a STALE row maximum; the tile is redone exactly (O and
//                 l rescaled in TMEM) only when its row sum shows that a score exceeded the reference by more than 2^8.
// The last KV tile is issued with N = round_up(remaining keys, 16), so 257 keys cost 2 x 128 + 16, not 3 x 128.
// Measured limits (profiles/): per (128 x 128) tile the softmax side costs 64 KB of tcgen05.ld (TMEM read port 64 B/clk/SM)
// and 16 K exp2 (16/clk/SM) against 512 clk of tensor work at hd 64 -- these, not the tensor pipe, bound the kernel.
#include "common.cuh"
#include "dropmask.cuh"

namespace mrb {

struct AttnTcParams {
  int B, H, Lq, Lk, hd;
  int dtype;
  float scale;
  const float* bias; int bias_len, bias_zero;   // [H, bias_len], index (j - i) + bias_zero
  const int* kmask;                              // [B, Lk] or null
  int kv_div, causal, q_pos0;
  void* o; long long o_bs, o_rs;
  float* lse;                                    // [B, H, Lq] or null
  // DROP instantiations only: attention-probability dropout (modeling_t5.py:600), masks of dropmask.cuh with
  // row = (b H + h) Lq + i, column = key j; thr7 = (round(256 p) / 2) * 0x01010101 (the SWAR compare below needs an even threshold)
  const uint32_t* drop_seed; uint32_t drop_site, drop_thr7; float drop_scale;
};

// Keep-masks of four consecutive keys as two AND-masks over the packed 16-bit pairs (keys 0,1 | keys 2,3):  draw >= thr with an
// even thr  <=>  (draw >> 1) >= thr / 2, evaluated for the four bytes at once -- (draw >> 1) | 0x80 minus thr / 2 keeps bit 7 of
// its byte exactly when the draw passes (no borrow crosses a byte) -- and PRMT replicates those sign bits over the half words.
__device__ __forceinline__ void drop_pair_masks(uint32_t w, uint32_t thr7, uint32_t& m01, uint32_t& m23) {
  const uint32_t t = (((w >> 1) & 0x7f7f7f7fu) | 0x80808080u) - thr7;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m01) : "r"(t), "r"(0u), "r"(0x9988u));
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m23) : "r"(t), "r"(0u), "r"(0xBBAAu));
}

constexpr int TQ = 128, TKV = 128;     // per softmax group: 128 query rows; a CTA runs two groups (256 rows)
constexpr float TAU = 8.0f;            // stale-max slack (log2 domain): P entries stay below 2^8

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// generic UMMA smem descriptor: lbo / sbo in bytes
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}
constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW64 = 4;

__host__ __device__ constexpr uint32_t idesc_f16(int fmt, int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
         (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// G = softmax groups (128 query rows each) per CTA.  G = 2: one CTA per SM, the tensor pipe works on one group while the
// other runs its softmax, K/V tiles in a 2-stage ring.  G = 1 (short sequences, i.e. the ViT's 257 tokens): half the shared
// memory and TMEM (one K and one V buffer with their own barriers), so TWO CTAs share an SM and one CTA's load / softmax
// latency is covered by the other CTA's tensor work.
#ifndef MRB_FWD_STAGES
#define MRB_FWD_STAGES 2
#endif
template <int HD, int G>   // HD 64, or 96 (= 64-wide SW128 atom + 32-wide SW64 atom)
struct TcSmem {
  static constexpr bool SPLIT = (HD == 96);
  static constexpr int Q_ONE = TQ * 64 * 2 + (SPLIT ? TQ * 32 * 2 : 0);       // one group's Q tile
  static constexpr int KV_ONE = TKV * 64 * 2 + (SPLIT ? TKV * 32 * 2 : 0);   // one of K or V
  static constexpr int STAGE_BYTES = 2 * KV_ONE;
  static constexpr int STAGES = (G == 1) ? 1 : (HD == 64 ? MRB_FWD_STAGES : 2);  // K+V stages held in shared memory (hd 96: 3 would need 268 KB)
  static constexpr int NBARST = (G == 1) ? 2 : STAGES;                         // full / empty barrier pairs (G = 1: K and V)
  static constexpr int P_BYTES = TQ * TKV * 2;                                 // two 64-key atoms, per group
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = G * Q_ONE;
  static constexpr int OFF_P = OFF_KV + STAGES * STAGE_BYTES;
  static constexpr int OFF_BIAS = OFF_P + G * P_BYTES;                         // 4 G warps x 160 floats
  static constexpr int OFF_BAR = OFF_BIAS + 4 * G * 160 * 4;
  static constexpr int NBAR = 1 + 2 * NBARST + 6;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16 + 1024;
  static constexpr int THREADS = (2 + 4 * G) * 32;
};

// ---- softmax building blocks (one thread = one query row; all tcgen05.ld calls are warp-uniform) ----
struct RowCtx {
  uint32_t s_addr;          // TMEM address of this row's S tile (lane + column base)
  float sl2;                // scale * log2(e)
  const float* wrow;        // this row's bias window in shared memory (already * log2 e), column c -> wrow[c]
  uint32_t mb[4];           // attendable-key bits of the tile
  int ncols, kv0, i_abs, causal;
  uint32_t dkey, drow, dthr7;   // DROP: site key, row base = ((b H + h) Lq + i) * ceil(Lk / 4), replicated threshold
};

// raw 32-column slice of this row's S tile -> log2-domain scores minus `sub` (scale, bias window, masks); sub = 0 gives the
// scores themselves, sub = the row's reference maximum gives the exp2 arguments with ONE FFMA per element
template <bool HAS_BIAS, bool MASKED>
__device__ __forceinline__ void scores_from(const RowCtx& x, int c, const uint32_t* v, float sub, float* sv) {
  const uint32_t mw = (c == 0) ? x.mb[0] : (c == 32) ? x.mb[1] : (c == 64) ? x.mb[2] : x.mb[3];
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    float s = HAS_BIAS ? fmaf(__uint_as_float(v[e]), x.sl2, x.wrow[c + e] - sub) : fmaf(__uint_as_float(v[e]), x.sl2, -sub);
    if (MASKED) {
      const bool ok = (c + e < x.ncols) && ((mw >> e) & 1u) && !(x.causal && x.kv0 + c + e > x.i_abs);
      s = ok ? s : -INFINITY;
    }
    sv[e] = s;
  }
}
template <bool HAS_BIAS, bool MASKED>
__device__ __forceinline__ void load_scores(const RowCtx& x, int c, float sub, float* sv) {
  uint32_t v[32];
  tmem_ld_32x32b_x32(x.s_addr + c, v);
  tmem_ld_wait();
  scores_from<HAS_BIAS, MASKED>(x, c, v, sub, sv);
}

template <bool HAS_BIAS, bool MASKED>
__device__ __forceinline__ float tile_max(const RowCtx& x, int nc32) {
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
  for (int c = 0; c < nc32; c += 32) {
    float sv[32];
    load_scores<HAS_BIAS, MASKED>(x, c, 0.f, sv);
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      mx0 = fmaxf(mx0, sv[e]); mx1 = fmaxf(mx1, sv[e + 1]); mx2 = fmaxf(mx2, sv[e + 2]); mx3 = fmaxf(mx3, sv[e + 3]);
    }
  }
  return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
}

// P = exp2(s - mref) -> 16 bit -> SWIZZLE_128B rows of the P tile; returns the row sum.  Per element: FFMA, MUFU, FADD and
// half a pack (DT is a compile-time dtype where the launcher knows it, so the pack is not predicated both ways).
// (Software-pipelining the tcgen05.ld of the next 32 columns under the current exponentials was measured and gained
// nothing -- 0.381 vs 0.378 ms on the T5 encoder shape -- so the simple load / wait / compute loop stays.)
template <bool HAS_BIAS, bool MASKED, int DT, bool DROP = false>
__device__ __forceinline__ float exp_store(const RowCtx& x, int nc32, int npad, float mref, uint8_t* prow, int r, int dtype) {
  const int dt = DT < 0 ? dtype : DT;
  float rs0 = 0.f, rs1 = 0.f;
  for (int c = 0; c < nc32; c += 32) {
    float sv[32];
    load_scores<HAS_BIAS, MASKED>(x, c, mref, sv);
    uint32_t pk[16];
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
      const float p0 = ex2(sv[e]), p1 = ex2(sv[e + 1]);
      rs0 += p0; rs1 += p1;
      pk[e >> 1] = pack2(p0, p1, dt);
    }
    if (DROP) {                                    // the row sum above is of the undropped P; 1 / (1 - p) is folded into the final 1 / l
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4) {
        uint32_t m01, m23;
        drop_pair_masks(drop_word(x.dkey, x.drow, static_cast<uint32_t>((x.kv0 + c) >> 2) + g4), x.dthr7, m01, m23);
        pk[2 * g4] &= m01;
        pk[2 * g4 + 1] &= m23;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {                  // 4 x 16-byte chunks of 8 keys
      const int key0 = c + q * 8;
      if (!MASKED || key0 < npad) {
        const int atom = key0 >> 6, chunk = (key0 & 63) >> 3;
        *reinterpret_cast<uint4*>(prow + atom * (TQ * 128) + ((chunk ^ (r & 7)) << 4)) =
            make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
    }
  }
  return rs0 + rs1;
}

// One KV tile of the online softmax for this row.  The first tile is exact (two passes).  Later tiles exponentiate against
// the STALE reference maximum m_ref in a single pass; no per-element maximum is tracked: a tile holds <= 128 keys, so a row
// sum above 2^TAU = 256 is the (conservative, warp-voted) sign that some score exceeded the reference by more than TAU --
// only then the tile maximum is computed, O and l are rescaled and the tile is redone exactly.
template <bool HAS_BIAS, bool MASKED, int DT, bool DROP = false>
__device__ __forceinline__ float softmax_tile(const RowCtx& x, int j, int nc32, int npad, uint8_t* prow, int r, int dtype,
                                              float& m_ref, float& l_run, uint32_t o_addr, int hd_cols) {
  // NOTE: tcgen05.ld is warp-collective (.sync.aligned): every branch around it must be warp-uniform.
  float rsum;
  if (j == 0) {                                      // no reference yet: exact two-pass tile
    const float mx = tile_max<HAS_BIAS, MASKED>(x, nc32);
    if (mx != -INFINITY) m_ref = mx;
    rsum = exp_store<HAS_BIAS, MASKED, DT, DROP>(x, nc32, npad, mx == -INFINITY ? 0.f : mx, prow, r, dtype);
  } else {
    rsum = exp_store<HAS_BIAS, MASKED, DT, DROP>(x, nc32, npad, m_ref == -INFINITY ? 0.f : m_ref, prow, r, dtype);
    if (__any_sync(0xffffffffu, !(rsum <= 256.0f))) {    // rare: some row's max jumped (or overflowed); redo the tile exactly
      const float tmax = tile_max<HAS_BIAS, MASKED>(x, nc32);
      const float m_new = fmaxf(m_ref, tmax);
      const float corr = (m_ref == -INFINITY) ? 0.f : ex2(m_ref - m_new);
      l_run *= corr;
      // s_full(j) completed => PV(j-1) completed (in-order tensor pipe); PV(j) is not issued before p_full(j):
      // this thread owns its TMEM lane of O right now.
      for (int c = 0; c < hd_cols; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(o_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * corr);
        tmem_st_32x32b_x32(o_addr + c, v);
      }
      tmem_st_wait();
      m_ref = m_new;
      rsum = exp_store<HAS_BIAS, MASKED, DT, DROP>(x, nc32, npad, m_ref == -INFINITY ? 0.f : m_ref, prow, r, dtype);
    }
  }
  return rsum;
}

template <int HD, int G, int DT, bool DROP = false>   // DT: 0 fp16 / 1 bf16 known at compile time, -1 = p.dtype
__global__ void __launch_bounds__((2 + 4 * G) * 32, G == 1 ? 2 : 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmQ2,
                   const __grid_constant__ CUtensorMap tmK2, const __grid_constant__ CUtensorMap tmV2,
                   const AttnTcParams p) {
  mrb::pdl_trigger();   // the successor may become resident and run its set-up; it blocks in its own pdl_wait()
  using S = TcSmem<HD, G>;
  constexpr bool SPLIT = S::SPLIT;
  constexpr int STAGES = S::STAGES;
  constexpr int NBARST = S::NBARST;
  constexpr uint32_t TMEM_COLS = (G == 2) ? 512 : 256;
  constexpr int S_COL = 0, O_COL = G * TKV;      // G = 2: S_A [0,128) S_B [128,256); O_A [256,256+HD) O_B [256+HD, 256+2HD)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + NBARST;
  uint64_t* s_full = kv_empty + NBARST;    // [2] per group
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + S::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * (G * TQ);
  const int bkv = b / p.kv_div;
  const int n_groups = (G == 2 && q0 + TQ < p.Lq) ? 2 : 1; // the second 128-row group may be empty
  int n_kv = (p.Lk + TKV - 1) / TKV;
  if (p.causal) n_kv = min(n_kv, (min(q0 + n_groups * TQ, p.Lq) - 1 + p.q_pos0) / TKV + 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < NBARST; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 4); mbar_init(&o_full[g], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  mrb::pdl_wait();      // set-up done; nothing above touches global memory (MRB_PDL, common.cuh)

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues: elect_one, common.cuh) =====================
    if (elect_one()) {
      mbar_expect_tx(q_full, n_groups * S::Q_ONE);
      for (int g = 0; g < n_groups; ++g) {
        tma_load_4d(smem + S::OFF_Q + g * S::Q_ONE, &tmQ, q_full, 0, h, q0 + g * TQ, b);
        if (SPLIT) tma_load_4d(smem + S::OFF_Q + g * S::Q_ONE + TQ * 128, &tmQ2, q_full, 64, h, q0 + g * TQ, b);
      }
    }
    __syncwarp();
    if (G == 1) {
      // one K and one V buffer, each with its own full / empty barrier pair ([0] = K, [1] = V): K_{j+1} streams in as
      // soon as S_j = Q K_j^T has retired (i.e. during softmax j), V_{j+1} as soon as O += P_j V_j has retired
      uint8_t* sk = smem + S::OFF_KV;
      uint8_t* sv = sk + S::KV_ONE;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&kv_empty[0], (j & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&kv_full[0], S::KV_ONE);
          tma_load_4d(sk, &tmK, &kv_full[0], 0, h, j * TKV, bkv);
          if (SPLIT) tma_load_4d(sk + TKV * 128, &tmK2, &kv_full[0], 64, h, j * TKV, bkv);
        }
        __syncwarp();
        mbar_wait(&kv_empty[1], (j & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&kv_full[1], S::KV_ONE);
          tma_load_4d(sv, &tmV, &kv_full[1], 0, h, j * TKV, bkv);
          if (SPLIT) tma_load_4d(sv + TKV * 128, &tmV2, &kv_full[1], 64, h, j * TKV, bkv);
        }
        __syncwarp();
      }
    }
    for (int j = 0; G == 2 && j < n_kv; ++j) {
      const int st = j % STAGES;
      mbar_wait(&kv_empty[st], ((j / STAGES) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* sk = smem + S::OFF_KV + st * S::STAGE_BYTES;
        uint8_t* sv = sk + S::KV_ONE;
        mbar_expect_tx(&kv_full[st], S::STAGE_BYTES);
        tma_load_4d(sk, &tmK, &kv_full[st], 0, h, j * TKV, bkv);
        tma_load_4d(sv, &tmV, &kv_full[st], 0, h, j * TKV, bkv);
        if (SPLIT) {
          tma_load_4d(sk + TKV * 128, &tmK2, &kv_full[st], 64, h, j * TKV, bkv);
          tma_load_4d(sv + TKV * 128, &tmV2, &kv_full[st], 64, h, j * TKV, bkv);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Converged warp, one elected lane issues; descriptors = kernel-lifetime bases + small offsets in the 14-bit address field
    // (16-byte units): K-major SWIZZLE_128B tiles step 2 per 16-element K step, the MN-major V tile steps 128 (SW128 atom,
    // 16 keys = 2048 B) or 64 (SW64 atom, 1024 B).  `if (lane == 0)` + descriptors rebuilt from addresses cost ~20 dependent
    // uniform-datapath instructions and an ELECT / BRA.U.ANY loop per tcgen05.mma (profiles/ncu_attn_issue_r02d.md).
    const int fmt = p.dtype == MRB_DT_BF16 ? 1 : 0;
    const uint32_t sb = smem_u32(smem);
    const uint64_t dk128 = umma_desc(sb, 16, 1024, LAYOUT_SW128);             // K-major, 64-wide atom
    const uint64_t dk64 = umma_desc(sb, 16, 512, LAYOUT_SW64);                // K-major, 32-wide atom (d 64..95)
    const uint64_t dv128 = umma_desc(sb, TKV * 128, 1024, LAYOUT_SW128);      // V [key][d] read MN-major
    const uint64_t dv64 = umma_desc(sb, TKV * 64, 512, LAYOUT_SW64);
    const uint32_t id64 = idesc_f16(fmt, TQ, 64, 0, 1);
    const uint32_t id32 = idesc_f16(fmt, TQ, 32, 0, 1);
    auto tail_n = [&](int j) {   // keys in KV tile j, rounded up to the UMMA N granularity (16)
      const int rem = p.Lk - j * TKV;
      return rem >= TKV ? TKV : ((rem + 15) & ~15);
    };
    auto issue_qk = [&](int g, int j) {
      const int st = j % STAGES;
      const uint32_t qo = static_cast<uint32_t>((S::OFF_Q + g * S::Q_ONE) >> 4), ko = static_cast<uint32_t>((S::OFF_KV + st * S::STAGE_BYTES) >> 4);
      const uint64_t qd = dk128 + qo, kd = dk128 + ko;
      const uint32_t d_s = tmem_base + S_COL + g * TKV;
      const uint32_t ids = idesc_f16(fmt, TQ, tail_n(j), 0, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(d_s, qd + 2 * k, kd + 2 * k, ids, k > 0 ? 1u : 0u);
      if (SPLIT) {
        const uint64_t qd2 = dk64 + (qo + ((TQ * 128) >> 4)), kd2 = dk64 + (ko + ((TKV * 128) >> 4));
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(d_s, qd2 + 2 * k, kd2 + 2 * k, ids, 1u);
      }
      umma_commit(&s_full[g]);
    };
    auto issue_pv = [&](int g, int j) {
      const int st = j % STAGES;
      const uint32_t po = static_cast<uint32_t>((S::OFF_P + g * S::P_BYTES) >> 4);
      const uint32_t vo = static_cast<uint32_t>((S::OFF_KV + st * S::STAGE_BYTES + S::KV_ONE) >> 4);
      const uint64_t pd = dk128 + po, vd = dv128 + vo, vd2 = dv64 + (vo + ((TKV * 128) >> 4));
      const uint32_t d_o = tmem_base + O_COL + g * HD;
      const int ksteps = tail_n(j) / 16;
      const uint32_t acc0 = j > 0 ? 1u : 0u;
      auto step = [&](int k) {
        // A = P (K-major, SW128): 64-key atom (k / 4) is TQ * 128 B further, 32-byte step inside the atom
        const uint64_t a = pd + static_cast<uint32_t>((k >> 2) * ((TQ * 128) >> 4) + (k & 3) * 2);
        umma_f16(d_o, a, vd + static_cast<uint32_t>(k * 128), id64, k > 0 ? 1u : acc0);
        if (SPLIT) umma_f16(d_o + 64, a, vd2 + static_cast<uint32_t>(k * 64), id32, k > 0 ? 1u : acc0);
      };
      if (ksteps == TKV / 16) {
#pragma unroll
        for (int k = 0; k < TKV / 16; ++k) step(k);
      } else {
        for (int k = 0; k < ksteps; ++k) step(k);
      }
      umma_commit(&o_full[g]);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    if (G == 1) {
      if (elect_one()) { issue_qk(0, 0); umma_commit(&kv_empty[0]); }     // K_0 is free once S_0 has retired
      __syncwarp();
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&p_full[0], j & 1);                                   // P_j written, S_j consumed
        mbar_wait(&kv_full[1], j & 1);                                  // V_j landed
        tc_fence_after();
        if (elect_one()) { issue_pv(0, j); umma_commit(&kv_empty[1]); }
        __syncwarp();
        if (j + 1 < n_kv) {
          mbar_wait(&kv_full[0], (j + 1) & 1);                          // K_{j+1} landed
          tc_fence_after();
          if (elect_one()) { issue_qk(0, j + 1); umma_commit(&kv_empty[0]); }
          __syncwarp();
        }
      }
    } else {
    if (elect_one())
      for (int g = 0; g < n_groups; ++g) issue_qk(g, 0);
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      for (int g = 0; g < n_groups; ++g) {
        mbar_wait(&p_full[g], j & 1);
        if (g == 0 && j + 1 < n_kv) mbar_wait(&kv_full[(j + 1) % STAGES], ((j + 1) / STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_pv(g, j);
          if (g == n_groups - 1) umma_commit(&kv_empty[j % STAGES]);   // K_j / V_j fully consumed
          if (j + 1 < n_kv) issue_qk(g, j + 1);
        }
        __syncwarp();
      }
    }
    }
  } else {
    // ===================== softmax groups: one thread per query row =====================
    const int g = (warp - 2) >> 2;                         // group 0: warps 2-5, group 1: warps 6-9
    if (g < n_groups) {
      const int quad = warp & 3;
      const int r = quad * 32 + lane;
      const int qg0 = q0 + g * TQ;
      const int i_abs = min(qg0 + r, p.Lq - 1) + p.q_pos0;     // rows past Lq are computed but never stored
      const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      const uint32_t s_addr = lane_base + S_COL + g * TKV;
      const uint32_t o_addr = lane_base + O_COL + g * HD;
      const float LOG2E = 1.4426950408889634f;
      const float sl2 = p.scale * LOG2E;
      const float* bhead = p.bias ? p.bias + static_cast<long long>(h) * p.bias_len : nullptr;
      const int* mrow = p.kmask ? p.kmask + static_cast<long long>(bkv) * p.Lk : nullptr;
      float* wbias = reinterpret_cast<float*>(smem + S::OFF_BIAS) + (warp - 2) * 160;   // this warp's bias window
      uint8_t* prow = smem + S::OFF_P + g * S::P_BYTES + r * 128;
      float m_ref = -INFINITY, l_run = 0.f;
      uint32_t dkey = 0, drow = 0;
      if (DROP) {
        dkey = drop_key(*p.drop_seed, p.drop_site);
        drow = (static_cast<uint32_t>(b * p.H + h) * static_cast<uint32_t>(p.Lq) + static_cast<uint32_t>(min(qg0 + r, p.Lq - 1))) *
               drop_groups(static_cast<uint32_t>(p.Lk));
      }
      // window index for (row r, column c): (kv0 + c) - i_abs + zero = w0 + (31 - lane) + c, w0 = bias index of
      // (column 0, last row of this warp)
      const int i_warp_last = min(qg0 + quad * 32 + 31, p.Lq - 1) + p.q_pos0;

      float bnext[5] = {0.f, 0.f, 0.f, 0.f, 0.f};     // raw bias window (5 x 32 lanes = 160 floats) of the next tile
      int mnext[4] = {1, 1, 1, 1};                    // key-mask values (4 x 32 lanes = 128 keys) of the next tile, 0 past Lk
      auto prefetch_tile = [&](int jn) {
        const int kvn = jn * TKV;
        if (bhead) {
          const int w0 = kvn - i_warp_last + p.bias_zero + lane;
#pragma unroll
          for (int t = 0; t < 5; ++t) {
            const int idx = w0 + 32 * t;
            bnext[t] = (idx >= 0 && idx < p.bias_len) ? __ldg(bhead + idx) : 0.f;
          }
        }
        if (mrow) {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int jj = kvn + lane + 32 * w;
            mnext[w] = (jj < p.Lk) ? __ldg(mrow + jj) : 0;
          }
        }
      };
      prefetch_tile(0);
      for (int j = 0; j < n_kv; ++j) {
        const int kv0 = j * TKV;
        const int ncols = min(TKV, p.Lk - kv0);
        const int nc32 = (ncols + 31) & ~31;
        const int npad = (ncols + 15) & ~15;               // the PV MMA reads keys [0, npad)
        // stage this warp's bias window (pre-multiplied by log2 e) while the QK MMA runs
        // this tile's bias window / key-mask words were fetched one tile ahead (the L2 round trip of these per-tile global
        // loads used to sit in front of every softmax); stage them, then start the next tile's fetch
        bool bias_const = false;
        float cbias = 0.f;
        if (bhead) {
          __syncwarp();
          bool same = true;
          const float first = bnext[0] * LOG2E;
#pragma unroll
          for (int t = 0; t < 5; ++t) {
            const float bv = bnext[t] * LOG2E;
            wbias[lane + 32 * t] = bv;
            same = same && (bv == first);
          }
          const float lane0 = __shfl_sync(0xffffffffu, first, 0);
          bias_const = __all_sync(0xffffffffu, same && first == lane0);
          cbias = bias_const ? lane0 : 0.f;
          __syncwarp();
        }
        RowCtx x;
        x.s_addr = s_addr; x.sl2 = sl2; x.ncols = ncols; x.kv0 = kv0; x.i_abs = i_abs; x.causal = p.causal;
        x.mb[0] = x.mb[1] = x.mb[2] = x.mb[3] = 0xffffffffu;
        if (DROP) { x.dkey = dkey; x.drow = drow; x.dthr7 = p.drop_thr7; }
        bool masked = (ncols < TKV) || (p.causal && kv0 + TKV - 1 > qg0 + quad * 32 + p.q_pos0);
        if (mrow) {
#pragma unroll
          for (int w = 0; w < 4; ++w) x.mb[w] = __ballot_sync(0xffffffffu, mnext[w] != 0);
          masked = masked || ((x.mb[0] & x.mb[1] & x.mb[2] & x.mb[3]) != 0xffffffffu);
        }
        if (j + 1 < n_kv) prefetch_tile(j + 1);
        // rows of a warp are consecutive: this row's window starts (i_warp_last - i_abs) floats into the warp window
        x.wrow = wbias + (i_warp_last - i_abs);

        mbar_wait(&s_full[g], j & 1);
        tc_fence_after();

        float rsum;
        if (bhead && !bias_const) {
          rsum = masked ? softmax_tile<true, true, DT, DROP>(x, j, nc32, npad, prow, r, p.dtype, m_ref, l_run, o_addr, HD)
                        : softmax_tile<true, false, DT, DROP>(x, j, nc32, npad, prow, r, p.dtype, m_ref, l_run, o_addr, HD);
        } else {
          // no bias, or one bias value for the whole tile (T5 buckets saturate 128 positions off the diagonal, i.e. for all but
          // ~3 of a row block's KV tiles): a constant shift of the scores = a shift of the reference maximum, zero per-element cost
          m_ref -= cbias;
          rsum = masked ? softmax_tile<false, true, DT, DROP>(x, j, nc32, npad, prow, r, p.dtype, m_ref, l_run, o_addr, HD)
                        : softmax_tile<false, false, DT, DROP>(x, j, nc32, npad, prow, r, p.dtype, m_ref, l_run, o_addr, HD);
          m_ref += cbias;
        }
        l_run += rsum;
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
      }
      // epilogue: O (accumulated in TMEM over all KV tiles) / l
      mbar_wait(&o_full[g], (n_kv - 1) & 1);
      tc_fence_after();
      const int i = qg0 + r;
      const int odt = DT < 0 ? p.dtype : DT;
      const float inv = (l_run > 0.f ? 1.f / l_run : 0.f) * (DROP ? p.drop_scale : 1.f);
      uint16_t* orow = static_cast<uint16_t*>(p.o) + b * p.o_bs + static_cast<long long>(min(i, p.Lq - 1)) * p.o_rs +
                       static_cast<long long>(h) * p.hd;
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(o_addr + c0, v);
        tmem_ld_wait();
        if (i < p.Lq) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            if (c0 + c < p.hd) {
              *reinterpret_cast<uint4*>(orow + c0 + c) =
                  make_uint4(pack2(__uint_as_float(v[c]) * inv, __uint_as_float(v[c + 1]) * inv, odt),
                             pack2(__uint_as_float(v[c + 2]) * inv, __uint_as_float(v[c + 3]) * inv, odt),
                             pack2(__uint_as_float(v[c + 4]) * inv, __uint_as_float(v[c + 5]) * inv, odt),
                             pack2(__uint_as_float(v[c + 6]) * inv, __uint_as_float(v[c + 7]) * inv, odt));
            }
          }
        }
      }
      if (i < p.Lq) {
        if (p.lse) p.lse[(static_cast<long long>(b) * p.H + h) * p.Lq + i] = (m_ref + log2f(l_run)) * 0.6931471805599453f;
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

// ---------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
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

// 4-D view (d, head, token, batch) of a [B, L, heads*hd] 16-bit tensor; box = box_d x 1 x 128 x 1
static int make_tmap4(CUtensorMap* map, const void* base, int dtype, int hd, int heads, int L, int B, long long rs,
                      long long bs, int box_d, bool sw64) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return MRB_ERR_CUDA;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(hd), static_cast<cuuint64_t>(heads), static_cast<cuuint64_t>(L),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(hd) * 2, static_cast<cuuint64_t>(rs) * 2, static_cast<cuuint64_t>(bs) * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_d), 1, 128, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, dtype == MRB_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                  const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MRB_OK : MRB_ERR_CUDA;
}

template <int HD, int G, int DT, bool DROP = false>
static int launch_tc(const CUtensorMap* maps, const AttnTcParams& p, cudaStream_t s) {
  using S = TcSmem<HD, G>;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<HD, G, DT, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return mrb_set_error(e);
    cfg = true;
  }
  dim3 grid((p.Lq + G * TQ - 1) / (G * TQ), p.H, p.B);
  MRB_LAUNCH((attn_fwd_tc_kernel<HD, G, DT, DROP>), grid, S::THREADS, S::TOTAL, s, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

}  // namespace mrb

using namespace mrb;

// Same contract as mrb_attention_fwd (include/mrblip_b200.h); requires hd == 64 or 64 < hd <= 96 with hd % 8 == 0.
static int attention_fwd_tc_impl(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                 const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                                 int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                                 int bias_len, int bias_zero, const int* kmask, int kv_div, int causal, int q_pos0,
                                 float* lse, const unsigned* drop_seed, unsigned drop_site, float drop_p, void* stream) {
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0) return MRB_OK;
  if (dtype != MRB_DT_F16 && dtype != MRB_DT_BF16) return MRB_ERR_ARG;
  if (!(hd == 64 || (hd > 64 && hd <= 96 && (hd & 7) == 0))) return MRB_ERR_UNSUPPORTED;
  if ((q_rs | k_rs | v_rs | o_rs | q_bs | k_bs | v_bs | o_bs) & 7) return MRB_ERR_ARG;
  if (kv_div <= 0) kv_div = 1;
  const int Bkv = (B + kv_div - 1) / kv_div;
  CUtensorMap maps[6];
  const bool split = hd > 64;
  int rc = make_tmap4(&maps[0], q, dtype, hd, H, Lq, B, q_rs, q_bs, 64, false);
  if (!rc) rc = make_tmap4(&maps[1], k, dtype, hd, H, Lk, Bkv, k_rs, k_bs, 64, false);
  if (!rc) rc = make_tmap4(&maps[2], v, dtype, hd, H, Lk, Bkv, v_rs, v_bs, 64, false);
  if (!rc && split) {
    rc = make_tmap4(&maps[3], q, dtype, hd, H, Lq, B, q_rs, q_bs, 32, true);
    if (!rc) rc = make_tmap4(&maps[4], k, dtype, hd, H, Lk, Bkv, k_rs, k_bs, 32, true);
    if (!rc) rc = make_tmap4(&maps[5], v, dtype, hd, H, Lk, Bkv, v_rs, v_bs, 32, true);
  } else if (!rc) {
    maps[3] = maps[0]; maps[4] = maps[1]; maps[5] = maps[2];
  }
  if (rc) return rc;
  AttnTcParams p{};
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.hd = hd; p.dtype = dtype; p.scale = scale;
  p.bias = bias; p.bias_len = bias_len; p.bias_zero = bias_zero; p.kmask = kmask; p.kv_div = kv_div;
  p.causal = causal; p.q_pos0 = q_pos0; p.o = o; p.o_bs = o_bs; p.o_rs = o_rs; p.lse = lse;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (drop_seed && drop_p > 0.f) {
    // T5 in train mode: hd 64, bf16.  The SWAR mask compare needs an even threshold (p = 0.1 -> 26 / 256).
    const DropSpec d = make_drop(drop_seed, drop_site, drop_p);
    if (split || dtype != MRB_DT_BF16 || (d.thr & 1u) || drop_p >= 1.f) return MRB_ERR_UNSUPPORTED;
    p.drop_seed = d.seed; p.drop_site = d.site; p.drop_thr7 = (d.thr >> 1) * 0x01010101u; p.drop_scale = d.scale;
    return launch_tc<64, 2, MRB_DT_BF16, true>(maps, p, s);
  }
  // short sequences (ViT: 257 keys = 3 K/V tiles): one softmax group per CTA, two CTAs per SM (MRB_ATTN_G1=0 disables)
  static int g1 = -1;
  if (g1 < 0) { const char* e = getenv("MRB_ATTN_G1"); g1 = (e && e[0] == '0') ? 0 : 1; }
  if (split) {
    if (g1 && Lk <= 4 * TKV) return dtype == MRB_DT_F16 ? launch_tc<96, 1, MRB_DT_F16>(maps, p, s) : launch_tc<96, 1, -1>(maps, p, s);
    return launch_tc<96, 2, -1>(maps, p, s);
  }
  return dtype == MRB_DT_BF16 ? launch_tc<64, 2, MRB_DT_BF16>(maps, p, s) : launch_tc<64, 2, -1>(maps, p, s);
}

extern "C" int mrb_attention_fwd_tc(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                    const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                                    int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                                    int bias_len, int bias_zero, const int* kmask, int kv_div, int causal, int q_pos0,
                                    float* lse, void* stream) {
  return attention_fwd_tc_impl(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, B, H, Lq, Lk, hd, dtype, scale, bias, bias_len,
                               bias_zero, kmask, kv_div, causal, q_pos0, lse, nullptr, 0u, 0.f, stream);
}

// mrb_attention_fwd_tc with train-mode dropout of the attention probabilities (modeling_t5.py:600): O = drop(softmax(S)) V, lse
// unchanged.  hd 64, bf16, round(256 p) even.  Mask: dropmask.cuh with row = (b H + h) Lq + i, column = key index.
extern "C" int mrb_attention_fwd_tc_drop(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                         const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                                         int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                                         int bias_len, int bias_zero, const int* kmask, int kv_div, int causal, int q_pos0,
                                         float* lse, const unsigned* seed, unsigned site, float p, void* stream) {
  if (!seed) return MRB_ERR_ARG;
  return attention_fwd_tc_impl(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, B, H, Lq, Lk, hd, dtype, scale, bias, bias_len,
                               bias_zero, kmask, kv_div, causal, q_pos0, lse, seed, site, p, stream);
}
