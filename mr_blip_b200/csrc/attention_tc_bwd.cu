// tcgen05 flash-attention backward for T5 (hd 64; modeling_t5.py:561-610 under autograd): two kernels built from one
// template.  Forward recap: S = scale*Q K^T + bias + mask, P = softmax(S), O = P V, lse = logsumexp(S).
//   MODE_DKV  CTA owns 2 x 128 keys (two softmax groups); streams 64-query tiles (Q_t, dO_t):
//               S^T  = K_g Q_t^T        dP^T = V_g dO_t^T           (128 x 64 fp32 tiles in TMEM)
//               P^T  = exp(S^T - lse),  dS^T = P^T * (dP^T - delta) * scale   -> 16-bit, packed back over the S^T / dP^T TMEM columns
//               dV_g += P^T dO_t        dK_g += dS^T Q_t             (accumulated in TMEM over all query tiles)
//   MODE_DQ   CTA owns 2 x 128 queries; streams 64-key tiles (K_t, V_t):
//               S = Q_g K_t^T, dP = dO_g V_t^T, dS = P * (dP - delta) * scale,  dQ_g += dS K_t
// Same warp roles and mbarrier protocol as attention_tc.cu (TMA producer / single-thread MMA issuer / 2 x 4 warps with
// one thread per TMEM lane).  The B operands of the accumulating MMAs (dO_t, Q_t, K_t) are read MN-major straight from
// their [row][d] tiles, so nothing is transposed in memory.
#include "common.cuh"
#include "dropmask.cuh"
#include "attn_delta.cuh"

namespace mrb {

struct AttnBwdParams {
  int B, H, Lq, Lk, dtype;
  float scale;
  const float* bias; int bias_len, bias_zero;
  const int* kmask;
  int causal, q_pos0;
  const float* lse; const float* delta;          // [B, H, Lq]
  void* out1; long long o1_bs, o1_rs;            // DKV: dK ; DQ: dQ
  void* out2; long long o2_bs, o2_rs;            // DKV: dV
};
struct AttnBwdDropParams : AttnBwdParams {
  // DROP instantiations only (attention-probability dropout, modeling_t5.py:600; masks of dropmask.cuh, row = (b H + h) Lq + i,
  // column = key j).  With keep mask m and s = 1 / (1 - p):  dV = s (m P)^T dO,  dS = P * (s m dP - delta) * scale  (delta is
  // rowsum(dO * O) of the DROPPED output, which is what the forward saved), dK / dQ from dS as before.
  const uint32_t* drop_seed; uint32_t drop_site, drop_thr; float drop_scale;
};

template <bool DROP> struct BwdParamsOf { typedef AttnBwdParams type; };
template <> struct BwdParamsOf<true> { typedef AttnBwdDropParams type; };

// what the element loop needs to recompute the mask (DROP only)
struct BwdDrop {
  uint32_t key, thr, thr7, ng;
  uint32_t rowbase;        // DQ: ((b H + h) Lq + i) * ng of this thread's query row;  DKV: ((b H + h) Lq) * ng of the head
  uint32_t jg, sh;         // DKV: word index (key >> 2) and bit offset 8 * (key & 3) of this thread's key
};
// 0xffffffff if bit 8 i + 7 of t is set (t from the SWAR compare of four draws, attention_tc.cu drop_pair_masks), else 0
template <int I>
__device__ __forceinline__ uint32_t sign_mask(uint32_t t) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(t), "r"(0u), "r"(0x8888u + 0x1111u * I));
  return m;
}

constexpr int MODE_DKV = 0, MODE_DQ = 1;
constexpr int TS = 128;      // stationary rows per group
constexpr int TT = 64;       // streamed rows per tile
constexpr int BHD = 64;

__device__ __forceinline__ void tma_load_4d_b(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ float ex2b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t udesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {   // SWIZZLE_128B
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_b(int fmt, int M, int N, int b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]: the 16-bit A operand (K-major: TMEM lane = row, one 32-bit column = two consecutive k) is read from
// tensor memory -- P^T / dS^T go from the elementwise warps' registers straight back over the S^T / dP^T columns they came from
// (tcgen05.st) and never touch shared memory: an SS-form MMA of this shape (M 128, N 64, K 16) reads 4 KB of A + 2 KB of B from
// shared memory for 32 clk of tensor work, i.e. 192 B/clk against the 128 B/clk a CTA's shared memory delivers.
__device__ __forceinline__ void bw_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void bw_tmem_st_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// first TMEM column (relative to the S^T resp. dP^T block of a group) of the packed 16-bit operand of streamed columns [c0, c0 + 16):
// each elementwise warp writes over columns it has already read -- warps 2-9 (streamed columns 0-31) into [0, 16), warps 10-17
// (columns 32-63) into [32, 48)
__host__ __device__ constexpr int bw_pcol(int c0) { return c0 < 32 ? (c0 >> 1) : 32 + ((c0 - 32) >> 1); }

#ifndef MRB_BWD_STAGES
#define MRB_BWD_STAGES 6      // streamed-tile ring: 6 x 16 KB since the P^T / dS^T tiles left shared memory (3 stages: 0.709 / 0.928 ms per
#endif                        // encoder layer without / with dropout, 6 stages: 0.659 / 0.876; unused padding instead: no change)
struct BwdSmem {
  static constexpr int X_BYTES = TS * 128;                 // one stationary tile (128 rows x 64 d)
  static constexpr int U_BYTES = TT * 128;                 // one streamed tile (64 rows x 64 d)
  static constexpr int STAGES = MRB_BWD_STAGES;
  static constexpr int OFF_X = 0;                          // [group][X, Y]
  static constexpr int OFF_U = 4 * X_BYTES;                // [stage][U, W]
  static constexpr int EW = 16;                            // elementwise warps: 2 groups x 4 TMEM lane quadrants x 2 column halves
  static constexpr int THREADS = (2 + EW) * 32;
  static constexpr int OFF_WIN = OFF_U + STAGES * 2 * U_BYTES;   // EW warps x (96 bias + 64 lse + 64 delta) floats (P^T / dS^T live in TMEM)
  static constexpr int WIN_FLOATS = 96 + 64 + 64;
  static constexpr int OFF_BAR = OFF_WIN + EW * WIN_FLOATS * 4;
  static constexpr int NBAR = 1 + 2 * STAGES + 6;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16 + 1024;
};

// Elementwise core of one streamed tile for one stationary row (thread pair): P = exp2(S*sl2 + bias - lse) and
// dS = P * (dP - delta) * scale for this warp's 32 of the 64 streamed columns, written as the 16-bit K-major TMEM operand of the
// accumulating MMAs.
//   off   : per-row constant added to every exp2 argument (DQ: [constant tile bias] - lse_row; DKV: 0, the per-column
//           window wlse[] already holds [constant tile bias] - lse)
//   dls   : DQ: delta_row * scale;  DKV: unused (wdl[] holds delta * scale)
// HAS_BIAS = the bias varies inside the tile (window lookup per element); MASKED = some element of the warp's tile is masked.
// Unmasked constant-bias tiles (all but the ~3 tiles next to the diagonal of a T5 encoder row block) cost
// FFMA + MUFU + FFMA + FMUL + the packs per element.
template <int MODE, bool HAS_BIAS, bool MASKED, bool DROP = false>
__device__ __forceinline__ void bwd_tile_rows(uint32_t lane_base, const float* wrow, const float* wlse, const float* wdl,
                                              float off, float dls, float sl2, float scale, uint32_t cm0, uint32_t cm1,
                                              bool row_key_ok, bool causal, int u0, int row_c, int q_pos0,
                                              int r, int dt, int cbeg, const BwdDrop& dd = BwdDrop()) {
  // this warp's half of the streamed tile: columns [cbeg, cbeg + 32) in two 16-column steps (48 live registers per step instead of
  // 96: the kernel runs 18 warps per CTA, i.e. at most 112 registers per thread)
#pragma unroll
  for (int cs = 0; cs < 32; cs += 16) {
    const int c0 = cbeg + cs;
    uint32_t sv[16], dv[16];
    tmem_ld_32x32b_x16(lane_base + c0, sv);
    tmem_ld_32x32b_x16(lane_base + 64 + c0, dv);
    tmem_ld_wait();
    const uint32_t cm = c0 < 32 ? cm0 : cm1;
    const int cb = c0 & 31;                   // bit of column c0 in cm
    uint32_t pp[8], pd[8];
#pragma unroll
    for (int e8 = 0; e8 < 16; e8 += 8) {
      float off8[8], dl8[8];                  // DKV: (bias const - lse) and delta * scale of 8 streamed queries, broadcast loads
      if (MODE == MODE_DKV) {
        const float4 la = *reinterpret_cast<const float4*>(wlse + c0 + e8), lb = *reinterpret_cast<const float4*>(wlse + c0 + e8 + 4);
        const float4 da = *reinterpret_cast<const float4*>(wdl + c0 + e8), db = *reinterpret_cast<const float4*>(wdl + c0 + e8 + 4);
        off8[0] = la.x; off8[1] = la.y; off8[2] = la.z; off8[3] = la.w; off8[4] = lb.x; off8[5] = lb.y; off8[6] = lb.z; off8[7] = lb.w;
        dl8[0] = da.x; dl8[1] = da.y; dl8[2] = da.z; dl8[3] = da.w; dl8[4] = db.x; dl8[5] = db.y; dl8[6] = db.z; dl8[7] = db.w;
      }
      uint32_t t8[2] = {0u, 0u};              // DROP, DQ: SWAR compare results of this thread's keys c0 + e8 .. + 7 (two words)
      if (DROP && MODE == MODE_DQ) {
#pragma unroll
        for (int w2 = 0; w2 < 2; ++w2) {
          const uint32_t w = drop_word(dd.key, dd.rowbase, static_cast<uint32_t>((u0 + c0 + e8) >> 2) + w2);
          t8[w2] = (((w >> 1) & 0x7f7f7f7fu) | 0x80808080u) - dd.thr7;
        }
      }
      // DROP, DKV: a thread owns one KEY and walks over queries, i.e. over mask rows: one word per element.  The four lanes of a
      // key group (keys 4m .. 4m+3 share their words) split the hashing -- lane k of the group evaluates the words of queries
      // c0 + e8 + k and c0 + e8 + 4 + k -- and exchange them by shuffles (the call is warp-uniform, every lane takes part).
      uint32_t wq[2] = {0u, 0u};
      if (DROP && MODE == MODE_DKV) {
#pragma unroll
        for (int w2 = 0; w2 < 2; ++w2)
          wq[w2] = drop_word(dd.key, dd.rowbase + static_cast<uint32_t>(u0 + c0 + e8 + 4 * w2 + (r & 3)) * dd.ng, dd.jg);
      }
#pragma unroll
      for (int e2 = 0; e2 < 8; e2 += 2) {
        const int e = e8 + e2;
        float pr[2], ds[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = c0 + e + q;
          uint32_t keepm = 0xffffffffu;
          if (DROP) {
            if (MODE == MODE_DQ) {
              const uint32_t t = t8[(e2 + q) >> 2];
              keepm = ((e2 + q) & 3) == 0 ? sign_mask<0>(t) : ((e2 + q) & 3) == 1 ? sign_mask<1>(t) : ((e2 + q) & 3) == 2 ? sign_mask<2>(t)
                                                                                                                     : sign_mask<3>(t);
            } else {                            // streamed query u0 + c, this thread's key
              const uint32_t w = __shfl_sync(0xffffffffu, wq[(e2 + q) >> 2], (r & 28) | ((e2 + q) & 3));
              keepm = (((w >> dd.sh) & 0xffu) >= dd.thr) ? 0xffffffffu : 0u;
            }
          }
          float o = (MODE == MODE_DQ) ? off : off8[e2 + q];
          if (HAS_BIAS) o += (MODE == MODE_DQ) ? wrow[c] : wrow[-c];
          float pv = ex2b(fmaf(__uint_as_float(sv[e + q]), sl2, o));
          if (MASKED) {
            bool ok;
            if (MODE == MODE_DQ) ok = ((cm >> (cb + e + q)) & 1u) && !(causal && u0 + c > row_c + q_pos0);
            else ok = row_key_ok && !(causal && row_c > u0 + c + q_pos0);
            pv = ok ? pv : 0.f;
          }
          pr[q] = DROP ? __uint_as_float(__float_as_uint(pv) & keepm) : pv;       // P^T operand of dV (DKV): dropped, unscaled
          ds[q] = pv * fmaf(__uint_as_float(DROP ? (dv[e + q] & keepm) : dv[e + q]), scale, -((MODE == MODE_DQ) ? dls : dl8[e2 + q]));
        }
        if (MODE == MODE_DKV) pp[e >> 1] = pack2(pr[0], pr[1], dt);
        pd[e >> 1] = pack2(ds[0], ds[1], dt);
      }
    }
    // packed pairs (streamed column 2c | 2c + 1 << 16) = the K-major TMEM operand layout; S^T / dP^T columns [c0, c0 + 16) were read above
    if (MODE == MODE_DKV) bw_tmem_st_x8(lane_base + bw_pcol(c0), pp);
    bw_tmem_st_x8(lane_base + 64 + bw_pcol(c0), pd);
  }
  tmem_st_wait();
}

// ENC = the encoder self-attention case that carries almost all of the time (bucketed bias, no causal mask, bf16): those
// three facts become compile-time constants so the per-element loop has no uniform branches left.
template <int MODE, bool ENC, bool DROP = false>
__global__ void __launch_bounds__(BwdSmem::THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                   const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmW,
                   const typename BwdParamsOf<DROP>::type p) {
  mrb::pdl_trigger();   // the successor may become resident and run its set-up; it blocks in its own pdl_wait()
  using S = BwdSmem;
  constexpr int STAGES = S::STAGES;
  constexpr uint32_t TMEM_COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* x_full = bars;
  uint64_t* st_full = bars + 1;
  uint64_t* st_empty = st_full + STAGES;
  uint64_t* t_full = st_empty + STAGES;
  uint64_t* e_full = t_full + 2;
  uint64_t* acc_full = e_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + S::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, x0 = blockIdx.x * (2 * TS);
  const int Lstat = (MODE == MODE_DKV) ? p.Lk : p.Lq;      // stationary / streamed sequence lengths
  const int Lstrm = (MODE == MODE_DKV) ? p.Lq : p.Lk;
  const int n_groups = (x0 + TS < Lstat) ? 2 : 1;
  int t_begin = 0, t_end = (Lstrm + TT - 1) / TT;
  if (p.causal) {
    if (MODE == MODE_DQ) t_end = min(t_end, (min(x0 + n_groups * TS, p.Lq) - 1 + p.q_pos0) / TT + 1);   // keys j <= i_max
    else t_begin = min(t_end - 1, max(0, (x0 - p.q_pos0) / TT));                                       // queries i >= j_min - q_pos0
  }
  const int n_t = t_end - t_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmY); tma_prefetch_desc(&tmU); tma_prefetch_desc(&tmW);
    mbar_init(x_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&st_full[s], 1); mbar_init(&st_empty[s], 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(&t_full[g], 1); mbar_init(&e_full[g], S::EW / 2); mbar_init(&acc_full[g], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  mrb::pdl_wait();      // set-up done; nothing above touches global memory (MRB_PDL, common.cuh)

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    if (elect_one()) {
      mbar_expect_tx(x_full, n_groups * 2 * S::X_BYTES);
      for (int g = 0; g < n_groups; ++g) {
        tma_load_4d_b(smem + S::OFF_X + (2 * g) * S::X_BYTES, &tmX, x_full, 0, h, x0 + g * TS, b);
        tma_load_4d_b(smem + S::OFF_X + (2 * g + 1) * S::X_BYTES, &tmY, x_full, 0, h, x0 + g * TS, b);
      }
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 0;
    for (int t = 0; t < n_t; ++t) {
      mbar_wait(&st_empty[st], ph ^ 1);
      if (elect_one()) {
        uint8_t* su = smem + S::OFF_U + st * 2 * S::U_BYTES;
        mbar_expect_tx(&st_full[st], 2 * S::U_BYTES);
        tma_load_4d_b(su, &tmU, &st_full[st], 0, h, (t_begin + t) * TT, b);
        tma_load_4d_b(su + S::U_BYTES, &tmW, &st_full[st], 0, h, (t_begin + t) * TT, b);
      }
      __syncwarp();
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this code converged and ONE elected lane issues (elect_one, common.cuh); every descriptor is a
    // kernel-lifetime base plus a small offset (the 14-bit address field counts 16-byte units: + 2 per 16-element K step of a
    // K-major tile, + 128 per 16-row step of an MN-major one).  With `if (lane == 0)` and descriptors rebuilt from addresses each
    // tcgen05.mma cost ~20 dependent uniform-datapath instructions plus an ELECT / BRA.U.ANY loop -- ~150 clk per instruction that
    // keeps the tensor pipe busy for 32 clk: the issuing warp, not ex2 or tensor memory, was what bounded the kernel
    // (profiles/ncu_attn_issue_r02d.md).
    const int fmt = p.dtype == MRB_DT_BF16 ? 1 : 0;
    const uint32_t id_t = idesc_b(fmt, TS, TT, 0);          // 128 x 64, both operands K-major (contraction over d)
    const uint32_t id_a = idesc_b(fmt, TS, BHD, 1);         // 128 x 64(d), B MN-major (contraction over streamed rows)
    const uint64_t dk0 = udesc(smem_u32(smem), 16, 1024);         // K-major SWIZZLE_128B tile at the start of shared memory
    const uint64_t dm0 = udesc(smem_u32(smem), TT * 128, 1024);   // MN-major view of a streamed [row][d] tile
    auto issue_T = [&](int g, int st) {
      const uint64_t xd = dk0 + static_cast<uint32_t>((S::OFF_X + (2 * g) * S::X_BYTES) >> 4), yd = xd + (S::X_BYTES >> 4);
      const uint64_t ud = dk0 + static_cast<uint32_t>((S::OFF_U + st * 2 * S::U_BYTES) >> 4), wd = ud + (S::U_BYTES >> 4);
      const uint32_t d1 = tmem_base + g * 256, d2 = d1 + 64;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(d1, xd + 2 * k, ud + 2 * k, id_t, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(d2, yd + 2 * k, wd + 2 * k, id_t, k > 0 ? 1u : 0u);
      umma_commit(&t_full[g]);
    };
    auto issue_acc = [&](int g, int st, uint32_t acc) {
      const uint64_t um = dm0 + static_cast<uint32_t>((S::OFF_U + st * 2 * S::U_BYTES) >> 4), wm = um + (S::U_BYTES >> 4);
      const uint32_t pt = tmem_base + g * 256, st_ = pt + 64;          // packed P^T over the S^T columns, dS^T over the dP^T columns
      const uint32_t a1 = tmem_base + g * 256 + 128, a2 = a1 + 64;
      if (MODE == MODE_DKV) {
#pragma unroll
        for (int k = 0; k < 4; ++k) bw_umma_ts(a1, pt + bw_pcol(16 * k), wm + 128 * k, id_a, (k > 0) ? 1u : acc);      // dV += P^T dO_t
#pragma unroll
        for (int k = 0; k < 4; ++k) bw_umma_ts(a2, st_ + bw_pcol(16 * k), um + 128 * k, id_a, (k > 0) ? 1u : acc);     // dK += dS^T Q_t
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) bw_umma_ts(a1, st_ + bw_pcol(16 * k), um + 128 * k, id_a, (k > 0) ? 1u : acc);     // dQ += dS K_t
      }
    };
    mbar_wait(x_full, 0);
    mbar_wait(&st_full[0], 0);
    tc_fence_after();
    if (elect_one())
      for (int g = 0; g < n_groups; ++g) issue_T(g, 0);
    __syncwarp();
    int st = 0, st_next = (STAGES > 1) ? 1 : 0;              // t % STAGES, (t + 1) % STAGES
    uint32_t ph_next = 0;                                    // ((t + 1) / STAGES) & 1
    for (int t = 0; t < n_t; ++t) {
      for (int g = 0; g < n_groups; ++g) {
        mbar_wait(&e_full[g], t & 1);
        if (g == 0 && t + 1 < n_t) mbar_wait(&st_full[st_next], ph_next);
        tc_fence_after();
        if (elect_one()) {
          issue_acc(g, st, t > 0 ? 1u : 0u);
          if (g == n_groups - 1) umma_commit(&st_empty[st]);
          if (t + 1 < n_t) issue_T(g, st_next);
          else umma_commit(&acc_full[g]);
        }
        __syncwarp();
      }
      st = st_next;
      if (++st_next == STAGES) { st_next = 0; ph_next ^= 1; }
    }
  } else {
    // ===================== elementwise groups: TWO threads per stationary row =====================
    // Warps 2-9 take columns [0, 32) of their group's 128 x 64 streamed tile, warps 10-17 columns [32, 64) (a warp reads the TMEM
    // lanes of quadrant warp % 4, so warps w and w + 8 share their 32 rows).  The backward has no row reductions -- every element
    // of P / dS is a function of its own S, dP and the row / column constants -- so the split needs no exchange; what it buys is
    // four resident elementwise warps per scheduler instead of two for a loop that was bound by its own latency (tensor pipe 31-34 %,
    // issue slots 40 %: profiles/ncu_attn_t5_r02d.md).
    const int ew = warp - 2, g = (ew >> 2) & 1, cbeg = (ew >> 3) * 32;
    if (g < n_groups) {
      const int quad = warp & 3;
      const int r = quad * 32 + lane;
      const int xg0 = x0 + g * TS;
      const int row = xg0 + r;                               // key index (DKV) or query index (DQ)
      const int row_c = min(row, Lstat - 1);
      const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + g * 256;
      const float LOG2E = 1.4426950408889634f;
      const float sl2 = p.scale * LOG2E;
      const float* bhead = p.bias ? p.bias + static_cast<long long>(h) * p.bias_len : nullptr;
      const bool has_bias = ENC ? true : (bhead != nullptr);
      const bool causal = ENC ? false : (p.causal != 0);
      const int dt = ENC ? MRB_DT_BF16 : p.dtype;
      const float scale = p.scale;
      const long long stat_off = (static_cast<long long>(b) * p.H + h) * p.Lq;
      float* win = reinterpret_cast<float*>(smem + S::OFF_WIN) + (warp - 2) * S::WIN_FLOATS;
      float* wlse = win + 96;
      float* wdl = win + 160;
      // per-row constants
      float lse_row = 0.f, dl_row = 0.f;
      bool row_key_ok = true;
      if (MODE == MODE_DQ) {
        lse_row = p.lse[stat_off + row_c] * LOG2E;
        dl_row = p.delta[stat_off + row_c];
      } else if (p.kmask) {
        row_key_ok = __ldg(p.kmask + static_cast<long long>(b) * p.Lk + row_c) != 0;
      }
      const int warp_row_first = min(xg0 + quad * 32, Lstat - 1), warp_row_last = min(xg0 + quad * 32 + 31, Lstat - 1);
      BwdDrop dd = BwdDrop();
      float tile_scale = scale;                                          // multiplies dP only (delta * scale is staged separately)
      if constexpr (DROP) {
        tile_scale = scale * p.drop_scale;
        dd.key = drop_key(*p.drop_seed, p.drop_site);
        dd.thr = p.drop_thr; dd.thr7 = (p.drop_thr >> 1) * 0x01010101u;
        dd.ng = drop_groups(static_cast<uint32_t>(p.Lk));
        const uint32_t head_row = static_cast<uint32_t>(b * p.H + h) * static_cast<uint32_t>(p.Lq);
        dd.rowbase = (MODE == MODE_DQ) ? (head_row + static_cast<uint32_t>(row_c)) * dd.ng : head_row * dd.ng;
        dd.jg = static_cast<uint32_t>(row_c) >> 2; dd.sh = 8u * (static_cast<uint32_t>(row_c) & 3u);
      }

      // per-tile global loads (bias window, lse / delta of the streamed queries, key mask) run one tile ahead of their use
      float bnext[3] = {0.f, 0.f, 0.f}, lnext[2] = {0.f, 0.f}, dnext[2] = {0.f, 0.f};
      int mnext[2] = {1, 1};
      auto prefetch_tile = [&](int tn) {
        const int un = (t_begin + tn) * TT;
        if (has_bias) {
          const int w0 = ((MODE == MODE_DQ) ? (un - (warp_row_last + p.q_pos0) + p.bias_zero)
                                            : (warp_row_first - (un + 63 + p.q_pos0) + p.bias_zero)) + lane;
#pragma unroll
          for (int tt = 0; tt < 3; ++tt) {
            const int idx = w0 + 32 * tt;
            bnext[tt] = (idx >= 0 && idx < p.bias_len) ? __ldg(bhead + idx) : 0.f;
          }
        }
        if (MODE == MODE_DKV) {
#pragma unroll
          for (int tt = 0; tt < 2; ++tt) {
            const int i = un + lane + 32 * tt;
            lnext[tt] = (i < p.Lq) ? p.lse[stat_off + i] : 0.f;
            dnext[tt] = (i < p.Lq) ? p.delta[stat_off + i] : 0.f;
          }
        } else {
          const int* mrow = p.kmask ? p.kmask + static_cast<long long>(b) * p.Lk : nullptr;
#pragma unroll
          for (int tt = 0; tt < 2; ++tt) {
            const int jn = un + lane + 32 * tt;
            mnext[tt] = (jn < p.Lk) ? (mrow ? __ldg(mrow + jn) : 1) : 0;
          }
        }
      };
      prefetch_tile(0);
      for (int t = 0; t < n_t; ++t) {
        const int u0 = (t_begin + t) * TT;                   // first streamed row (query for DKV, key for DQ)
        __syncwarp();
        // ---- stage the per-warp windows (their global loads were issued one tile ahead): bias (pre-multiplied by log2 e)
        //      and, for DKV, lse / delta of the 64 streamed queries
        bool bias_const = false;
        float cbias = 0.f;
        if (has_bias) {
          bool same = true;
          const float first = bnext[0] * LOG2E;
#pragma unroll
          for (int tt = 0; tt < 3; ++tt) {
            const float bv = bnext[tt] * LOG2E;
            win[lane + 32 * tt] = bv;
            same = same && (bv == first);
          }
          const float lane0 = __shfl_sync(0xffffffffu, first, 0);
          bias_const = __all_sync(0xffffffffu, same && first == lane0);   // T5 buckets saturate 128 positions off the diagonal
          cbias = bias_const ? lane0 : 0.f;
        }
        uint32_t cm0 = 0xffffffffu, cm1 = 0xffffffffu;       // per-column validity bits (keys for DQ)
        bool masked = causal;
        if (MODE == MODE_DKV) {
#pragma unroll
          for (int tt = 0; tt < 2; ++tt) {
            wlse[lane + 32 * tt] = cbias - lnext[tt] * LOG2E;    // exp2 argument offset of streamed query k
            wdl[lane + 32 * tt] = dnext[tt] * scale;
          }
          masked = masked || !__all_sync(0xffffffffu, row_key_ok);
        } else {
          cm0 = __ballot_sync(0xffffffffu, mnext[0] != 0);
          cm1 = __ballot_sync(0xffffffffu, mnext[1] != 0);
          masked = masked || (cm0 & cm1) != 0xffffffffu;
        }
        if (t + 1 < n_t) prefetch_tile(t + 1);
        __syncwarp();
        const float* wrow = (MODE == MODE_DQ) ? (win + (warp_row_last - row_c)) : (win + (row_c - warp_row_first) + 63);
        const float off = (MODE == MODE_DQ) ? (cbias - lse_row) : 0.f;
        const float dls = dl_row * scale;
        const bool vbias = has_bias && !bias_const;

        mbar_wait(&t_full[g], t & 1);
        tc_fence_after();
        // warp-uniform dispatch (tcgen05.ld inside is warp-collective)
        if (vbias) {
          if (masked) bwd_tile_rows<MODE, true, true, DROP>(lane_base, wrow, wlse, wdl, off, dls, sl2, tile_scale, cm0, cm1, row_key_ok, causal, u0, row_c, p.q_pos0, r, dt, cbeg, dd);
          else bwd_tile_rows<MODE, true, false, DROP>(lane_base, wrow, wlse, wdl, off, dls, sl2, tile_scale, cm0, cm1, row_key_ok, causal, u0, row_c, p.q_pos0, r, dt, cbeg, dd);
        } else {
          if (masked) bwd_tile_rows<MODE, false, true, DROP>(lane_base, wrow, wlse, wdl, off, dls, sl2, tile_scale, cm0, cm1, row_key_ok, causal, u0, row_c, p.q_pos0, r, dt, cbeg, dd);
          else bwd_tile_rows<MODE, false, false, DROP>(lane_base, wrow, wlse, wdl, off, dls, sl2, tile_scale, cm0, cm1, row_key_ok, causal, u0, row_c, p.q_pos0, r, dt, cbeg, dd);
        }
        tc_fence_before();                                  // the operand tiles were written with tcgen05.st (waited for in bwd_tile_rows)
        __syncwarp();
        if (lane == 0) mbar_arrive(&e_full[g]);
      }
      // ---- write the accumulators
      mbar_wait(&acc_full[g], 0);
      tc_fence_after();
      const int n_out = (MODE == MODE_DKV) ? 2 : 1;
      for (int a = 0; a < n_out; ++a) {
        // DKV: TMEM acc1 = dV -> out2, acc2 = dK -> out1 ; DQ: acc1 = dQ -> out1
        void* outp = (MODE == MODE_DKV) ? (a == 0 ? p.out2 : p.out1) : p.out1;
        const long long bs = (MODE == MODE_DKV && a == 0) ? p.o2_bs : p.o1_bs;
        const long long rs = (MODE == MODE_DKV && a == 0) ? p.o2_rs : p.o1_rs;
        uint16_t* orow = static_cast<uint16_t*>(outp) + b * bs + static_cast<long long>(row_c) * rs + static_cast<long long>(h) * BHD;
        {
          const int c0 = cbeg;                                  // this warp's half of the 64 accumulator columns
          uint32_t v[32];
          tmem_ld_32x32b_x32(lane_base + 128 + a * 64 + c0, v);
          tmem_ld_wait();
          if constexpr (DROP) {
            if (MODE == MODE_DKV && a == 0) {                   // dV = 1 / (1 - p) * (m P)^T dO
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) * p.drop_scale);
            }
          }
          if (row < Lstat) {
#pragma unroll
            for (int c = 0; c < 32; c += 8)
              *reinterpret_cast<uint4*>(orow + c0 + c) =
                  make_uint4(pack2(__uint_as_float(v[c]), __uint_as_float(v[c + 1]), dt),
                             pack2(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]), dt),
                             pack2(__uint_as_float(v[c + 4]), __uint_as_float(v[c + 5]), dt),
                             pack2(__uint_as_float(v[c + 6]), __uint_as_float(v[c + 7]), dt));
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

// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]  (16-bit inputs, hd = 64).  Thread = 8 consecutive d of one (b, i, h): one 16-byte
// load of O and of dO, the 8 lanes of a head meet in three shuffles; a warp covers 4 heads = 512 contiguous bytes of a token row.
// (The first version gave a warp one (b, h, i) row: 4-byte loads, 128 B per warp instruction, 47 us for 134 MB at the QVH shape.)
__global__ void __launch_bounds__(256) attn_delta64_kernel(const uint16_t* __restrict__ o, long long o_bs, long long o_rs,
                                                           const uint16_t* __restrict__ d_o, long long do_bs, long long do_rs,
                                                           float* __restrict__ delta, int B, int H, int Lq, int dtype) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;      // ((b Lq + i) H + h) 8 + chunk
  const long long total = static_cast<long long>(B) * Lq * H * 8;
  const bool live = t < total;
  const long long tt = live ? t : total - 1;            // every lane takes part in the shuffles
  const int chunk = static_cast<int>(tt & 7), h = static_cast<int>((tt >> 3) % H);
  const long long bi = (tt >> 3) / H;
  const int i = static_cast<int>(bi % Lq), b = static_cast<int>(bi / Lq);
  const uint4 ov = *reinterpret_cast<const uint4*>(o + b * o_bs + static_cast<long long>(i) * o_rs + h * 64 + chunk * 8);
  const uint4 dv = *reinterpret_cast<const uint4*>(d_o + b * do_bs + static_cast<long long>(i) * do_rs + h * 64 + chunk * 8);
  const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) acc += unpack_lo(ow[e], dtype) * unpack_lo(dw[e], dtype) + unpack_hi(ow[e], dtype) * unpack_hi(dw[e], dtype);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (live && chunk == 0) delta[(static_cast<long long>(b) * H + h) * Lq + i] = acc;
}

// ---------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn_b() {
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
static int make_tmap4b(CUtensorMap* map, const void* base, int dtype, int heads, int L, int B, long long rs, long long bs,
                       int box_rows) {
  EncodeTiledFn fn = encode_fn_b();
  if (!fn) return MRB_ERR_CUDA;
  cuuint64_t gdim[4] = {64, static_cast<cuuint64_t>(heads), static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[3] = {128, static_cast<cuuint64_t>(rs) * 2, static_cast<cuuint64_t>(bs) * 2};
  cuuint32_t box[4] = {64, 1, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, dtype == MRB_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                  const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MRB_OK : MRB_ERR_CUDA;
}

template <int MODE, bool ENC, bool DROP = false>
static int launch_bwd_tc2(const CUtensorMap& x, const CUtensorMap& y, const CUtensorMap& u, const CUtensorMap& w,
                          const AttnBwdDropParams& p, int Lstat, cudaStream_t s) {
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel<MODE, ENC, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::TOTAL);
    if (e != cudaSuccess) return mrb_set_error(e);
    cfg = true;
  }
  dim3 grid((Lstat + 2 * TS - 1) / (2 * TS), p.H, p.B);
  MRB_LAUNCH((attn_bwd_tc_kernel<MODE, ENC, DROP>), grid, BwdSmem::THREADS, BwdSmem::TOTAL, s, x, y, u, w,
             static_cast<const typename BwdParamsOf<DROP>::type&>(p));
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
template <int MODE>
static int launch_bwd_tc(const CUtensorMap& x, const CUtensorMap& y, const CUtensorMap& u, const CUtensorMap& w,
                         const AttnBwdDropParams& p, int Lstat, cudaStream_t s) {
  static int spec = -1;                   // MRB_ATTN_BWD_ENC=0 keeps the generic instantiation (A/B measurements)
  if (spec < 0) { const char* e = getenv("MRB_ATTN_BWD_ENC"); spec = (e && e[0] == '0') ? 0 : 1; }
  if (p.drop_seed) {                      // train-mode dropout of the probabilities: same two specialisations
    if (p.bias && !p.causal && p.dtype == MRB_DT_BF16) return launch_bwd_tc2<MODE, true, true>(x, y, u, w, p, Lstat, s);
    return launch_bwd_tc2<MODE, false, true>(x, y, u, w, p, Lstat, s);
  }
  if (spec && p.bias && !p.causal && p.dtype == MRB_DT_BF16) return launch_bwd_tc2<MODE, true>(x, y, u, w, p, Lstat, s);
  return launch_bwd_tc2<MODE, false>(x, y, u, w, p, Lstat, s);
}

}  // namespace mrb

using namespace mrb;

static int attention_bwd_tc_impl(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                 const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                                 const void* dout, long long do_bs, long long do_rs, void* dq, void* dk, void* dv,
                                 int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                                 int bias_len, int bias_zero, const int* kmask, int causal, int q_pos0, const float* lse,
                                 float* delta_ws, const unsigned* drop_seed, unsigned drop_site, float drop_p, void* stream) {
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0) return MRB_OK;
  if (hd != 64) return MRB_ERR_UNSUPPORTED;
  if (dtype != MRB_DT_F16 && dtype != MRB_DT_BF16) return MRB_ERR_ARG;
  if ((q_rs | k_rs | v_rs | o_rs | do_rs | q_bs | k_bs | v_bs | o_bs | do_bs) & 7) return MRB_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (Lq <= DELTA_EXACT_MAX_LQ) {
    // few query rows (the decoder's cross-attention): delta = sum_j P_ij dP_ij recomputed exactly, see attn_delta.cuh
    DeltaExactParams d{};
    d.q = static_cast<const uint16_t*>(q); d.k = static_cast<const uint16_t*>(k); d.v = static_cast<const uint16_t*>(v);
    d.dout = static_cast<const uint16_t*>(dout);
    d.q_bs = q_bs; d.q_rs = q_rs; d.k_bs = k_bs; d.k_rs = k_rs; d.v_bs = v_bs; d.v_rs = v_rs; d.do_bs = do_bs; d.do_rs = do_rs;
    d.B = B; d.H = H; d.Lq = Lq; d.Lk = Lk; d.dtype = dtype; d.scale = scale; d.bias = bias; d.bias_len = bias_len; d.bias_zero = bias_zero;
    d.kmask = kmask; d.causal = causal; d.q_pos0 = q_pos0; d.lse = lse; d.delta = delta_ws;
    if (drop_seed && drop_p > 0.f) {
      const DropSpec ds = make_drop(drop_seed, drop_site, drop_p);
      d.drop_seed = ds.seed; d.drop_site = ds.site; d.drop_thr = ds.thr; d.drop_scale = ds.scale;
    }
    if (int rc = launch_delta_exact(d, s)) return rc;
  } else {
    const int rows = B * H * Lq;
    MRB_LAUNCH((attn_delta64_kernel), static_cast<unsigned>((static_cast<long long>(rows) * 8 + 255) / 256), 256, 0, s, static_cast<const uint16_t*>(o), o_bs, o_rs,
                                                         static_cast<const uint16_t*>(dout), do_bs, do_rs, delta_ws, B, H, Lq, dtype);
    MRB_CHECK_LAUNCH();
  }
  AttnBwdDropParams p{};
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.dtype = dtype; p.scale = scale;
  p.bias = bias; p.bias_len = bias_len; p.bias_zero = bias_zero; p.kmask = kmask; p.causal = causal; p.q_pos0 = q_pos0;
  p.lse = lse; p.delta = delta_ws;
  if (drop_seed && drop_p > 0.f) {
    const DropSpec d = make_drop(drop_seed, drop_site, drop_p);
    if ((d.thr & 1u) || drop_p >= 1.f) return MRB_ERR_UNSUPPORTED;        // the SWAR compare of the dQ kernel needs an even threshold
    p.drop_seed = d.seed; p.drop_site = d.site; p.drop_thr = d.thr; p.drop_scale = d.scale;
  }
  CUtensorMap mq128, mk128, mv128, mdo128, mq64, mk64, mv64, mdo64;
  int rc = make_tmap4b(&mq128, q, dtype, H, Lq, B, q_rs, q_bs, TS);
  if (!rc) rc = make_tmap4b(&mk128, k, dtype, H, Lk, B, k_rs, k_bs, TS);
  if (!rc) rc = make_tmap4b(&mv128, v, dtype, H, Lk, B, v_rs, v_bs, TS);
  if (!rc) rc = make_tmap4b(&mdo128, dout, dtype, H, Lq, B, do_rs, do_bs, TS);
  if (!rc) rc = make_tmap4b(&mq64, q, dtype, H, Lq, B, q_rs, q_bs, TT);
  if (!rc) rc = make_tmap4b(&mk64, k, dtype, H, Lk, B, k_rs, k_bs, TT);
  if (!rc) rc = make_tmap4b(&mv64, v, dtype, H, Lk, B, v_rs, v_bs, TT);
  if (!rc) rc = make_tmap4b(&mdo64, dout, dtype, H, Lq, B, do_rs, do_bs, TT);
  if (rc) return rc;
  // dK, dV: stationary K, V; streamed Q, dO
  p.out1 = dk; p.o1_bs = k_bs; p.o1_rs = k_rs; p.out2 = dv; p.o2_bs = v_bs; p.o2_rs = v_rs;
  rc = launch_bwd_tc<MODE_DKV>(mk128, mv128, mq64, mdo64, p, Lk, s);
  if (rc) return rc;
  // dQ: stationary Q, dO; streamed K, V
  p.out1 = dq; p.o1_bs = q_bs; p.o1_rs = q_rs; p.out2 = nullptr;
  return launch_bwd_tc<MODE_DQ>(mq128, mdo128, mk64, mv64, p, Lq, s);
}

// Same contract as mrb_attention_bwd; hd must be 64.
extern "C" int mrb_attention_bwd_tc(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                    const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                                    const void* dout, long long do_bs, long long do_rs, void* dq, void* dk, void* dv,
                                    int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                                    int bias_len, int bias_zero, const int* kmask, int causal, int q_pos0, const float* lse,
                                    float* delta_ws, void* stream) {
  return attention_bwd_tc_impl(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, dout, do_bs, do_rs, dq, dk, dv, B, H, Lq, Lk, hd,
                               dtype, scale, bias, bias_len, bias_zero, kmask, causal, q_pos0, lse, delta_ws, nullptr, 0u, 0.f, stream);
}

// Backward of mrb_attention_fwd_tc_drop (same seed word, site and p: the mask is recomputed, nothing was stored).
extern "C" int mrb_attention_bwd_tc_drop(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                         const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                                         const void* dout, long long do_bs, long long do_rs, void* dq, void* dk, void* dv,
                                         int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                                         int bias_len, int bias_zero, const int* kmask, int causal, int q_pos0, const float* lse,
                                         float* delta_ws, const unsigned* seed, unsigned site, float p, void* stream) {
  if (!seed) return MRB_ERR_ARG;
  return attention_bwd_tc_impl(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, dout, do_bs, do_rs, dq, dk, dv, B, H, Lq, Lk, hd,
                               dtype, scale, bias, bias_len, bias_zero, kmask, causal, q_pos0, lse, delta_ws, seed, site, p, stream);
}
