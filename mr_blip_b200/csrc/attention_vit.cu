// EVA ViT self-attention (eva_vit.py:128-145) as ONE persistent tcgen05 kernel: 257 tokens = CLS + 256 patches, hd 88, no mask.
//
// Why a second forward kernel next to attention_tc.cu: with 257 keys the generic flash kernel runs three KV tiles per 128-row
// CTA whose QK -> softmax -> PV chain is strictly serial, pays TMEM allocation / barrier set-up / first-load latency per CTA
// (12 us of CTA lifetime for ~1.3 us of tensor work: profiles/ncu_attn_vit_fwd_r02a.md, tensor pipe 15 % active) and needs a
// separate CUDA-core kernel for the 257th query row.  Here:
//   * one CTA per SM loops over (frame, head) items; the next item's Q / K (then V) stream in by TMA while the current item is
//     in its softmax / PV / epilogue phase, so load latency and set-up are paid once per CTA, not once per item;
//   * the 256 patch tokens are the tensor-core problem: two groups of 128 query rows, S_g = Q_g K^T is ONE 128 x 256 MMA per
//     16-wide d step (fp32 in TMEM, 256 columns per group = all 512 columns), so the row softmax sees all its keys at once:
//     no online rescaling of O, one barrier round trip per group instead of three;
//   * P never touches shared memory: it is packed to 16 bit and written back over the S columns it came from (tcgen05.st),
//     and O += P V takes its A operand from TMEM (tcgen05.mma, A in tensor memory), V read MN-major from its [key][d] tile;
//     O then reuses the dead upper half of the S columns;
//   * the CLS token is rank-1 work on CUDA cores: as a KEY its score q_i . k_0 is one 88-long dot product per row thread and
//     its PV term p_i0 * v_0 is added in the epilogue; as a QUERY (one row against 257 keys) every row thread also forms the
//     score of ITS OWN key, q_0 . k_j (the K tile row it already sits next to), and two spare warps do that row's softmax and
//     its P V from the V tile in shared memory.  No padding to 384 rows, no second kernel.
// Softmax: single pass against the exact maximum of the row's first 32 patch scores and its CLS score; a later 32-key chunk
// whose sum exceeds 2^13 (a score more than ~8 above the reference, log2 domain) raises the reference and rescales what
// was written so far -- the usual online-softmax correction, applied to the 16-bit P in TMEM (warp-voted, rare).
#include "common.cuh"

namespace mrb {

constexpr int VT_ROWS = 128;                 // query rows per softmax group
constexpr int VT_KEYS = 256;                 // patch keys (tokens 1..256); token 0 = CLS
constexpr int VT_D1 = 64, VT_D2 = 32;        // head dim as a 64-wide SW128 atom + a 32-wide SW64 atom (d >= hd zero-filled by TMA)
constexpr int VT_HD = VT_D1 + VT_D2;
constexpr float VT_REDO = 8192.0f;

struct VitAttnParams {
  int frames, H, L, hd, dtype;
  float scale;
  const uint16_t* q; const uint16_t* k; const uint16_t* v;      // token 0, head 0 of frame 0
  long long q_bs, q_rs, k_bs, k_rs, v_bs, v_rs;                  // frame / token strides in elements
  uint16_t* o; long long o_bs, o_rs;
};

struct VitSmem {
  static constexpr int Q1 = VT_ROWS * 128, Q2 = VT_ROWS * 64;   // one group's Q tile: SW128 part, SW64 part
  static constexpr int K1 = VT_KEYS * 128, K2 = VT_KEYS * 64;   // K (or V) tile
  static constexpr int OFF_Q = 0;                                // [g][Q1 | Q2]
  static constexpr int OFF_K = 2 * (Q1 + Q2);
  static constexpr int OFF_V = OFF_K + K1 + K2;
  static constexpr int OFF_X = OFF_V + K1 + K2;                  // 3 x [q0 | k0 | v0] fp32, 96 each (triple buffered by item)
  static constexpr int X_ONE = 3 * VT_HD * 4;
  static constexpr int OFF_PROB = OFF_X + 3 * X_ONE;             // CLS query row: 2 x 257 scores -> probabilities (+ pad), by item parity
  static constexpr int PROB_ONE = 264 * 4;
  static constexpr int OFF_PART = OFF_PROB + 2 * PROB_ONE;       // 5 x 96 partial O rows
  static constexpr int OFF_RED = OFF_PART + 5 * VT_HD * 4;       // cross-warp max / sum
  static constexpr int OFF_BAR = OFF_RED + 64;
  static constexpr int NBAR = 3 + 3 + 2 + 2 + 2 + 2 + 3 + 1;     // q/k/v full, q/k/v empty, s_full, p_full, o_full, o_empty, x_full, c_full
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16 + 1024;
  static constexpr int THREADS = 12 * 32;                        // TMA, MMA, 2 x 4 softmax, 2 CLS-row warps
};

__device__ __forceinline__ void vt_tma_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ uint64_t vt_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}
constexpr uint32_t VT_SW128 = 2, VT_SW64 = 4;
__host__ __device__ constexpr uint32_t vt_idesc(int fmt, int M, int N, int b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]: A (16 bit, K-major: lane = row, 32-bit column = two consecutive k) read from tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ float vt_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 16-byte chunk c of row r of a K-major tile as TMA wrote it: 128-byte rows / SWIZZLE_128B, or 64-byte rows / SWIZZLE_64B
__device__ __forceinline__ const uint4* vt_chunk128(const uint8_t* tile, int r, int c) {
  return reinterpret_cast<const uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ const uint4* vt_chunk64(const uint8_t* tile, int r, int c) {
  return reinterpret_cast<const uint4*>(tile + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
}
// sum_e x[e] * w[e] over the 8 16-bit values of one chunk
__device__ __forceinline__ float vt_dot8(uint4 x, const float* w, int dt) {
  return unpack_lo(x.x, dt) * w[0] + unpack_hi(x.x, dt) * w[1] + unpack_lo(x.y, dt) * w[2] + unpack_hi(x.y, dt) * w[3] +
         unpack_lo(x.z, dt) * w[4] + unpack_hi(x.z, dt) * w[5] + unpack_lo(x.w, dt) * w[6] + unpack_hi(x.w, dt) * w[7];
}
__device__ __forceinline__ void vt_bar_rows() { asm volatile("bar.sync 1, 64;" ::: "memory"); }   // the two CLS-row warps

template <int DT>      // MRB_DT_F16 / MRB_DT_BF16 at compile time: the unpack / pack sequences are not predicated both ways
__global__ void __launch_bounds__(VitSmem::THREADS, 1)
attn_vit_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmQ2,
                const __grid_constant__ CUtensorMap tmK2, const __grid_constant__ CUtensorMap tmV2, const VitAttnParams p) {
  mrb::pdl_trigger();
  using S = VitSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* q_full = bars + 0;  uint64_t* k_full = bars + 1;  uint64_t* v_full = bars + 2;
  uint64_t* q_empty = bars + 3; uint64_t* k_empty = bars + 4; uint64_t* v_empty = bars + 5;
  uint64_t* s_full = bars + 6;  uint64_t* p_full = bars + 8;  uint64_t* o_full = bars + 10; uint64_t* o_empty = bars + 12;
  uint64_t* x_full = bars + 14;      // [3]
  uint64_t* c_full = bars + 17;      // the 256 patch-key scores of the CLS query row are in shared memory
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + S::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.frames * p.H;
  constexpr int dt = DT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmQ2); tma_prefetch_desc(&tmK2); tma_prefetch_desc(&tmV2);
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(v_full, 1);
    mbar_init(q_empty, 1 + 8);          // S MMAs retired + the 8 softmax warps read their Q rows (CLS-key score)
    mbar_init(k_empty, 1 + 8 + 2);      // S MMAs retired + the 8 softmax warps read their K rows (score of the CLS query) + the CLS-row
                                        // warps took those scores (keeps c_full at most one phase ahead of them)
    mbar_init(c_full, 8);
    mbar_init(v_empty, 1 + 2);          // PV MMAs retired + the 2 CLS-row warps read V
    for (int g = 0; g < 2; ++g) { mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 4); mbar_init(&o_full[g], 1); mbar_init(&o_empty[g], 4); }
    for (int x = 0; x < 3; ++x) mbar_init(&x_full[x], 2);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  mrb::pdl_wait();

  uint8_t* sQ = smem + S::OFF_Q;
  uint8_t* sK = smem + S::OFF_K;
  uint8_t* sV = smem + S::OFF_V;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const int b = item / p.H, h = item % p.H;
        const uint32_t ph = (n & 1) ^ 1;                       // parity of the PREVIOUS item's "empty" phase (first wait passes)
        mbar_wait(q_empty, ph);
        mbar_expect_tx(q_full, 2 * (S::Q1 + S::Q2));
        for (int g = 0; g < 2; ++g) {
          vt_tma_4d(sQ + g * (S::Q1 + S::Q2), &tmQ, q_full, 0, h, 1 + g * VT_ROWS, b);
          vt_tma_4d(sQ + g * (S::Q1 + S::Q2) + S::Q1, &tmQ2, q_full, VT_D1, h, 1 + g * VT_ROWS, b);
        }
        mbar_wait(k_empty, ph);
        mbar_expect_tx(k_full, S::K1 + S::K2);
        for (int g = 0; g < 2; ++g) {
          vt_tma_4d(sK + g * (VT_ROWS * 128), &tmK, k_full, 0, h, 1 + g * VT_ROWS, b);
          vt_tma_4d(sK + S::K1 + g * (VT_ROWS * 64), &tmK2, k_full, VT_D1, h, 1 + g * VT_ROWS, b);
        }
        mbar_wait(v_empty, ph);
        mbar_expect_tx(v_full, S::K1 + S::K2);
        for (int g = 0; g < 2; ++g) {
          vt_tma_4d(sV + g * (VT_ROWS * 128), &tmV, v_full, 0, h, 1 + g * VT_ROWS, b);
          vt_tma_4d(sV + S::K1 + g * (VT_ROWS * 64), &tmV2, v_full, VT_D1, h, 1 + g * VT_ROWS, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const int fmt = dt == MRB_DT_BF16 ? 1 : 0;
    const uint32_t id_s = vt_idesc(fmt, VT_ROWS, VT_KEYS, 0);      // S = Q K^T: 128 x 256, both K-major
    const uint32_t id_o1 = vt_idesc(fmt, VT_ROWS, VT_D1, 1);       // O[:, :64] += P V: B = V MN-major
    const uint32_t id_o2 = vt_idesc(fmt, VT_ROWS, VT_D2, 1);
    int n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const uint32_t ph = n & 1;
      mbar_wait(q_full, ph);
      mbar_wait(k_full, ph);
      for (int g = 0; g < 2; ++g) {
        if (n > 0) mbar_wait(&o_empty[g], ph ^ 1);                 // the previous item's O of this group has been read out
        tc_fence_after();
        if (lane == 0) {
          const uint32_t q_addr = smem_u32(sQ + g * (S::Q1 + S::Q2));
          const uint32_t k_addr = smem_u32(sK);
          const uint32_t d_s = tmem_base + g * 256;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(d_s, vt_desc(q_addr + k * 32, 16, 1024, VT_SW128), vt_desc(k_addr + k * 32, 16, 1024, VT_SW128), id_s, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_f16(d_s, vt_desc(q_addr + S::Q1 + k * 32, 16, 512, VT_SW64), vt_desc(k_addr + S::K1 + k * 32, 16, 512, VT_SW64), id_s, 1u);
          umma_commit(&s_full[g]);
        }
        __syncwarp();
      }
      if (lane == 0) { umma_commit(q_empty); umma_commit(k_empty); }
      __syncwarp();
      mbar_wait(v_full, ph);
      for (int g = 0; g < 2; ++g) {
        mbar_wait(&p_full[g], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t v_addr = smem_u32(sV);
          const uint32_t a_p = tmem_base + g * 256;                 // P: keys 2c, 2c+1 in column c
          const uint32_t d_o = tmem_base + g * 256 + 128;
#pragma unroll 4
          for (int k = 0; k < VT_KEYS / 16; ++k) {
            umma_f16_ts(d_o, a_p + k * 8, vt_desc(v_addr + k * 2048, VT_KEYS * 128, 1024, VT_SW128), id_o1, k > 0 ? 1u : 0u);
            umma_f16_ts(d_o + VT_D1, a_p + k * 8, vt_desc(v_addr + S::K1 + k * 1024, VT_KEYS * 64, 512, VT_SW64), id_o2, k > 0 ? 1u : 0u);
          }
          umma_commit(&o_full[g]);
        }
        __syncwarp();
      }
      if (lane == 0) umma_commit(v_empty);
      __syncwarp();
    }
  } else if (warp < 10) {
    // ===================== softmax groups: one thread per patch query row =====================
    const int g = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                                // row of the group's 128 x 256 S tile
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + g * 256;
    const float sl2 = p.scale * 1.4426950408889634f;
    const uint8_t* q1 = sQ + g * (S::Q1 + S::Q2);
    const uint8_t* q2 = q1 + S::Q1;
    int n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const int b = item / p.H, h = item % p.H;
      const uint32_t ph = n & 1;
      const float* xk0 = reinterpret_cast<const float*>(smem + S::OFF_X + (n % 3) * S::X_ONE) + VT_HD;
      const float* xv0 = xk0 + VT_HD;
      // ---- the two rank-1 pieces on CUDA cores: the score of the CLS KEY for this thread's query row, q_r . k_0, and the score
      //      of this thread's KEY (tile row g * 128 + r) for the CLS QUERY, q_0 . k_j -- 16-byte chunks of the swizzled tiles
      //      against fp32 vectors broadcast from shared memory, four independent partial sums each
      mbar_wait(&x_full[n % 3], (n / 3) & 1);
      mbar_wait(q_full, ph);
      mbar_wait(k_full, ph);
      const float* xq0 = xk0 - VT_HD;
      const uint8_t* k1 = sK;
      const uint8_t* k2 = sK + S::K1;
      const int jr = g * VT_ROWS + r;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll
      for (int c = 0; c < 8; c += 4) {
        a0 += vt_dot8(*vt_chunk128(q1, r, c), xk0 + c * 8, dt);
        a1 += vt_dot8(*vt_chunk128(q1, r, c + 1), xk0 + c * 8 + 8, dt);
        a2 += vt_dot8(*vt_chunk128(q1, r, c + 2), xk0 + c * 8 + 16, dt);
        a3 += vt_dot8(*vt_chunk128(q1, r, c + 3), xk0 + c * 8 + 24, dt);
        b0 += vt_dot8(*vt_chunk128(k1, jr, c), xq0 + c * 8, dt);
        b1 += vt_dot8(*vt_chunk128(k1, jr, c + 1), xq0 + c * 8 + 8, dt);
        b2 += vt_dot8(*vt_chunk128(k1, jr, c + 2), xq0 + c * 8 + 16, dt);
        b3 += vt_dot8(*vt_chunk128(k1, jr, c + 3), xq0 + c * 8 + 24, dt);
      }
      a0 += vt_dot8(*vt_chunk64(q2, r, 0), xk0 + VT_D1, dt);
      a1 += vt_dot8(*vt_chunk64(q2, r, 1), xk0 + VT_D1 + 8, dt);
      a2 += vt_dot8(*vt_chunk64(q2, r, 2), xk0 + VT_D1 + 16, dt);
      a3 += vt_dot8(*vt_chunk64(q2, r, 3), xk0 + VT_D1 + 24, dt);
      b0 += vt_dot8(*vt_chunk64(k2, jr, 0), xq0 + VT_D1, dt);
      b1 += vt_dot8(*vt_chunk64(k2, jr, 1), xq0 + VT_D1 + 8, dt);
      b2 += vt_dot8(*vt_chunk64(k2, jr, 2), xq0 + VT_D1 + 16, dt);
      b3 += vt_dot8(*vt_chunk64(k2, jr, 3), xq0 + VT_D1 + 24, dt);
      a0 += a2; a1 += a3; b0 += b2; b1 += b3;
      const float t_cls = (a0 + a1) * sl2;
      reinterpret_cast<float*>(smem + S::OFF_PROB + (n & 1) * S::PROB_ONE)[1 + jr] = (b0 + b1) * sl2;
      __syncwarp();
      if (lane == 0) { mbar_arrive(q_empty); mbar_arrive(k_empty); mbar_arrive(c_full); }

      mbar_wait(&s_full[g], ph);
      tc_fence_after();
      float m_ref = t_cls, l = 0.f;
      for (int c = 0; c < VT_KEYS; c += 32) {
        uint32_t sv[32];
        tmem_ld_32x32b_x32(lane_base + c, sv);
        tmem_ld_wait();
        if (c == 0) {                                              // reference = exact maximum of the first 32 patch scores and the CLS score
          float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; e += 2) { mx0 = fmaxf(mx0, __uint_as_float(sv[e])); mx1 = fmaxf(mx1, __uint_as_float(sv[e + 1])); }
          m_ref = fmaxf(m_ref, fmaxf(mx0, mx1) * sl2);             // sl2 > 0
        }
        float t[32];
        float cs0 = 0.f, cs1 = 0.f;
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          t[e] = fmaf(__uint_as_float(sv[e]), sl2, -m_ref);
          t[e + 1] = fmaf(__uint_as_float(sv[e + 1]), sl2, -m_ref);
          const float p0 = vt_ex2(t[e]), p1 = vt_ex2(t[e + 1]);
          cs0 += p0; cs1 += p1;
          pk[e >> 1] = pack2(p0, p1, dt);
        }
        float csum = cs0 + cs1;
        if (__any_sync(0xffffffffu, !(csum <= VT_REDO))) {
          // rare: some score of this chunk is far above the reference.  Raise the reference of the rows concerned, rescale their
          // running sum and the P columns written so far, redo the chunk.  (warp-uniform: tcgen05.ld / st are collective)
          float cm = t[0];
#pragma unroll
          for (int e = 1; e < 32; ++e) cm = fmaxf(cm, t[e]);
          cm = fmaxf(cm, 0.f);
          const float corr = vt_ex2(-cm);
          m_ref += cm;
          l *= corr;
          tmem_st_wait();                                           // the P columns written so far are read back below
          for (int c2 = 0; c2 < c; c2 += 32) {
            uint32_t w[16];
            tmem_ld_32x32b_x16(lane_base + (c2 >> 1), w);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) w[e] = pack2(unpack_lo(w[e], dt) * corr, unpack_hi(w[e], dt) * corr, dt);
            tmem_st_32x32b_x16(lane_base + (c2 >> 1), w);
          }
          cs0 = 0.f; cs1 = 0.f;
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float p0 = vt_ex2(t[e] - cm), p1 = vt_ex2(t[e + 1] - cm);
            cs0 += p0; cs1 += p1;
            pk[e >> 1] = pack2(p0, p1, dt);
          }
          csum = cs0 + cs1;
        }
        l += csum;
        tmem_st_32x32b_x16(lane_base + (c >> 1), pk);              // P over the S columns it came from (always behind the reads)
      }
      const float p_cls = vt_ex2(t_cls - m_ref);                    // <= 1: the reference is >= the CLS score
      l += p_cls;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);

      // ---- epilogue: O (TMEM) + p_cls * v_0, / l
      mbar_wait(&o_full[g], ph);
      tc_fence_after();
      const float inv = 1.f / l;
      const float pc = p_cls * inv;
      uint16_t* orow = p.o + b * p.o_bs + static_cast<long long>(1 + g * VT_ROWS + r) * p.o_rs + static_cast<long long>(h) * p.hd;
#pragma unroll
      for (int c0 = 0; c0 < VT_HD; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(lane_base + 128 + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          if (c0 + c < p.hd) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = fmaf(__uint_as_float(v[c + e]), inv, pc * xv0[c0 + c + e]);
            *reinterpret_cast<uint4*>(orow + c0 + c) = make_uint4(pack2(o8[0], o8[1], dt), pack2(o8[2], o8[3], dt),
                                                                  pack2(o8[4], o8[5], dt), pack2(o8[6], o8[7], dt));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[g]);
    }
  } else {
    // ===================== CLS query row: 1 x 257 attention on CUDA cores from the tiles in shared memory =====================
    const int t = threadIdx.x - 10 * 32;                            // 0..63
    const int rw = warp - 10;
    float* part = reinterpret_cast<float*>(smem + S::OFF_PART);
    float* red = reinterpret_cast<float*>(smem + S::OFF_RED);
    const float sl2 = p.scale * 1.4426950408889634f;
    auto load_x = [&](int item, float* xb) {                       // q_0, k_0, v_0 of the item -> fp32 (d >= hd: zero)
      const int b = item / p.H, h = item % p.H;
      for (int i = t; i < 3 * VT_HD; i += 64) {
        const int which = i / VT_HD, d = i % VT_HD;
        const uint16_t* src = which == 0 ? p.q + b * p.q_bs : which == 1 ? p.k + b * p.k_bs : p.v + b * p.v_bs;
        float val = 0.f;
        if (d < p.hd) {
          const uint16_t raw = src[static_cast<long long>(h) * p.hd + d];
          val = dt == MRB_DT_F16 ? __half2float(__ushort_as_half(raw)) : __uint_as_float(static_cast<uint32_t>(raw) << 16);
        }
        xb[i] = val;
      }
    };
    if (blockIdx.x < n_items) {
      load_x(blockIdx.x, reinterpret_cast<float*>(smem + S::OFF_X));
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_full[0]);
    }
    int n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const int b = item / p.H, h = item % p.H;
      const uint32_t ph = n & 1;
      const float* xq0 = reinterpret_cast<const float*>(smem + S::OFF_X + (n % 3) * S::X_ONE);
      const float* xk0 = xq0 + VT_HD;
      const float* xv0 = xk0 + VT_HD;
      // the next item's q_0 / k_0 / v_0 (its buffer was last read two items ago, whose epilogue finished before this item's S MMA)
      if (item + gridDim.x < n_items) {
        load_x(item + gridDim.x, reinterpret_cast<float*>(smem + S::OFF_X + ((n + 1) % 3) * S::X_ONE));
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_full[(n + 1) % 3]);
      }
      vt_bar_rows();                                               // this item's x buffer is complete for both warps (item 0: written above)
      // ---- scores of the 256 patch keys come from the softmax threads (one key each); thread t takes t, t + 64, t + 128, t + 192
      float* prob = reinterpret_cast<float*>(smem + S::OFF_PROB + (n & 1) * S::PROB_ONE);     // [0] = CLS key, [1 + j] = patch key j
      float s0 = xq0[lane] * xk0[lane] + xq0[lane + 32] * xk0[lane + 32] + xq0[lane + 64] * xk0[lane + 64];
      s0 = warp_sum(s0) * sl2;                                     // CLS query against the CLS key
      mbar_wait(c_full, ph);
      __syncwarp();
      if (lane == 0) mbar_arrive(k_empty);
      float sc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) sc[i] = prob[1 + t + 64 * i];
      float mx = fmaxf(fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3])), s0);
      mx = warp_max(mx);
      if (lane == 0) red[rw] = mx;
      vt_bar_rows();
      mx = fmaxf(red[0], red[1]);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float e = vt_ex2(sc[i] - mx);
        prob[1 + t + 64 * i] = e;
        sum += e;
      }
      const float e0 = vt_ex2(s0 - mx);
      if (t == 0) sum += e0;
      sum = warp_sum(sum);
      if (lane == 0) red[2 + rw] = sum;
      vt_bar_rows();
      const float inv = 1.f / (red[2] + red[3]);
      // ---- O_0 = sum_j p_j v_j: thread = (8-wide d chunk dc, key residue kp mod 5); 60 threads active
      mbar_wait(v_full, ph);
      const int dc = t % 12, kp = t / 12;
      if (kp < 5) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const uint8_t* vt = dc < 8 ? sV : sV + S::K1;
#pragma unroll 4
        for (int j = kp; j < VT_KEYS; j += 5) {
          const uint4 x = dc < 8 ? *vt_chunk128(vt, j, dc) : *vt_chunk64(vt, j, dc - 8);
          const float pj = prob[1 + j];
          acc[0] = fmaf(pj, unpack_lo(x.x, dt), acc[0]); acc[1] = fmaf(pj, unpack_hi(x.x, dt), acc[1]);
          acc[2] = fmaf(pj, unpack_lo(x.y, dt), acc[2]); acc[3] = fmaf(pj, unpack_hi(x.y, dt), acc[3]);
          acc[4] = fmaf(pj, unpack_lo(x.z, dt), acc[4]); acc[5] = fmaf(pj, unpack_hi(x.z, dt), acc[5]);
          acc[6] = fmaf(pj, unpack_lo(x.w, dt), acc[6]); acc[7] = fmaf(pj, unpack_hi(x.w, dt), acc[7]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) part[kp * VT_HD + dc * 8 + e] = acc[e];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(v_empty);
      vt_bar_rows();
      uint16_t* orow = p.o + b * p.o_bs + static_cast<long long>(h) * p.hd;       // token 0
      for (int d = t; d < p.hd; d += 64) {
        float o = e0 * xv0[d];
#pragma unroll
        for (int q5 = 0; q5 < 5; ++q5) o += part[q5 * VT_HD + d];
        o *= inv;
        orow[d] = dt == MRB_DT_F16 ? __half_as_ushort(__float2half_rn(o)) : __bfloat16_as_ushort(__float2bfloat16_rn(o));
      }
      vt_bar_rows();                                               // prob / part / red are rewritten by the next item
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- host
typedef CUresult (*VtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static VtEncodeFn vt_encode_fn() {
  static VtEncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<VtEncodeFn>(ptr);
  }
  return fn;
}
// 4-D view (d, head, token, frame) of a [frames, L, heads * hd] 16-bit tensor; box = box_d x 1 x 128 x 1
static int vt_tmap(CUtensorMap* map, const void* base, int dtype, int hd, int heads, int L, int frames, long long rs, long long bs,
                   int box_d, bool sw64) {
  VtEncodeFn fn = vt_encode_fn();
  if (!fn) return MRB_ERR_CUDA;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(hd), static_cast<cuuint64_t>(heads), static_cast<cuuint64_t>(L),
                        static_cast<cuuint64_t>(frames)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(hd) * 2, static_cast<cuuint64_t>(rs) * 2, static_cast<cuuint64_t>(bs) * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_d), 1, VT_ROWS, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, dtype == MRB_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                  const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MRB_OK : MRB_ERR_CUDA;
}

}  // namespace mrb

using namespace mrb;

// Self-attention of the EVA ViT over every (frame, head): out = softmax(scale * Q K^T) V with L = 257 tokens (CLS + 256 patches)
// and 64 < hd <= 96, hd % 8 == 0 (the ViT-g has hd 88); q / k / v / out point at token 0, head 0 of frame 0, strides in elements.
// Replaces Attention.forward of lavis/models/eva_vit.py:128-145 between the qkv and proj Linears (no relative position bias:
// use_rel_pos_bias is off at eva_vit.py:416-428).
extern "C" int mrb_attention_vit(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                 const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                                 int frames, int H, int L, int hd, int dtype, float scale, void* stream) {
  if (frames <= 0 || H <= 0) return MRB_OK;
  if (dtype != MRB_DT_F16 && dtype != MRB_DT_BF16) return MRB_ERR_ARG;
  if (L != VT_KEYS + 1 || !(hd > VT_D1 && hd <= VT_HD && (hd & 7) == 0)) return MRB_ERR_UNSUPPORTED;
  if ((q_rs | k_rs | v_rs | o_rs | q_bs | k_bs | v_bs | o_bs) & 7) return MRB_ERR_ARG;
  if (!(scale > 0.f)) return MRB_ERR_ARG;
  CUtensorMap maps[6];
  int rc = vt_tmap(&maps[0], q, dtype, hd, H, L, frames, q_rs, q_bs, VT_D1, false);
  if (!rc) rc = vt_tmap(&maps[1], k, dtype, hd, H, L, frames, k_rs, k_bs, VT_D1, false);
  if (!rc) rc = vt_tmap(&maps[2], v, dtype, hd, H, L, frames, v_rs, v_bs, VT_D1, false);
  if (!rc) rc = vt_tmap(&maps[3], q, dtype, hd, H, L, frames, q_rs, q_bs, VT_D2, true);
  if (!rc) rc = vt_tmap(&maps[4], k, dtype, hd, H, L, frames, k_rs, k_bs, VT_D2, true);
  if (!rc) rc = vt_tmap(&maps[5], v, dtype, hd, H, L, frames, v_rs, v_bs, VT_D2, true);
  if (rc) return rc;
  VitAttnParams p{};
  p.frames = frames; p.H = H; p.L = L; p.hd = hd; p.dtype = dtype; p.scale = scale;
  p.q = static_cast<const uint16_t*>(q); p.k = static_cast<const uint16_t*>(k); p.v = static_cast<const uint16_t*>(v);
  p.q_bs = q_bs; p.q_rs = q_rs; p.k_bs = k_bs; p.k_rs = k_rs; p.v_bs = v_bs; p.v_rs = v_rs;
  p.o = static_cast<uint16_t*>(o); p.o_bs = o_bs; p.o_rs = o_rs;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(attn_vit_kernel<MRB_DT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, VitSmem::TOTAL);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_vit_kernel<MRB_DT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, VitSmem::TOTAL);
    if (e != cudaSuccess) { sms = 0; return mrb_set_error(e); }
  }
  const int items = frames * H;
  const int grid = items < sms ? items : sms;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == MRB_DT_F16)
    MRB_LAUNCH((attn_vit_kernel<MRB_DT_F16>), grid, VitSmem::THREADS, VitSmem::TOTAL, s, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
  else
    MRB_LAUNCH((attn_vit_kernel<MRB_DT_BF16>), grid, VitSmem::THREADS, VitSmem::TOTAL, s, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
