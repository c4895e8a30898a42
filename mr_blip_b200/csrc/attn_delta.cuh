// delta_i of the flash-attention backward, computed EXACTLY as sum_j P_ij * dP_ij (P, dP recomputed from q, k, v, dO and lse) for
// the few-query-rows shapes (Lq <= 64: the T5 decoder's cross-attention, 16 target rows against ~2000 encoder keys, and its self-
// attention), instead of the usual rowsum(dO * O).
//
// Why: dS_ij = P_ij (dP_ij - delta_i) must sum to zero over j.  rowsum(dO * O) equals sum_j P_ij dP_ij only for the exact O; the
// forward stores O in 16 bit, and the rounding error of O_i enters delta_i once and then dS_ij of EVERY key j of the row with the
// same sign: dQ_i = sum_j dS_ij K_j picks up -err_i * (P-weighted mean of K), which does not average out, while the true dQ_i is a
// sum of ~L incoherent deviations and shrinks like 1 / sqrt(L).  Measured at full depth (profiles/full_depth_parity_r02a.json): the
// LoRA gradients of the decoder's cross-attention q / k were 3-8 x further from the fp32 oracle than the reference's own autocast
// regime (0.15 vs 0.018 rel. error at L_enc 2033) while every other gradient was closer than it.  Eager PyTorch does not have the
// problem because its softmax backward forms sum_j P_ij dP_ij from the same dP it multiplies with.  With many query rows per key the
// same error is incoherent over i in dK and was not visible; the long-Lq shapes keep the one-pass rowsum(dO * O).
#pragma once
#include "common.cuh"
#include "dropmask.cuh"

namespace mrb {

constexpr int DELTA_EXACT_MAX_LQ = 64;

struct DeltaExactParams {
  const uint16_t* q; const uint16_t* k; const uint16_t* v; const uint16_t* dout;
  long long q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, do_bs, do_rs;      // batch / row strides in elements; head h at column h * 64
  int B, H, Lq, Lk, dtype;
  float scale;
  const float* bias; int bias_len, bias_zero;                      // [H, bias_len], index (j - i_abs) + bias_zero
  const int* kmask;                                                // [B, Lk] or null
  int causal, q_pos0;
  const float* lse;                                                // [B, H, Lq] (natural log)
  float* delta;                                                    // [B, H, Lq]
  const uint32_t* drop_seed; uint32_t drop_site, drop_thr; float drop_scale;   // drop_seed null: no dropout of the probabilities
};

__device__ __forceinline__ void delta_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1, int dt) {
#ifdef MRB_HOST_SHIM      // tests/cuda_host_shim: the CPU suite runs this source with an emulated warp
  shim::mma_m16n8k16(c, a, b0, b1, dt);
#else
  if (dt == MRB_DT_F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
#endif
}

// Block = (head, batch), 8 warps; a warp takes every 8th 8-key tile and, per 16-row query tile, forms
// S = Q K^T and dP = dO V^T with mma.sync m16n8k16 (hd = 64 = 4 k-steps).  The k slots are permuted so that a lane's fragment of a
// K / V / Q / dO row is the 32 consecutive bytes [32 t, 32 t + 32) of that row (two 16-byte loads; a dot product does not care in
// which order d runs): K and V of a head are read from L2 exactly once per query tile.  (A first version with one block per query
// row re-read them Lq times: 1 GB of L2 traffic, 106 us per decoder layer.)  The warps' partial sums meet in shared memory in a
// fixed order: no atomics, the result is bit-reproducible from run to run (graph replay vs eager launches compare exactly).
// DELTA_WARPS = 16 for the decoder's cross-attention (~2000 keys: 16 key tiles per warp instead of 32, the launch is a chain of
// dependent L2 round trips on 128 blocks -- 51.6 us per decoder layer with 8 warps, profiles/launch_summary_r02f.csv), 8 otherwise.
template <int DELTA_WARPS>
static __global__ void __launch_bounds__(DELTA_WARPS * 32) attn_delta_exact_kernel(const DeltaExactParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[DELTA_WARPS][16];
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int key_lo = 0, key_hi = p.Lk;
  const float LOG2E = 1.4426950408889634f;
  const float sl2 = p.scale * LOG2E;
  const float* bhead = p.bias ? p.bias + static_cast<long long>(h) * p.bias_len : nullptr;
  const int* mrow = p.kmask ? p.kmask + static_cast<long long>(b) * p.Lk : nullptr;
  const long long row0 = (static_cast<long long>(b) * p.H + h) * p.Lq;
  uint32_t dkey = 0, dng = 0;
  if (p.drop_seed) { dkey = drop_key(*p.drop_seed, p.drop_site); dng = drop_groups(static_cast<uint32_t>(p.Lk)); }
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int i0 = 0; i0 < p.Lq; i0 += 16) {
    // A fragments of this query tile: rows i0 + g and i0 + g + 8, bytes [32 t, 32 t + 32) of the q and dO rows
    const int ia = i0 + g, ib = i0 + g + 8;
    const bool oka = ia < p.Lq, okb = ib < p.Lq;
    uint4 qa[2], qb[2], da[2], db[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      qa[c] = oka ? *reinterpret_cast<const uint4*>(p.q + b * p.q_bs + static_cast<long long>(ia) * p.q_rs + h * 64 + 16 * t + 8 * c) : zero4;
      qb[c] = okb ? *reinterpret_cast<const uint4*>(p.q + b * p.q_bs + static_cast<long long>(ib) * p.q_rs + h * 64 + 16 * t + 8 * c) : zero4;
      da[c] = oka ? *reinterpret_cast<const uint4*>(p.dout + b * p.do_bs + static_cast<long long>(ia) * p.do_rs + h * 64 + 16 * t + 8 * c) : zero4;
      db[c] = okb ? *reinterpret_cast<const uint4*>(p.dout + b * p.do_bs + static_cast<long long>(ib) * p.do_rs + h * 64 + 16 * t + 8 * c) : zero4;
    }
    const float lse_a = oka ? p.lse[row0 + ia] * LOG2E : 0.f, lse_b = okb ? p.lse[row0 + ib] * LOG2E : 0.f;
    float acc_a = 0.f, acc_b = 0.f;
#pragma unroll 2
    for (int j0 = key_lo + 8 * warp; j0 < key_hi; j0 += 8 * DELTA_WARPS) {
      // B fragments: key j0 + g, bytes [32 t, 32 t + 32) of its K and V rows
      const int jk = j0 + g;
      const bool okk = jk < key_hi;
      uint4 kf[2], vf[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        kf[c] = okk ? *reinterpret_cast<const uint4*>(p.k + b * p.k_bs + static_cast<long long>(jk) * p.k_rs + h * 64 + 16 * t + 8 * c) : zero4;
        vf[c] = okk ? *reinterpret_cast<const uint4*>(p.v + b * p.v_bs + static_cast<long long>(jk) * p.v_rs + h * 64 + 16 * t + 8 * c) : zero4;
      }
      float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
      // k-step ks: slots (2t, 2t+1) = the lane's values 4 ks, 4 ks + 1; slots (2t+8, 2t+9) = values 4 ks + 2, 4 ks + 3
      {
        const uint32_t a0[4] = {qa[0].x, qb[0].x, qa[0].y, qb[0].y}; delta_mma(s, a0, kf[0].x, kf[0].y, p.dtype);
        const uint32_t a1[4] = {qa[0].z, qb[0].z, qa[0].w, qb[0].w}; delta_mma(s, a1, kf[0].z, kf[0].w, p.dtype);
        const uint32_t a2[4] = {qa[1].x, qb[1].x, qa[1].y, qb[1].y}; delta_mma(s, a2, kf[1].x, kf[1].y, p.dtype);
        const uint32_t a3[4] = {qa[1].z, qb[1].z, qa[1].w, qb[1].w}; delta_mma(s, a3, kf[1].z, kf[1].w, p.dtype);
        const uint32_t e0[4] = {da[0].x, db[0].x, da[0].y, db[0].y}; delta_mma(dp, e0, vf[0].x, vf[0].y, p.dtype);
        const uint32_t e1[4] = {da[0].z, db[0].z, da[0].w, db[0].w}; delta_mma(dp, e1, vf[0].z, vf[0].w, p.dtype);
        const uint32_t e2[4] = {da[1].x, db[1].x, da[1].y, db[1].y}; delta_mma(dp, e2, vf[1].x, vf[1].y, p.dtype);
        const uint32_t e3[4] = {da[1].z, db[1].z, da[1].w, db[1].w}; delta_mma(dp, e3, vf[1].z, vf[1].w, p.dtype);
      }
      // accumulators: s[0], s[1] = (row ia, keys j0 + 2t, + 1); s[2], s[3] = (row ib, same keys)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = (e < 2) ? ia : ib, j = j0 + 2 * t + (e & 1);
        const bool rok = (e < 2) ? oka : okb;
        if (!rok || j >= key_hi || (mrow && mrow[j] == 0) || (p.causal && j > i + p.q_pos0)) continue;
        float x = s[e] * sl2 - ((e < 2) ? lse_a : lse_b);
        if (bhead) {
          const int idx = j - (i + p.q_pos0) + p.bias_zero;
          if (idx >= 0 && idx < p.bias_len) x += bhead[idx] * LOG2E;
        }
        float d1 = dp[e];
        if (p.drop_seed) {     // delta of the DROPPED output: sum_j P_ij * (s m_ij dP_ij)
          const uint32_t w = drop_word(dkey, static_cast<uint32_t>(row0 + i) * dng, static_cast<uint32_t>(j) >> 2);
          d1 = drop_keep(w, j, p.drop_thr) ? d1 * p.drop_scale : 0.f;
        }
        const float pr = exp2f(x) * d1;
        if (e < 2) acc_a += pr; else acc_b += pr;
      }
    }
    // the four lanes of a row group hold partial sums of the same two rows
    acc_a += __shfl_xor_sync(0xffffffffu, acc_a, 1); acc_a += __shfl_xor_sync(0xffffffffu, acc_a, 2);
    acc_b += __shfl_xor_sync(0xffffffffu, acc_b, 1); acc_b += __shfl_xor_sync(0xffffffffu, acc_b, 2);
    __syncthreads();                                             // the previous query tile's sums have been read
    if (t == 0) { red[warp][g] = acc_a; red[warp][g + 8] = acc_b; }
    __syncthreads();
    if (threadIdx.x < 16 && i0 + threadIdx.x < p.Lq) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < DELTA_WARPS; ++w) sum += red[w][threadIdx.x];
      p.delta[row0 + i0 + threadIdx.x] = sum;
    }
  }
}

static inline int launch_delta_exact(const DeltaExactParams& d, cudaStream_t s) {
  static int wide = -1;                   // MRB_DELTA_WIDE=0: 8 warps for every key count (A/B measurements)
  if (wide < 0) { const char* e = getenv("MRB_DELTA_WIDE"); wide = (e && e[0] == '0') ? 0 : 1; }
  if (wide && d.Lk >= 512) MRB_LAUNCH((attn_delta_exact_kernel<16>), dim3(d.H, d.B), 16 * 32, 0, s, d);
  else MRB_LAUNCH((attn_delta_exact_kernel<8>), dim3(d.H, d.B), 8 * 32, 0, s, d);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

}  // namespace mrb
