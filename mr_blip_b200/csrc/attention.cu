// Fused flash-style attention (softmax(scale*Q K^T + bias + mask) V) for the four attention shapes of
// the path: EVA ViT (eva_vit.py:118-148, 257 tokens, hd 88), Q-Former self/cross (Qformer.py:169-275,
// 32x32 / 32x257, hd 64), T5 encoder/decoder self + cross (modeling_t5.py:474-620, no 1/sqrt(d),
// bucketed relative bias, padding / causal mask), plus the T5 backward (dQ and dK/dV kernels).
// Scores never touch HBM.  This first version drives the tensor cores through mma.sync m16n8k16
// (HMMA); the tcgen05 port of the long-sequence T5 case is tracked in DESIGN.md.
#include <type_traits>
#include "common.cuh"
#include "dropmask.cuh"
#include "attn_delta.cuh"

namespace mrb {

struct AttnParams {
  const void* q; const void* k; const void* v; void* o;
  long long q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;   // batch / row strides (elements); head h at column h*hd
  int B, H, Lq, Lk, hd;
  float scale;
  const float* bias;        // [H, bias_len] indexed by (j - i_abs) + bias_zero, or null
  int bias_len, bias_zero;
  const int* kmask;         // [B, Lk] 1 = attend, or null
  int kv_div;               // K/V (and kmask) batch index = b / kv_div  (beams sharing one encoder output)
  int causal;               // key j allowed iff j <= i + q_pos0
  int q_pos0;               // absolute position of query row 0
  float* lse;               // [B, H, Lq] or null
  // backward
  const void* dout; long long do_bs, do_rs;
  const float* delta;       // [B, H, Lq] rowsum(dO * O)
  void* dq; void* dk; void* dv;   // same layout/strides as q/k/v
};
// DROP instantiations only: train-mode dropout of the attention probabilities (modeling_t5.py:600, Qformer.py:258); masks of
// dropmask.cuh with row = (b H + h) Lq + i, column = key j.  O = (m s P) V with s = 1 / (1 - p) folded into the final 1 / l;
// backward: dV = s (m P)^T dO, dS = P * (s m dP - delta) * scale.
struct AttnDropParams : AttnParams {
  const uint32_t* drop_seed; uint32_t drop_site, drop_thr; float drop_scale;
};
template <bool DROP> struct AttnParamsOf { typedef AttnParams type; };
template <> struct AttnParamsOf<true> { typedef AttnDropParams type; };

constexpr int BQ = 64, BKV = 64, NTHREADS = 128;

template <typename T> struct MmaType;
template <> struct MmaType<__half> {
  static __device__ __forceinline__ void mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
#ifdef MRB_HOST_SHIM      // tests/cuda_host_shim: the CPU suite runs these kernels with an emulated warp (no effect on the CUDA build)
    shim::mma_m16n8k16(c, a, b0, b1, MRB_DT_F16);
#else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
#endif
  }
  static __device__ __forceinline__ uint32_t pack(float x, float y) { __half2 h = __floats2half2_rn(x, y); return *reinterpret_cast<uint32_t*>(&h); }
};
template <> struct MmaType<__nv_bfloat16> {
  static __device__ __forceinline__ void mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
#ifdef MRB_HOST_SHIM
    shim::mma_m16n8k16(c, a, b0, b1, MRB_DT_BF16);
#else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
#endif
  }
  static __device__ __forceinline__ uint32_t pack(float x, float y) { __nv_bfloat162 h = __floats2bfloat162_rn(x, y); return *reinterpret_cast<uint32_t*>(&h); }
};

__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t addr) {
#ifdef MRB_HOST_SHIM
  shim::ldsm_x4(r, addr, false);
#else
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
#endif
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, uint32_t addr) {
#ifdef MRB_HOST_SHIM
  shim::ldsm_x4(r, addr, true);
#else
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
#endif
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
#ifdef MRB_HOST_SHIM
  shim::cp_async16(dst, src, sz);
#else
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
#endif
}
#ifdef MRB_HOST_SHIM      // the emulated copies complete at once
__device__ __forceinline__ void cp_async_commit() {}
template <int N> __device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// Load a [ROWS x HD] tile (row stride HD+8 elements in smem) of a [L, *] matrix; rows >= L and columns >= hd zero-filled.
template <typename T, int HD, int ROWS>
__device__ __forceinline__ void load_tile(uint32_t smem, const T* g, long long rs, int row0, int L, int hd) {
  constexpr int CH = HD / 8;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < ROWS * CH; i += NTHREADS) {
    const int r = i / CH, c = i - r * CH;
    const bool ok = (row0 + r < L) && (c * 8 < hd);
    const T* src = ok ? g + static_cast<long long>(row0 + r) * rs + c * 8 : g;
    cp_async16(smem + (r * (HD + 8) + c * 8) * 2, src, ok);
  }
}

// score post-processing shared by forward and backward: scale, relative bias, masks
struct ScoreCtx {
  float scale; const float* bias; int bias_zero; const int* kmask; int causal, q_pos0, Lk, Lq;
  __device__ __forceinline__ float apply(float s, int i, int j) const {
    if (j >= Lk) return -INFINITY;
    i = min(i, Lq - 1);                      // rows past Lq are never stored; keep their bias index in range
    s *= scale;
    if (bias) s += __ldg(bias + (j - (i + q_pos0)) + bias_zero);
    if (kmask && __ldg(kmask + j) == 0) return -INFINITY;
    if (causal && j > i + q_pos0) return -INFINITY;
    return s;
  }
};

template <typename T, int HD, bool DROP = false>
__global__ void __launch_bounds__(NTHREADS) attn_fwd_kernel(const typename AttnParamsOf<DROP>::type p) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  constexpr int LDS = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  T* sQ = reinterpret_cast<T*>(smem_attn);
  T* sK = sQ + BQ * LDS;
  T* sV = sK + 2 * BKV * LDS;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const T* gq = static_cast<const T*>(p.q) + b * p.q_bs + static_cast<long long>(h) * p.hd;
  const T* gk = static_cast<const T*>(p.k) + (b / p.kv_div) * p.k_bs + static_cast<long long>(h) * p.hd;
  const T* gv = static_cast<const T*>(p.v) + (b / p.kv_div) * p.v_bs + static_cast<long long>(h) * p.hd;
  ScoreCtx sc{p.scale, p.bias ? p.bias + static_cast<long long>(h) * p.bias_len : nullptr, p.bias_zero,
              p.kmask ? p.kmask + static_cast<long long>(b / p.kv_div) * p.Lk : nullptr, p.causal, p.q_pos0, p.Lk, p.Lq};

  int n_kv = (p.Lk + BKV - 1) / BKV;
  if (p.causal) n_kv = min(n_kv, (q0 + BQ - 1 + p.q_pos0) / BKV + 1);

  load_tile<T, HD, BQ>(smem_u32(sQ), gq, p.q_rs, q0, p.Lq, p.hd);
  load_tile<T, HD, BKV>(smem_u32(sK), gk, p.k_rs, 0, p.Lk, p.hd);
  load_tile<T, HD, BKV>(smem_u32(sV), gv, p.v_rs, 0, p.Lk, p.hd);
  cp_async_commit();

  float o_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[HD / 16][4];
  const float LOG2E = 1.4426950408889634f;
  uint32_t dkey = 0, drow[2] = {0u, 0u}, dthr = 0;
  float dscale = 1.f;
  if constexpr (DROP) {
    dkey = drop_key(*p.drop_seed, p.drop_site);
    dthr = p.drop_thr; dscale = p.drop_scale;
#pragma unroll
    for (int r = 0; r < 2; ++r)
      drow[r] = (static_cast<uint32_t>(b * p.H + h) * static_cast<uint32_t>(p.Lq) +
                 static_cast<uint32_t>(min(q0 + warp * 16 + g + 8 * r, p.Lq - 1))) * drop_groups(static_cast<uint32_t>(p.Lk));
  }

  for (int kv = 0; kv < n_kv; ++kv) {
    const int buf = kv & 1;
    if (kv + 1 < n_kv) {
      load_tile<T, HD, BKV>(smem_u32(sK + (buf ^ 1) * BKV * LDS), gk, p.k_rs, (kv + 1) * BKV, p.Lk, p.hd);
      load_tile<T, HD, BKV>(smem_u32(sV + (buf ^ 1) * BKV * LDS), gv, p.v_rs, (kv + 1) * BKV, p.Lk, p.hd);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kv == 0) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk)
        ldsm_x4(qf[kk], smem_u32(sQ + (warp * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8));
    }
    const T* cK = sK + buf * BKV * LDS;
    const T* cV = sV + buf * BKV * LDS;
    // S = Q K^T  (16 x 64 per warp)
    float s[BKV / 8][4];
#pragma unroll
    for (int i = 0; i < BKV / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
      for (int nb = 0; nb < BKV / 16; ++nb) {
        uint32_t kf[4];
        ldsm_x4(kf, smem_u32(cK + (nb * 16 + (lane >> 4) * 8 + (lane & 7)) * LDS + kk * 16 + ((lane >> 3) & 1) * 8));
        MmaType<T>::mma(s[2 * nb], qf[kk], kf[0], kf[1]);
        MmaType<T>::mma(s[2 * nb + 1], qf[kk], kf[2], kf[3]);
      }
    }
    // scale / bias / mask, online softmax
    const int i0 = q0 + warp * 16 + g;
    float m_new[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nb = 0; nb < BKV / 8; ++nb) {
      const int j = kv * BKV + nb * 8 + 2 * t4;
      s[nb][0] = sc.apply(s[nb][0], i0, j);
      s[nb][1] = sc.apply(s[nb][1], i0, j + 1);
      s[nb][2] = sc.apply(s[nb][2], i0 + 8, j);
      s[nb][3] = sc.apply(s[nb][3], i0 + 8, j + 1);
      m_new[0] = fmaxf(m_new[0], fmaxf(s[nb][0], s[nb][1]));
      m_new[1] = fmaxf(m_new[1], fmaxf(s[nb][2], s[nb][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 1));
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 2));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      msc[r] = (m_new[r] == -INFINITY) ? 0.f : m_new[r] * LOG2E;
      corr[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f(m_run[r] * LOG2E - msc[r]);
      m_run[r] = m_new[r];
      l_run[r] *= corr[r];
    }
    uint32_t pf[BKV / 16][4];
#pragma unroll
    for (int nb = 0; nb < BKV / 8; ++nb) {
      const float p0 = exp2f(s[nb][0] * LOG2E - msc[0]), p1 = exp2f(s[nb][1] * LOG2E - msc[0]);
      const float p2 = exp2f(s[nb][2] * LOG2E - msc[1]), p3 = exp2f(s[nb][3] * LOG2E - msc[1]);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      if constexpr (DROP) {                      // the row sums stay those of the undropped P
        const int j = kv * BKV + nb * 8 + 2 * t4;
        const uint32_t w0 = drop_word(dkey, drow[0], j >> 2), w1 = drop_word(dkey, drow[1], j >> 2);
        pf[nb >> 1][(nb & 1) * 2] = MmaType<T>::pack(drop_keep(w0, j, dthr) ? p0 : 0.f, drop_keep(w0, j + 1, dthr) ? p1 : 0.f);
        pf[nb >> 1][(nb & 1) * 2 + 1] = MmaType<T>::pack(drop_keep(w1, j, dthr) ? p2 : 0.f, drop_keep(w1, j + 1, dthr) ? p3 : 0.f);
        continue;
      }
      pf[nb >> 1][(nb & 1) * 2] = MmaType<T>::pack(p0, p1);
      pf[nb >> 1][(nb & 1) * 2 + 1] = MmaType<T>::pack(p2, p3);
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o_acc[i][0] *= corr[0]; o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1]; o_acc[i][3] *= corr[1];
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {
#pragma unroll
      for (int db = 0; db < HD / 16; ++db) {
        uint32_t vf[4];
        ldsm_x4_t(vf, smem_u32(cV + (kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LDS + db * 16 + (lane >> 4) * 8));
        MmaType<T>::mma(o_acc[2 * db], pf[kk], vf[0], vf[1]);
        MmaType<T>::mma(o_acc[2 * db + 1], pf[kk], vf[2], vf[3]);
      }
    }
    __syncthreads();
  }
  // finalize
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  T* go = static_cast<T*>(p.o) + b * p.o_bs + static_cast<long long>(h) * p.hd;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = q0 + warp * 16 + g + r * 8;
    if (i >= p.Lq) continue;
    const float inv = (l_run[r] > 0.f ? 1.f / l_run[r] : 0.f) * dscale;
#pragma unroll
    for (int nb = 0; nb < HD / 8; ++nb) {
      const int c = nb * 8 + 2 * t4;
      if (c < p.hd) {
        const uint32_t w = MmaType<T>::pack(o_acc[nb][2 * r] * inv, o_acc[nb][2 * r + 1] * inv);
        *reinterpret_cast<uint32_t*>(go + static_cast<long long>(i) * p.o_rs + c) = w;
      }
    }
    if (p.lse && t4 == 0)
      p.lse[(static_cast<long long>(b) * p.H + h) * p.Lq + i] = m_run[r] + logf(l_run[r]);
  }
}

// ---------------------------------------------------------------- few queries x a few hundred keys (Q-Former cross-attention)
// Qformer.py:198-268 with encoder_hidden_states: 32 query tokens attend to the 257 ViT tokens of their frame, hd 64, no bias,
// all-ones mask.  One CTA per (frame, head): the whole K and V (257 x 64 each) are fetched with ONE burst of cp.async (~66 KB
// in flight per CTA, two CTAs per SM -- the op is a pure HBM stream of the batched K/V projection), the four warps split the
// keys in 16-key groups, each computes S / softmax partials / P.V for all 32 queries over its keys with mma.sync, and the
// partial (max, sum, O) triples are merged through shared memory.  No loop-carried barriers, no running rescale.
constexpr int XQ_MAX_LK = 320;
// XW warps: 4 (five 16-key groups per warp) or 8 (three): the eight-warp form halves every warp's serial chain (S, softmax, P.V,
// merge) and the cp.async issue per thread at the same shared memory per CTA -- but its 116-127 registers x 256 threads leave two
// CTAs per SM instead of three, and it measured SLOWER on a B200 (path 1.80-1.83 -> 2.0-2.2 ms per step, call 26): the four-warp
// form stays the default, MRB_XQ_WARPS=8 selects the other.
template <typename T, bool DROP = false, int XW = 4>
__global__ void __launch_bounds__(32 * XW) attn_xq_kernel(const typename AttnParamsOf<DROP>::type p) {
  constexpr int XQ_MAXG = (XQ_MAX_LK / 16 + XW - 1) / XW, NT = 32 * XW;
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  // rows are 128 B (64 x 16 bit), 16-byte chunk c of row r lives at chunk (c ^ (r & 7)): conflict-free ldmatrix without
  // padding, so Q + K + V of 257 keys take 72 KB and THREE CTAs share an SM
  constexpr int HD = 64, LDS = HD, LQ = 32;
  extern __shared__ __align__(128) uint8_t smem_attn[];
  const int LkP = (p.Lk + 15) & ~15;
  T* sQ = reinterpret_cast<T*>(smem_attn);
  T* sK = sQ + LQ * LDS;
  T* sV = sK + LkP * LDS;
  auto sw = [](int r, int c) { return r * LDS + ((c ^ (r & 7)) << 3); };     // element offset of chunk c in row r
  const int b = blockIdx.y, h = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const T* gq = static_cast<const T*>(p.q) + b * p.q_bs + static_cast<long long>(h) * HD;
  const T* gk = static_cast<const T*>(p.k) + b * p.k_bs + static_cast<long long>(h) * HD;
  const T* gv = static_cast<const T*>(p.v) + b * p.v_bs + static_cast<long long>(h) * HD;
  // two commit groups: Q and K first (the scores and the softmax start as soon as they have landed), V behind them
  for (int i = threadIdx.x; i < LQ * 8; i += NT) {
    const int r = i >> 3, c = i & 7;
    const bool ok = r < p.Lq;
    cp_async16(smem_u32(sQ + sw(r, c)), ok ? gq + static_cast<long long>(r) * p.q_rs + c * 8 : gq, ok);
  }
  for (int i = threadIdx.x; i < LkP * 8; i += NT) {
    const int r = i >> 3, c = i & 7;
    const bool ok = r < p.Lk;
    cp_async16(smem_u32(sK + sw(r, c)), ok ? gk + static_cast<long long>(r) * p.k_rs + c * 8 : gk, ok);
  }
  cp_async_commit();
  for (int i = threadIdx.x; i < LkP * 8; i += NT) {
    const int r = i >> 3, c = i & 7;
    const bool ok = r < p.Lk;
    cp_async16(smem_u32(sV + sw(r, c)), ok ? gv + static_cast<long long>(r) * p.v_rs + c * 8 : gv, ok);
  }
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();

  // this warp's contiguous range of 16-key groups
  const int NG = LkP >> 4, base = NG / XW, rem = NG % XW;
  const int ng = base + (warp < rem ? 1 : 0);
  const int g0 = warp * base + min(warp, rem);
  const int krow0 = g0 * 16;

  float sc[2][2 * XQ_MAXG][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nb = 0; nb < 2 * XQ_MAXG; ++nb) sc[mt][nb][0] = sc[mt][nb][1] = sc[mt][nb][2] = sc[mt][nb][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
    uint32_t qf[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
      ldsm_x4(qf[mt], smem_u32(sQ + sw(mt * 16 + (lane & 15), kk * 2 + (lane >> 4))));
#pragma unroll
    for (int gi = 0; gi < XQ_MAXG; ++gi) {
      if (gi < ng) {
        uint32_t kf[4];
        ldsm_x4(kf, smem_u32(sK + sw(krow0 + gi * 16 + (lane >> 4) * 8 + (lane & 7), kk * 2 + ((lane >> 3) & 1))));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          MmaType<T>::mma(sc[mt][2 * gi], qf[mt], kf[0], kf[1]);
          MmaType<T>::mma(sc[mt][2 * gi + 1], qf[mt], kf[2], kf[3]);
        }
      }
    }
  }
  // softmax partials over this warp's keys (scores in the log2 domain)
  const float sl2 = p.scale * 1.4426950408889634f;
  float mx[2][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}}, ls[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nb = 0; nb < 2 * XQ_MAXG; ++nb) {
      const int j = (g0 + (nb >> 1)) * 16 + (nb & 1) * 8 + 2 * t4;
      const bool in = (nb >> 1) < ng;
      sc[mt][nb][0] = (in && j < p.Lk) ? sc[mt][nb][0] * sl2 : -INFINITY;
      sc[mt][nb][1] = (in && j + 1 < p.Lk) ? sc[mt][nb][1] * sl2 : -INFINITY;
      sc[mt][nb][2] = (in && j < p.Lk) ? sc[mt][nb][2] * sl2 : -INFINITY;
      sc[mt][nb][3] = (in && j + 1 < p.Lk) ? sc[mt][nb][3] * sl2 : -INFINITY;
      mx[mt][0] = fmaxf(mx[mt][0], fmaxf(sc[mt][nb][0], sc[mt][nb][1]));
      mx[mt][1] = fmaxf(mx[mt][1], fmaxf(sc[mt][nb][2], sc[mt][nb][3]));
    }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[mt][r] = fmaxf(mx[mt][r], __shfl_xor_sync(0xffffffffu, mx[mt][r], 1));
      mx[mt][r] = fmaxf(mx[mt][r], __shfl_xor_sync(0xffffffffu, mx[mt][r], 2));
    }
  uint32_t pf[2][XQ_MAXG][4];
  uint32_t dkey = 0, dhead = 0, dng = 0, dthr = 0;
  float dscale = 1.f;
  if constexpr (DROP) {                          // Qformer.py:258 in train mode: masks of dropmask.cuh, row = (b H + h) Lq + i
    dkey = drop_key(*p.drop_seed, p.drop_site);
    dthr = p.drop_thr; dscale = p.drop_scale;
    dhead = static_cast<uint32_t>(b * p.H + h) * static_cast<uint32_t>(p.Lq);
    dng = drop_groups(static_cast<uint32_t>(p.Lk));
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const float m0 = mx[mt][0] == -INFINITY ? 0.f : mx[mt][0], m1 = mx[mt][1] == -INFINITY ? 0.f : mx[mt][1];
#pragma unroll
    for (int nb = 0; nb < 2 * XQ_MAXG; ++nb) {
      const float p0 = exp2f(sc[mt][nb][0] - m0), p1 = exp2f(sc[mt][nb][1] - m0);
      const float p2 = exp2f(sc[mt][nb][2] - m1), p3 = exp2f(sc[mt][nb][3] - m1);
      ls[mt][0] += p0 + p1;
      ls[mt][1] += p2 + p3;
      if constexpr (DROP) {                      // the sums stay those of the undropped P; 1 / (1 - p) goes into the final 1 / L
        const int j = (g0 + (nb >> 1)) * 16 + (nb & 1) * 8 + 2 * t4;
        const uint32_t r0 = static_cast<uint32_t>(min(mt * 16 + g, p.Lq - 1)), r1 = static_cast<uint32_t>(min(mt * 16 + g + 8, p.Lq - 1));
        const uint32_t w0 = drop_word(dkey, (dhead + r0) * dng, static_cast<uint32_t>(j) >> 2);
        const uint32_t w1 = drop_word(dkey, (dhead + r1) * dng, static_cast<uint32_t>(j) >> 2);
        pf[mt][nb >> 1][(nb & 1) * 2] = MmaType<T>::pack(drop_keep(w0, j, dthr) ? p0 : 0.f, drop_keep(w0, j + 1, dthr) ? p1 : 0.f);
        pf[mt][nb >> 1][(nb & 1) * 2 + 1] = MmaType<T>::pack(drop_keep(w1, j, dthr) ? p2 : 0.f, drop_keep(w1, j + 1, dthr) ? p3 : 0.f);
        continue;
      }
      pf[mt][nb >> 1][(nb & 1) * 2] = MmaType<T>::pack(p0, p1);
      pf[mt][nb >> 1][(nb & 1) * 2 + 1] = MmaType<T>::pack(p2, p3);
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      ls[mt][r] += __shfl_xor_sync(0xffffffffu, ls[mt][r], 1);
      ls[mt][r] += __shfl_xor_sync(0xffffffffu, ls[mt][r], 2);
    }
  cp_async_wait<0>();
  __syncthreads();                               // V has landed (every thread's copies)
  // O partial = P V over this warp's keys
  float oa[2][HD / 8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) oa[mt][i][0] = oa[mt][i][1] = oa[mt][i][2] = oa[mt][i][3] = 0.f;
#pragma unroll
  for (int gi = 0; gi < XQ_MAXG; ++gi) {
    if (gi < ng) {
#pragma unroll
      for (int db = 0; db < HD / 16; ++db) {
        uint32_t vf[4];
        ldsm_x4_t(vf, smem_u32(sV + sw(krow0 + gi * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), db * 2 + (lane >> 4))));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          MmaType<T>::mma(oa[mt][2 * db], pf[mt][gi], vf[0], vf[1]);
          MmaType<T>::mma(oa[mt][2 * db + 1], pf[mt][gi], vf[2], vf[3]);
        }
      }
    }
  }
  __syncthreads();                               // every warp is done with K / V: reuse their space for the partials
  constexpr int LDO = HD + 2;
  float* sO = reinterpret_cast<float*>(sK);      // [XW warps][32 rows][LDO]  (the launcher sizes the K / V region for it)
  float* sM = sO + XW * LQ * LDO;                // [XW][32] running max (log2 domain), [XW][32] sums
  float* sL = sM + XW * LQ;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = mt * 16 + g + r * 8;
#pragma unroll
      for (int nb = 0; nb < HD / 8; ++nb)
        *reinterpret_cast<float2*>(sO + (warp * LQ + row) * LDO + nb * 8 + 2 * t4) = make_float2(oa[mt][nb][2 * r], oa[mt][nb][2 * r + 1]);
      if (t4 == 0) { sM[warp * LQ + row] = mx[mt][r]; sL[warp * LQ + row] = ls[mt][r]; }
    }
  __syncthreads();
  {
    constexpr int TPR = NT / LQ, CW = HD / TPR;  // threads per row (4 / 8), columns per thread (16 / 8)
    const int row = threadIdx.x / TPR, c0 = (threadIdx.x % TPR) * CW;
    if (row < p.Lq) {
      float m = -INFINITY;
#pragma unroll
      for (int w = 0; w < XW; ++w) m = fmaxf(m, sM[w * LQ + row]);
      float f[XW], L = 0.f;
#pragma unroll
      for (int w = 0; w < XW; ++w) {
        const float mw = sM[w * LQ + row];
        f[w] = mw == -INFINITY ? 0.f : exp2f(mw - m);
        L += f[w] * sL[w * LQ + row];
      }
      const float inv = (L > 0.f ? 1.f / L : 0.f) * dscale;
      T* go = static_cast<T*>(p.o) + b * p.o_bs + static_cast<long long>(row) * p.o_rs + static_cast<long long>(h) * HD + c0;
      uint32_t w8[CW / 2];
#pragma unroll
      for (int c = 0; c < CW; c += 2) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int w = 0; w < XW; ++w) {
          const float2 v = *reinterpret_cast<const float2*>(sO + (w * LQ + row) * LDO + c0 + c);
          a0 = fmaf(f[w], v.x, a0);
          a1 = fmaf(f[w], v.y, a1);
        }
        w8[c >> 1] = MmaType<T>::pack(a0 * inv, a1 * inv);
      }
#pragma unroll
      for (int c = 0; c < CW / 8; ++c)
        *reinterpret_cast<uint4*>(go + 8 * c) = make_uint4(w8[4 * c], w8[4 * c + 1], w8[4 * c + 2], w8[4 * c + 3]);
    }
  }
}

// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]
template <typename T>
__global__ void attn_delta_kernel(const AttnParams p) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int total = p.B * p.H * p.Lq;
  if (row >= total) return;
  const int i = row % p.Lq, h = (row / p.Lq) % p.H, b = row / (p.Lq * p.H);
  const T* o = static_cast<const T*>(p.o) + b * p.o_bs + static_cast<long long>(i) * p.o_rs + h * p.hd;
  const T* d = static_cast<const T*>(p.dout) + b * p.do_bs + static_cast<long long>(i) * p.do_rs + h * p.hd;
  float acc = 0.f;
  for (int c = lane; c < p.hd; c += 32) acc += to_f32(o[c]) * to_f32(d[c]);
  acc = warp_sum(acc);
  if (lane == 0) const_cast<float*>(p.delta)[row] = acc;
}

// dQ = scale * sum_j dS_ij K_j   with dS = P * (dO V^T - delta).  One CTA per 64 query rows.
template <typename T, int HD, bool DROP = false>
__global__ void __launch_bounds__(NTHREADS) attn_bwd_dq_kernel(const typename AttnParamsOf<DROP>::type p) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  constexpr int LDS = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  T* sQ = reinterpret_cast<T*>(smem_attn);
  T* sdO = sQ + BQ * LDS;
  T* sK = sdO + BQ * LDS;
  T* sV = sK + 2 * BKV * LDS;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const T* gq = static_cast<const T*>(p.q) + b * p.q_bs + static_cast<long long>(h) * p.hd;
  const T* gk = static_cast<const T*>(p.k) + (b / p.kv_div) * p.k_bs + static_cast<long long>(h) * p.hd;
  const T* gv = static_cast<const T*>(p.v) + (b / p.kv_div) * p.v_bs + static_cast<long long>(h) * p.hd;
  const T* gdo = static_cast<const T*>(p.dout) + b * p.do_bs + static_cast<long long>(h) * p.hd;
  ScoreCtx sc{p.scale, p.bias ? p.bias + static_cast<long long>(h) * p.bias_len : nullptr, p.bias_zero,
              p.kmask ? p.kmask + static_cast<long long>(b / p.kv_div) * p.Lk : nullptr, p.causal, p.q_pos0, p.Lk, p.Lq};
  int n_kv = (p.Lk + BKV - 1) / BKV;
  if (p.causal) n_kv = min(n_kv, (q0 + BQ - 1 + p.q_pos0) / BKV + 1);

  load_tile<T, HD, BQ>(smem_u32(sQ), gq, p.q_rs, q0, p.Lq, p.hd);
  load_tile<T, HD, BQ>(smem_u32(sdO), gdo, p.do_rs, q0, p.Lq, p.hd);
  load_tile<T, HD, BKV>(smem_u32(sK), gk, p.k_rs, 0, p.Lk, p.hd);
  load_tile<T, HD, BKV>(smem_u32(sV), gv, p.v_rs, 0, p.Lk, p.hd);
  cp_async_commit();

  const long long stat = (static_cast<long long>(b) * p.H + h) * p.Lq;
  float lse[2], dl[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = q0 + warp * 16 + g + r * 8;
    lse[r] = i < p.Lq ? p.lse[stat + i] : 0.f;
    dl[r] = i < p.Lq ? p.delta[stat + i] : 0.f;
  }
  float dq_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) dq_acc[i][0] = dq_acc[i][1] = dq_acc[i][2] = dq_acc[i][3] = 0.f;
  uint32_t qf[HD / 16][4], dof[HD / 16][4];
  uint32_t dkey = 0, drow[2] = {0u, 0u}, dthr = 0;
  float dscale = 1.f;
  if constexpr (DROP) {
    dkey = drop_key(*p.drop_seed, p.drop_site);
    dthr = p.drop_thr; dscale = p.drop_scale;
#pragma unroll
    for (int r = 0; r < 2; ++r)
      drow[r] = (static_cast<uint32_t>(b * p.H + h) * static_cast<uint32_t>(p.Lq) +
                 static_cast<uint32_t>(min(q0 + warp * 16 + g + 8 * r, p.Lq - 1))) * drop_groups(static_cast<uint32_t>(p.Lk));
  }

  for (int kv = 0; kv < n_kv; ++kv) {
    const int buf = kv & 1;
    if (kv + 1 < n_kv) {
      load_tile<T, HD, BKV>(smem_u32(sK + (buf ^ 1) * BKV * LDS), gk, p.k_rs, (kv + 1) * BKV, p.Lk, p.hd);
      load_tile<T, HD, BKV>(smem_u32(sV + (buf ^ 1) * BKV * LDS), gv, p.v_rs, (kv + 1) * BKV, p.Lk, p.hd);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kv == 0) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        ldsm_x4(qf[kk], smem_u32(sQ + (warp * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8));
        ldsm_x4(dof[kk], smem_u32(sdO + (warp * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8));
      }
    }
    const T* cK = sK + buf * BKV * LDS;
    const T* cV = sV + buf * BKV * LDS;
    float s[BKV / 8][4], dp[BKV / 8][4];
#pragma unroll
    for (int i = 0; i < BKV / 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
      for (int nb = 0; nb < BKV / 16; ++nb) {
        uint32_t kf[4], vf[4];
        const int off = (nb * 16 + (lane >> 4) * 8 + (lane & 7)) * LDS + kk * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(kf, smem_u32(cK + off));
        ldsm_x4(vf, smem_u32(cV + off));
        MmaType<T>::mma(s[2 * nb], qf[kk], kf[0], kf[1]);
        MmaType<T>::mma(s[2 * nb + 1], qf[kk], kf[2], kf[3]);
        MmaType<T>::mma(dp[2 * nb], dof[kk], vf[0], vf[1]);
        MmaType<T>::mma(dp[2 * nb + 1], dof[kk], vf[2], vf[3]);
      }
    }
    const int i0 = q0 + warp * 16 + g;
    uint32_t dsf[BKV / 16][4];
#pragma unroll
    for (int nb = 0; nb < BKV / 8; ++nb) {
      const int j = kv * BKV + nb * 8 + 2 * t4;
      float ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const float sv = sc.apply(s[nb][e], i0 + r * 8, j + (e & 1));
        const float pr = (sv == -INFINITY) ? 0.f : __expf(sv - lse[r]);
        float dpe = dp[nb][e];
        if constexpr (DROP) dpe = drop_keep(drop_word(dkey, drow[r], j >> 2), j + (e & 1), dthr) ? dpe * dscale : 0.f;
        ds[e] = pr * (dpe - dl[r]) * p.scale;
      }
      dsf[nb >> 1][(nb & 1) * 2] = MmaType<T>::pack(ds[0], ds[1]);
      dsf[nb >> 1][(nb & 1) * 2 + 1] = MmaType<T>::pack(ds[2], ds[3]);
    }
    // dQ += dS K   (B operand = K [key x d] -> transposed ldmatrix)
#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {
#pragma unroll
      for (int db = 0; db < HD / 16; ++db) {
        uint32_t kf[4];
        ldsm_x4_t(kf, smem_u32(cK + (kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LDS + db * 16 + (lane >> 4) * 8));
        MmaType<T>::mma(dq_acc[2 * db], dsf[kk], kf[0], kf[1]);
        MmaType<T>::mma(dq_acc[2 * db + 1], dsf[kk], kf[2], kf[3]);
      }
    }
    __syncthreads();
  }
  T* gdq = static_cast<T*>(p.dq) + b * p.q_bs + static_cast<long long>(h) * p.hd;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = q0 + warp * 16 + g + r * 8;
    if (i >= p.Lq) continue;
#pragma unroll
    for (int nb = 0; nb < HD / 8; ++nb) {
      const int c = nb * 8 + 2 * t4;
      if (c < p.hd)
        *reinterpret_cast<uint32_t*>(gdq + static_cast<long long>(i) * p.q_rs + c) =
            MmaType<T>::pack(dq_acc[nb][2 * r], dq_acc[nb][2 * r + 1]);
    }
  }
}

// dK_j = scale * sum_i dS_ij Q_i,  dV_j = sum_i P_ij dO_i.  One CTA per 64 keys; works on the transposed
// problem (keys are the MMA M dimension) so P^T / dS^T fragments feed the second MMAs directly.
template <typename T, int HD, bool DROP = false>
__global__ void __launch_bounds__(NTHREADS) attn_bwd_dkv_kernel(const typename AttnParamsOf<DROP>::type p) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  constexpr int LDS = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  T* sK = reinterpret_cast<T*>(smem_attn);
  T* sV = sK + BKV * LDS;
  T* sQ = sV + BKV * LDS;
  T* sdO = sQ + 2 * BQ * LDS;
  float* sLse = reinterpret_cast<float*>(sdO + 2 * BQ * LDS);
  float* sDl = sLse + 2 * BQ;
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * BKV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const T* gq = static_cast<const T*>(p.q) + b * p.q_bs + static_cast<long long>(h) * p.hd;
  const T* gk = static_cast<const T*>(p.k) + (b / p.kv_div) * p.k_bs + static_cast<long long>(h) * p.hd;
  const T* gv = static_cast<const T*>(p.v) + (b / p.kv_div) * p.v_bs + static_cast<long long>(h) * p.hd;
  const T* gdo = static_cast<const T*>(p.dout) + b * p.do_bs + static_cast<long long>(h) * p.hd;
  ScoreCtx sc{p.scale, p.bias ? p.bias + static_cast<long long>(h) * p.bias_len : nullptr, p.bias_zero,
              p.kmask ? p.kmask + static_cast<long long>(b / p.kv_div) * p.Lk : nullptr, p.causal, p.q_pos0, p.Lk, p.Lq};
  const long long stat = (static_cast<long long>(b) * p.H + h) * p.Lq;
  const int n_q = (p.Lq + BQ - 1) / BQ;
  int q_begin = 0;
  if (p.causal) q_begin = max(0, (k0 - p.q_pos0) / BQ);   // rows i with i + q_pos0 >= k0

  auto load_q = [&](int qt, int buf) {
    load_tile<T, HD, BQ>(smem_u32(sQ + buf * BQ * LDS), gq, p.q_rs, qt * BQ, p.Lq, p.hd);
    load_tile<T, HD, BQ>(smem_u32(sdO + buf * BQ * LDS), gdo, p.do_rs, qt * BQ, p.Lq, p.hd);
    if (threadIdx.x < BQ) {
      const int i = qt * BQ + threadIdx.x;
      sLse[buf * BQ + threadIdx.x] = i < p.Lq ? p.lse[stat + i] : 0.f;
      sDl[buf * BQ + threadIdx.x] = i < p.Lq ? p.delta[stat + i] : 0.f;
    }
  };
  load_tile<T, HD, BKV>(smem_u32(sK), gk, p.k_rs, k0, p.Lk, p.hd);
  load_tile<T, HD, BKV>(smem_u32(sV), gv, p.v_rs, k0, p.Lk, p.hd);
  if (q_begin < n_q) load_q(q_begin, 0);
  cp_async_commit();

  float dk_acc[HD / 8][4], dv_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    dk_acc[i][0] = dk_acc[i][1] = dk_acc[i][2] = dk_acc[i][3] = 0.f;
    dv_acc[i][0] = dv_acc[i][1] = dv_acc[i][2] = dv_acc[i][3] = 0.f;
  }
  uint32_t kf_a[HD / 16][4], vf_a[HD / 16][4];
  uint32_t dkey = 0, dhead = 0, dng = 0, dthr = 0;
  float dscale = 1.f;
  if constexpr (DROP) {
    dkey = drop_key(*p.drop_seed, p.drop_site);
    dthr = p.drop_thr; dscale = p.drop_scale;
    dhead = static_cast<uint32_t>(b * p.H + h) * static_cast<uint32_t>(p.Lq);
    dng = drop_groups(static_cast<uint32_t>(p.Lk));
  }

  for (int qt = q_begin; qt < n_q; ++qt) {
    const int buf = (qt - q_begin) & 1;
    if (qt + 1 < n_q) {
      load_q(qt + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (qt == q_begin) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        ldsm_x4(kf_a[kk], smem_u32(sK + (warp * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8));
        ldsm_x4(vf_a[kk], smem_u32(sV + (warp * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8));
      }
    }
    const T* cQ = sQ + buf * BQ * LDS;
    const T* cdO = sdO + buf * BQ * LDS;
    // S^T = K Q^T, dP^T = V dO^T   (16 keys x 64 queries per warp)
    float st[BQ / 8][4], dpt[BQ / 8][4];
#pragma unroll
    for (int i = 0; i < BQ / 8; ++i) {
      st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
      dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
      for (int nb = 0; nb < BQ / 16; ++nb) {
        uint32_t qf[4], dof[4];
        const int off = (nb * 16 + (lane >> 4) * 8 + (lane & 7)) * LDS + kk * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(qf, smem_u32(cQ + off));
        ldsm_x4(dof, smem_u32(cdO + off));
        MmaType<T>::mma(st[2 * nb], kf_a[kk], qf[0], qf[1]);
        MmaType<T>::mma(st[2 * nb + 1], kf_a[kk], qf[2], qf[3]);
        MmaType<T>::mma(dpt[2 * nb], vf_a[kk], dof[0], dof[1]);
        MmaType<T>::mma(dpt[2 * nb + 1], vf_a[kk], dof[2], dof[3]);
      }
    }
    const int j0 = k0 + warp * 16 + g;
    uint32_t ptf[BQ / 16][4], dstf[BQ / 16][4];
#pragma unroll
    for (int nb = 0; nb < BQ / 8; ++nb) {
      float pr[4], ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int il = nb * 8 + 2 * t4 + (e & 1);       // query within tile (column of S^T)
        const int i = qt * BQ + il;
        const int j = j0 + (e >> 1) * 8;
        const float sv = (i < p.Lq) ? sc.apply(st[nb][e], i, j) : -INFINITY;
        pr[e] = (sv == -INFINITY) ? 0.f : __expf(sv - sLse[buf * BQ + il]);
        float dpe = dpt[nb][e];
        bool keep = true;
        if constexpr (DROP) {
          keep = drop_keep(drop_word(dkey, (dhead + static_cast<uint32_t>(min(i, p.Lq - 1))) * dng, static_cast<uint32_t>(j) >> 2), j, dthr);
          dpe = keep ? dpe * dscale : 0.f;
        }
        ds[e] = pr[e] * (dpe - sDl[buf * BQ + il]) * p.scale;
        if constexpr (DROP) pr[e] = keep ? pr[e] : 0.f;      // P^T operand of dV: dropped; 1 / (1 - p) is applied to dV at the end
      }
      ptf[nb >> 1][(nb & 1) * 2] = MmaType<T>::pack(pr[0], pr[1]);
      ptf[nb >> 1][(nb & 1) * 2 + 1] = MmaType<T>::pack(pr[2], pr[3]);
      dstf[nb >> 1][(nb & 1) * 2] = MmaType<T>::pack(ds[0], ds[1]);
      dstf[nb >> 1][(nb & 1) * 2 + 1] = MmaType<T>::pack(ds[2], ds[3]);
    }
    // dV += P^T dO ; dK += dS^T Q    (B operands [query x d] -> transposed ldmatrix)
#pragma unroll
    for (int kk = 0; kk < BQ / 16; ++kk) {
#pragma unroll
      for (int db = 0; db < HD / 16; ++db) {
        uint32_t f1[4], f2[4];
        const int off = (kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LDS + db * 16 + (lane >> 4) * 8;
        ldsm_x4_t(f1, smem_u32(cdO + off));
        ldsm_x4_t(f2, smem_u32(cQ + off));
        MmaType<T>::mma(dv_acc[2 * db], ptf[kk], f1[0], f1[1]);
        MmaType<T>::mma(dv_acc[2 * db + 1], ptf[kk], f1[2], f1[3]);
        MmaType<T>::mma(dk_acc[2 * db], dstf[kk], f2[0], f2[1]);
        MmaType<T>::mma(dk_acc[2 * db + 1], dstf[kk], f2[2], f2[3]);
      }
    }
    __syncthreads();
  }
  T* gdk = static_cast<T*>(p.dk) + b * p.k_bs + static_cast<long long>(h) * p.hd;
  T* gdv = static_cast<T*>(p.dv) + b * p.v_bs + static_cast<long long>(h) * p.hd;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int j = k0 + warp * 16 + g + r * 8;
    if (j >= p.Lk) continue;
#pragma unroll
    for (int nb = 0; nb < HD / 8; ++nb) {
      const int c = nb * 8 + 2 * t4;
      if (c < p.hd) {
        *reinterpret_cast<uint32_t*>(gdk + static_cast<long long>(j) * p.k_rs + c) = MmaType<T>::pack(dk_acc[nb][2 * r], dk_acc[nb][2 * r + 1]);
        *reinterpret_cast<uint32_t*>(gdv + static_cast<long long>(j) * p.v_rs + c) =
            MmaType<T>::pack(dv_acc[nb][2 * r] * dscale, dv_acc[nb][2 * r + 1] * dscale);
      }
    }
  }
}

// Single-query-row attention (the 257th ViT token: 257 = 2 x 128 + 1, the two full tiles go through the tcgen05 kernel).
// One warp per (batch, head); lane l owns head-dim elements [4l, 4l+4).  No bias / mask.
template <typename T>
__global__ void __launch_bounds__(128) attn_row_kernel(const T* __restrict__ q, long long q_bs, const T* __restrict__ k,
                                                       long long k_bs, long long k_rs, const T* __restrict__ v, long long v_bs,
                                                       long long v_rs, T* __restrict__ o, long long o_bs, int B, int H, int Lk,
                                                       int hd, float scale) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  extern __shared__ float sp[];                       // [4 warps][Lk] scores / probabilities
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x * 4 + warp;
  if (bh >= B * H) return;
  const int b = bh / H, h = bh % H;
  float* p = sp + warp * Lk;
  const bool act = lane * 4 < hd;
  const int d0 = lane * 4;
  float qv[4] = {0.f, 0.f, 0.f, 0.f};
  if (act) {
    const uint2 w = *reinterpret_cast<const uint2*>(q + b * q_bs + h * hd + d0);
    const T* e = reinterpret_cast<const T*>(&w);
#pragma unroll
    for (int i = 0; i < 4; ++i) qv[i] = to_f32(e[i]) * scale;
  }
  // scores: lane = key; the query row is broadcast from shared memory
  float* sq = sp + 4 * Lk + warp * 128;
  if (act) { sq[d0] = qv[0]; sq[d0 + 1] = qv[1]; sq[d0 + 2] = qv[2]; sq[d0 + 3] = qv[3]; }
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < Lk; j += 32) {
    const T* kr = k + b * k_bs + static_cast<long long>(j) * k_rs + h * hd;
    float s = 0.f;
    for (int d = 0; d < hd; d += 8) {
      const uint4 w = *reinterpret_cast<const uint4*>(kr + d);
      const T* e = reinterpret_cast<const T*>(&w);
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(sq[d + i], to_f32(e[i]), s);
    }
    p[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  __syncwarp();
  float sum = 0.f;
  for (int j = lane; j < Lk; j += 32) {
    const float e = __expf(p[j] - mx);
    p[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const T* vb = v + b * v_bs + h * hd + d0;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (act) {
#pragma unroll 8
    for (int j = 0; j < Lk; ++j) {
      const uint2 w = *reinterpret_cast<const uint2*>(vb + static_cast<long long>(j) * v_rs);
      const T* e = reinterpret_cast<const T*>(&w);
      const float pj = p[j];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(pj, to_f32(e[i]), acc[i]);
    }
    const float inv = 1.f / sum;
    T r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = from_f32<T>(acc[i] * inv);
    *reinterpret_cast<uint2*>(o + b * o_bs + h * hd + d0) = *reinterpret_cast<const uint2*>(r);
  }
}

// ---------------------------------------------------------------- few query rows x thousands of keys (T5 decoder cross-attention)
// modeling_t5.py:474-620 with key_value_states = the encoder output: the ~16 target tokens of a clip attend to its ~2000 encoder
// positions.  The long-sequence kernels tile the QUERY rows over CTAs and walk the keys serially, so this shape ran on one CTA per
// (clip, head) with 16 of its 128 / 256 rows in use: 49 us forward, 77 + 73 us backward per decoder layer on the step's latency-
// bound chain.  Here the KEYS are split instead: a cluster of FQ_SPLIT CTAs per (16-row block, head, clip), every CTA streams its
// share of the 64-key tiles through a 3-deep cp.async ring, its four warps take 16 keys of a tile each (mma.sync m16n8k16, the 16
// query rows are the M dimension), and the partial results -- (max, sum, O) triples in the forward, dQ in the backward -- meet in
// shared memory inside the CTA and through distributed shared memory inside the cluster, always in the same order (no atomics: run-
// to-run reproducible).  dK / dV of this shape come from attn_bwd_dkv_kernel, which already tiles the keys over CTAs.
#ifdef MRB_HOST_SHIM          // tests/cuda_host_shim has no clusters: one CTA takes all keys, the merge code below runs with one part
constexpr int FQ_SPLIT = 1;
#define MRB_FQ_CLUSTER
__device__ __forceinline__ void fq_cluster_sync() { __syncthreads(); }      // what the cluster barrier is for ONE CTA
__device__ __forceinline__ float fq_ld_part(const float* local, int) { return *local; }
#else
constexpr int FQ_SPLIT = 4;
#define MRB_FQ_CLUSTER __cluster_dims__(FQ_SPLIT, 1, 1)
__device__ __forceinline__ void fq_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float fq_ld_part(const float* local, int rank) {     // the same shared-memory variable in CTA `rank` of the cluster
  uint32_t remote;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}
#endif
constexpr int FQ_ROWS = 16, FQ_ST = 3;

// shared-memory layout of both kernels: Q (and dO) tile(s), then the K / V ring; once the key loop is over the ring is dead and holds
// the warps' partial results [4][16][64] (+ max / sum [4][16] each) and the CTA's partial [16][64] (+ [16] + [16])
struct FqSmem {
  static constexpr int LDS = 64 + 8;
  static constexpr int RING = FQ_ST * BKV * LDS;                      // elements, one of K / V
  static constexpr int W_O = 0, W_M = 4 * FQ_ROWS * 64, W_L = W_M + 4 * FQ_ROWS, C_O = W_L + 4 * FQ_ROWS, C_M = C_O + FQ_ROWS * 64,
                       C_L = C_M + FQ_ROWS, PART_FLOATS = C_L + FQ_ROWS;
  static_assert(PART_FLOATS * 4 <= RING * 2, "partials must fit into the K ring");
  static constexpr int bytes(int q_tiles) { return (q_tiles * FQ_ROWS * LDS + 2 * RING) * 2; }
};

template <typename T, bool DROP = false>
__global__ void MRB_FQ_CLUSTER __launch_bounds__(NTHREADS) attn_fq_fwd_kernel(const typename AttnParamsOf<DROP>::type p) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  constexpr int HD = 64, LDS = FqSmem::LDS;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  T* sQ = reinterpret_cast<T*>(smem_attn);
  T* sK = sQ + FQ_ROWS * LDS;
  T* sV = sK + FqSmem::RING;
  float* part = reinterpret_cast<float*>(sK);
  const int split = blockIdx.x, h = blockIdx.y;
  const int n_rb = (p.Lq + FQ_ROWS - 1) / FQ_ROWS, b = blockIdx.z / n_rb, q0 = (blockIdx.z % n_rb) * FQ_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const T* gq = static_cast<const T*>(p.q) + b * p.q_bs + static_cast<long long>(h) * p.hd;
  const T* gk = static_cast<const T*>(p.k) + (b / p.kv_div) * p.k_bs + static_cast<long long>(h) * p.hd;
  const T* gv = static_cast<const T*>(p.v) + (b / p.kv_div) * p.v_bs + static_cast<long long>(h) * p.hd;
  ScoreCtx sc{p.scale, p.bias ? p.bias + static_cast<long long>(h) * p.bias_len : nullptr, p.bias_zero,
              p.kmask ? p.kmask + static_cast<long long>(b / p.kv_div) * p.Lk : nullptr, p.causal, p.q_pos0, p.Lk, p.Lq};
  const int n_kv = (p.Lk + BKV - 1) / BKV, per = (n_kv + FQ_SPLIT - 1) / FQ_SPLIT;
  const int t0 = split * per, n_t = max(0, min(n_kv, t0 + per) - t0);

  auto load_kv = [&](int i) {              // tile i of this CTA's share -> ring slot i % FQ_ST (one commit group per call, possibly empty)
    if (i < n_t) {
      load_tile<T, HD, BKV>(smem_u32(sK + (i % FQ_ST) * BKV * LDS), gk, p.k_rs, (t0 + i) * BKV, p.Lk, p.hd);
      load_tile<T, HD, BKV>(smem_u32(sV + (i % FQ_ST) * BKV * LDS), gv, p.v_rs, (t0 + i) * BKV, p.Lk, p.hd);
    }
    cp_async_commit();
  };
  load_tile<T, HD, FQ_ROWS>(smem_u32(sQ), gq, p.q_rs, q0, p.Lq, p.hd);
  load_kv(0);
  load_kv(1);

  float o_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[HD / 16][4];
  const float LOG2E = 1.4426950408889634f;
  uint32_t dkey = 0, drow[2] = {0u, 0u}, dthr = 0;
  float dscale = 1.f;
  if constexpr (DROP) {
    dkey = drop_key(*p.drop_seed, p.drop_site);
    dthr = p.drop_thr; dscale = p.drop_scale;
#pragma unroll
    for (int r = 0; r < 2; ++r)
      drow[r] = (static_cast<uint32_t>(b * p.H + h) * static_cast<uint32_t>(p.Lq) +
                 static_cast<uint32_t>(min(q0 + g + 8 * r, p.Lq - 1))) * drop_groups(static_cast<uint32_t>(p.Lk));
  }

  for (int i = 0; i < n_t; ++i) {
    load_kv(i + 2);
    cp_async_wait<2>();
    __syncthreads();
    if (i == 0) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) ldsm_x4(qf[kk], smem_u32(sQ + (lane & 15) * LDS + kk * 16 + (lane >> 4) * 8));
    }
    const T* cK = sK + (i % FQ_ST) * BKV * LDS;
    const T* cV = sV + (i % FQ_ST) * BKV * LDS;
    // S = Q K^T for this warp's 16 keys of the tile
    float s[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      uint32_t kf[4];
      ldsm_x4(kf, smem_u32(cK + (warp * 16 + (lane >> 4) * 8 + (lane & 7)) * LDS + kk * 16 + ((lane >> 3) & 1) * 8));
      MmaType<T>::mma(s[0], qf[kk], kf[0], kf[1]);
      MmaType<T>::mma(s[1], qf[kk], kf[2], kf[3]);
    }
    const int i0 = q0 + g, jw = (t0 + i) * BKV + warp * 16;
    float m_new[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      const int j = jw + n * 8 + 2 * t4;
      s[n][0] = sc.apply(s[n][0], i0, j);
      s[n][1] = sc.apply(s[n][1], i0, j + 1);
      s[n][2] = sc.apply(s[n][2], i0 + 8, j);
      s[n][3] = sc.apply(s[n][3], i0 + 8, j + 1);
      m_new[0] = fmaxf(m_new[0], fmaxf(s[n][0], s[n][1]));
      m_new[1] = fmaxf(m_new[1], fmaxf(s[n][2], s[n][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 1));
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 2));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      msc[r] = (m_new[r] == -INFINITY) ? 0.f : m_new[r] * LOG2E;
      corr[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f(m_run[r] * LOG2E - msc[r]);
      m_run[r] = m_new[r];
      l_run[r] *= corr[r];
    }
    uint32_t pf[4];
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      const float p0 = exp2f(s[n][0] * LOG2E - msc[0]), p1 = exp2f(s[n][1] * LOG2E - msc[0]);
      const float p2 = exp2f(s[n][2] * LOG2E - msc[1]), p3 = exp2f(s[n][3] * LOG2E - msc[1]);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      if constexpr (DROP) {                      // the row sums stay those of the undropped P
        const int j = jw + n * 8 + 2 * t4;
        const uint32_t w0 = drop_word(dkey, drow[0], j >> 2), w1 = drop_word(dkey, drow[1], j >> 2);
        pf[2 * n] = MmaType<T>::pack(drop_keep(w0, j, dthr) ? p0 : 0.f, drop_keep(w0, j + 1, dthr) ? p1 : 0.f);
        pf[2 * n + 1] = MmaType<T>::pack(drop_keep(w1, j, dthr) ? p2 : 0.f, drop_keep(w1, j + 1, dthr) ? p3 : 0.f);
      } else {
        pf[2 * n] = MmaType<T>::pack(p0, p1);
        pf[2 * n + 1] = MmaType<T>::pack(p2, p3);
      }
    }
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) {
      o_acc[d][0] *= corr[0]; o_acc[d][1] *= corr[0];
      o_acc[d][2] *= corr[1]; o_acc[d][3] *= corr[1];
    }
    // O += P V over the same 16 keys
#pragma unroll
    for (int db = 0; db < HD / 16; ++db) {
      uint32_t vf[4];
      ldsm_x4_t(vf, smem_u32(cV + (warp * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LDS + db * 16 + (lane >> 4) * 8));
      MmaType<T>::mma(o_acc[2 * db], pf, vf[0], vf[1]);
      MmaType<T>::mma(o_acc[2 * db + 1], pf, vf[2], vf[3]);
    }
    __syncthreads();                             // the slot is refilled by the next iteration's load_kv
  }
  cp_async_wait<0>();
  __syncthreads();
  // ---- warp partials -> shared memory (the ring is dead)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    const int row = g + 8 * r;
    if (t4 == 0) {
      part[FqSmem::W_M + warp * FQ_ROWS + row] = m_run[r];
      part[FqSmem::W_L + warp * FQ_ROWS + row] = l_run[r];
    }
#pragma unroll
    for (int d = 0; d < HD / 8; ++d)
      *reinterpret_cast<float2*>(part + FqSmem::W_O + (warp * FQ_ROWS + row) * 64 + d * 8 + 2 * t4) = make_float2(o_acc[d][2 * r], o_acc[d][2 * r + 1]);
  }
  __syncthreads();
  // ---- CTA partial: thread = (row, 8 columns); softmax partials merge with exp2((m_w - M) log2 e) weights
  const int row = threadIdx.x >> 3, c0 = (threadIdx.x & 7) * 8;
  {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) M = fmaxf(M, part[FqSmem::W_M + w * FQ_ROWS + row]);
    float L = 0.f, acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float mw = part[FqSmem::W_M + w * FQ_ROWS + row];
      const float f = (mw == -INFINITY) ? 0.f : exp2f((mw - M) * LOG2E);
      L += f * part[FqSmem::W_L + w * FQ_ROWS + row];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] += f * part[FqSmem::W_O + (w * FQ_ROWS + row) * 64 + c0 + c];
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) part[FqSmem::C_O + row * 64 + c0 + c] = acc[c];
    if ((threadIdx.x & 7) == 0) { part[FqSmem::C_M + row] = M; part[FqSmem::C_L + row] = L; }
  }
  fq_cluster_sync();
  // ---- cluster: CTA 0 merges the FQ_SPLIT partials in rank order and writes O / lse
  if (split == 0) {
    float M = -INFINITY;
#pragma unroll
    for (int sp = 0; sp < FQ_SPLIT; ++sp) M = fmaxf(M, fq_ld_part(part + FqSmem::C_M + row, sp));
    float L = 0.f, acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll
    for (int sp = 0; sp < FQ_SPLIT; ++sp) {
      const float ms = fq_ld_part(part + FqSmem::C_M + row, sp);
      const float f = (ms == -INFINITY) ? 0.f : exp2f((ms - M) * LOG2E);
      L += f * fq_ld_part(part + FqSmem::C_L + row, sp);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] += f * fq_ld_part(part + FqSmem::C_O + row * 64 + c0 + c, sp);
    }
    const int i = q0 + row;
    if (i < p.Lq) {
      const float inv = (L > 0.f ? 1.f / L : 0.f) * dscale;
      T* go = static_cast<T*>(p.o) + b * p.o_bs + static_cast<long long>(i) * p.o_rs + static_cast<long long>(h) * p.hd + c0;
      *reinterpret_cast<uint4*>(go) = make_uint4(MmaType<T>::pack(acc[0] * inv, acc[1] * inv), MmaType<T>::pack(acc[2] * inv, acc[3] * inv),
                                                 MmaType<T>::pack(acc[4] * inv, acc[5] * inv), MmaType<T>::pack(acc[6] * inv, acc[7] * inv));
      if (p.lse && (threadIdx.x & 7) == 0) p.lse[(static_cast<long long>(b) * p.H + h) * p.Lq + i] = M + logf(L);
    }
  }
  fq_cluster_sync();                            // the other CTAs' shared memory stays alive until CTA 0 has read it
}

// dQ of the same shape: dQ_i = scale * sum_j P_ij (dP_ij - delta_i) K_j, keys split as in the forward; partial dQ tiles are summed.
template <typename T, bool DROP = false>
__global__ void MRB_FQ_CLUSTER __launch_bounds__(NTHREADS) attn_fq_dq_kernel(const typename AttnParamsOf<DROP>::type p) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  constexpr int HD = 64, LDS = FqSmem::LDS;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  T* sQ = reinterpret_cast<T*>(smem_attn);
  T* sdO = sQ + FQ_ROWS * LDS;
  T* sK = sdO + FQ_ROWS * LDS;
  T* sV = sK + FqSmem::RING;
  float* part = reinterpret_cast<float*>(sK);
  const int split = blockIdx.x, h = blockIdx.y;
  const int n_rb = (p.Lq + FQ_ROWS - 1) / FQ_ROWS, b = blockIdx.z / n_rb, q0 = (blockIdx.z % n_rb) * FQ_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const T* gq = static_cast<const T*>(p.q) + b * p.q_bs + static_cast<long long>(h) * p.hd;
  const T* gk = static_cast<const T*>(p.k) + (b / p.kv_div) * p.k_bs + static_cast<long long>(h) * p.hd;
  const T* gv = static_cast<const T*>(p.v) + (b / p.kv_div) * p.v_bs + static_cast<long long>(h) * p.hd;
  const T* gdo = static_cast<const T*>(p.dout) + b * p.do_bs + static_cast<long long>(h) * p.hd;
  ScoreCtx sc{p.scale, p.bias ? p.bias + static_cast<long long>(h) * p.bias_len : nullptr, p.bias_zero,
              p.kmask ? p.kmask + static_cast<long long>(b / p.kv_div) * p.Lk : nullptr, p.causal, p.q_pos0, p.Lk, p.Lq};
  const int n_kv = (p.Lk + BKV - 1) / BKV, per = (n_kv + FQ_SPLIT - 1) / FQ_SPLIT;
  const int t0 = split * per, n_t = max(0, min(n_kv, t0 + per) - t0);

  auto load_kv = [&](int i) {
    if (i < n_t) {
      load_tile<T, HD, BKV>(smem_u32(sK + (i % FQ_ST) * BKV * LDS), gk, p.k_rs, (t0 + i) * BKV, p.Lk, p.hd);
      load_tile<T, HD, BKV>(smem_u32(sV + (i % FQ_ST) * BKV * LDS), gv, p.v_rs, (t0 + i) * BKV, p.Lk, p.hd);
    }
    cp_async_commit();
  };
  load_tile<T, HD, FQ_ROWS>(smem_u32(sQ), gq, p.q_rs, q0, p.Lq, p.hd);
  load_tile<T, HD, FQ_ROWS>(smem_u32(sdO), gdo, p.do_rs, q0, p.Lq, p.hd);
  load_kv(0);
  load_kv(1);

  const long long stat = (static_cast<long long>(b) * p.H + h) * p.Lq;
  float lse[2], dl[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = q0 + g + r * 8;
    lse[r] = i < p.Lq ? p.lse[stat + i] : 0.f;
    dl[r] = i < p.Lq ? p.delta[stat + i] : 0.f;
  }
  float dq_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) dq_acc[i][0] = dq_acc[i][1] = dq_acc[i][2] = dq_acc[i][3] = 0.f;
  uint32_t qf[HD / 16][4], dof[HD / 16][4];
  uint32_t dkey = 0, drow[2] = {0u, 0u}, dthr = 0;
  float dscale = 1.f;
  if constexpr (DROP) {
    dkey = drop_key(*p.drop_seed, p.drop_site);
    dthr = p.drop_thr; dscale = p.drop_scale;
#pragma unroll
    for (int r = 0; r < 2; ++r)
      drow[r] = (static_cast<uint32_t>(b * p.H + h) * static_cast<uint32_t>(p.Lq) +
                 static_cast<uint32_t>(min(q0 + g + 8 * r, p.Lq - 1))) * drop_groups(static_cast<uint32_t>(p.Lk));
  }

  for (int i = 0; i < n_t; ++i) {
    load_kv(i + 2);
    cp_async_wait<2>();
    __syncthreads();
    if (i == 0) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        ldsm_x4(qf[kk], smem_u32(sQ + (lane & 15) * LDS + kk * 16 + (lane >> 4) * 8));
        ldsm_x4(dof[kk], smem_u32(sdO + (lane & 15) * LDS + kk * 16 + (lane >> 4) * 8));
      }
    }
    const T* cK = sK + (i % FQ_ST) * BKV * LDS;
    const T* cV = sV + (i % FQ_ST) * BKV * LDS;
    float s[2][4], dp[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
      dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      uint32_t kf[4], vf[4];
      const int off = (warp * 16 + (lane >> 4) * 8 + (lane & 7)) * LDS + kk * 16 + ((lane >> 3) & 1) * 8;
      ldsm_x4(kf, smem_u32(cK + off));
      ldsm_x4(vf, smem_u32(cV + off));
      MmaType<T>::mma(s[0], qf[kk], kf[0], kf[1]);
      MmaType<T>::mma(s[1], qf[kk], kf[2], kf[3]);
      MmaType<T>::mma(dp[0], dof[kk], vf[0], vf[1]);
      MmaType<T>::mma(dp[1], dof[kk], vf[2], vf[3]);
    }
    const int i0 = q0 + g, jw = (t0 + i) * BKV + warp * 16;
    uint32_t dsf[4];
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      const int j = jw + n * 8 + 2 * t4;
      float ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const float sv = sc.apply(s[n][e], i0 + r * 8, j + (e & 1));
        const float pr = (sv == -INFINITY) ? 0.f : __expf(sv - lse[r]);
        float dpe = dp[n][e];
        if constexpr (DROP) dpe = drop_keep(drop_word(dkey, drow[r], j >> 2), j + (e & 1), dthr) ? dpe * dscale : 0.f;
        ds[e] = pr * (dpe - dl[r]) * p.scale;
      }
      dsf[2 * n] = MmaType<T>::pack(ds[0], ds[1]);
      dsf[2 * n + 1] = MmaType<T>::pack(ds[2], ds[3]);
    }
    // dQ += dS K over this warp's 16 keys (B operand = K [key x d] -> transposed ldmatrix)
#pragma unroll
    for (int db = 0; db < HD / 16; ++db) {
      uint32_t kf[4];
      ldsm_x4_t(kf, smem_u32(cK + (warp * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LDS + db * 16 + (lane >> 4) * 8));
      MmaType<T>::mma(dq_acc[2 * db], dsf, kf[0], kf[1]);
      MmaType<T>::mma(dq_acc[2 * db + 1], dsf, kf[2], kf[3]);
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = g + 8 * r;
#pragma unroll
    for (int d = 0; d < HD / 8; ++d)
      *reinterpret_cast<float2*>(part + FqSmem::W_O + (warp * FQ_ROWS + row) * 64 + d * 8 + 2 * t4) = make_float2(dq_acc[d][2 * r], dq_acc[d][2 * r + 1]);
  }
  __syncthreads();
  const int row = threadIdx.x >> 3, c0 = (threadIdx.x & 7) * 8;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) a += part[FqSmem::W_O + (w * FQ_ROWS + row) * 64 + c0 + c];
    part[FqSmem::C_O + row * 64 + c0 + c] = a;
  }
  fq_cluster_sync();
  if (split == 0) {
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      acc[c] = 0.f;
#pragma unroll
      for (int sp = 0; sp < FQ_SPLIT; ++sp) acc[c] += fq_ld_part(part + FqSmem::C_O + row * 64 + c0 + c, sp);
    }
    const int i = q0 + row;
    if (i < p.Lq) {
      T* gdq = static_cast<T*>(p.dq) + b * p.q_bs + static_cast<long long>(i) * p.q_rs + static_cast<long long>(h) * p.hd + c0;
      *reinterpret_cast<uint4*>(gdq) = make_uint4(MmaType<T>::pack(acc[0], acc[1]), MmaType<T>::pack(acc[2], acc[3]),
                                                  MmaType<T>::pack(acc[4], acc[5]), MmaType<T>::pack(acc[6], acc[7]));
    }
  }
  fq_cluster_sync();
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  return e == cudaSuccess ? MRB_OK : mrb_set_error(e);
}

template <typename T, int HD, bool DROP = false>
static int launch_fwd(const typename AttnParamsOf<DROP>::type& p, cudaStream_t s) {
  const int smem = (BQ + 4 * BKV) * (HD + 8) * 2;
  static bool cfg = false;
  if (!cfg) { if (int rc = set_smem(attn_fwd_kernel<T, HD, DROP>, smem)) return rc; cfg = true; }
  dim3 grid((p.Lq + BQ - 1) / BQ, p.H, p.B);
  MRB_LAUNCH((attn_fwd_kernel<T, HD, DROP>), grid, NTHREADS, smem, s, p);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

template <typename T, bool DROP, int XW>
static int launch_xq_w(const typename AttnParamsOf<DROP>::type& p, cudaStream_t s) {
  const int LkP = (p.Lk + 15) & ~15;
  const int kv = 2 * LkP * 64 * 2, part = XW * 32 * (64 + 2) * 4 + 2 * XW * 32 * 4;     // K + V, reused for the warps' partials
  const int smem = 32 * 64 * 2 + (kv > part ? kv : part);
  static int cfg = 0;
  if (cfg < smem) { if (int rc = set_smem(attn_xq_kernel<T, DROP, XW>, smem)) return rc; cfg = smem; }
  MRB_LAUNCH((attn_xq_kernel<T, DROP, XW>), dim3(p.H, p.B), 32 * XW, smem, s, p);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

template <typename T, bool DROP = false>
static int launch_xq(const typename AttnParamsOf<DROP>::type& p, cudaStream_t s) {
  static int warps = 0;                   // MRB_XQ_WARPS=8: the eight-warp form (A/B measurements)
  if (!warps) { const char* e = getenv("MRB_XQ_WARPS"); warps = (e && e[0] == '8') ? 8 : 4; }
  return warps == 4 ? launch_xq_w<T, DROP, 4>(p, s) : launch_xq_w<T, DROP, 8>(p, s);
}

// few query rows x >= 512 keys (the T5 decoder's cross-attention), hd = 64
static inline bool fq_shape(int Lq, int Lk, int hd) {
  static int use = -1;                    // MRB_ATTN_FQ=0 keeps the generic kernels (A/B measurements)
  if (use < 0) { const char* e = getenv("MRB_ATTN_FQ"); use = (e && e[0] == '0') ? 0 : 1; }
  return use && hd == 64 && Lq <= 2 * FQ_ROWS && Lk >= 512;
}
template <typename T, bool DROP = false>
static int launch_fq_fwd(const typename AttnParamsOf<DROP>::type& p, cudaStream_t s) {
  const int smem = FqSmem::bytes(1);
  static bool cfg = false;
  if (!cfg) { if (int rc = set_smem(attn_fq_fwd_kernel<T, DROP>, smem)) return rc; cfg = true; }
  dim3 grid(FQ_SPLIT, p.H, p.B * ((p.Lq + FQ_ROWS - 1) / FQ_ROWS));
  MRB_LAUNCH((attn_fq_fwd_kernel<T, DROP>), grid, NTHREADS, smem, s, p);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
template <typename T, bool DROP = false>
static int launch_fq_dq(const typename AttnParamsOf<DROP>::type& p, cudaStream_t s) {
  const int smem = FqSmem::bytes(2);
  static bool cfg = false;
  if (!cfg) { if (int rc = set_smem(attn_fq_dq_kernel<T, DROP>, smem)) return rc; cfg = true; }
  dim3 grid(FQ_SPLIT, p.H, p.B * ((p.Lq + FQ_ROWS - 1) / FQ_ROWS));
  MRB_LAUNCH((attn_fq_dq_kernel<T, DROP>), grid, NTHREADS, smem, s, p);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

template <typename T, int HD, bool DROP = false>
static int launch_bwd(const typename AttnParamsOf<DROP>::type& p, cudaStream_t s) {
  if (HD == 64 && p.Lq <= DELTA_EXACT_MAX_LQ) {
    // few query rows (T5 decoder): delta = sum_j P_ij dP_ij recomputed exactly, see attn_delta.cuh
    DeltaExactParams d{};
    d.q = static_cast<const uint16_t*>(p.q); d.k = static_cast<const uint16_t*>(p.k); d.v = static_cast<const uint16_t*>(p.v);
    d.dout = static_cast<const uint16_t*>(p.dout);
    d.q_bs = p.q_bs; d.q_rs = p.q_rs; d.k_bs = p.k_bs; d.k_rs = p.k_rs; d.v_bs = p.v_bs; d.v_rs = p.v_rs; d.do_bs = p.do_bs; d.do_rs = p.do_rs;
    d.B = p.B; d.H = p.H; d.Lq = p.Lq; d.Lk = p.Lk; d.dtype = sizeof(T) == 2 && std::is_same<T, __half>::value ? MRB_DT_F16 : MRB_DT_BF16;
    d.scale = p.scale; d.bias = p.bias; d.bias_len = p.bias_len; d.bias_zero = p.bias_zero; d.kmask = p.kmask; d.causal = p.causal;
    d.q_pos0 = p.q_pos0; d.lse = p.lse; d.delta = const_cast<float*>(p.delta);
    if constexpr (DROP) { d.drop_seed = p.drop_seed; d.drop_site = p.drop_site; d.drop_thr = p.drop_thr; d.drop_scale = p.drop_scale; }
    if (int rc = launch_delta_exact(d, s)) return rc;
  } else {
    const int rows = p.B * p.H * p.Lq;
    MRB_LAUNCH((attn_delta_kernel<T>), (rows + 7) / 8, 256, 0, s, static_cast<const AttnParams&>(p));
    MRB_CHECK_LAUNCH();
  }
  if (HD == 64 && fq_shape(p.Lq, p.Lk, p.hd)) {          // keys split over a cluster (attn_fq_dq_kernel)
    if (int rc = launch_fq_dq<T, DROP>(p, s)) return rc;
  } else {
    const int smem = (2 * BQ + 4 * BKV) * (HD + 8) * 2;
    static bool cfg = false;
    if (!cfg) { if (int rc = set_smem(attn_bwd_dq_kernel<T, HD, DROP>, smem)) return rc; cfg = true; }
    dim3 grid((p.Lq + BQ - 1) / BQ, p.H, p.B);
    MRB_LAUNCH((attn_bwd_dq_kernel<T, HD, DROP>), grid, NTHREADS, smem, s, p);
    MRB_CHECK_LAUNCH();
  }
  {
    const int smem = (2 * BKV + 4 * BQ) * (HD + 8) * 2 + 4 * BQ * 4;
    static bool cfg = false;
    if (!cfg) { if (int rc = set_smem(attn_bwd_dkv_kernel<T, HD, DROP>, smem)) return rc; cfg = true; }
    dim3 grid((p.Lk + BKV - 1) / BKV, p.H, p.B);
    MRB_LAUNCH((attn_bwd_dkv_kernel<T, HD, DROP>), grid, NTHREADS, smem, s, p);
    MRB_CHECK_LAUNCH();
  }
  return MRB_OK;
}

}  // namespace mrb

using namespace mrb;

static int check_attn(const AttnParams& p, int dtype) {
  if (dtype != MRB_DT_F16 && dtype != MRB_DT_BF16) return MRB_ERR_ARG;
  if (p.hd <= 0 || p.hd > 96 || (p.hd & 7)) return MRB_ERR_UNSUPPORTED;
  if ((p.q_rs | p.k_rs | p.v_rs | p.o_rs | p.q_bs | p.k_bs | p.v_bs | p.o_bs) & 7) return MRB_ERR_ARG;
  return MRB_OK;
}

static int attention_fwd_impl(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                              const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                              int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                              int bias_len, int bias_zero, const int* kmask, int kv_div, int causal, int q_pos0, float* lse,
                              const unsigned* drop_seed, unsigned drop_site, float drop_p, void* stream) {
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0) return MRB_OK;
  AttnDropParams p{};
  p.q = q; p.k = k; p.v = v; p.o = o;
  p.q_bs = q_bs; p.q_rs = q_rs; p.k_bs = k_bs; p.k_rs = k_rs; p.v_bs = v_bs; p.v_rs = v_rs; p.o_bs = o_bs; p.o_rs = o_rs;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.hd = hd; p.scale = scale;
  p.bias = bias; p.bias_len = bias_len; p.bias_zero = bias_zero; p.kmask = kmask; p.causal = causal; p.q_pos0 = q_pos0; p.kv_div = kv_div > 0 ? kv_div : 1;
  p.lse = lse;
  if (int rc = check_attn(p, dtype)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // few queries x a few hundred keys, no bias / mask (Q-Former cross-attention): one-shot K/V fetch, keys split over the warps
  static int use_xq = -1;                 // MRB_ATTN_XQ=0 keeps the generic kernel (A/B measurements)
  if (use_xq < 0) { const char* e = getenv("MRB_ATTN_XQ"); use_xq = (e && e[0] == '0') ? 0 : 1; }
  const bool xq_shape = use_xq && hd == 64 && Lq <= 32 && Lk > 64 && Lk <= XQ_MAX_LK && !bias && !kmask && !causal && !lse &&
                        p.kv_div == 1 && 4 * 32 * (64 + 2) * 4 + 8 * 32 * 4 <= 2 * ((Lk + 15) & ~15) * 64 * 2;
  const bool fq = fq_shape(Lq, Lk, hd);   // few query rows x thousands of keys (T5 decoder cross-attention): keys split over a cluster
  if (drop_seed && drop_p > 0.f) {
    if (hd > 64 || drop_p >= 1.f) return MRB_ERR_UNSUPPORTED;
    const DropSpec d = make_drop(drop_seed, drop_site, drop_p);
    p.drop_seed = d.seed; p.drop_site = d.site; p.drop_thr = d.thr; p.drop_scale = d.scale;
    if (fq) return dtype == MRB_DT_F16 ? launch_fq_fwd<__half, true>(p, s) : launch_fq_fwd<__nv_bfloat16, true>(p, s);
    if (xq_shape) return dtype == MRB_DT_F16 ? launch_xq<__half, true>(p, s) : launch_xq<__nv_bfloat16, true>(p, s);
    return dtype == MRB_DT_F16 ? launch_fwd<__half, 64, true>(p, s) : launch_fwd<__nv_bfloat16, 64, true>(p, s);
  }
  if (fq) return dtype == MRB_DT_F16 ? launch_fq_fwd<__half>(p, s) : launch_fq_fwd<__nv_bfloat16>(p, s);
  if (xq_shape) return dtype == MRB_DT_F16 ? launch_xq<__half>(p, s) : launch_xq<__nv_bfloat16>(p, s);
  if (dtype == MRB_DT_F16) return hd <= 64 ? launch_fwd<__half, 64>(p, s) : launch_fwd<__half, 96>(p, s);
  return hd <= 64 ? launch_fwd<__nv_bfloat16, 64>(p, s) : launch_fwd<__nv_bfloat16, 96>(p, s);
}

extern "C" int mrb_attention_fwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                 const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                                 int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                                 int bias_len, int bias_zero, const int* kmask, int kv_div, int causal, int q_pos0, float* lse,
                                 void* stream) {
  return attention_fwd_impl(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, B, H, Lq, Lk, hd, dtype, scale, bias, bias_len,
                            bias_zero, kmask, kv_div, causal, q_pos0, lse, nullptr, 0u, 0.f, stream);
}
// mrb_attention_fwd with train-mode dropout of the attention probabilities (hd <= 64): O = drop(softmax(S)) V, lse unchanged
extern "C" int mrb_attention_fwd_drop(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                      const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                                      int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                                      int bias_len, int bias_zero, const int* kmask, int kv_div, int causal, int q_pos0, float* lse,
                                      const unsigned* seed, unsigned site, float p, void* stream) {
  if (!seed) return MRB_ERR_ARG;
  return attention_fwd_impl(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, B, H, Lq, Lk, hd, dtype, scale, bias, bias_len,
                            bias_zero, kmask, kv_div, causal, q_pos0, lse, seed, site, p, stream);
}

static int attention_bwd_impl(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                              const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                              const void* dout, long long do_bs, long long do_rs, void* dq, void* dk, void* dv,
                              int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                              int bias_len, int bias_zero, const int* kmask, int causal, int q_pos0, const float* lse,
                              float* delta_ws, const unsigned* drop_seed, unsigned drop_site, float drop_p, void* stream) {
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0) return MRB_OK;
  AttnDropParams p{};
  p.q = q; p.k = k; p.v = v; p.o = const_cast<void*>(o);
  p.q_bs = q_bs; p.q_rs = q_rs; p.k_bs = k_bs; p.k_rs = k_rs; p.v_bs = v_bs; p.v_rs = v_rs; p.o_bs = o_bs; p.o_rs = o_rs;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.hd = hd; p.scale = scale;
  p.bias = bias; p.bias_len = bias_len; p.bias_zero = bias_zero; p.kmask = kmask; p.causal = causal; p.q_pos0 = q_pos0; p.kv_div = 1;
  p.lse = const_cast<float*>(lse); p.dout = dout; p.do_bs = do_bs; p.do_rs = do_rs; p.delta = delta_ws;
  p.dq = dq; p.dk = dk; p.dv = dv;
  if (int rc = check_attn(p, dtype)) return rc;
  if (hd > 64 || ((do_bs | do_rs) & 7)) return MRB_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (drop_seed && drop_p > 0.f) {
    if (drop_p >= 1.f) return MRB_ERR_UNSUPPORTED;
    const DropSpec d = make_drop(drop_seed, drop_site, drop_p);
    p.drop_seed = d.seed; p.drop_site = d.site; p.drop_thr = d.thr; p.drop_scale = d.scale;
    return dtype == MRB_DT_F16 ? launch_bwd<__half, 64, true>(p, s) : launch_bwd<__nv_bfloat16, 64, true>(p, s);
  }
  if (dtype == MRB_DT_F16) return launch_bwd<__half, 64>(p, s);
  return launch_bwd<__nv_bfloat16, 64>(p, s);
}

extern "C" int mrb_attention_bwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                 const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                                 const void* dout, long long do_bs, long long do_rs, void* dq, void* dk, void* dv,
                                 int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                                 int bias_len, int bias_zero, const int* kmask, int causal, int q_pos0, const float* lse,
                                 float* delta_ws, void* stream) {
  return attention_bwd_impl(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, dout, do_bs, do_rs, dq, dk, dv, B, H, Lq, Lk, hd, dtype,
                            scale, bias, bias_len, bias_zero, kmask, causal, q_pos0, lse, delta_ws, nullptr, 0u, 0.f, stream);
}
// Backward of mrb_attention_fwd_drop (same seed word, site and p: the mask is recomputed)
extern "C" int mrb_attention_bwd_drop(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                      const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                                      const void* dout, long long do_bs, long long do_rs, void* dq, void* dk, void* dv,
                                      int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                                      int bias_len, int bias_zero, const int* kmask, int causal, int q_pos0, const float* lse,
                                      float* delta_ws, const unsigned* seed, unsigned site, float p, void* stream) {
  if (!seed) return MRB_ERR_ARG;
  return attention_bwd_impl(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, dout, do_bs, do_rs, dq, dk, dv, B, H, Lq, Lk, hd, dtype,
                            scale, bias, bias_len, bias_zero, kmask, causal, q_pos0, lse, delta_ws, seed, site, p, stream);
}

// One query row per (batch, head): q/o point at that row of batch 0 (batch strides apply).  No bias / mask.
extern "C" int mrb_attention_row(const void* q, long long q_bs, const void* k, long long k_bs, long long k_rs, const void* v,
                                 long long v_bs, long long v_rs, void* o, long long o_bs, int B, int H, int Lk, int hd,
                                 int dtype, float scale, void* stream) {
  if (B <= 0 || H <= 0 || Lk <= 0) return MRB_OK;
  if ((hd & 7) || hd > 128 || Lk > 4096 || ((q_bs | k_bs | k_rs | v_bs | v_rs | o_bs) & 3)) return MRB_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int blocks = (B * H + 3) / 4, smem = 4 * Lk * 4 + 4 * 128 * 4;
  if (dtype == MRB_DT_F16)
    MRB_LAUNCH((attn_row_kernel<__half>), blocks, 128, smem, s, static_cast<const __half*>(q), q_bs, static_cast<const __half*>(k), k_bs, k_rs,
                                                     static_cast<const __half*>(v), v_bs, v_rs, static_cast<__half*>(o), o_bs, B, H, Lk, hd, scale);
  else if (dtype == MRB_DT_BF16)
    MRB_LAUNCH((attn_row_kernel<__nv_bfloat16>), blocks, 128, smem, s, static_cast<const __nv_bfloat16*>(q), q_bs, static_cast<const __nv_bfloat16*>(k), k_bs, k_rs,
                                                            static_cast<const __nv_bfloat16*>(v), v_bs, v_rs, static_cast<__nv_bfloat16*>(o), o_bs, B, H, Lk, hd, scale);
  else return MRB_ERR_ARG;
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
