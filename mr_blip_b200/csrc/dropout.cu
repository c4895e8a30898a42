// Train-mode dropout outside the attention kernels (masks: dropmask.cuh): the residual-branch / embedding / final-norm / FF-inner
// sites of T5 (modeling_t5.py:327,346,652,690,1149,1258) and of the Q-Former (Qformer.py:107,287,373) as one-pass HBM-bound
// kernels, and peft's LoRA input dropout (blip2_mr.py:197: y = W x + B A drop_j(x), an independent mask per adapted Linear),
// which breaks the "LoRA folded into the GEMM's K extension" layout of the backward pass and gets three CUDA-core kernels:
//   forward   u_j  = drop_j(x) A_j^T                          (mrb_lora_down_drop: fills the 32 extension columns of x_ext)
//   backward  dA_j += (dy sB_j)^T drop_j(x)                   (mrb_lora_wgrad_drop)
//             dx   += sum_j mask_j * ((dy sB_j) A_j)          (mrb_lora_dx_drop; the dense dy W part stays one tcgen05 GEMM over K = N)
// dB_j = dy_j^T u_j needs no change: u_j already carries the mask.  Every backward kernel recomputes its mask from
// (seed, site, row, column).
#include <stdlib.h>
#include "common.cuh"
#include "dropmask.cuh"

namespace mrb {

__device__ __forceinline__ float drop1(float v, uint32_t w, int i, const DropSpec& d) {
  return ((w >> (8 * i)) & 0xffu) >= d.thr ? v * d.scale : 0.f;
}
__device__ __forceinline__ float keep1(float v, uint32_t w, int i, uint32_t thr) {        // mask without the scale
  return ((w >> (8 * i)) & 0xffu) >= thr ? v : 0.f;
}

// ---------------------------------------------------------------- out = drop(x), 4 columns per thread
// DT: 0 = 16-bit in / 16-bit out (dtype), 1 = fp32 / fp32, 2 = fp32 in / 16-bit out (the cast that builds a dgrad operand)
template <int DT>
__global__ void __launch_bounds__(256) dropout_kernel(const void* __restrict__ x, long long ldx, void* __restrict__ out, long long ldo,
                                                      int rows, int cols, int dtype, const DropSpec d) {
  pdl_trigger();
  pdl_wait();
  const int ng = cols >> 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * ng) return;
  const int g = static_cast<int>(idx % ng);
  const long long r = idx / ng;
  const uint32_t w = drop_word(drop_key(*d.seed, d.site), static_cast<uint32_t>(r) * static_cast<uint32_t>(ng), g);
  float v[4];
  if (DT == 0) {
    const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(x) + r * ldx + 4 * g);
    v[0] = unpack_lo(u.x, dtype); v[1] = unpack_hi(u.x, dtype); v[2] = unpack_lo(u.y, dtype); v[3] = unpack_hi(u.y, dtype);
  } else {
    const float4 f = *reinterpret_cast<const float4*>(static_cast<const float*>(x) + r * ldx + 4 * g);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = drop1(v[i], w, i, d);
  if (DT == 1) {
    *reinterpret_cast<float4*>(static_cast<float*>(out) + r * ldo + 4 * g) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    *reinterpret_cast<uint2*>(static_cast<uint16_t*>(out) + r * ldo + 4 * g) =
        make_uint2(pack2(v[0], v[1], dtype), pack2(v[2], v[3], dtype));
  }
}

// out = resid + drop(branch)   (fp32 [rows, cols], contiguous):  hidden + dropout(sublayer(hidden))
__global__ void __launch_bounds__(256) dropout_add_kernel(const float4* __restrict__ resid, const float4* __restrict__ branch,
                                                          float4* __restrict__ out, int rows, int cols, const DropSpec d) {
  pdl_trigger();
  pdl_wait();
  const int ng = cols >> 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * ng) return;
  const int g = static_cast<int>(idx % ng);
  const long long r = idx / ng;
  const uint32_t w = drop_word(drop_key(*d.seed, d.site), static_cast<uint32_t>(r) * static_cast<uint32_t>(ng), g);
  const float4 a = resid[idx], b = branch[idx];
  out[idx] = make_float4(a.x + drop1(b.x, w, 0, d), a.y + drop1(b.y, w, 1, d), a.z + drop1(b.z, w, 2, d), a.w + drop1(b.w, w, 3, d));
}

// ---------------------------------------------------------------- gated GELU with the FF-inner dropout (modeling_t5.py:323-327)
// h = drop(gelu(a) * b);   backward: dh' = mask * scale * dh, then dab[:, :F] = dh' b gelu'(a), dab[:, F:] = dh' gelu(a)
__global__ void gated_gelu_fwd_drop_kernel(const uint4* __restrict__ ab, uint4* __restrict__ h, int M, int F, long long ldh, int dtype,
                                           const DropSpec d) {
  pdl_trigger();
  pdl_wait();
  const int fv = F >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(M) * fv) return;
  const int c = idx % fv;
  const long long m = idx / fv;
  const uint32_t key = drop_key(*d.seed, d.site), rowbase = static_cast<uint32_t>(m) * static_cast<uint32_t>(F >> 2);
  const uint32_t w[2] = {drop_word(key, rowbase, 2 * c), drop_word(key, rowbase, 2 * c + 1)};
  const uint4 a = ab[m * (2 * fv) + c], b = ab[m * (2 * fv) + fv + c];
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    o[i] = pack2(drop1(gelu_erf(unpack_lo(aw[i], dtype)) * unpack_lo(bw[i], dtype), w[i >> 1], (2 * i) & 3, d),
                 drop1(gelu_erf(unpack_hi(aw[i], dtype)) * unpack_hi(bw[i], dtype), w[i >> 1], (2 * i + 1) & 3, d), dtype);
  h[m * (ldh >> 3) + c] = make_uint4(o[0], o[1], o[2], o[3]);
}
__global__ void gated_gelu_bwd_drop_kernel(const uint4* __restrict__ ab, const uint4* __restrict__ dh, long long lddh,
                                           uint4* __restrict__ dab, long long lddab, int M, int F, int dtype, const DropSpec d) {
  pdl_trigger();
  pdl_wait();
  const int fv = F >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(M) * fv) return;
  const int c = idx % fv;
  const long long m = idx / fv;
  const uint32_t key = drop_key(*d.seed, d.site), rowbase = static_cast<uint32_t>(m) * static_cast<uint32_t>(F >> 2);
  const uint32_t w[2] = {drop_word(key, rowbase, 2 * c), drop_word(key, rowbase, 2 * c + 1)};
  const uint4 a = ab[m * (2 * fv) + c], b = ab[m * (2 * fv) + fv + c], dv = dh[m * (lddh >> 3) + c];
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
  uint32_t oa[4], ob[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a0 = unpack_lo(aw[i], dtype), a1 = unpack_hi(aw[i], dtype);
    const float b0 = unpack_lo(bw[i], dtype), b1 = unpack_hi(bw[i], dtype);
    const float d0 = drop1(unpack_lo(dw[i], dtype), w[i >> 1], (2 * i) & 3, d);
    const float d1 = drop1(unpack_hi(dw[i], dtype), w[i >> 1], (2 * i + 1) & 3, d);
    float g0, g1, dg0, dg1;
    gelu_erf_both(a0, g0, dg0);
    gelu_erf_both(a1, g1, dg1);
    oa[i] = pack2(d0 * b0 * dg0, d1 * b1 * dg1, dtype);
    ob[i] = pack2(d0 * g0, d1 * g1, dtype);
  }
  dab[m * (lddab >> 3) + c] = make_uint4(oa[0], oa[1], oa[2], oa[3]);
  dab[m * (lddab >> 3) + fv + c] = make_uint4(ob[0], ob[1], ob[2], ob[3]);
}

// ---------------------------------------------------------------- LoRA forward: u[m, 8j + r] = scale * sum_k keep_j(m,k) x[m,k] A[8j + r, k]
// x 16-bit [M, K] (ldx), A 16-bit [8 NL (.. 32), K] (lda) = the group's stacked lora_A, out 16-bit [M, 32] (ldo; columns >= 8 NL are
// written as zeros: they are the unused K-extension columns of x_ext).  LANES threads share a row: lane l takes the 16-byte
// chunks l, l + LANES, ...; A is staged per 256-column tile in shared memory as fp32, chunk-major with a 16-byte pad so that the
// 8 lanes of a quarter warp (8 consecutive chunks) read conflict-free; the 8 NL partial sums meet in xor shuffles.
template <int NL, int LANES>
__global__ void __launch_bounds__(128) lora_down_drop_kernel(const uint16_t* __restrict__ x, long long ldx, const uint16_t* __restrict__ A,
                                                             long long lda, int M, int K, uint16_t* __restrict__ out, long long ldo,
                                                             int dtype, const DropSpec d) {
  pdl_trigger();
  pdl_wait();
  constexpr int R = 8 * NL, KT = 256, CH = KT / 8, AS = R * 8 + 4;
  __shared__ __align__(16) float A_s[CH * AS];
  const int lane_k = threadIdx.x % LANES, row_in = threadIdx.x / LANES;
  const int m = blockIdx.x * (128 / LANES) + row_in;
  const bool live = m < M;
  uint32_t key[NL];
#pragma unroll
  for (int j = 0; j < NL; ++j) key[j] = drop_key(*d.seed, d.site + j);
  const uint32_t rowbase = static_cast<uint32_t>(m) * static_cast<uint32_t>(K >> 2);
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  for (int k0 = 0; k0 < K; k0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < R * CH; i += 128) {            // one 16-byte chunk of one A row per step (coalesced along k)
      const int r = i / CH, c = i % CH;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (k0 + c * 8 < K) v = *reinterpret_cast<const uint4*>(A + static_cast<long long>(r) * lda + k0 + c * 8);
      float* dst = A_s + c * AS + r * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(unpack_lo(v.x, dtype), unpack_hi(v.x, dtype), unpack_lo(v.y, dtype), unpack_hi(v.y, dtype));
      *reinterpret_cast<float4*>(dst + 4) = make_float4(unpack_lo(v.z, dtype), unpack_hi(v.z, dtype), unpack_lo(v.w, dtype), unpack_hi(v.w, dtype));
    }
    __syncthreads();
    if (live) {
      for (int c = lane_k; c < CH && k0 + c * 8 < K; c += LANES) {
        const int k = k0 + c * 8;
        const uint4 xv = *reinterpret_cast<const uint4*>(x + static_cast<long long>(m) * ldx + k);
        const float xf[8] = {unpack_lo(xv.x, dtype), unpack_hi(xv.x, dtype), unpack_lo(xv.y, dtype), unpack_hi(xv.y, dtype),
                             unpack_lo(xv.z, dtype), unpack_hi(xv.z, dtype), unpack_lo(xv.w, dtype), unpack_hi(xv.w, dtype)};
        const float* as = A_s + c * AS;
#pragma unroll
        for (int j = 0; j < NL; ++j) {
          const uint32_t w0 = drop_word(key[j], rowbase, k >> 2), w1 = drop_word(key[j], rowbase, (k >> 2) + 1);
          float xm[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) { xm[i] = keep1(xf[i], w0, i, d.thr); xm[4 + i] = keep1(xf[4 + i], w1, i, d.thr); }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float4 a0 = *reinterpret_cast<const float4*>(as + (8 * j + r) * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(as + (8 * j + r) * 8 + 4);
            acc[8 * j + r] += xm[0] * a0.x + xm[1] * a0.y + xm[2] * a0.z + xm[3] * a0.w + xm[4] * a1.x + xm[5] * a1.y + xm[6] * a1.z +
                              xm[7] * a1.w;
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
  }
  if (live && lane_k == 0) {
    uint32_t o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = 2 * i < R ? pack2(acc[(2 * i) % R] * d.scale, acc[(2 * i + 1) % R] * d.scale, dtype) : 0u;
    uint4* op = reinterpret_cast<uint4*>(out + static_cast<long long>(m) * ldo);
#pragma unroll
    for (int i = 0; i < 4; ++i) op[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
  }
}

// ---------------------------------------------------------------- LoRA backward, weight gradient: dA[r, k] += scale * sum_m keep(m,k) x[m,k] q[m,r]
// x 16-bit [M, K] (ldx), q 16-bit [M, 8] (ldq) = dy sB_j, dA fp32 [8, K].  Thread = (8 columns, row slice), block = 256 columns x
// rows_per_block rows; row slices meet in shared-memory atomics, then one global atomic per output element and block.
__global__ void __launch_bounds__(256) lora_wgrad_drop_kernel(const uint16_t* __restrict__ x, long long ldx, const uint16_t* __restrict__ q,
                                                              long long ldq, int M, int K, float* __restrict__ dA, int dtype,
                                                              int rows_per_block, const DropSpec d) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[64 * 32];
  const int cg = threadIdx.x & 31, rs = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + cg * 8;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  const uint32_t key = drop_key(*d.seed, d.site), ng = static_cast<uint32_t>(K >> 2);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[i][r] = 0.f;
  for (int i = threadIdx.x; i < 64 * 32; i += 256) red[i] = 0.f;
  if (c0 < K) {
#pragma unroll 2
    for (int m = m0 + rs; m < m1; m += 8) {
      const uint4 pv = *reinterpret_cast<const uint4*>(x + static_cast<long long>(m) * ldx + c0);
      const uint4 qv = *reinterpret_cast<const uint4*>(q + static_cast<long long>(m) * ldq);
      const uint32_t rowbase = static_cast<uint32_t>(m) * ng;
      const uint32_t w0 = drop_word(key, rowbase, c0 >> 2), w1 = drop_word(key, rowbase, (c0 >> 2) + 1);
      const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w}, qw[4] = {qv.x, qv.y, qv.z, qv.w};
      float pf[8], qf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pf[2 * i] = keep1(unpack_lo(pw[i], dtype), i < 2 ? w0 : w1, (2 * i) & 3, d.thr);
        pf[2 * i + 1] = keep1(unpack_hi(pw[i], dtype), i < 2 ? w0 : w1, (2 * i + 1) & 3, d.thr);
        qf[2 * i] = unpack_lo(qw[i], dtype); qf[2 * i + 1] = unpack_hi(qw[i], dtype);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[i][r] = fmaf(pf[i], qf[r], acc[i][r]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int r = 0; r < 8; ++r) atomicAdd(&red[(i * 8 + r) * 32 + cg], acc[i][r]);
  __syncthreads();
  for (int idx = threadIdx.x; idx < 64 * 32; idx += 256) {
    const int g = idx & 31, ir = idx >> 5, i = ir >> 3, r = ir & 7;
    const int c = blockIdx.x * 256 + g * 8 + i;
    if (c < K) atomicAdd(dA + static_cast<long long>(r) * K + c, red[idx] * d.scale);
  }
}

// ---------------------------------------------------------------- LoRA backward, input gradient: dx[m,k] += scale * sum_j keep_j(m,k) sum_r q[m, 8j+r] A[8j+r, k]
// q 16-bit [M, >= 8 NL] (ldq) = the extension columns dy sB of the dgrad operand, A 16-bit [8 NL, K] (lda), dx 16-bit or fp32 [M, K]
// (read-modify-write after the dense dgrad GEMM).  Block = 256 columns x rows_per_block rows, A tile in shared memory as fp32 split
// into the low / high four columns of every 8-column group (lane stride 16 bytes: conflict-free); warp = one row at a time.
template <int NL, bool F32OUT>
__global__ void __launch_bounds__(256) lora_dx_drop_kernel(const uint16_t* __restrict__ q, long long ldq, const uint16_t* __restrict__ A,
                                                           long long lda, void* __restrict__ dx, long long lddx, int M, int K, int dtype,
                                                           int rows_per_block, const DropSpec d) {
  pdl_trigger();
  pdl_wait();
  constexpr int R = 8 * NL;
  __shared__ __align__(16) float A_s[R * 2 * 32 * 4];
  const int cg = threadIdx.x & 31, rs = threadIdx.x >> 5;
  const int cb = blockIdx.x * 256, c0 = cb + cg * 8;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  for (int i = threadIdx.x; i < R * 32; i += 256) {
    const int r = i >> 5, g = i & 31;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (cb + g * 8 < K) v = *reinterpret_cast<const uint4*>(A + static_cast<long long>(r) * lda + cb + g * 8);
    *reinterpret_cast<float4*>(A_s + ((r * 2 + 0) * 32 + g) * 4) =
        make_float4(unpack_lo(v.x, dtype), unpack_hi(v.x, dtype), unpack_lo(v.y, dtype), unpack_hi(v.y, dtype));
    *reinterpret_cast<float4*>(A_s + ((r * 2 + 1) * 32 + g) * 4) =
        make_float4(unpack_lo(v.z, dtype), unpack_hi(v.z, dtype), unpack_lo(v.w, dtype), unpack_hi(v.w, dtype));
  }
  __syncthreads();
  if (c0 >= K) return;
  uint32_t key[NL];
#pragma unroll
  for (int j = 0; j < NL; ++j) key[j] = drop_key(*d.seed, d.site + j);
  const uint32_t ng = static_cast<uint32_t>(K >> 2);
  for (int m = m0 + rs; m < m1; m += 8) {
    const uint32_t rowbase = static_cast<uint32_t>(m) * ng;
    float sum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const uint4 qv = *reinterpret_cast<const uint4*>(q + static_cast<long long>(m) * ldq + 8 * j);      // same address in every lane
      const float qf[8] = {unpack_lo(qv.x, dtype), unpack_hi(qv.x, dtype), unpack_lo(qv.y, dtype), unpack_hi(qv.y, dtype),
                           unpack_lo(qv.z, dtype), unpack_hi(qv.z, dtype), unpack_lo(qv.w, dtype), unpack_hi(qv.w, dtype)};
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 lo = *reinterpret_cast<const float4*>(A_s + (((8 * j + r) * 2 + 0) * 32 + cg) * 4);
        const float4 hi = *reinterpret_cast<const float4*>(A_s + (((8 * j + r) * 2 + 1) * 32 + cg) * 4);
        t[0] = fmaf(qf[r], lo.x, t[0]); t[1] = fmaf(qf[r], lo.y, t[1]); t[2] = fmaf(qf[r], lo.z, t[2]); t[3] = fmaf(qf[r], lo.w, t[3]);
        t[4] = fmaf(qf[r], hi.x, t[4]); t[5] = fmaf(qf[r], hi.y, t[5]); t[6] = fmaf(qf[r], hi.z, t[6]); t[7] = fmaf(qf[r], hi.w, t[7]);
      }
      const uint32_t w0 = drop_word(key[j], rowbase, c0 >> 2), w1 = drop_word(key[j], rowbase, (c0 >> 2) + 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) { sum[i] += keep1(t[i], w0, i, d.thr); sum[4 + i] += keep1(t[4 + i], w1, i, d.thr); }
    }
    if (F32OUT) {
      float4* p = reinterpret_cast<float4*>(static_cast<float*>(dx) + static_cast<long long>(m) * lddx + c0);
      float4 a = p[0], b = p[1];
      a.x += sum[0] * d.scale; a.y += sum[1] * d.scale; a.z += sum[2] * d.scale; a.w += sum[3] * d.scale;
      b.x += sum[4] * d.scale; b.y += sum[5] * d.scale; b.z += sum[6] * d.scale; b.w += sum[7] * d.scale;
      p[0] = a; p[1] = b;
    } else {
      uint4* p = reinterpret_cast<uint4*>(static_cast<uint16_t*>(dx) + static_cast<long long>(m) * lddx + c0);
      const uint4 v = *p;
      *p = make_uint4(pack2(unpack_lo(v.x, dtype) + sum[0] * d.scale, unpack_hi(v.x, dtype) + sum[1] * d.scale, dtype),
                      pack2(unpack_lo(v.y, dtype) + sum[2] * d.scale, unpack_hi(v.y, dtype) + sum[3] * d.scale, dtype),
                      pack2(unpack_lo(v.z, dtype) + sum[4] * d.scale, unpack_hi(v.z, dtype) + sum[5] * d.scale, dtype),
                      pack2(unpack_lo(v.w, dtype) + sum[6] * d.scale, unpack_hi(v.w, dtype) + sum[7] * d.scale, dtype));
    }
  }
}


// ================================================================ tensor-core versions of the three LoRA-dropout kernels
// The CUDA-core kernels above spend 24 FMAs + the mask arithmetic per element and ran at 77 / 39 / 35 us per encoder-sized call
// (28.5 ms of a 283 ms step: profiles/launch_summary_r02a_dropout.csv).  The products below are rank-8 contractions, i.e. tensor
// core work (mma.sync m16n8k16 -- three Linears x 8 ranks is far too thin a problem for a tcgen05 tile); what stays on the CUDA
// cores is the mask itself: one hash per four elements and Linear, applied to PACKED 16-bit pairs with two AND masks per word.
// All three take the operand straight from global memory in the layout they find it: the fragment slots of an mma are permuted
// so that a thread's 8 or 16 consecutive bytes of a row ARE its fragment (a dot product does not care in which order k runs).
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1, int dt) {
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
// m16n8k8: the dx kernel contracts over the 8 LoRA ranks only
__device__ __forceinline__ void mma1688(float* c, uint32_t a0, uint32_t a1, uint32_t b0, int dt) {
#ifdef MRB_HOST_SHIM
  const uint32_t a[4] = {a0, a1, 0u, 0u};
  shim::mma_m16n8k16(c, a, b0, 0u, dt);
#else
  if (dt == MRB_DT_F16)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
  else
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
#endif
}
// The four draws of a mask word against ANY threshold, as AND masks over the two packed 16-bit pairs they cover (columns 0,1 and
// 2,3 of the word's group).  draw >= thr <=> draw + (256 - thr) carries into bit 8; even and odd bytes are summed separately so no
// carry crosses a draw.  add = (256 - thr) * 0x00010001.
__device__ __forceinline__ void pair_masks(uint32_t w, uint32_t add, uint32_t& m01, uint32_t& m23) {
  const uint32_t te = (w & 0x00ff00ffu) + add;                  // bit 8: draw 0 kept, bit 24: draw 2 kept
  const uint32_t to = ((w >> 8) & 0x00ff00ffu) + add;           // bit 8: draw 1 kept, bit 24: draw 3 kept
#ifdef MRB_HOST_SHIM
  m01 = (((te >> 8) & 1u) ? 0x0000ffffu : 0u) | (((to >> 8) & 1u) ? 0xffff0000u : 0u);
  m23 = (((te >> 24) & 1u) ? 0x0000ffffu : 0u) | (((to >> 24) & 1u) ? 0xffff0000u : 0u);
#else
  const uint32_t e7 = te << 7, o7 = to << 7;                    // the flags are now the sign bits of bytes 1 and 3
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m01) : "r"(e7), "r"(o7), "r"(0xDD99u));
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m23) : "r"(e7), "r"(o7), "r"(0xFFBBu));
#endif
}
__device__ __forceinline__ uint32_t flag_mask(uint32_t t, int bit) { return 0u - ((t >> bit) & 1u); }   // 0xffffffff if the bit is set

// ---- forward: u[m, 8j + r] = scale * sum_k keep_j(m,k) x[m,k] A[8j + r, k].  Block = 16 rows, its 8 warps split K; lane (g, t)
//      loads 16 bytes of rows g and g + 8 (columns k0 + 8t ..): fragment slots (2t, 2t+1 | 2t+8, 2t+9) of two mma's are those 8
//      columns, and the B fragment is the matching 16 bytes of A row 8j + g.  Partial sums of the warps meet in shared memory.
// NW warps split K: 8 for the encoder-sized calls, 16 for decoder-sized inputs (M <= 128: four blocks, the launch is one serial chain
// of K / (32 NW) load-hash-mma steps per warp -- 8.5 us at K 2048, 17.5 us at K 5120 with 8 warps, profiles/launch_summary_r02f.csv).
template <int NL, int NW = 8>
__global__ void __launch_bounds__(32 * NW) lora_down_drop_mma_kernel(const uint16_t* __restrict__ x, long long ldx, const uint16_t* __restrict__ A,
                                                                 long long lda, int M, int K, uint16_t* __restrict__ out, long long ldo,
                                                                 int dtype, const DropSpec d) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[NW * NL * 16 * 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.x * 16, r0 = m0 + g, r1 = r0 + 8;
  const bool ok0 = r0 < M, ok1 = r1 < M;
  const int steps = (K + 31) >> 5, per = (steps + NW - 1) / NW;
  const int ks = warp * per * 32, ke = min(K, ks + per * 32);
  const uint32_t ng = static_cast<uint32_t>(K >> 2), rb0 = static_cast<uint32_t>(r0) * ng, rb1 = static_cast<uint32_t>(r1) * ng;
  const uint32_t add = (256u - d.thr) * 0x00010001u;
  uint32_t key[NL];
  float acc[NL][4];
#pragma unroll
  for (int j = 0; j < NL; ++j) {
    key[j] = drop_key(*d.seed, d.site + j);
    acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  }
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int k0 = ks; k0 < ke; k0 += 32) {
    const int kk = k0 + 8 * t;
    const bool inb = kk < K;
    const uint4 xa = (ok0 && inb) ? *reinterpret_cast<const uint4*>(x + static_cast<long long>(r0) * ldx + kk) : zero4;
    const uint4 xb = (ok1 && inb) ? *reinterpret_cast<const uint4*>(x + static_cast<long long>(r1) * ldx + kk) : zero4;
    const uint32_t gq = static_cast<uint32_t>(kk) >> 2;
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const uint4 aj = inb ? *reinterpret_cast<const uint4*>(A + static_cast<long long>(8 * j + g) * lda + kk) : zero4;
      uint32_t a01, a23, a45, a67, b01, b23, b45, b67;
      pair_masks(drop_word(key[j], rb0, gq), add, a01, a23);
      pair_masks(drop_word(key[j], rb0, gq + 1u), add, a45, a67);
      pair_masks(drop_word(key[j], rb1, gq), add, b01, b23);
      pair_masks(drop_word(key[j], rb1, gq + 1u), add, b45, b67);
      const uint32_t f1[4] = {xa.x & a01, xb.x & b01, xa.y & a23, xb.y & b23};
      mma16816(acc[j], f1, aj.x, aj.y, dtype);
      const uint32_t f2[4] = {xa.z & a45, xb.z & b45, xa.w & a67, xb.w & b67};
      mma16816(acc[j], f2, aj.z, aj.w, dtype);
    }
  }
#pragma unroll
  for (int j = 0; j < NL; ++j) {
    float* rj = red + ((warp * NL + j) * 16) * 8;
    rj[g * 8 + 2 * t] = acc[j][0]; rj[g * 8 + 2 * t + 1] = acc[j][1];
    rj[(g + 8) * 8 + 2 * t] = acc[j][2]; rj[(g + 8) * 8 + 2 * t + 1] = acc[j][3];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 16 * 16; idx += 32 * NW) {        // (row, pair of output columns): 16 rows x 16 pairs = 32 columns
    const int row = idx >> 4, c2 = (idx & 15) * 2, m = m0 + row;
    if (m >= M) continue;
    float s0 = 0.f, s1 = 0.f;
    if (c2 < 8 * NL) {
      const int j = c2 >> 3, n = c2 & 7;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        s0 += red[((w * NL + j) * 16 + row) * 8 + n];
        s1 += red[((w * NL + j) * 16 + row) * 8 + n + 1];
      }
    }
    *reinterpret_cast<uint32_t*>(out + static_cast<long long>(m) * ldo + c2) = pack2(s0 * d.scale, s1 * d.scale, dtype);
  }
}

// ---- weight gradient: dA[r, k] += scale * sum_m keep(m,k) x[m,k] q[m,r].  The contraction runs over ROWS of x, whose pairs are
//      not adjacent in memory; so the two halves of a B register (x[m,k], x[m,k+1]) are made two different OUTPUT columns instead:
//      the 16 output rows of the mma are (rank r, parity), the A fragment holds q[m, r] in the half that matches its parity and 0 in
//      the other, and one mma contracts 8 rows of x for 16 columns.  Lane (g, t) loads 8 bytes of rows m + t and m + 4 + t (columns
//      kc0 + 4g ..): one mask word per load.  Block = 32 columns x rows_per_block rows, its 8 warps split the rows and meet in
//      shared memory: one fp32 atomic per output element and block.
__global__ void __launch_bounds__(256) lora_wgrad_drop_mma_kernel(const uint16_t* __restrict__ x, long long ldx, const uint16_t* __restrict__ q,
                                                                  long long ldq, int M, int K, float* __restrict__ dA, int dtype,
                                                                  int rows_per_block, const DropSpec d) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8 * 8 * 32];                              // [warp][rank][column]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int kc0 = blockIdx.x * 32, kk = kc0 + 4 * g;
  const bool inb = kk < K;                                       // K is a multiple of 8, kk of 4: the 8-byte load is in or out as a whole
  const int per_warp = rows_per_block >> 3;                      // multiple of 8
  const int m0 = blockIdx.y * rows_per_block + warp * per_warp, m1 = min(M, m0 + per_warp);
  const uint32_t key = drop_key(*d.seed, d.site), ng = static_cast<uint32_t>(K >> 2), gq = static_cast<uint32_t>(kk) >> 2;
  const uint32_t add = (256u - d.thr) * 0x00010001u;
  const int ra = g >> 1, sh = (g & 1) * 16;                      // rank of output row g (row g + 8: rank ra + 4), parity -> half
  float c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int mm = m0; mm < m1; mm += 8) {
    const int ma = mm + t, mb = mm + 4 + t;
    const bool oka = ma < m1, okb = mb < m1;
    uint2 xa = make_uint2(0u, 0u), xb = make_uint2(0u, 0u);
    uint4 qa = make_uint4(0u, 0u, 0u, 0u), qb = make_uint4(0u, 0u, 0u, 0u);
    if (oka) { qa = *reinterpret_cast<const uint4*>(q + static_cast<long long>(ma) * ldq); if (inb) xa = *reinterpret_cast<const uint2*>(x + static_cast<long long>(ma) * ldx + kk); }
    if (okb) { qb = *reinterpret_cast<const uint4*>(q + static_cast<long long>(mb) * ldq); if (inb) xb = *reinterpret_cast<const uint2*>(x + static_cast<long long>(mb) * ldx + kk); }
    uint32_t a01, a23, b01, b23;
    pair_masks(drop_word(key, static_cast<uint32_t>(ma) * ng, gq), add, a01, a23);
    pair_masks(drop_word(key, static_cast<uint32_t>(mb) * ng, gq), add, b01, b23);
    // q[m, ra] and q[m, ra + 4] (ranks 0..3 live in words x, y; 4..7 in z, w), moved into the half of this row's parity
    const uint32_t qa_lo = (((ra & 2) ? qa.y : qa.x) >> ((ra & 1) * 16)) & 0xffffu, qa_hi = (((ra & 2) ? qa.w : qa.z) >> ((ra & 1) * 16)) & 0xffffu;
    const uint32_t qb_lo = (((ra & 2) ? qb.y : qb.x) >> ((ra & 1) * 16)) & 0xffffu, qb_hi = (((ra & 2) ? qb.w : qb.z) >> ((ra & 1) * 16)) & 0xffffu;
    const uint32_t af[4] = {qa_lo << sh, qa_hi << sh, qb_lo << sh, qb_hi << sh};
    mma16816(c1, af, xa.x & a01, xb.x & b01, dtype);             // columns kc0 + 4n + parity
    mma16816(c2, af, xa.y & a23, xb.y & b23, dtype);             // columns kc0 + 4n + 2 + parity
  }
  // thread holds output rows g (rank ra) and g + 8 (rank ra + 4), mma columns n = 2t, 2t + 1 -> column 4n + parity (+ 2)
  const int par = g & 1;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int r = ra + (e >> 1) * 4, n = 2 * t + (e & 1);
    red[(warp * 8 + r) * 32 + 4 * n + par] = c1[e];
    red[(warp * 8 + r) * 32 + 4 * n + par + 2] = c2[e];
  }
  __syncthreads();
  {
    const int r = threadIdx.x >> 5, c = threadIdx.x & 31;          // 8 ranks x 32 columns = 256 threads
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[(w * 8 + r) * 32 + c];
    if (kc0 + c < K) atomicAdd(dA + static_cast<long long>(r) * K + kc0 + c, sum * d.scale);
  }
}

// ---- input gradient: dx[m,k] += scale * sum_j keep_j(m,k) (q_j[m,:] . A_j[:,k]).  One mma per (16 rows, 8 columns, Linear) with the 8
//      ranks in the lower half of its k slots; the 8 mma columns are permuted (n -> 4 (n / 2) + n % 2, second tile + 2) so that a
//      thread's accumulators of a tile pair are 4 CONSECUTIVE columns of rows g and g + 8: one mask word, one 8 / 16-byte
//      read-modify-write (mma m16n8k8: the k slots are exactly the 8 ranks).  A^T of the block's 256 columns sits in shared memory ([Linear][column][rank], 16 bit).
template <int NL, bool F32OUT>
__global__ void __launch_bounds__(256) lora_dx_drop_mma_kernel(const uint16_t* __restrict__ q, long long ldq, const uint16_t* __restrict__ A,
                                                               long long lda, void* __restrict__ dx, long long lddx, int M, int K, int dtype,
                                                               int rows_per_block, const DropSpec d) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) uint16_t At[NL * 256 * 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int cb = blockIdx.x * 256;
  for (int i = threadIdx.x; i < NL * 8 * 32; i += 256) {          // one 16-byte chunk (8 columns) of one A row -> 8 transposed entries
    const int r = i >> 5, c8 = (i & 31) * 8;                       // r = 8 j + rank
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (cb + c8 < K) v = *reinterpret_cast<const uint4*>(A + static_cast<long long>(r) * lda + cb + c8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint16_t* dst = At + ((r >> 3) * 256 + c8) * 8 + (r & 7);
#pragma unroll
    for (int e = 0; e < 4; ++e) { dst[(2 * e) * 8] = static_cast<uint16_t>(w[e] & 0xffffu); dst[(2 * e + 1) * 8] = static_cast<uint16_t>(w[e] >> 16); }
  }
  __syncthreads();
  const int kc0 = cb + warp * 32;
  if (kc0 >= K) return;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  const uint32_t ng = static_cast<uint32_t>(K >> 2), add = (256u - d.thr) * 0x00010001u;
  uint32_t key[NL];
#pragma unroll
  for (int j = 0; j < NL; ++j) key[j] = drop_key(*d.seed, d.site + j);
  const int coln = 4 * (g >> 1) + (g & 1);                          // column (inside a 16-column tile pair) that mma column g stands for
  for (int mm = m0; mm < m1; mm += 16) {
    const int r0 = mm + g, r1 = mm + g + 8;
    const bool ok0 = r0 < m1, ok1 = r1 < m1;
    uint32_t qf[NL][2];
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      qf[j][0] = ok0 ? *reinterpret_cast<const uint32_t*>(q + static_cast<long long>(r0) * ldq + 8 * j + 2 * t) : 0u;
      qf[j][1] = ok1 ? *reinterpret_cast<const uint32_t*>(q + static_cast<long long>(r1) * ldq + 8 * j + 2 * t) : 0u;
    }
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {                                // two 16-column tile pairs of the warp's 32 columns
      const int kp0 = kc0 + 16 * pr, kth = kp0 + 4 * t;             // this thread's 4 output columns
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      // the read half of the read-modify-write goes out BEFORE the mask / mma work below (it used to follow it: every warp then sat
      // on one exposed global round trip per 16 x 16 outputs)
      float4 old32[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
      uint2 old16[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
      if (kth < K) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = h ? r1 : r0;
          if (!(h ? ok1 : ok0)) continue;
          if (F32OUT) old32[h] = *reinterpret_cast<const float4*>(static_cast<const float*>(dx) + static_cast<long long>(r) * lddx + kth);
          else old16[h] = *reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(dx) + static_cast<long long>(r) * lddx + kth);
        }
      }
#pragma unroll
      for (int j = 0; j < NL; ++j) {
        const uint16_t* at = At + (j * 256 + (kp0 - cb) + coln) * 8 + 2 * t;
        const uint32_t bA = *reinterpret_cast<const uint32_t*>(at), bB = *reinterpret_cast<const uint32_t*>(at + 2 * 8);
        float ca[4] = {0.f, 0.f, 0.f, 0.f}, cc[4] = {0.f, 0.f, 0.f, 0.f};
        mma1688(ca, qf[j][0], qf[j][1], bA, dtype);                 // columns kth, kth + 1   (rows g: ca[0..1], g + 8: ca[2..3])
        mma1688(cc, qf[j][0], qf[j][1], bB, dtype);                 // columns kth + 2, kth + 3
        const uint32_t w0 = drop_word(key[j], static_cast<uint32_t>(r0) * ng, static_cast<uint32_t>(kth) >> 2);
        const uint32_t w1 = drop_word(key[j], static_cast<uint32_t>(r1) * ng, static_cast<uint32_t>(kth) >> 2);
        const uint32_t e0 = (w0 & 0x00ff00ffu) + add, o0 = ((w0 >> 8) & 0x00ff00ffu) + add;
        const uint32_t e1 = (w1 & 0x00ff00ffu) + add, o1 = ((w1 >> 8) & 0x00ff00ffu) + add;
        s0[0] += __uint_as_float(__float_as_uint(ca[0]) & flag_mask(e0, 8));  s0[1] += __uint_as_float(__float_as_uint(ca[1]) & flag_mask(o0, 8));
        s0[2] += __uint_as_float(__float_as_uint(cc[0]) & flag_mask(e0, 24)); s0[3] += __uint_as_float(__float_as_uint(cc[1]) & flag_mask(o0, 24));
        s1[0] += __uint_as_float(__float_as_uint(ca[2]) & flag_mask(e1, 8));  s1[1] += __uint_as_float(__float_as_uint(ca[3]) & flag_mask(o1, 8));
        s1[2] += __uint_as_float(__float_as_uint(cc[2]) & flag_mask(e1, 24)); s1[3] += __uint_as_float(__float_as_uint(cc[3]) & flag_mask(o1, 24));
      }
      if (kth >= K) continue;                                       // K is a multiple of 8 and kth of 4: all four columns in or out
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = h ? r1 : r0;
        if (!(h ? ok1 : ok0)) continue;
        const float* sv = h ? s1 : s0;
        if (F32OUT) {
          float4* pp = reinterpret_cast<float4*>(static_cast<float*>(dx) + static_cast<long long>(r) * lddx + kth);
          float4 a = old32[h];
          a.x += sv[0] * d.scale; a.y += sv[1] * d.scale; a.z += sv[2] * d.scale; a.w += sv[3] * d.scale;
          *pp = a;
        } else {
          uint2* pp = reinterpret_cast<uint2*>(static_cast<uint16_t*>(dx) + static_cast<long long>(r) * lddx + kth);
          const uint2 v = old16[h];
          *pp = make_uint2(pack2(unpack_lo(v.x, dtype) + sv[0] * d.scale, unpack_hi(v.x, dtype) + sv[1] * d.scale, dtype),
                           pack2(unpack_lo(v.y, dtype) + sv[2] * d.scale, unpack_hi(v.y, dtype) + sv[3] * d.scale, dtype));
        }
      }
    }
  }
}

}  // namespace mrb

using namespace mrb;
#define STREAM static_cast<cudaStream_t>(stream)
// MRB_LORA_DROP_MMA=0 keeps the CUDA-core LoRA-dropout kernels of round 1 (A/B measurements)
static bool lora_drop_mma() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MRB_LORA_DROP_MMA"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
static inline unsigned blocks_for(long long n, int per) { return static_cast<unsigned>((n + per - 1) / per); }
static inline bool bad16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }
static inline bool half_dt(int dt) { return dt == MRB_DT_F16 || dt == MRB_DT_BF16; }

extern "C" int mrb_dropout(const void* x, long long ldx, void* out, long long ldo, int rows, int cols, int dtype, int out_dtype,
                           const unsigned* seed, unsigned site, float p, void* stream) {
  if (rows <= 0 || cols <= 0) return MRB_OK;
  if (!seed || (cols & 3) || (ldx & 3) || (ldo & 3) || p < 0.f || p >= 1.f) return MRB_ERR_ARG;
  const DropSpec d = make_drop(seed, site, p);
  const unsigned grid = blocks_for(static_cast<long long>(rows) * (cols >> 2), 256);
  if (dtype == MRB_DT_F32 && out_dtype == MRB_DT_F32) {
    if (bad16(x) || bad16(out)) return MRB_ERR_ARG;
    MRB_LAUNCH((dropout_kernel<1>), grid, 256, 0, STREAM, x, ldx, out, ldo, rows, cols, dtype, d);
  } else if (dtype == MRB_DT_F32 && half_dt(out_dtype)) {
    if (bad16(x) || (reinterpret_cast<uintptr_t>(out) & 7)) return MRB_ERR_ARG;
    MRB_LAUNCH((dropout_kernel<2>), grid, 256, 0, STREAM, x, ldx, out, ldo, rows, cols, out_dtype, d);
  } else if (half_dt(dtype) && out_dtype == dtype) {
    if ((reinterpret_cast<uintptr_t>(x) & 7) || (reinterpret_cast<uintptr_t>(out) & 7)) return MRB_ERR_ARG;
    MRB_LAUNCH((dropout_kernel<0>), grid, 256, 0, STREAM, x, ldx, out, ldo, rows, cols, dtype, d);
  } else {
    return MRB_ERR_ARG;
  }
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_dropout_add(const float* resid, const float* branch, float* out, int rows, int cols, const unsigned* seed,
                               unsigned site, float p, void* stream) {
  if (rows <= 0 || cols <= 0) return MRB_OK;
  if (!seed || (cols & 3) || bad16(resid) || bad16(branch) || bad16(out) || p < 0.f || p >= 1.f) return MRB_ERR_ARG;
  const DropSpec d = make_drop(seed, site, p);
  MRB_LAUNCH((dropout_add_kernel), blocks_for(static_cast<long long>(rows) * (cols >> 2), 256), 256, 0, STREAM,
             reinterpret_cast<const float4*>(resid), reinterpret_cast<const float4*>(branch), reinterpret_cast<float4*>(out), rows, cols, d);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_gated_gelu_fwd_drop(const void* ab, void* h, int M, int F, long long ldh, int dtype, const unsigned* seed,
                                       unsigned site, float p, void* stream) {
  if (M <= 0) return MRB_OK;
  if (!seed || (F & 7) || (ldh & 7) || !half_dt(dtype) || p < 0.f || p >= 1.f) return MRB_ERR_ARG;
  const DropSpec d = make_drop(seed, site, p);
  MRB_LAUNCH((gated_gelu_fwd_drop_kernel), blocks_for(static_cast<long long>(M) * (F >> 3), 256), 256, 0, STREAM,
             static_cast<const uint4*>(ab), static_cast<uint4*>(h), M, F, ldh, dtype, d);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
extern "C" int mrb_gated_gelu_bwd_drop(const void* ab, const void* dh, long long lddh, void* dab, long long lddab, int M, int F,
                                       int dtype, const unsigned* seed, unsigned site, float p, void* stream) {
  if (M <= 0) return MRB_OK;
  if (!seed || (F & 7) || (lddh & 7) || (lddab & 7) || lddab < 2 * F || !half_dt(dtype) || p < 0.f || p >= 1.f) return MRB_ERR_ARG;
  const DropSpec d = make_drop(seed, site, p);
  MRB_LAUNCH((gated_gelu_bwd_drop_kernel), blocks_for(static_cast<long long>(M) * (F >> 3), 256), 256, 0, STREAM,
             static_cast<const uint4*>(ab), static_cast<const uint4*>(dh), lddh, static_cast<uint4*>(dab), lddab, M, F, dtype, d);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

template <int LANES>
static int launch_down(int nlin, unsigned grid, cudaStream_t s, const uint16_t* x, long long ldx, const uint16_t* A, long long lda, int M,
                       int K, uint16_t* out, long long ldo, int dtype, const DropSpec& d) {
  switch (nlin) {
    case 1: MRB_LAUNCH((lora_down_drop_kernel<1, LANES>), grid, 128, 0, s, x, ldx, A, lda, M, K, out, ldo, dtype, d); break;
    case 2: MRB_LAUNCH((lora_down_drop_kernel<2, LANES>), grid, 128, 0, s, x, ldx, A, lda, M, K, out, ldo, dtype, d); break;
    case 3: MRB_LAUNCH((lora_down_drop_kernel<3, LANES>), grid, 128, 0, s, x, ldx, A, lda, M, K, out, ldo, dtype, d); break;
    default: return MRB_ERR_ARG;
  }
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_lora_down_drop(const void* x, long long ldx, const void* A, long long lda, int M, int K, int nlin, void* out,
                                  long long ldo, int dtype, const unsigned* seed, unsigned site0, float p, void* stream) {
  if (M <= 0) return MRB_OK;
  if (!seed || K <= 0 || (K & 7) || (ldx & 7) || (lda & 7) || (ldo & 7) || !half_dt(dtype) || bad16(x) || bad16(A) || bad16(out) ||
      p < 0.f || p >= 1.f)
    return MRB_ERR_ARG;
  const DropSpec d = make_drop(seed, site0, p);
  const uint16_t* xp = static_cast<const uint16_t*>(x);
  const uint16_t* Ap = static_cast<const uint16_t*>(A);
  uint16_t* op = static_cast<uint16_t*>(out);
  if (lora_drop_mma()) {
    const unsigned grid = blocks_for(M, 16);
    static int wide = -1;                 // MRB_LORA_DOWN_WIDE=0: 8 warps per block for every M (A/B measurements)
    if (wide < 0) { const char* e = getenv("MRB_LORA_DOWN_WIDE"); wide = (e && e[0] == '0') ? 0 : 1; }
    if (wide && M <= 128 && K >= 1024) {
      switch (nlin) {
        case 1: MRB_LAUNCH((lora_down_drop_mma_kernel<1, 16>), grid, 512, 0, STREAM, xp, ldx, Ap, lda, M, K, op, ldo, dtype, d); break;
        case 2: MRB_LAUNCH((lora_down_drop_mma_kernel<2, 16>), grid, 512, 0, STREAM, xp, ldx, Ap, lda, M, K, op, ldo, dtype, d); break;
        case 3: MRB_LAUNCH((lora_down_drop_mma_kernel<3, 16>), grid, 512, 0, STREAM, xp, ldx, Ap, lda, M, K, op, ldo, dtype, d); break;
        default: return MRB_ERR_ARG;
      }
      MRB_CHECK_LAUNCH();
      return MRB_OK;
    }
    switch (nlin) {
      case 1: MRB_LAUNCH((lora_down_drop_mma_kernel<1>), grid, 256, 0, STREAM, xp, ldx, Ap, lda, M, K, op, ldo, dtype, d); break;
      case 2: MRB_LAUNCH((lora_down_drop_mma_kernel<2>), grid, 256, 0, STREAM, xp, ldx, Ap, lda, M, K, op, ldo, dtype, d); break;
      case 3: MRB_LAUNCH((lora_down_drop_mma_kernel<3>), grid, 256, 0, STREAM, xp, ldx, Ap, lda, M, K, op, ldo, dtype, d); break;
      default: return MRB_ERR_ARG;
    }
    MRB_CHECK_LAUNCH();
    return MRB_OK;
  }
  // few rows (decoder steps): a warp per row so that the launch still covers the SMs; otherwise 8 lanes per row
  if (M <= 2048) return launch_down<32>(nlin, blocks_for(M, 4), STREAM, xp, ldx, Ap, lda, M, K, op, ldo, dtype, d);
  return launch_down<8>(nlin, blocks_for(M, 16), STREAM, xp, ldx, Ap, lda, M, K, op, ldo, dtype, d);
}

extern "C" int mrb_lora_wgrad_drop(const void* x, long long ldx, const void* q, long long ldq, int M, int K, float* dA, int dtype,
                                   const unsigned* seed, unsigned site, float p, void* stream) {
  if (M <= 0 || K <= 0) return MRB_OK;
  if (!seed || (K & 7) || (ldx & 7) || (ldq & 7) || !half_dt(dtype) || bad16(x) || bad16(q) || p < 0.f || p >= 1.f) return MRB_ERR_ARG;
  const DropSpec d = make_drop(seed, site, p);
  if (lora_drop_mma()) {
    const int rpb = M <= 512 ? 64 : (M <= 4096 ? 512 : 2048);       // 8 warps x (rpb / 8) rows; a multiple of 64
    dim3 grid(blocks_for(K, 32), blocks_for(M, rpb));
    MRB_LAUNCH((lora_wgrad_drop_mma_kernel), grid, 256, 0, STREAM, static_cast<const uint16_t*>(x), ldx, static_cast<const uint16_t*>(q), ldq,
               M, K, dA, dtype, rpb, d);
    MRB_CHECK_LAUNCH();
    return MRB_OK;
  }
  const int rows_per_block = M <= 1024 ? 64 : 256;
  dim3 grid(blocks_for(K, 256), blocks_for(M, rows_per_block));
  MRB_LAUNCH((lora_wgrad_drop_kernel), grid, 256, 0, STREAM, static_cast<const uint16_t*>(x), ldx, static_cast<const uint16_t*>(q), ldq, M, K,
             dA, dtype, rows_per_block, d);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_lora_dx_drop(const void* q, long long ldq, const void* A, long long lda, int nlin, void* dx, long long lddx,
                                int dx_dtype, int M, int K, int dtype, const unsigned* seed, unsigned site0, float p, void* stream) {
  if (M <= 0 || K <= 0) return MRB_OK;
  if (!seed || (K & 7) || (ldq & 7) || (lda & 7) || (lddx & 7) || !half_dt(dtype) || (dx_dtype != MRB_DT_F32 && dx_dtype != dtype) ||
      bad16(q) || bad16(A) || bad16(dx) || p < 0.f || p >= 1.f)
    return MRB_ERR_ARG;
  const DropSpec d = make_drop(seed, site0, p);
  const int rows_per_block = M <= 1024 ? 16 : 128;       // the A^T staging of a block is amortised over 8 row steps
  dim3 grid(blocks_for(K, 256), blocks_for(M, rows_per_block));
  const uint16_t* qp = static_cast<const uint16_t*>(q);
  const uint16_t* Ap = static_cast<const uint16_t*>(A);
  if (lora_drop_mma()) {
#define MRB_DXM(NLV)                                                                                                               \
  if (dx_dtype == MRB_DT_F32) MRB_LAUNCH((lora_dx_drop_mma_kernel<NLV, true>), grid, 256, 0, STREAM, qp, ldq, Ap, lda, dx, lddx, M, K,   \
                                         dtype, rows_per_block, d);                                                               \
  else MRB_LAUNCH((lora_dx_drop_mma_kernel<NLV, false>), grid, 256, 0, STREAM, qp, ldq, Ap, lda, dx, lddx, M, K, dtype, rows_per_block, d)
    switch (nlin) {
      case 1: MRB_DXM(1); break;
      case 2: MRB_DXM(2); break;
      case 3: MRB_DXM(3); break;
      default: return MRB_ERR_ARG;
    }
#undef MRB_DXM
    MRB_CHECK_LAUNCH();
    return MRB_OK;
  }
#define MRB_DX(NLV)                                                                                                                \
  if (dx_dtype == MRB_DT_F32) MRB_LAUNCH((lora_dx_drop_kernel<NLV, true>), grid, 256, 0, STREAM, qp, ldq, Ap, lda, dx, lddx, M, K, dtype, \
                                         rows_per_block, d);                                                                      \
  else MRB_LAUNCH((lora_dx_drop_kernel<NLV, false>), grid, 256, 0, STREAM, qp, ldq, Ap, lda, dx, lddx, M, K, dtype, rows_per_block, d)
  switch (nlin) {
    case 1: MRB_DX(1); break;
    case 2: MRB_DX(2); break;
    case 3: MRB_DX(3); break;
    default: return MRB_ERR_ARG;
  }
#undef MRB_DX
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
