// Train-mode dropout outside the attention kernels (masks: dropmask.cuh): the residual-branch / embedding / final-norm / FF-inner
// sites of T5 (modeling_t5.py:327,346,652,690,1149,1258) and of the Q-Former (Qformer.py:107,287,373) as one-pass HBM-bound
// kernels, and peft's LoRA input dropout (blip2_mr.py:197: y = W x + B A drop_j(x), an independent mask per adapted Linear),
// which breaks the "LoRA folded into the GEMM's K extension" layout of the backward pass and gets three CUDA-core kernels:
//   forward   u_j  = drop_j(x) A_j^T                          (mrb_lora_down_drop: fills the 32 extension columns of x_ext)
//   backward  dA_j += (dy sB_j)^T drop_j(x)                   (mrb_lora_wgrad_drop)
//             dx   += sum_j mask_j * ((dy sB_j) A_j)          (mrb_lora_dx_drop; the dense dy W part stays one tcgen05 GEMM over K = N)
// dB_j = dy_j^T u_j needs no change: u_j already carries the mask.  Every backward kernel recomputes its mask from
// (seed, site, row, column).
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
    oa[i] = pack2(d0 * b0 * gelu_erf_grad(a0), d1 * b1 * gelu_erf_grad(a1), dtype);
    ob[i] = pack2(d0 * gelu_erf(a0), d1 * gelu_erf(a1), dtype);
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

}  // namespace mrb

using namespace mrb;
#define STREAM static_cast<cudaStream_t>(stream)
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
  // few rows (decoder steps): a warp per row so that the launch still covers the SMs; otherwise 8 lanes per row
  if (M <= 2048) return launch_down<32>(nlin, blocks_for(M, 4), STREAM, xp, ldx, Ap, lda, M, K, op, ldo, dtype, d);
  return launch_down<8>(nlin, blocks_for(M, 16), STREAM, xp, ldx, Ap, lda, M, K, op, ldo, dtype, d);
}

extern "C" int mrb_lora_wgrad_drop(const void* x, long long ldx, const void* q, long long ldq, int M, int K, float* dA, int dtype,
                                   const unsigned* seed, unsigned site, float p, void* stream) {
  if (M <= 0 || K <= 0) return MRB_OK;
  if (!seed || (K & 7) || (ldx & 7) || (ldq & 7) || !half_dt(dtype) || bad16(x) || bad16(q) || p < 0.f || p >= 1.f) return MRB_ERR_ARG;
  const DropSpec d = make_drop(seed, site, p);
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
  const int rows_per_block = M <= 1024 ? 16 : 64;
  dim3 grid(blocks_for(K, 256), blocks_for(M, rows_per_block));
  const uint16_t* qp = static_cast<const uint16_t*>(q);
  const uint16_t* Ap = static_cast<const uint16_t*>(A);
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
