// HBM-bound kernels of the path: LayerNorm / T5 RMSNorm (+backward), patch extraction for the ViT
// patch-embed GEMM, cls/pos rows, gated-GELU, the interleave gather that assembles inputs_embeds
// (blip2_mr.py:691-783), embedding lookup, cross-entropy, LoRA down-projection and LoRA weight
// gradients, casts/transposes.  All are one pass over their operands with 16-byte accesses.
#include "common.cuh"
#include "dropmask.cuh"

namespace mrb {

// ---------------------------------------------------------------- LayerNorm / RMSNorm (fp32 in)
// One warp per row, row cached in registers (C <= 2048, C % 4 == 0).
// mode 0: LayerNorm(x [+ add]) * w + b (eva_vit.py:175-176, blip2.py:113-119, Qformer.py:278-289)
// mode 1: T5 RMSNorm (modeling_t5.py:263-277)
template <int MAXV>
__global__ void __launch_bounds__(256) norm_kernel(const float* __restrict__ x, const float* __restrict__ add,
                                                   const float* __restrict__ w, const float* __restrict__ bias,
                                                   float eps, int rows, int C, int mode, float* __restrict__ out_f32,
                                                   void* __restrict__ out_h, int h_dtype, long long ld_h,
                                                   float* __restrict__ sum_out) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // 8 warps per block; 1 for decoder-sized inputs (norm_warps)
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nv = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * C);
  const float4* ar = add ? reinterpret_cast<const float4*>(add + static_cast<long long>(row) * C) : nullptr;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      v[i] = xr[c];
      if (ar) { const float4 a = ar[c]; v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w; }
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  if (sum_out) {   // x + add is also the next residual stream
    float4* so = reinterpret_cast<float4*>(sum_out + static_cast<long long>(row) * C);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) { const int c = lane + i * 32; if (c < nv) so[c] = v[i]; }
  }
  float mean = 0.f;
  if (mode == 0) mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + cc * cc + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  const float4* wr = reinterpret_cast<const float4*>(w);
  const float4* br = bias ? reinterpret_cast<const float4*>(bias) : nullptr;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float4 ww = wr[c];
      float4 y;
      y.x = (v[i].x - mean) * rstd * ww.x; y.y = (v[i].y - mean) * rstd * ww.y;
      y.z = (v[i].z - mean) * rstd * ww.z; y.w = (v[i].w - mean) * rstd * ww.w;
      if (br) { const float4 bb = br[c]; y.x += bb.x; y.y += bb.y; y.z += bb.z; y.w += bb.w; }
      if (out_f32) reinterpret_cast<float4*>(out_f32 + static_cast<long long>(row) * C)[c] = y;
      if (out_h) {
        uint2 pk;
        pk.x = pack2(y.x, y.y, h_dtype);
        pk.y = pack2(y.z, y.w, h_dtype);
        reinterpret_cast<uint2*>(static_cast<uint16_t*>(out_h) + static_cast<long long>(row) * ld_h)[c] = pk;
      }
    }
  }
}

// Same arithmetic (two-pass mean / centred variance), one BLOCK per row for the encoder- / ViT-sized calls: 256 threads x <= 2 float4,
// every load of a thread (x, add, w, bias) issued up front, 64 warps per SM instead of 8-16 (the warp-per-row kernel holds a whole
// row in up to 64 registers per lane), row sums through shared memory in a fixed order.
__global__ void __launch_bounds__(256) norm_row_kernel(const float* __restrict__ x, const float* __restrict__ add,
                                                       const float* __restrict__ w, const float* __restrict__ bias, float eps,
                                                       int rows, int C, int mode, float* __restrict__ out_f32,
                                                       void* __restrict__ out_h, int h_dtype, long long ld_h,
                                                       float* __restrict__ sum_out, const DropSpec dsp) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  __shared__ float red[2][8];
  const int row = blockIdx.x, nv = C >> 2, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // dsp.seed != null: `add` is a sublayer output under train-mode dropout -- x + drop(add) is the new residual stream (sum_out)
  // and its norm the next sublayer's input: dropout_add_kernel + this kernel in one pass (same arithmetic, same order)
  const uint32_t dkey = dsp.seed ? drop_key(*dsp.seed, dsp.site) : 0u, drow = static_cast<uint32_t>(row) * static_cast<uint32_t>(nv);
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * C);
  const float4* ar = add ? reinterpret_cast<const float4*>(add + static_cast<long long>(row) * C) : nullptr;
  const float4* wr = reinterpret_cast<const float4*>(w);
  const float4* br = bias ? reinterpret_cast<const float4*>(bias) : nullptr;
  float4 v[2], ww[2], bb[2];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = threadIdx.x + i * 256;
    v[i] = ww[i] = bb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < nv) {
      v[i] = xr[c];
      ww[i] = wr[c];
      if (br) bb[i] = br[c];
      if (ar) {
        float4 a = ar[c];
        if (dsp.seed) {
          const uint32_t w = drop_word(dkey, drow, static_cast<uint32_t>(c));
          a.x = drop_keep(w, 0, dsp.thr) ? a.x * dsp.scale : 0.f; a.y = drop_keep(w, 1, dsp.thr) ? a.y * dsp.scale : 0.f;
          a.z = drop_keep(w, 2, dsp.thr) ? a.z * dsp.scale : 0.f; a.w = drop_keep(w, 3, dsp.thr) ? a.w * dsp.scale : 0.f;
        }
        v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w;
      }
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  if (sum_out) {   // x + add is also the next residual stream
    float4* so = reinterpret_cast<float4*>(sum_out + static_cast<long long>(row) * C);
#pragma unroll
    for (int i = 0; i < 2; ++i) { const int c = threadIdx.x + i * 256; if (c < nv) so[c] = v[i]; }
  }
  float mean = 0.f;
  if (mode == 0) {
    s = warp_sum(s);
    if (lane == 0) red[0][warp] = s;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[0][k];
    mean = t / C;
  }
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + cc * cc + d * d;
    }
  }
  q = warp_sum(q);
  if (lane == 0) red[1][warp] = q;
  __syncthreads();
  float qt = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) qt += red[1][k];
  const float rstd = rsqrtf(qt / C + eps);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < nv) {
      float4 y;
      y.x = (v[i].x - mean) * rstd * ww[i].x + bb[i].x; y.y = (v[i].y - mean) * rstd * ww[i].y + bb[i].y;
      y.z = (v[i].z - mean) * rstd * ww[i].z + bb[i].z; y.w = (v[i].w - mean) * rstd * ww[i].w + bb[i].w;
      if (out_f32) reinterpret_cast<float4*>(out_f32 + static_cast<long long>(row) * C)[c] = y;
      if (out_h) {
        uint2 pk;
        pk.x = pack2(y.x, y.y, h_dtype);
        pk.y = pack2(y.z, y.w, h_dtype);
        reinterpret_cast<uint2*>(static_cast<uint16_t*>(out_h) + static_cast<long long>(row) * ld_h)[c] = pk;
      }
    }
  }
}

// RMSNorm backward: dres[row] += rstd * (w*dy) - x * rstd^3 * mean(w*dy*x)       (weights frozen: no dw)
// dy: fp32 or 16-bit [rows, ld]; when A != null the LoRA input gradient is folded in first:
//     dy_eff[k] = dy[k] + sum_r dy[C + r] * A[r, k]      (dy is then an "extended" dgrad buffer [rows, C+32])
template <int MAXV>
__global__ void __launch_bounds__(256) rmsnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const void* __restrict__ dy, int dy_dtype, long long ld,
                                                          const float* __restrict__ A, int R, float eps, int rows, int C,
                                                          float* __restrict__ dres) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nv = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * C);
  const float4* wr = reinterpret_cast<const float4*>(w);
  float dxa = 0.f;   // lane r holds dy[C + r]
  if (A && lane < R) {
    const uint16_t v = static_cast<const uint16_t*>(dy)[static_cast<long long>(row) * ld + C + lane];
    dxa = (dy_dtype == MRB_DT_F16) ? __half2float(__ushort_as_half(v)) : __uint_as_float(static_cast<uint32_t>(v) << 16);
  }
  float4 xv[MAXV], gv[MAXV];
  float ss = 0.f, dot = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      xv[i] = xr[c];
      float4 d;
      if (dy_dtype == MRB_DT_F32) {
        d = reinterpret_cast<const float4*>(static_cast<const float*>(dy) + static_cast<long long>(row) * ld)[c];
      } else {
        const uint2 pk = reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(dy) + static_cast<long long>(row) * ld)[c];
        d = make_float4(unpack_lo(pk.x, dy_dtype), unpack_hi(pk.x, dy_dtype), unpack_lo(pk.y, dy_dtype), unpack_hi(pk.y, dy_dtype));
      }
      if (A) {
        for (int r = 0; r < R; ++r) {
          const float g = __shfl_sync(0xffffffffu, dxa, r);
          const float4 a = reinterpret_cast<const float4*>(A + static_cast<long long>(r) * C)[c];
          d.x += g * a.x; d.y += g * a.y; d.z += g * a.z; d.w += g * a.w;
        }
      }
      const float4 ww = wr[c];
      gv[i] = make_float4(d.x * ww.x, d.y * ww.y, d.z * ww.z, d.w * ww.w);
      ss += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
      dot += gv[i].x * xv[i].x + gv[i].y * xv[i].y + gv[i].z * xv[i].z + gv[i].w * xv[i].w;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / C + eps);
  const float coef = warp_sum(dot) / C * rstd * rstd * rstd;
  float4* out = reinterpret_cast<float4*>(dres + static_cast<long long>(row) * C);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      float4 o = out[c];
      o.x += gv[i].x * rstd - xv[i].x * coef; o.y += gv[i].y * rstd - xv[i].y * coef;
      o.z += gv[i].z * rstd - xv[i].z * coef; o.w += gv[i].w * rstd - xv[i].w * coef;
      out[c] = o;
    }
  }
}

// Same arithmetic, one BLOCK per row (256 threads x <= 2 float4 each) for the encoder-sized calls (8192 rows x 2048): the warp-per-row
// kernel above keeps x and w*dy of a whole row in 128 registers, runs 8 warps per SM and reads dres only after both reductions --
// 103 us for 235 MB (2.3 TB/s, profiles/launch_summary_r02d.csv).  Here every load of a thread (x, dy, w, dres) is issued up front,
// 64 warps fit an SM, and the two row sums meet in shared memory in a fixed order.
__global__ void __launch_bounds__(256) rmsnorm_bwd_row_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const void* __restrict__ dy, int dy_dtype, long long ld, float eps,
                                                              int rows, int C, float* __restrict__ dres, const DropSpec dsp,
                                                              uint16_t* __restrict__ dy_next, int next_dtype, long long ld_next) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  __shared__ float red[2][8];
  const int row = blockIdx.x, nv = C >> 2, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * C);
  const float4* wr = reinterpret_cast<const float4*>(w);
  float4* out = reinterpret_cast<float4*>(dres + static_cast<long long>(row) * C);
  float4 xv[2], gv[2], ov[2];
  float ss = 0.f, dot = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = threadIdx.x + i * 256;
    xv[i] = gv[i] = ov[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < nv) {
      xv[i] = xr[c];
      ov[i] = out[c];
      float4 d;
      if (dy_dtype == MRB_DT_F32) {
        d = reinterpret_cast<const float4*>(static_cast<const float*>(dy) + static_cast<long long>(row) * ld)[c];
      } else {
        const uint2 pk = reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(dy) + static_cast<long long>(row) * ld)[c];
        d = make_float4(unpack_lo(pk.x, dy_dtype), unpack_hi(pk.x, dy_dtype), unpack_lo(pk.y, dy_dtype), unpack_hi(pk.y, dy_dtype));
      }
      const float4 ww = wr[c];
      gv[i] = make_float4(d.x * ww.x, d.y * ww.y, d.z * ww.z, d.w * ww.w);
      ss += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
      dot += gv[i].x * xv[i].x + gv[i].y * xv[i].y + gv[i].z * xv[i].z + gv[i].w * xv[i].w;
    }
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  if (lane == 0) { red[0][warp] = ss; red[1][warp] = dot; }
  __syncthreads();
  float sst = 0.f, dott = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { sst += red[0][k]; dott += red[1][k]; }
  const float rstd = rsqrtf(sst / C + eps);
  const float coef = dott / C * rstd * rstd * rstd;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < nv) {
      float4 o = ov[i];
      o.x += gv[i].x * rstd - xv[i].x * coef; o.y += gv[i].y * rstd - xv[i].y * coef;
      o.z += gv[i].z * rstd - xv[i].z * coef; o.w += gv[i].w * rstd - xv[i].w * coef;
      out[c] = o;
      if (dy_next) {
        // the next sublayer's backward starts from this gradient under ITS residual-dropout mask, as a 16-bit dgrad operand
        // (dropout_kernel<2> of csrc/dropout.cu, here without re-reading the fp32 gradient)
        const uint32_t w = drop_word(drop_key(*dsp.seed, dsp.site), static_cast<uint32_t>(row) * static_cast<uint32_t>(nv),
                                     static_cast<uint32_t>(c));
        const float m0 = drop_keep(w, 0, dsp.thr) ? o.x * dsp.scale : 0.f, m1 = drop_keep(w, 1, dsp.thr) ? o.y * dsp.scale : 0.f;
        const float m2 = drop_keep(w, 2, dsp.thr) ? o.z * dsp.scale : 0.f, m3 = drop_keep(w, 3, dsp.thr) ? o.w * dsp.scale : 0.f;
        *reinterpret_cast<uint2*>(dy_next + static_cast<long long>(row) * ld_next + 4 * c) =
            make_uint2(pack2(m0, m1, next_dtype), pack2(m2, m3, next_dtype));
      }
    }
  }
}

// LoRA input-gradient fix-up on an extended dgrad buffer x [M, K+32] (16-bit):
//   y[m, k] = x[m, k] + sum_r x[m, K + r] * A[r, k];   in place, or accumulated into fp32 acc [M, K] when acc != null
__global__ void __launch_bounds__(256) lora_up_add_kernel(uint16_t* __restrict__ x, long long ldx, const float* __restrict__ A,
                                                           int R, int M, int K, int dtype, float* __restrict__ acc) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const int kv = K >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(M) * kv) return;
  const int c = (idx % kv) * 8;
  const long long m = idx / kv;
  uint16_t* xr = x + m * ldx;
  const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  float y[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { y[2 * i] = unpack_lo(w[i], dtype); y[2 * i + 1] = unpack_hi(w[i], dtype); }
  for (int r = 0; r < R; r += 2) {
    const uint32_t gw = *reinterpret_cast<const uint32_t*>(xr + K + r);
    const float g0 = unpack_lo(gw, dtype), g1 = unpack_hi(gw, dtype);
    const float4 a0 = *reinterpret_cast<const float4*>(A + static_cast<long long>(r) * K + c);
    const float4 a1 = *reinterpret_cast<const float4*>(A + static_cast<long long>(r) * K + c + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(A + static_cast<long long>(r + 1) * K + c);
    const float4 b1 = *reinterpret_cast<const float4*>(A + static_cast<long long>(r + 1) * K + c + 4);
    y[0] += g0 * a0.x + g1 * b0.x; y[1] += g0 * a0.y + g1 * b0.y; y[2] += g0 * a0.z + g1 * b0.z; y[3] += g0 * a0.w + g1 * b0.w;
    y[4] += g0 * a1.x + g1 * b1.x; y[5] += g0 * a1.y + g1 * b1.y; y[6] += g0 * a1.z + g1 * b1.z; y[7] += g0 * a1.w + g1 * b1.w;
  }
  if (acc) {
    float4* ap = reinterpret_cast<float4*>(acc + m * K + c);
    float4 p0 = ap[0], p1 = ap[1];
    p0.x += y[0]; p0.y += y[1]; p0.z += y[2]; p0.w += y[3];
    p1.x += y[4]; p1.y += y[5]; p1.z += y[6]; p1.w += y[7];
    ap[0] = p0; ap[1] = p1;
  } else {
    *reinterpret_cast<uint4*>(xr + c) = make_uint4(pack2(y[0], y[1], dtype), pack2(y[2], y[3], dtype),
                                                   pack2(y[4], y[5], dtype), pack2(y[6], y[7], dtype));
  }
}

// ---------------------------------------------------------------- ViT patch extraction (eva_vit.py:196-203)
// frames fp32 [F,3,S,S] -> A [F*G*G, ldA] half with column k = c*P*P + i*P + j (the Conv2d weight's flattened
// order); columns [3*P*P, ldA) are zero.  One thread per (patch, c, i): P contiguous pixels.
__global__ void patchify_kernel(const float* __restrict__ img, void* __restrict__ out, int dtype, int F, int S, int P,
                                int ldA) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const int G = S / P;
  const long long total = static_cast<long long>(F) * G * G * 3 * P;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  // fastest: patch column px (coalesced reads along an image row), then i, c, patch row, frame
  const int px = idx % G;
  long long r = idx / G;
  const int i = r % P; r /= P;
  const int c = r % 3; r /= 3;
  const int py = r % G;
  const int f = r / G;
  const float* src = img + ((static_cast<long long>(f) * 3 + c) * S + (py * P + i)) * S + px * P;
  uint16_t* dst = static_cast<uint16_t*>(out) + (static_cast<long long>(f) * G * G + py * G + px) * ldA + (c * P + i) * P;
  for (int j = 0; j + 1 < P; j += 2) {
    const float2 v = *reinterpret_cast<const float2*>(src + j);
    *reinterpret_cast<uint32_t*>(dst + j) = pack2(v.x, v.y, dtype);
  }
  if (P & 1) dst[P - 1] = static_cast<uint16_t>(pack2(src[P - 1], 0.f, dtype) & 0xffff);
  if (c == 2 && i == P - 1) {
    for (int k = 3 * P * P; k < ldA; ++k) dst[k - (c * P + i) * P] = 0;
  }
}

// Same patch extraction from RAW uint8 frames [F,3,S,S] with the processors' normalisation fused in
// (lavis/processors/blip_processors.py:61-70, 355-395: ToTensorVideo x/255, then NormalizeVideo (x - mean) / std, in that
// operation order so the fp32 value -- and hence the fp16 patch matrix -- is bit-identical to the host-normalised path).
// A clip then crosses PCIe as 9 MB instead of 36 MB.
struct NormConst { float mean[3], std[3]; };
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ img, void* __restrict__ out, int dtype, int F, int S, int P,
                                   int ldA, NormConst nc) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const int G = S / P;
  const long long total = static_cast<long long>(F) * G * G * 3 * P;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int px = idx % G;
  long long r = idx / G;
  const int i = r % P; r /= P;
  const int c = r % 3; r /= 3;
  const int py = r % G;
  const int f = r / G;
  const uint8_t* src = img + ((static_cast<long long>(f) * 3 + c) * S + (py * P + i)) * S + px * P;
  uint16_t* dst = static_cast<uint16_t*>(out) + (static_cast<long long>(f) * G * G + py * G + px) * ldA + (c * P + i) * P;
  const float mean = nc.mean[c], sd = nc.std[c];
  for (int j = 0; j + 1 < P; j += 2) {
    const float a = __fdiv_rn(__fdiv_rn(static_cast<float>(src[j]), 255.0f) - mean, sd);
    const float b = __fdiv_rn(__fdiv_rn(static_cast<float>(src[j + 1]), 255.0f) - mean, sd);
    *reinterpret_cast<uint32_t*>(dst + j) = pack2(a, b, dtype);
  }
  if (P & 1) dst[P - 1] = static_cast<uint16_t>(pack2(__fdiv_rn(__fdiv_rn(static_cast<float>(src[P - 1]), 255.0f) - mean, sd), 0.f, dtype) & 0xffff);
  if (c == 2 && i == P - 1) {
    for (int k = 3 * P * P; k < ldA; ++k) dst[k - (c * P + i) * P] = 0;
  }
}

// x[f, 0, :] = cls + pos[0]   (eva_vit.py:328-331)
__global__ void cls_pos_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x,
                               int F, int tokens, int C) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(F) * C) return;
  const int c = idx % C;
  const int f = idx / C;
  x[(static_cast<long long>(f) * tokens) * C + c] = cls[c] + pos[c];
}

// ---------------------------------------------------------------- gated GELU (modeling_t5.py:323-329)
// ab [M, 2F] half (columns [0,F) = wi_0 x, [F,2F) = wi_1 x) -> h [M, ldh] half = gelu(a) * b
__global__ void gated_gelu_fwd_kernel(const uint4* __restrict__ ab, uint4* __restrict__ h, int M, int F, long long ldh,
                                      int dtype) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const int fv = F >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(M) * fv) return;
  const int c = idx % fv;
  const long long m = idx / fv;
  const uint4 a = ab[m * (2 * fv) + c], b = ab[m * (2 * fv) + fv + c];
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    o[i] = pack2(gelu_erf(unpack_lo(aw[i], dtype)) * unpack_lo(bw[i], dtype),
                 gelu_erf(unpack_hi(aw[i], dtype)) * unpack_hi(bw[i], dtype), dtype);
  h[m * (ldh >> 3) + c] = make_uint4(o[0], o[1], o[2], o[3]);
}
// dab[:, :F] = dh * b * gelu'(a) ; dab[:, F:] = dh * gelu(a)
__global__ void gated_gelu_bwd_kernel(const uint4* __restrict__ ab, const uint4* __restrict__ dh, long long lddh,
                                      uint4* __restrict__ dab, long long lddab, int M, int F, int dtype) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const int fv = F >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(M) * fv) return;
  const int c = idx % fv;
  const long long m = idx / fv;
  const uint4 a = ab[m * (2 * fv) + c], b = ab[m * (2 * fv) + fv + c], d = dh[m * (lddh >> 3) + c];
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, dw[4] = {d.x, d.y, d.z, d.w};
  uint32_t oa[4], ob[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a0 = unpack_lo(aw[i], dtype), a1 = unpack_hi(aw[i], dtype);
    const float b0 = unpack_lo(bw[i], dtype), b1 = unpack_hi(bw[i], dtype);
    const float d0 = unpack_lo(dw[i], dtype), d1 = unpack_hi(dw[i], dtype);
    float g0, g1, dg0, dg1;
    gelu_erf_both(a0, g0, dg0);
    gelu_erf_both(a1, g1, dg1);
    oa[i] = pack2(d0 * b0 * dg0, d1 * b1 * dg1, dtype);
    ob[i] = pack2(d0 * g0, d1 * g1, dtype);
  }
  dab[m * (lddab >> 3) + c] = make_uint4(oa[0], oa[1], oa[2], oa[3]);
  dab[m * (lddab >> 3) + fv + c] = make_uint4(ob[0], ob[1], ob[2], ob[3]);
}

// ---------------------------------------------------------------- interleave gather (blip2_mr.py:691-783)
// out[r, :] (fp32) = table row:  idx >= 0 -> emb[idx] (fp32 embedding table)
//                               idx <  0 and != INT_MIN -> frames[-(idx+1)] (fp32 frame tokens)
//                               idx == INT_MIN -> zeros (left padding)
__global__ void gather_rows_kernel(const int* __restrict__ idx, const float* __restrict__ emb,
                                   const float* __restrict__ frames, float* __restrict__ out, int rows, int C) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const int r = blockIdx.x;
  if (r >= rows) return;
  const int id = idx[r];
  const float4* src = nullptr;
  if (id >= 0) src = reinterpret_cast<const float4*>(emb + static_cast<long long>(id) * C);
  else if (id != INT_MIN) src = reinterpret_cast<const float4*>(frames + static_cast<long long>(-(id + 1)) * C);
  float4* dst = reinterpret_cast<float4*>(out + static_cast<long long>(r) * C);
  for (int c = threadIdx.x; c < (C >> 2); c += blockDim.x) dst[c] = src ? src[c] : make_float4(0.f, 0.f, 0.f, 0.f);
}
// backward of the gather w.r.t. the frame tokens: dframes[-(idx+1)] = dout[r]  (each frame row appears once)
__global__ void scatter_frames_kernel(const int* __restrict__ idx, const float* __restrict__ dout,
                                      float* __restrict__ dframes, int rows, int C) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const int r = blockIdx.x;
  if (r >= rows) return;
  const int id = idx[r];
  if (id >= 0 || id == INT_MIN) return;
  const float4* src = reinterpret_cast<const float4*>(dout + static_cast<long long>(r) * C);
  float4* dst = reinterpret_cast<float4*>(dframes + static_cast<long long>(-(id + 1)) * C);
  for (int c = threadIdx.x; c < (C >> 2); c += blockDim.x) dst[c] = src[c];
}

// mean over groups of n consecutive rows (frame_token_aggregation == "mean", blip2_mr.py:493-498)
__global__ void group_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int groups, int n, int C) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(groups) * C) return;
  const int c = idx % C;
  const long long gidx = idx / C;
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += x[(gidx * n + i) * C + c];
  out[idx] = s / n;
}
__global__ void group_mean_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int groups, int n, int C) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(groups) * n * C) return;
  const int c = idx % C;
  const long long row = idx / C;
  dx[idx] = dout[(row / n) * C + c] / n;
}

// ---------------------------------------------------------------- cross entropy (modeling_t5.py:1872-1875)
// logits fp32 [rows, V]; labels int64 (-100 = ignore).  loss_sum += -log p[label]; dlogits = (p - onehot) * gscale.
// gscale < 0: mean reduction, 1 / #valid targets counted on the device (keeps the step free of host syncs / graph-capturable)
__global__ void __launch_bounds__(1024) ce_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                  int rows, int V, float* __restrict__ row_loss, void* __restrict__ dlogits,
                                                  int d_dtype, long long ldd, float gscale,
                                                  float* __restrict__ loss_sum) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  __shared__ float red[32];
  __shared__ float bval;
  const int row = blockIdx.x;
  if (gscale < 0.f) {
    int cnt = 0;
    for (int r0 = 0; r0 < rows; r0 += blockDim.x) {
      const int r = r0 + threadIdx.x;
      cnt += __syncthreads_count(r < rows && labels[r] >= 0);
    }
    gscale = 1.f / static_cast<float>(max(cnt, 1));
  }
  const float* lr = logits + static_cast<long long>(row) * V;
  const long long lab = labels[row];
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < V; c += blockDim.x) mx = fmaxf(mx, lr[c]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
    v = warp_max(v);
    if (threadIdx.x == 0) bval = v;
  }
  __syncthreads();
  mx = bval;
  float sm = 0.f;
  for (int c = threadIdx.x; c < V; c += blockDim.x) sm += __expf(lr[c] - mx);
  sm = warp_sum(sm);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sm;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) bval = v;
  }
  __syncthreads();
  const float lse = mx + logf(bval);
  if (threadIdx.x == 0) {
    const float l = (lab >= 0) ? (lse - lr[lab]) : 0.f;
    if (row_loss) row_loss[row] = l;
    if (loss_sum && lab >= 0) atomicAdd(loss_sum, l * gscale);   // gscale = 1 / #valid targets -> mean loss
  }
  if (dlogits) {
    uint16_t* dr = static_cast<uint16_t*>(dlogits) + static_cast<long long>(row) * ldd;
    const float gs = (lab >= 0) ? gscale : 0.f;
    for (int c = threadIdx.x * 2; c < V; c += blockDim.x * 2) {
      float p0 = __expf(lr[c] - lse), p1 = (c + 1 < V) ? __expf(lr[c + 1] - lse) : 0.f;
      if (c == lab) p0 -= 1.f;
      if (c + 1 == lab) p1 -= 1.f;
      *reinterpret_cast<uint32_t*>(dr + c) = pack2(p0 * gs, p1 * gs, d_dtype);
    }
  }
}

// ---------------------------------------------------------------- LoRA down-projection
// xa[m, j] = sum_k x[m, k] * A[j, k], j < R (R <= 32), written (zero-padded to 32) at x[m, K .. K+32).
// x is the [M, K+32] "extended" activation buffer consumed by the tcgen05 GEMM with [W | B | 0] weights.
template <int R>
__global__ void __launch_bounds__(256) lora_down_kernel(uint16_t* __restrict__ x, long long ldx, const float* __restrict__ A,
                                                         int M, int K, int dtype) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  uint16_t* xr = x + static_cast<long long>(row) * ldx;
  float acc[R];
#pragma unroll
  for (int j = 0; j < R; ++j) acc[j] = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + k);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float xv[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { xv[2 * i] = unpack_lo(w[i], dtype); xv[2 * i + 1] = unpack_hi(w[i], dtype); }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const float4 a0 = *reinterpret_cast<const float4*>(A + static_cast<long long>(j) * K + k);
      const float4 a1 = *reinterpret_cast<const float4*>(A + static_cast<long long>(j) * K + k + 4);
      acc[j] += xv[0] * a0.x + xv[1] * a0.y + xv[2] * a0.z + xv[3] * a0.w + xv[4] * a1.x + xv[5] * a1.y + xv[6] * a1.z + xv[7] * a1.w;
    }
  }
#pragma unroll
  for (int j = 0; j < R; ++j) acc[j] = warp_sum(acc[j]);
  float mine = 0.f;
#pragma unroll
  for (int j = 0; j < R; ++j) if (lane == j) mine = acc[j];
  xr[K + lane] = static_cast<uint16_t>(pack2(mine, 0.f, dtype) & 0xffff);
}

// ---------------------------------------------------------------- skinny wgrad: out[c, r] (+)= sum_m P[m, c] * Q[m, r]
// P half [M, C] (ldp), Q half [M, 8] (ldq): LoRA dB = dY^T xa, dA = dxa^T x (transposed store), fp32 atomics.
__global__ void __launch_bounds__(256) skinny_wgrad_kernel(const uint16_t* __restrict__ P, long long ldp,
                                                           const uint16_t* __restrict__ Q, long long ldq, int M, int C,
                                                           float* __restrict__ out, int transposed_out, int dtype,
                                                           int rows_per_block) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  // thread = (column group of 8 columns, row slice); block = 256 columns x rows_per_block rows.
  // Row slices are combined with shared-memory atomics (layout [i][r][cg]: conflict-free), then one global atomic
  // per output element and block.
  __shared__ float red[64 * 32];
  const int cg = threadIdx.x & 31, rs = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + cg * 8;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[i][r] = 0.f;
  for (int i = threadIdx.x; i < 64 * 32; i += 256) red[i] = 0.f;
  if (c0 < C) {
#pragma unroll 4
    for (int m = m0 + rs; m < m1; m += 8) {
      const uint4 pv = *reinterpret_cast<const uint4*>(P + static_cast<long long>(m) * ldp + c0);
      const uint4 qv = *reinterpret_cast<const uint4*>(Q + static_cast<long long>(m) * ldq);
      const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w}, qw[4] = {qv.x, qv.y, qv.z, qv.w};
      float pf[8], qf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pf[2 * i] = unpack_lo(pw[i], dtype); pf[2 * i + 1] = unpack_hi(pw[i], dtype);
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
    if (c < C) {
      if (transposed_out) atomicAdd(out + static_cast<long long>(r) * C + c, red[idx]);
      else atomicAdd(out + static_cast<long long>(c) * 8 + r, red[idx]);
    }
  }
}

// small-M variant (decoder steps): out += P^T Q without atomics.  A block owns 64 columns of P (lane -> 2 columns); its
// 8 warps split the M rows (all loads independent: the kernel is pure latency), then reduce through shared memory.
__global__ void __launch_bounds__(256) skinny_wgrad_small_kernel(const uint16_t* __restrict__ P, long long ldp,
                                                                 const uint16_t* __restrict__ Q, long long ldq, int M, int C,
                                                                 float* __restrict__ out, int transposed_out, int dtype) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  __shared__ float red[8][64 * 8 + 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 64 + lane * 2;
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[i][r] = 0.f;
  if (c < C) {
#pragma unroll 8
    for (int m = warp; m < M; m += 8) {
      const uint32_t pw = *reinterpret_cast<const uint32_t*>(P + static_cast<long long>(m) * ldp + c);
      const uint4 qv = *reinterpret_cast<const uint4*>(Q + static_cast<long long>(m) * ldq);
      const float p0 = unpack_lo(pw, dtype), p1 = unpack_hi(pw, dtype);
      const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float q0 = unpack_lo(qw[i], dtype), q1 = unpack_hi(qw[i], dtype);
        acc[0][2 * i] = fmaf(p0, q0, acc[0][2 * i]);     acc[0][2 * i + 1] = fmaf(p0, q1, acc[0][2 * i + 1]);
        acc[1][2 * i] = fmaf(p1, q0, acc[1][2 * i]);     acc[1][2 * i + 1] = fmaf(p1, q1, acc[1][2 * i + 1]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int r = 0; r < 8; ++r) red[warp][(lane * 2 + i) * 8 + r] = acc[i][r];
  __syncthreads();
  for (int idx = threadIdx.x; idx < 64 * 8; idx += 256) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][idx];
    const int cc = blockIdx.x * 64 + (idx >> 3), r = idx & 7;
    if (cc < C) {
      if (transposed_out) out[static_cast<long long>(r) * C + cc] += v;
      else out[static_cast<long long>(cc) * 8 + r] += v;
    }
  }
}

// ---------------------------------------------------------------- LoRA re-pack after an optimiser step
// For every LoRA-wrapped Linear j of a LoraGroup (mr_blip_b200/t5.py) the trainable fp32 A [8,K] / B [N,8] are copied as
// 16-bit values into their four slots: [W | sB] columns, B_down rows, A_down rows and [W^T | A^T] columns.  One launch
// for all Linears (descriptor table in device memory) instead of ~1700 strided copies per step.
struct LoraPackDesc {
  const float* A; const float* B;
  uint16_t* ext_slot; long long ld_ext;        // [N, 8]: ext_slot[n * ld + r]   = s * B[n, r]
  uint16_t* bdown_slot; long long ld_bdown;    // [8, N]: bdown_slot[r * ld + n] = s * B[n, r]
  uint16_t* adown_slot; long long ld_adown;    // [8, K]: adown_slot[r * ld + k] = A[r, k]
  uint16_t* extb_slot; long long ld_extb;      // [K, 8]: extb_slot[k * ld + r]  = A[r, k]
  long long K, N;
  double scale;
};
__global__ void __launch_bounds__(256) lora_pack_kernel(const LoraPackDesc* __restrict__ descs, int dtype) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const LoraPackDesc d = descs[blockIdx.x];
  const float s = static_cast<float>(d.scale);
  for (long long i = blockIdx.y * 256 + threadIdx.x; i < max(d.K, d.N); i += static_cast<long long>(gridDim.y) * 256) {
    if (i < d.N) {
      const float4 b0 = *reinterpret_cast<const float4*>(d.B + i * 8), b1 = *reinterpret_cast<const float4*>(d.B + i * 8 + 4);
      const float v[8] = {b0.x * s, b0.y * s, b0.z * s, b0.w * s, b1.x * s, b1.y * s, b1.z * s, b1.w * s};
      *reinterpret_cast<uint4*>(d.ext_slot + i * d.ld_ext) =
          make_uint4(pack2(v[0], v[1], dtype), pack2(v[2], v[3], dtype), pack2(v[4], v[5], dtype), pack2(v[6], v[7], dtype));
#pragma unroll
      for (int r = 0; r < 8; ++r) d.bdown_slot[r * d.ld_bdown + i] = static_cast<uint16_t>(pack2(v[r], 0.f, dtype) & 0xffff);
    }
    if (i < d.K) {
      float v[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = d.A[r * d.K + i];
#pragma unroll
      for (int r = 0; r < 8; ++r) d.adown_slot[r * d.ld_adown + i] = static_cast<uint16_t>(pack2(v[r], 0.f, dtype) & 0xffff);
      *reinterpret_cast<uint4*>(d.extb_slot + i * d.ld_extb) =
          make_uint4(pack2(v[0], v[1], dtype), pack2(v[2], v[3], dtype), pack2(v[4], v[5], dtype), pack2(v[6], v[7], dtype));
    }
  }
}

// ---------------------------------------------------------------- LoRA down-projection for small M
// out[m, r] = sum_k x[m, k] * W[r, k], r < 32 (16-bit x / W / out): one block per row.  Used when M is too small for the
// tensor-core path to spread over the SMs (decoder steps).  The block has ceil(K / 2048) x 256 threads (<= 1024), so a thread
// makes one pass (two for K = 10240) with its 32 W loads in flight together; a warp's 32 sums are folded across its lanes by a
// butterfly that halves the live values per step (31 shuffles, lane r ends with sum r) instead of 32 full warp reductions.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) small_down_kernel(const uint16_t* __restrict__ x, long long ldx,
                                                             const uint16_t* __restrict__ W, long long ldw, int K,
                                                             uint16_t* __restrict__ out, long long ldo, int dtype) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  __shared__ float red[THREADS / 32][33];
  const int m = blockIdx.x;
  constexpr int nwarps = THREADS / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) acc[r] = 0.f;
  for (int k = threadIdx.x * 8; k < K; k += THREADS * 8) {
    const uint4 xv = *reinterpret_cast<const uint4*>(x + static_cast<long long>(m) * ldx + k);
    const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
    float xf[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { xf[2 * i] = unpack_lo(xw[i], dtype); xf[2 * i + 1] = unpack_hi(xw[i], dtype); }
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const uint4 wv = *reinterpret_cast<const uint4*>(W + static_cast<long long>(r) * ldw + k);
      const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[r] += xf[2 * i] * unpack_lo(ww[i], dtype) + xf[2 * i + 1] * unpack_hi(ww[i], dtype);
    }
  }
  // butterfly: after the step with offset `off`, acc[i] (i < off) of a lane is a partial sum of output i + (lane & off ? off : 0) + ...
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? acc[i] : acc[i + off];
      const float keep = up ? acc[i + off] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  red[warp][lane] = acc[0];                  // lane r: this warp's sum for output r
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < nwarps; ++w) v += red[w][threadIdx.x];
    out[static_cast<long long>(m) * ldo + threadIdx.x] = static_cast<uint16_t>(pack2(v, 0.f, dtype) & 0xffff);
  }
}

// ---------------------------------------------------------------- casts / transpose / column sums
__global__ void cast_f32_to_h_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4, int dtype) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n4) return;
  const float4 v = in[i];
  out[i] = make_uint2(pack2(v.x, v.y, dtype), pack2(v.z, v.w, dtype));
}
// 2D strided cast: out[r, c] = in[r, c] for c < cols (ld_in / ld_out element strides)
__global__ void cast2d_f32_to_h_kernel(const float* __restrict__ in, long long ld_in, uint16_t* __restrict__ out,
                                       long long ld_out, int rows, int cols, int dtype) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int cv = cols >> 1;
  if (idx >= static_cast<long long>(rows) * cv) return;
  const int c = (idx % cv) * 2;
  const long long r = idx / cv;
  const float2 v = *reinterpret_cast<const float2*>(in + r * ld_in + c);
  *reinterpret_cast<uint32_t*>(out + r * ld_out + c) = pack2(v.x, v.y, dtype);
}
// out[c, r] = in[r, c]  (16-bit elements), 32x32 tiles through shared memory
__global__ void transpose16_kernel(const uint16_t* __restrict__ in, long long ld_in, uint16_t* __restrict__ out,
                                   long long ld_out, int rows, int cols) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  __shared__ uint16_t tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[static_cast<long long>(r) * ld_in + c] : 0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[static_cast<long long>(c) * ld_out + r] = tile[threadIdx.x][i];
  }
}
// out[c] (+)= sum_r in[r, c]   (fp32), used for the t5_proj bias gradient
__global__ void colsum_kernel(const float* __restrict__ in, int rows, int C, float* __restrict__ out, int rows_per_block) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += in[static_cast<long long>(r) * C + c];
  atomicAdd(out + c, s);
}
// y = a*x + b*y over fp32 vectors (grad accumulation / residual sums on the flat streams)
__global__ void axpby_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long n4, float a, float b) {
  mrb::pdl_trigger();
  mrb::pdl_wait();      // before the first global access (common.cuh, MRB_PDL)
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n4) return;
  const float4 xv = x[i];
  float4 yv = y[i];
  yv.x = a * xv.x + b * yv.x; yv.y = a * xv.y + b * yv.y; yv.z = a * xv.z + b * yv.z; yv.w = a * xv.w + b * yv.w;
  y[i] = yv;
}

}  // namespace mrb

using namespace mrb;
#define STREAM static_cast<cudaStream_t>(stream)
static inline unsigned blocks_for(long long n, int per) { return static_cast<unsigned>((n + per - 1) / per); }
// Row kernels (one warp per row): 8 rows per block, but decoder-sized inputs (64 rows) would then sit on 8 of the 148 SMs and run
// at the latency of 8 serial row passes per SM -- one row per block spreads them.
static inline int norm_warps(int rows) { return rows <= 2 * 148 ? 1 : 8; }
static inline bool small_rows_take_row_kernel() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MRB_NORM_ROW_SMALL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

extern "C" int mrb_norm(const float* x, const float* add, const float* w, const float* bias, float eps, int rows, int C,
                        int mode, float* out_f32, void* out_h, int h_dtype, long long ld_h, float* sum_out, void* stream) {
  if (rows <= 0) return MRB_OK;
  if ((C & 3) || C > 2048 || (out_h && (ld_h & 3))) return MRB_ERR_ARG;
  static int row_kernel = -1;             // MRB_NORM_ROW=0 keeps the warp-per-row kernel for large inputs (A/B measurements)
  if (row_kernel < 0) { const char* e = getenv("MRB_NORM_ROW"); row_kernel = (e && e[0] == '0') ? 0 : 1; }
  // decoder-sized inputs (64 rows x 2048): the one-warp-per-row kernel walks its 16 float4 per lane as dependent round trips
  // (10 / 18 us per launch forward / backward, profiles/launch_summary_r02f.csv); a block per row issues every load at once.
  // MRB_NORM_ROW_SMALL=0 restores the warp-per-row kernels for them (A/B measurements).
  if (row_kernel && (rows > 2 * 148 || (C >= 1024 && small_rows_take_row_kernel()))) {
    DropSpec none;
    none.seed = nullptr; none.site = 0; none.thr = 0; none.scale = 1.f;
    MRB_LAUNCH((norm_row_kernel), rows, 256, 0, STREAM, x, add, w, bias, eps, rows, C, mode, out_f32, out_h, h_dtype, ld_h, sum_out, none);
    MRB_CHECK_LAUNCH();
    return MRB_OK;
  }
  const int nw = norm_warps(rows);
  const unsigned grid = blocks_for(rows, nw);
  if (C <= 1024) MRB_LAUNCH((norm_kernel<8>), grid, 32 * nw, 0, STREAM, x, add, w, bias, eps, rows, C, mode, out_f32, out_h, h_dtype, ld_h, sum_out);
  else if (C <= 1536) MRB_LAUNCH((norm_kernel<12>), grid, 32 * nw, 0, STREAM, x, add, w, bias, eps, rows, C, mode, out_f32, out_h, h_dtype, ld_h, sum_out);
  else MRB_LAUNCH((norm_kernel<16>), grid, 32 * nw, 0, STREAM, x, add, w, bias, eps, rows, C, mode, out_f32, out_h, h_dtype, ld_h, sum_out);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_rmsnorm_bwd(const float* x, const float* w, const void* dy, int dy_dtype, long long ld_dy,
                               const float* lora_A, int R, float eps, int rows, int C, float* dres, void* stream) {
  if (rows <= 0) return MRB_OK;
  if ((C & 3) || C > 2048 || (ld_dy & 3) || R > 32 || (lora_A && dy_dtype == MRB_DT_F32)) return MRB_ERR_ARG;
  static int row_kernel = -1;             // MRB_RMSNORM_BWD_ROW=0 keeps the warp-per-row kernel for large inputs (A/B measurements)
  if (row_kernel < 0) { const char* e = getenv("MRB_RMSNORM_BWD_ROW"); row_kernel = (e && e[0] == '0') ? 0 : 1; }
  if (row_kernel && !lora_A && (rows > 2 * 148 || (C >= 1024 && small_rows_take_row_kernel()))) {
    DropSpec none;
    none.seed = nullptr; none.site = 0; none.thr = 0; none.scale = 1.f;
    MRB_LAUNCH((rmsnorm_bwd_row_kernel), rows, 256, 0, STREAM, x, w, dy, dy_dtype, ld_dy, eps, rows, C, dres, none,
               static_cast<uint16_t*>(nullptr), 0, 0LL);
    MRB_CHECK_LAUNCH();
    return MRB_OK;
  }
  MRB_LAUNCH((rmsnorm_bwd_kernel<16>), blocks_for(rows, norm_warps(rows)), 32 * norm_warps(rows), 0, STREAM, x, w, dy, dy_dtype, ld_dy, lora_A, R, eps, rows, C, dres);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

// sum_out = x + drop(add), out_h = RMSNorm(sum_out) * w: the residual add under train-mode dropout (modeling_t5.py:346,652,690)
// and the next sublayer's T5LayerNorm (:263-277) in one pass over the row -- mrb_dropout_add followed by mrb_norm(mode 1), same
// arithmetic in the same order (bit-identical), one launch and one read of the residual stream less.
extern "C" int mrb_dropout_add_norm(const float* x, const float* add, const float* w, float eps, int rows, int C, void* out_h,
                                    int h_dtype, long long ld_h, float* sum_out, const unsigned* seed, unsigned site, float p,
                                    void* stream) {
  if (rows <= 0) return MRB_OK;
  if (!x || !add || !w || !out_h || !sum_out || !seed || (C & 3) || C > 2048 || (ld_h & 3) || p < 0.f || p >= 1.f) return MRB_ERR_ARG;
  const DropSpec d = make_drop(seed, site, p);
  MRB_LAUNCH((norm_row_kernel), rows, 256, 0, STREAM, x, add, w, static_cast<const float*>(nullptr), eps, rows, C, 1,
             static_cast<float*>(nullptr), out_h, h_dtype, ld_h, sum_out, d);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

// mrb_rmsnorm_bwd (weights frozen, no LoRA fold) that also writes dy_next = drop_site(dres) as a 16-bit operand [rows, ld_next]:
// the gradient the NEXT sublayer's backward starts from (mrb_dropout fp32 -> 16 bit of the updated residual-stream gradient).
extern "C" int mrb_rmsnorm_bwd_drop(const float* x, const float* w, const void* dy, int dy_dtype, long long ld_dy, float eps, int rows,
                                    int C, float* dres, void* dy_next, int next_dtype, long long ld_next, const unsigned* seed,
                                    unsigned site, float p, void* stream) {
  if (rows <= 0) return MRB_OK;
  if ((C & 3) || C > 2048 || (ld_dy & 3) || (ld_next & 3) || !dy_next || !seed || p < 0.f || p >= 1.f ||
      (next_dtype != MRB_DT_F16 && next_dtype != MRB_DT_BF16))
    return MRB_ERR_ARG;
  const DropSpec d = make_drop(seed, site, p);
  MRB_LAUNCH((rmsnorm_bwd_row_kernel), rows, 256, 0, STREAM, x, w, dy, dy_dtype, ld_dy, eps, rows, C, dres, d,
             static_cast<uint16_t*>(dy_next), next_dtype, ld_next);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_lora_up_add(void* x_ext, long long ldx, const float* A, int R, int M, int K, int dtype, float* acc,
                               void* stream) {
  if (M <= 0) return MRB_OK;
  if ((K & 7) || (ldx & 7) || (R & 1) || R > 32) return MRB_ERR_ARG;
  MRB_LAUNCH((lora_up_add_kernel), blocks_for(static_cast<long long>(M) * (K >> 3), 256), 256, 0, STREAM, 
      static_cast<uint16_t*>(x_ext), ldx, A, R, M, K, dtype, acc);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_patchify(const float* img, void* out, int dtype, int frames, int img_size, int patch, int ldA, void* stream) {
  if (frames <= 0) return MRB_OK;
  if (img_size % patch || (patch & 1) || ldA < 3 * patch * patch || (ldA & 7)) return MRB_ERR_ARG;
  const int G = img_size / patch;
  const long long total = static_cast<long long>(frames) * G * G * 3 * patch;
  MRB_LAUNCH((patchify_kernel), blocks_for(total, 256), 256, 0, STREAM, img, out, dtype, frames, img_size, patch, ldA);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_patchify_u8(const unsigned char* img, void* out, int dtype, int frames, int img_size, int patch, int ldA,
                               float mean0, float mean1, float mean2, float std0, float std1, float std2, void* stream) {
  if (frames <= 0) return MRB_OK;
  if (img_size % patch || (patch & 1) || ldA < 3 * patch * patch || (ldA & 7)) return MRB_ERR_ARG;
  if (!(std0 > 0.f && std1 > 0.f && std2 > 0.f)) return MRB_ERR_ARG;
  const int G = img_size / patch;
  const long long total = static_cast<long long>(frames) * G * G * 3 * patch;
  NormConst nc{{mean0, mean1, mean2}, {std0, std1, std2}};
  MRB_LAUNCH((patchify_u8_kernel), blocks_for(total, 256), 256, 0, STREAM, img, out, dtype, frames, img_size, patch, ldA, nc);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_cls_pos(const float* cls, const float* pos, float* x, int frames, int tokens, int C, void* stream) {
  if (frames <= 0) return MRB_OK;
  MRB_LAUNCH((cls_pos_kernel), blocks_for(static_cast<long long>(frames) * C, 256), 256, 0, STREAM, cls, pos, x, frames, tokens, C);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_gated_gelu_fwd(const void* ab, void* h, int M, int F, long long ldh, int dtype, void* stream) {
  if (M <= 0) return MRB_OK;
  if ((F & 7) || (ldh & 7)) return MRB_ERR_ARG;
  MRB_LAUNCH((gated_gelu_fwd_kernel), blocks_for(static_cast<long long>(M) * (F >> 3), 256), 256, 0, STREAM, 
      static_cast<const uint4*>(ab), static_cast<uint4*>(h), M, F, ldh, dtype);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
extern "C" int mrb_gated_gelu_bwd(const void* ab, const void* dh, long long lddh, void* dab, long long lddab, int M, int F,
                                  int dtype, void* stream) {
  if (M <= 0) return MRB_OK;
  if ((F & 7) || (lddh & 7) || (lddab & 7) || lddab < 2 * F) return MRB_ERR_ARG;
  MRB_LAUNCH((gated_gelu_bwd_kernel), blocks_for(static_cast<long long>(M) * (F >> 3), 256), 256, 0, STREAM, 
      static_cast<const uint4*>(ab), static_cast<const uint4*>(dh), lddh, static_cast<uint4*>(dab), lddab, M, F, dtype);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_gather_rows(const int* idx, const float* emb, const float* frames, float* out, int rows, int C, void* stream) {
  if (rows <= 0) return MRB_OK;
  if (C & 3) return MRB_ERR_ARG;
  MRB_LAUNCH((gather_rows_kernel), rows, 256, 0, STREAM, idx, emb, frames, out, rows, C);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
extern "C" int mrb_scatter_frames(const int* idx, const float* dout, float* dframes, int rows, int C, void* stream) {
  if (rows <= 0) return MRB_OK;
  if (C & 3) return MRB_ERR_ARG;
  MRB_LAUNCH((scatter_frames_kernel), rows, 256, 0, STREAM, idx, dout, dframes, rows, C);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
extern "C" int mrb_group_mean(const float* x, float* out, int groups, int n, int C, void* stream) {
  if (groups <= 0) return MRB_OK;
  MRB_LAUNCH((group_mean_kernel), blocks_for(static_cast<long long>(groups) * C, 256), 256, 0, STREAM, x, out, groups, n, C);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
extern "C" int mrb_group_mean_bwd(const float* dout, float* dx, int groups, int n, int C, void* stream) {
  if (groups <= 0) return MRB_OK;
  MRB_LAUNCH((group_mean_bwd_kernel), blocks_for(static_cast<long long>(groups) * n * C, 256), 256, 0, STREAM, dout, dx, groups, n, C);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_cross_entropy(const float* logits, const long long* labels, int rows, int V, float* row_loss,
                                 void* dlogits, int d_dtype, long long ldd, float gscale, float* loss_sum, void* stream) {
  if (rows <= 0) return MRB_OK;
  if (dlogits && (ldd & 1)) return MRB_ERR_ARG;
  MRB_LAUNCH((ce_kernel), rows, 1024, 0, STREAM, logits, labels, rows, V, row_loss, dlogits, d_dtype, ldd, gscale, loss_sum);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_lora_down(void* x_ext, long long ldx, const float* A, int M, int K, int R, int dtype, void* stream) {
  if (M <= 0) return MRB_OK;
  if ((K & 255) || (ldx & 7) || ldx < K + 32) return MRB_ERR_ARG;
  const unsigned grid = blocks_for(M, 8);
  uint16_t* x = static_cast<uint16_t*>(x_ext);
  switch (R) {
    case 8: MRB_LAUNCH((lora_down_kernel<8>), grid, 256, 0, STREAM, x, ldx, A, M, K, dtype); break;
    case 16: MRB_LAUNCH((lora_down_kernel<16>), grid, 256, 0, STREAM, x, ldx, A, M, K, dtype); break;
    case 24: MRB_LAUNCH((lora_down_kernel<24>), grid, 256, 0, STREAM, x, ldx, A, M, K, dtype); break;
    default: return MRB_ERR_UNSUPPORTED;
  }
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_skinny_wgrad(const void* P, long long ldp, const void* Q, long long ldq, int M, int C, float* out,
                                int transposed_out, int dtype, void* stream) {
  if (M <= 0 || C <= 0) return MRB_OK;
  if ((ldq & 7) || (ldp & 7) || (C & 7) || (reinterpret_cast<uintptr_t>(P) & 15) || (reinterpret_cast<uintptr_t>(Q) & 15)) return MRB_ERR_ARG;
  if (M <= 256) {
    MRB_LAUNCH((skinny_wgrad_small_kernel), blocks_for(C, 64), 256, 0, STREAM, static_cast<const uint16_t*>(P), ldp,
                                                                    static_cast<const uint16_t*>(Q), ldq, M, C, out,
                                                                    transposed_out, dtype);
    MRB_CHECK_LAUNCH();
    return MRB_OK;
  }
  const int rows_per_block = 256;
  dim3 grid(blocks_for(C, 256), blocks_for(M, rows_per_block));
  MRB_LAUNCH((skinny_wgrad_kernel), grid, 256, 0, STREAM, static_cast<const uint16_t*>(P), ldp, static_cast<const uint16_t*>(Q), ldq,
                                                M, C, out, transposed_out, dtype, rows_per_block);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_lora_pack(const void* descs, int n, int blocks_per_linear, int dtype, void* stream) {
  if (n <= 0) return MRB_OK;
  if (blocks_per_linear <= 0 || (dtype != MRB_DT_F16 && dtype != MRB_DT_BF16)) return MRB_ERR_ARG;
  MRB_LAUNCH((lora_pack_kernel), dim3(n, blocks_per_linear), 256, 0, STREAM, static_cast<const LoraPackDesc*>(descs), dtype);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_small_down(const void* x, long long ldx, const void* W, long long ldw, int M, int K, void* out,
                              long long ldo, int dtype, void* stream) {
  if (M <= 0) return MRB_OK;
  if ((K & 7) || (ldx & 7) || (ldw & 7)) return MRB_ERR_ARG;
  static int wide = -1;                   // MRB_SMALL_DOWN_WIDE=0: 256 threads for every K (A/B measurements)
  if (wide < 0) { const char* e = getenv("MRB_SMALL_DOWN_WIDE"); wide = (e && e[0] == '0') ? 0 : 1; }
  int threads = 256;
  if (wide) { threads = ((K + 2047) / 2048) * 256; if (threads > 1024) threads = 1024; }
  const uint16_t* xp = static_cast<const uint16_t*>(x);
  const uint16_t* wp = static_cast<const uint16_t*>(W);
  uint16_t* op = static_cast<uint16_t*>(out);
  if (threads <= 256) MRB_LAUNCH((small_down_kernel<256>), M, 256, 0, STREAM, xp, ldx, wp, ldw, K, op, ldo, dtype);
  else if (threads <= 512) MRB_LAUNCH((small_down_kernel<512>), M, 512, 0, STREAM, xp, ldx, wp, ldw, K, op, ldo, dtype);
  else MRB_LAUNCH((small_down_kernel<1024>), M, 1024, 0, STREAM, xp, ldx, wp, ldw, K, op, ldo, dtype);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}

extern "C" int mrb_cast_f32_to_h(const float* in, void* out, long long n, int dtype, void* stream) {
  if (n <= 0) return MRB_OK;
  if (n & 3) return MRB_ERR_ARG;
  MRB_LAUNCH((cast_f32_to_h_kernel), blocks_for(n >> 2, 256), 256, 0, STREAM, reinterpret_cast<const float4*>(in),
                                                                    static_cast<uint2*>(out), n >> 2, dtype);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
extern "C" int mrb_cast2d_f32_to_h(const float* in, long long ld_in, void* out, long long ld_out, int rows, int cols,
                                   int dtype, void* stream) {
  if (rows <= 0 || cols <= 0) return MRB_OK;
  if ((cols & 1) || (ld_in & 1) || (ld_out & 1)) return MRB_ERR_ARG;
  MRB_LAUNCH((cast2d_f32_to_h_kernel), blocks_for(static_cast<long long>(rows) * (cols >> 1), 256), 256, 0, STREAM, 
      in, ld_in, static_cast<uint16_t*>(out), ld_out, rows, cols, dtype);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
extern "C" int mrb_transpose16(const void* in, long long ld_in, void* out, long long ld_out, int rows, int cols, void* stream) {
  if (rows <= 0 || cols <= 0) return MRB_OK;
  dim3 grid(blocks_for(cols, 32), blocks_for(rows, 32)), block(32, 8);
  MRB_LAUNCH((transpose16_kernel), grid, block, 0, STREAM, static_cast<const uint16_t*>(in), ld_in, static_cast<uint16_t*>(out), ld_out, rows, cols);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
extern "C" int mrb_colsum(const float* in, int rows, int C, float* out, void* stream) {
  if (rows <= 0) return MRB_OK;
  dim3 grid(blocks_for(C, 256), blocks_for(rows, 256));
  MRB_LAUNCH((colsum_kernel), grid, 256, 0, STREAM, in, rows, C, out, 256);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
extern "C" int mrb_axpby(const float* x, float* y, long long n, float a, float b, void* stream) {
  if (n <= 0) return MRB_OK;
  if (n & 3) return MRB_ERR_ARG;
  MRB_LAUNCH((axpby_kernel), blocks_for(n >> 2, 256), 256, 0, STREAM, reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n >> 2, a, b);
  MRB_CHECK_LAUNCH();
  return MRB_OK;
}
