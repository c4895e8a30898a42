// Train-mode dropout masks as a PURE FUNCTION of (step seed, site, row, column): the same function is evaluated by the forward
// kernel, by the backward kernels (which recompute the mask instead of storing it) and by the CPU oracle
// (oracle/dropout.py restates it in numpy), so train-mode parity is a bit-comparable test and not a statistical one.
//
//   key   = mix(seed + 0x9E3779B9 * (site + 1))                 one per launch (seed: a device word the host rewrites per step,
//                                                               so that a captured CUDA graph draws new masks on every replay)
//   word  = mix((row * ceil(cols / 4) + (col >> 2)) ^ key)      four 8-bit draws: columns 4g .. 4g+3 of `row`   (mod 2^32)
//   keep  = byte (col & 3) of word  >=  thr,   thr = round(256 p)
//   y     = keep ? x * 256 / (256 - thr) : 0                    (the drop rate is p quantised to 1/256: 26/256 for the reference's
//                                                               0.1, 13/256 for LoRA's 0.05; the scale uses the quantised rate, so
//                                                               E[y] = x exactly)
// "row" is the flat row index of the 2-D operand the site masks ([M, cols]); for attention probabilities it is
// (b * H + h) * Lq + i and cols = Lk.  mix() is the 32-bit finaliser "lowbias32" (two multiplies, three xor-shifts).
// The reference draws its masks from torch's Philox stream (nn.Dropout: modeling_t5.py:303,320,341,600,630,664,967;
// Qformer.py:66,135,283,369; peft lora_dropout, blip2_mr.py:197): same distribution per element up to the 1/256 quantisation,
// different stream -- no implementation outside torch can reproduce that stream, which is why parity is defined on this one.
// Compiles as plain C++ too (tests build a host harness from this header and compare it with the numpy restatement).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define MRB_HD __host__ __device__ __forceinline__
#else
#define MRB_HD inline
#endif

namespace mrb {

MRB_HD uint32_t drop_mix(uint32_t x) {
  x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15;
  return x;
}
MRB_HD uint32_t drop_key(uint32_t seed, uint32_t site) { return drop_mix(seed + 0x9E3779B9u * (site + 1u)); }
MRB_HD uint32_t drop_groups(uint32_t cols) { return (cols + 3u) >> 2; }
// draws of columns 4g .. 4g+3 of `row`; rowbase = row * drop_groups(cols)
MRB_HD uint32_t drop_word(uint32_t key, uint32_t rowbase, uint32_t g) { return drop_mix((rowbase + g) ^ key); }
MRB_HD bool drop_keep(uint32_t word, int col, uint32_t thr) { return ((word >> (8 * (col & 3))) & 0xffu) >= thr; }

// what a launch needs: the device word holding the step seed, the site id, the threshold and the keep scale
struct DropSpec {
  const uint32_t* seed;
  uint32_t site, thr;
  float scale;
};
inline DropSpec make_drop(const unsigned* seed, unsigned site, float p) {
  DropSpec d;
  d.seed = seed; d.site = site;
  int t = static_cast<int>(256.f * p + 0.5f);
  d.thr = static_cast<uint32_t>(t < 0 ? 0 : (t > 255 ? 255 : t));
  d.scale = 256.f / static_cast<float>(256u - d.thr);
  return d;
}

}  // namespace mrb
