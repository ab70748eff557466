/* rb_split.cuh — the operand format of the tensor engine.
 *
 * Every FP32 operand a of the three big contractions is kept as TWO FP16
 * numbers ("planes"), written once where the operand is produced:
 *
 *     hi = fp16(s * a)               s: a power of two per plane set
 *     lo = fp16((s * a - hi) * 2048)
 *
 * so that s * a = hi + lo / 2048 to 22 significant bits (FP16 carries 11),
 * and lo sits 11 binades above the residual it stands for: values down to
 * 2^-25 / s keep a usable low part instead of vanishing into FP16's subnormal
 * range.  A product is three tcgen05.mma.kind::f16 into two FP32 accumulators
 * in tensor memory,
 *
 *     main += hi_a * hi_b        corr += hi_a * lo_b + lo_a * hi_b
 *     a.b   = (main + corr / 2048) / (s_a * s_b)
 *
 * (the dropped lo*lo term is 2^-22 relative).  Against TF32 planes this halves
 * the bytes per operand and doubles the MMA rate at the same 22 bits.
 *
 * Scales (powers of two, so exact): ring rows 1 (bounded by the input soft
 * clip, recur-nn.c:68-81); weights 64; error rows by the bound the top
 * layer's own soft clip puts on them (recur-nn.c:720: top <= 1.68 * 2 *
 * h_size, elements of E(0) <= top, elements of later rows <= sqrt(2 top + 1)
 * while the walk goes on, recur-nn.c:387).  Conversions saturate.
 */
#ifndef RB_SPLIT_CUH
#define RB_SPLIT_CUH

#include <stdint.h>

typedef unsigned short rb_h16; /* FP16 bits */

#define RB_LO_GAIN 2048.0f
#define RB_LO_UNGAIN (1.0f / 2048.0f)
#define RB_W_SCALE 64.0f

__device__ __forceinline__ rb_h16
rb_f2h(float a)
{
  rb_h16 h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(a));
  return h;
}

__device__ __forceinline__ float
rb_h2f(rb_h16 h)
{
  float f;
  asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
  return f;
}

/* a is already multiplied by the plane scale */
__device__ __forceinline__ void
rb_split_f16(float a, rb_h16 &hi, rb_h16 &lo)
{
  hi = rb_f2h(a);
  lo = rb_f2h((a - rb_h2f(hi)) * RB_LO_GAIN);
}

__device__ __forceinline__ float
rb_join_f16(rb_h16 hi, rb_h16 lo)
{
  return fmaf(rb_h2f(lo), RB_LO_UNGAIN, rb_h2f(hi));
}

__device__ __forceinline__ uint32_t
rb_pack2(rb_h16 a, rb_h16 b)
{
  return (uint32_t)a | ((uint32_t)b << 16);
}

/* two consecutive values (already multiplied by the plane scale) -> 4 bytes of
   each plane; the packed conversions round like rb_f2h */
__device__ __forceinline__ void
rb_split2(float a0, float a1, uint32_t &hi, uint32_t &lo)
{
  float f0, f1;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(a1), "f"(a0));
  asm("{\n\t"
      ".reg .b16 l, h;\n\t"
      "mov.b32 {l, h}, %2;\n\t"
      "cvt.f32.f16 %0, l;\n\t"
      "cvt.f32.f16 %1, h;\n\t"
      "}" : "=f"(f0), "=f"(f1) : "r"(hi));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;"
      : "=r"(lo) : "f"((a1 - f1) * RB_LO_GAIN), "f"((a0 - f0) * RB_LO_GAIN));
}

/* four consecutive values -> 8 bytes of each plane */
__device__ __forceinline__ void
rb_split4(float4 a, float scale, uint2 &hi, uint2 &lo)
{
  rb_split2(a.x * scale, a.y * scale, hi.x, lo.x);
  rb_split2(a.z * scale, a.w * scale, hi.y, lo.y);
}

#endif
