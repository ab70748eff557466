/* rb_rng.h — the random streams recur's nets carry.
 *
 * A saved net stores its generator state (recur-nn-io.c:89 "net.rng") and
 * weight initialisation, cloning (sub-seeds), random damage and presynaptic
 * noise all draw from it, so the sequences have to match the reference's
 * draw for draw or seeds stop reproducing.  The generator is Bob Jenkins'
 * 64-bit "small noncryptographic PRNG" (public domain,
 * burtleburtle.net/bob/rand/smallprng.html), which is what reference
 * recur-rng.h:24-43 implements; the derived distributions follow
 * recur-rng.h:62-101 (doubles, small ints) and recur-rng.h:179-200 (the
 * Irwin-Hall "cheap gaussian").
 *
 * Usable from C, C++ and CUDA device code.
 */
#ifndef RB_RNG_H
#define RB_RNG_H

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define RB_HD __host__ __device__ __forceinline__
#else
#define RB_HD static inline
#endif

typedef struct rb_rng_state {
  uint64_t a, b, c, d;
} rb_rng_state;

#define RB_ROTL64(x, k) (((x) << (k)) | ((x) >> (64 - (k))))

RB_HD uint64_t
rb_rng_next(rb_rng_state *s)
{
  uint64_t e = s->a - RB_ROTL64(s->b, 7);
  s->a = s->b ^ RB_ROTL64(s->c, 13);
  s->b = s->c + RB_ROTL64(s->d, 37);
  s->c = s->d + e;
  s->d = e + s->a;
  return s->d;
}

RB_HD void
rb_rng_seed(rb_rng_state *s, uint64_t seed)
{
  s->a = 0xf1ea5eedull;
  s->b = s->c = s->d = seed;
  for (int i = 0; i < 20; i++)
    (void)rb_rng_next(s);
}

/* [0, 1): 52 random mantissa bits under exponent 0, minus one
   (recur-rng.h:62-75) */
RB_HD double
rb_rng_double(rb_rng_state *s)
{
  uint64_t bits = (rb_rng_next(s) & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
  double d;
#ifdef __CUDA_ARCH__
  d = __longlong_as_double((long long)bits);
#else
  memcpy(&d, &bits, sizeof(d));
#endif
  return d - 1.0;
}

/* recur-rng.h:93-98 */
RB_HD int
rb_rng_small_int(rb_rng_state *s, int cap)
{
  return (int)(rb_rng_double(s) * cap);
}

/* Sum of twelve 16-bit uniforms taken from three draws, centred and scaled
   to unit variance (recur-rng.h:179-200). */
RB_HD float
rb_rng_cheap_gaussian(rb_rng_state *s)
{
  int64_t acc = 0;
  for (int draw = 0; draw < 3; draw++) {
    uint64_t r = rb_rng_next(s);
    acc += (int64_t)(r & 0xffff) + (int64_t)((r >> 16) & 0xffff) +
      (int64_t)((r >> 32) & 0xffff) + (int64_t)(r >> 48);
  }
  return (float)(acc - 0xffff * 6) / (0xffff);
}

#endif
