/* rb_devmath.cuh - the reference's small numeric helpers as device functions,
 * shared by rb_kernels.cu and rb_cells.cu.  Each keeps the reference's order of
 * operations (cited), because results are compared with it to the last bits. */
#ifndef RB_DEVMATH_CUH
#define RB_DEVMATH_CUH

/* recur-nn-helpers.h:104-113: 2x / (1 + x^2 (0.99 + x^2/100)), x = sum/halfmax */
__device__ __forceinline__ float
soft_clip_dev(float sum, float halfmax)
{
  if (halfmax == 0.0f)
    return sum;
  float x = sum / halfmax;
  float fudge = (float)(0.99 + (double)(x * x / 100.0f));
  return 2.0f * x / (1.0f + x * x * fudge);
}

/* badmaths.h:14-29: Pade(2,2) of exp on |x| < 0.2 after dividing by 8^count,
   then count rounds of three squarings.  The comparison in the reference is
   made in double against 0.2, which for a float means >= 0.2f. */
__device__ __forceinline__ float
fast_expf_dev(float x)
{
  int count = 0;
  while (fabsf(x) >= 0.2f && count < 48) {
    x *= 0.125f;
    count++;
  }
  float a = ((x + 3.0f) * (x + 3.0f) + 3.0f) / ((x - 3.0f) * (x - 3.0f) + 3.0f);
  while (count) {
    a *= a;
    a *= a;
    a *= a;
    count--;
  }
  return a;
}

/* get_offset_point, gstrnnca.c:644-667: the neighbour (dx, dy) of cell (cx, cy),
   clamped to the frame (edges) or wrapped around it */
__device__ __forceinline__ int
rnnca_offset_point(const int *off, int cx, int cy, int w, int h, int edges)
{
  int x = cx + off[0], y = cy + off[1];
  if (edges) {
    y = max(0, min(h - 1, y));
    x = max(0, min(w - 1, x));
  }
  else {
    if (y < 0)
      y += h;
    else if (y >= h)
      y -= h;
    if (x < 0)
      x += w;
    else if (x >= w)
      x -= w;
  }
  return y * w + x;
}

#endif
