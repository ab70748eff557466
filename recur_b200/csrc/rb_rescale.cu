/* rb_rescale.cu - recur_adaptive_downscale (reference rescale.c:88-256) on the
 * device: the box filter that shrinks every incoming video plane to the
 * automaton's size before the rnnca trainers read their targets from it
 * (remember_frame, gstrnnca.c:619-637).  Byte work, bit-exact.
 *
 * The reference walks the source with two Bresenham-style accumulators (rows,
 * then columns): a run of source rows is summed column by column into 16-bit
 * sums, and when the accumulator crosses 0x20000 the runs of columns are
 * summed, rounded and divided into one destination row.  Which source rows and
 * columns end up in which destination pixel depends only on the four sizes,
 * so the host replays the accumulators once into two small tables of runs
 * (the reference's integer arithmetic, including what it leaves unwritten at
 * the right and bottom when the steps round down), and the kernel does the
 * sums: a block per destination row, column sums in shared memory with
 * coalesced reads, then a thread per destination pixel.
 *
 * Three modes, as in the reference (rescale.c:240-256): shrinking by four or
 * more both ways uses every second row and column ("skipping"); equal sizes
 * are a plain copy; the rest is exact.
 */
#include "rb_internal.h"
#include "rb_kernels.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned char u8;

#define RS_CUDA(call) do {                                              \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess)                                              \
      rb_die("recur-b200: %s failed at %s:%d: %s", #call, __FILE__,     \
          __LINE__, cudaGetErrorString(e_));                            \
  } while (0)

/* runs[2 * k], runs[2 * k + 1]: first and one-past-last source index (in
   units of `stride` samples) of destination index k.  Replays the loop shared
   by consolidate_exact_row / consolidate_skipped_row and the row loops of
   recur_exact_downscale / recur_skipping_downscale: flush when acc >= 0x20000,
   then take the sample, then acc += step; one last flush for what is left if
   there is still a destination to put it in. */
static int
make_runs(int *runs, int n_src, int n_dst, int step, unsigned int acc)
{
  int k = 0, first = 0;
  for (int i = 0; i < n_src; i++) {
    if (acc >= 0x20000) {
      if (k >= n_dst)
        return -1; /* the reference would write past the destination */
      runs[2 * k] = first;
      runs[2 * k + 1] = i;
      k++;
      acc -= 0x20000;
      first = i;
    }
    acc += step;
  }
  if (k < n_dst && first < n_src) {
    runs[2 * k] = first;
    runs[2 * k + 1] = n_src;
    k++;
  }
  return k;
}

/* one block per destination row */
__global__ void __launch_bounds__(256)
k_downscale(const u8 *__restrict__ src, int s_stride, u8 *__restrict__ dst, int d_stride,
    const int *__restrict__ row_runs, const int *__restrict__ col_runs, int n_lanes, int n_cols,
    int sample)
{
  extern __shared__ unsigned int colsum[]; /* [n_lanes] */
  const int r = blockIdx.x;
  const int y0 = row_runs[2 * r], y1 = row_runs[2 * r + 1];
  /* the reference's temporary row holds 16-bit sums */
  for (int i = threadIdx.x; i < n_lanes; i += blockDim.x) {
    unsigned int s = 0;
    for (int y = y0; y < y1; y++)
      s += src[(size_t)(y * sample) * s_stride + i * sample];
    colsum[i] = s & 0xffffu;
  }
  __syncthreads();
  const unsigned int n_rows = (unsigned int)(y1 - y0);
  for (int j = threadIdx.x; j < n_cols; j += blockDim.x) {
    const int i0 = col_runs[2 * j], i1 = col_runs[2 * j + 1];
    unsigned int sum = 0;
    for (int i = i0; i < i1; i++)
      sum += colsum[i];
    const unsigned int n_samples = (unsigned int)(i1 - i0) * n_rows;
    sum += n_samples / 2;
    dst[(size_t)r * d_stride + j] = (u8)(sum / n_samples);
  }
}

static int *runs_dev = NULL;
static size_t runs_cap = 0;

/* both planes in device memory; queued on the library's stream.  Returns 0,
   or -1 (nothing done, one line on stderr) for what the reference itself does
   not handle: enlarging, and shrink factors whose 16-bit sums it lets
   overflow into each other. */
extern "C" int
rnn_b200_adaptive_downscale_device(const unsigned char *src, int s_width, int s_height,
    int s_stride, unsigned char *dst, int d_width, int d_height, int d_stride)
{
  rb_require_device("rnn_b200_adaptive_downscale");
  if (s_width < 1 || s_height < 1 || d_width < 1 || d_height < 1)
    return -1;
  if (s_width == d_width && s_height == d_height) {
    /* rescale.c:250-252: one memcpy of width * height bytes, strides ignored */
    RS_CUDA(cudaMemcpyAsync(dst, src, (size_t)s_width * s_height, cudaMemcpyDeviceToDevice,
            rb_stream));
    return 0;
  }
  if (d_width > s_width || d_height > s_height) {
    fprintf(stderr, "rnn_b200_adaptive_downscale: %dx%d -> %dx%d is not a downscale\n", s_width,
        s_height, d_width, d_height);
    return -1;
  }
  const int skipping = s_width >= d_width * 4 && s_height >= d_height * 4;
  const int sample = skipping ? 2 : 1;
  /* rescale.c:103-104 and :214-215 (int arithmetic, as there) */
  const int y_step = 0x20000 * sample * d_height / s_height;
  const int x_step = 0x20000 * sample * d_width / s_width;
  /* rows the loops visit: every one, or every second (rescale.c:224);
     columns: every byte, or the low byte of every 16-bit word (:163-178) */
  const int n_rows_src = skipping ? (s_height + 1) / 2 : s_height;
  const int n_lanes = skipping ? s_width / 2 : s_width;
  const unsigned int acc0_y = skipping ? y_step / 4 : y_step / 2;
  const unsigned int acc0_x = skipping ? x_step / 4 : x_step / 2;
  int *runs = (int *)malloc((size_t)2 * (d_height + d_width) * sizeof(int));
  const int n_out_rows = make_runs(runs, n_rows_src, d_height, y_step, acc0_y);
  const int n_out_cols = make_runs(runs + 2 * d_height, n_lanes, d_width, x_step, acc0_x);
  int longest = 0;
  for (int r = 0; r < n_out_rows; r++)
    if (runs[2 * r + 1] - runs[2 * r] > longest)
      longest = runs[2 * r + 1] - runs[2 * r];
  if (n_out_rows < 1 || n_out_cols < 1 || longest > 257) {
    fprintf(stderr, "rnn_b200_adaptive_downscale: %dx%d -> %dx%d is outside what the reference's "
        "16-bit sums hold\n", s_width, s_height, d_width, d_height);
    free(runs);
    return -1;
  }
  const size_t bytes = (size_t)2 * (d_height + d_width) * sizeof(int);
  if (bytes > runs_cap) {
    cudaFree(runs_dev);
    RS_CUDA(cudaMalloc((void **)&runs_dev, bytes));
    runs_cap = bytes;
  }
  /* (the tables are reused by the next call: wait for this stream's last use) */
  RS_CUDA(cudaStreamSynchronize(rb_stream));
  RS_CUDA(cudaMemcpyAsync(runs_dev, runs, bytes, cudaMemcpyHostToDevice, rb_stream));
  RS_CUDA(cudaStreamSynchronize(rb_stream));
  free(runs);
  const size_t sh = (size_t)n_lanes * sizeof(unsigned int);
  if (sh > 200 * 1024) {
    fprintf(stderr, "rnn_b200_adaptive_downscale: rows of %d samples do not fit\n", n_lanes);
    return -1;
  }
  if (sh > 48 * 1024)
    RS_CUDA(cudaFuncSetAttribute(k_downscale, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int)sh));
  k_downscale<<<n_out_rows, 256, sh, rb_stream>>>(src, s_stride, dst, d_stride, runs_dev,
      runs_dev + 2 * d_height, n_lanes, n_out_cols, sample);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    rb_die("recur-b200: launch of k_downscale failed: %s", cudaGetErrorString(e));
  rb_count_launch(1);
  return 0;
}

/* host planes in and out, like the reference's call (synchronises).  What the
   reference leaves unwritten in `dst` stays as it was. */
extern "C" int
rnn_b200_adaptive_downscale(const unsigned char *src, int s_width, int s_height, int s_stride,
    unsigned char *dst, int d_width, int d_height, int d_stride)
{
  rb_require_device("rnn_b200_adaptive_downscale");
  if (s_width < 1 || s_height < 1 || d_width < 1 || d_height < 1)
    return -1;
  const int same = s_width == d_width && s_height == d_height;
  const size_t s_bytes = same ? (size_t)s_width * s_height
      : (size_t)s_stride * (s_height - 1) + s_width;
  const size_t d_bytes = same ? s_bytes : (size_t)d_stride * (d_height - 1) + d_width;
  static u8 *s_dev = NULL, *d_dev = NULL;
  static size_t s_cap = 0, d_cap = 0;
  if (s_bytes > s_cap) {
    cudaFree(s_dev);
    RS_CUDA(cudaMalloc((void **)&s_dev, s_bytes));
    s_cap = s_bytes;
  }
  if (d_bytes > d_cap) {
    cudaFree(d_dev);
    RS_CUDA(cudaMalloc((void **)&d_dev, d_bytes));
    d_cap = d_bytes;
  }
  RS_CUDA(cudaMemcpyAsync(s_dev, src, s_bytes, cudaMemcpyHostToDevice, rb_stream));
  RS_CUDA(cudaMemcpyAsync(d_dev, dst, d_bytes, cudaMemcpyHostToDevice, rb_stream));
  int r = rnn_b200_adaptive_downscale_device(s_dev, s_width, s_height, s_stride, d_dev, d_width,
      d_height, d_stride);
  if (r == 0)
    RS_CUDA(cudaMemcpyAsync(dst, d_dev, d_bytes, cudaMemcpyDeviceToHost, rb_stream));
  RS_CUDA(cudaStreamSynchronize(rb_stream));
  return r;
}
