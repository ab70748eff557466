/* rb_mfcc.cu - the audio front end of gstclassify on the device:
 * reference mfcc.c:9-94, recur_extract_log_freq_bins and recur_extract_mfccs,
 * for all channels' windows in one launch.
 *
 *   window (mfcc.c:57-74)  x[i] * mask[i]
 *   real FFT (:80-83)      N samples -> N/2 + 1 complex bins, unscaled
 *   mel bins (:9-52)       overlapping triangles described by n_bins + 1
 *                          slopes; every bin is the falling half of one slope
 *                          plus the rising half of the next; log(1 + power)
 *   DCT (:403-437)         recur_dct_cached's table walk, optional
 *
 * The set-up (window masks :272-305, slopes :134-178 with the iterative
 * mel -> Hz inverse :115-133) runs once on the host, in the reference's order
 * of operations: the tables decide which FFT bin a fraction belongs to, so
 * they have to come out the same.
 *
 * One block per window.  Everything lives in shared memory: the FFT is a
 * radix-2 pass structure over float2 with twiddles from a table computed in
 * double; a window is 1-2 KB, so the kernel is bound by reading the samples
 * once (4 bytes per sample in, 4 bytes per bin out).
 */
#include "rb_internal.h"
#include "rb_kernels.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MFCC_MAX_WINDOW 4096
#define MFCC_MAX_BINS 128

typedef struct MfccSlope {
  int left, right;
  float left_fraction, right_fraction, slope;
} MfccSlope;

struct RnnMfcc {
  int window_size, n_bins, window_type, log2n;
  float *mask_dev;       /* [window_size] */
  float2 *twiddle_dev;   /* [window_size / 2]: exp(-2 pi i k / N) */
  MfccSlope *slopes_dev; /* [n_bins + 1] */
  float *cos_dev;        /* [2 n_bins + 1]: recur_dct_cached's table */
  float *in_dev, *out_dev;
  size_t in_cap, out_cap;
  /* host copies, for inspection by tests */
  float *mask;
  MfccSlope *slopes;
};

#define MFCC_CUDA(call) do {                                            \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess)                                              \
      rb_die("recur-b200: %s failed at %s:%d: %s", #call, __FILE__,     \
          __LINE__, cudaGetErrorString(e_));                            \
  } while (0)

/* ---- set-up on the host ---------------------------------------------------
 * The tables decide which FFT bin a fraction of a triangle belongs to, so they
 * are computed with the reference's arithmetic (float where it uses float,
 * double where it uses double, same sequence): the warped mel scale
 * (mfcc.c:101-108), its inverse by damped fixed-point iteration (:115-133),
 * the slope edges (:134-178), the window masks (:272-305). */

static float
warped_mel(float hz, float knee, float focus)
{
  const float plain = 1127.0f * logf(1.0f + hz / knee);
  if (!focus)
    return plain;
  return plain / (1.0f + expf(3.0f * (1.0f - hz / focus)));
}

/* the frequency whose warped_mel is `target`: start from (target / 34)^2 and
   step by gain * residual, halving the gain whenever the residual changes
   sign; stop within 1e-4 mel or when the estimate no longer moves */
static float
warped_mel_inverse(float target, float knee, float focus)
{
  float hz = (target / 34) * (target / 34);
  float gain = 2.0f;
  float last = warped_mel(hz, knee, focus) - 1;
  for (;;) {
    const float now = warped_mel(hz, knee, focus);
    if (fabs(target - now) < 0.0001 || last == now)
      return hz;
    const float moved = hz + gain * (target - now);
    hz = moved > 0 ? moved : 0;
    if ((last > target) != (now > target))
      gain *= 0.5;
    last = now;
  }
}

/* n_bins + 1 slopes between n_bins + 2 edges equally spaced in warped mel;
   edges in units of FFT bins (fft_len * 2 / rate per Hz) */
static void
make_slopes(MfccSlope *slopes, int n_bins, int fft_len, float fmin, float fmax, float fknee,
    float ffocus, float audio_rate)
{
  const int n_slopes = n_bins + 1;
  const float mel_lo = warped_mel(fmin, fknee, ffocus);
  const float mel_hi = warped_mel(fmax, fknee, ffocus);
  const float mel_step = (mel_hi - mel_lo) / n_slopes;
  const float bins_per_hz = fft_len * 2 / audio_rate;
  float *edge = (float *)malloc((n_slopes + 1) * sizeof(float));
  float mel = mel_lo;
  edge[0] = fmin * bins_per_hz;
  for (int i = 1; i <= n_slopes; i++) {
    mel += mel_step;
    edge[i] = warped_mel_inverse(mel, fknee, ffocus) * bins_per_hz;
  }
  for (int i = 0; i < n_slopes; i++) {
    MfccSlope *s = &slopes[i];
    const float from = edge[i], to = edge[i + 1];
    s->left = (int)from;
    s->right = (int)to;
    s->slope = 1.0 / (to - from);
    if (s->left != s->right) {
      s->left_fraction = 1.0 - (from - s->left); /* what of bin `left` lies inside */
      s->right_fraction = to - s->right;
    }
    else { /* the whole triangle side inside one bin */
      s->left_fraction = to - from;
      s->right_fraction = 0;
    }
  }
  free(edge);
}

static void
make_window(float *mask, int len, int type, float scale)
{
  const double pi = 3.1415926535897932384626433832795028841971693993751;
  const double per_sample = pi / len;
  for (int i = 0; i < len; i++) {
    double w = 1.0;
    if (type == 1)                      /* Hann */
      w = 0.5 - 0.5 * cos(2.0 * per_sample * i);
    else if (type == 3)                 /* MP3's sine window */
      w = sin(per_sample * (i + 0.5f));
    else if (type == 2) {               /* Vorbis' power-complementary window */
      const double z = per_sample * (i + 0.5);
      w = sin(pi * 0.5 * sin(z) * sin(z));
    }
    mask[i] = (type >= 1 && type <= 3) ? (float)(w * scale) : 1.0f;
  }
}

__global__ void
k_mfcc(const float *__restrict__ pcm, float *__restrict__ out, const float *__restrict__ mask,
    const float2 *__restrict__ twiddle, const MfccSlope *__restrict__ slopes,
    const float *__restrict__ cos_lut, int N, int log2n, int n_bins, int windowed, int dct)
{
  extern __shared__ float2 z[];          /* [N] */
  float *power = (float *)(z + N);       /* [N / 2 + 1] */
  float *sum_left = power + N / 2 + 1;   /* [n_bins + 1] */
  float *bins = sum_left + n_bins + 1;   /* [n_bins] */
  const float *x = pcm + (size_t)blockIdx.x * N;
  /* window, into bit-reversed order */
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float v = windowed ? x[i] * mask[i] : x[i];
    z[__brev((unsigned)i) >> (32 - log2n)] = make_float2(v, 0.0f);
  }
  __syncthreads();
  /* decimation in time: log2 N passes of N / 2 butterflies */
  for (int s = 1; s <= log2n; s++) {
    const int half = 1 << (s - 1);
    for (int b = threadIdx.x; b < N / 2; b += blockDim.x) {
      const int k = b & (half - 1);
      const int i0 = ((b >> (s - 1)) << s) + k, i1 = i0 + half;
      const float2 w = twiddle[k << (log2n - s)];
      const float2 a = z[i0], c = z[i1];
      const float tr = c.x * w.x - c.y * w.y, ti = c.x * w.y + c.y * w.x;
      z[i0] = make_float2(a.x + tr, a.y + ti);
      z[i1] = make_float2(a.x - tr, a.y - ti);
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j <= N / 2; j += blockDim.x)
    power[j] = z[j].x * z[j].x + z[j].y * z[j].y;
  __syncthreads();
  /* recur_bin_complex (mfcc.c:9-52).  The reference carries sum_left from one
     slope into the next slope's sum_right; a slope's sum_left depends on that
     slope alone, so all of them first, then every bin in the reference's order
     of additions. */
  for (int i = threadIdx.x; i <= n_bins; i += blockDim.x) {
    const MfccSlope sl = slopes[i];
    int j = sl.left;
    float mul = sl.slope * sl.left_fraction;
    float p = power[j] * sl.left_fraction;
    float left = mul * p;
    if (sl.left != sl.right) {
      for (j = sl.left + 1; j < sl.right; j++) {
        mul += sl.slope;
        left += mul * power[j];
      }
    }
    mul += sl.slope * sl.right_fraction;
    p = power[j] * sl.right_fraction;
    left += mul * p;
    sum_left[i] = left;
  }
  __syncthreads();
  for (int i = 1 + threadIdx.x; i <= n_bins; i += blockDim.x) {
    const MfccSlope sl = slopes[i];
    int j = sl.left;
    float mul = sl.slope * sl.left_fraction;
    float p = power[j] * sl.left_fraction;
    float right = sum_left[i - 1] + (1.0f - mul) * p;
    if (sl.left != sl.right) {
      for (j = sl.left + 1; j < sl.right; j++) {
        mul += sl.slope;
        right += (1.0f - mul) * power[j];
      }
    }
    mul += sl.slope * sl.right_fraction;
    p = power[j] * sl.right_fraction;
    right += (1.0f - mul) * p;
    bins[i - 1] = logf(right + 1);
  }
  __syncthreads();
  float *row = out + (size_t)blockIdx.x * n_bins;
  if (!dct) {
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
      row[i] = bins[i];
    return;
  }
  /* recur_dct_cached (mfcc.c:403-437): the same walk over the same table */
  const int cos_len = 2 * n_bins;
  for (int j = threadIdx.x; j < n_bins; j += blockDim.x) {
    float a = 0.0f;
    int step = j * 2, i = j;
    for (int k = 0; k < n_bins; k++) {
      a += bins[k] * cos_lut[i];
      i += step;
      if (i > cos_len) {
        i = 2 * cos_len - i;
        step = -step;
      }
      else if (i < 0) {
        i = -i;
        step = -step;
      }
    }
    row[j] = j ? a : a * 0.7071067811865476f;
  }
}

/* recur_audio_binner_new (mfcc.c:308-337) */
extern "C" RnnMfcc *
rnn_mfcc_new(int window_size, int window_type, int n_bins, float min_freq, float max_freq,
    float knee_freq, float focus_freq, float audio_rate, float scale, int value_size)
{
  int log2n = 0;
  while ((1 << log2n) < window_size)
    log2n++;
  if (window_size < 16 || window_size > MFCC_MAX_WINDOW || (1 << log2n) != window_size ||
      n_bins < 1 || n_bins > MFCC_MAX_BINS || (value_size != 1 && value_size != 2)) {
    fprintf(stderr, "rnn_mfcc_new: windows are powers of two from 16 to %d samples, "
        "bins 1 to %d, value_size 1 or 2\n", MFCC_MAX_WINDOW, MFCC_MAX_BINS);
    return NULL;
  }
  RnnMfcc *m = (RnnMfcc *)calloc(1, sizeof(RnnMfcc));
  m->window_size = window_size;
  m->n_bins = n_bins;
  m->window_type = window_type;
  m->log2n = log2n;
  m->mask = (float *)malloc(window_size * sizeof(float));
  make_window(m->mask, window_size, window_type, scale);
  m->slopes = (MfccSlope *)calloc(n_bins + 1, sizeof(MfccSlope));
  make_slopes(m->slopes, n_bins, window_size / value_size, min_freq, max_freq, knee_freq,
      focus_freq, audio_rate);
  for (int i = 0; i <= n_bins; i++) {
    if (m->slopes[i].left < 0 || m->slopes[i].right > window_size / 2 ||
        m->slopes[i].left > m->slopes[i].right) {
      fprintf(stderr, "rnn_mfcc_new: the frequency range does not fit the window's %d bins\n",
          window_size / 2 + 1);
      free(m->mask);
      free(m->slopes);
      free(m);
      return NULL;
    }
  }
  return m;
}

/* the tables' way to the device, at the first extraction */
static void
mfcc_upload(RnnMfcc *m)
{
  if (m->mask_dev)
    return;
  rb_require_device("rnn_mfcc_extract");
  const int window_size = m->window_size, n_bins = m->n_bins;
  float2 *tw = (float2 *)malloc(window_size / 2 * sizeof(float2));
  const double pi = 3.14159265358979323846;
  for (int k = 0; k < window_size / 2; k++) {
    tw[k].x = (float)cos(-2.0 * pi * k / window_size);
    tw[k].y = (float)sin(-2.0 * pi * k / window_size);
  }
  const int cos_len = 2 * n_bins;
  float *lut = (float *)malloc((cos_len + 1) * sizeof(float));
  for (int j = 0; j <= cos_len; j++)
    lut[j] = cos(pi / cos_len * j);
  MFCC_CUDA(cudaMalloc((void **)&m->mask_dev, window_size * sizeof(float)));
  MFCC_CUDA(cudaMalloc((void **)&m->twiddle_dev, window_size / 2 * sizeof(float2)));
  MFCC_CUDA(cudaMalloc((void **)&m->slopes_dev, (n_bins + 1) * sizeof(MfccSlope)));
  MFCC_CUDA(cudaMalloc((void **)&m->cos_dev, (cos_len + 1) * sizeof(float)));
  MFCC_CUDA(cudaMemcpy(m->mask_dev, m->mask, window_size * sizeof(float), cudaMemcpyHostToDevice));
  MFCC_CUDA(cudaMemcpy(m->twiddle_dev, tw, window_size / 2 * sizeof(float2),
          cudaMemcpyHostToDevice));
  MFCC_CUDA(cudaMemcpy(m->slopes_dev, m->slopes, (n_bins + 1) * sizeof(MfccSlope),
          cudaMemcpyHostToDevice));
  MFCC_CUDA(cudaMemcpy(m->cos_dev, lut, (cos_len + 1) * sizeof(float), cudaMemcpyHostToDevice));
  free(tw);
  free(lut);
}

extern "C" void
rnn_mfcc_delete(RnnMfcc *m)
{
  if (!m)
    return;
  if (m->mask_dev)
    cudaStreamSynchronize(rb_stream);
  cudaFree(m->mask_dev);
  cudaFree(m->twiddle_dev);
  cudaFree(m->slopes_dev);
  cudaFree(m->cos_dev);
  cudaFree(m->in_dev);
  cudaFree(m->out_dev);
  free(m->mask);
  free(m->slopes);
  free(m);
}

static void
mfcc_launch(RnnMfcc *m, const float *pcm_dev, int n_windows, float *out_dev, int dct)
{
  mfcc_upload(m);
  const int N = m->window_size;
  const size_t sh = N * sizeof(float2) + (N / 2 + 1 + 2 * m->n_bins + 1) * sizeof(float);
  const int threads = N / 2 < 256 ? (N / 2 < 32 ? 32 : N / 2) : 256;
  if (sh > 48 * 1024)
    MFCC_CUDA(cudaFuncSetAttribute(k_mfcc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  k_mfcc<<<n_windows, threads, sh, rb_stream>>>(pcm_dev, out_dev, m->mask_dev, m->twiddle_dev,
      m->slopes_dev, m->cos_dev, N, m->log2n, m->n_bins, m->window_type != 0, dct);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    rb_die("recur-b200: launch of k_mfcc failed: %s", cudaGetErrorString(e));
  rb_count_launch(1);
}

/* n_windows windows of window_size samples (host) -> n_windows rows of n_bins
   floats (host): recur_extract_log_freq_bins (dct == 0) or recur_extract_mfccs */
extern "C" void
rnn_mfcc_extract(RnnMfcc *m, const float *pcm, int n_windows, float *out, int dct)
{
  if (n_windows < 1)
    return;
  mfcc_upload(m);
  const size_t in_bytes = (size_t)n_windows * m->window_size * sizeof(float);
  const size_t out_bytes = (size_t)n_windows * m->n_bins * sizeof(float);
  if (in_bytes > m->in_cap) {
    cudaFree(m->in_dev);
    MFCC_CUDA(cudaMalloc((void **)&m->in_dev, in_bytes));
    m->in_cap = in_bytes;
  }
  if (out_bytes > m->out_cap) {
    cudaFree(m->out_dev);
    MFCC_CUDA(cudaMalloc((void **)&m->out_dev, out_bytes));
    m->out_cap = out_bytes;
  }
  MFCC_CUDA(cudaMemcpyAsync(m->in_dev, pcm, in_bytes, cudaMemcpyHostToDevice, rb_stream));
  mfcc_launch(m, m->in_dev, n_windows, m->out_dev, dct);
  MFCC_CUDA(cudaMemcpyAsync(out, m->out_dev, out_bytes, cudaMemcpyDeviceToHost, rb_stream));
  MFCC_CUDA(cudaStreamSynchronize(rb_stream));
}

/* the same with both ends on the device (the feature rows can go straight
   into rnn_batch_set_inputs' staging), queued on the library stream */
extern "C" void
rnn_mfcc_extract_device(RnnMfcc *m, const float *pcm_dev, int n_windows, float *out_dev, int dct)
{
  if (n_windows >= 1)
    mfcc_launch(m, pcm_dev, n_windows, out_dev, dct);
}

/* the set-up tables (tests compare them with the reference's) */
extern "C" void
rnn_mfcc_tables(RnnMfcc *m, float *mask, int *left, int *right, float *left_fraction,
    float *right_fraction, float *slope)
{
  memcpy(mask, m->mask, m->window_size * sizeof(float));
  for (int i = 0; i <= m->n_bins; i++) {
    left[i] = m->slopes[i].left;
    right[i] = m->slopes[i].right;
    left_fraction[i] = m->slopes[i].left_fraction;
    right_fraction[i] = m->slopes[i].right_fraction;
    slope[i] = m->slopes[i].slope;
  }
}
