/* oracle/ref_mfcc_glue.c - TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiled together with the UNMODIFIED reference mfcc.c (oracle/Makefile ->
 * oracle/_ref/libmfcc_ref.so).  It supplies the FFT mfcc.c expects from
 * GStreamer (see shim/gst/fft/gstfftf32.h) and a driver that runs the
 * reference's own recur_extract_log_freq_bins / recur_extract_mfccs over a
 * batch of windows, the way gstclassify's pcm_to_features does per channel
 * (gstclassify.c:1984-1995).
 */
#include "mfcc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

GstFFTF32 *
gst_fft_f32_new(int len, gboolean inverse)
{
  GstFFTF32 *f = calloc(1, sizeof(*f));
  f->len = len;
  f->inverse = inverse;
  return f;
}

void
gst_fft_f32_free(GstFFTF32 *self)
{
  free(self);
}

/* the DFT by its definition, in double */
void
gst_fft_f32_fft(GstFFTF32 *self, const float *t, GstFFTF32Complex *f)
{
  const int n = self->len;
  for (int j = 0; j <= n / 2; j++){
    double re = 0.0, im = 0.0;
    for (int k = 0; k < n; k++){
      /* reduce jk mod n exactly before the trig call */
      const double a = -2.0 * M_PI * (double)(((long long)j * k) % n) / n;
      re += t[k] * cos(a);
      im += t[k] * sin(a);
    }
    f[j].r = (float)re;
    f[j].i = (float)im;
  }
}

RecurAudioBinner *
ref_mfcc_new(int window_size, int window_type, int n_bins, float min_freq, float max_freq,
    float knee_freq, float focus_freq, float audio_rate, float scale, int value_size)
{
  return recur_audio_binner_new(window_size, window_type, n_bins, min_freq, max_freq, knee_freq,
      focus_freq, audio_rate, scale, value_size);
}

void
ref_mfcc_delete(RecurAudioBinner *ab)
{
  recur_audio_binner_delete(ab);
}

/* n_windows windows of window_size samples -> n_windows rows of n_bins floats */
void
ref_mfcc_extract(RecurAudioBinner *ab, const float *pcm, int n_windows, float *out, int dct)
{
  float *tmp = malloc((ab->window_size + 2) * sizeof(float));
  for (int w = 0; w < n_windows; w++){
    memcpy(tmp, pcm + (size_t)w * ab->window_size, ab->window_size * sizeof(float));
    float *row = dct ? recur_extract_mfccs(ab, tmp) : recur_extract_log_freq_bins(ab, tmp);
    memcpy(out + (size_t)w * ab->n_bins, row, ab->n_bins * sizeof(float));
  }
  free(tmp);
}

/* the set-up tables, for tests of the product's own set-up */
void
ref_mfcc_tables(RecurAudioBinner *ab, float *mask, int *left, int *right, float *left_fraction,
    float *right_fraction, float *slope)
{
  memcpy(mask, ab->mask, ab->window_size * sizeof(float));
  for (int i = 0; i <= ab->n_bins; i++){
    left[i] = ab->slopes[i].left;
    right[i] = ab->slopes[i].right;
    left_fraction[i] = ab->slopes[i].left_fraction;
    right_fraction[i] = ab->slopes[i].right_fraction;
    slope[i] = ab->slopes[i].slope;
  }
}
