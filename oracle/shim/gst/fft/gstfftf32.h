/* oracle/shim/gst/fft/gstfftf32.h - TEST INFRASTRUCTURE ONLY.
 *
 * The reference's mfcc.c (audio front end of gstclassify) takes
 * its real FFT from GStreamer (gst_fft_f32_*, a KISS FFT), which is not in this
 * image.  This header gives the unmodified mfcc.c the four names it uses; the
 * transform behind them (oracle/ref_mfcc_glue.c) is the discrete Fourier
 * transform by its definition, summed in double: N real samples in, N/2 + 1
 * complex bins out, unscaled, forward sign exp(-2 pi i jk/N) - what
 * gst_fft_f32_fft documents.  Any correct FFT differs from it by float
 * rounding only.
 */
#ifndef RB_SHIM_GSTFFTF32_H
#define RB_SHIM_GSTFFTF32_H

typedef int gboolean;
#ifndef FALSE
#define FALSE 0
#endif
#ifndef TRUE
#define TRUE 1
#endif
#ifndef G_PI
#define G_PI 3.1415926535897932384626433832795028841971693993751
#endif

typedef struct _GstFFTF32Complex {
  float r, i;
} GstFFTF32Complex;

typedef struct _GstFFTF32 {
  int len;
  gboolean inverse;
} GstFFTF32;

GstFFTF32 *gst_fft_f32_new(int len, gboolean inverse);
void gst_fft_f32_fft(GstFFTF32 *self, const float *timedata, GstFFTF32Complex *freqdata);
void gst_fft_f32_free(GstFFTF32 *self);

#endif
