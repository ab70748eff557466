/* oracle/ref_glue.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiled together with the UNMODIFIED reference sources where they lie
 * under $(RECUR_REF) (see oracle/Makefile) into oracle/_ref/.  The reference
 * keeps several pieces of the hot path as `static inline` helpers in headers
 * (badmaths.h, recur-nn-helpers.h, recur-rng.h, charmodel-helpers.h); they have
 * no linkable symbol, so this file gives each one an exported wrapper, and
 * adds drivers that replay the reference's own stream loops over the
 * reference's rnn_* API so tests and the CPU baseline can call them in one
 * piece through ctypes.
 *
 * Nothing here re-implements arithmetic: every number is produced by
 * reference code.
 */
#include "recur-nn.h"
#include "badmaths.h"
#include "recur-nn-helpers.h"
#include <math.h>
#include <time.h>

/* ---- header-only helpers, exported --------------------------------------- */

float ref_fast_expf(float x){ return fast_expf(x); }
float ref_fast_sigmoid(float x){ return fast_sigmoid(x); }
float ref_fast_tanhf(float x){ return fast_tanhf(x); }
float ref_soft_clip(float sum, float halfmax){ return soft_clip(sum, halfmax); }

void ref_softmax(float *dest, const float *src, int len){ softmax(dest, src, len); }

int ref_softmax_best_guess(float *error, const float *src, int len){
  return softmax_best_guess(error, src, len);
}

void ref_init_rand64(rand_ctx *ctx, u64 seed){ init_rand64(ctx, seed); }
u64 ref_rand64(rand_ctx *ctx){ return rand64(ctx); }
double ref_rand_double(rand_ctx *ctx){ return rand_double(ctx); }
float ref_cheap_gaussian_noise(rand_ctx *ctx){ return cheap_gaussian_noise(ctx); }
int ref_rand_small_int(rand_ctx *ctx, int cap){ return rand_small_int(ctx, cap); }

size_t ref_sizeof_RecurNN(void){ return sizeof(RecurNN); }
size_t ref_sizeof_RecurNNBPTT(void){ return sizeof(RecurNNBPTT); }
size_t ref_sizeof_RecurExtraLayer(void){ return sizeof(RecurExtraLayer); }

/* ---- the text-predict inner step, as charmodel-predict.c:18-27 and
        charmodel-helpers.h:16-33 spell it (one_hot_opinion + net_error_bptt),
        driven through the reference API ---------------------------------- */

static inline float
capped_log2f_(float x){ /* charmodel-helpers.h:11-14 */
  return (x < 1e-30f) ? -100.0f : log2f(x);
}

float
ref_one_hot_error(RecurNN *net, int c, int next, int *correct)
{
  float *inputs = net->bottom_layer ? net->bottom_layer->inputs : net->real_inputs;
  int len = net->bottom_layer ? net->bottom_layer->input_size : net->input_size;
  memset(inputs, 0, len * sizeof(float));
  inputs[c] = 1.0f;
  float *answer = rnn_opinion(net, NULL, net->presynaptic_noise);
  float *error = net->bptt->o_error;
  int winner = softmax_best_guess(error, answer, net->output_size);
  *correct = (winner == next);
  error[next] += 1.0f;
  return error[next];
}

/* The body of rnn_char_multi_cross_entropy's loop (charmodel-multi-predict.c:
   361-368) for output vectors that did not come from rnn_opinion on a
   reference net: per class group the reference's own softmax and capped log. */
void
ref_multi_entropy_step(const float *answer, int n_classes, int alphabet_len, int next,
    double *entropy)
{
  float error[alphabet_len];
  for (int j = 0; j < n_classes; j++){
    const float *group = answer + alphabet_len * j;
    softmax(error, group, alphabet_len);
    float e = error[next];
    entropy[j] -= capped_log2f_(e);
  }
}

/* Replays the synchronic multi-tap loop of rnn_char_epoch
   (charmodel-predict.c:288-311) for `steps` character positions starting at
   text position `start`.  Accumulates the same three report sums.  Returns
   elapsed seconds (CLOCK_MONOTONIC), which is what bench.py's CPU baseline
   reads. */
double
ref_multi_tap_train(RecurNN **nets, int n_nets, const u8 *text, int len,
    int start, int steps, int learning_style, float momentum,
    float momentum_soft_start,
    double *sum_error, double *sum_entropy, int *sum_correct)
{
  struct timespec t0, t1;
  RecurNN *net = nets[0];
  int spacing = (len - 1) / n_nets;
  float error = 0, entropy = 0;
  int correct = 0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int i = start;
  for (int s = 0; s < steps; s++, i++){
    if (i >= len - 1)
      i = 0;
    float m = rnn_calculate_momentum_soft_start(net->generation, momentum,
        momentum_soft_start);
    for (int j = 0; j < n_nets; j++){
      RecurNN *n = nets[j];
      int c;
      int offset = i + j * spacing;
      if (offset >= len - 1)
        offset -= len - 1;
      rnn_bptt_advance(n);
      float e = ref_one_hot_error(n, text[offset], text[offset + 1], &c);
      correct += c;
      error += e;
      entropy += capped_log2f_(1.0f - e);
      rnn_bptt_calc_deltas(n, j ? 1 : 0, NULL);
    }
    rnn_apply_learning(net, learning_style, m);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (sum_error) *sum_error = error;
  if (sum_entropy) *sum_entropy = entropy;
  if (sum_correct) *sum_correct = correct;
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* The single-net fast path of the same function (charmodel-predict.c:313-322). */
double
ref_single_net_train(RecurNN *net, const u8 *text, int len, int start, int steps,
    float momentum, float momentum_soft_start, uint batch_size,
    double *sum_error, double *sum_entropy, int *sum_correct)
{
  struct timespec t0, t1;
  float error = 0, entropy = 0;
  int correct = 0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int i = start;
  for (int s = 0; s < steps; s++, i++){
    int c;
    if (i >= len - 1)
      i = 0;
    net->bptt->momentum = rnn_calculate_momentum_soft_start(net->generation,
        momentum, momentum_soft_start);
    rnn_bptt_advance(net);
    float e = ref_one_hot_error(net, text[i], text[i + 1], &c);
    rnn_bptt_calculate(net, batch_size);
    correct += c;
    error += e;
    entropy += capped_log2f_(1.0f - e);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (sum_error) *sum_error = error;
  if (sum_entropy) *sum_entropy = entropy;
  if (sum_correct) *sum_correct = correct;
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* Forward-only steps/sec: rnn_opinion over a one-hot symbol stream. */
double
ref_opinion_steps(RecurNN *net, const u8 *text, int len, int steps)
{
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int s = 0; s < steps; s++){
    memset(net->real_inputs, 0, net->input_size * sizeof(float));
    net->real_inputs[text[s % len]] = 1.0f;
    rnn_opinion(net, NULL, 0);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ---- gstrnnca's fill_frame (gstrnnca.c:805-830) over a band of cells -------
   gstrnnca.c needs GStreamer and cannot be compiled here; its frame loop is
   rnn_opinion + fast_sigmoid_array + UNIT_TO_BYTE around fill_net_inputs.
   The first three are reference code, called here; fill_net_inputs is the
   oracle's restatement, handed in as a function pointer; UNIT_TO_BYTE is a
   one-line macro of gstrnnca.c (:642), spelled out below.  Cells first .. first + n - 1 of a
   w x h frame; clones[i] is the net of cell first + i.  Returns seconds. */
typedef void (*ref_fill_inputs_fn)(const unsigned char *frame, int w, int h, int cx, int cy,
    const int *offsets_y, int len_y, const int *offsets_c, int len_c, int len_pos, int edges,
    float *inputs);

double
ref_rnnca_cells(RecurNN **clones, int first, int n, const unsigned char *frame,
    unsigned char *frame_out, int w, int h, ref_fill_inputs_fn fill, const int *offsets_y,
    int len_y, const int *offsets_c, int len_c, int len_pos, int edges)
{
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int plane = w * h;
  for (int i = 0; i < n; i++){
    int cell = first + i;
    RecurNN *net = clones[i];
    fill(frame, w, h, cell % w, cell / w, offsets_y, len_y, offsets_c, len_c, len_pos, edges,
        net->real_inputs);
    float *answer = rnn_opinion(net, NULL, 0);
    fast_sigmoid_array(answer, answer, 3);
  }
  for (int i = 0; i < n; i++){
    float *yuv = clones[i]->output_layer;
    for (int k = 0; k < 3; k++){
      frame_out[k * plane + first + i] = (unsigned char)(yuv[k] * 255.9f); /* UNIT_TO_BYTE :642 */
    }
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
