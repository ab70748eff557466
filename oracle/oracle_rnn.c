/* oracle/oracle_rnn.c — TEST INFRASTRUCTURE ONLY (never linked into or called
 * by the product; see oracle/__init__.py).
 *
 * A plain-C restatement of recur's RNN-core hot path, written from the
 * algorithm's description in SURVEY.md §8a ("batched restatement") and the
 * reference lines cited at each function.  It works on bare arrays, one
 * stream at a time, in straightforward loops (no row skipping, no vector
 * types), so that it can serve as an independent checker of the CUDA path
 * at sizes it finishes in seconds.
 *
 * PINNING: tests/test_oracle.py checks every function here against the
 * unmodified reference compiled in place (oracle/_ref/librecur_ref_strict.so,
 * IEEE build) on seeded inputs, and against the golden vectors under
 * tests/golden/ that were generated from that reference
 * (tests/golden/make_golden.py).  The reference's own tests hold no numeric
 * vectors for this path, so those two are what pins it.
 *
 * Layouts (reference recur-nn-init.c:110-126, SURVEY.md §8 "Memory layout"):
 *   x row      [i_size]  = [1 | hidden(t-1)[1..hs] | inputs[is] | 0 pad]
 *   hidden     [h_size]  = [1 | h[1..hs] | 0 pad]
 *   Wih        [i_size][h_size]  row = source node, column = destination
 *   Who        [h_size][o_size]
 *   history    [depth][i_size]   ring of x rows; `index` is the newest
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct OracleDims {
  int i_size, h_size, o_size;
  int input_size, hidden_size, output_size;
} OracleDims;

enum { ACT_RELU = 1, ACT_RESQRT = 2, ACT_RECLIP20 = 5 };

/* badmaths.h:14-29 */
float
oracle_fast_expf(float x)
{
  int count = 0;
  while ((double)fabsf(x) > 0.2) {
    x = (float)(x * 0.125);
    count++;
  }
  float num = (x + 3.0f) * (x + 3.0f) + 3.0f;
  float den = (x - 3.0f) * (x - 3.0f) + 3.0f;
  float a = num / den;
  for (; count; count--) {
    a *= a;
    a *= a;
    a *= a;
  }
  return a;
}

/* recur-nn-helpers.h:104-113 */
float
oracle_soft_clip(float sum, float halfmax)
{
  if (halfmax == 0)
    return sum;
  float x = sum / halfmax;
  float fudge = (float)(0.99 + (double)(x * x / 100));
  return 2.0f * x / (1 + x * x * fudge);
}

/* badmaths.h:71-141 + charmodel-predict.c:18-27: err = onehot(target) - p.
   Returns err[target]; *winner = first index of the largest p. */
float
oracle_softmax_error(const float *y, int len, int target, float *err, int *winner)
{
  float mx = y[0], mn = y[0];
  for (int i = 1; i < len; i++) {
    if (y[i] > mx) mx = y[i];
    if (y[i] < mn) mn = y[i];
  }
  float adj = 0.0f;
  if (mx > 50.0f)
    adj = 50.0f - mx;
  else if (mn < -60.0f)
    adj = (-60.0f - mn < 50.0f - mx) ? -60.0f - mn : 50.0f - mx;
  float sum = 0.0f;
  for (int i = 0; i < len; i++) {
    err[i] = oracle_fast_expf(y[i] + adj);
    sum += err[i];
  }
  int best = 0;
  float best_p = -1.0f;
  for (int i = 0; i < len; i++) {
    float p = err[i] / sum;
    if (p > best_p) {
      best_p = p;
      best = i;
    }
    err[i] = -p;
  }
  if (winner)
    *winner = best;
  if (target >= 0) {
    err[target] += 1.0f;
    return err[target];
  }
  return 0.0f;
}

/* recur-nn.c:18-48 as a dense product: out[x] = sum_y in[y] * W[y][x] */
static void
interlayer(const float *in, int n_in, float *out, int n_out, const float *W)
{
  for (int x = 0; x < n_out; x++)
    out[x] = 0.0f;
  for (int y = 0; y < n_in; y++) {
    float v = in[y];
    for (int x = 0; x < n_out; x++)
      out[x] += v * W[(size_t)y * n_out + x];
  }
}

/* rnn_opinion, recur-nn.c:83-154, for a net without bottom layer and with
   the inputs already written into x[hs+1 ...].  `noise` (h_size floats or
   NULL) is added to hidden[1..] before the activation (recur-nn.c:120). */
void
oracle_forward(const OracleDims *d, const float *Wih, const float *Who, float *x,
    float *hidden, float *out, int activation, const float *noise)
{
  const int hs1 = d->hidden_size + 1;
  memcpy(x, hidden, hs1 * sizeof(float));
  x[0] = 1.0f;
  /* recur-nn.c:68-81 */
  float softclip = d->i_size * 16.0f;
  float sum = 0.0f;
  for (int i = 0; i < d->i_size; i++)
    sum += x[i];
  if (sum > softclip) {
    float scale = oracle_soft_clip(sum, softclip);
    for (int i = 0; i < d->i_size; i++)
      x[i] *= scale;
  }
  interlayer(x, d->i_size, hidden, d->h_size, Wih);
  if (noise) {
    for (int i = 1; i < d->h_size; i++)
      hidden[i] += noise[i];
  }
  if (activation == ACT_RESQRT) {
    for (int i = 0; i < d->h_size; i++)
      hidden[i] = (hidden[i] > 0.0f) ? sqrtf(hidden[i] + 1.0f) - 1.0f : 0.0f;
  }
  else if (activation == ACT_RECLIP20) {
    for (int i = 1; i < d->h_size; i++) {
      float h = hidden[i];
      h = h < 20.0f ? h : 20.0f;
      hidden[i] = (h > 0.0f) ? h : 0.0f;
    }
  }
  else {
    for (int i = 1; i < d->h_size; i++)
      hidden[i] = (hidden[i] > 0.0f) ? hidden[i] : 0.0f;
  }
  hidden[0] = 1.0f;
  interlayer(hidden, d->h_size, out, d->o_size, Who);
}

typedef struct OracleBpttResult {
  float top_raw, top_scaled, err_sum, ih_scale, min_error_factor;
  float cum_error, min_error_sum;
  int n_steps, t_left;
} OracleBpttResult;

/* rnn_bptt_calc_deltas, recur-nn.c:707-772, with bptt_and_accumulate_error
   (303-450), backprop_single_layer (199-228), single_layer_sgd (256-273) and
   the delta fold (734-748), for one stream, no bottom layer, no error ranges.
   ih_delta and ho_delta are ADDED to (zero them first for accumulate == 0):
   ih_delta += ih_scale * sum_k x_k^T E_k,  ho_delta += hidden^T o_error.
   `scratch` needs 3 * i_size + ih_size floats. */
void
oracle_calc_deltas(const OracleDims *d, const float *Wih, const float *Who,
    const float *history, int depth, int index, const float *hidden,
    const float *o_error, float learn_rate, float min_error_factor, int adaptive,
    int activation, float *ih_delta, float *ho_delta, float *scratch,
    OracleBpttResult *res)
{
  const int I = d->i_size, H = d->h_size, O = d->o_size;
  const int hs1 = d->hidden_size + 1;
  float *h_error = scratch;
  float *i_error = scratch + I;
  float *G = scratch + 3 * (size_t)I;
  memset(h_error, 0, 2 * (size_t)I * sizeof(float));
  memset(G, 0, (size_t)I * H * sizeof(float));

  /* top layer */
  float top = 0.0f;
  for (int y = 1; y < H; y++) {
    float e = 0.0f;
    if (hidden[y] != 0.0f) {
      for (int x = 0; x < O; x++)
        e += Who[(size_t)y * O + x] * o_error[x];
      top += fabsf(e);
    }
    h_error[y] = e;
  }
  float top_scaled = top;
  float halfmax = H * 2.0f;
  if (top > halfmax) {
    float scale = oracle_soft_clip(top, halfmax);
    for (int y = 0; y < H; y++)
      h_error[y] *= scale;
    top_scaled = scale * top;
  }
  for (int y = 0; y < H; y++) {
    if (hidden[y] != 0.0f)
      for (int x = 0; x < O; x++)
        ho_delta[(size_t)y * O + x] += o_error[x] * hidden[y];
  }

  /* the walk back through the ring */
  float max_error_sum = 2.0f * top_scaled + 1;
  float ceiling = 1.0f * top_scaled;
  float min_error_gain = 1e-8f * top_scaled;
  float min_error_sum = min_error_factor / learn_rate;
  if (min_error_gain < min_error_sum)
    min_error_sum = min_error_gain;
  float error_sum = 0.0f, cum_error = 0.0f;
  int offset = index;
  int t, n_steps = 0;
  for (t = depth; t > 0; t--, offset += offset ? -1 : depth - 1) {
    const float *x = history + (size_t)offset * I;
    error_sum = 0.0f;
    h_error[0] = 0.0f;
    for (int i = hs1; i < H; i++)
      h_error[i] = 0.0f;
    for (int y = 0; y < I; y++) {
      float in = x[y];
      float e = 0.0f;
      if (in != 0.0f && (activation != ACT_RECLIP20 || in < 20.0f)) {
        for (int c = 0; c < H; c++) {
          G[(size_t)y * H + c] += h_error[c] * in;
          e += Wih[(size_t)y * H + c] * h_error[c];
        }
        if (activation == ACT_RESQRT)
          e /= 2 * (in + 1.0f);
        error_sum += e * e;
      }
      i_error[y] = e;
    }
    cum_error += sqrtf(error_sum);
    n_steps++;
    float *tmp = h_error;
    h_error = i_error;
    i_error = tmp;
    if (error_sum <= min_error_sum || error_sum > max_error_sum)
      break;
  }
  float ih_scale = 1.0f;
  if (error_sum > ceiling) {
    ih_scale = oracle_soft_clip(error_sum, max_error_sum);
  }
  else if (adaptive) {
    int depth_error = depth / 4 - t;
    if (min_error_factor < 1e-2f && (min_error_gain != min_error_sum || depth_error < 0))
      min_error_factor = (float)(min_error_factor * (1.0f + depth_error * 1e-3));
    if (min_error_factor < 1e-20f)
      min_error_factor = 1e-20f;
  }
  for (size_t i = 0; i < (size_t)I * H; i++)
    ih_delta[i] += G[i] * ih_scale;
  res->top_raw = top;
  res->top_scaled = top_scaled;
  res->err_sum = error_sum;
  res->ih_scale = ih_scale;
  res->min_error_factor = min_error_factor;
  res->cum_error = cum_error;
  res->min_error_sum = min_error_sum;
  res->n_steps = n_steps;
  res->t_left = t;
}

/* rnn_apply_learning's elementwise bodies, recur-nn.c:454-593.
   method: 0 weighted (also simplified Nesterov / classical through
   momentum_weight), 1 Nesterov, 4 adagrad, 5 adadelta, 6 rprop. */
void
oracle_apply_learning(int method, float *w, const float *delta, float *mom,
    float *aux, int size, float rate, float momentum, float momentum_weight)
{
  for (int i = 0; i < size; i++) {
    float dl = delta[i];
    if (method == 1) {
      float t = dl * rate;
      w[i] += t;
      mom[i] += t;
      mom[i] *= momentum;
      w[i] += mom[i];
    }
    else if (method == 4) {
      float a = mom[i] + dl * dl;
      w[i] += dl * rate / sqrtf(a);
      mom[i] = a;
    }
    else if (method == 5) {
      float renewal = 1.0f - momentum;
      float g = mom[i] * momentum, s = aux[i] * momentum;
      g += fabsf(dl) * renewal + rate;
      float step = s / g * dl;
      s += fabsf(step) * renewal + rate;
      mom[i] = g;
      aux[i] = s;
      w[i] += step;
    }
    else if (method == 6) {
      float max_step = rate, min_step = (float)(1e-6 * rate);
      float p = mom[i], step = aux[i];
      if (dl * p > 0.0f) {
        step = step * 1.2f;
        if (step > max_step) step = max_step;
      }
      else if (dl * p < 0.0f) {
        step = step * 0.5f;
        if (step < min_step) step = min_step;
        dl = 0;
      }
      if (dl > 0.0f) w[i] += step; else w[i] -= step;
      aux[i] = step;
      mom[i] = dl;
    }
    else {
      float t = dl * rate;
      float m = mom[i];
      w[i] += t + m * momentum_weight;
      mom[i] = (m + t) * momentum;
    }
  }
}

/* recur-nn.c:595-599 */
float
oracle_momentum_soft_start(float generation, float max_momentum, float x)
{
  float m = 1.0f - x / (1.0f + generation + 2.0f * x);
  return max_momentum < m ? max_momentum : m;
}

/* ---- a whole synchronic training set on bare arrays ------------------------ */

typedef struct OracleSet {
  OracleDims d;
  int n, depth, activation, adaptive;
  float *Wih, *Who, *ih_mom, *ho_mom, *ih_delta, *ho_delta;
  float *history; /* [n][depth][i_size] */
  float *hidden;  /* [n][h_size] */
  float *out;     /* [n][o_size] */
  float *o_error; /* [n][o_size] */
  int *index;     /* [n] */
  float *mef;     /* [n] */
  float *lr;      /* [n] */
  float *scratch;
  float ho_scale, momentum_weight;
  uint32_t generation;
} OracleSet;

static int
align4(int n)
{
  return (n + 3) & ~3;
}

/* weights are copied in from ih/ho arrays in the padded layout */
OracleSet *
oracle_set_new(int input_size, int hidden_size, int output_size, int n, int depth,
    float learn_rate, int activation, int adaptive, const float *Wih, const float *Who)
{
  OracleSet *s = calloc(1, sizeof(*s));
  OracleDims *d = &s->d;
  d->input_size = input_size;
  d->hidden_size = hidden_size;
  d->output_size = output_size;
  d->i_size = align4(hidden_size + input_size + 1);
  d->h_size = align4(hidden_size + 1);
  d->o_size = align4(output_size);
  size_t ih = (size_t)d->i_size * d->h_size, ho = (size_t)d->h_size * d->o_size;
  s->n = n;
  s->depth = depth;
  s->activation = activation;
  s->adaptive = adaptive;
  s->Wih = malloc(ih * sizeof(float));
  s->Who = malloc(ho * sizeof(float));
  memcpy(s->Wih, Wih, ih * sizeof(float));
  memcpy(s->Who, Who, ho * sizeof(float));
  s->ih_mom = calloc(ih, sizeof(float));
  s->ho_mom = calloc(ho, sizeof(float));
  s->ih_delta = calloc(ih, sizeof(float));
  s->ho_delta = calloc(ho, sizeof(float));
  s->history = calloc((size_t)n * depth * d->i_size, sizeof(float));
  s->hidden = calloc((size_t)n * d->h_size, sizeof(float));
  s->out = calloc((size_t)n * d->o_size, sizeof(float));
  s->o_error = calloc((size_t)n * d->o_size, sizeof(float));
  s->index = calloc(n, sizeof(int));
  s->mef = malloc(n * sizeof(float));
  s->lr = malloc(n * sizeof(float));
  s->scratch = malloc((3 * (size_t)d->i_size + ih) * sizeof(float));
  for (int j = 0; j < n; j++) {
    s->index[j] = 1; /* rnn_new advances once, recur-nn-init.c:133 */
    s->mef[j] = 1e-12f * d->h_size;
    s->lr[j] = learn_rate;
  }
  s->ho_scale = 1.0f;
  s->momentum_weight = 0.5f;
  return s;
}

void
oracle_set_delete(OracleSet *s)
{
  free(s->Wih); free(s->Who); free(s->ih_mom); free(s->ho_mom);
  free(s->ih_delta); free(s->ho_delta); free(s->history); free(s->hidden);
  free(s->out); free(s->o_error); free(s->index); free(s->mef); free(s->lr);
  free(s->scratch);
  free(s);
}

float *oracle_set_wih(OracleSet *s){ return s->Wih; }
float *oracle_set_who(OracleSet *s){ return s->Who; }
float *oracle_set_ih_delta(OracleSet *s){ return s->ih_delta; }
float *oracle_set_ho_delta(OracleSet *s){ return s->ho_delta; }
float *oracle_set_hidden(OracleSet *s){ return s->hidden; }
float *oracle_set_out(OracleSet *s){ return s->out; }
float *oracle_set_o_error(OracleSet *s){ return s->o_error; }
float *oracle_set_mef(OracleSet *s){ return s->mef; }

/* One position of the multi-tap loop (charmodel-predict.c:293-311), WEIGHTED
   style.  Adds to the three report sums. */
void
oracle_set_char_step(OracleSet *s, const uint8_t *cur, const uint8_t *next,
    float momentum, double *sum_error, double *sum_entropy, int *sum_correct)
{
  const OracleDims *d = &s->d;
  size_t ih = (size_t)d->i_size * d->h_size, ho = (size_t)d->h_size * d->o_size;
  memset(s->ih_delta, 0, ih * sizeof(float));
  memset(s->ho_delta, 0, ho * sizeof(float));
  for (int j = 0; j < s->n; j++) {
    s->index[j] = (s->index[j] + 1) % s->depth;
    float *ring = s->history + (size_t)j * s->depth * d->i_size;
    float *x = ring + (size_t)s->index[j] * d->i_size;
    float *in = x + d->hidden_size + 1;
    memset(in, 0, d->input_size * sizeof(float));
    in[cur[j]] = 1.0f;
    float *hid = s->hidden + (size_t)j * d->h_size;
    float *out = s->out + (size_t)j * d->o_size;
    float *err = s->o_error + (size_t)j * d->o_size;
    oracle_forward(d, s->Wih, s->Who, x, hid, out, s->activation, NULL);
    int winner;
    float e = oracle_softmax_error(out, d->output_size, next[j], err, &winner);
    if (sum_correct) *sum_correct += (winner == next[j]);
    if (sum_error) *sum_error += e;
    if (sum_entropy) {
      float p = 1.0f - e;
      *sum_entropy += (p < 1e-30f) ? -100.0f : log2f(p);
    }
    OracleBpttResult r;
    oracle_calc_deltas(d, s->Wih, s->Who, ring, s->depth, s->index[j], hid, err,
        s->lr[j], s->mef[j], s->adaptive, s->activation, s->ih_delta, s->ho_delta,
        s->scratch, &r);
    s->mef[j] = r.min_error_factor;
  }
  oracle_apply_learning(0, s->Who, s->ho_delta, s->ho_mom, NULL, (int)ho,
      s->lr[0] * s->ho_scale, momentum, s->momentum_weight);
  oracle_apply_learning(0, s->Wih, s->ih_delta, s->ih_mom, NULL, (int)ih,
      s->lr[0], momentum, s->momentum_weight);
  s->generation++;
}

/* `steps` positions over a text, as rnn_char_epoch walks it; returns seconds */
double
oracle_set_text_train(OracleSet *s, const uint8_t *text, int len, int start, int steps,
    float momentum, float momentum_soft_start, double *sum_error, double *sum_entropy,
    int *sum_correct)
{
  struct timespec t0, t1;
  uint8_t *cur = malloc(s->n), *next = malloc(s->n);
  int spacing = (len - 1) / s->n;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int i = start;
  for (int k = 0; k < steps; k++, i++) {
    if (i >= len - 1)
      i = 0;
    for (int j = 0; j < s->n; j++) {
      int off = i + j * spacing;
      if (off >= len - 1)
        off -= len - 1;
      cur[j] = text[off];
      next[j] = text[off + 1];
    }
    float m = oracle_momentum_soft_start(s->generation, momentum, momentum_soft_start);
    oracle_set_char_step(s, cur, next, m, sum_error, sum_entropy, sum_correct);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(cur);
  free(next);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}


/* ---- f4: gstrnnca's neighbourhood gather ------------------------------------
 * Restated from reference gstrnnca.c:644-667 (get_offset_point) and :670-691
 * (fill_net_inputs).  gstrnnca.c itself needs GStreamer and cannot be compiled
 * here, so for these forty lines parity is UNPINNED by the live reference: the
 * arithmetic is index clamping / wrapping and a scale by 1/255.  The sigmoid
 * that follows IS the reference's (oracle/_ref exports badmaths.h's
 * fast_sigmoid as ref_fast_sigmoid). */

static int
rnnca_offset_point(const int *offset, int cx, int cy, int w, int h, int edges)
{
  int x = cx + offset[0];
  int y = cy + offset[1];
  if (edges) {
    y = y < 0 ? 0 : (y > h - 1 ? h - 1 : y);
    x = x < 0 ? 0 : (x > w - 1 ? w - 1 : x);
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

void
oracle_rnnca_fill_inputs(const uint8_t *frame, int w, int h, int cx, int cy,
    const int *offsets_y, int len_y, const int *offsets_c, int len_c, int len_pos, int edges,
    float *inputs)
{
  const uint8_t *Y = frame, *Cb = frame + w * h, *Cr = frame + 2 * w * h;
  int i = 0;
  for (int j = 0; j < len_y; j++) {
    int off = rnnca_offset_point(offsets_y + j * 2, cx, cy, w, h, edges);
    inputs[i++] = Y[off] * (1.0f / 255.0f);
  }
  for (int j = 0; j < len_c; j++) {
    int off = rnnca_offset_point(offsets_c + j * 2, cx, cy, w, h, edges);
    inputs[i] = Cb[off] * (1.0f / 255.0f);
    inputs[i + 1] = Cr[off] * (1.0f / 255.0f);
    i += 2;
  }
  float xx = cx * 1.0f / w;
  float yy = cy * 1.0f / h;
  inputs[i] = xx;
  inputs[i + 1] = yy;
  if (len_pos == 3)
    inputs[i + 2] = 0.5 - ((yy - 0.5) * (yy - 0.5) + (xx - 0.5) * (xx - 0.5));
}

/* UNIT_TO_BYTE, gstrnnca.c:642 */
uint8_t
oracle_rnnca_unit_to_byte(float x)
{
  return (uint8_t)(x * 255.9f);
}
