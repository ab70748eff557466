/* rb_init.c — host-side weight initialisation and weight surgery.
 *
 * These run once (or rarely) and consume the net's random stream, so they
 * stay on the CPU and follow the reference draw for draw: the same
 * generator calls in the same order give the same weights for the same
 * seed (SURVEY.md §8a "init stays host-side CPU (RNG-sequence-exact)").
 * They write straight into the managed weight matrices; the next compute
 * call prefetches those back to the GPU (rb_matrices_to_device).
 *
 * Reference: recur-nn-init.c:352-735 (initialisation), recur-nn.c:857-883
 * (weight noise), recur-nn.c:1027-1145 (gain scaling, diagonal zapping),
 * recur-nn-helpers.h:84-102 (perforation).
 */
#include "rb_internal.h"
#include "rb_host.h"
#include "rb_rng.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* rand_ctx and rb_rng_state are the same four words */
#define RNG(ctx) ((rb_rng_state *)(ctx))

void
rb_init_rand64_maybe_randomly(rand_ctx *ctx, u64 seed)
{
  if (seed == RECUR_RNG_RANDOM_SEED) { /* recur-rng.h:45-56 */
    struct timespec t;
    clock_gettime(CLOCK_REALTIME, &t);
    seed = (u64)(((u64)t.tv_nsec << 20) + t.tv_sec) ^ (u64)((uintptr_t)ctx);
    fprintf(stderr, "seeding with %llx\n\n", (unsigned long long)seed);
  }
  rb_rng_seed(RNG(ctx), seed);
}

u64 rb_rand64(rand_ctx *x){ return rb_rng_next(RNG(x)); }
float rb_cheap_gaussian_noise(rand_ctx *ctx){ return rb_rng_cheap_gaussian(RNG(ctx)); }
double rb_rand_double(rand_ctx *ctx){ return rb_rng_double(RNG(ctx)); }
int rb_rand_small_int(rand_ctx *ctx, int cap){ return rb_rng_small_int(RNG(ctx), cap); }

/* badmaths.h:14-29 on the host (used by the log-normal initialisers) */
float
rb_fast_expf(float x)
{
  int count = 0;
  while (fabsf(x) > 0.2) {
    x *= 0.125;
    count++;
  }
  float a = ((x + 3) * (x + 3) + 3) / ((x - 3) * (x - 3) + 3);
  while (count) {
    a *= a;
    a *= a;
    a *= a;
    count--;
  }
  return a;
}

static inline int
rand_int_between(rand_ctx *rng, int start, int cap)
{
  return start + rb_rand_small_int(rng, cap - start);
}

/* ---- flat: every weight from one distribution, some left at zero ---------
   recur-nn-init.c:505-591 */

static void
fill_flat(rand_ctx *rng, float *m, int width, int height, int stride, int offset,
    float variance, rnn_init_distribution shape, double perforation)
{
  const float stddev = sqrtf(variance);
  for (int y = 0; y < height; y++) {
    float *row = m + (size_t)y * stride;
    for (int x = offset; x < width + offset; x++) {
      if (perforation != 0 && !(rb_rand_double(rng) > perforation))
        continue;
      switch (shape) {
      case RNN_INIT_DIST_UNIFORM: {
        const double range = sqrtf(12.0f * variance);
        row[x] = range * rb_rand_double(rng) - range * 0.5;
      } break;
      case RNN_INIT_DIST_LOG_NORMAL: {
        float a = rb_cheap_gaussian_noise(rng) * 0.33;
        float b = 0.9 * stddev * rb_fast_expf(a);
        row[x] = (rb_rand64(rng) & 1) ? b : -b;
      } break;
      case RNN_INIT_DIST_SEMICIRCLE: {
        double a, b;
        do {
          a = rb_rand_double(rng) * 2.0 - 1.0;
          b = rb_rand_double(rng);
        } while (a * a + b * b > 1.0);
        row[x] = stddev * 2 * a;
      } break;
      case RNN_INIT_DIST_GAUSSIAN:
      default:
        row[x] = stddev * rb_cheap_gaussian_noise(rng);
        break;
      }
    }
  }
}

static void
init_flat(RecurNN *net, float variance, rnn_init_distribution shape, double perforation)
{
  memset(net->ih_weights, 0, net->ih_size * sizeof(float));
  memset(net->ho_weights, 0, net->ho_size * sizeof(float));
  if (perforation < 0)
    perforation = 0;
  else if (perforation >= 1.0)
    return;
  fill_flat(&net->rng, net->ih_weights, net->hidden_size,
      net->input_size + net->hidden_size + 1, net->h_size, 1, variance, shape, perforation);
  fill_flat(&net->rng, net->ho_weights, net->output_size, net->hidden_size + 1,
      net->o_size, 0, variance, shape, perforation);
  if (net->bottom_layer) {
    RecurExtraLayer *bl = net->bottom_layer;
    memset(bl->weights, 0, (size_t)bl->i_size * bl->o_size * sizeof(float));
    fill_flat(&net->rng, bl->weights, bl->output_size, bl->input_size, bl->o_size, 1,
        variance, shape, perforation);
  }
}

/* ---- fan-in: each destination node gets inputs summing to about `sum` ----
   recur-nn-init.c:593-644 */

static void
fill_fan_in(rand_ctx *rng, float *m, int width, int height, int stride, float sum,
    float kurtosis, float margin)
{
  for (int x = 0; x < width; x++) {
    float remainder = sum + margin;
    for (int i = 0; i < height * 2 && remainder > margin; i++) {
      int y = rb_rand_small_int(rng, height);
      float *w = m + (size_t)y * stride + x;
      if (*w == 0) {
        float v = (rb_rand_double(rng) * 2 - 1) * remainder * kurtosis;
        *w += v;
        remainder -= fabsf(v);
      }
    }
  }
}

static void
init_fan_in(RecurNN *net, float sum, float kurtosis, float margin, float inputs_ratio)
{
  memset(net->ih_weights, 0, net->ih_size * sizeof(float));
  memset(net->ho_weights, 0, net->ho_size * sizeof(float));
  int hsize = 1 + net->hidden_size;
  if (inputs_ratio > 0) {
    fill_fan_in(&net->rng, net->ih_weights + 1, net->hidden_size, hsize, net->h_size,
        sum, kurtosis, margin);
    fill_fan_in(&net->rng, net->ih_weights + (size_t)hsize * net->h_size + 1,
        net->hidden_size, net->input_size, net->h_size, sum * inputs_ratio, kurtosis, margin);
  }
  else {
    fill_fan_in(&net->rng, net->ih_weights + 1, net->hidden_size,
        hsize + net->input_size, net->h_size, sum, kurtosis, margin);
  }
  fill_fan_in(&net->rng, net->ho_weights, net->output_size, net->hidden_size,
      net->o_size, sum, kurtosis, margin);
  if (net->bottom_layer) {
    RecurExtraLayer *bl = net->bottom_layer;
    memset(bl->weights, 0, (size_t)bl->i_size * bl->o_size * sizeof(float));
    fill_fan_in(&net->rng, bl->weights, bl->output_size, bl->input_size + 1, bl->o_size,
        sum, kurtosis, margin);
  }
}

/* ---- runs: chains / loops of hidden nodes ---------------------------------
   recur-nn-init.c:384-503 */

static float
log_normal_random_sign(rand_ctx *rng, float mean, float stddev, float bound)
{
  float x;
  do {
    x = rb_cheap_gaussian_noise(rng);
  } while (fabsf(x) > bound);
  float w = mean * rb_fast_expf(x * stddev);
  return (rb_rand64(rng) & 1) ? w : -w;
}

static void
link_random_input(RecurNN *net, int dest, float deviation)
{
  int input = rand_int_between(&net->rng, 0, net->input_size);
  net->ih_weights[(size_t)(net->hidden_size + 1 + input) * net->h_size + dest] =
    rb_cheap_gaussian_noise(&net->rng) * deviation;
}

static void
link_hidden(RecurNN *net, int from, int to, float gain, float input_probability,
    float input_magnitude)
{
  float weight = log_normal_random_sign(&net->rng, gain, 0.25, 3.0);
  net->ih_weights[(size_t)from * net->h_size + to] = weight;
  if (rb_rand_double(&net->rng) < input_probability)
    link_random_input(net, to, input_magnitude);
}

static void
init_runs(RecurNN *net, int n_loops, int len_mean, int len_stddev, float gain,
    float input_probability, float input_magnitude, int loop, int crossing_paths,
    int inputs_miss, int input_at_start)
{
  fprintf(stderr, "n_loops %d len_mean %d, len_stddev %d, gain %g, "
      "input_probability %g, input_magnitude %g loop %d crossing_paths %d,"
      " inputs_miss %d input_at_start %d\n",
      n_loops, len_mean, len_stddev, gain, input_probability, input_magnitude,
      loop, crossing_paths, inputs_miss, input_at_start);
  const int bound = net->hidden_size + 1;
  int *unused = (int *)malloc(bound * sizeof(int));
  int i = bound;
  int sum = 0;
  double linked_input_p = inputs_miss ? 0 : input_probability;
  double missing_input_p = inputs_miss ? input_probability : 0;

  for (int count = 0; count < n_loops; count++) {
    int len = rb_cheap_gaussian_noise(&net->rng) * len_stddev + len_mean + 0.5;
    if (len < 2)
      len = 2;
    if (len > net->hidden_size)
      len = net->hidden_size;
    if (i + len + inputs_miss >= bound || crossing_paths) {
      for (int k = 0; k < bound; k++)
        unused[k] = k;
      i = 1;
    }
    int j = rand_int_between(&net->rng, i, bound);
    int e = unused[j];
    const int beginning = e;
    if (input_at_start && input_magnitude)
      link_random_input(net, e, input_magnitude);
    for (int m = 0; m < len; m++, i++) {
      unused[j] = unused[i];
      int s = e;
      if (crossing_paths == 2) {
        e = rand_int_between(&net->rng, 1, bound);
      }
      else {
        j = rand_int_between(&net->rng, i, bound);
        e = unused[j];
      }
      link_hidden(net, s, e, gain, linked_input_p, input_magnitude);
    }
    if (loop)
      link_hidden(net, e, beginning, gain, linked_input_p, input_magnitude);
    if (rb_rand_double(&net->rng) < missing_input_p && i < bound) {
      j = rand_int_between(&net->rng, i, bound);
      e = unused[j];
      unused[j] = unused[i];
      i++;
      link_random_input(net, e, input_magnitude);
    }
    sum += len;
  }
  fprintf(stderr, "mean loop len %3g\n", (double)sum / n_loops);
  free(unused);
}

/* recur-nn-init.c:648-671 */
static void
runs_prepare_with_submethod(RecurNN *net, struct RecurInitialisationParameters *p)
{
  if (p->submethod != p->method) {
    p->method = p->submethod;
    rnn_randomise_weights_clever(net, p);
    p->method = RNN_INIT_RUNS;
  }
  float *mem = net->ih_weights;
  size_t rows = p->inputs_use_submethod ? net->h_size : net->i_size;
  if (p->bias_uses_submethod) {
    rows--;
    mem += net->h_size;
  }
  memset(mem, 0, rows * net->h_size * sizeof(float));
}

/* recur-nn-init.c:674-710 */
void
rnn_randomise_weights_clever(RecurNN *net, struct RecurInitialisationParameters *p)
{
  rb_host_will_touch_matrices(net);
  if (p->method == RNN_INIT_ZERO) {
    memset(net->ih_weights, 0, net->ih_size * sizeof(float));
    memset(net->ho_weights, 0, net->ho_size * sizeof(float));
  }
  else if (p->method == RNN_INIT_FAN_IN) {
    init_fan_in(net, p->fan_in_sum, p->fan_in_step, p->fan_in_min, p->fan_in_ratio);
  }
  else if (p->method == RNN_INIT_FLAT) {
    init_flat(net, p->flat_variance, p->flat_shape, p->flat_perforation);
  }
  else if (p->method == RNN_INIT_RUNS) {
    runs_prepare_with_submethod(net, p);
    init_runs(net, p->run_n, p->run_len_mean, p->run_len_stddev, p->run_gain,
        p->run_input_probability, p->run_input_magnitude, p->run_loop,
        p->run_crossing_paths, p->run_inputs_miss, p->run_input_at_start);
  }
  rb_weights_changed(net);
}

/* recur-nn-init.c:712-748 */
void
rnn_init_default_weight_parameters(RecurNN *net, struct RecurInitialisationParameters *q)
{
  memset(q, 0, sizeof(*q));
  q->method = RNN_INIT_FLAT;
  q->submethod = RNN_INIT_FLAT;
  q->bias_uses_submethod = 0;
  q->inputs_use_submethod = 0;

  q->fan_in_ratio = net->input_size * 1.0f / net->hidden_size;
  q->fan_in_sum = 3.0;
  q->fan_in_step = 0.3;
  q->fan_in_min = 0.1;

  q->flat_variance = RNN_INITIAL_WEIGHT_VARIANCE_FACTOR / net->h_size;
  q->flat_shape = RNN_INIT_DIST_UNIFORM;
  q->flat_perforation = 0.7;

  q->run_input_probability = .17;
  q->run_input_magnitude = 0.2;
  q->run_gain = 0.17;
  q->run_len_mean = net->hidden_size / 1.0;
  q->run_len_stddev = net->hidden_size / 3.0f;
  q->run_n = net->h_size * 0.085;
  q->run_loop = 1;
  q->run_crossing_paths = 0;
  q->run_inputs_miss = 0;
  q->run_input_at_start = 0;
}

void
rnn_randomise_weights_simple(RecurNN *net, const rnn_init_method method)
{
  struct RecurInitialisationParameters p;
  rnn_init_default_weight_parameters(net, &p);
  p.method = method;
  rnn_randomise_weights_clever(net, &p);
}

void
rnn_randomise_weights_auto(RecurNN *net)
{
  rnn_randomise_weights_simple(net, RNN_INIT_FLAT);
}

/* ---- weight surgery ---------------------------------------------------------- */

/* recur-nn-helpers.h:84-102 */
static void
perforate(float *a, int len, float dropout, rand_ctx *rng)
{
  int i;
  if (dropout == 0.5f) {
    for (i = 0; i < len;) {
      u64 bits = rb_rand64(rng);
      int end = i + ((len - i < 64) ? len - i : 64);
      for (; i < end; i++)
        a[i] = (bits & 1) ? a[i] : 0;
    }
  }
  else {
    for (i = 0; i < len; i++)
      a[i] = (rb_rand_double(rng) > dropout) ? a[i] : 0.0f;
  }
}

/* recur-nn-init.c:752-755 */
void
rnn_perforate_weights(RecurNN *net, float p)
{
  rb_host_will_touch_matrices(net);
  perforate(net->ih_weights, net->ih_size, p, &net->rng);
  perforate(net->ho_weights, net->ho_size, p, &net->rng);
  rb_weights_changed(net);
}

static void
noise_block(rand_ctx *rng, float *w, int width, int stride, int height, float deviation)
{
  for (int y = 0; y < height; y++) {
    float *row = w + (size_t)y * stride;
    for (int i = 0; i < width; i++)
      row[i] += rb_cheap_gaussian_noise(rng) * deviation;
  }
}

/* recur-nn.c:857-883 */
void
rnn_weight_noise(RecurNN *net, float deviation)
{
  rb_host_will_touch_matrices(net);
  noise_block(&net->rng, net->ih_weights + 1, net->hidden_size, net->h_size,
      net->hidden_size + 1 + net->input_size, deviation);
  noise_block(&net->rng, net->ho_weights, net->output_size, net->o_size,
      net->hidden_size + 1, deviation);
  if (net->bottom_layer) {
    RecurExtraLayer *bl = net->bottom_layer;
    noise_block(&net->rng, bl->weights + 1, bl->input_size, bl->i_size,
        bl->output_size, deviation);
  }
  rb_weights_changed(net);
}

/* recur-nn.c:1027-1076: nudge ih_weights until a rectified gaussian input
   comes out with about target_gain times its energy */
void
rnn_scale_initial_weights(RecurNN *net, float target_gain)
{
  rb_host_will_touch_matrices(net);
  const int h_size = net->h_size;
  float *in = (float *)malloc(h_size * sizeof(float));
  float *out = (float *)malloc(h_size * sizeof(float));
  double net_adjustment = 1.0, tail_in = 0, tail_out = 0;
  const double generations = 10000;
  for (double j = 1; j < generations; j++) {
    float sum_out, sum_in = 1;
    in[0] = 1;
    for (int i = 1; i < net->hidden_size; i++) {
      /* The reference writes MAX(cheap_gaussian_noise(rng), 0) with a macro
         that evaluates its argument twice (recur-common.h:183,
         recur-nn.c:1042): one draw decides the sign test and, when it is
         non-negative, a SECOND draw becomes the value (which may then be
         negative).  Reproduced so the generator stays in step. */
      float n = 0;
      if (rb_cheap_gaussian_noise(&net->rng) >= 0)
        n = rb_cheap_gaussian_noise(&net->rng);
      in[i] = n;
      sum_in += n * n;
    }
    for (int i = net->hidden_size; i < h_size; i++) {
      in[i] = 0;
      out[i] = 0;
    }
    memset(out, 0, h_size * sizeof(float));
    for (int y = 0; y < net->hidden_size + 1; y++) {
      float v = in[y];
      if (v) {
        const float *row = net->ih_weights + (size_t)y * h_size;
        for (int x = 0; x < h_size; x++)
          out[x] += v * row[x];
      }
    }
    out[0] = 1.0f;
    sum_out = 0;
    for (int i = 0; i < net->hidden_size; i++) {
      float h = out[i];
      h = (h > 0.0f) ? h : 0.0f;
      out[i] = h;
      sum_out += h * h;
    }
    double ratio = sum_out / sum_in;
    double adj = (target_gain * 10 + j) / (ratio * 10 + j);
    net_adjustment *= adj;
    float fadj = adj;
    for (int i = 0; i < net->ih_size; i++)
      net->ih_weights[i] *= fadj;
    if (j > generations * 0.95) {
      tail_in += sum_in;
      tail_out += sum_out;
    }
  }
  fprintf(stderr, "scaled toward target gain %.3f; hit roughly %.3f; adjusted by %.3f\n",
      target_gain, tail_out / tail_in, net_adjustment);
  free(in);
  free(out);
  rb_weights_changed(net);
}

/* recur-nn.c:1078-1134 */
void
rnn_zap_non_diagonals(RecurNN *net, int start, int stop, int friend_n)
{
  int h_end = net->hidden_size + 1;
  int friend_start = start - friend_n;
  if (start >= h_end || start < 0)
    return;
  if (start > stop) {
    fprintf(stderr, "diagonal zap start is %d, stop is %d; doing nothing\n", start, stop);
    return;
  }
  if (stop > h_end) {
    fprintf(stderr, "net->hidden size is %d, diagonal zap stop is %d; truncating\n",
        net->hidden_size, stop);
    stop = h_end;
  }
  if (friend_n > stop - start || friend_start <= 0) {
    fprintf(stderr, "diagonal friend parameter %d is stupid: start is %d stop %d, "
        "size %d ...ignoring it\n", friend_n, start, stop, h_end);
    friend_n = 0;
  }
  rb_host_will_touch_matrices(net);
  float *row_start = net->ih_weights + start;
  int zero_len = stop - start;
  int stride = net->h_size;
  for (int y = 0; y < h_end; y++) {
    if (y < friend_start || y >= stop) {
      memset(row_start, 0, zero_len * sizeof(float));
    }
    else {
      int x = (y < start) ? y - friend_start : y - start;
      memset(row_start, 0, x * sizeof(float));
      memset(row_start + x + 1, 0, (zero_len - x - 1) * sizeof(float));
    }
    row_start += stride;
  }
  rb_weights_changed(net);
}

/* recur-nn.c:1136-1145 */
void
rnn_clear_diagonal_only_section(RecurNN *net, uint len, uint friends)
{
  int h_end = net->hidden_size + 1;
  int start = h_end - len;
  int stop = h_end;
  if (friends > len)
    friends = len;
  rnn_zap_non_diagonals(net, start, stop, friends);
}

/* ---- inspection -------------------------------------------------------------- */

static void
mean_and_variance(const float *a, int width, int height, int stride, int offset,
    const char *name)
{
  float mean = 0, var = 0, n = 0;
  for (int y = 0; y < height; y++) {
    for (int x = offset; x < width + offset; x++) {
      n++;
      float val = a[(size_t)y * stride + x];
      float delta = val - mean;
      mean += delta / n;
      var += delta * (val - mean);
    }
  }
  var /= n;
  fprintf(stderr, "%s: mean %3g variance %3g (std dev %3g) n %d\n", name, mean, var,
      sqrt(var), (int)n);
}

/* recur-nn-init.c:845-861 */
void
rnn_print_net_stats(RecurNN *net)
{
  rb_host_will_touch_matrices(net);
  mean_and_variance(net->ih_weights, net->hidden_size,
      net->hidden_size + net->input_size + 1, net->h_size, 1, "ih_weights");
  mean_and_variance(net->ho_weights, net->output_size, net->hidden_size + 1,
      net->o_size, 0, "ho_weights");
  if (net->bottom_layer) {
    RecurExtraLayer *bl = net->bottom_layer;
    mean_and_variance(bl->weights, bl->output_size, bl->input_size, bl->o_size, 1,
        "bottom weights");
  }
}

/* One signed matrix as a P6 image: negative red, positive green, zero blue,
   scaled to the largest magnitude (the reference's weight pictures,
   pgm_dump.h:125-167,215-221; file name "images/<desc>-<gen>-<w>x<h>.ppm"). */
static void
dump_signed_ppm(const float *w, int width, int height, const char *desc, int id)
{
  char name[200];
  snprintf(name, sizeof(name), "images/%s-%08d-%dx%d.ppm", desc, id, width, height);
  float biggest = 1e-35f;
  size_t n = (size_t)width * height;
  for (size_t i = 0; i < n; i++) {
    float f = fabsf(w[i]);
    if (f > biggest)
      biggest = f;
  }
  float scale = 255.99f / biggest;
  FILE *fh = fopen(name, "w");
  if (!fh) {
    fprintf(stderr, "could not open '%s' for writing\n", name);
    return;
  }
  fprintf(fh, "P6\n%u %u\n255\n", width, height);
  for (size_t i = 0; i < n; i++) {
    float f = w[i] * scale;
    u8 b = fabsf(f);
    u8 rgb[3] = {0, 0, 0};
    if (f < 0.0)
      rgb[0] = b;
    else if (f > 0.0)
      rgb[1] = b;
    else
      rgb[2] = 180;
    fwrite(rgb, 1, 3, fh);
  }
  fclose(fh);
}

/* recur-nn-init.c:757-823: dumpees is a space separated list of three-letter
   codes: <from><to><what>, e.g. "ihw hod" */
void
rnn_multi_pgm_dump(RecurNN *net, const char *dumpees, const char *basename)
{
  RecurNNBPTT *bptt = net->bptt;
  rb_host_will_touch_matrices(net);
  char *copy = strdup(dumpees);
  char *working = copy;
  char *token;
  while ((token = strsep(&working, " "))) {
    int x = 0, y = 0;
    float *array = NULL;
    if (strlen(token) != 3)
      continue;
    char in = token[0], out = token[1], v = token[2];
    int aux_ok = (net->flags & RNN_NET_FLAG_AUX_ARRAYS) != 0;
    if (out == 'h') {
      x = net->h_size;
      if (in == 'i')
        y = net->i_size;
      else if (in == 'h')
        y = net->hidden_size;
      else
        continue;
      if (v == 'w') array = net->ih_weights;
      else if (v == 'm' && bptt) array = bptt->ih_momentum;
      else if (v == 'd' && bptt) array = bptt->ih_delta;
      else if (v == 't' && bptt) array = bptt->ih_delta_tmp;
      else if (v == 'a' && bptt && aux_ok) array = bptt->ih_aux;
      else continue;
    }
    else if (in == 'h' && out == 'o') {
      x = net->o_size;
      y = net->h_size;
      if (v == 'w') array = net->ho_weights;
      else if (v == 'm' && bptt) array = bptt->ho_momentum;
      else if (v == 'd' && bptt) array = bptt->ho_delta;
      else if (v == 'a' && bptt && aux_ok) array = bptt->ho_aux;
      else continue;
    }
    else if (in == 'b' && out == 'i') {
      RecurExtraLayer *b = net->bottom_layer;
      if (!b)
        continue;
      x = b->o_size;
      y = b->i_size;
      if (v == 'w') array = b->weights;
      else if (v == 'm') array = b->momentums;
      else if (v == 'd') array = b->delta;
      else if (v == 'a' && aux_ok) array = b->aux;
      else continue;
    }
    if (array) {
      if (!basename || !basename[0])
        basename = "untitled";
      char name[160];
      snprintf(name, sizeof(name), "%s-%s", basename, token);
      dump_signed_ppm(array, x, y, name, net->generation);
    }
  }
  free(copy);
}
