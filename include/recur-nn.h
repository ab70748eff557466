/* recur-nn.h — the RecurNN C API, served by librecur_b200.so.
 *
 * This header is the drop-in boundary.  It declares the same
 * C symbols, the same flag/enum values and byte-identical struct layouts as
 * the reference's recur-nn.h (reference recur-nn.h:15-334), so that
 * text-predict, charmodel, gstclassify and gstrnnca compile and link against
 * librecur_b200.so unchanged.  Behind it, rnn_opinion / rnn_bptt_* /
 * rnn_apply_learning run as sm_100a CUDA kernels; there is no CPU compute
 * path: a compute call without a usable CUDA device aborts.
 *
 * What the pointers in the structs mean here:
 *   - ih_weights, ho_weights, *_momentum, *_delta, ih_delta_tmp, *_aux and
 *     the bottom layer's matrices are CUDA managed allocations: valid in
 *     host code exactly as in the reference (callers poke them, e.g.
 *     text-predict.c:462-468) and the same address is what the kernels use.
 *   - input_layer, hidden_layer, output_layer, real_inputs, i/h/o_error and
 *     history are pinned host mirrors of one stream's slot in the device
 *     stream pool.  The per-net calls (rnn_opinion, rnn_bptt_calc_deltas ...)
 *     copy them in on entry and out on exit, so the reference protocol
 *     "write real_inputs -> rnn_opinion -> read output_layer -> write
 *     o_error -> rnn_bptt_calc_deltas" holds unchanged.
 *   - the array-of-nets calls in recur_b200.h keep the state on the device
 *     between calls and refresh the mirrors only on request.
 *
 * Each declaration cites the reference line it replaces.
 */
#ifndef RECUR_B200_RECUR_NN_H
#define RECUR_B200_RECUR_NN_H 1

#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#ifdef __cplusplus
extern "C" {
#endif

/* short integer names, as reference recur-common.h:80-88 */
#ifndef RECUR_B200_HAVE_SHORT_TYPES
#define RECUR_B200_HAVE_SHORT_TYPES 1
typedef uint64_t u64;
typedef int64_t s64;
typedef uint32_t u32;
typedef int32_t s32;
typedef uint16_t u16;
typedef int16_t s16;
typedef uint8_t u8;
typedef int8_t s8;
typedef unsigned int uint;
#endif

/* Jenkins small PRNG state; reference recur-rng.h:17-22 */
typedef struct _rand_ctx {
  u64 a;
  u64 b;
  u64 c;
  u64 d;
} rand_ctx;

#define RECUR_RNG_RANDOM_SEED (-1ULL) /* recur-rng.h:15 */
#define RECUR_RNG_SUBSEED (-2ULL)     /* recur-nn.h:15 */

/* ---- tuning constants (reference recur-nn.h:17-57) ---------------------- */
#define RANDOM_DAMAGE_FACTOR 0.5f
#define MAX_TOP_ERROR_FACTOR 2.0f
#define MAX_ERROR_GAIN 2.0f
#define ERROR_GAIN_CEILING 1.0f
#define BASE_MIN_ERROR_FACTOR 1e-12f
#define MAX_MIN_ERROR_FACTOR 1e-2f
#define ABS_MIN_ERROR_FACTOR 1e-20f
#define MIN_ERROR_GAIN 1e-8f
#define RNN_HIDDEN_PENALTY 0.0f
#define HIDDEN_MEAN_SOFT_TOP 16.0f
#define INPUT_MEAN_SOFT_TOP 16.0f
#define RNN_INITIAL_WEIGHT_VARIANCE_FACTOR 2.0f
#define WEIGHT_SCALE (1.0f - 1e-6f)
#define RNN_CONDITIONING_INTERVAL 8
#define RNN_TALL_POPPY_THRESHOLD 1.0f
#define RNN_TALL_POPPY_SCALE 0.99f
#define RNN_LAWN_MOWER_THRESHOLD 10.0f
#define RNN_MOMENTUM_WEIGHT 0.5f

/* ---- conditioning schedule (reference recur-nn.h:59-76) ------------------ */
#define RNN_COND_USE_OFFSET 16
enum {
  RNN_COND_BIT_SCALE = 0U,
  RNN_COND_BIT_ZERO = 2U,
  RNN_COND_BIT_LAWN_MOWER = 3U,
  RNN_COND_BIT_TALL_POPPY = 4U,
  RNN_COND_BIT_RAND = 6U
};

/* ---- net flags (reference recur-nn.h:78-103) ----------------------------- */
enum {
  RNN_NET_FLAG_OWN_BPTT = 1,
  RNN_NET_FLAG_OWN_WEIGHTS = 2,
  RNN_NET_FLAG_LOG_APPEND = 8,
  RNN_NET_FLAG_LOG_HIDDEN_SUM = 16,
  RNN_NET_FLAG_LOG_WEIGHT_SUM = 32,
  RNN_NET_FLAG_BPTT_ADAPTIVE_MIN_ERROR = 64,
  RNN_NET_FLAG_NO_MOMENTUMS = 128,
  RNN_NET_FLAG_NO_DELTAS = 256,
  RNN_NET_FLAG_BOTTOM_LAYER = 1024,
  RNN_NET_FLAG_AUX_ARRAYS = 2048,

  RNN_COND_USE_SCALE = (1 << (RNN_COND_BIT_SCALE + RNN_COND_USE_OFFSET)),
  RNN_COND_USE_ZERO = (1 << (RNN_COND_BIT_ZERO + RNN_COND_USE_OFFSET)),
  RNN_COND_USE_LAWN_MOWER = (1 << (RNN_COND_BIT_LAWN_MOWER + RNN_COND_USE_OFFSET)),
  RNN_COND_USE_TALL_POPPY = (1 << (RNN_COND_BIT_TALL_POPPY + RNN_COND_USE_OFFSET)),
  RNN_COND_USE_RAND = (1 << (RNN_COND_BIT_RAND + RNN_COND_USE_OFFSET)),

  RNN_NET_FLAG_STANDARD = (RNN_NET_FLAG_OWN_BPTT | RNN_NET_FLAG_OWN_WEIGHTS |
      RNN_COND_USE_ZERO | RNN_NET_FLAG_LOG_HIDDEN_SUM)
};

/* reference recur-nn.h:109-119 */
typedef enum {
  RNN_MOMENTUM_WEIGHTED = 0,
  RNN_MOMENTUM_NESTEROV,
  RNN_MOMENTUM_SIMPLIFIED_NESTEROV,
  RNN_MOMENTUM_CLASSICAL,
  RNN_ADAGRAD,
  RNN_ADADELTA,
  RNN_RPROP,
  RNN_LAST_LEARNING_METHOD
} rnn_learning_method;

/* reference recur-nn.h:121-128 */
typedef enum {
  RNN_INIT_ZERO = 0,
  RNN_INIT_FLAT,
  RNN_INIT_FAN_IN,
  RNN_INIT_RUNS,
  RNN_INIT_LAST
} rnn_init_method;

/* reference recur-nn.h:130-140 */
typedef enum {
  RNN_RELU = 1,
  RNN_RESQRT,
  RNN_RESERVED_ACTIVATION_1,
  RNN_RESERVED_ACTIVATION_2,
  RNN_RECLIP20 = 5,
  RNN_ACTIVATION_LAST
} rnn_activation;

/* reference recur-nn.h:142-151 */
typedef enum {
  RNN_INIT_DIST_UNIFORM = 1,
  RNN_INIT_DIST_GAUSSIAN,
  RNN_INIT_DIST_LOG_NORMAL,
  RNN_INIT_DIST_SEMICIRCLE,
  RNN_INIT_DIST_DEFAULT
} rnn_init_distribution;

typedef struct _RecurNN RecurNN;
typedef struct _RecurNNBPTT RecurNNBPTT;
typedef struct _RecurExtraLayer RecurExtraLayer;

/* One stream's view of a net; reference recur-nn.h:158-186.  Sizes with a
   leading letter are padded to a multiple of 4 floats (i_size covers bias +
   hidden feedback + inputs). */
struct _RecurNN {
  int i_size;
  int h_size;
  int o_size;
  int input_size;
  int hidden_size;
  int output_size;
  int ih_size;
  int ho_size;
  u32 flags;
  FILE *log;
  float *mem;
  float *input_layer;
  float *hidden_layer;
  float *output_layer;
  float *ih_weights;
  float *ho_weights;
  float *real_inputs;
  rand_ctx rng;
  RecurNNBPTT *bptt;
  RecurExtraLayer *bottom_layer;
  char *metadata;
  u32 generation;
  float presynaptic_noise;
  rnn_activation activation;
};

/* Training state; reference recur-nn.h:188-209 */
struct _RecurNNBPTT {
  int depth;
  int index;
  float *i_error;
  float *h_error;
  float *o_error;
  float *ih_momentum;
  float *ho_momentum;
  float *history;
  float *ih_delta;
  float *ho_delta;
  float *ih_delta_tmp;
  float *ih_aux;
  float *ho_aux;
  float *mem;
  float learn_rate;
  float ih_scale;
  float ho_scale;
  float momentum;
  float momentum_weight;
  float min_error_factor;
};

/* Optional layer below the recurrent one; reference recur-nn.h:211-227 */
struct _RecurExtraLayer {
  float *mem;
  float *weights;
  float *momentums;
  float *aux;
  float *delta;
  float *inputs;
  float *outputs;
  float *i_error;
  float *o_error;
  float learn_rate_scale;
  int input_size;
  int output_size;
  int i_size;
  int o_size;
  int overlap;
};

/* reference recur-nn.h:230-258 */
struct RecurInitialisationParameters {
  rnn_init_method method;
  rnn_init_method submethod;
  int bias_uses_submethod;
  int inputs_use_submethod;

  float fan_in_sum;
  float fan_in_step;
  float fan_in_min;
  float fan_in_ratio;

  float flat_variance;
  rnn_init_distribution flat_shape;
  double flat_perforation;

  float run_input_probability;
  float run_input_magnitude;
  float run_gain;
  float run_len_mean;
  float run_len_stddev;
  int run_n;
  int run_loop;
  int run_crossing_paths;
  int run_inputs_miss;
  int run_input_at_start;
};

/* A span of output columns carrying error (multi-head nets);
   reference recur-nn.h:260-265.  Lists end with start < 0. */
typedef struct _RecurErrorRange RecurErrorRange;
struct _RecurErrorRange {
  int start;
  int len;
};

/* ---- construction / destruction (reference recur-nn.h:269-300) ----------- */
RecurNN *rnn_new(uint input_size, uint hidden_size, uint output_size,
    u32 flags, u64 rng_seed, const char *log_file, int depth, float learn_rate,
    float momentum, float presynaptic_noise, rnn_activation activation);

RecurNN *rnn_clone(RecurNN *parent, u32 flags, u64 rng_seed, const char *log_file);

RecurExtraLayer *rnn_new_extra_layer(int input_size, int output_size, int overlap,
    u32 flags);

RecurNN *rnn_new_with_bottom_layer(int n_inputs, int r_input_size,
    int hidden_size, int output_size, u32 flags, u64 rng_seed,
    const char *log_file, int bptt_depth, float learn_rate,
    float momentum, float presynaptic_noise,
    rnn_activation activation, int convolutional_overlap);

void rnn_set_log_file(RecurNN *net, const char *log_file, int append_dont_truncate);

void rnn_randomise_weights_clever(RecurNN *net, struct RecurInitialisationParameters *p);
void rnn_randomise_weights_simple(RecurNN *net, const rnn_init_method method);
void rnn_randomise_weights_auto(RecurNN *net);
void rnn_init_default_weight_parameters(RecurNN *net,
    struct RecurInitialisationParameters *q);
void rnn_scale_initial_weights(RecurNN *net, float target_gain);
void rnn_print_net_stats(RecurNN *net);

void rnn_delete_net(RecurNN *net);
RecurNN **rnn_new_training_set(RecurNN *prototype, int n_nets);
void rnn_delete_training_set(RecurNN **nets, int n_nets, int leave_prototype);

/* ---- the hot path (reference recur-nn.h:302-322) ------------------------- */
float *rnn_opinion(RecurNN *net, const float *inputs, float presynaptic_noise);

void rnn_multi_pgm_dump(RecurNN *net, const char *dumpees, const char *basename);

RecurNN *rnn_load_net(const char *filename);
int rnn_save_net(RecurNN *net, const char *filename, int backup);

void rnn_bptt_clear_deltas(RecurNN *net);
void rnn_bptt_advance(RecurNN *net);
void rnn_bptt_calculate(RecurNN *net, uint batch_size);
void rnn_apply_learning(RecurNN *net, int learning_style, float momentum);
float rnn_calculate_momentum_soft_start(float generation, float momentum,
    float momentum_soft_start);

void rnn_bptt_calc_deltas(RecurNN *net, int accumulate_delta,
    RecurErrorRange *top_error_ranges);

void rnn_condition_net(RecurNN *net);
void rnn_log_net(RecurNN *net);
void rnn_forget_history(RecurNN *net, int bptt_too);

/* ---- weight surgery (reference recur-nn.h:324-334) ----------------------- */
void rnn_perforate_weights(RecurNN *net, float p);
void rnn_weight_noise(RecurNN *net, float deviation);
void rnn_set_momentum_values(RecurNN *net, float x);
void rnn_set_aux_values(RecurNN *net, float x);
void rnn_zap_non_diagonals(RecurNN *net, int start, int stop, int n_friends);
void rnn_clear_diagonal_only_section(RecurNN *net, uint len, uint friends);

/* log-line helpers callers use directly (reference recur-nn.h:337-349) */
static inline void
rnn_log_float(RecurNN *net, const char *name, float value)
{
  if (net->log)
    fprintf(net->log, "%s %.5g\n", name, value);
}

static inline void
rnn_log_int(RecurNN *net, const char *name, int value)
{
  if (net->log)
    fprintf(net->log, "%s %d\n", name, value);
}

#ifdef __cplusplus
}
#endif
#endif
